"""ctypes binding of libapproxfun_b200.so (the C ABI declared in include/approxfun_b200.h).

The product library is the ONLY thing `load()` will open: there is no CPU fallback and no alternate
backend.  A missing library or a missing CUDA device raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libapproxfun_b200.so")

AFB_OK, AFB_BAD_ARG, AFB_BAD_SIZE, AFB_NO_DEVICE, AFB_CUDA_ERROR, AFB_NCCL_ERROR, AFB_OOM, AFB_UNSUPPORTED = range(8)
CHEB1_FWD, CHEB1_INV, FOURIER_FWD, FOURIER_INV, LAURENT_FWD, LAURENT_INV, CHEB1_2D_FWD, CHEB1_2D_INV = range(8)
CHEB2_FWD, CHEB2_INV, SIN_FWD, SIN_INV = 8, 9, 10, 11
TAYLOR_FWD, TAYLOR_INV, HARDYM_FWD, HARDYM_INV = 12, 13, 14, 15
PADUA_FWD, PADUA_INV = 16, 17
FOURIER_CHEB1_2D_FWD, FOURIER_CHEB1_2D_INV, LAURENT_CHEB1_2D_FWD, LAURENT_CHEB1_2D_INV = 18, 19, 20, 21
PLAN_2D_TRANSPOSE = 1  # afb_plan_create flag: explicit-transpose row pass for 2D kinds

# every symbol include/approxfun_b200.h declares (tests check the .so exports each of them)
SYMBOLS = [
    "afb_version", "afb_last_error", "afb_init", "afb_shutdown", "afb_device_info",
    "afb_malloc", "afb_free", "afb_host_alloc", "afb_host_free", "afb_host_register", "afb_host_unregister", "afb_memcpy_h2d", "afb_memcpy_d2h",
    "afb_memcpy_d2d", "afb_stream_sync",
    "afb_plan_create", "afb_plan_execute_host", "afb_plan_execute_device", "afb_plan_destroy", "afb_plan_launches",
    "afb_plan_io_doubles", "afb_plan_execute_buf",
    "afb_buf_create", "afb_buf_destroy", "afb_buf_size", "afb_buf_ptr", "afb_buf_upload", "afb_buf_download", "afb_buf_unary",
    "afb_buf_binary", "afb_buf_pad", "afb_buf_chebpoints", "afb_buf_clenshaw", "afb_buf_tail",
    "afb_chebpoints", "afb_chebpoints_device", "afb_fourierpoints", "afb_paduapoints",
    "afb_clenshaw_cheb", "afb_clenshaw_cheb_device", "afb_clenshaw_cheb_cols", "afb_clenshaw_cheb_cols_device",
    "afb_clenshaw_cheb_multi", "afb_clenshaw_cheb_multi_device", "afb_clenshaw_rec", "afb_clenshaw_rec_device",
    "afb_cheb_normcumsum", "afb_cheb_normcumsum_device", "afb_cheb_bisect", "afb_cheb_bisect_device",
    "afb_cheb_bisect_cols", "afb_cheb_bisect_cols_device", "afb_samplecdf_cheb", "afb_sample_cheb", "afb_sample_cheb_device",
    "afb_philox_uniform_device", "afb_fp64_probe_device",
    "afb_clenshaw_cheb2d", "afb_clenshaw_cheb2d_device", "afb_piecewise_clenshaw", "afb_piecewise_bisect",
    "afb_clenshaw_fourier_cheb2d", "afb_clenshaw_fourier_cheb2d_device",
    "afb_sample2d_cheb_lowrank", "afb_sample2d_cheb_lowrank_device", "afb_dgemm_nn", "afb_dgemm_nn_device", "afb_dgemm_nn_dmma_device",
    "afb_comm_unique_id", "afb_comm_create", "afb_comm_destroy",
    "afb_sharded2d_create", "afb_sharded2d_execute_device", "afb_sharded2d_destroy",
    "afb_sharded2d_ipc_export", "afb_sharded2d_ipc_open", "afb_sharded2d_set_chunks", "afb_sharded2d_set_phases", "afb_sharded2d_set_tuning",
]


class AfbError(RuntimeError):
    """A non-zero status from the C ABI (the Julia shim turns the same codes into `error(msg)`)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libapproxfun_b200 status {code}: {msg}")
        self.code = code


class DimensionMismatch(AfbError, ValueError):
    """AFB_BAD_SIZE -- mirrors Julia's DimensionMismatch / the `@assert` of src/Extras/sample.jl:85."""


def bind(lib: C.CDLL) -> C.CDLL:
    i32, i64, u32, u64, vp, dbl, sz = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_double, C.c_size_t
    sig = {
        "afb_version": [],
        "afb_last_error": [C.c_char_p, sz],
        "afb_init": [C.POINTER(i32), i32],
        "afb_shutdown": [],
        "afb_device_info": [C.POINTER(i32), C.POINTER(i64), C.c_char_p, sz],
        "afb_malloc": [C.POINTER(vp), sz],
        "afb_free": [vp],
        "afb_host_alloc": [C.POINTER(vp), sz],
        "afb_host_free": [vp],
        "afb_host_register": [vp, sz],
        "afb_host_unregister": [vp],
        "afb_memcpy_h2d": [vp, vp, sz, vp],
        "afb_memcpy_d2h": [vp, vp, sz, vp],
        "afb_memcpy_d2d": [vp, vp, sz, vp],
        "afb_stream_sync": [vp],
        "afb_plan_create": [C.POINTER(vp), i32, i64, i64, i64, i64, u32],
        "afb_plan_execute_host": [vp, vp, vp],
        "afb_plan_execute_device": [vp, vp, vp, vp],
        "afb_plan_destroy": [vp],
        "afb_plan_launches": [vp, C.POINTER(i32)],
        "afb_plan_io_doubles": [vp, C.POINTER(i64), C.POINTER(i64)],
        "afb_plan_execute_buf": [vp, vp, vp],
        "afb_buf_create": [C.POINTER(vp), i64],
        "afb_buf_destroy": [vp],
        "afb_buf_size": [vp, C.POINTER(i64)],
        "afb_buf_ptr": [vp, C.POINTER(vp)],
        "afb_buf_upload": [vp, vp, i64, i64],
        "afb_buf_download": [vp, vp, i64, i64],
        "afb_buf_unary": [i32, vp, vp, i64, dbl, dbl],
        "afb_buf_binary": [i32, vp, vp, vp, i64],
        "afb_buf_pad": [vp, i64, vp, i64],
        "afb_buf_chebpoints": [vp, i64, dbl, dbl],
        "afb_buf_clenshaw": [vp, i64, vp, vp, i64],
        "afb_buf_tail": [vp, i64, i64, dbl, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)],
        "afb_chebpoints": [vp, i64, dbl, dbl],
        "afb_chebpoints_device": [vp, i64, dbl, dbl, vp],
        "afb_fourierpoints": [vp, i64, dbl, dbl],
        "afb_paduapoints": [vp, i64, dbl, dbl, dbl, dbl, C.POINTER(i64)],
        "afb_clenshaw_cheb": [vp, i64, vp, vp, i64],
        "afb_clenshaw_cheb_device": [vp, i64, vp, vp, i64, vp],
        "afb_clenshaw_cheb_cols": [vp, i64, i64, vp, vp, i64],
        "afb_clenshaw_cheb_cols_device": [vp, i64, i64, vp, vp, i64, vp],
        "afb_clenshaw_cheb_multi": [vp, i64, i64, vp, vp, i64],
        "afb_clenshaw_rec": [vp, i64, vp, vp, vp, vp, vp, i64],
        "afb_clenshaw_rec_device": [vp, i64, vp, vp, vp, vp, vp, i64, vp],
        "afb_clenshaw_cheb_multi_device": [vp, i64, i64, vp, vp, i64, vp],
        "afb_cheb_normcumsum": [vp, i64, i64, i64, i32],
        "afb_cheb_normcumsum_device": [vp, i64, i64, i64, i32, vp],
        "afb_cheb_bisect": [vp, i64, vp, vp, i64, i32],
        "afb_cheb_bisect_device": [vp, i64, vp, vp, i64, i32, dbl, dbl, vp],
        "afb_cheb_bisect_cols": [vp, i64, i64, vp, vp, i64, i32],
        "afb_cheb_bisect_cols_device": [vp, i64, i64, vp, vp, i64, i32, vp],
        "afb_samplecdf_cheb": [vp, i64, dbl, dbl, vp, vp, i64],
        "afb_sample_cheb": [vp, i64, dbl, dbl, u64, vp, i64],
        "afb_sample_cheb_device": [vp, i64, dbl, dbl, u64, u64, vp, i64, vp],
        "afb_philox_uniform_device": [u64, u64, vp, i64, vp],
        "afb_fp64_probe_device": [i32, i32, i32, vp, vp],
        "afb_piecewise_clenshaw": [vp, vp, vp, i64, vp, vp, i64],
        "afb_piecewise_bisect": [vp, vp, vp, i64, vp, vp, i64, i32],
        "afb_clenshaw_cheb2d": [vp, i64, i64, vp, vp, vp, i64],
        "afb_clenshaw_fourier_cheb2d": [vp, i64, i64, vp, vp, vp, i64],
        "afb_clenshaw_fourier_cheb2d_device": [vp, i64, i64, vp, vp, vp, i64, vp],
        "afb_clenshaw_cheb2d_device": [vp, i64, i64, vp, vp, vp, i64, vp],
        "afb_sample2d_cheb_lowrank": [vp, i64, vp, i64, i64, dbl, dbl, dbl, dbl, vp, vp, vp, i64],
        "afb_sample2d_cheb_lowrank_device": [vp, i64, vp, i64, i64, dbl, dbl, dbl, dbl, vp, vp, vp, i64, vp],
        "afb_dgemm_nn": [vp, vp, vp, i64, i64, i64],
        "afb_dgemm_nn_device": [vp, vp, vp, i64, i64, i64, vp],
        "afb_dgemm_nn_dmma_device": [vp, vp, vp, i64, i64, i64, vp],
        "afb_comm_unique_id": [vp],
        "afb_comm_create": [vp, i32, i32],
        "afb_comm_destroy": [],
        "afb_sharded2d_create": [C.POINTER(vp), i32, i64, i64],
        "afb_sharded2d_execute_device": [vp, vp, vp, vp],
        "afb_sharded2d_destroy": [vp],
        "afb_sharded2d_ipc_export": [vp, vp],
        "afb_sharded2d_ipc_open": [vp, vp, i32],
        "afb_sharded2d_set_chunks": [vp, i32],
        "afb_sharded2d_set_phases": [vp, i32],
        "afb_sharded2d_set_tuning": [vp, i32, i32, i32],
    }
    assert sorted(sig) == sorted(SYMBOLS)
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    return lib


_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Open the product library (built by `python __graft_entry__.py` / csrc/Makefile)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `make -C approxfun.jl_b200/csrc` (nvcc, sm_100a). "
                    "There is no CPU fallback.")
            _lib = bind(C.CDLL(LIB_PATH))
        return _lib


def init(lib: C.CDLL, devices=None) -> None:
    """`afb_init(devices, ndev)`: None keeps the current device; one ordinal selects it (one process per GPU); several
    ordinals make the host-pointer entry points fan out over those GPUs from this one process."""
    if devices is None:
        check(lib, lib.afb_init(None, 0))
        return
    devs = [int(devices)] if isinstance(devices, (int, np.integer)) else [int(d) for d in devices]
    arr = (C.c_int32 * len(devs))(*devs)
    check(lib, lib.afb_init(arr, len(devs)))


def check(lib: C.CDLL, status: int) -> None:
    if status == AFB_OK:
        return
    buf = C.create_string_buffer(1024)
    lib.afb_last_error(buf, 1024)
    msg = buf.value.decode(errors="replace")
    if status == AFB_BAD_SIZE:
        raise DimensionMismatch(status, msg)
    raise AfbError(status, msg)


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class LowLevel:
    """NumPy-level calls into the C ABI with HOST buffers (each call: H2D, kernels, D2H)."""

    def __init__(self, lib: C.CDLL):
        self.lib = lib

    # -- plans ------------------------------------------------------------------------------------
    def plan_create(self, kind: int, n: int, n2: int = 0, batch: int = 1, stride: int | None = None, flags: int = 0):
        h = C.c_void_p()
        check(self.lib, self.lib.afb_plan_create(C.byref(h), kind, n, n2, batch, n if stride is None else stride, flags))
        return h

    def plan_destroy(self, h):
        check(self.lib, self.lib.afb_plan_destroy(h))

    def plan_execute_host(self, h, src: np.ndarray, dst: np.ndarray):
        check(self.lib, self.lib.afb_plan_execute_host(h, _ptr(src), _ptr(dst)))

    def plan_launches(self, h) -> int:
        v = C.c_int32()
        check(self.lib, self.lib.afb_plan_launches(h, C.byref(v)))
        return int(v.value)

    def transform(self, kind: int, v: np.ndarray) -> np.ndarray:
        """v: (n,) or (n, B) array in the reference's column-major sense (each column one transform)."""
        cplx = kind in (LAURENT_FWD, LAURENT_INV, TAYLOR_FWD, TAYLOR_INV, HARDYM_FWD, HARDYM_INV)
        dt = np.complex128 if cplx else np.float64
        v = np.asarray(v, dtype=dt)
        one = v.ndim == 1
        if v.size == 0:
            return v.copy()
        m = v.reshape(v.shape[0], -1)
        n, B = m.shape
        buf = np.ascontiguousarray(m.T)  # B lines of length n
        out = np.empty_like(buf)
        h = self.plan_create(kind, n, 0, B, n)
        try:
            if n and B:
                self.plan_execute_host(h, buf, out)
        finally:
            self.plan_destroy(h)
        res = np.ascontiguousarray(out.T)
        return res[:, 0] if one else res.reshape(v.shape)

    def transform2d(self, kind: int, V: np.ndarray, flags: int = 0) -> np.ndarray:
        """V[k, j] (n x n2) in the reference's column-major sense."""
        V = np.asarray(V, dtype=np.float64)
        n, n2 = V.shape
        buf = np.ascontiguousarray(V.T)  # memory order of a Julia n x n2 matrix
        out = np.empty_like(buf)
        h = self.plan_create(kind, n, n2, 1, n, flags)
        try:
            if n and n2:
                self.plan_execute_host(h, buf, out)
        finally:
            self.plan_destroy(h)
        return np.ascontiguousarray(out.T)

    # -- points -------------------------------------------------------------------------------------
    def chebpoints(self, n: int, a: float = -1.0, b: float = 1.0) -> np.ndarray:
        out = np.empty(max(n, 0))
        check(self.lib, self.lib.afb_chebpoints(_ptr(out), n, a, b))
        return out

    def clenshaw_rec(self, c, A, B, Cc, x) -> np.ndarray:
        c, A, B, Cc, x = (np.ascontiguousarray(a, dtype=np.float64) for a in (c, A, B, Cc, x))
        if A.size < c.size or B.size < c.size or Cc.size < c.size + 1:
            raise DimensionMismatch(AFB_BAD_SIZE, "clenshaw_rec: need len(A), len(B) >= N and len(C) >= N + 1")
        y = np.empty_like(x)
        check(self.lib, self.lib.afb_clenshaw_rec(_ptr(c), c.size, _ptr(A), _ptr(B), _ptr(Cc), _ptr(x), _ptr(y), x.size))
        return y

    def sample_cheb(self, cdf_c, a: float, b: float, seed: int, n: int) -> np.ndarray:
        """n samples from the normalised cumulative `cdf_c` on a..b, uniforms drawn on the device (Philox, seed)."""
        cdf_c = np.ascontiguousarray(cdf_c, dtype=np.float64)
        out = np.empty(n)
        check(self.lib, self.lib.afb_sample_cheb(_ptr(cdf_c), cdf_c.size, a, b, seed, _ptr(out), n))
        return out

    def paduapoints(self, N: int, ax=-1.0, bx=1.0, ay=-1.0, by=1.0) -> np.ndarray:
        """Np x 2 array [x y] of the Padua points carrying N coefficients (Np = next triangular number >= N)."""
        npts = C.c_int64()
        check(self.lib, self.lib.afb_paduapoints(None, N, ax, bx, ay, by, C.byref(npts)))
        out = np.empty((2, npts.value))  # column-major Np x 2
        if npts.value:
            check(self.lib, self.lib.afb_paduapoints(_ptr(out), N, ax, bx, ay, by, C.byref(npts)))
        return np.ascontiguousarray(out.T)

    def padua(self, kind: int, v: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(v)
        h = self.plan_create(kind, v.size, 0, 1, v.size)
        try:
            if v.size:
                self.plan_execute_host(h, v, out)
        finally:
            self.plan_destroy(h)
        return out

    def fourierpoints(self, n: int, a: float = 0.0, b: float = 2 * np.pi) -> np.ndarray:
        out = np.empty(max(n, 0))
        check(self.lib, self.lib.afb_fourierpoints(_ptr(out), n, a, b))
        return out

    # -- Clenshaw ---------------------------------------------------------------------------------------
    def clenshaw(self, c, x) -> np.ndarray:
        c = np.ascontiguousarray(c, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        check(self.lib, self.lib.afb_clenshaw_cheb(_ptr(c), c.size, _ptr(x), _ptr(y), x.size))
        return y

    def clenshaw_cols(self, Cm, x) -> np.ndarray:
        Cm = np.asarray(Cm, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        N, P = Cm.shape
        flat = np.ascontiguousarray(Cm.T)
        y = np.empty(x.size)
        check(self.lib, self.lib.afb_clenshaw_cheb_cols(_ptr(flat), N, N, _ptr(x), _ptr(y), x.size if P == x.size else -1))
        return y

    def clenshaw_multi(self, Cm, x) -> np.ndarray:
        Cm = np.asarray(Cm, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        N, M = Cm.shape
        flat = np.ascontiguousarray(Cm.T)
        Y = np.empty((x.size, M))  # column-major M x P
        check(self.lib, self.lib.afb_clenshaw_cheb_multi(_ptr(flat), N, M, _ptr(x), _ptr(Y), x.size))
        return np.ascontiguousarray(Y.T)

    # -- sampling ------------------------------------------------------------------------------------------
    def cheb_normcumsum(self, Cm, grow: bool | None = None) -> np.ndarray:
        Cm = np.asarray(Cm, dtype=np.float64)
        if Cm.ndim == 1:
            g = True if grow is None else grow
            N = Cm.size
            buf = np.zeros(N + 1)
            buf[:N] = Cm
            check(self.lib, self.lib.afb_cheb_normcumsum(_ptr(buf), N, 1, N + 1, int(g)))
            return buf if g else buf[:N].copy()
        g = False if grow is None else grow
        N, P = Cm.shape
        ld = N + 1
        buf = np.zeros((P, ld))
        buf[:, :N] = Cm.T
        check(self.lib, self.lib.afb_cheb_normcumsum(_ptr(buf), N, P, ld, int(g)))
        return np.ascontiguousarray(buf[:, : (N + 1 if g else N)].T)

    def cheb_bisect(self, c, u, numits: int = 47) -> np.ndarray:
        c = np.ascontiguousarray(c, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty_like(u)
        check(self.lib, self.lib.afb_cheb_bisect(_ptr(c), c.size, _ptr(u), _ptr(out), u.size, numits))
        return out

    def cheb_bisect_cols(self, Cm, u, numits: int = 47) -> np.ndarray:
        Cm = np.asarray(Cm, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        N, P = Cm.shape
        flat = np.ascontiguousarray(Cm.T)
        out = np.empty(u.size)
        check(self.lib, self.lib.afb_cheb_bisect_cols(_ptr(flat), N, N, _ptr(u), _ptr(out), u.size if P == u.size else -1, numits))
        return out

    def clenshaw2d(self, Cm, x, y) -> np.ndarray:
        """Cm[i, j] multiplies T_i(x) T_j(y); x, y canonical, same length."""
        Cm = np.asarray(Cm, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if x.size != y.size:
            raise DimensionMismatch(AFB_BAD_SIZE, "clenshaw2d: x and y differ in length")
        flat = np.ascontiguousarray(Cm.T)  # column-major n1 x n2
        out = np.empty_like(x)
        check(self.lib, self.lib.afb_clenshaw_cheb2d(_ptr(flat), Cm.shape[0], Cm.shape[1], _ptr(x), _ptr(y), _ptr(out), x.size))
        return out

    def piecewise(self, pieces, breaks, xu, bisect: bool, numits: int = 47) -> np.ndarray:
        """pieces: list of coefficient vectors; breaks: len(pieces)+1 increasing break points; evaluation (bisect=False)
        or the generic bisection of sample.jl:23-36 (bisect=True)."""
        offs = np.zeros(len(pieces) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(p) for p in pieces])
        coef = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in pieces])) if offs[-1] else np.zeros(1)
        brk = np.ascontiguousarray(breaks, dtype=np.float64)
        xu = np.ascontiguousarray(xu, dtype=np.float64)
        out = np.empty_like(xu)
        if bisect:
            check(self.lib, self.lib.afb_piecewise_bisect(_ptr(coef), _ptr(offs), _ptr(brk), len(pieces), _ptr(xu), _ptr(out), xu.size, numits))
        else:
            check(self.lib, self.lib.afb_piecewise_clenshaw(_ptr(coef), _ptr(offs), _ptr(brk), len(pieces), _ptr(xu), _ptr(out), xu.size))
        return out

    def fourier_cheb2d(self, Cm, t, y) -> np.ndarray:
        """Cm[i, j]: coefficient of F_i(t) T_j(y) (Fourier order in i); t in [0, 2 pi), y in [-1, 1]."""
        Cm = np.asarray(Cm, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if t.size != y.size:
            raise DimensionMismatch(AFB_BAD_SIZE, "fourier_cheb2d: t and y differ in length")
        flat = np.ascontiguousarray(Cm.T)
        out = np.empty_like(t)
        check(self.lib, self.lib.afb_clenshaw_fourier_cheb2d(_ptr(flat), Cm.shape[0], Cm.shape[1], _ptr(t), _ptr(y), _ptr(out), t.size))
        return out

    def sample2d_lowrank(self, A, CB, ya, yb, xa, xb, ry, u) -> np.ndarray:
        """A: (NA, R), CB: (N, R) coefficient matrices (columns = functions); returns rx."""
        A = np.asarray(A, dtype=np.float64)
        CB = np.asarray(CB, dtype=np.float64)
        ry = np.ascontiguousarray(ry, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        if A.shape[1] != CB.shape[1] or ry.size != u.size:
            raise DimensionMismatch(AFB_BAD_SIZE, "sample2d_lowrank: rank / sample-count mismatch")
        Af, Cf = np.ascontiguousarray(A.T), np.ascontiguousarray(CB.T)
        rx = np.empty_like(u)
        check(self.lib, self.lib.afb_sample2d_cheb_lowrank(_ptr(Af), A.shape[0], _ptr(Cf), CB.shape[0], A.shape[1],
                                                           ya, yb, xa, xb, _ptr(ry), _ptr(u), _ptr(rx), u.size))
        return rx

    def dgemm_nn(self, CB, fA) -> np.ndarray:
        CB = np.asarray(CB, dtype=np.float64)
        fA = np.asarray(fA, dtype=np.float64)
        N, R = CB.shape
        R2, P = fA.shape
        if R != R2:
            raise DimensionMismatch(AFB_BAD_SIZE, "dgemm_nn: inner dimensions differ")
        a, b = np.ascontiguousarray(CB.T), np.ascontiguousarray(fA.T)
        out = np.empty((P, N))
        check(self.lib, self.lib.afb_dgemm_nn(_ptr(a), _ptr(b), _ptr(out), N, R, P))
        return np.ascontiguousarray(out.T)

    def samplecdf(self, cdf_c, a: float, b: float, u) -> np.ndarray:
        c = np.ascontiguousarray(cdf_c, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty_like(u)
        check(self.lib, self.lib.afb_samplecdf_cheb(_ptr(c), c.size, a, b, _ptr(u), _ptr(out), u.size))
        return out
