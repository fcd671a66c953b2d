"""ctypes front end of oracle/liboracle_cpu.so (TEST INFRASTRUCTURE ONLY, see oracle_cpu.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_cpu.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_cpu.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i64, i32 = C.c_int64, C.c_int
        L.orc_clenshaw.argtypes = [_dp, i64, _dp, _dp, i64, i32]
        L.orc_clenshaw_planned.argtypes = [_dp, i64, _dp, _dp, i64, _dp, i32]
        L.orc_clenshaw_cols.argtypes = [_dp, i64, i64, _dp, _dp, i64, i32]
        L.orc_clenshaw_multi.argtypes = [_dp, i64, i64, _dp, _dp, i64, i32]
        L.orc_cheb_bisect.argtypes = [_dp, i64, _dp, _dp, i64, i32, i32]
        L.orc_cheb_bisect_cols.argtypes = [_dp, i64, i64, _dp, _dp, i64, i32, i32]
        L.orc_cheb_bisect_planned.argtypes = [_dp, i64, _dp, _dp, i64, i32, _dp, i32]
        L.orc_cheb_normcumsum.argtypes = [_dp, i64, i64, i64, i32]
        L.orc_sample2d_lowrank.argtypes = [_dp, i64, _dp, i64, i64, C.c_double, C.c_double, C.c_double, C.c_double,
                                           _dp, _dp, _dp, i64, i32]
        L.orc_sample2d_lowrank.restype = None
        L.orc_clenshaw2d.argtypes = [_dp, i64, i64, _dp, _dp, _dp, i64, i32]
        L.orc_clenshaw2d.restype = None
        L.orc_clenshaw_rec.argtypes = [_dp, i64, _dp, _dp, _dp, _dp, _dp, i64, i32]
        L.orc_clenshaw_rec.restype = None
        L.orc_max_threads.restype = i32
        for f in (L.orc_clenshaw, L.orc_clenshaw_planned, L.orc_clenshaw_cols, L.orc_clenshaw_multi,
                  L.orc_cheb_bisect, L.orc_cheb_bisect_cols, L.orc_cheb_bisect_planned,
                  L.orc_cheb_normcumsum):
            f.restype = None
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def clenshaw(c, x, threads: int = 1):
    c, x = _c(c), _c(x)
    y = np.empty_like(x)
    lib().orc_clenshaw(c, c.size, x.ravel(), y.ravel(), x.size, threads)
    return y


def clenshaw_planned(c, x, threads: int = 1):
    """Reference loop order (k outer, i inner) through plan work vectors."""
    c, x = _c(c), _c(x).copy()
    y = np.empty_like(x)
    work = np.empty(3 * x.size)
    lib().orc_clenshaw_planned(c, c.size, x, y, x.size, work, threads)
    return y


def clenshaw_cols(Cm, x, threads: int = 1):
    """Cm: (N, P) array in the reference's column-major sense -> pass as Fortran order."""
    Cm = np.asarray(Cm, dtype=np.float64)
    N, P = Cm.shape
    flat = np.ascontiguousarray(Cm.T).ravel()  # column j contiguous
    x = _c(x)
    y = np.empty(P)
    lib().orc_clenshaw_cols(flat, N, N, x, y, P, threads)
    return y


def clenshaw_multi(Cm, x, threads: int = 1):
    Cm = np.asarray(Cm, dtype=np.float64)
    N, M = Cm.shape
    flat = np.ascontiguousarray(Cm.T).ravel()
    x = _c(x)
    Y = np.empty((M, x.size))
    lib().orc_clenshaw_multi(flat, N, M, x, Y.ravel(), x.size, threads)
    return Y


def clenshaw_rec(c, A, B, Cc, x, threads: int = 1):
    """General three-term Clenshaw with true fma (UPSTREAM `clenshaw(sp::PolynomialSpace, c, x)` operation order)."""
    c, A, B, Cc, x = _c(c), _c(A), _c(B), _c(Cc), _c(x)
    assert A.size >= c.size and B.size >= c.size and Cc.size >= c.size + 1
    y = np.empty_like(x)
    lib().orc_clenshaw_rec(c, c.size, A, B, Cc, x, y, x.size, threads)
    return y


def cheb_bisect(c, u, numits: int = 47, threads: int = 1):
    c, u = _c(c), _c(u)
    out = np.empty_like(u)
    lib().orc_cheb_bisect(c, c.size, u, out, u.size, numits, threads)
    return out


def cheb_bisect_planned(c, u, numits: int = 47, threads: int = 1):
    c, u = _c(c), _c(u)
    out = np.empty_like(u)
    work = np.empty(7 * u.size)
    lib().orc_cheb_bisect_planned(c, c.size, u, out, u.size, numits, work, threads)
    return out


def cheb_bisect_cols(Cm, u, numits: int = 47, threads: int = 1):
    Cm = np.asarray(Cm, dtype=np.float64)
    N, P = Cm.shape
    flat = np.ascontiguousarray(Cm.T).ravel()
    u = _c(u)
    out = np.empty(P)
    lib().orc_cheb_bisect_cols(flat, N, N, u, out, P, numits, threads)
    return out


def cheb_normcumsum(Cm, grow: bool | None = None):
    """Vector (grows by one) or (N, ncols) matrix (does not grow)."""
    Cm = np.asarray(Cm, dtype=np.float64)
    if Cm.ndim == 1:
        g = True if grow is None else grow
        N = Cm.size
        buf = np.zeros(N + 1)
        buf[:N] = Cm
        lib().orc_cheb_normcumsum(buf, N, 1, N + 1, int(g))
        return buf if g else buf[:N]
    g = False if grow is None else grow
    N, P = Cm.shape
    ld = N + 1
    buf = np.zeros((P, ld))
    buf[:, :N] = Cm.T
    lib().orc_cheb_normcumsum(buf.ravel(), N, P, ld, int(g))
    L = N + 1 if g else N
    return np.ascontiguousarray(buf[:, :L].T)


def sample2d_lowrank(A, CB, ya, yb, xa, xb, ry, u, threads: int = 1):
    """A: (NA, R), CB: (N, R) with columns = functions; returns rx (sample.jl:208-213 given ry and u)."""
    A = np.asarray(A, dtype=np.float64)
    CB = np.asarray(CB, dtype=np.float64)
    ry, u = _c(ry), _c(u)
    rx = np.empty_like(u)
    Af, Cf = np.ascontiguousarray(A.T).ravel(), np.ascontiguousarray(CB.T).ravel()
    lib().orc_sample2d_lowrank(Af if Af.size else np.zeros(1), A.shape[0], Cf if Cf.size else np.zeros(1), CB.shape[0],
                               A.shape[1], ya, yb, xa, xb, ry, u, rx, u.size, threads)
    return rx


def clenshaw2d(Cm, x, y, threads: int = 1):
    """Cm[i, j] multiplies T_i(x) T_j(y) (ProductFun evaluation order)."""
    Cm = np.asarray(Cm, dtype=np.float64)
    x, y = _c(x), _c(y)
    out = np.empty_like(x)
    flat = np.ascontiguousarray(Cm.T).ravel()
    lib().orc_clenshaw2d(flat if flat.size else np.zeros(1), Cm.shape[0], Cm.shape[1], x, y, out, x.size, threads)
    return out
