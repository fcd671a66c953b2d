"""Host-side mirror of the ApproxFun API for the transform / evaluate / sample path.

Same names, argument meaning and error behaviour as the reference (file:line under /root/reference):
`plan_transform` / `plan_itransform` / `plan*v` (NEWS.md:90, src/ApproxFun.jl:22-23), `transform`,
`itransform`, `points` (NEWS.md:92), `Fun(f, space, n)`, `coefficients`, `values`
(docs/src/faq.md:14-17, ext/ApproxFunDualNumbersExt.jl:79-82), `clenshaw`, `ClenshawPlan`
(src/ApproxFun.jl:26), `bisectioninv`, `chebbisectioninv`, `chebnormalizedcumsum`, `normalizedcumsum`,
`samplecdf`, `sample` (src/Extras/sample.jl).  All arithmetic on the hot path runs in
libapproxfun_b200.so on the GPU; this layer only orchestrates (as the Julia layer does upstream).

Julia is not available in this image, so this Python layer is what the parity tests drive; the Julia
binding a maintainer would add is julia/ApproxFunB200.jl.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L

_ll = None


def lowlevel() -> L.LowLevel:
    global _ll
    if _ll is None:
        lib = L.load()
        L.init(lib, None)
        _ll = L.LowLevel(lib)
    return _ll


def use_devices(devices=None) -> None:
    """Select the GPUs of this process (`afb_init`): an int, a list of ordinals (single-process multi-GPU: host-array calls
    fan out over them), or None for the current device."""
    L.init(lowlevel().lib, devices)


class pinned:
    """`with pinned(a, b): ...` page-locks caller-owned NumPy arrays in place (`afb_host_register`) for the duration of the
    block, so host-array calls DMA them directly instead of staging them through the library's pinned ring; they are
    unregistered on exit.  Also usable without `with`: `h = pinned(a); ...; h.release()`."""

    def __init__(self, *arrays):
        import ctypes as C
        self._lib = lowlevel().lib
        self._ptrs = []
        try:
            for a in arrays:
                if not (a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]):
                    raise L.AfbError(L.AFB_BAD_ARG, "pinned: array must be contiguous")
                L.check(self._lib, self._lib.afb_host_register(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes)))
                self._ptrs.append(a.ctypes.data)
        except Exception:
            self.release()
            raise

    def release(self) -> None:
        import ctypes as C
        while self._ptrs:
            L.check(self._lib, self._lib.afb_host_unregister(C.c_void_p(self._ptrs.pop())))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()
        return False


# ------------------------------------------------------------------------------------------------------
# spaces
# ------------------------------------------------------------------------------------------------------
class Space:
    pass


class Chebyshev(Space):
    """`Chebyshev(a..b)`; canonical domain -1..1."""

    def __init__(self, a: float = -1.0, b: float = 1.0):
        self.a, self.b = float(a), float(b)

    def __repr__(self):
        return f"Chebyshev({self.a}..{self.b})"

    def fromcanonical(self, x):
        return (self.a + self.b) / 2.0 + (self.b - self.a) * np.asarray(x, dtype=np.float64) / 2.0

    def tocanonical(self, x):
        return (2.0 * np.asarray(x, dtype=np.float64) - self.a - self.b) / (self.b - self.a)

    def __pow__(self, k):
        if k != 2:
            raise ValueError("only Chebyshev()^2 is supported")
        return TensorSpace(self, self)

    def __mul__(self, other):
        return TensorSpace(self, other)


class PolynomialSpace(Space):
    """Orthogonal-polynomial spaces evaluated through the general three-term Clenshaw kernel (`afb_clenshaw_rec`):
    p_{j+1} = (recA(j) x + recB(j)) p_j - recC(j) p_{j-1}  (UPSTREAM ApproxFunOrthogonalPolynomials recA/recB/recC)."""

    def __init__(self, a: float = -1.0, b: float = 1.0):
        self.a, self.b = float(a), float(b)

    fromcanonical = Chebyshev.fromcanonical
    tocanonical = Chebyshev.tocanonical

    def rec(self, N: int):
        """(A[0..N-1], B[0..N-1], C[0..N])"""
        raise NotImplementedError


class Jacobi(PolynomialSpace):
    """`Jacobi(b, a)`: P_n^{(a,b)} (the reference's argument order), DLMF 18.9.2."""

    def __init__(self, b: float, a: float, lo: float = -1.0, hi: float = 1.0):
        super().__init__(lo, hi)
        self.jb, self.ja = float(b), float(a)

    def rec(self, N: int):
        a, b = self.ja, self.jb
        n = np.arange(N + 1, dtype=np.float64)
        s = 2 * n + a + b
        with np.errstate(divide="ignore", invalid="ignore"):
            A = (s + 1) * (s + 2) / (2 * (n + 1) * (n + a + b + 1))
            Bv = (a * a - b * b) * (s + 1) / (2 * (n + 1) * (n + a + b + 1) * s)
            Cv = (n + a) * (n + b) * (s + 2) / ((n + 1) * (n + a + b + 1) * s)
        A[0], Bv[0], Cv[0] = (a + b + 2) / 2.0, (a - b) / 2.0, 0.0
        return A[:N].copy(), Bv[:N].copy(), Cv


class Legendre(Jacobi):
    def __init__(self, lo: float = -1.0, hi: float = 1.0):
        super().__init__(0.0, 0.0, lo, hi)


class Ultraspherical(PolynomialSpace):
    """`Ultraspherical(lam)`: C_n^{(lam)}; A_n = 2(n+lam)/(n+1), B_n = 0, C_n = (n+2lam-1)/(n+1)."""

    def __init__(self, lam: float, lo: float = -1.0, hi: float = 1.0):
        super().__init__(lo, hi)
        self.lam = float(lam)

    def rec(self, N: int):
        n = np.arange(N + 1, dtype=np.float64)
        return (2 * (n + self.lam) / (n + 1))[:N].copy(), np.zeros(N), (n + 2 * self.lam - 1) / (n + 1)


class Fourier(Space):
    def __init__(self, a: float = 0.0, b: float = 2.0 * math.pi):
        self.a, self.b = float(a), float(b)

    def tocanonical(self, x):
        """PeriodicSegment(a, b) -> [0, 2 pi)"""
        return 2.0 * math.pi * (np.asarray(x, dtype=np.float64) - self.a) / (self.b - self.a)

    def __mul__(self, other):
        return TensorSpace(self, other)


class Laurent(Space):
    def __init__(self, a: float = 0.0, b: float = 2.0 * math.pi):
        self.a, self.b = float(a), float(b)

    def tocanonical(self, x):
        """PeriodicSegment(a, b) -> [0, 2 pi)"""
        return 2.0 * math.pi * (np.asarray(x, dtype=np.float64) - self.a) / (self.b - self.a)

    def __mul__(self, other):
        return TensorSpace(self, other)


class CosSpace(Space):
    """`CosSpace()`: f(theta) = sum_k f_k cos(k theta); same coefficients and transform as `Chebyshev`
    (docs/src/usage/spaces.md:39-49), on theta_j = pi (j + 1/2) / n."""

    def __init__(self, a: float = 0.0, b: float = 2.0 * math.pi):
        self.a, self.b = float(a), float(b)


class SinSpace(Space):
    """`SinSpace()`: f(theta) = sum_k f_k sin((k+1) theta) on theta_j = pi j/(n+1), j = 1..n
    (plans: src/Extras/fftGeneric.jl:74-81)."""

    def __init__(self, a: float = 0.0, b: float = 2.0 * math.pi):
        self.a, self.b = float(a), float(b)


class Taylor(Space):
    """`Taylor()` = `Hardy{true}`: f(t) = sum_k f_k e^{ikt} (docs/src/usage/spaces.md:100-104)."""

    def __init__(self, a: float = 0.0, b: float = 2.0 * math.pi):
        self.a, self.b = float(a), float(b)


class Hardy(Space):
    """`Hardy{false}`: f(t) = sum_k f_k e^{-i(k+1)t} (docs/src/usage/spaces.md:106-110)."""

    def __init__(self, a: float = 0.0, b: float = 2.0 * math.pi):
        self.a, self.b = float(a), float(b)


class TensorSpace(Space):
    """`S1 ⊗ S2` (`S1 * S2` here): Chebyshev ⊗ Chebyshev, and the periodic-times-interval spaces Fourier ⊗ Chebyshev and
    Laurent ⊗ Chebyshev (`PeriodicSegment() × ChebyshevInterval()`, test/PeriodicXIntervalTest.jl:22, LaplaceInAStripTest.jl:5)."""

    def __init__(self, s1: Space, s2: Space):
        if not (isinstance(s1, (Chebyshev, Fourier, Laurent)) and isinstance(s2, Chebyshev)):
            raise TypeError("TensorSpace: GPU paths exist for (Chebyshev | Fourier | Laurent) (x) Chebyshev")
        self.factors = (s1, s2)

    def kinds(self):
        s1 = self.factors[0]
        if isinstance(s1, Fourier):
            return L.FOURIER_CHEB1_2D_FWD, L.FOURIER_CHEB1_2D_INV
        if isinstance(s1, Laurent):
            return L.LAURENT_CHEB1_2D_FWD, L.LAURENT_CHEB1_2D_INV
        return L.CHEB1_2D_FWD, L.CHEB1_2D_INV


_FWD = {Chebyshev: L.CHEB1_FWD, Fourier: L.FOURIER_FWD, Laurent: L.LAURENT_FWD, TensorSpace: L.CHEB1_2D_FWD,
        CosSpace: L.CHEB1_FWD, SinSpace: L.SIN_FWD, Taylor: L.TAYLOR_FWD, Hardy: L.HARDYM_FWD}
_INV = {Chebyshev: L.CHEB1_INV, Fourier: L.FOURIER_INV, Laurent: L.LAURENT_INV, TensorSpace: L.CHEB1_2D_INV,
        CosSpace: L.CHEB1_INV, SinSpace: L.SIN_INV, Taylor: L.TAYLOR_INV, Hardy: L.HARDYM_INV}


def points(space: Space, n: int, m: int | None = None):
    """`points(space, n)`.  TensorSpace: `points(sp, N, M)` returns the two 1-D grids of the tensor grid;
    `points(sp, N)` returns the Padua points carrying N coefficients as an Np x 2 array [x y] (NEWS.md:85)."""
    ll = lowlevel()
    if isinstance(space, Chebyshev):
        return ll.chebpoints(n, space.a, space.b)
    if isinstance(space, (Fourier, Laurent, Taylor, Hardy)):
        return ll.fourierpoints(n, space.a, space.b)
    if isinstance(space, TensorSpace):
        s1, s2 = space.factors
        if m is None:
            return ll.paduapoints(int(n), s1.a, s1.b, s2.a, s2.b)
        return points(s1, n), points(s2, m)
    if isinstance(space, CosSpace):
        return math.pi * (np.arange(n) + 0.5) / n
    if isinstance(space, SinSpace):
        return math.pi * np.arange(1, n + 1) / (n + 1)
    raise TypeError(space)


# ------------------------------------------------------------------------------------------------------
# plans:  plan = plan_transform(S, v_or_n);  coefficients = plan * v
# ------------------------------------------------------------------------------------------------------
class DualArray:
    """Vector of dual numbers value + epsilon*eps (DualNumbers.jl).  Mirrors ext/ApproxFunDualNumbersExt.jl:44-63: a plan
    for a Dual vector is the plan of its real part, and `P * v = dual.(P*realpart.(v), P*dualpart.(v))`."""

    def __init__(self, value, epsilon):
        self.value = np.asarray(value, dtype=np.float64)
        self.epsilon = np.asarray(epsilon, dtype=np.float64)
        if self.value.shape != self.epsilon.shape:
            raise L.DimensionMismatch(L.AFB_BAD_SIZE, "DualArray: value and epsilon shapes differ")

    @property
    def shape(self):
        return self.value.shape


def realpart(d: DualArray):
    return d.value


def dualpart(d: DualArray):
    return d.epsilon


class TransformPlan:
    """`TransformPlan` / `ITransformPlan`: owns a C-ABI plan handle; `plan * v` executes it."""

    inverse = False

    def __init__(self, space: Space, shape):
        self.space = space
        self.shape = tuple(int(s) for s in shape)
        self._ll = lowlevel()
        table = _INV if self.inverse else _FWD
        self.kind = table[type(space)]
        if isinstance(space, TensorSpace) and len(self.shape) == 1:
            # vector of values at the Padua points <-> flat coefficients ordered by total degree
            self.kind = L.PADUA_INV if self.inverse else L.PADUA_FWD
            self._h = self._ll.plan_create(self.kind, self.shape[0], 0, 1, self.shape[0])
        elif isinstance(space, TensorSpace):
            if len(self.shape) != 2:
                raise L.DimensionMismatch(L.AFB_BAD_SIZE, "TensorSpace plans take an n x m matrix or a Padua vector")
            n, n2 = self.shape
            self.kind = space.kinds()[1 if self.inverse else 0]
            self._h = self._ll.plan_create(self.kind, n, n2, 1, n)
        else:
            n = self.shape[0]
            batch = int(np.prod(self.shape[1:])) if len(self.shape) > 1 else 1
            self._h = self._ll.plan_create(self.kind, n, 0, batch, n)
        cplx = isinstance(space, (Laurent, Taylor, Hardy)) or (isinstance(space, TensorSpace) and isinstance(space.factors[0], Laurent))
        self.dtype = np.complex128 if cplx else np.float64

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            try:
                self._ll.plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def __mul__(self, v):
        if isinstance(v, DualArray):
            return DualArray(self * v.value, self * v.epsilon)
        v = np.asarray(v)
        if v.shape != self.shape:
            raise L.DimensionMismatch(L.AFB_BAD_SIZE, f"plan built for shape {self.shape}, got {v.shape}")
        if np.iscomplexobj(v) and self.dtype == np.float64:
            # linear: transform real and imaginary parts separately (what the reference does for Dual/complex)
            return (self * v.real) + 1j * (self * v.imag)
        if v.size == 0:
            return v.astype(self.dtype)
        a = np.asarray(v, dtype=self.dtype)
        buf = np.ascontiguousarray(a.T) if a.ndim > 1 else np.ascontiguousarray(a)
        out = np.empty_like(buf)
        self._ll.plan_execute_host(self._h, buf, out)
        return np.ascontiguousarray(out.T) if a.ndim > 1 else out

    @property
    def launches(self) -> int:
        return self._ll.plan_launches(self._h)


class ITransformPlan(TransformPlan):
    inverse = True


def _shape_of(v_or_n):
    if isinstance(v_or_n, (int, np.integer)):
        return (int(v_or_n),)
    if isinstance(v_or_n, tuple):
        return v_or_n
    if isinstance(v_or_n, DualArray):
        return v_or_n.shape
    return np.asarray(v_or_n).shape


def plan_transform(space: Space, v_or_n) -> TransformPlan:
    return TransformPlan(space, _shape_of(v_or_n))


def plan_itransform(space: Space, v_or_n) -> ITransformPlan:
    return ITransformPlan(space, _shape_of(v_or_n))


def chebyshevtransform(v, kind: int = 1):
    """UPSTREAM FastTransforms `chebyshevtransform(v, Val(kind))`: kind 1 = first-kind points (REDFT10),
    kind 2 = second-kind points cos(pi j/(n-1)) (REDFT00)."""
    return lowlevel().transform(L.CHEB1_FWD if kind == 1 else L.CHEB2_FWD, v)


def ichebyshevtransform(c, kind: int = 1):
    return lowlevel().transform(L.CHEB1_INV if kind == 1 else L.CHEB2_INV, c)


_plan_cache: dict = {}


def _cached_plan(cls, space: Space, shape):
    """`transform(S, v)` builds a plan per call in the reference (FFTW ESTIMATE plans are cheap); here plans own
    device tables, so the convenience entry points keep the most recent ones."""
    key = (cls, type(space), tuple(type(f) for f in getattr(space, "factors", ())), tuple(shape))
    p = _plan_cache.get(key)
    if p is None:
        if len(_plan_cache) >= 64:
            _plan_cache.pop(next(iter(_plan_cache)))
        p = _plan_cache[key] = cls(space, shape)
    return p


def transform(space: Space, v):
    """`transform(S, vals) = plan_transform(S, vals) * vals`."""
    return _cached_plan(TransformPlan, space, _shape_of(v)) * v


def itransform(space: Space, c):
    return _cached_plan(ITransformPlan, space, _shape_of(c)) * c


# ------------------------------------------------------------------------------------------------------
# Clenshaw
# ------------------------------------------------------------------------------------------------------
class ClenshawPlan:
    """`ClenshawPlan(Float64, Chebyshev(), N, n)`: kept for signature parity; the GPU kernel needs no
    host work vectors, and unlike the reference the result never aliases plan storage."""

    def __init__(self, T=np.float64, space: Space | None = None, N: int = 0, n: int = 0):
        self.space = space if space is not None else Chebyshev()
        self.N, self.n = int(N), int(n)


def clenshaw(*args):
    """`clenshaw(sp::Chebyshev, c, x)` (x scalar or vector, canonical variable) or
    `clenshaw(c, x, plan::ClenshawPlan)` with c a vector (all points) or an N x n matrix (column j at x[j])."""
    if isinstance(args[0], Space):
        _, c, x = args
    else:
        c, x = args[0], args[1]
    ll = lowlevel()
    c = np.asarray(c, dtype=np.float64)
    scalar = np.ndim(x) == 0
    xv = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if c.ndim == 2:
        if c.shape[1] != xv.size:
            raise L.DimensionMismatch(L.AFB_BAD_SIZE, "clenshaw(c::Matrix, x): size(c,2) != length(x)")
        y = ll.clenshaw_cols(c, xv)
    else:
        y = ll.clenshaw(c, xv.ravel()).reshape(xv.shape)
    return float(y[0]) if scalar else y


# ------------------------------------------------------------------------------------------------------
# Fun
# ------------------------------------------------------------------------------------------------------
class Fun:
    """`Fun(f, space, n)`, `Fun(space, coefficients)`, adaptive `Fun(f, space)`; `f(x)` evaluates."""

    def __init__(self, *args):
        if len(args) == 2 and isinstance(args[0], Space):
            cplx = isinstance(args[0], Laurent) or (isinstance(args[0], TensorSpace) and isinstance(args[0].factors[0], Laurent))
            self.space, self.coefficients = args[0], np.array(args[1], dtype=np.complex128 if cplx else np.float64)
        elif len(args) == 3:
            f, space, n = args
            self.space = space
            if isinstance(space, TensorSpace) and isinstance(n, (int, np.integer)):
                # Fun(f, Chebyshev()^2, N): values at the Padua points -> flat coefficients by total degree
                pts = points(space, int(n))
                self.coefficients = transform(space, np.asarray(f(pts[:, 0], pts[:, 1]), dtype=np.float64))
            elif isinstance(space, TensorSpace):
                n1, n2 = n  # tensor grid, coefficient MATRIX (ProductFun layout)
                x, y = points(space, n1, n2)
                self.coefficients = transform(space, f(x[:, None], y[None, :]))
            else:
                self.coefficients = transform(space, np.asarray(f(points(space, int(n)))))
        elif len(args) == 2 and callable(args[0]):
            f, space = args
            self.space = space
            self.coefficients = _adaptive(f, space)
        else:
            raise TypeError("Fun(f, space, n) | Fun(f, space) | Fun(space, coefficients)")

    def __call__(self, x, y=None):
        sp = self.space
        if isinstance(sp, Chebyshev):
            xa = np.asarray(x, dtype=np.float64)
            if np.any(xa < min(sp.a, sp.b)) or np.any(xa > max(sp.a, sp.b)):
                # the reference returns zero outside the domain
                inside = (xa >= min(sp.a, sp.b)) & (xa <= max(sp.a, sp.b))
                res = np.zeros_like(xa)
                if np.any(inside):
                    res[inside] = clenshaw(sp, self.coefficients, sp.tocanonical(xa[inside]))
                return res
            return clenshaw(sp, self.coefficients, sp.tocanonical(x))
        if isinstance(sp, PolynomialSpace):
            xa = np.atleast_1d(np.asarray(x, dtype=np.float64))
            A, Bv, Cv = sp.rec(len(self.coefficients))
            out = lowlevel().clenshaw_rec(self.coefficients, A, Bv, Cv, sp.tocanonical(xa.ravel())).reshape(xa.shape)
            return float(out[0]) if np.ndim(x) == 0 else out
        if isinstance(sp, TensorSpace):
            s1, s2 = sp.factors
            xs, ys = np.broadcast_arrays(np.atleast_1d(np.asarray(x, dtype=np.float64)),
                                         np.atleast_1d(np.asarray(y, dtype=np.float64)))
            if isinstance(s1, Fourier):
                out = lowlevel().fourier_cheb2d(self.coefficients, s1.tocanonical(xs.ravel()), s2.tocanonical(ys.ravel()))
                out = out.reshape(xs.shape)
                return float(out.ravel()[0]) if (np.ndim(x) == 0 and np.ndim(y) == 0) else out
            if isinstance(s1, Laurent):
                raise NotImplementedError("evaluation of Laurent (x) Chebyshev Funs: transform plans only")
            cm = self.coefficients if self.coefficients.ndim == 2 else totaldegree_to_matrix(self.coefficients)
            out = lowlevel().clenshaw2d(cm, s1.tocanonical(xs.ravel()), s2.tocanonical(ys.ravel()))
            out = out.reshape(xs.shape)
            return float(out.ravel()[0]) if (np.ndim(x) == 0 and np.ndim(y) == 0) else out
        raise NotImplementedError("evaluation is GPU-accelerated for Chebyshev spaces only")

    def __len__(self):
        return len(self.coefficients)


def totaldegree_to_matrix(c) -> np.ndarray:
    """Flat TensorSpace coefficients (T0T0, T0(x)T1(y), T1(x)T0(y), ...: docs/src/usage/constructors.md:98-116) ->
    coefficient matrix C[i, j] of T_i(x) T_j(y) (`coefficientmatrix(f)`); zero-padded to the next full diagonal."""
    c = np.asarray(c, dtype=np.float64)
    d = 0
    while (d + 1) * (d + 2) // 2 < c.size:
        d += 1
    s = np.repeat(np.arange(d + 1), np.arange(1, d + 2))[: c.size]   # diagonal of every entry
    i = np.arange(c.size) - s * (s + 1) // 2                           # x-degree within the diagonal
    cm = np.zeros((d + 1, d + 1))
    cm[i, s - i] = c
    return cm


def matrix_to_totaldegree(cm) -> np.ndarray:
    """Inverse of `totaldegree_to_matrix` (entries with i + j beyond the last full diagonal must be zero)."""
    cm = np.asarray(cm, dtype=np.float64)
    d = cm.shape[0] + cm.shape[1] - 2
    s = np.repeat(np.arange(d + 1), np.arange(1, d + 2))
    i = np.arange(s.size) - s * (s + 1) // 2
    j = s - i
    ok = (i < cm.shape[0]) & (j < cm.shape[1])
    out = np.zeros(s.size)
    out[ok] = cm[i[ok], j[ok]]
    return out


def coefficients(f: Fun):
    return f.coefficients


def ncoefficients(f: Fun) -> int:
    return len(f.coefficients)


def space(f: Fun) -> Space:
    return f.space


def values(f: Fun):
    """`values(f) = itransform(space(f), coefficients(f))` (src/Plot/Plot.jl:20, test/ExtrasTest.jl:61)."""
    return itransform(f.space, f.coefficients)


def _adaptive(f, sp: Space):
    """Adaptive constructor ladder n = 2^4 .. 2^20, then 2^21 (in-tree mirror ext/ApproxFunDualNumbersExt.jl:98-110)."""
    if not isinstance(sp, Chebyshev):
        raise NotImplementedError("adaptive construction: Chebyshev only")
    tol = 20 * np.finfo(np.float64).eps
    for logn in range(4, 21):
        c = transform(sp, np.asarray(f(points(sp, 2 ** logn))))
        mx = float(np.max(np.abs(c)))
        if mx == 0.0:
            return np.zeros(1)
        if float(np.max(np.abs(c[-8:]))) < tol * mx:
            nz = np.nonzero(np.abs(c) > tol * mx)[0]
            return c[: nz[-1] + 1].copy()
    return transform(sp, np.asarray(f(points(sp, 2 ** 21))))


# ------------------------------------------------------------------------------------------------------
# sampling (src/Extras/sample.jl)
# ------------------------------------------------------------------------------------------------------
def chebnormalizedcumsum(f, grow=None):
    """`chebnormalizedcumsum!(f)` (sample.jl:170-175); returns the new array (vectors grow by one)."""
    return lowlevel().cheb_normcumsum(f, grow)


def chebbisectioninv(c, x, plan: ClenshawPlan | None = None, numits: int = 47):
    """`chebbisectioninv(c, x[, plan])` for vector or matrix `c` (sample.jl:41-101)."""
    ll = lowlevel()
    c = np.asarray(c, dtype=np.float64)
    scalar = np.ndim(x) == 0
    xv = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if c.ndim == 2:
        if c.shape[1] != xv.size:
            raise L.DimensionMismatch(L.AFB_BAD_SIZE, "chebbisectioninv: size(c)[2] == length(xl) violated (sample.jl:85)")
        out = ll.cheb_bisect_cols(c, xv, numits)
    else:
        out = ll.cheb_bisect(c, xv, numits)
    return float(out[0]) if scalar else out


def cumsum(f: Fun) -> Fun:
    """`cumsum(f::Fun{<:Chebyshev})` = integrate(f) - first(integrate(f)) (host: O(N) on a handful of numbers)."""
    sp = f.space
    c = np.asarray(f.coefficients, dtype=np.float64)
    n = c.size
    v = c.copy()
    if n == 2:
        v[1] /= 2
    elif n > 2:
        v[0] -= v[2] / 2
        v[1:n - 2] = (v[1:n - 2] - v[3:n]) / 2
        v[n - 2] /= 2
        v[n - 1] /= 2
    g = np.zeros(n + 1)
    g[1:] = v / np.arange(1, n + 1)
    g *= (sp.b - sp.a) / 2.0
    sgn = np.where(np.arange(n + 1) % 2 == 0, 1.0, -1.0)
    g[0] -= float(np.sum(sgn * g))
    return Fun(sp, g)


def normalizedcumsum(f):
    """sample.jl:114-119: cf = cumsum(f); cf / last(cf)."""
    if isinstance(f, PiecewiseFun):
        cf = _piecewise_cumsum(f)
        last = float(np.sum(cf.pieces[-1].coefficients))
        return PiecewiseFun([Fun(p.space, p.coefficients / last) for p in cf.pieces])
    cf = cumsum(f)
    return Fun(cf.space, cf.coefficients / float(np.sum(cf.coefficients)))


def bisectioninv(cf: Fun, x, numits: int = 47):
    """sample.jl:103-109: fromcanonical.(space(cf), chebbisectioninv(coefficients(cf), x))."""
    if isinstance(cf, PiecewiseFun):   # no Chebyshev dispatch: the generic method (sample.jl:8, 23-36)
        return bisectioninv_generic(cf, x, numits)
    sp = cf.space
    if not isinstance(sp, Chebyshev):
        raise NotImplementedError("bisectioninv: the GPU path covers Fun{<:Chebyshev,Float64}")
    scalar = np.ndim(x) == 0
    xv = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if numits == 47:
        out = lowlevel().samplecdf(cf.coefficients, sp.a, sp.b, xv)
    else:
        out = sp.fromcanonical(chebbisectioninv(cf.coefficients, xv, numits=numits))
    return float(out[0]) if scalar else out


def samplecdf(cf: Fun, n: int, rng=None, uniforms=None):
    """sample.jl:184: bisectioninv(cf, rand(n)).  `uniforms` replaces `rand(n)` for reproducible parity."""
    if uniforms is None:
        uniforms = (rng or np.random.default_rng()).random(n)
    return bisectioninv(cf, np.asarray(uniforms, dtype=np.float64))


def genmidpoint(a: float, b: float) -> float:
    """sample.jl:11-21: midpoint that also works for infinite end points."""
    if math.isinf(a) and math.isinf(b):
        return 0.0
    if math.isinf(a):
        return b - 100
    if math.isinf(b):
        return a + 100
    return (a + b) / 2


def bisectioninv_generic(f, x, numits: int = 47):
    """sample.jl:8, 23-36: the generic (any-space) `bisectioninv(f::Fun, x; numits=47)` -- bisection in the DOMAIN variable
    between minimum(d) and maximum(d) with `genmidpoint`; one fused GPU kernel (`afb_piecewise_bisect`) for Chebyshev and
    piecewise-Chebyshev Funs."""
    scalar = np.ndim(x) == 0
    xv = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if isinstance(f, PiecewiseFun):
        pieces, breaks = [p.coefficients for p in f.pieces], f.breaks
    elif isinstance(f.space, Chebyshev):
        if not f.space.a < f.space.b:
            raise NotImplementedError("bisectioninv: reversed segments")
        pieces, breaks = [f.coefficients], [f.space.a, f.space.b]
    else:
        raise NotImplementedError("bisectioninv: the GPU path covers Chebyshev and piecewise-Chebyshev Funs")
    out = lowlevel().piecewise(pieces, breaks, xv.ravel(), True, numits).reshape(xv.shape)
    return float(out[0]) if scalar else out


class PiecewiseFun:
    """A Fun on a `PiecewiseSpace` of Chebyshev segments (what `abs(Fun(sin, -5..5))` returns, test/PiecewiseSampleTest.jl:5):
    `pieces[k]` is a Fun on Chebyshev(breaks[k]..breaks[k+1]).  Evaluation picks the first piece whose closed interval
    contains x (zero outside the domain) and runs on the GPU (`afb_piecewise_clenshaw`)."""

    def __init__(self, pieces):
        self.pieces = list(pieces)
        self.breaks = [self.pieces[0].space.a] + [p.space.b for p in self.pieces]
        for p, q in zip(self.pieces[:-1], self.pieces[1:]):
            if p.space.b != q.space.a:
                raise ValueError("PiecewiseFun: pieces must share break points")

    def __call__(self, x):
        xa = np.atleast_1d(np.asarray(x, dtype=np.float64))
        out = lowlevel().piecewise([p.coefficients for p in self.pieces], self.breaks, xa.ravel(), False).reshape(xa.shape)
        return float(out[0]) if np.ndim(x) == 0 else out


def roots(f: Fun) -> np.ndarray:
    """Real roots of a Chebyshev Fun inside its domain (host: colleague-matrix eigenvalues, as upstream; O(N^3) on a
    handful of numbers -- not part of the data-parallel path)."""
    c = np.trim_zeros(np.asarray(f.coefficients, dtype=np.float64), "b")
    if c.size < 2:
        return np.zeros(0)
    r = np.polynomial.chebyshev.chebroots(c)
    r = np.sort(r[np.abs(r.imag) < 1e-10].real)
    r = r[(r > -1 - 1e-12) & (r < 1 + 1e-12)]
    return f.space.fromcanonical(np.clip(r, -1, 1))


def abs_fun(f: Fun) -> PiecewiseFun:
    """`abs(f::Fun)`: split the domain at the roots of f and take |f| on every piece (each piece built by the adaptive
    constructor through the GPU transform).  Device-resident variant: `DeviceFun.abs`."""
    sp = f.space
    rts = [r for r in roots(f) if sp.a < r < sp.b]
    brk = [sp.a] + rts + [sp.b]
    pieces = []
    for lo, hi in zip(brk[:-1], brk[1:]):
        sgn = 1.0 if float(f(0.5 * (lo + hi))) >= 0 else -1.0
        pieces.append(Fun(lambda t, s=sgn: s * f(t), Chebyshev(lo, hi)))
    return PiecewiseFun(pieces)


def _piecewise_cumsum(f: PiecewiseFun) -> PiecewiseFun:
    """cumsum on a PiecewiseSpace: every piece is integrated and shifted by the total of the pieces to its left."""
    acc, out = 0.0, []
    for p in f.pieces:
        g = cumsum(p)
        c = g.coefficients.copy()
        c[0] += acc
        acc = float(np.sum(c))   # value at the right end point: sum of Chebyshev coefficients
        out.append(Fun(p.space, c))
    return PiecewiseFun(out)


def integrate(f):
    """`integrate(f)` with the reference's normalisation F(left end) = 0 (= cumsum)."""
    return _piecewise_cumsum(f) if isinstance(f, PiecewiseFun) else cumsum(f)


class ProductFun:
    """`ProductFun(f, sp, N, M)`: values on the tensor grid -> coefficient matrix C[i,j] (T_i(x) T_j(y))."""

    def __init__(self, f, sp: "TensorSpace", N: int | None = None, M: int | None = None):
        if isinstance(f, Fun):
            c = np.asarray(f.coefficients)
            self.space, self.coefficients = f.space, (c if c.ndim == 2 else totaldegree_to_matrix(c))
        else:
            g = Fun(f, sp, (N, N if M is None else M))
            self.space, self.coefficients = g.space, g.coefficients

    def __call__(self, x, y):
        return Fun(self.space, self.coefficients)(x, y)


def coefficientmatrix(funs) -> np.ndarray:
    """`coefficientmatrix(::Vector{Fun})`: columns = zero-padded coefficient vectors."""
    n = max(len(f.coefficients) for f in funs)
    out = np.zeros((n, len(funs)))
    for r, f in enumerate(funs):
        out[: len(f.coefficients), r] = f.coefficients
    return out


class LowRankFun:
    """f(x,y) = sum_r A_r(x) B_r(y).  UPSTREAM ApproxFunBase builds it by adaptive cross approximation; here the
    factors come from the SVD of the coefficient matrix (same function, same conditional densities)."""

    def __init__(self, f, tol: float = 1e-15):
        pf = f if isinstance(f, ProductFun) else ProductFun(f, f.space)
        Cm = np.asarray(pf.coefficients, dtype=np.float64)
        U, S, Vt = np.linalg.svd(Cm, full_matrices=False)
        R = max(1, int(np.sum(S > tol * S[0])))
        s1, s2 = pf.space.factors
        self.space = pf.space
        self.A = [Fun(s1, U[:, r] * S[r]) for r in range(R)]
        self.B = [Fun(s2, Vt[r].copy()) for r in range(R)]

    def __call__(self, x, y):
        return sum(a(x) * b(y) for a, b in zip(self.A, self.B))


def _definite_integral(f: Fun) -> float:
    c = np.asarray(f.coefficients, dtype=np.float64)
    k = np.arange(c.size)
    with np.errstate(divide="ignore"):
        w = np.where(k % 2 == 0, 2.0 / (1.0 - k.astype(np.float64) ** 2), 0.0)
    return float(w @ c) * (f.space.b - f.space.a) / 2.0


def lowrank_sum(f: LowRankFun, dim: int) -> Fun:
    """`sum(f::LowRankFun, dim)`: integrate out variable `dim` (1 = x, 2 = y)."""
    if dim == 1:
        w = np.array([_definite_integral(a) for a in f.A])
        return Fun(f.B[0].space, coefficientmatrix(f.B) @ w)
    w = np.array([_definite_integral(b) for b in f.B])
    return Fun(f.A[0].space, coefficientmatrix(f.A) @ w)


def sample2d(f: LowRankFun, n: int, rng=None, uniforms=None) -> np.ndarray:
    """`sample(f::LowRankFun{Chebyshev,Chebyshev}, n)` (sample.jl:205-214), line for line; the conditional stage
    (lines 208-213) is ONE fused GPU kernel that never materialises the N x n matrix AB.  Returns [rx ry]."""
    if uniforms is None:
        rng = rng or np.random.default_rng()
        uniforms = (rng.random(n), rng.random(n))
    u_y, u_x = (np.asarray(u, dtype=np.float64) for u in uniforms)
    ry = sample(lowrank_sum(f, 1), n, uniforms=u_y)                 # ry = sample(sum(f,1), n)
    s1 = f.A[0].space
    rx = lowlevel().sample2d_lowrank(coefficientmatrix(f.A), coefficientmatrix(f.B), s1.a, s1.b, s1.a, s1.b, ry, u_x)
    return np.stack([rx, ry], axis=1)


def sample(f, n: int | None = None, rng=None, uniforms=None, seed: int | None = None):
    """sample.jl:182: samplecdf(normalizedcumsum(f), n); `sample(f)` returns one sample.
    2-D (sample.jl:196,218): Fun on a TensorSpace / ProductFun -> LowRankFun -> `sample2d`.
    `seed` (1-D Chebyshev Funs): draw the uniforms on the device instead of `rand(n)` on the host (Philox4x32-10,
    reproducible; no host-to-device traffic)."""
    if isinstance(f, (ProductFun, LowRankFun)) or (isinstance(f, Fun) and isinstance(f.space, TensorSpace)):
        lr = f if isinstance(f, LowRankFun) else LowRankFun(f)
        out = sample2d(lr, 1 if n is None else n, rng, uniforms)
        return out[0] if n is None else out
    one = n is None
    if seed is not None and uniforms is None and isinstance(f, Fun) and isinstance(f.space, Chebyshev):
        cf = normalizedcumsum(f)
        out = lowlevel().sample_cheb(cf.coefficients, f.space.a, f.space.b, int(seed), 1 if one else int(n))
        return float(out[0]) if one else out
    out = samplecdf(normalizedcumsum(f), 1 if one else n, rng, uniforms)
    return float(out[0]) if one else out


# ------------------------------------------------------------------------------------------------------
# device-resident use (torch tensors or raw pointers): what bench.py times
# ------------------------------------------------------------------------------------------------------
class DevicePlan:
    """Plan executed on device buffers: `execute(in_ptr, out_ptr, stream_ptr)` is asynchronous."""

    def __init__(self, kind: int, n: int, n2: int = 0, batch: int = 1, stride: int | None = None, flags: int = 0):
        self._ll = lowlevel()
        self._h = self._ll.plan_create(kind, n, n2, batch, stride, flags)
        self.kind, self.n, self.n2, self.batch = kind, n, n2, batch

    def execute(self, in_ptr: int, out_ptr: int, stream: int = 0):
        lib = self._ll.lib
        L.check(lib, lib.afb_plan_execute_device(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), C.c_void_p(stream)))

    def execute_host(self, src: np.ndarray, dst: np.ndarray):
        self._ll.plan_execute_host(self._h, src, dst)

    @property
    def launches(self) -> int:
        return self._ll.plan_launches(self._h)

    def close(self):
        if self._h is not None:
            self._ll.plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def clenshaw_device(c_ptr: int, N: int, x_ptr: int, y_ptr: int, P: int, stream: int = 0):
    lib = lowlevel().lib
    L.check(lib, lib.afb_clenshaw_cheb_device(C.c_void_p(c_ptr), N, C.c_void_p(x_ptr), C.c_void_p(y_ptr), P, C.c_void_p(stream)))


def sample_device(cdf_ptr: int, N: int, a: float, b: float, seed: int, offset: int, out_ptr: int, P: int, stream: int = 0):
    lib = lowlevel().lib
    L.check(lib, lib.afb_sample_cheb_device(C.c_void_p(cdf_ptr), N, a, b, seed, offset, C.c_void_p(out_ptr), P, C.c_void_p(stream)))


def bisect_device(c_ptr: int, N: int, u_ptr: int, out_ptr: int, P: int, a: float = -1.0, b: float = 1.0, numits: int = 47, stream: int = 0):
    lib = lowlevel().lib
    L.check(lib, lib.afb_cheb_bisect_device(C.c_void_p(c_ptr), N, C.c_void_p(u_ptr), C.c_void_p(out_ptr), P, numits, a, b, C.c_void_p(stream)))


def philox_device(seed: int, offset: int, out_ptr: int, P: int, stream: int = 0):
    lib = lowlevel().lib
    L.check(lib, lib.afb_philox_uniform_device(seed, offset, C.c_void_p(out_ptr), P, C.c_void_p(stream)))
