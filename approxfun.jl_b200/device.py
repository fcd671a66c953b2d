"""Device-resident Fun algebra: `values -> point-wise op -> transform` chains and the adaptive constructor ladder without a
PCIe round trip per step (reference: the ladder of ext/ApproxFunDualNumbersExt.jl:96-111 /
UPSTREAM ApproxFunBase `default_Fun`, `f*g`, `sin(f)`, `abs`).

`DeviceArray` wraps an opaque `afb_buf` (C ABI) and overloads + - * / and sin, cos, exp, abs, sqrt, log, tanh, so the SAME
closure a user passes to `Fun(f, S)` -- e.g. `lambda x: sin(x**2)` with this module's `sin` -- runs as point-wise CUDA kernels
on the device-resident grid.  `DeviceFun` keeps the coefficients in a buffer; `f * g`, `f + g`, `compose(op, f)` stay on the
device; only `.coefficients` downloads."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .api import Chebyshev, Fun, lowlevel

OP = {"ident": 0, "sin": 1, "cos": 2, "exp": 3, "abs": 4, "sqrt": 5, "square": 6, "recip": 7, "log": 8, "tanh": 9}
BOP = {"add": 0, "sub": 1, "mul": 2, "div": 3}

# bytes moved across PCIe by this module (uploads, downloads), for tests and bench.py
traffic = {"h2d": 0, "d2h": 0}


def _lib():
    return lowlevel().lib


class DeviceArray:
    """n doubles in device memory."""

    def __init__(self, n: int, handle=None):
        self.n = int(n)
        if handle is None:
            handle = C.c_void_p()
            L.check(_lib(), _lib().afb_buf_create(C.byref(handle), self.n))
        self._h = handle

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            try:
                _lib().afb_buf_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- host <-> device ---------------------------------------------------------------------------------------
    @staticmethod
    def from_host(a) -> "DeviceArray":
        a = np.ascontiguousarray(a, dtype=np.float64)
        d = DeviceArray(a.size)
        L.check(_lib(), _lib().afb_buf_upload(d._h, C.c_void_p(a.ctypes.data), a.size, 0))
        traffic["h2d"] += a.nbytes
        return d

    def to_host(self, n: int | None = None) -> np.ndarray:
        n = self.n if n is None else int(n)
        out = np.empty(n)
        L.check(_lib(), _lib().afb_buf_download(self._h, C.c_void_p(out.ctypes.data), n, 0))
        traffic["d2h"] += out.nbytes
        return out

    @property
    def ptr(self) -> int:
        p = C.c_void_p()
        L.check(_lib(), _lib().afb_buf_ptr(self._h, C.byref(p)))
        return int(p.value or 0)

    # ---- point-wise kernels ------------------------------------------------------------------------------------------
    def _unary(self, op: str, alpha: float = 1.0, beta: float = 0.0) -> "DeviceArray":
        out = DeviceArray(self.n)
        L.check(_lib(), _lib().afb_buf_unary(OP[op], self._h, out._h, self.n, float(alpha), float(beta)))
        return out

    def _binary(self, op: str, other) -> "DeviceArray":
        if not isinstance(other, DeviceArray):
            raise TypeError(other)
        if other.n != self.n:
            raise L.DimensionMismatch(L.AFB_BAD_SIZE, "DeviceArray: lengths differ")
        out = DeviceArray(self.n)
        L.check(_lib(), _lib().afb_buf_binary(BOP[op], self._h, other._h, out._h, self.n))
        return out

    def __add__(self, o):
        return self._binary("add", o) if isinstance(o, DeviceArray) else self._unary("ident", 1.0, float(o))

    __radd__ = __add__

    def __sub__(self, o):
        return self._binary("sub", o) if isinstance(o, DeviceArray) else self._unary("ident", 1.0, -float(o))

    def __rsub__(self, o):
        return self._unary("ident", -1.0, float(o))

    def __mul__(self, o):
        return self._binary("mul", o) if isinstance(o, DeviceArray) else self._unary("ident", float(o), 0.0)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._binary("div", o) if isinstance(o, DeviceArray) else self._unary("ident", 1.0 / float(o), 0.0)

    def __rtruediv__(self, o):
        r = self._unary("recip")
        return r if float(o) == 1.0 else r._unary("ident", float(o), 0.0)

    def __neg__(self):
        return self._unary("ident", -1.0, 0.0)

    def __pow__(self, k):
        if k == 2:
            return self._unary("square")
        if k == 1:
            return self._unary("ident")
        if isinstance(k, (int, np.integer)) and k > 2:
            r = self._unary("square")
            for _ in range(int(k) - 2):
                r = r * self
            return r
        if k == 0.5:
            return self._unary("sqrt")
        raise NotImplementedError("DeviceArray ** k for integer k >= 1 or k = 0.5")

    def __abs__(self):
        return self._unary("abs")

    def pad(self, n_in: int, n_out: int) -> "DeviceArray":
        out = DeviceArray(n_out)
        L.check(_lib(), _lib().afb_buf_pad(self._h, int(n_in), out._h, int(n_out)))
        return out


def _elementwise(name):
    def fn(x):
        if isinstance(x, DeviceArray):
            return x._unary(name)
        return getattr(np, {"abs": "abs"}.get(name, name))(x)
    fn.__name__ = name
    return fn


sin, cos, exp, sqrt, log, tanh = (_elementwise(n) for n in ("sin", "cos", "exp", "sqrt", "log", "tanh"))
absolute = _elementwise("abs")


def chebpoints_device(space: Chebyshev, n: int) -> DeviceArray:
    """`points(space, n)` produced on the device (never leaves it)."""
    d = DeviceArray(n)
    L.check(_lib(), _lib().afb_buf_chebpoints(d._h, n, space.a, space.b))
    return d


_plans: dict = {}


def _plan(kind: int, n: int):
    key = (kind, n)
    p = _plans.get(key)
    if p is None:
        if len(_plans) >= 64:
            _plans.pop(next(iter(_plans)))
        p = _plans[key] = lowlevel().plan_create(kind, n, 0, 1, n)
    return p


def transform_device(vals: DeviceArray, n: int, inverse: bool = False) -> DeviceArray:
    out = DeviceArray(n)
    L.check(_lib(), _lib().afb_plan_execute_buf(_plan(L.CHEB1_INV if inverse else L.CHEB1_FWD, n), vals._h, out._h))
    return out


def _tail(c: DeviceArray, n: int, rel: float):
    mx, tl, ln = C.c_double(), C.c_double(), C.c_int64()
    L.check(_lib(), _lib().afb_buf_tail(c._h, n, 8, rel, C.byref(mx), C.byref(tl), C.byref(ln)))
    traffic["d2h"] += 24
    return mx.value, tl.value, int(ln.value)


class DeviceFun:
    """A Chebyshev Fun whose coefficients live on the device."""

    def __init__(self, space: Chebyshev, coeffs: DeviceArray, n: int):
        self.space, self.c, self.n = space, coeffs, int(n)

    # ---- constructors ---------------------------------------------------------------------------------------------
    @staticmethod
    def from_function(f, space: Chebyshev, n: int | None = None) -> "DeviceFun":
        """`Fun(f, space, n)` or the adaptive `Fun(f, space)`; `f` maps a DeviceArray of points to a DeviceArray of values
        (built from + - * / ** and this module's sin, cos, exp, ...).  Ladder, tail test and chop as the host constructor
        (`api._adaptive`): n = 2^4 .. 2^20, tolerance 20 eps max|c| on the last 8 coefficients, then 2^21."""
        if n is not None:
            return DeviceFun(space, transform_device(f(chebpoints_device(space, n)), n), n)
        tol = 20 * np.finfo(np.float64).eps
        for logn in range(4, 21):
            m = 2 ** logn
            c = transform_device(f(chebpoints_device(space, m)), m)
            mx, tl, ln = _tail(c, m, tol)
            if mx == 0.0:
                return DeviceFun(space, DeviceArray.from_host(np.zeros(1)), 1)
            if tl < tol * mx:
                return DeviceFun(space, c, max(ln, 1))
        m = 2 ** 21
        return DeviceFun(space, transform_device(f(chebpoints_device(space, m)), m), m)

    @staticmethod
    def from_fun(f: Fun) -> "DeviceFun":
        return DeviceFun(f.space, DeviceArray.from_host(f.coefficients), len(f.coefficients))

    # ---- device-side pieces ------------------------------------------------------------------------------------------
    def values_device(self, m: int | None = None) -> DeviceArray:
        """`values(pad(f, m))`: the function on the m-point grid, on the device."""
        m = self.n if m is None else int(m)
        return transform_device(self.c.pad(min(self.n, m), m), m, inverse=True)

    def compose(self, op, other: "DeviceFun | None" = None) -> "DeviceFun":
        """Adaptive `Fun(x -> op(f(x)[, g(x)]), space)` with f (and g) evaluated from their coefficients on every rung."""
        tol = 20 * np.finfo(np.float64).eps
        start = max(4, int(np.ceil(np.log2(max(self.n, other.n if other else 1, 16)))))
        c, m = None, 0
        for logn in range(start, 22):
            m = 2 ** logn
            v = op(self.values_device(m)) if other is None else op(self.values_device(m), other.values_device(m))
            c = transform_device(v, m)
            if logn == 21:
                break
            mx, tl, ln = _tail(c, m, tol)
            if mx == 0.0:
                return DeviceFun(self.space, DeviceArray.from_host(np.zeros(1)), 1)
            if tl < tol * mx:
                return DeviceFun(self.space, c, max(ln, 1))
        return DeviceFun(self.space, c, m)

    def _product(self, other: "DeviceFun") -> "DeviceFun":
        """`f * g`: exact on n = ncoefficients(f) + ncoefficients(g) - 1 points (values multiply, degrees add), then chop."""
        m = max(self.n + other.n - 1, 1)
        c = transform_device(self.values_device(m) * other.values_device(m), m)
        mx, _, ln = _tail(c, m, 20 * np.finfo(np.float64).eps)
        return DeviceFun(self.space, c, max(ln, 1) if mx > 0 else 1)

    def _linear(self, other: "DeviceFun", sign: float) -> "DeviceFun":
        m = max(self.n, other.n)
        a, b = self.c.pad(self.n, m), other.c.pad(other.n, m)
        return DeviceFun(self.space, a + b if sign > 0 else a - b, m)

    def __mul__(self, o):
        return self._product(o) if isinstance(o, DeviceFun) else DeviceFun(self.space, self.c * float(o), self.n)

    __rmul__ = __mul__

    def __add__(self, o):
        if isinstance(o, DeviceFun):
            return self._linear(o, 1.0)
        one = np.zeros(self.n)
        one[0] = float(o)
        return DeviceFun(self.space, self.c.pad(self.n, self.n) + DeviceArray.from_host(one), self.n)

    def __sub__(self, o):
        return self._linear(o, -1.0) if isinstance(o, DeviceFun) else self + (-float(o))

    def __call__(self, x):
        """Evaluate at host points (uploads x, downloads f(x)); `evaluate_device` keeps both on the device."""
        xa = np.atleast_1d(np.asarray(x, dtype=np.float64))
        y = self.evaluate_device(DeviceArray.from_host(self.space.tocanonical(xa.ravel()))).to_host().reshape(xa.shape)
        return float(y[0]) if np.ndim(x) == 0 else y

    def evaluate_device(self, xc: DeviceArray) -> DeviceArray:
        y = DeviceArray(xc.n)
        L.check(_lib(), _lib().afb_buf_clenshaw(self.c._h, self.n, xc._h, y._h, xc.n))
        return y

    @property
    def coefficients(self) -> np.ndarray:
        return self.c.to_host(self.n)

    def to_fun(self) -> Fun:
        return Fun(self.space, self.coefficients)


def sin_fun(f: DeviceFun) -> DeviceFun:
    return f.compose(sin)


def exp_fun(f: DeviceFun) -> DeviceFun:
    return f.compose(exp)


def abs_values(f: DeviceFun, m: int) -> DeviceArray:
    """|f| on the m-point grid (`abs(f)` proper is piecewise: `api.abs_fun`)."""
    return absolute(f.values_device(m))
