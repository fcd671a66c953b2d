"""CPU oracle for the ApproxFun transform / evaluate / sample hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``approxfun.jl_b200``) never does and
fails loudly when the CUDA library is missing.

It restates, in NumPy / SciPy (pocketfft), the conventions of the reference
(ApproxFun.jl v0.13.28 at /root/reference).  Citations are ``path:line`` under
/root/reference; UPSTREAM marks arithmetic that lives in packages that are *not*
vendored there (FastTransforms 0.13-0.17, FFTW.jl 1.x, ApproxFunBase 0.8/0.9,
ApproxFunOrthogonalPolynomials 0.5/0.6, ApproxFunFourier 0.3 -- Project.toml:27-51); for those
the published algorithm is restated and parity is anchored on the reference's own call sites,
doctests and golden values (tests/test_oracle_goldens.py).

Parity status: PINNED to the in-tree golden values / known-answer tests that exist for this
path: test/SamplingTest.jl:9 (0.2004758624973169), test/SamplingTest.jl:12,
test/ReadmeTest.jl:23, docs/src/usage/spaces.md:94-98 and :121-125 (Fourier / Laurent
ordering), test/NumberTypeTest.jl:43-44 via src/Extras/fftGeneric.jl:47-64 (restated
here verbatim as ``fourier_transform_generic``), test/ExtrasTest.jl:61 (round trip).
UNPINNED (no stored vectors exist in the reference): element-wise coefficient vectors at
n = 4096 / 2^20 (pocketfft stands in for FFTW; both are O(eps log n) accurate), the last
bit of ``points`` values, whether Julia's ``muladd`` fused (we use a true fma in the C
oracle; this NumPy file uses separate multiply/add and is therefore compared with a tolerance).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.fft as _sfft

# --------------------------------------------------------------------------------------
# a1  points
# --------------------------------------------------------------------------------------


def sinpi(x):
    """sin(pi*x) with exact argument reduction (Julia ``sinpi``)."""
    x = np.asarray(x, dtype=np.float64)
    # reduce to r in [-0.5, 0.5]: sin(pi x) = (-1)^k sin(pi r), x = k + r
    k = np.rint(x)
    r = x - k
    s = np.sin(np.pi * r)
    # exact values where representable
    s = np.where(r == 0.0, 0.0, s)
    s = np.where(np.abs(r) == 0.5, np.sign(r), s)
    return np.where(np.mod(k, 2.0) == 0.0, s, -s)


def chebpoints_canonical(n: int) -> np.ndarray:
    """First-kind Chebyshev points on [-1,1], DESCENDING (NEWS.md:92).

    UPSTREAM FastTransforms.chebyshevpoints(Float64, n, Val(1)):
    ``x_k = sinpi((n-2k-1)/(2n))``, k = 0..n-1  (= cos(pi (2k+1)/(2n))).
    In-tree users: docs/src/faq.md:15, ext/ApproxFunDualNumbersExt.jl:80.
    """
    k = np.arange(n, dtype=np.float64)
    return sinpi((n - 2.0 * k - 1.0) / (2.0 * n))


def fromcanonical(a: float, b: float, x):
    """UPSTREAM ApproxFunBase ``fromcanonical(d::Segment,x) = mean(d) + complexlength(d)x/2``.

    Inverse of M(x) = (2x-b-a)/(b-a) (docs/src/usage/spaces.md:137-139).
    """
    return (a + b) / 2.0 + (b - a) * np.asarray(x, dtype=np.float64) / 2.0


def tocanonical(a: float, b: float, x):
    """M(x) = (2x-b-a)/(b-a)  (docs/src/usage/spaces.md:137-139)."""
    x = np.asarray(x, dtype=np.float64)
    return (2.0 * x - a - b) / (b - a)


def chebpoints(n: int, a: float = -1.0, b: float = 1.0) -> np.ndarray:
    """``points(Chebyshev(a..b), n)``."""
    return fromcanonical(a, b, chebpoints_canonical(n))


def fourierpoints(n: int, a: float = 0.0, b: float = 2.0 * math.pi) -> np.ndarray:
    """``points(Fourier()/Laurent(), n)``: theta_j = 2 pi j / n mapped affinely to [a,b).

    UPSTREAM ApproxFunFourier (canonical periodic grid); plans are domain free.
    """
    j = np.arange(n, dtype=np.float64)
    return a + (b - a) * j / n


# --------------------------------------------------------------------------------------
# a2 / a3  Chebyshev transforms (first-kind points):  FFTW REDFT10 / REDFT01
# --------------------------------------------------------------------------------------


def cheb_transform(v: np.ndarray, axis: int = 0, workers: int | None = None) -> np.ndarray:
    """``plan_transform(Chebyshev(), n) * v``.

    UPSTREAM FastTransforms.plan_chebyshevtransform(v, Val(1)) (boundary seen in-tree at
    ext/ApproxFunDualNumbersExt.jl:10-12,44,49): FFTW REDFT10 ``Y_k = 2 sum_j v_j cos(pi k (j+1/2)/n)``
    then ``Y[1] /= 2; Y ./= n``  =>  c_0 = mean(v), c_k = (2/n) sum_j v_j cos(pi k (j+1/2)/n).
    scipy ``dct(type=2)`` (pocketfft) is the same unnormalised REDFT10.
    """
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[axis]
    if n == 0:
        return v.copy()
    y = _sfft.dct(v, type=2, axis=axis, workers=workers)
    y /= n
    sl = [slice(None)] * y.ndim
    sl[axis] = 0
    y[tuple(sl)] /= 2.0
    return y


def cheb_itransform(c: np.ndarray, axis: int = 0, workers: int | None = None) -> np.ndarray:
    """``plan_itransform(Chebyshev(), n) * c``  (``values(f)``, src/Plot/Plot.jl:20, test/ExtrasTest.jl:61).

    UPSTREAM FastTransforms.plan_ichebyshevtransform(c, Val(1)): ``c[1] *= 2``; FFTW REDFT01
    ``Y_j = X_0 + 2 sum_{k>=1} X_k cos(pi k (j+1/2)/n)``; ``Y ./ 2``  =>  v_j = sum_k c_k cos(pi k (j+1/2)/n).
    scipy ``dct(type=3)`` is the same unnormalised REDFT01.
    """
    c = np.array(c, dtype=np.float64, copy=True)
    n = c.shape[axis]
    if n == 0:
        return c
    sl = [slice(None)] * c.ndim
    sl[axis] = 0
    c[tuple(sl)] *= 2.0
    y = _sfft.dct(c, type=3, axis=axis, workers=workers)
    y /= 2.0
    return y


def cheb_transform_direct(v: np.ndarray) -> np.ndarray:
    """O(n^2) long-double evaluation of the a2 definition (ground truth for small n)."""
    v = np.asarray(v, dtype=np.longdouble)
    n = v.shape[0]
    j = np.arange(n, dtype=np.longdouble)
    k = np.arange(n, dtype=np.longdouble)
    # cos(pi k (2j+1) / (2n)) with integer argument reduction mod 4n
    m = (np.outer(np.arange(n), 2 * np.arange(n) + 1) % (4 * n)).astype(np.longdouble)
    C = np.cos(np.longdouble(np.pi) * m / (2 * np.longdouble(n)))
    del j, k
    c = (C @ v) * (np.longdouble(2) / n)
    c[0] /= 2
    return c.astype(np.float64)


def cheb_itransform_direct(c: np.ndarray) -> np.ndarray:
    c = np.asarray(c, dtype=np.longdouble)
    n = c.shape[0]
    m = (np.outer(2 * np.arange(n) + 1, np.arange(n)) % (4 * n)).astype(np.longdouble)
    C = np.cos(np.longdouble(np.pi) * m / (2 * np.longdouble(n)))
    return (C @ c).astype(np.float64)


# --------------------------------------------------------------------------------------
# a4  Fourier  (R2HC/HC2R + shuffles upstream; in-tree BigFloat twin fftGeneric.jl:47-64)
# --------------------------------------------------------------------------------------


def _interlace(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ApproxFunBase.interlace for vectors with len(a) in {len(b), len(b)+1}."""
    out = np.empty(a.shape[0] + b.shape[0], dtype=np.result_type(a, b))
    out[0::2] = a
    out[1::2] = b
    return out


def fourier_transform_generic(x: np.ndarray) -> np.ndarray:
    """Line-by-line restatement of src/Extras/fftGeneric.jl:47-54 (1-D only).

    l = length(x); n = div(l+1,2); v = fft(x); v *= 2/l; v[1] /= 2;
    odd  l: interlace(real(v[1:n]), -imag(v[2:n]))
    even l: [interlace(real(v[1:n]), -imag(v[2:n])); -real(v[n+1])/2]
    """
    x = np.asarray(x, dtype=np.float64)
    l = x.shape[0]
    if l == 0:
        return x.copy()
    n = (l + 1) // 2
    v = np.fft.fft(x)
    v = v * (2.0 / l)
    v[0] /= 2
    if l % 2 == 1:
        return _interlace(v.real[0:n], -v.imag[1:n])
    return np.concatenate([_interlace(v.real[0:n], -v.imag[1:n]), [-v.real[n] / 2]])


def fourier_itransform_generic(x: np.ndarray) -> np.ndarray:
    """Line-by-line restatement of src/Extras/fftGeneric.jl:56-64 (1-D only)."""
    x = np.asarray(x, dtype=np.float64)
    l = x.shape[0]
    if l == 0:
        return x.copy()
    if l == 1:
        return x.copy()
    # Julia 1-based slices -> 0-based
    if l % 2 == 1:
        re = np.concatenate([[2 * x[0]], x[2::2], x[2::2][::-1]])
        im = np.concatenate([[0.0], -x[1::2], x[1::2][::-1]])
    else:
        # complex([2x[1]; x[3:2:end-1]; -2x[end]; x[end-1:-2:3]], [0; -x[2:2:end]; x[end-2:-2:2]])
        xr = x[2:l - 1:2]          # x[3:2:end-1]
        xi = x[1::2]               # x[2:2:end]   (length l/2, includes x[end])
        re = np.concatenate([[2 * x[0]], xr, [-2 * x[-1]], xr[::-1]])
        im = np.concatenate([[0.0], -xi, x[1:l - 2:2][::-1]])  # x[end-2:-2:2]
    v = (re + 1j * im) * (l / 2.0)
    return np.real(np.fft.ifft(v))


def fourier_transform(v: np.ndarray, axis: int = 0, workers: int | None = None) -> np.ndarray:
    """Batched ``plan_transform(Fourier(), n) * v`` (same numbers as the generic form above).

    out = [a0, b1, a1, b2, a2, ...], a0 = F0/n, a_k = (2/n) Re F_k, b_k = -(2/n) Im F_k,
    F_k = sum_j v_j e^{-2 pi i jk/n}; even n: last slot = -(1/n) Re F_{n/2}
    (fftGeneric.jl:53; ordering doctest docs/src/usage/spaces.md:94-98).
    """
    v = np.moveaxis(np.asarray(v, dtype=np.float64), axis, 0)
    l = v.shape[0]
    if l == 0:
        return np.moveaxis(v.copy(), 0, axis)
    n = (l + 1) // 2
    F = _sfft.rfft(v, axis=0, workers=workers)
    out = np.empty_like(v)
    out[0] = F[0].real / l
    out[2:2 * n:2] = F[1:n].real * (2.0 / l)       # a_k at index 2k
    out[1:2 * n - 1:2] = -F[1:n].imag * (2.0 / l)  # b_k at index 2k-1
    if l % 2 == 0:
        out[l - 1] = -F[n].real / l
    return np.moveaxis(out, 0, axis)


def fourier_itransform(c: np.ndarray, axis: int = 0, workers: int | None = None) -> np.ndarray:
    """Batched ``plan_itransform(Fourier(), n) * c`` (fftGeneric.jl:56-64)."""
    c = np.moveaxis(np.asarray(c, dtype=np.float64), axis, 0)
    l = c.shape[0]
    if l == 0:
        return np.moveaxis(c.copy(), 0, axis)
    n = (l + 1) // 2
    F = np.zeros((l // 2 + 1,) + c.shape[1:], dtype=np.complex128)
    F[0] = c[0] * l
    # a_k = c[2k], b_k = c[2k-1]:  F_k = (l/2)(a_k - i b_k)
    F[1:n] = (c[2:2 * n:2] - 1j * c[1:2 * n - 1:2]) * (l / 2.0)
    if l % 2 == 0:
        F[n] = -c[l - 1] * l
    out = _sfft.irfft(F, n=l, axis=0, workers=workers)
    return np.moveaxis(out, 0, axis)


# --------------------------------------------------------------------------------------
# a5  Laurent  (c2c FFT + reorder; doctest docs/src/usage/spaces.md:113-125)
# --------------------------------------------------------------------------------------


def laurent_index_map(n: int) -> np.ndarray:
    """m[i] = FFT bin feeding output slot i:  [0, -1, 1, -2, 2, ...] mod n."""
    i = np.arange(n)
    k = np.where(i % 2 == 0, i // 2, -((i + 1) // 2))
    return np.mod(k, n)


def laurent_transform(v: np.ndarray, axis: int = 0, workers: int | None = None) -> np.ndarray:
    """``plan_transform(Laurent(), n) * v``: f_k = F_k/n ordered [f0, f-1, f1, f-2, f2, ...]."""
    v = np.moveaxis(np.asarray(v, dtype=np.complex128), axis, 0)
    n = v.shape[0]
    if n == 0:
        return np.moveaxis(v.copy(), 0, axis)
    F = _sfft.fft(v, axis=0, workers=workers) / n
    return np.moveaxis(F[laurent_index_map(n)], 0, axis)


def laurent_itransform(c: np.ndarray, axis: int = 0, workers: int | None = None) -> np.ndarray:
    c = np.moveaxis(np.asarray(c, dtype=np.complex128), axis, 0)
    n = c.shape[0]
    if n == 0:
        return np.moveaxis(c.copy(), 0, axis)
    F = np.empty_like(c)
    F[laurent_index_map(n)] = c
    return np.moveaxis(_sfft.ifft(F * n, axis=0, workers=workers), 0, axis)


# --------------------------------------------------------------------------------------
# a7-a9  Clenshaw (UPSTREAM AFOP; in-tree call sites src/Extras/sample.jl:47,56,68,80,94)
# --------------------------------------------------------------------------------------


def clenshaw(c: np.ndarray, x) -> np.ndarray:
    """Chebyshev Clenshaw, same recurrence/order as the reference (non-fused arithmetic).

    x2=2x; b1=b2=0; for k=N:-1:2: (b2,b1)=(b1, muladd(x2,b1,c[k]-b2)); muladd(x2/2,b1,c[1]-b2).
    N == 0 -> 0, N == 1 -> c[1].
    """
    c = np.asarray(c, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    N = c.shape[0]
    if N == 0:
        return np.zeros_like(x)
    if N == 1:
        return np.full_like(x, c[0])
    x2 = 2.0 * x
    b1 = np.zeros_like(x)
    b2 = np.zeros_like(x)
    for k in range(N - 1, 0, -1):
        b2, b1 = b1, x2 * b1 + (c[k] - b2)
    return (x2 / 2.0) * b1 + (c[0] - b2)


def clenshaw_cols(C: np.ndarray, x: np.ndarray) -> np.ndarray:
    """``clenshaw(c::Matrix, x::Vector, plan)``: column j of C (N x n) at x[j] (sample.jl:80,94)."""
    C = np.asarray(C, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    N, n = C.shape
    assert n == x.shape[0]
    if N == 0:
        return np.zeros_like(x)
    if N == 1:
        return C[0].copy()
    x2 = 2.0 * x
    b1 = np.zeros_like(x)
    b2 = np.zeros_like(x)
    for k in range(N - 1, 0, -1):
        b2, b1 = b1, x2 * b1 + (C[k] - b2)
    return (x2 / 2.0) * b1 + (C[0] - b2)


def clenshaw_multi(C: np.ndarray, x: np.ndarray) -> np.ndarray:
    """``evaluate.(f.A, ry')`` (sample.jl:208): every column of C (N x M) at every x -> (M, P)."""
    C = np.asarray(C, dtype=np.float64)
    return np.stack([clenshaw(C[:, m], x) for m in range(C.shape[1])], axis=0)


def clenshaw_mp(c, x, dps: int = 50):
    """mpmath ground truth  sum_k c_k T_k(x)  (error budgets)."""
    import mpmath as mp

    mp.mp.dps = dps
    xx = mp.mpf(float(x))
    b1 = mp.mpf(0)
    b2 = mp.mpf(0)
    for k in range(len(c) - 1, 0, -1):
        b2, b1 = b1, 2 * xx * b1 + (mp.mpf(float(c[k])) - b2)
    return xx * b1 + (mp.mpf(float(c[0])) - b2) if len(c) else mp.mpf(0)


# --------------------------------------------------------------------------------------
# a12 / a13  normalised cumulative sums
# --------------------------------------------------------------------------------------


def ultraconversion(v: np.ndarray) -> np.ndarray:
    """UPSTREAM AFOP ``ultraconversion!`` (T -> U coefficients), per column for matrices.

    n<=1 nothing; n==2: v2/=2; else v1 -= v3/2; v_j=(v_j - v_{j+2})/2 (j=2..n-2); v_{n-1}/=2; v_n/=2.
    In-tree call site src/Extras/sample.jl:171.
    """
    v = np.array(v, dtype=np.float64, copy=True)
    n = v.shape[0]
    if n <= 1:
        return v
    if n == 2:
        v[1] /= 2
        return v
    v[0] -= v[2] / 2
    v[1:n - 2] = (v[1:n - 2] - v[3:n]) / 2
    v[n - 2] /= 2
    v[n - 1] /= 2
    return v


def ultraint(v: np.ndarray, grow: bool) -> np.ndarray:
    """UPSTREAM AFOP ``ultraint!``: v_k = v_{k-1}/(k-1), v_1 = 0.

    The Vector method resizes by +1 (``grow=True``); the Matrix method does not and drops the
    top term (``grow=False``).  In-tree call site src/Extras/sample.jl:172.
    """
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[0]
    if grow:
        out = np.zeros((n + 1,) + v.shape[1:], dtype=np.float64)
        k = np.arange(1, n + 1, dtype=np.float64).reshape((-1,) + (1,) * (v.ndim - 1))
        out[1:] = v / k
        return out
    out = np.zeros_like(v)
    if n > 1:
        k = np.arange(1, n, dtype=np.float64).reshape((-1,) + (1,) * (v.ndim - 1))
        out[1:] = v[:-1] / k
    return out


def subtract_zeroatleft(f: np.ndarray) -> np.ndarray:
    """src/Extras/sample.jl:121-135: f[1] += sum_{k>=2} (-1)^k f[k] (sequential order)."""
    f = np.array(f, dtype=np.float64, copy=True)
    for k in range(2, f.shape[0] + 1):
        f[0] = f[0] + ((-1.0) ** k) * f[k - 1]
    return f


def multiply_oneatright(f: np.ndarray) -> np.ndarray:
    """src/Extras/sample.jl:137-168: val = sum_k f[k] (sequential); f *= 1/val."""
    f = np.array(f, dtype=np.float64, copy=True)
    val = np.zeros(f.shape[1:], dtype=np.float64)
    for k in range(f.shape[0]):
        val = val + f[k]
    val = 1.0 / val
    return f * val


def chebnormalizedcumsum(f: np.ndarray, grow: bool | None = None) -> np.ndarray:
    """src/Extras/sample.jl:170-175.  Vector input grows by one; matrix input does not."""
    f = np.asarray(f, dtype=np.float64)
    if grow is None:
        grow = f.ndim == 1
    return multiply_oneatright(subtract_zeroatleft(ultraint(ultraconversion(f), grow)))


def cheb_cumsum(c: np.ndarray, a: float = -1.0, b: float = 1.0) -> np.ndarray:
    """UPSTREAM ``cumsum(f::Fun{<:Chebyshev})`` = integrate(f) - first(integrate(f)).

    integrate = ((b-a)/2) * ultraint!(ultraconversion!(copy(c))); first(g) = g(a) = sum_k (-1)^k g_k.
    In-tree call sites src/Extras/sample.jl:115, test/SamplingTest.jl:11.
    """
    g = ultraint(ultraconversion(c), True) * ((b - a) / 2.0)
    sgn = np.where(np.arange(g.shape[0]) % 2 == 0, 1.0, -1.0)
    g[0] -= float(np.sum(sgn * g))
    return g


def normalizedcumsum(c: np.ndarray, a: float = -1.0, b: float = 1.0) -> np.ndarray:
    """src/Extras/sample.jl:114-119: cf = cumsum(f); cf/last(cf)  (last(cf) = cf(b) = sum of coefficients)."""
    g = cheb_cumsum(c, a, b)
    return g / float(np.sum(g))


def cheb_definite_integrals(N: int) -> np.ndarray:
    """int_{-1}^{1} T_k(x) dx = 2/(1-k^2) for even k, 0 for odd k (used by ``sum``)."""
    k = np.arange(N, dtype=np.float64)
    w = np.zeros(N)
    ev = (np.arange(N) % 2) == 0
    w[ev] = 2.0 / (1.0 - k[ev] ** 2)
    return w


# --------------------------------------------------------------------------------------
# a14  bisection
# --------------------------------------------------------------------------------------


def chebbisectioninv(c: np.ndarray, u, numits: int = 47) -> np.ndarray:
    """src/Extras/sample.jl:41-52 (scalar) and :58-75 (vector): 47 Clenshaw bisections on [-1,1]."""
    u = np.atleast_1d(np.asarray(u, dtype=np.float64))
    a = -np.ones_like(u)
    b = np.ones_like(u)
    for _ in range(numits):
        m = 0.5 * (a + b)
        vals = clenshaw(c, m)
        le = vals <= u
        a = np.where(le, m, a)
        b = np.where(le, b, m)
    return 0.5 * (a + b)


def chebbisectioninv_cols(C: np.ndarray, u: np.ndarray, numits: int = 47) -> np.ndarray:
    """src/Extras/sample.jl:82-101: matrix variant, column j against u[j]."""
    C = np.asarray(C, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    assert C.shape[1] == u.shape[0]  # sample.jl:85
    a = -np.ones_like(u)
    b = np.ones_like(u)
    for _ in range(numits):
        m = 0.5 * (a + b)
        vals = clenshaw_cols(C, m)
        le = vals <= u
        a = np.where(le, m, a)
        b = np.where(le, b, m)
    return 0.5 * (a + b)


def bisectioninv(cf: np.ndarray, a: float, b: float, u) -> np.ndarray:
    """src/Extras/sample.jl:103-109: fromcanonical.(space, chebbisectioninv(coefficients(cf), u))."""
    return fromcanonical(a, b, chebbisectioninv(cf, u))


def sample_from_uniforms(pdf_c: np.ndarray, a: float, b: float, u: np.ndarray) -> np.ndarray:
    """src/Extras/sample.jl:182-184 with caller-supplied uniforms in place of ``rand(n)``."""
    return bisectioninv(normalizedcumsum(pdf_c, a, b), a, b, u)


# --------------------------------------------------------------------------------------
# a10  tensor-product transform
# --------------------------------------------------------------------------------------


def cheb2_transform(V: np.ndarray, workers: int | None = None) -> np.ndarray:
    """UPSTREAM ApproxFunBase ``transform!(::TensorSpace, V)``: factor-1 plan on every column, then
    factor-2 plan on every row: C = T_x V T_y^T  (docs/src/internals/multivariate.md:30-44).

    V is indexed V[k, j] = f(x_k, y_j); with NumPy the "column" direction is axis 0.
    """
    return cheb_transform(cheb_transform(V, axis=0, workers=workers), axis=1, workers=workers)


def cheb2_itransform(C: np.ndarray, workers: int | None = None) -> np.ndarray:
    return cheb_itransform(cheb_itransform(C, axis=0, workers=workers), axis=1, workers=workers)


def tensor_interlace(C: np.ndarray) -> np.ndarray:
    """Coefficient matrix -> flat TensorSpace ``Fun`` coefficients, ordered by total degree
    T0T0, T0T1, T1T0, T0T2, T1T1, T2T0, ... (docs/src/usage/constructors.md:98-116)."""
    n, m = C.shape
    out = []
    for d in range(n + m - 1):
        for i in range(0, d + 1):
            j = d - i
            if i < n and j < m:
                out.append(C[i, j])
            else:
                out.append(0.0)
    return np.array(out)


def tensor_deinterlace(c: np.ndarray, n: int | None = None) -> np.ndarray:
    """Inverse of :func:`tensor_interlace` for a triangular-length vector: (n+1) x (n+1) matrix C[i, j]
    (x-degree i, y-degree j), zero for i + j > n."""
    N = len(c)
    n = padua_degree(N) if n is None else n
    C = np.zeros((n + 1, n + 1), dtype=np.asarray(c).dtype)
    l = 0
    for d in range(n + 1):
        for i in range(d + 1):
            if l < N:
                C[i, d - i] = c[l]
            l += 1
    return C


# --------------------------------------------------------------------------------------
# 8(f) rank 1: Padua points / transform  (`Fun(f, Chebyshev()^2, N)`, NEWS.md:85)
#
# UPSTREAM FastTransforms `paduapoints`, `plan_paduatransform!`, `plan_ipaduatransform!` (not in
# /root/reference).  PARITY UNPINNED for the ORDER of the points: it is restated from the published
# algorithm (Caliari, De Marchi, Vianello 2008; the upstream loop nest k = n:-1:0, j = NN+delta:-1:1)
# and no in-tree vector fixes it.  What IS pinned in-tree: the coefficient order
# T0T0, T0(x)T1(y), T1(x)T0(y), ... (docs/src/usage/constructors.md:98-116) and the functional
# identities (`Fun(f, Chebyshev()^2)(x, y) == f(x, y)`, examples/PDE_Helmholtz.jl:33-36), which any
# consistent (points, transform) pair with that coefficient order satisfies.
# --------------------------------------------------------------------------------------
def padua_degree(N: int) -> int:
    """n with (n+1)(n+2)/2 >= N: `Int(cld(-3 + sqrt(1 + 8N), 2))`."""
    n = int(math.ceil((-3.0 + math.sqrt(1.0 + 8.0 * N)) / 2.0 - 1e-12))
    while (n + 1) * (n + 2) // 2 < N:
        n += 1
    while n > 0 and n * (n + 1) // 2 >= N:
        n -= 1
    return n


def padua_grid_index(n: int):
    """(ix, iy) of the Padua points in emission order on the tensor Chebyshev-Lobatto grid
    x = cos(ix pi / n) (ix = 0..n), y = cos(iy pi / (n+1)) (iy = 0..n+1): the points with ix + iy = n+1 mod 2."""
    ix, iy = [], []
    NN = n // 2 + 1
    for k in range(n, -1, -1):
        delta = (k % 2) if (n % 2 == 1) else 0
        for j in range(NN + delta, 0, -1):
            ix.append(n - k)                                   # x = -cospi(k/n) = cos((n-k) pi/n)
            iy.append(n + 2 - 2 * j if (n - k) % 2 == 1 else n + 3 - 2 * j)
    return np.array(ix), np.array(iy)


def paduapoints(n: int) -> np.ndarray:
    """N x 2 array [x y] of the (n+1)(n+2)/2 Padua points of degree n on [-1,1]^2."""
    if n == 0:
        return np.array([[0.0, 0.0]]) * 0.0 + np.array([[-1.0, -1.0]])  # -cospi(0/0) is undefined upstream; one point
    ix, iy = padua_grid_index(n)
    x = -_cospi((n - ix) / n)
    y = -_cospi((n + 1 - iy) / (n + 1))
    return np.stack([x, y], axis=1)


def _cospi(t):
    t = np.asarray(t, dtype=np.float64)
    return np.sin(np.pi * (0.5 - np.abs(((t + 1.0) % 2.0) - 1.0)))  # cos(pi t) through a reduced argument


def padua_transform_direct(vals: np.ndarray) -> np.ndarray:
    """Ground truth: the unique total-degree-n interpolant at the Padua points, by solving the N x N collocation
    system in the ApproxFun coefficient order (small n only)."""
    N = len(vals)
    n = padua_degree(N)
    P = paduapoints(n)
    Tx = np.cos(np.outer(np.arccos(np.clip(P[:, 0], -1, 1)), np.arange(n + 1)))
    Ty = np.cos(np.outer(np.arccos(np.clip(P[:, 1], -1, 1)), np.arange(n + 1)))
    cols = [Tx[:, i] * Ty[:, d - i] for d in range(n + 1) for i in range(d + 1)]
    return np.linalg.solve(np.stack(cols, axis=1), np.asarray(vals, dtype=np.float64))


def padua_transform(vals: np.ndarray) -> np.ndarray:
    """`plan_paduatransform!(v, Val{false}) * v`: values at :func:`paduapoints` -> coefficients ordered by total
    degree.  Zero-filled (n+2) x (n+1) grid -> 2-D REDFT00 -> 2/(n(n+1)), first/last row and column halved ->
    triangle."""
    vals = np.asarray(vals, dtype=np.float64)
    N = len(vals)
    n = padua_degree(N)
    if (n + 1) * (n + 2) // 2 != N:
        raise ValueError("Padua transforms can only be applied to vectors of length (n+1)*(n+2)/2")
    if n == 0:
        return vals.copy()
    ix, iy = padua_grid_index(n)
    W = np.zeros((n + 2, n + 1))
    W[iy, ix] = vals
    C = _sfft.dct(_sfft.dct(W, type=1, axis=0), type=1, axis=1) * (2.0 / (n * (n + 1)))
    C[0, :] *= 0.5
    C[-1, :] *= 0.5
    C[:, 0] *= 0.5
    C[:, -1] *= 0.5
    return np.array([C[d - i, i] for d in range(n + 1) for i in range(d + 1)])


def padua_itransform(cfs: np.ndarray) -> np.ndarray:
    """`plan_ipaduatransform!`: coefficients (total-degree order) -> values at the Padua points."""
    cfs = np.asarray(cfs, dtype=np.float64)
    N = len(cfs)
    n = padua_degree(N)
    if (n + 1) * (n + 2) // 2 != N:
        raise ValueError("Padua transforms can only be applied to vectors of length (n+1)*(n+2)/2")
    if n == 0:
        return cfs.copy()
    C = tensor_deinterlace(cfs, n)                      # C[i, j]: x-degree i, y-degree j
    ix, iy = padua_grid_index(n)
    Tx = np.cos(np.pi * np.outer(np.arange(n + 1), np.arange(n + 1)) / n)            # T_i(x_ix)
    Ty = np.cos(np.pi * np.outer(np.arange(n + 1), np.arange(n + 2)) / (n + 1))      # T_j(y_iy)
    V = Tx.T @ C @ Ty                                   # V[ix, iy]
    return V[ix, iy]


# --------------------------------------------------------------------------------------
# 8(f) rank 4: three-term recurrence tables  p_{j+1} = (A_j x + B_j) p_j - C_j p_{j-1}  (DLMF 18.9.1-2),
# the `recA / recB / recC` of UPSTREAM ApproxFunOrthogonalPolynomials (Jacobi(b, a) = P^{(a,b)}, Ultraspherical(l) = C^{(l)})
# --------------------------------------------------------------------------------------
def jacobi_rec(a: float, b: float, N: int):
    """A[0..N-1], B[0..N-1], C[0..N] for P_n^{(a,b)}."""
    n = np.arange(N + 1, dtype=np.float64)
    s = 2 * n + a + b
    with np.errstate(divide="ignore", invalid="ignore"):
        A = (s + 1) * (s + 2) / (2 * (n + 1) * (n + a + b + 1))
        Bv = (a * a - b * b) * (s + 1) / (2 * (n + 1) * (n + a + b + 1) * s)
        Cv = (n + a) * (n + b) * (s + 2) / ((n + 1) * (n + a + b + 1) * s)
    A[0] = (a + b + 2) / 2.0                      # n = 0 limits (s may vanish)
    Bv[0] = (a - b) / 2.0
    Cv[0] = 0.0
    return A[:N].copy(), Bv[:N].copy(), Cv


def ultraspherical_rec(lam: float, N: int):
    """A, B, C for C_n^{(lam)}: A_n = 2(n+lam)/(n+1), B_n = 0, C_n = (n+2lam-1)/(n+1)."""
    n = np.arange(N + 1, dtype=np.float64)
    return (2 * (n + lam) / (n + 1))[:N].copy(), np.zeros(N), (n + 2 * lam - 1) / (n + 1)


def chebyshev_rec(N: int):
    A = np.full(N, 2.0)
    if N:
        A[0] = 1.0
    return A, np.zeros(N), np.ones(N + 1)


def clenshaw_rec(c, A, B, Cc, x):
    """numpy restatement of the general Clenshaw (no fma; :mod:`oracle_c` has the fused one)."""
    c = np.asarray(c, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    N = c.size
    if N == 0:
        return np.zeros_like(x)
    bk1 = np.zeros_like(x)
    bk2 = np.zeros_like(x)
    for j in range(N - 1, 0, -1):
        bk2, bk1 = bk1, (A[j] * x + B[j]) * bk1 + (c[j] - Cc[j + 1] * bk2)
    return (A[0] * x + B[0]) * bk1 + (c[0] - Cc[1] * bk2)


# --------------------------------------------------------------------------------------
# a6  constructors (host orchestration)
# --------------------------------------------------------------------------------------


def fun_coefficients(f, n: int, a: float = -1.0, b: float = 1.0) -> np.ndarray:
    """``Fun(f, Chebyshev(a..b), n)`` -> coefficients: points -> f.(pts) -> transform
    (mirror ext/ApproxFunDualNumbersExt.jl:79-82, docs/src/faq.md:14-17)."""
    return cheb_transform(f(chebpoints(n, a, b)))


def chop(c: np.ndarray, tol: float) -> np.ndarray:
    """UPSTREAM ``chop!(f, tol)``: drop the trailing coefficients with |c| <= tol."""
    nz = np.nonzero(np.abs(c) > tol)[0]
    return c[: (nz[-1] + 1 if nz.size else 0)].copy()


def adaptive_fun_coefficients(f, a: float = -1.0, b: float = 1.0):
    """Adaptive ``Fun(f, Chebyshev(a..b))``: ladder n = 2^4..2^20 (mirror ext/...:98-110), tail test
    on the last 8 coefficients against 20 eps * max|c| (UPSTREAM default_Fun), then chop."""
    tol = 20 * np.finfo(np.float64).eps
    for logn in range(4, 21):
        n = 2 ** logn
        c = fun_coefficients(f, n, a, b)
        mx = np.max(np.abs(c))
        if mx == 0:
            return np.zeros(1)
        if np.max(np.abs(c[-8:])) < tol * mx:
            return chop(c, tol * mx)
    return fun_coefficients(f, 2 ** 21, a, b)


# --------------------------------------------------------------------------------------
# a11 / a15  2-D sampling through a low-rank representation (src/Extras/sample.jl:205-214)
# --------------------------------------------------------------------------------------


def lowrank_factors(Cm: np.ndarray, tol: float = 1e-15):
    """Low-rank factors of a coefficient matrix C[i,j] (T_i(x) T_j(y)):  f = sum_r A_r(x) B_r(y).

    UPSTREAM ApproxFunBase builds LowRankFun by adaptive cross approximation; any factorisation gives the same
    conditional densities, so this restatement uses the SVD (A = U*S, B = V) truncated at tol * s_1.
    Returns (A, B): coefficient matrices (n1, R) and (n2, R)."""
    U, S, Vt = np.linalg.svd(np.asarray(Cm, dtype=np.float64), full_matrices=False)
    R = max(1, int(np.sum(S > tol * S[0])))
    return U[:, :R] * S[:R], Vt[:R].T.copy()


def sample2d_lowrank(A, B, dom1, dom2, u_y, u_x) -> np.ndarray:
    """Literal restatement of sample(f::LowRankFun{Chebyshev,Chebyshev}, n) (sample.jl:205-214) with the two
    `rand(n)` draws supplied by the caller.  A: (n1, R) coefficients of f.A on dom1, B: (n2, R) of f.B on dom2.
    Returns the n x 2 matrix [rx ry]."""
    a1, b1 = dom1
    a2, b2 = dom2
    w = cheb_definite_integrals(A.shape[0]) * (b1 - a1) / 2
    marg = B @ (w @ A)                                   # sum(f,1) = sum_r (int A_r) B_r : a Fun on dom2
    ry = sample_from_uniforms(marg, a2, b2, u_y)         # ry = sample(sum(f,1), n)
    inside = (ry >= min(a1, b1)) & (ry <= max(a1, b1))
    fA = clenshaw_multi(A, tocanonical(a1, b1, ry)) * inside      # fA = evaluate.(f.A, ry')
    AB = B @ fA                                          # CB = coefficientmatrix(f.B); AB = CB*fA
    AB = chebnormalizedcumsum(AB, grow=False)
    rx = chebbisectioninv_cols(AB, u_x)
    return np.stack([fromcanonical(a1, b1, rx), ry], axis=1)   # [fromcanonical.(domain(f,1), rx) ry]


# --------------------------------------------------------------------------------------
# "next" rows: second-kind Chebyshev points and SinSpace
# --------------------------------------------------------------------------------------


def cheb2points_canonical(n: int) -> np.ndarray:
    """UPSTREAM FastTransforms.chebyshevpoints(Float64, n, Val(2)): x_j = cos(pi j/(n-1)), descending."""
    if n == 1:
        return np.zeros(1)
    j = np.arange(n, dtype=np.float64)
    return sinpi((n - 2.0 * j - 1.0) / (2.0 * n - 2.0))


def chebkind2_transform(v: np.ndarray, axis: int = 0) -> np.ndarray:
    """UPSTREAM FastTransforms plan_chebyshevtransform(v, Val(2)): FFTW REDFT00, /(n-1), first and last halved
    (the Val(kind) boundary is visible at ext/ApproxFunDualNumbersExt.jl:44-45).  n == 1: identity."""
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[axis]
    if n <= 1:
        return v.copy()
    y = np.moveaxis(_sfft.dct(v, type=1, axis=axis), axis, 0) / (n - 1)
    y[0] /= 2
    y[-1] /= 2
    return np.moveaxis(y, 0, axis)


def chebkind2_itransform(c: np.ndarray, axis: int = 0) -> np.ndarray:
    """plan_ichebyshevtransform(c, Val(2)): ends doubled, REDFT00, /2  =>  v_j = sum_k c_k cos(pi j k/(n-1))."""
    c = np.moveaxis(np.array(c, dtype=np.float64, copy=True), axis, 0)
    n = c.shape[0]
    if n <= 1:
        return np.moveaxis(c, 0, axis)
    c[0] *= 2
    c[-1] *= 2
    return np.moveaxis(_sfft.dct(c, type=1, axis=0) / 2, 0, axis)


def sin_transform_generic(x: np.ndarray) -> np.ndarray:
    """Verbatim src/Extras/fftGeneric.jl:74-78:  v = imag(fft([0;-x;0;reverse(x)]))[2:length(x)+1]; v/(length(x)+1)."""
    x = np.asarray(x, dtype=np.float64)
    s = np.concatenate([[0.0], -x, [0.0], x[::-1]])
    return np.imag(np.fft.fft(s))[1:x.size + 1] / (x.size + 1)


def sin_itransform_generic(x: np.ndarray) -> np.ndarray:
    """Verbatim src/Extras/fftGeneric.jl:80-81:  imag(fft([0;-x;0;reverse(x)]))[2:length(x)+1]/2."""
    x = np.asarray(x, dtype=np.float64)
    s = np.concatenate([[0.0], -x, [0.0], x[::-1]])
    return np.imag(np.fft.fft(s))[1:x.size + 1] / 2


def sinpoints(n: int) -> np.ndarray:
    """Grid of the SinSpace plans above: theta_j = pi j/(n+1), j = 1..n (f_k <-> sin(k theta), k = 1..n)."""
    return np.pi * np.arange(1, n + 1) / (n + 1)


# ------------------------------------------------------------------------------------------------------
# generic bisection and piecewise Funs  (/root/reference/src/Extras/sample.jl:8-36; test/PiecewiseSampleTest.jl:5-10)
# ------------------------------------------------------------------------------------------------------
def genmidpoint(a: float, b: float) -> float:
    """sample.jl:11-21, verbatim: a midpoint that also works for infinite end points."""
    if math.isinf(a) and math.isinf(b):
        return 0.0
    if math.isinf(a):
        return b - 100
    if math.isinf(b):
        return a + 100
    return (a + b) / 2


def piecewise_eval(pieces, breaks, x, clenshaw_fn=None):
    """Evaluation of a Fun on a PiecewiseSpace of Chebyshev segments: the FIRST piece whose closed interval contains x
    (UPSTREAM ApproxFunBase evaluate for PiecewiseSpace loops over the components in order), evaluated at
    tocanonical(piece, x) = (2x - lo - hi)/(hi - lo); zero outside the domain."""
    cl = clenshaw_fn or clenshaw
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    out = np.zeros_like(x)
    done = ~((x >= breaks[0]) & (x <= breaks[-1]))
    for k, c in enumerate(pieces):
        lo, hi = breaks[k], breaks[k + 1]
        sel = (~done) & (x <= hi)
        if np.any(sel):
            c = np.asarray(c, dtype=np.float64)
            xc = (2.0 * x[sel] - lo - hi) / (hi - lo)
            out[sel] = 0.0 if c.size == 0 else (c[0] if c.size == 1 else cl(c, xc))
            done |= sel
    return out


def bisectioninv_generic(pieces, breaks, u, numits: int = 47, clenshaw_fn=None) -> np.ndarray:
    """sample.jl:23-36 applied entry by entry (sample.jl:36), all entries advanced together:
    a = minimum(d); b = maximum(d); numits x { m = genmidpoint(a,b); val = f(m); (val <= x) ? a = m : b = m }; .5*(a+b)."""
    u = np.atleast_1d(np.asarray(u, dtype=np.float64))
    a = np.full_like(u, breaks[0])
    b = np.full_like(u, breaks[-1])
    for _ in range(numits):
        if np.isinf(a).any() or np.isinf(b).any():
            m = np.array([genmidpoint(ai, bi) for ai, bi in zip(a, b)])
        else:
            m = (a + b) / 2
        val = piecewise_eval(pieces, breaks, m, clenshaw_fn)
        le = val <= u
        a = np.where(le, m, a)
        b = np.where(le, b, m)
    return 0.5 * (a + b)


# ------------------------------------------------------------------------------------------------------
# mixed tensor spaces: factor-1 plan on every column, factor-2 (Chebyshev) plan on every row
# (UPSTREAM ApproxFunBase transform!(::TensorSpace, M); callers /root/reference/test/PeriodicXIntervalTest.jl:22-24,
#  test/LaplaceInAStripTest.jl:5-8)
# ------------------------------------------------------------------------------------------------------
def fourier_cheb_transform(V: np.ndarray) -> np.ndarray:
    return cheb_transform(fourier_transform(V, axis=0), axis=1)


def fourier_cheb_itransform(C: np.ndarray) -> np.ndarray:
    return fourier_itransform(cheb_itransform(C, axis=1), axis=0)


def laurent_cheb_transform(V: np.ndarray) -> np.ndarray:
    A = laurent_transform(V, axis=0)
    return cheb_transform(A.real, axis=1) + 1j * cheb_transform(A.imag, axis=1)


def laurent_cheb_itransform(C: np.ndarray) -> np.ndarray:
    A = cheb_itransform(C.real, axis=1) + 1j * cheb_itransform(C.imag, axis=1)
    return laurent_itransform(A, axis=0)


def fourier_basis(n: int, t: np.ndarray) -> np.ndarray:
    """[1, sin t, cos t, sin 2t, cos 2t, ...] (docs/src/usage/spaces.md:86-98), shape (n, len(t))."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    F = np.empty((n, t.size))
    for i in range(n):
        k = (i + 1) // 2
        F[i] = np.sin(k * t) if (i % 2 == 1) else np.cos(k * t)
    return F


def fourier_cheb_eval(Cm: np.ndarray, t, y) -> np.ndarray:
    """sum_{i,j} C[i,j] F_i(t) T_j(y), direct summation (t canonical in [0, 2 pi), y in [-1, 1])."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    F = fourier_basis(Cm.shape[0], t)                                  # n1 x P
    T = np.cos(np.arange(Cm.shape[1])[:, None] * np.arccos(np.clip(y, -1, 1))[None, :])   # n2 x P
    return np.einsum("ip,ij,jp->p", F, Cm, T)
