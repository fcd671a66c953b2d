"""Parity tests proper: the CUDA path (through the C ABI, ctypes) against the CPU oracle.  Needs a B200.

Tolerances are BASELINE.json's: coefficients  max-abs <= 1e-13 * ||v||inf * log2(n);
evaluations <= 1e-13 * ||f||inf (in fact bit-exact against the fma oracle); ordering exact."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import approxfun_oracle as orc  # noqa: E402
from oracle import oracle_c as orcc  # noqa: E402

EPS = np.finfo(np.float64).eps
RNG = np.random.default_rng(2024)


@pytest.fixture(scope="module")
def af():
    import approxfun_b200 as m

    m.api.lowlevel()  # raises if the CUDA library or device is missing: no fallback
    return m


def coef_tol(n, v):
    return 1e-13 * float(np.max(np.abs(v))) * max(1.0, math.log2(max(n, 2)))


SIZES = [1, 2, 3, 5, 8, 16, 31, 32, 33, 64, 100, 128, 256, 500, 512, 1000, 1024, 2048, 4096, 8192, 16384, 32768, 65536]


@pytest.mark.parametrize("n", SIZES)
def test_chebyshev_transform_pair(af, n):
    nb = 4 if n <= 8192 else 2
    v = RNG.standard_normal((n, nb))
    S = af.Chebyshev()
    c = af.plan_transform(S, v) * v
    want = orc.cheb_transform(v, axis=0)
    assert np.max(np.abs(c - want)) <= coef_tol(n, v)
    vi = af.plan_itransform(S, v) * v
    wanti = orc.cheb_itransform(v, axis=0)
    assert np.max(np.abs(vi - wanti)) <= coef_tol(n, wanti)
    back = af.itransform(S, c)
    assert np.max(np.abs(back - v)) <= coef_tol(n, v)  # test/ExtrasTest.jl:61 identity


@pytest.mark.parametrize("n", SIZES)
def test_fourier_transform_pair(af, n):
    nb = 3 if n <= 8192 else 1
    v = RNG.standard_normal((n, nb))
    S = af.Fourier()
    c = af.transform(S, v)
    assert np.max(np.abs(c - orc.fourier_transform(v, axis=0))) <= coef_tol(n, v)
    if n <= 128:  # in-tree BigFloat twin, restated verbatim (src/Extras/fftGeneric.jl:47-54)
        assert np.max(np.abs(c[:, 0] - orc.fourier_transform_generic(v[:, 0]))) <= coef_tol(n, v)
    vi = af.itransform(S, v)
    wanti = orc.fourier_itransform(v, axis=0)
    assert np.max(np.abs(vi - wanti)) <= coef_tol(n, wanti)


def test_long_lines_persistent_tma_staged(af):
    # n = 16384 with more lines than SMs: persistent CTAs, next line staged by a TMA bulk copy (line_kernel_staged);
    # 301 lines = two full rounds of 148 CTAs plus a ragged third
    from approxfun_b200 import _lib as L
    ll = af.api.lowlevel()
    for n, nb in ((16384, 301), (8192, 601)):     # 8192: plain kernel, many lines
        v = RNG.standard_normal((n, nb))
        for kind, ref in ((L.CHEB1_FWD, orc.cheb_transform), (L.CHEB1_INV, orc.cheb_itransform),
                          (L.FOURIER_FWD, orc.fourier_transform), (L.FOURIER_INV, orc.fourier_itransform)):
            want = ref(v, axis=0)
            assert np.max(np.abs(ll.transform(kind, v) - want)) <= coef_tol(n, want), (kind, n)


@pytest.mark.parametrize("n", SIZES)
def test_laurent_transform_pair(af, n):
    nb = 2 if n <= 8192 else 1
    v = RNG.standard_normal((n, nb)) + 1j * RNG.standard_normal((n, nb))
    S = af.Laurent()
    c = af.transform(S, v)
    assert np.max(np.abs(c - orc.laurent_transform(v, axis=0))) <= coef_tol(n, v)
    vi = af.itransform(S, v)
    wanti = orc.laurent_itransform(v, axis=0)
    assert np.max(np.abs(vi - wanti)) <= coef_tol(n, wanti)


def test_config1_sin_x2_2pow20(af):
    # BASELINE.json configs[0]: transform + itransform of sin(x^2) on 0..10 at n = 2^20
    n = 2 ** 20
    S = af.Chebyshev(0, 10)
    x = af.points(S, n)
    xo = orc.chebpoints(n, 0.0, 10.0)
    assert np.all(np.diff(x) < 0)                        # descending, NEWS.md:92
    assert np.max(np.abs(x - xo)) <= 10 * 2.3e-16 * 10   # values to an ulp of the interval scale
    v = np.sin(xo ** 2)
    c = af.plan_transform(S, v) * v
    assert np.max(np.abs(c - orc.cheb_transform(v))) <= 1e-13 * 20
    back = af.plan_itransform(S, c) * c
    assert np.max(np.abs(back - v)) <= 1e-13 * 20
    f = af.Fun(S, c)
    assert abs(f(0.1) - math.sin(0.01)) < 1000 * EPS      # test/ReadmeTest.jl:19-23


def test_very_long_single_transform(af):
    # n = 2^24 (128 MiB): the four-step path far above config 1; coefficients of a known series, round trip, oracle subsample
    n = 1 << 24
    S = af.Chebyshev()
    x = af.points(S, n)
    v = np.cos(3 * np.arccos(np.clip(x, -1, 1))) * 2.0 + 0.5 + 0.25 * x      # 0.5 T0 + 0.25 T1 + 2 T3
    c = af.transform(S, v)
    want = np.zeros(8)
    want[[0, 1, 3]] = [0.5, 0.25, 2.0]
    assert np.max(np.abs(c[:8] - want)) < 1e-12 and np.max(np.abs(c[8:])) < 1e-12
    r = RNG.standard_normal(n)
    cr = af.transform(S, r)
    assert abs(cr[0] - r.mean()) < 1e-13
    assert np.max(np.abs(af.itransform(S, cr) - r)) < coef_tol(n, r)
    assert np.max(np.abs(cr - orc.cheb_transform(r))) < coef_tol(n, r)


def test_golden_fixture(af):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
    for n in (8, 64, 100, 4096):
        v = g[f"v_{n}"]
        assert np.max(np.abs(af.transform(af.Chebyshev(), v) - g[f"cheb_fwd_{n}"])) <= coef_tol(n, v)
        assert np.max(np.abs(af.itransform(af.Chebyshev(), v) - g[f"cheb_inv_{n}"])) <= coef_tol(n, g[f"cheb_inv_{n}"])
        assert np.max(np.abs(af.transform(af.Fourier(), v) - g[f"fourier_fwd_{n}"])) <= coef_tol(n, v)
        assert np.max(np.abs(af.itransform(af.Fourier(), v) - g[f"fourier_inv_{n}"])) <= coef_tol(n, g[f"fourier_inv_{n}"])
    assert np.array_equal(af.clenshaw(af.Chebyshev(), g["cl_c"], g["cl_x"]), g["cl_y"])
    assert np.array_equal(af.chebbisectioninv(g["bis_c"], g["bis_u"]), g["bis_x"])
    assert np.array_equal(af.chebnormalizedcumsum(g["ncs_in"]), g["ncs_out"])


def test_doctest_orderings(af):
    # docs/src/usage/spaces.md:94-98 and :121-125
    th = orc.fourierpoints(5)
    c = np.array([1.0, 2.0, 3.0, 4.0, 0.0])
    assert np.allclose(af.itransform(af.Fourier(), c), 1 + 2 * np.sin(th) + 3 * np.cos(th) + 4 * np.sin(2 * th), atol=50 * EPS)
    cl = np.array([1, 2, 3, 4, 0], dtype=complex)
    want = 1 + 2 * np.exp(-1j * th) + 3 * np.exp(1j * th) + 4 * np.exp(-2j * th)
    assert np.allclose(af.itransform(af.Laurent(), cl), want, atol=50 * EPS)
    th = orc.fourierpoints(20)  # even-n Nyquist rule, fftGeneric.jl:53
    c = af.transform(af.Fourier(), 5 * np.cos(10 * th))
    assert abs(c[-1] + 5) < 50 * EPS


def test_fun_exp_bessel_and_readme(af):
    from scipy import special
    f = af.Fun(np.exp, af.Chebyshev(), 32)
    assert abs(f.coefficients[0] - special.iv(0, 1.0)) <= 2 * EPS
    assert abs(f.coefficients[1] - 2 * special.iv(1, 1.0)) <= 4 * EPS
    g = af.Fun(lambda x: np.sin(x * x), af.Chebyshev(0, 10))       # adaptive ladder
    assert abs(g(0.1) - math.sin(0.01)) < 1000 * EPS               # test/ReadmeTest.jl:23
    assert np.max(np.abs(af.transform(g.space, af.values(g)) - g.coefficients)) < 10 * EPS  # ExtrasTest.jl:61


@pytest.mark.parametrize("N", [0, 1, 2, 3, 66, 1000, 10000, 12289, 40000])
def test_clenshaw_bit_exact(af, N):
    c = RNG.standard_normal(N) / (1 + np.arange(N)) ** 1.2
    x = RNG.uniform(-1, 1, 100003)
    x[:2] = [-1.0, 1.0]
    got = af.clenshaw(af.Chebyshev(), c, x)
    assert np.array_equal(got, orcc.clenshaw(c, x, threads=8))
    if N == 66:
        assert af.clenshaw(af.Chebyshev(), c, 0.3) == orcc.clenshaw(c, np.array([0.3]))[0]


def test_clenshaw_config3_subsample(af):
    # config 3 shape: N = 10^4 coefficients of a function that needs them; 10^6-point subsample
    N = 10000
    xs = orc.chebpoints_canonical(N)
    c = orc.cheb_transform(np.sin(3000 * xs ** 2) + np.cos(977 * xs))
    x = np.random.default_rng(0).uniform(-1, 1, 10 ** 6)
    got = af.clenshaw(af.Chebyshev(), c, x)
    want = orcc.clenshaw(c, x, threads=8)
    assert np.array_equal(got, want)
    truth = np.sin(3000 * x ** 2) + np.cos(977 * x)
    assert np.max(np.abs(got - truth)) <= 1e-13 * 2 * 50   # interpolant accuracy, not parity


def test_clenshaw_matrix_variants(af):
    for N, P in [(0, 5), (1, 5), (7, 1000), (150, 257), (3000, 9)]:
        Cm = RNG.standard_normal((N, P))
        x = RNG.uniform(-1, 1, P)
        want = orcc.clenshaw_cols(Cm, x) if N else np.zeros(P)
        assert np.array_equal(af.clenshaw(Cm, x, af.ClenshawPlan(float, af.Chebyshev(), N, P)), want)
    Cm = RNG.standard_normal((40, 6))
    x = RNG.uniform(-1, 1, 5000)
    assert np.array_equal(af.api.lowlevel().clenshaw_multi(Cm, x), orcc.clenshaw_multi(Cm, x))
    with pytest.raises(af.DimensionMismatch):
        af.clenshaw(RNG.standard_normal((4, 6)), np.zeros(5))


def test_sampling_golden_and_parity(af):
    # test/SamplingTest.jl:10-12
    f = af.Fun(lambda x: np.exp(-x * x / 2), af.Chebyshev(-5, 5))
    g = af.cumsum(f)
    assert abs(g(af.bisectioninv(g, 0.5)) - 0.5) < 1e-12
    # config 5 shape: f = exp(-x^2) on -5..5, caller-supplied uniforms, bit-exact against the oracle
    f = af.Fun(lambda x: np.exp(-x * x), af.Chebyshev(-5, 5))
    cf = af.normalizedcumsum(f)
    assert np.max(np.abs(cf.coefficients - orc.normalizedcumsum(f.coefficients, -5.0, 5.0))) < 1e-15
    u = np.random.default_rng(5).random(10 ** 6)
    s = af.sample(f, u.size, uniforms=u)
    want = orc.fromcanonical(-5.0, 5.0, orcc.cheb_bisect(cf.coefficients, u, threads=8))
    assert np.array_equal(s, want)
    assert s.min() >= -5 and s.max() <= 5                      # test/PiecewiseSampleTest.jl:5-10 range check
    from scipy import special
    assert np.max(np.abs((1 + special.erf(s)) / 2 - u)) < 1e-12
    assert len(af.sample(af.Fun(np.exp, af.Chebyshev()), 100000)) == 100000   # test/SampleTest.jl:5-7 smoke


def test_normcumsum_and_matrix_bisection(af):
    for N in (1, 2, 3, 4, 9, 66, 300):
        c = np.abs(RNG.standard_normal(N)) / (1 + np.arange(N)) ** 2
        c[0] += 2.0
        assert np.array_equal(af.chebnormalizedcumsum(c), orcc.cheb_normcumsum(c))
    M = np.abs(RNG.standard_normal((30, 5000))) / (1 + np.arange(30))[:, None] ** 2
    M[0] += 2
    gm = af.chebnormalizedcumsum(M)
    assert np.array_equal(gm, orcc.cheb_normcumsum(M))
    u = RNG.random(5000)
    assert np.array_equal(af.chebbisectioninv(gm, u), orcc.cheb_bisect_cols(gm, u))
    with pytest.raises(af.DimensionMismatch):                       # src/Extras/sample.jl:85
        af.chebbisectioninv(gm, u[:10])


def test_philox_device_sampler_matches_host_uniform_path(af):
    import torch
    f = af.Fun(lambda x: np.exp(-x * x), af.Chebyshev(-5, 5))
    cf = af.normalizedcumsum(f).coefficients
    P = 200000
    d_c = torch.tensor(cf, device="cuda")
    d_u = torch.empty(P, dtype=torch.float64, device="cuda")
    d_s = torch.empty(P, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    af.api.philox_device(7, 1000, d_u.data_ptr(), P, st)
    af.api.sample_device(d_c.data_ptr(), cf.size, -5.0, 5.0, 7, 1000, d_s.data_ptr(), P, st)
    torch.cuda.synchronize()
    u = d_u.cpu().numpy()
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert np.array_equal(d_s.cpu().numpy(), orc.fromcanonical(-5.0, 5.0, orcc.cheb_bisect(cf, u, threads=8)))
    # host-pointer sampler with device uniforms (afb_sample_cheb): counter i for sample i, chunked over two streams
    Ph = (1 << 23) + 12345                                   # more than one chunk
    d_u2 = torch.empty(Ph, dtype=torch.float64, device="cuda")
    af.api.philox_device(11, 0, d_u2.data_ptr(), Ph, st)
    torch.cuda.synchronize()
    hs = af.api.lowlevel().sample_cheb(cf, -5.0, 5.0, 11, Ph)
    pick = np.r_[0:1000, (1 << 23) - 500:(1 << 23) + 500, Ph - 1000:Ph]
    want = orc.fromcanonical(-5.0, 5.0, orcc.cheb_bisect(cf, d_u2.cpu().numpy()[pick], threads=8))
    assert np.array_equal(hs[pick], want)
    assert np.array_equal(af.sample(f, 5000, seed=11), hs[:5000])


def test_tensor_transform_and_sampling_golden(af):
    # test/SamplingTest.jl:5-9 golden through the GPU 2-D transform + GPU Clenshaw
    n = 96
    a, b = -4.0, 4.0
    S = af.Chebyshev(a, b)
    x = af.points(S, n)
    V = (x[:, None] - x[None, :]) ** 2 * np.exp(-x[:, None] ** 2 / 2 - x[None, :] ** 2 / 2)
    Cm = af.transform(S ** 2, V)
    assert np.max(np.abs(Cm - orc.cheb2_transform(V))) < 1e-13
    w = orc.cheb_definite_integrals(n) * (b - a) / 2
    marg = w @ Cm
    total = float(marg @ w)
    assert abs(af.Fun(S, marg)(0.1) / total - 0.2004758624973169) < 4 * EPS
    V2 = RNG.standard_normal((300, 128))
    C2 = af.transform(af.Chebyshev() ** 2, V2)
    assert np.max(np.abs(C2 - orc.cheb2_transform(V2))) < 1e-13 * 4 * 9
    assert np.max(np.abs(af.itransform(af.Chebyshev() ** 2, C2) - V2)) < 1e-12
    F2 = af.Fun(lambda xx, yy: np.cos(xx + 2 * yy), af.Chebyshev() ** 2, (32, 32))
    assert abs(F2(0.1, 0.2) - math.cos(0.5)) < 1e-13


@pytest.mark.parametrize("shape", [(2, 1024), (30, 2048), (256, 4096), (130, 8192), (64, 16384), (4, 1 << 16), (2, 1 << 18)])
def test_tensor_row_pass_without_transposes(af, shape):
    # n2 = 2^10..2^18, n even: the row pass is the strided four-step on row pairs (afb_rows.cu); flag 1 forces the
    # transpose route.  Both against the oracle, and against each other.
    from approxfun_b200 import _lib as L
    ll = af.api.lowlevel()
    V = RNG.standard_normal(shape)
    want = orc.cheb2_transform(V)
    tol = 1e-13 * np.log2(shape[1]) * np.max(np.abs(V))
    got = ll.transform2d(L.CHEB1_2D_FWD, V)
    alt = ll.transform2d(L.CHEB1_2D_FWD, V, L.PLAN_2D_TRANSPOSE)
    assert np.max(np.abs(got - want)) < tol and np.max(np.abs(alt - want)) < tol
    back = ll.transform2d(L.CHEB1_2D_INV, want)
    alt = ll.transform2d(L.CHEB1_2D_INV, want, L.PLAN_2D_TRANSPOSE)
    vtol = 1e-13 * np.log2(shape[1]) * max(1.0, np.max(np.abs(want))) * np.sqrt(shape[0] * shape[1])
    assert np.max(np.abs(back - V)) < vtol and np.max(np.abs(alt - V)) < vtol


@pytest.mark.parametrize("d", [0, 1, 2, 3, 6, 7, 30, 31, 100, 255, 1000])
def test_padua_points_and_transform(af, d):
    # Fun(f, Chebyshev()^2, N) = values at the Padua points -> coefficients by total degree (NEWS.md:85).
    # Oracle: zero-filled grid + scipy REDFT00 (pinned against the N x N collocation solve in test_oracle_goldens.py).
    from approxfun_b200 import _lib as L
    ll = af.api.lowlevel()
    N = (d + 1) * (d + 2) // 2
    P = ll.paduapoints(N)
    Po = orc.paduapoints(d)
    assert P.shape == (N, 2) and np.max(np.abs(P - Po)) <= 5e-16
    v = np.exp(0.3 * Po[:, 0] - 0.7 * Po[:, 1]) + Po[:, 0] * Po[:, 1] ** 2
    c = ll.padua(L.PADUA_FWD, v)
    co = orc.padua_transform(v)
    tol = 1e-13 * max(1.0, np.log2(d + 2)) * np.max(np.abs(v))
    assert np.max(np.abs(c - co)) < tol
    assert np.max(np.abs(ll.padua(L.PADUA_INV, co) - v)) < 10 * tol
    w = RNG.standard_normal(N)
    assert np.max(np.abs(ll.padua(L.PADUA_FWD, w) - orc.padua_transform(w))) < 1e-13 * np.log2(d + 2) * 4
    if d in (7, 100):
        S2 = af.Chebyshev(-1, 2) * af.Chebyshev(0, 3)
        f = lambda x, y: np.cos(x + 2 * y) * np.exp(-x * y / 5)  # noqa: E731
        F = af.Fun(f, S2, N) if d == 100 else af.Fun(lambda x, y: 1 + x + 2 * y + x * y ** 2, S2, N)
        assert F.coefficients.shape == (N,)
        xs, ys = RNG.uniform(-1, 2, 50), RNG.uniform(0, 3, 50)
        want = f(xs, ys) if d == 100 else 1 + xs + 2 * ys + xs * ys ** 2
        assert np.max(np.abs(F(xs, ys) - want)) < 1e-12
        assert np.max(np.abs(af.values(F) - F(*af.points(S2, N).T))) < 1e-12
    with pytest.raises(af.DimensionMismatch):
        ll.padua(L.PADUA_FWD, np.zeros(N + 1))


@pytest.mark.parametrize("n", [1, 2, 3, 5, 17, 64, 65, 100, 1025, 4097])
def test_sibling_spaces_second_kind_sin_cos(af, n):
    # Val(2) Chebyshev plans (REDFT00) and SinSpace (in-tree spec src/Extras/fftGeneric.jl:74-81)
    v = RNG.standard_normal((n, 3))
    c = af.chebyshevtransform(v, kind=2)
    assert np.max(np.abs(c - orc.chebkind2_transform(v, axis=0))) <= coef_tol(n, v)
    assert np.max(np.abs(af.ichebyshevtransform(c, kind=2) - v)) <= 4 * coef_tol(n, v)
    S = af.SinSpace()
    s = af.transform(S, v)
    assert np.max(np.abs(s[:, 0] - orc.sin_transform_generic(v[:, 0]))) <= coef_tol(n, v)
    si = af.itransform(S, v)
    assert np.max(np.abs(si[:, 1] - orc.sin_itransform_generic(v[:, 1]))) <= coef_tol(n, si)
    # docs/src/usage/spaces.md:39-49: CosSpace shares coefficients and transform with Chebyshev
    assert np.array_equal(af.transform(af.CosSpace(), v), af.transform(af.Chebyshev(), v))
    th = af.points(af.CosSpace(), n)
    assert np.allclose(np.cos(th), af.points(af.Chebyshev(), n), atol=4e-16)


def test_dual_number_plans(af):
    # ext/ApproxFunDualNumbersExt.jl:44-63: the plan of a Dual vector is the plan of its real part and acts on both parts
    n = 200
    d = af.DualArray(RNG.standard_normal(n), RNG.standard_normal(n))
    for S, fwd in ((af.Chebyshev(), orc.cheb_transform), (af.Fourier(), orc.fourier_transform)):
        P = af.plan_transform(S, d)
        c = P * d
        assert np.max(np.abs(af.realpart(c) - fwd(d.value))) < 1e-14 and np.max(np.abs(af.dualpart(c) - fwd(d.epsilon))) < 1e-14
        back = af.plan_itransform(S, c) * c
        assert np.max(np.abs(back.value - d.value)) < 1e-13 and np.max(np.abs(back.epsilon - d.epsilon)) < 1e-13


@pytest.mark.parametrize("n", [1, 5, 64, 100, 4096])
def test_taylor_hardy(af, n):
    # docs/src/usage/spaces.md:100-110
    v = RNG.standard_normal((n, 2)) + 1j * RNG.standard_normal((n, 2))
    F = np.fft.fft(v, axis=0) / n
    assert np.max(np.abs(af.transform(af.Taylor(), v) - F)) <= coef_tol(n, v)
    assert np.max(np.abs(af.transform(af.Hardy(), v) - F[::-1])) <= coef_tol(n, v)
    assert np.max(np.abs(af.itransform(af.Taylor(), af.transform(af.Taylor(), v)) - v)) <= 4 * coef_tol(n, v)
    assert np.max(np.abs(af.itransform(af.Hardy(), af.transform(af.Hardy(), v)) - v)) <= 4 * coef_tol(n, v)


def test_bivariate_clenshaw(af):
    for n1, n2, P in [(1, 1, 5), (9, 7, 100000), (96, 96, 20000), (130, 100, 500)]:
        Cm = RNG.standard_normal((n1, n2)) / (1 + np.arange(n1))[:, None]
        x, y = RNG.uniform(-1, 1, P), RNG.uniform(-1, 1, P)
        assert np.array_equal(af.api.lowlevel().clenshaw2d(Cm, x, y), orcc.clenshaw2d(Cm, x, y, threads=8))
    # examples/PDE_Helmholtz.jl:33-36 shape: evaluate a 2-D Fun at single points and on a grid
    F = af.Fun(lambda xx, yy: np.cos(xx + 2 * yy) * np.exp(xx * yy), af.Chebyshev() ** 2, (48, 48))
    assert abs(F(0.1, 0.2) - math.cos(0.5) * math.exp(0.02)) < 1e-13
    # the same function through the Padua constructor Fun(f, Chebyshev()^2, N): flat coefficients, same values
    Fp = af.Fun(lambda xx, yy: np.cos(xx + 2 * yy) * np.exp(xx * yy), af.Chebyshev() ** 2, 1326)
    assert Fp.coefficients.shape == (1326,) and abs(Fp(0.1, 0.2) - F(0.1, 0.2)) < 1e-13
    gx = np.linspace(-1, 1, 50)
    G = F(gx[:, None], gx[None, :])
    assert np.max(np.abs(G - np.cos(gx[:, None] + 2 * gx[None, :]) * np.exp(gx[:, None] * gx[None, :]))) < 1e-12


def test_general_three_term_clenshaw(af):
    # Jacobi / Ultraspherical / Legendre evaluation through recA/recB/recC; bit-exact against the
    # fused C oracle, and against scipy's polynomials for the mathematics
    import scipy.special as sps
    ll = af.api.lowlevel()
    x = RNG.uniform(-1, 1, 20001)
    for N in (0, 1, 2, 3, 65, 3072, 3073, 10000):
        c = RNG.standard_normal(N) / (1 + np.arange(N))
        for A, Bv, Cv in (orc.jacobi_rec(0.5, 1.5, N), orc.ultraspherical_rec(1.0, N), orc.chebyshev_rec(N)):
            assert np.array_equal(ll.clenshaw_rec(c, A, Bv, Cv, x), orcc.clenshaw_rec(c, A, Bv, Cv, x, threads=8))
    N = 30
    c = RNG.standard_normal(N) / (1 + np.arange(N)) ** 2
    xs = x[:500]
    f = af.Fun(af.Jacobi(1.5, 0.5), c)   # Jacobi(b, a) = P^{(a,b)} = P^{(0.5, 1.5)}
    assert np.max(np.abs(f(xs) - sum(c[n] * sps.eval_jacobi(n, 0.5, 1.5, xs) for n in range(N)))) < 1e-12
    g = af.Fun(af.Ultraspherical(2.0, 0.0, 4.0), c)
    assert np.max(np.abs(g(2 + 2 * xs) - sum(c[n] * sps.eval_gegenbauer(n, 2.0, xs) for n in range(N)))) < 1e-12
    h = af.Fun(af.Legendre(), c)
    assert abs(h(0.3) - sum(c[n] * sps.eval_legendre(n, 0.3) for n in range(N))) < 1e-14
    with pytest.raises(af.DimensionMismatch):
        ll.clenshaw_rec(np.ones(5), np.ones(4), np.ones(5), np.ones(6), xs)


def test_sampling_2d_lowrank(af):
    # test/SamplingTest.jl:5-8: f = (x-y)^2 exp(-x^2/2-y^2/2) on (-4..4)^2; r = sample(f, 5000)
    S = af.Chebyshev(-4, 4)
    f = af.Fun(lambda x, y: (x - y) ** 2 * np.exp(-x ** 2 / 2 - y ** 2 / 2), S ** 2, (96, 96))
    r = af.sample(f, 5000)
    assert r.shape == (5000, 2) and np.all(np.abs(r) <= 4)
    # parity with caller-supplied uniforms, same low-rank factors on both sides
    lr = af.LowRankFun(f)
    A, B = af.coefficientmatrix(lr.A), af.coefficientmatrix(lr.B)
    rng = np.random.default_rng(11)
    n = 200000
    u_y, u_x = rng.random(n), rng.random(n)
    got = af.sample(lr, n, uniforms=(u_y, u_x))
    marg = af.api.lowrank_sum(lr, 1)
    cdf = af.normalizedcumsum(marg).coefficients
    assert np.max(np.abs(cdf - orc.normalizedcumsum(marg.coefficients, -4.0, 4.0))) < 1e-15
    ry = orc.fromcanonical(-4.0, 4.0, orcc.cheb_bisect(cdf, u_y, threads=8))
    assert np.array_equal(got[:, 1], ry)                                     # 1-D stage: bit exact
    rx_c = orcc.sample2d_lowrank(A, B, -4.0, 4.0, -4.0, 4.0, ry, u_x, threads=8)
    assert np.array_equal(got[:, 0], rx_c)                                    # fused conditional stage: bit exact
    want = orc.sample2d_lowrank(A, B, (-4.0, 4.0), (-4.0, 4.0), u_y, u_x)    # literal sample.jl:205-214 with a BLAS dgemm
    d = np.abs(got - want)
    assert np.mean(d[:, 0] < 1e-9) > 0.999 and np.max(d[:, 1]) < 1e-9         # dgemm summation order may flip rare ties
    # the samples follow f: compare E[x y] with quadrature of the density
    x = af.points(S, 96)
    V = (x[:, None] - x[None, :]) ** 2 * np.exp(-x[:, None] ** 2 / 2 - x[None, :] ** 2 / 2)
    w = orc.cheb_definite_integrals(96) * 4.0
    Cm = orc.cheb2_transform(V)
    Cxy = orc.cheb2_transform(V * x[:, None] * x[None, :])
    exy = float(w @ Cxy @ w) / float(w @ Cm @ w)
    assert abs(np.mean(got[:, 0] * got[:, 1]) - exy) < 0.02
    # stand-alone dgemm entry point (sample.jl:210)
    fA = rng.standard_normal((B.shape[1], 1000))
    assert np.max(np.abs(af.api.lowlevel().dgemm_nn(B, fA) - B @ fA)) < 1e-12


def test_config2_full_size_properties(af):
    # BASELINE.json configs[1] at full size, device resident: round trip + linearity + column subsample vs oracle
    import torch
    n, B = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(0)
    v = torch.randn(B, n, dtype=torch.float64, device="cuda", generator=g)
    c = torch.empty_like(v)
    st = torch.cuda.current_stream().cuda_stream
    fwd = af.api.DevicePlan(af._lib.CHEB1_FWD, n, 0, B, n)
    inv = af.api.DevicePlan(af._lib.CHEB1_INV, n, 0, B, n)
    fwd.execute(v.data_ptr(), c.data_ptr(), st)
    torch.cuda.synchronize()
    idx = [0, 1, 4095, 32768, 65535]
    want = orc.cheb_transform(v[idx].cpu().numpy().T, axis=0).T
    vmax = float(v.abs().max())
    assert np.max(np.abs(c[idx].cpu().numpy() - want)) <= 1e-13 * vmax * 12
    # c_0 is the mean of every column: a size-independent checksum
    assert float((c[:, 0] - v.mean(dim=1)).abs().max()) <= 1e-13 * vmax
    back = torch.empty_like(v)
    inv.execute(c.data_ptr(), back.data_ptr(), st)
    torch.cuda.synchronize()
    assert float((back - v).abs().max()) <= 1e-13 * vmax * 12
    # Fourier at full size: round trip
    ff = af.api.DevicePlan(af._lib.FOURIER_FWD, n, 0, B, n)
    fi = af.api.DevicePlan(af._lib.FOURIER_INV, n, 0, B, n)
    ff.execute(v.data_ptr(), c.data_ptr(), st)
    fi.execute(c.data_ptr(), back.data_ptr(), st)
    torch.cuda.synchronize()
    assert float((back - v).abs().max()) <= 1e-13 * vmax * 12
    wantf = orc.fourier_transform(v[idx].cpu().numpy().T, axis=0).T
    assert np.max(np.abs(c[idx].cpu().numpy() - wantf)) <= 1e-13 * vmax * 12


def test_config4_full_size_properties(af):
    # BASELINE.json configs[3] at full size (16384 x 16384, device resident): a rank-one grid V = a b^T must give
    # C = T(a) T(b)^T (checked on sampled rows / columns against the 1-D oracle), a random grid must survive the round
    # trip, and the row-pair route must agree with the explicit-transpose route.
    import torch
    n = 16384
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(4)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    ca, cb = orc.cheb_transform(a), orc.cheb_transform(b)
    # column-major n x n2 matrix V[k, j] = a_k b_j  <->  torch (n2, n) row-major with entry [j, k]
    V = torch.outer(torch.tensor(b, device="cuda"), torch.tensor(a, device="cuda")).contiguous()
    C = torch.empty_like(V)
    fwd = af.api.DevicePlan(af._lib.CHEB1_2D_FWD, n, n, 1, n)
    fwd.execute(V.data_ptr(), C.data_ptr(), st)
    torch.cuda.synchronize()
    scale = float(np.max(np.abs(ca)) * np.max(np.abs(cb)))
    for j in (0, 1, 77, 8191, 8192, 16383):          # coefficient columns C[:, j] = ca * cb[j]
        assert np.max(np.abs(C[j].cpu().numpy() - ca * cb[j])) <= 1e-13 * 14 * scale
    for k in (0, 3, 4096, 16382):                     # coefficient rows C[k, :] = ca[k] * cb
        assert np.max(np.abs(C[:, k].cpu().numpy() - ca[k] * cb)) <= 1e-13 * 14 * scale
    alt = af.api.DevicePlan(af._lib.CHEB1_2D_FWD, n, n, 1, n, af._lib.PLAN_2D_TRANSPOSE)
    C2 = torch.empty_like(V)
    alt.execute(V.data_ptr(), C2.data_ptr(), st)
    torch.cuda.synchronize()
    assert float((C - C2).abs().max()) <= 1e-13 * 14 * scale
    del C2, alt
    g = torch.Generator(device="cuda").manual_seed(1)
    V.normal_(generator=g)
    vmax = float(V.abs().max())
    inv = af.api.DevicePlan(af._lib.CHEB1_2D_INV, n, n, 1, n)
    fwd.execute(V.data_ptr(), C.data_ptr(), st)
    inv.execute(C.data_ptr(), C.data_ptr(), st)     # in place
    torch.cuda.synchronize()
    assert float((C - V).abs().max()) <= 1e-13 * 28 * vmax


def test_plain_c_caller(af, tmp_path):
    # the same C program the CPU suite links: on the GPU it must pass every check through the raw C ABI
    import subprocess
    from tests.test_abi import _build_c_smoke
    out = subprocess.run([_build_c_smoke(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "abi smoke ok" in out.stdout, (out.returncode, out.stdout, out.stderr)


# ---------------------------------------------------------------------------------------------------------------------
# round 2: cases the first round only covered on the CPU emulator, generic / piecewise sampling, host-call pipelines
# ---------------------------------------------------------------------------------------------------------------------
def test_laurent_n65536(af):
    # largest in-CTA complex length + 1 step: the generic path for the complex kinds
    n = 65536
    v = RNG.standard_normal((n, 1)) + 1j * RNG.standard_normal((n, 1))
    c = af.transform(af.Laurent(), v)
    assert np.max(np.abs(c - orc.laurent_transform(v, axis=0))) <= coef_tol(n, v)
    vi = af.itransform(af.Laurent(), v)
    wanti = orc.laurent_itransform(v, axis=0)
    assert np.max(np.abs(vi - wanti)) <= coef_tol(n, wanti)


@pytest.mark.parametrize("n,stride", [(64, 70), (256, 256 + 34), (4096, 4096 + 2), (1000, 1003), (16384, 16384 + 16)])
def test_strided_and_inplace_device_plans(af, n, stride):
    # stride > n (a column range of a larger Julia matrix) and in == out on the device, every real kind
    import torch
    from approxfun_b200 import _lib as L
    nb = 37
    st = torch.cuda.current_stream().cuda_stream
    host = RNG.standard_normal((nb, stride))
    for kind, ref in ((L.CHEB1_FWD, orc.cheb_transform), (L.CHEB1_INV, orc.cheb_itransform),
                      (L.FOURIER_FWD, orc.fourier_transform), (L.FOURIER_INV, orc.fourier_itransform)):
        want = ref(host[:, :n], axis=1)
        plan = af.api.DevicePlan(kind, n, 0, nb, stride)
        d_in = torch.tensor(host, device="cuda")
        d_out = torch.full_like(d_in, -7.0)
        plan.execute(d_in.data_ptr(), d_out.data_ptr(), st)          # out of place
        plan.execute(d_in.data_ptr(), d_in.data_ptr(), st)           # in place
        torch.cuda.synchronize()
        got, got_ip = d_out.cpu().numpy(), d_in.cpu().numpy()
        assert np.max(np.abs(got[:, :n] - want)) <= coef_tol(n, want), (kind, n)
        assert np.array_equal(got[:, n:], np.full((nb, stride - n), -7.0))       # the padding is never written
        assert np.array_equal(got_ip[:, :n], got[:, :n]) and np.array_equal(got_ip[:, n:], host[:, n:])
        # the host-pointer call takes the same strided layout
        hout = np.full_like(host, -7.0)
        plan.execute_host(host, hout)
        assert np.array_equal(hout[:, :n], got[:, :n]) and np.array_equal(hout[:, n:], np.full((nb, stride - n), -7.0))
        plan.close()


@pytest.mark.parametrize("n,batch", [(32, 300000), (8, 1100000), (31, 290000)])
def test_tiny_lines_multi_chunk_host_pipeline(af, n, batch):
    # n <= 32 runs the O(n^2) kernel, which is not in-place safe; batch * n * 8 B > 64 MiB makes afb_plan_execute_host
    # pipeline several chunks over several streams (the round-1 advisor's race: every chunk now has its own in / out halves)
    v = RNG.standard_normal((batch, n))
    from approxfun_b200 import _lib as L
    plan = af.api.DevicePlan(L.CHEB1_FWD, n, 0, batch, n)
    out = np.empty_like(v)
    for _ in range(2):
        plan.execute_host(v, out)
        assert np.max(np.abs(out - orc.cheb_transform(v, axis=1))) <= coef_tol(n, v)
    plan.execute_host(v, v)                                              # in == out on the host
    assert np.array_equal(v, out)
    plan.close()


def test_concurrent_host_calls_on_different_plans(af):
    # include/approxfun_b200.h "Threading": different plans may execute concurrently from different threads
    import threading
    from approxfun_b200 import _lib as L
    n, B = 4096, 8192     # 256 MiB per call: several pipeline chunks each
    va, vb = RNG.standard_normal((B, n)), RNG.standard_normal((B, n))
    pa, pb = af.api.DevicePlan(L.CHEB1_FWD, n, 0, B, n), af.api.DevicePlan(L.FOURIER_FWD, n, 0, B, n)
    oa, ob = np.empty_like(va), np.empty_like(vb)
    errs = []

    def run(plan, src, dst):
        try:
            for _ in range(3):
                plan.execute_host(src, dst)
        except Exception as exc:  # noqa: BLE001
            errs.append(exc)

    ts = [threading.Thread(target=run, args=(pa, va, oa)), threading.Thread(target=run, args=(pb, vb, ob))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    sa, sb = np.empty_like(va), np.empty_like(vb)
    pa.execute_host(va, sa)
    pb.execute_host(vb, sb)
    assert np.array_equal(oa, sa) and np.array_equal(ob, sb)             # same bits as the serial calls
    idx = [0, 17, B - 1]
    assert np.max(np.abs(oa[idx] - orc.cheb_transform(va[idx], axis=1))) <= coef_tol(n, va)
    assert np.max(np.abs(ob[idx] - orc.fourier_transform(vb[idx], axis=1))) <= coef_tol(n, vb)
    pa.close()
    pb.close()


def test_generic_bisection_and_genmidpoint(af):
    # src/Extras/sample.jl:8-36: bisectioninv(f::Fun, x) in the domain variable between minimum(d) and maximum(d)
    f = af.Fun(lambda x: np.exp(-x * x), af.Chebyshev(-5, 5))
    cf = af.normalizedcumsum(f)
    u = np.concatenate([RNG.random(20000), [0.0, 1.0, 0.5, 1e-300, 1 - 2.0 ** -53]])
    got = af.api.bisectioninv_generic(cf, u)
    want = orc.bisectioninv_generic([cf.coefficients], [-5.0, 5.0], u, clenshaw_fn=lambda c, x: orcc.clenshaw(c, x))
    assert np.array_equal(got, want)                                      # bit exact against the fma oracle
    # the Chebyshev dispatch (sample.jl:103-109) bisects the canonical variable: same root, bracket 10 * 2^-47 wide
    assert np.max(np.abs(got - af.bisectioninv(cf, u))) <= 10 * 2.0 ** -45
    assert isinstance(af.api.bisectioninv_generic(cf, 0.5), float)
    assert abs(af.api.bisectioninv_generic(cf, 0.5)) < 1e-12             # median of the Gaussian
    for numits in (0, 1, 5):
        assert np.array_equal(af.api.bisectioninv_generic(cf, u[:100], numits=numits),
                              orc.bisectioninv_generic([cf.coefficients], [-5.0, 5.0], u[:100], numits=numits,
                                                       clenshaw_fn=lambda c, x: orcc.clenshaw(c, x)))
    # genmidpoint (sample.jl:11-21), host mirror and oracle
    inf = float("inf")
    for a, b, w in ((-inf, inf, 0.0), (-inf, 3.0, -97.0), (2.0, inf, 102.0), (1.0, 4.0, 2.5)):
        assert af.api.genmidpoint(a, b) == w == orc.genmidpoint(a, b)


def test_piecewise_sample(af):
    # test/PiecewiseSampleTest.jl:5-10 (#635)
    f = af.api.abs_fun(af.Fun(np.sin, af.Chebyshev(-5, 5)))
    assert len(f.pieces) == 4 and np.allclose(f.breaks, [-5, -math.pi, 0, math.pi, 5], atol=1e-13)
    F = af.api.integrate(f)
    assert abs(F(-4.0) - (-(math.cos(-4.0) - math.cos(-5.0)))) < 1e-13
    assert abs(F(-3.0) - (-(math.cos(-math.pi) - math.cos(-5.0)) + math.cos(-3.0) - math.cos(-math.pi))) < 1e-13
    r = af.sample(f, 10)
    assert r.shape == (10,) and r.max() <= 5 and r.min() >= -5
    # parity on caller-supplied uniforms: GPU generic bisection of the piecewise CDF against the oracle, bit for bit
    cf = af.normalizedcumsum(f)
    u = RNG.random(50000)
    got = af.sample(f, u.size, uniforms=u)
    pieces = [p.coefficients for p in cf.pieces]
    want = orc.bisectioninv_generic(pieces, cf.breaks, u, clenshaw_fn=lambda c, x: orcc.clenshaw(c, x))
    assert np.array_equal(got, want)
    x = RNG.uniform(-6, 6, 10000)                                         # evaluation, including outside the domain
    assert np.array_equal(f(x), orc.piecewise_eval([p.coefficients for p in f.pieces], f.breaks, x,
                                                   clenshaw_fn=lambda c, xx: orcc.clenshaw(c, xx)))
    assert np.max(np.abs(f(x) - np.where(np.abs(x) <= 5, np.abs(np.sin(x)), 0.0))) < 1e-13
    # the samples follow |sin| / 4(ish): the CDF at the samples is uniform
    assert abs(np.mean(cf(got)) - 0.5) < 0.01


def test_sampling_2d_long_b_factors(af):
    # B factors with more than 198 coefficients: the fused conditional kernel's column tile no longer fits shared memory and
    # the library runs the reference's own sequence (sample.jl:208-213) as separate kernels; same bits as the C oracle
    rng = np.random.default_rng(5)
    NA, N, R, n = 40, 300, 3, 20000
    decay = np.exp(-0.05 * np.arange(N))[:, None]
    A = rng.standard_normal((NA, R)) * np.exp(-0.3 * np.arange(NA))[:, None]
    B = np.abs(rng.standard_normal((N, R))) * decay
    B[0] += 5.0                                                          # a positive density in x for every y
    A[0] = np.abs(A[0]) + 3.0
    A[1:] *= 0.05
    ry, u = rng.uniform(-1, 1, n), rng.random(n)
    got = af.api.lowlevel().sample2d_lowrank(A, B, -1.0, 1.0, -2.0, 3.0, ry, u)
    want = orcc.sample2d_lowrank(A, B, -1.0, 1.0, -2.0, 3.0, ry, u, threads=8)
    assert np.array_equal(got, want)
    small = af.api.lowlevel().sample2d_lowrank(A, B[:150], -1.0, 1.0, -2.0, 3.0, ry, u)      # fused kernel, same formulae
    assert np.array_equal(small, orcc.sample2d_lowrank(A, B[:150], -1.0, 1.0, -2.0, 3.0, ry, u, threads=8))


MIXED_SIZES = [36, 48, 60, 96, 100, 120, 243, 245, 343, 360, 500, 1000, 1029, 1536, 2187, 2400, 3000, 3125, 5000, 6000, 6144]


@pytest.mark.parametrize("n", MIXED_SIZES)
def test_mixed_radix_lengths(af, n):
    # n = 2^a 3^b 5^c 7^d (not a power of two): ONE shared-memory mixed-radix FFT per pair of lines (afb_mixed.cu)
    from approxfun_b200 import _lib as L
    nb = 37
    v = RNG.standard_normal((n, nb))
    for S, fwd, inv in ((af.Chebyshev(), orc.cheb_transform, orc.cheb_itransform),
                        (af.Fourier(), orc.fourier_transform, orc.fourier_itransform)):
        c = af.transform(S, v)
        assert np.max(np.abs(c - fwd(v, axis=0))) <= coef_tol(n, v), (type(S).__name__, n)
        wi = inv(v, axis=0)
        assert np.max(np.abs(af.itransform(S, v) - wi)) <= coef_tol(n, wi), (type(S).__name__, n)
        assert np.max(np.abs(af.itransform(S, c) - v)) <= coef_tol(n, v)
    z = RNG.standard_normal((n, 5)) + 1j * RNG.standard_normal((n, 5))
    assert np.max(np.abs(af.transform(af.Laurent(), z) - orc.laurent_transform(z, axis=0))) <= coef_tol(n, z)
    wi = orc.laurent_itransform(z, axis=0)
    assert np.max(np.abs(af.itransform(af.Laurent(), z) - wi)) <= coef_tol(n, wi)
    assert af.api.DevicePlan(L.CHEB1_FWD, n, 0, nb, n).launches == 1


@pytest.mark.parametrize("n", [8000, 8002, 13824, 16000, 40000, 7875, 12003, 16807])
def test_split_plan_lengths(af, n):
    # n = S h with h inside a one-CTA engine (mixed radix h <= 6912 or fused Bluestein h <= 4096), S = 2, 3, 4, 5, 7, 8, 16: one radix-S
    # pass fused into the pre kernel, S one-CTA transforms per line, de-interleave fused into the post kernel: 3 launches
    # instead of Bluestein's padded multi-kernel route (8000 / 8002 are the Padua lengths of degree 4000)
    from approxfun_b200 import _lib as L
    nb = 9
    v = RNG.standard_normal((n, nb))
    for S, fwd, inv in ((af.Chebyshev(), orc.cheb_transform, orc.cheb_itransform),
                        (af.Fourier(), orc.fourier_transform, orc.fourier_itransform)):
        c = af.transform(S, v)
        assert np.max(np.abs(c - fwd(v, axis=0))) <= coef_tol(n, v), (type(S).__name__, n)
        wi = inv(v, axis=0)
        assert np.max(np.abs(af.itransform(S, v) - wi)) <= coef_tol(n, wi), (type(S).__name__, n)
    z = RNG.standard_normal((n, 3)) + 1j * RNG.standard_normal((n, 3))
    assert np.max(np.abs(af.transform(af.Laurent(), z) - orc.laurent_transform(z, axis=0))) <= coef_tol(n, z)
    wi = orc.laurent_itransform(z, axis=0)
    assert np.max(np.abs(af.itransform(af.Laurent(), z) - wi)) <= coef_tol(n, wi)
    assert af.api.DevicePlan(L.CHEB1_FWD, n, 0, nb, n).launches == 3


@pytest.mark.parametrize("n", [4097, 5003, 10006, 40009, 70001])
def test_bluestein_beyond_one_cta(af, n):
    # non-smooth n whose padded length M exceeds one CTA: M = S x 8192 (S <= 16) runs as chirp + radix-S pass, length-8192
    # one-CTA convolutions, radix-S pass + chirp (5 launches with the real pre / post kernels); 70001 (M = 2^18) keeps the
    # padded four-step route
    from approxfun_b200 import _lib as L
    nb = 7
    v = RNG.standard_normal((n, nb))
    for S, fwd, inv in ((af.Chebyshev(), orc.cheb_transform, orc.cheb_itransform),
                        (af.Fourier(), orc.fourier_transform, orc.fourier_itransform)):
        c = af.transform(S, v)
        assert np.max(np.abs(c - fwd(v, axis=0))) <= coef_tol(n, v), (type(S).__name__, n)
        wi = inv(v, axis=0)
        assert np.max(np.abs(af.itransform(S, v) - wi)) <= coef_tol(n, wi), (type(S).__name__, n)
    z = RNG.standard_normal((n, 3)) + 1j * RNG.standard_normal((n, 3))
    assert np.max(np.abs(af.transform(af.Laurent(), z) - orc.laurent_transform(z, axis=0))) <= coef_tol(n, z)
    wi = orc.laurent_itransform(z, axis=0)
    assert np.max(np.abs(af.itransform(af.Laurent(), z) - wi)) <= coef_tol(n, wi)
    if n < 70001:
        assert af.api.DevicePlan(L.CHEB1_FWD, n, 0, nb, n).launches == 5


@pytest.mark.parametrize("n", [4001, 4002, 6913])
def test_split_plan_extension_kinds(af, n):
    # second-kind points / SinSpace: the even / odd extension (length 2(n-1) / 2n+2) through a split plan
    from approxfun_b200 import _lib as L
    v = RNG.standard_normal((n, 5))
    got = af.api.lowlevel().transform(L.CHEB2_FWD, v)
    assert np.max(np.abs(got - orc.chebkind2_transform(v, axis=0))) <= coef_tol(n, v)
    assert np.max(np.abs(af.api.lowlevel().transform(L.CHEB2_INV, got) - v)) <= 50 * coef_tol(n, v)
    s = af.api.lowlevel().transform(L.SIN_FWD, v[:n - 2])
    assert np.max(np.abs(af.api.lowlevel().transform(L.SIN_INV, s) - v[:n - 2])) <= 50 * coef_tol(n, v)


def test_registered_caller_arrays_take_the_direct_path(af):
    # afb_host_register: a caller-owned NumPy array page-locked in place is DMA-ed directly (no pinned ring); same bits as the
    # pageable route; registering twice is a CUDA error that comes back as a status, not a crash
    import ctypes
    from approxfun_b200 import _lib as L
    n, nb = 4096, 4096                              # 128 MiB
    host = RNG.standard_normal((nb, n))
    want = np.empty_like(host)
    pl = af.api.DevicePlan(L.CHEB1_FWD, n, 0, nb, n)
    pl.execute_host(host, want)                     # pageable: through the ring
    got = np.empty_like(host)
    with af.api.pinned(host, got) as h:
        pl.execute_host(host, got)
        lib = af.api.lowlevel().lib
        assert lib.afb_host_register(ctypes.c_void_p(host.ctypes.data), ctypes.c_size_t(host.nbytes)) == L.AFB_CUDA_ERROR
        assert len(h._ptrs) == 2
    assert np.array_equal(got, want)
    assert np.max(np.abs(got[:64] - orc.cheb_transform(host[:64], axis=1))) <= coef_tol(n, host)
    pl.execute_host(host, got)                      # unregistered again: back through the ring
    assert np.array_equal(got, want)
    pl.close()


def test_pageable_2d_and_padua_host_calls_through_the_pinned_ring(af):
    # the 2-D / Padua host calls move one large array each way: pageable arrays go through the ring in 32 MiB chunks (with a
    # ragged last chunk), registered ones directly -- same bits, and the oracle on a small corner
    from approxfun_b200 import _lib as L
    n1, n2 = 2048, 2306                               # 37.8 MB: one full chunk + a tail
    M = RNG.standard_normal((n2, n1))                 # column-major n1 x n2
    pl = af.api.DevicePlan(L.CHEB1_2D_FWD, n1, n2, 1, n1)
    got = np.empty_like(M)
    pl.execute_host(M, got)
    reg = np.empty_like(M)
    with af.api.pinned(M, reg):
        pl.execute_host(M, reg)
    assert np.array_equal(got, reg)
    want = orc.cheb_transform(orc.cheb_transform(M, axis=1), axis=0)
    assert np.max(np.abs(got - want)) <= coef_tol(max(n1, n2), M)
    pl.close()


def test_pageable_strided_host_arrays_through_the_pinned_ring(af):
    # a pageable caller array with stride > n, large enough (> 4 MiB, several 32 MiB chunks) for the library's own staging:
    # host copy team -> pinned ring -> DMA; the padding columns of the destination must stay untouched
    from approxfun_b200 import _lib as L
    n, stride, nb = 4096, 4096 + 6, 9000         # 295 MB
    host = RNG.standard_normal((nb, stride))
    out = np.full_like(host, -3.0)
    plan = af.api.DevicePlan(L.CHEB1_FWD, n, 0, nb, stride)
    plan.execute_host(host, out)
    idx = RNG.integers(0, nb, 64)
    assert np.max(np.abs(out[idx, :n] - orc.cheb_transform(host[idx, :n], axis=1))) <= coef_tol(n, host)
    assert np.array_equal(out[:, n:], np.full((nb, stride - n), -3.0))
    ref = np.empty((nb, n))
    plan2 = af.api.DevicePlan(L.CHEB1_FWD, n, 0, nb, n)
    plan2.execute_host(np.ascontiguousarray(host[:, :n]), ref)
    assert np.array_equal(out[:, :n], ref)
    plan.execute_host(host, host)               # in place on the host
    assert np.array_equal(host[:, :n], ref)
    plan.close()
    plan2.close()


@pytest.mark.parametrize("shape", [(5, 7), (20, 33), (64, 1024), (100, 60), (1000, 256), (4096, 2048)])
def test_mixed_tensor_spaces_transform(af, shape):
    # Fourier (x) Chebyshev and Laurent (x) Chebyshev: factor-1 plan on the columns, Chebyshev plan on the rows
    n, n2 = shape
    V = RNG.standard_normal(shape)
    S = af.Fourier() * af.Chebyshev()
    C = af.transform(S, V)
    want = orc.fourier_cheb_transform(V)
    tol = 1e-13 * np.max(np.abs(V)) * (math.log2(max(n, 2)) + math.log2(max(n2, 2)))
    assert np.max(np.abs(C - want)) <= tol
    assert np.max(np.abs(af.itransform(S, C) - V)) <= tol
    wi = orc.fourier_cheb_itransform(V)
    assert np.max(np.abs(af.itransform(S, V) - wi)) <= tol * max(1.0, np.max(np.abs(wi)))
    Z = V + 1j * RNG.standard_normal(shape)
    SL = af.Laurent() * af.Chebyshev()
    CL = af.transform(SL, Z)
    assert np.max(np.abs(CL - orc.laurent_cheb_transform(Z))) <= 2 * tol
    assert np.max(np.abs(af.itransform(SL, CL) - Z)) <= 2 * tol


def test_periodic_x_interval_fun(af):
    # test/PeriodicXIntervalTest.jl:22-24 and test/LaplaceInAStripTest.jl:5-8:
    #   d = PeriodicSegment() x ChebyshevInterval(); u_ex = Fun((x,y) -> real(cos(x + im*y)), d)
    #   u_ex(1.0, 0.1) ~ real(cos(1.0 + 0.1im))  atol = 10 eps
    S = af.Fourier() * af.Chebyshev()
    f = lambda x, y: np.cos(x) * np.cosh(y)              # real(cos(x + i y))
    u = af.Fun(f, S, (16, 24))
    assert abs(u(1.0, 0.1) - math.cos(1.0) * math.cosh(0.1)) <= 10 * EPS
    assert abs(u(0.1, 1.0) - math.cos(0.1) * math.cosh(1.0)) <= 10 * EPS       # LaplaceInAStripTest.jl:7
    assert abs(u(0.1, -1.0) - math.cos(0.1) * math.cosh(-1.0)) <= 10 * EPS
    # a periodic segment that is not the default one, many points, against direct summation of the series
    S2 = af.Fourier(-2.0, 2.0) * af.Chebyshev(0.0, 1.0)                         # PeriodicXIntervalTest.jl:8-10
    g = lambda th, t: np.exp(-20 * np.sin(math.pi * th / 4) ** 2) * (1 + t * t)
    w = af.Fun(g, S2, (101, 20))
    x, y = RNG.uniform(-2, 2, 5000), RNG.uniform(0, 1, 5000)
    got = w(x, y)
    s1, s2 = S2.factors
    want = orc.fourier_cheb_eval(w.coefficients, s1.tocanonical(x), s2.tocanonical(y))
    assert np.max(np.abs(got - want)) <= 1e-13 * 20
    assert np.max(np.abs(got - g(x, y))) <= 1e-12


def test_device_resident_fun_chains(af):
    # Fun(x -> sin(x^2), 0..10) built by the adaptive ladder ON the device (points, point-wise ops,
    # transforms, tail test and chop), f * g and exp(f) without a PCIe round trip per step
    from approxfun_b200 import device as dv
    S = af.Chebyshev(0, 10)
    dv.traffic["h2d"] = dv.traffic["d2h"] = 0
    f = dv.DeviceFun.from_function(lambda x: dv.sin(x ** 2), S)
    ladder_d2h = dv.traffic["d2h"]
    assert dv.traffic["h2d"] == 0 and ladder_d2h <= 24 * 17            # 24 bytes of scalars per rung, no samples, no coefficients
    want = orc.adaptive_fun_coefficients(lambda x: np.sin(x ** 2), 0.0, 10.0)
    host = af.Fun(lambda x: np.sin(x ** 2), S)                          # the host-array route through the same library
    c = f.coefficients                                                  # the ONE download of the result
    assert dv.traffic["d2h"] == ladder_d2h + 8 * f.n
    # the chop point sits in the rounding noise (coefficients ~ 1e-16 relative), where the device's sin and NumPy's differ by
    # an ulp: lengths agree to a few terms, coefficients to tolerance once zero-extended
    assert abs(f.n - len(want)) <= 24 and abs(f.n - len(host.coefficients)) <= 24
    m = max(f.n, len(want), len(host.coefficients))
    ext = lambda v: np.concatenate([v, np.zeros(m - len(v))])
    assert np.max(np.abs(ext(c) - ext(want))) <= 1e-13 * math.log2(256)
    assert np.max(np.abs(ext(c) - ext(host.coefficients))) <= 1e-13 * 8
    assert abs(f(0.1) - math.sin(0.01)) < 1000 * EPS                    # test/ReadmeTest.jl:19-23 on the device-built Fun
    # f * g: values on n_f + n_g - 1 points, multiplied and transformed on the device
    g = dv.DeviceFun.from_function(lambda x: dv.exp(-0.3 * x) + 2.0, S)
    before = dict(dv.traffic)
    h = f * g
    assert dv.traffic["h2d"] == before["h2d"] and dv.traffic["d2h"] - before["d2h"] == 24
    x = RNG.uniform(0, 10, 2000)
    assert np.max(np.abs(h(x) - np.sin(x ** 2) * (np.exp(-0.3 * x) + 2.0))) <= 1e-13 * 3 * 10
    hc = orc.adaptive_fun_coefficients(lambda t: np.sin(t ** 2) * (np.exp(-0.3 * t) + 2.0), 0.0, 10.0)
    assert abs(h.n - len(hc)) <= 24 and np.max(np.abs(h.coefficients[:min(h.n, len(hc))] - hc[:min(h.n, len(hc))])) <= 1e-13 * 30
    # composition exp(f) (adaptive) and linear algebra of Funs
    e = dv.exp_fun(f)
    assert np.max(np.abs(e(x) - np.exp(np.sin(x ** 2)))) <= 1e-13 * 3 * 10
    s = f + g - f
    assert np.max(np.abs(s(x) - (np.exp(-0.3 * x) + 2.0))) <= 1e-13 * 30
    # a Fun built on the host moves to the device once and back once
    hf = dv.DeviceFun.from_fun(host)
    assert np.array_equal(hf.coefficients, host.coefficients)
    # fixed-length constructor Fun(f, S, n) and values(f) round trip on the device
    f1k = dv.DeviceFun.from_function(lambda t: dv.cos(3.0 * t) * t, S, 1024)
    assert np.max(np.abs(f1k.coefficients - orc.fun_coefficients(lambda t: np.cos(3 * t) * t, 1024, 0.0, 10.0))) <= 1e-13 * 10 * 10
    v = f1k.values_device().to_host()
    xs = orc.chebpoints(1024, 0.0, 10.0)
    assert np.max(np.abs(v - np.cos(3 * xs) * xs)) <= 1e-13 * 10 * 10


@pytest.mark.parametrize("N,R,P", [(66, 8, 5000), (198, 32, 3001), (100, 50, 4097), (7, 3, 33), (64, 64, 1000)])
def test_dgemm_dmma_variant(af, N, R, P):
    # AB = CB * fA (sample.jl:210) on the FP64 tensor cores (mma.sync m8n8k4) against the FMA kernel and NumPy
    import ctypes as C
    import torch
    from approxfun_b200 import _lib as L
    lib = af.api.lowlevel().lib
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + R)
    CB = torch.randn(R, N, dtype=torch.float64, device="cuda", generator=g)        # column-major N x R
    fA = torch.randn(P, R, dtype=torch.float64, device="cuda", generator=g)        # column-major R x P
    a0 = torch.empty(P, N, dtype=torch.float64, device="cuda")
    a1 = torch.full((P, N), 7.0, dtype=torch.float64, device="cuda")
    vp = C.c_void_p
    L.check(lib, lib.afb_dgemm_nn_device(vp(CB.data_ptr()), vp(fA.data_ptr()), vp(a0.data_ptr()), N, R, P, vp(0)))
    L.check(lib, lib.afb_dgemm_nn_dmma_device(vp(CB.data_ptr()), vp(fA.data_ptr()), vp(a1.data_ptr()), N, R, P, vp(0)))
    torch.cuda.synchronize()
    want = fA.cpu().numpy() @ CB.cpu().numpy()
    tol = 1e-13 * R * float(CB.abs().max()) * float(fA.abs().max())
    assert np.max(np.abs(a0.cpu().numpy() - want)) <= tol
    assert np.max(np.abs(a1.cpu().numpy() - want)) <= tol


def test_host_point_sets_pipelined(af):
    # more than 2^22 points through the host-pointer calls: chunks pipelined over three streams (H2D | kernel | D2H);
    # results equal those of small calls, and the oracle on a subsample
    ll = af.api.lowlevel()
    P = 3 * (1 << 22) + 17
    rng = np.random.default_rng(77)
    c = rng.standard_normal(50) / (1 + np.arange(50))
    x = rng.uniform(-1, 1, P)
    y = ll.clenshaw(c, x)
    idx = np.concatenate([np.arange(0, 3000), np.arange((1 << 22) - 1500, (1 << 22) + 1500), np.arange(P - 3000, P)])
    assert np.array_equal(y[idx], orcc.clenshaw(c, x[idx]))
    assert np.array_equal(y[5_000_000:5_010_000], ll.clenshaw(c, x[5_000_000:5_010_000]))
    f = af.Fun(lambda t: np.exp(-t * t), af.Chebyshev(-5, 5))
    cdf = af.normalizedcumsum(f).coefficients
    u = rng.random(P)
    s = ll.samplecdf(cdf, -5.0, 5.0, u)
    assert np.array_equal(s[idx], orc.fromcanonical(-5.0, 5.0, orcc.cheb_bisect(cdf, u[idx], threads=8)))
    A, Bv, Cv = af.Legendre().rec(50)
    yr = ll.clenshaw_rec(c, A, Bv, Cv, x)
    assert np.array_equal(yr[idx], orcc.clenshaw_rec(c, A, Bv, Cv, x[idx]))


def test_cluster_line_kernel_variant(af):
    # the 2-CTA cluster / distributed-shared-memory kernel for n = 16384 lines (AFB_LINE_CLUSTER=1; measured slower than the
    # persistent TMA-staged kernel and therefore off by default): same results
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, sys\n"
        "import approxfun_b200 as af\n"
        "from approxfun_b200 import _lib as L\n"
        "v = np.random.default_rng(9).standard_normal((16384, 151))\n"
        "np.save(sys.argv[1], af.api.lowlevel().transform(L.CHEB1_FWD, v))\n")
    res = []
    for flag in ("0", "1"):
        path = f"/tmp/afb_cluster_{flag}.npy"
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, AFB_LINE_CLUSTER=flag), cwd=root)
        res.append(np.load(path))
    v = np.random.default_rng(9).standard_normal((16384, 151))
    assert np.max(np.abs(res[1] - orc.cheb_transform(v, axis=0))) <= coef_tol(16384, v)
    assert np.array_equal(res[0], res[1])       # same radix schedule and twiddle tables: identical bits


def test_short_line_tile_kernel_variant(af):
    # line_kernel_tile for n <= 256 (AFB_LINE_TILE=1: the CTA's lines move as one coalesced tile through shared memory; measured
    # slower than the plain kernel -- n = 128: 0.53 against 0.755 of HBM -- and therefore off by default): same bits
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, sys\n"
        "import approxfun_b200 as af\n"
        "from approxfun_b200 import _lib as L\n"
        "rng = np.random.default_rng(12)\n"
        "out = []\n"
        "for n in (64, 128, 256):\n"
        "    v = rng.standard_normal((n, 1003))\n"
        "    for kind in (L.CHEB1_FWD, L.CHEB1_INV, L.FOURIER_FWD, L.FOURIER_INV):\n"
        "        out.append(af.api.lowlevel().transform(kind, v).ravel())\n"
        "np.save(sys.argv[1], np.concatenate(out))\n")
    res = []
    for flag in ("0", "1"):
        path = f"/tmp/afb_tile_{flag}.npy"
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, AFB_LINE_TILE=flag), cwd=root)
        res.append(np.load(path))
    assert np.array_equal(res[0], res[1])
    v = np.random.default_rng(12).standard_normal((64, 1003))
    assert np.max(np.abs(res[1][:64 * 1003].reshape(64, 1003) - orc.cheb_transform(v, axis=0))) <= coef_tol(64, v)
