"""Pins the CPU oracle to every golden value / known-answer test the reference holds for the hot
path.  CPU only."""
import math

import numpy as np
import pytest
from scipy import special

from oracle import approxfun_oracle as orc
from oracle import oracle_c as orcc

EPS = np.finfo(np.float64).eps


def test_points_descending_and_formula():
    # NEWS.md:92 (descending order); UPSTREAM chebyshevpoints(T,n,Val(1))
    for n in (1, 2, 5, 16, 4096):
        x = orc.chebpoints_canonical(n)
        assert x.shape == (n,)
        assert np.all(np.diff(x) < 0) or n == 1
        k = np.arange(n)
        assert np.allclose(x, np.cos(np.pi * (2 * k + 1) / (2 * n)), rtol=0, atol=2 * EPS)
        assert np.array_equal(x, -x[::-1])  # exact antisymmetry of sinpi form
    p = orc.chebpoints(20, 0.0, 10.0)
    assert p[0] > p[-1] and 0 < p[-1] and p[0] < 10


def test_fun_exp_coefficients_bessel():
    # Fun(exp, Chebyshev()): c0 = I0(1), ck = 2 Ik(1)  (README.md:30-35 usage)
    c = orc.fun_coefficients(np.exp, 32)
    assert abs(c[0] - special.iv(0, 1.0)) <= 2 * EPS
    for k in range(1, 12):
        assert abs(c[k] - 2 * special.iv(k, 1.0)) <= 4 * EPS


def test_readme_sin_x2():
    # test/ReadmeTest.jl:19-23: x = Fun(identity,0..10); f = sin(x^2); f(.1) ~ sin(.01) atol 1000eps
    c = orc.adaptive_fun_coefficients(lambda x: np.sin(x * x), 0.0, 10.0)
    val = orcc.clenshaw(c, orc.tocanonical(0.0, 10.0, np.array([0.1])))[0]
    assert abs(val - math.sin(0.01)) < 1000 * EPS
    # config 1 shape: fixed n = 2^20 transform + itransform round trip
    n = 2 ** 20
    v = np.sin(orc.chebpoints(n, 0.0, 10.0) ** 2)
    cc = orc.cheb_transform(v)
    assert np.max(np.abs(orc.cheb_itransform(cc) - v)) < 1e-13 * 20


def test_extras_roundtrip():
    # test/ExtrasTest.jl:56-61: transform(space(f), values(f)) ~ coefficients(f) to 10eps
    c = orc.adaptive_fun_coefficients(lambda x: np.exp(x) * np.cos(3 * x))
    v = orc.cheb_itransform(c)
    assert np.max(np.abs(orc.cheb_transform(v) - c)) < 10 * EPS


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 17, 64, 100])
def test_cheb_transform_vs_direct_sum(n):
    rng = np.random.default_rng(n)
    v = rng.standard_normal(n)
    c = orc.cheb_transform(v)
    assert np.max(np.abs(c - orc.cheb_transform_direct(v))) < 1e-13 * np.max(np.abs(v)) * max(1, math.log2(n))
    vv = orc.cheb_itransform(c)
    assert np.max(np.abs(vv - v)) < 1e-13 * np.max(np.abs(v)) * max(1, math.log2(n))
    assert np.max(np.abs(orc.cheb_itransform_direct(c) - v)) < 1e-12


def test_fourier_doctest_ordering():
    # docs/src/usage/spaces.md:94-98: Fun(Fourier(),[1,2,3,4])(0.1) = 1+2sin(.1)+3cos(.1)+4sin(.2)
    # on a grid wide enough that slot 4 is a genuine sine slot (n = 5)
    c = np.array([1.0, 2.0, 3.0, 4.0, 0.0])
    th = orc.fourierpoints(5)
    want = 1 + 2 * np.sin(th) + 3 * np.cos(th) + 4 * np.sin(2 * th)
    assert np.allclose(orc.fourier_itransform(c), want, atol=20 * EPS)
    assert np.allclose(orc.fourier_transform(want), c, atol=20 * EPS)
    assert np.allclose(orc.fourier_itransform_generic(c), want, atol=20 * EPS)
    assert np.allclose(orc.fourier_transform_generic(want), c, atol=20 * EPS)


def test_fourier_even_n_nyquist_rule():
    # src/Extras/fftGeneric.jl:53: even n -> last slot = -Re F_{n/2}/n  (5cos(10 theta), n=20 -> -5)
    th = orc.fourierpoints(20)
    c = orc.fourier_transform(5 * np.cos(10 * th))
    assert abs(c[-1] + 5) < 50 * EPS and np.max(np.abs(c[:-1])) < 50 * EPS
    assert np.allclose(orc.fourier_itransform(c), 5 * np.cos(10 * th), atol=100 * EPS)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8, 20, 33, 128])
def test_fourier_batched_equals_generic(n):
    # test/NumberTypeTest.jl:32-45 pins the Float64 path to fftGeneric.jl:47-64 (restated verbatim)
    rng = np.random.default_rng(100 + n)
    v = rng.standard_normal((n, 3))
    got = orc.fourier_transform(v, axis=0)
    for b in range(3):
        ref = orc.fourier_transform_generic(v[:, b])
        assert np.max(np.abs(got[:, b] - ref)) < 10 * EPS * max(1.0, np.max(np.abs(ref))) * 4
        back = orc.fourier_itransform_generic(ref)
        assert np.max(np.abs(back - v[:, b])) < 200 * EPS
    assert np.max(np.abs(orc.fourier_itransform(got, axis=0) - v)) < 200 * EPS


def test_fourier_exp_cos_numbertype():
    # test/NumberTypeTest.jl:32-45: f = exp(cos(theta)) on 0..2pi: a_k = 2 I_k(1), a_0 = I_0(1), b_k = 0
    n = 64
    th = orc.fourierpoints(n)
    c = orc.fourier_transform(np.exp(np.cos(th)))
    assert abs(c[0] - special.iv(0, 1.0)) < 10 * EPS
    for k in range(1, 15):
        assert abs(c[2 * k] - 2 * special.iv(k, 1.0)) < 10 * EPS
        assert abs(c[2 * k - 1]) < 10 * EPS


def test_laurent_doctest_ordering():
    # docs/src/usage/spaces.md:121-125: [1,2,3,4] <-> 1 + 2e^{-i t} + 3e^{i t} + 4e^{-2i t}
    c = np.array([1, 2, 3, 4, 0], dtype=np.complex128)
    th = orc.fourierpoints(5)
    want = 1 + 2 * np.exp(-1j * th) + 3 * np.exp(1j * th) + 4 * np.exp(-2j * th)
    assert np.allclose(orc.laurent_itransform(c), want, atol=50 * EPS)
    assert np.allclose(orc.laurent_transform(want), c, atol=50 * EPS)
    assert list(orc.laurent_index_map(6)) == [0, 5, 1, 4, 2, 3]


def test_fft_sign_convention():
    # test/NumberTypeTest.jl:99-105 + src/Extras/fftBigFloat.jl:14: F_k = sum v_j cis(-2 pi jk/n)
    for n in range(1, 11):
        v = np.arange(1, n + 1, dtype=np.float64)
        j = np.arange(n)
        F = np.array([np.sum(v * np.exp(-2j * np.pi * j * k / n)) for k in range(n)])
        got = orc.laurent_transform(v) * n
        assert np.allclose(got, F[orc.laurent_index_map(n)], rtol=1e-8)


def test_sampling_golden_2d_marginal():
    # test/SamplingTest.jl:5-9: f=(x-y)^2 exp(-x^2/2-y^2/2) on (-4..4)^2;
    # g = sum(f,1)/sum(f);  g(0.1) ~ 0.2004758624973169
    n = 96
    a, b = -4.0, 4.0
    x = orc.chebpoints(n, a, b)
    V = (x[:, None] - x[None, :]) ** 2 * np.exp(-x[:, None] ** 2 / 2 - x[None, :] ** 2 / 2)
    Cm = orc.cheb2_transform(V)
    w = orc.cheb_definite_integrals(n) * (b - a) / 2
    marg = w @ Cm              # integrate over the first variable -> Chebyshev series in y
    total = float(marg @ w)
    g01 = orcc.clenshaw(marg, orc.tocanonical(a, b, np.array([0.1])))[0] / total
    assert abs(g01 - 0.2004758624973169) < 4 * EPS
    assert np.max(np.abs(orc.cheb2_itransform(Cm) - V)) < 1e-13


def test_sampling_golden_bisection():
    # test/SamplingTest.jl:10-12: g = cumsum(Fun(exp(-x^2/2), -5..5)); g(bisectioninv(g,0.5)) ~ 0.5
    a, b = -5.0, 5.0
    c = orc.adaptive_fun_coefficients(lambda x: np.exp(-x * x / 2), a, b)
    g = orc.cheb_cumsum(c, a, b)
    assert abs(orcc.clenshaw(g, np.array([-1.0]))[0]) < 1e-15
    assert abs(orcc.clenshaw(g, np.array([1.0]))[0] - math.sqrt(2 * math.pi) * math.erf(5 / math.sqrt(2))) < 1e-14
    xs = orcc.cheb_bisect(g, np.array([0.5]))
    val = orcc.clenshaw(g, xs)[0]
    assert abs(val - 0.5) < 1e-12       # Julia's default isapprox rtol = sqrt(eps)
    xs_np = orc.chebbisectioninv(g, np.array([0.5]))
    assert abs(xs_np[0] - xs[0]) < 1e-13


def test_sample_distribution_gaussian():
    # examples/Sampling1.jl / config 5 shape: samples of exp(-x^2) on -5..5 follow the truncated normal
    a, b = -5.0, 5.0
    c = orc.adaptive_fun_coefficients(lambda x: np.exp(-x * x), a, b)
    assert 40 < c.size < 100            # N ~ 65
    cf = orc.normalizedcumsum(c, a, b)
    u = np.random.default_rng(0).random(20000)
    s = orc.fromcanonical(a, b, orcc.cheb_bisect(cf, u))
    assert s.min() >= a and s.max() <= b
    # exact CDF of N(0, 1/2) (truncation at +-5 is below 1e-12): F(sample) must reproduce u
    assert np.max(np.abs((1 + special.erf(s)) / 2 - u)) < 1e-12


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(7)
    for N in (0, 1, 2, 3, 10, 65, 300):
        c = rng.standard_normal(N) / (1 + np.arange(N)) ** 2
        x = rng.uniform(-1, 1, 257)
        x[:2] = [-1.0, 1.0]
        y_np = orc.clenshaw(c, x)
        y_c = orcc.clenshaw(c, x)
        assert np.max(np.abs(y_np - y_c)) <= 1e-13 * max(1.0, np.max(np.abs(y_np)) if N else 0)
        # reference loop order (k outer, i inner) is bit-identical to the point-outer order
        assert np.array_equal(orcc.clenshaw_planned(c, x), y_c)
    Cm = rng.standard_normal((12, 40))
    x = rng.uniform(-1, 1, 40)
    assert np.max(np.abs(orc.clenshaw_cols(Cm, x) - orcc.clenshaw_cols(Cm, x))) < 1e-13
    assert np.max(np.abs(orc.clenshaw_multi(Cm[:, :5], x) - orcc.clenshaw_multi(Cm[:, :5], x))) < 1e-13


def test_c_oracle_normcumsum_and_bisect():
    rng = np.random.default_rng(11)
    for N in (1, 2, 3, 4, 9, 66):
        c = np.abs(rng.standard_normal(N)) / (1 + np.arange(N)) ** 2
        c[0] += 2.0
        got = orcc.cheb_normcumsum(c)
        want = orc.chebnormalizedcumsum(c)
        assert got.shape == (N + 1,)
        assert np.max(np.abs(got - want)) < 1e-15
        assert abs(orcc.clenshaw(got, np.array([1.0]))[0] - 1) < 1e-14
        assert abs(orcc.clenshaw(got, np.array([-1.0]))[0]) < 1e-14
    M = np.abs(rng.standard_normal((9, 6))) / (1 + np.arange(9))[:, None] ** 2
    M[0] += 2
    gm = orcc.cheb_normcumsum(M)
    assert gm.shape == M.shape
    assert np.max(np.abs(gm - orc.chebnormalizedcumsum(M))) < 1e-15
    u = rng.random(6)
    r1 = orcc.cheb_bisect_cols(gm, u)
    r2 = orc.chebbisectioninv_cols(gm, u)
    assert np.max(np.abs(r1 - r2)) < 1e-12
    cf = orcc.cheb_normcumsum(np.array([1.0, 0.0, -0.5]))
    u = rng.random(100)
    assert np.array_equal(orcc.cheb_bisect(cf, u), orcc.cheb_bisect_planned(cf, u))
    assert np.max(np.abs(orcc.cheb_bisect(cf, u) - orc.chebbisectioninv(cf, u))) < 1e-12


def test_tensor_interlace_order():
    # docs/src/usage/constructors.md:98-116
    Cm = np.array([[1.0, 2.0], [3.0, 0.0]])
    assert list(orc.tensor_interlace(Cm)[:3]) == [1.0, 2.0, 3.0]


def test_padua_oracle_against_collocation_and_documented_order():
    # SURVEY 8f rank 1.  The DCT-based restatement must reproduce the unique interpolant (N x N collocation solve), and
    # the flat coefficient order must be the documented one: T0T0, T0(x)T1(y), T1(x)T0(y), ... (constructors.md:98-116).
    for d in (1, 2, 3, 4, 7, 12):
        P = orc.paduapoints(d)
        N = (d + 1) * (d + 2) // 2
        assert P.shape == (N, 2) and len({(round(a, 12), round(b, 12)) for a, b in P}) == N
        ix, iy = orc.padua_grid_index(d)
        assert np.all((ix + iy) % 2 == (d + 1) % 2)            # the checkerboard subset of the Lobatto tensor grid
        v = np.exp(0.3 * P[:, 0] - 0.7 * P[:, 1]) + P[:, 0] * P[:, 1] ** 2
        c = orc.padua_transform(v)
        assert np.max(np.abs(c - orc.padua_transform_direct(v))) < 1e-14
        assert np.max(np.abs(orc.padua_itransform(c) - v)) < 1e-14
    P = orc.paduapoints(1)
    assert np.allclose(orc.padua_itransform(np.array([1.0, 2.0, 3.0])), 1 + 2 * P[:, 1] + 3 * P[:, 0])
    # published Padua set of degree 2 (Caliari-De Marchi-Vianello): x in {1, 0, -1}, y in {1, 1/2, -1/2, -1}
    assert np.allclose(sorted(map(tuple, np.round(orc.paduapoints(2), 12))),
                       sorted([(1, .5), (1, -1), (0, 1), (0, -.5), (-1, .5), (-1, -1)]))
    assert [orc.padua_degree(N) for N in (1, 2, 3, 4, 6, 7, 10, 32, 36, 37)] == [0, 1, 1, 2, 2, 3, 3, 7, 7, 8]
    with pytest.raises(ValueError):
        orc.padua_transform(np.zeros(5))


def test_three_term_recurrence_tables_against_scipy():
    # recA/recB/recC restated from DLMF 18.9.1-2, checked on scipy's Jacobi / Gegenbauer / Legendre
    import scipy.special as sps
    from oracle import oracle_c as orcc
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 300)
    N = 40
    c = rng.standard_normal(N) / (1 + np.arange(N))
    for a, b in [(0.0, 0.0), (0.5, -0.5), (1.5, 2.0), (0.0, 1.0)]:
        A, B, C = orc.jacobi_rec(a, b, N)
        want = sum(c[n] * sps.eval_jacobi(n, a, b, x) for n in range(N))
        tol = 1e-13 * max(1.0, np.max(np.abs(want)))   # evaluations: 1e-13 ||f||_inf (BASELINE.json tolerance)
        assert np.max(np.abs(orc.clenshaw_rec(c, A, B, C, x) - want)) < tol
        assert np.max(np.abs(orcc.clenshaw_rec(c, A, B, C, x) - want)) < tol
    for lam in (0.5, 1.0, 2.0):
        A, B, C = orc.ultraspherical_rec(lam, N)
        want = sum(c[n] * sps.eval_gegenbauer(n, lam, x) for n in range(N))
        assert np.max(np.abs(orcc.clenshaw_rec(c, A, B, C, x) - want)) < 1e-13 * max(1.0, np.max(np.abs(want)))
    # the Chebyshev table through the general recurrence is the first-kind series of the hot path
    A, B, C = orc.chebyshev_rec(N)
    assert np.max(np.abs(orcc.clenshaw_rec(c, A, B, C, x) - orcc.clenshaw(c, x))) < 1e-14
