"""CPU-side unit tests of the CUDA kernel sources through the fiber emulator (tests/emu).

The emulator compiles the very same .cu files with g++ and runs every CUDA thread as a fiber; it is
a debugging aid for index arithmetic (this container has no GPU) and is not part of the product:
the shipped library is nvcc-built only and the `-m gpu` tests are the parity tests proper."""
import math

import numpy as np
import pytest

from oracle import approxfun_oracle as orc
from oracle import oracle_c as orcc
from tests.emu_harness import emu_lowlevel

B, ll = emu_lowlevel()
RNG = np.random.default_rng(1234)


def tol(n, scale):
    return 1e-13 * scale * max(1.0, math.log2(max(n, 2)))


REAL_KINDS = [
    (B.CHEB1_FWD, lambda v: orc.cheb_transform(v, axis=0)),
    (B.CHEB1_INV, lambda v: orc.cheb_itransform(v, axis=0)),
    (B.FOURIER_FWD, lambda v: orc.fourier_transform(v, axis=0)),
    (B.FOURIER_INV, lambda v: orc.fourier_itransform(v, axis=0)),
]
CPLX_KINDS = [
    (B.LAURENT_FWD, lambda v: orc.laurent_transform(v, axis=0)),
    (B.LAURENT_INV, lambda v: orc.laurent_itransform(v, axis=0)),
]

# direct (<=32), line (pow2 64..16384), generic Bluestein (odd sizes), generic four-step (2^15)
SIZES = [1, 2, 3, 4, 5, 7, 8, 16, 17, 31, 32, 33, 50, 64, 100, 127, 128, 256, 500, 512, 1024, 2048, 4096, 8192,
         16384, 32768]


@pytest.mark.parametrize("n", SIZES)
def test_real_transforms(n):
    nb = 3 if n <= 4096 else 1
    v = RNG.standard_normal((n, nb))
    for kind, ref in REAL_KINDS:
        want = ref(v)
        got = ll.transform(kind, v)
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(v)))), (kind, n)


@pytest.mark.parametrize("n", SIZES[:-2] + [16384])
def test_complex_transforms(n):
    nb = 2 if n <= 4096 else 1
    v = RNG.standard_normal((n, nb)) + 1j * RNG.standard_normal((n, nb))
    for kind, ref in CPLX_KINDS:
        want = ref(v)
        got = ll.transform(kind, v)
        assert np.max(np.abs(got - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(v)))), (kind, n)


@pytest.mark.parametrize("n,nb", [(16384, 11), (8192, 19)])
def test_long_lines_staged_persistent_kernel(n, nb):
    # n = 16384 with more lines than (emulated, 4) SMs: persistent CTAs, next line staged by a bulk copy (3/4 of the line,
    # the rest read from global memory); n = 8192 takes the plain kernel
    v = RNG.standard_normal((n, nb))
    got = ll.transform(B.CHEB1_FWD, v)
    want = orc.cheb_transform(v, axis=0)
    assert np.max(np.abs(got - want)) <= tol(n, np.max(np.abs(v)))
    back = ll.transform(B.CHEB1_INV, got)
    assert np.max(np.abs(back - v)) <= 50 * tol(n, np.max(np.abs(v)))
    assert np.max(np.abs(ll.transform(B.CHEB1_INV, v) - orc.cheb_itransform(v, axis=0))) <= tol(n, 200 * np.max(np.abs(v)))
    fo = ll.transform(B.FOURIER_FWD, v)
    assert np.max(np.abs(fo - orc.fourier_transform(v, axis=0))) <= tol(n, np.max(np.abs(v)))
    assert np.max(np.abs(ll.transform(B.FOURIER_INV, fo) - v)) <= 50 * tol(n, np.max(np.abs(v)))
    assert np.max(np.abs(ll.transform(B.FOURIER_INV, v) - orc.fourier_itransform(v, axis=0))) <= tol(n, 200 * np.max(np.abs(v)))


def test_empty_and_vector_shapes():
    assert ll.transform(B.CHEB1_FWD, np.zeros(0)).shape == (0,)
    assert ll.transform(B.CHEB1_FWD, np.zeros((8, 0))).shape == (8, 0)
    v = RNG.standard_normal(64)
    assert np.allclose(ll.transform(B.CHEB1_FWD, v), orc.cheb_transform(v), atol=1e-14)


def test_batch_not_multiple_of_lines_per_cta():
    v = RNG.standard_normal((64, 7))  # 4 lines per CTA for N=32
    assert np.allclose(ll.transform(B.CHEB1_FWD, v), orc.cheb_transform(v, axis=0), atol=1e-14)
    assert np.allclose(ll.transform(B.FOURIER_INV, v), orc.fourier_itransform(v, axis=0), atol=1e-13)


def test_strided_plan_and_inplace():
    n, nb, stride = 128, 5, 160
    buf = RNG.standard_normal((nb, stride))
    want = orc.cheb_transform(buf[:, :n].T, axis=0).T
    h = ll.plan_create(B.CHEB1_FWD, n, 0, nb, stride)
    work = buf.copy()
    ll.plan_execute_host(h, work, work)
    ll.plan_destroy(h)
    assert np.max(np.abs(work[:, :n] - want)) < 1e-14
    assert np.array_equal(work[:, n:], buf[:, n:])  # gaps untouched


@pytest.mark.parametrize("shape", [(8, 8), (64, 128), (100, 36), (256, 64)])
def test_tensor2d(shape):
    V = RNG.standard_normal(shape)
    got = ll.transform2d(B.CHEB1_2D_FWD, V)
    want = orc.cheb2_transform(V)
    assert np.max(np.abs(got - want)) < 1e-13
    back = ll.transform2d(B.CHEB1_2D_INV, got)
    assert np.max(np.abs(back - V)) < 1e-12


@pytest.mark.parametrize("shape", [(2, 1024), (6, 2048), (34, 1024), (10, 4096), (2, 16384)])
def test_tensor2d_row_pairs_four_step(shape):
    # even n, n2 = 2^10..2^18: the row pass runs on row pairs without transposes; flag 1 = transpose route
    V = RNG.standard_normal(shape)
    want = orc.cheb2_transform(V)
    got = ll.transform2d(B.CHEB1_2D_FWD, V)
    assert np.max(np.abs(got - want)) < 1e-14
    assert np.max(np.abs(ll.transform2d(B.CHEB1_2D_FWD, V, B.PLAN_2D_TRANSPOSE) - want)) < 1e-14
    assert np.max(np.abs(ll.transform2d(B.CHEB1_2D_INV, want) - V)) < 1e-13
    assert np.max(np.abs(ll.transform2d(B.CHEB1_2D_INV, want, B.PLAN_2D_TRANSPOSE) - V)) < 1e-13


@pytest.mark.parametrize("d", [0, 1, 2, 5, 8, 31, 63])
def test_padua(d):
    N = (d + 1) * (d + 2) // 2
    P, Po = ll.paduapoints(N), orc.paduapoints(d)
    assert np.max(np.abs(P - Po)) <= 5e-16
    assert ll.paduapoints(max(N - 1, 1)).shape[0] == (N if d > 0 and N - 1 > d * (d + 1) // 2 else max(N - 1, 1))
    v = RNG.standard_normal(N)
    c = ll.padua(B.PADUA_FWD, v)
    assert np.max(np.abs(c - orc.padua_transform(v))) < 1e-13
    assert np.max(np.abs(ll.padua(B.PADUA_INV, c) - v)) < 1e-12
    with pytest.raises(Exception):
        ll.padua(B.PADUA_FWD, np.zeros(N + 1))


def test_batched_four_step_under_bluestein():
    # odd lengths above 4096: Bluestein pads to M >= 2^14, whose sub-FFT is the four-step run on whole chunks of lines
    v = RNG.standard_normal((4500, 3))
    assert np.max(np.abs(ll.transform(B.CHEB2_FWD, v) - orc.chebkind2_transform(v, axis=0))) < 1e-14
    v = RNG.standard_normal((9001, 2))
    assert np.max(np.abs(ll.transform(B.CHEB1_FWD, v) - orc.cheb_transform(v, axis=0))) < 1e-14
    z = RNG.standard_normal((32768, 2)) + 1j * RNG.standard_normal((32768, 2))
    assert np.max(np.abs(ll.transform(B.LAURENT_FWD, z) - orc.laurent_transform(z, axis=0))) < 1e-14


def test_general_three_term_clenshaw():
    from oracle import oracle_c as orcc
    x = RNG.uniform(-1, 1, 1500)
    for N in (0, 1, 2, 40, 3100):
        c = RNG.standard_normal(N)
        for A, Bv, Cv in (orc.jacobi_rec(0.5, 1.5, N), orc.ultraspherical_rec(2.0, N)):
            assert np.array_equal(ll.clenshaw_rec(c, A, Bv, Cv, x), orcc.clenshaw_rec(c, A, Bv, Cv, x))


def test_points():
    for n in (1, 2, 5, 64, 1000):
        x = ll.chebpoints(n)
        want = orc.chebpoints_canonical(n)
        assert np.all(np.diff(x) < 0) or n == 1
        assert np.max(np.abs(x - want)) <= 2.3e-16
        assert np.array_equal(ll.chebpoints(n, 0.0, 10.0), orc.fromcanonical(0.0, 10.0, x))
    assert np.allclose(ll.fourierpoints(8), orc.fourierpoints(8), atol=0)


@pytest.mark.parametrize("N", [0, 1, 2, 3, 4, 5, 66, 1000, 12288, 12289, 30001])
def test_clenshaw_bit_exact(N):
    c = RNG.standard_normal(N) / (1 + np.arange(N)) ** 1.5
    P = 700 if N < 10000 else 9
    x = RNG.uniform(-1, 1, P)
    x[:2] = [-1.0, 1.0]
    assert np.array_equal(ll.clenshaw(c, x), orcc.clenshaw(c, x))


def test_clenshaw_tile_boundaries():
    c = RNG.standard_normal(9)
    for P in (0, 1, 511, 512, 2047, 2048, 2049, 5000):
        x = RNG.uniform(-1, 1, P)
        assert np.array_equal(ll.clenshaw(c, x), orcc.clenshaw(c, x))


def test_clenshaw_cols_and_multi():
    for N, P in [(0, 5), (1, 5), (7, 130), (300, 3)]:
        Cm = RNG.standard_normal((N, P))
        x = RNG.uniform(-1, 1, P)
        assert np.array_equal(ll.clenshaw_cols(Cm, x), orcc.clenshaw_cols(Cm, x) if N else np.zeros(P))
    Cm = RNG.standard_normal((9, 4))
    x = RNG.uniform(-1, 1, 300)
    assert np.array_equal(ll.clenshaw_multi(Cm, x), orcc.clenshaw_multi(Cm, x))
    with pytest.raises(B.DimensionMismatch):
        ll.clenshaw_cols(RNG.standard_normal((4, 6)), np.zeros(5))  # src/Extras/sample.jl:85


def test_normcumsum_and_bisection_bit_exact():
    for N in (1, 2, 3, 4, 9, 66):
        c = np.abs(RNG.standard_normal(N)) / (1 + np.arange(N)) ** 2
        c[0] += 2.0
        cf = ll.cheb_normcumsum(c)
        assert np.array_equal(cf, orcc.cheb_normcumsum(c))
        u = RNG.random(600)
        assert np.array_equal(ll.cheb_bisect(cf, u), orcc.cheb_bisect(cf, u))
        assert np.array_equal(ll.samplecdf(cf, -5.0, 5.0, u), orc.fromcanonical(-5.0, 5.0, orcc.cheb_bisect(cf, u)))
    M = np.abs(RNG.standard_normal((9, 140))) / (1 + np.arange(9))[:, None] ** 2
    M[0] += 2
    gm = ll.cheb_normcumsum(M)
    assert np.array_equal(gm, orcc.cheb_normcumsum(M))
    u = RNG.random(140)
    assert np.array_equal(ll.cheb_bisect_cols(gm, u), orcc.cheb_bisect_cols(gm, u))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 17, 33, 64, 65, 100, 129, 1025])
def test_second_kind_and_sinspace(n):
    v3 = RNG.standard_normal((n, 3))   # odd batch: two lines share one complex FFT, the last one rides alone
    assert np.max(np.abs(ll.transform(B.CHEB2_FWD, v3) - orc.chebkind2_transform(v3, axis=0))) <= tol(n, np.max(np.abs(v3)))
    s3 = ll.transform(B.SIN_INV, v3)
    for b in range(3):
        assert np.max(np.abs(s3[:, b] - orc.sin_itransform_generic(v3[:, b]))) <= tol(n, max(1.0, np.max(np.abs(s3))))
    v = RNG.standard_normal((n, 2))
    got = ll.transform(B.CHEB2_FWD, v)
    assert np.max(np.abs(got - orc.chebkind2_transform(v, axis=0))) <= tol(n, np.max(np.abs(v)))
    back = ll.transform(B.CHEB2_INV, got)
    assert np.max(np.abs(back - v)) <= tol(n, np.max(np.abs(v))) * 4
    s = ll.transform(B.SIN_FWD, v)
    for b in range(2):
        assert np.max(np.abs(s[:, b] - orc.sin_transform_generic(v[:, b]))) <= tol(n, np.max(np.abs(v)))
    si = ll.transform(B.SIN_INV, v)
    assert np.max(np.abs(si[:, 0] - orc.sin_itransform_generic(v[:, 0]))) <= tol(n, np.max(np.abs(si)))
    # definitions: values of sum_k c_k T_k at the second-kind points, and of a sine series on its grid
    if n >= 2:
        c = RNG.standard_normal(n)
        x = orc.cheb2points_canonical(n)
        assert np.max(np.abs(ll.transform(B.CHEB2_INV, c) - orc.clenshaw(c, x))) < 1e-12 * n
    th = orc.sinpoints(n)
    f = RNG.standard_normal(n)
    vals = sum(f[k] * np.sin((k + 1) * th) for k in range(n))
    assert np.max(np.abs(ll.transform(B.SIN_FWD, vals) - f)) < 1e-12 * n


@pytest.mark.parametrize("n", [1, 2, 5, 32, 64, 100, 1024])
def test_taylor_and_hardy(n):
    # docs/src/usage/spaces.md:100-110: Taylor f = sum f_k e^{ikt}; Hardy{false} f = sum f_k e^{-i(k+1)t}; t_j = 2 pi j/n
    v = RNG.standard_normal((n, 2)) + 1j * RNG.standard_normal((n, 2))
    F = np.fft.fft(v, axis=0) / n
    assert np.max(np.abs(ll.transform(B.TAYLOR_FWD, v) - F)) <= tol(n, np.max(np.abs(v)))
    assert np.max(np.abs(ll.transform(B.HARDYM_FWD, v) - F[::-1])) <= tol(n, np.max(np.abs(v)))
    th = orc.fourierpoints(n)
    f = RNG.standard_normal(n) + 1j * RNG.standard_normal(n)
    k = np.arange(n)
    vt = (f[None, :] * np.exp(1j * np.outer(th, k))).sum(axis=1)
    vh = (f[None, :] * np.exp(-1j * np.outer(th, k + 1))).sum(axis=1)
    assert np.max(np.abs(ll.transform(B.TAYLOR_INV, f) - vt)) < 1e-12 * n
    assert np.max(np.abs(ll.transform(B.HARDYM_INV, f) - vh)) < 1e-12 * n
    assert np.max(np.abs(ll.transform(B.HARDYM_FWD, vh) - f)) < 1e-12 * n


def test_bivariate_clenshaw_bit_exact():
    for n1, n2, P in [(1, 1, 5), (5, 1, 40), (1, 6, 40), (9, 7, 1500), (40, 33, 300), (130, 100, 9)]:
        Cm = RNG.standard_normal((n1, n2)) / (1 + np.arange(n1))[:, None]
        x = RNG.uniform(-1, 1, P)
        y = RNG.uniform(-1, 1, P)
        got = ll.clenshaw2d(Cm, x, y)
        assert np.array_equal(got, orcc.clenshaw2d(Cm, x, y))
        want = np.polynomial.chebyshev.chebval2d(x, y, Cm)
        assert np.max(np.abs(got - want)) < 1e-12 * max(1.0, np.sum(np.abs(Cm)))


def test_sample2d_fused_kernel_bit_exact():
    # src/Extras/sample.jl:208-213 as one kernel: evaluate.(f.A, ry'), CB*fA, chebnormalizedcumsum!, chebbisectioninv
    for NA, N, R, P in [(12, 9, 3, 300), (1, 1, 1, 5), (30, 66, 7, 140), (5, 2, 2, 130)]:
        A = RNG.standard_normal((NA, R)) / (1 + np.arange(NA))[:, None]
        CB = np.abs(RNG.standard_normal((N, R))) / (1 + np.arange(N))[:, None] ** 2
        CB[0] += 3.0
        A[0] = np.abs(A[0]) + 2.0            # positive weights keep every conditional density positive
        ry = RNG.uniform(-4, 4, P)
        ry[0] = 4.5                          # outside the domain: evaluate returns 0 there
        u = RNG.random(P)
        got = ll.sample2d_lowrank(A, CB, -4.0, 4.0, -4.0, 4.0, ry, u)
        want = orcc.sample2d_lowrank(A, CB, -4.0, 4.0, -4.0, 4.0, ry, u)
        ok = np.isfinite(want)
        assert np.array_equal(got[ok], want[ok])
    CB = RNG.standard_normal((7, 4))
    fA = RNG.standard_normal((4, 50))
    assert np.max(np.abs(ll.dgemm_nn(CB, fA) - CB @ fA)) < 1e-14


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_barrier_placement_under_other_fiber_orders(order):
    """Poor man's racecheck (compute-sanitizer is closed on the GPU pool): rerun a representative subset with the
    emulator scheduling the fibers of every barrier interval in reverse / permuted order.  A missing
    __syncthreads then reads stale or future data and changes the answer."""
    import subprocess
    import sys
    import os

    code = r'''
import numpy as np, sys
sys.path.insert(0, %r)
from tests.emu_harness import emu_lowlevel
from oracle import approxfun_oracle as orc
B, ll = emu_lowlevel()
rng = np.random.default_rng(3)
for n in (64, 256, 1024, 4096, 16384):
    v = rng.standard_normal((n, 2))
    assert np.max(np.abs(ll.transform(B.CHEB1_FWD, v) - orc.cheb_transform(v, axis=0))) < 1e-13
    assert np.max(np.abs(ll.transform(B.CHEB1_INV, v) - orc.cheb_itransform(v, axis=0))) < 1e-11
    assert np.max(np.abs(ll.transform(B.FOURIER_FWD, v) - orc.fourier_transform(v, axis=0))) < 1e-13
    assert np.max(np.abs(ll.transform(B.FOURIER_INV, v) - orc.fourier_itransform(v, axis=0))) < 1e-11
    z = v + 1j * rng.standard_normal((n, 2))
    if n <= 4096:
        assert np.max(np.abs(ll.transform(B.LAURENT_FWD, z) - orc.laurent_transform(z, axis=0))) < 1e-13
v = rng.standard_normal((16384, 6))
assert np.max(np.abs(ll.transform(B.CHEB1_FWD, v) - orc.cheb_transform(v, axis=0))) < 1e-13
assert np.max(np.abs(ll.transform(B.CHEB1_INV, v) - orc.cheb_itransform(v, axis=0))) < 1e-11
assert np.max(np.abs(ll.transform(B.FOURIER_FWD, v) - orc.fourier_transform(v, axis=0))) < 1e-13
assert np.max(np.abs(ll.transform(B.FOURIER_INV, v) - orc.fourier_itransform(v, axis=0))) < 1e-11
v = rng.standard_normal((100, 2))
assert np.max(np.abs(ll.transform(B.CHEB1_FWD, v) - orc.cheb_transform(v, axis=0))) < 1e-13
V = rng.standard_normal((64, 128))
assert np.max(np.abs(ll.transform2d(B.CHEB1_2D_FWD, V) - orc.cheb2_transform(V))) < 1e-13
for shape in ((100, 3), (1025, 2)):   # extension kinds in one kernel: Bluestein and power-of-two extension lengths
    v = rng.standard_normal(shape)
    assert np.max(np.abs(ll.transform(B.CHEB2_FWD, v) - orc.chebkind2_transform(v, axis=0))) < 1e-13
    assert np.max(np.abs(ll.transform(B.SIN_FWD, v)[:, 0] - orc.sin_transform_generic(v[:, 0]))) < 1e-13
z = rng.standard_normal((100, 2)) + 1j * rng.standard_normal((100, 2))
assert np.max(np.abs(ll.transform(B.LAURENT_FWD, z) - orc.laurent_transform(z, axis=0))) < 1e-13
V = rng.standard_normal((6, 2048))
assert np.max(np.abs(ll.transform2d(B.CHEB1_2D_FWD, V) - orc.cheb2_transform(V))) < 1e-13
assert np.max(np.abs(ll.transform2d(B.CHEB1_2D_INV, V) - orc.cheb2_itransform(V))) < 1e-11
c = rng.standard_normal(13000); x = rng.uniform(-1, 1, 600)
from oracle import oracle_c as orcc
assert np.array_equal(ll.clenshaw(c, x), orcc.clenshaw(c, x))
M = np.abs(rng.standard_normal((9, 140))) + 1
assert np.array_equal(ll.clenshaw_cols(M, x[:140]), orcc.clenshaw_cols(M, x[:140]))
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AFB_EMU_ORDER=order)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_host_sampler_with_device_uniforms():
    # afb_sample_cheb: Philox counter i for sample i, chunked pipeline; in the emulator "device" pointers are host pointers,
    # so the uniforms the kernel draws can be read back with afb_philox_uniform_device and pushed through the C oracle
    import ctypes as C
    from oracle import oracle_c as orcc
    c = orc.fun_coefficients(lambda x: np.exp(-x * x), 64, -5.0, 5.0)
    cf = orc.normalizedcumsum(c, -5.0, 5.0)
    n = 3001
    s = ll.sample_cheb(cf, -5.0, 5.0, 42, n)
    u = np.empty(n)
    assert ll.lib.afb_philox_uniform_device(42, 0, C.c_void_p(u.ctypes.data), n, None) == 0
    assert 0.0 <= u.min() and u.max() < 1.0
    assert np.array_equal(s, orc.fromcanonical(-5.0, 5.0, orcc.cheb_bisect(cf, u)))
    assert np.array_equal(ll.sample_cheb(cf, -5.0, 5.0, 42, 10), s[:10])      # reproducible prefix
    assert not np.array_equal(ll.sample_cheb(cf, -5.0, 5.0, 43, 10), s[:10])  # seed matters


# mixed-radix engine (afb_mixed.cu): n = 2^a 3^b 5^c 7^d in one shared-memory FFT instead of Bluestein
MIXED_SIZES = [6, 10, 12, 14, 15, 21, 35, 36, 48, 60, 63, 70, 96, 100, 105, 120, 125, 243, 245, 343, 360, 500, 1000, 1029,
               1536, 2187, 2400, 3000, 3125, 5000, 6000, 6144]


@pytest.mark.parametrize("n", MIXED_SIZES)
def test_mixed_radix_real_and_complex(n):
    nb = 5 if n <= 1000 else 3        # odd line counts: the last pair of a CTA is half empty
    v = RNG.standard_normal((n, nb))
    for kind, ref in REAL_KINDS:
        want = ref(v)
        got = ll.transform(kind, v)
        assert np.max(np.abs(got - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(v)))), (kind, n)
    z = RNG.standard_normal((n, 2)) + 1j * RNG.standard_normal((n, 2))
    for kind, ref in CPLX_KINDS:
        want = ref(z)
        got = ll.transform(kind, z)
        assert np.max(np.abs(got - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(z)))), (kind, n)


def test_mixed_radix_many_lines_per_cta():
    # short smooth lengths pack up to 16 pairs of lines into a CTA: pooled butterflies cross line boundaries
    for n, nb in ((12, 67), (60, 41), (100, 33)):
        v = RNG.standard_normal((n, nb))
        assert np.max(np.abs(ll.transform(B.CHEB1_FWD, v) - orc.cheb_transform(v, axis=0))) <= tol(n, np.max(np.abs(v)))
        assert np.max(np.abs(ll.transform(B.FOURIER_INV, v) - orc.fourier_itransform(v, axis=0))) <= tol(n, 20 * np.max(np.abs(v)))


def test_device_resident_fun_layer_on_the_emulator():
    # host logic of approxfun_b200.device (ladder, tail test, chop, product) with the kernels running in the emulator
    import approxfun_b200 as af
    from approxfun_b200 import api, device as dv
    keep = api._ll
    api._ll = ll
    try:
        S = af.Chebyshev(0, 3)
        f = dv.DeviceFun.from_function(lambda x: dv.sin(x ** 2), S)
        want = orc.adaptive_fun_coefficients(lambda x: np.sin(x ** 2), 0.0, 3.0)
        assert f.n == len(want) and np.max(np.abs(f.coefficients - want)) < 1e-14
        g = dv.DeviceFun.from_function(lambda x: dv.exp(-x), S)
        x = np.linspace(0, 3, 50)
        assert np.max(np.abs((f * g)(x) - np.sin(x ** 2) * np.exp(-x))) < 1e-13
        assert np.max(np.abs(dv.exp_fun(f)(x) - np.exp(np.sin(x ** 2)))) < 1e-13
        assert np.max(np.abs((f + g)(x) - np.sin(x ** 2) - np.exp(-x))) < 1e-13
    finally:
        api._ll = keep
        dv._plans.clear()


def test_mixed_radix_register_passes_match_shared_memory_passes():
    # forward kinds: first pass from global memory / last pass in registers (AFB_MIXED_REG_PASSES=1) against the all-shared-
    # memory passes (AFB_MIXED_SMEM_PASSES=1), in fresh processes because the switches are read once: same bits
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np\n"
        "from tests.emu_harness import emu_lowlevel\n"
        "B, ll = emu_lowlevel()\n"
        "rng = np.random.default_rng(5)\n"
        "out = []\n"
        "for n in (36, 48, 96, 100, 360, 1000, 1536, 3000):\n"
        "    v = rng.standard_normal((n, 5))\n"
        "    out += [ll.transform(B.CHEB1_FWD, v).ravel(), ll.transform(B.FOURIER_FWD, v).ravel()]\n"
        "np.save(__import__('sys').argv[1], np.concatenate(out))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for key in ("AFB_MIXED_REG_PASSES", "AFB_MIXED_SMEM_PASSES"):
        path = f"/tmp/afb_mixed_{key}.npy"
        env = dict(os.environ, **{key: "1"})
        env.pop("AFB_MIXED_SMEM_PASSES" if key == "AFB_MIXED_REG_PASSES" else "AFB_MIXED_REG_PASSES", None)
        subprocess.check_call([sys.executable, "-c", code, path], env=env, cwd=root)
        res.append(np.load(path))
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("n", [8000, 8002, 16000, 7875, 12003])
def test_split_plans(n):
    # n = S h, h inside a one-CTA engine (8000 -> 2 x 4000 mixed radix, 8002 -> 2 x 4001 fused Bluestein, 16000 -> 4 x 4000):
    # radix-S pass in the pre kernel, de-interleave in the post kernel (afb_large.cu, CfftPlan mode 4)
    v = RNG.standard_normal((n, 3))
    for kind, ref in REAL_KINDS:
        want = ref(v)
        got = ll.transform(kind, v)
        assert np.max(np.abs(got - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(v)))), (kind, n)
    z = RNG.standard_normal((n, 2)) + 1j * RNG.standard_normal((n, 2))
    for kind, ref in CPLX_KINDS:
        want = ref(z)
        assert np.max(np.abs(ll.transform(kind, z) - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(z)))), (kind, n)
    if n % 2:
        return                          # odd n (S = 3, 5, 7): no even / odd extension of that length
    m = n // 2 + 1                      # extension kinds: 2 (m - 1) = n
    w = RNG.standard_normal((m, 3))
    c2 = ll.transform(B.CHEB2_FWD, w)
    assert np.max(np.abs(c2 - orc.chebkind2_transform(w, axis=0))) <= tol(n, np.max(np.abs(w)))
    assert np.max(np.abs(ll.transform(B.CHEB2_INV, c2) - w)) <= 50 * tol(n, np.max(np.abs(w)))
    s = ll.transform(B.SIN_FWD, w[:m - 2])     # 2 (m - 2) + 2 = n
    assert np.max(np.abs(ll.transform(B.SIN_INV, s) - w[:m - 2])) <= 50 * tol(n, np.max(np.abs(w)))


@pytest.mark.parametrize("n", [4097, 10006])
def test_bluestein_padded_length_above_one_cta(n):
    # M = 16384 / 32768: radix-S pass in the chirp kernels around length-8192 one-CTA convolutions (blue_split_*_kernel)
    v = RNG.standard_normal((n, 3))
    for kind, ref in REAL_KINDS:
        want = ref(v)
        got = ll.transform(kind, v)
        assert np.max(np.abs(got - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(v)))), (kind, n)
    z = RNG.standard_normal((n, 2)) + 1j * RNG.standard_normal((n, 2))
    for kind, ref in CPLX_KINDS:
        want = ref(z)
        assert np.max(np.abs(ll.transform(kind, z) - want)) <= tol(n, max(np.max(np.abs(want)), np.max(np.abs(z)))), (kind, n)


def test_short_line_tile_kernel_variant():
    # line_kernel_tile (AFB_LINE_TILE=1: the CTA's lines staged as one coalesced tile; measured slower on the B200 and therefore
    # off by default) gives the same bits as the plain kernel; the switch is read once per process
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from tests.emu_harness import emu_lowlevel\n"
        "B, ll = emu_lowlevel()\n"
        "rng = np.random.default_rng(3)\n"
        "out = []\n"
        "for n in (64, 128, 256):\n"
        "    for nb in (1, 37, 100):\n"
        "        v = rng.standard_normal((n, nb))\n"
        "        for kind in (B.CHEB1_FWD, B.CHEB1_INV, B.FOURIER_FWD, B.FOURIER_INV):\n"
        "            out.append(ll.transform(kind, v).ravel())\n"
        "np.save(sys.argv[1], np.concatenate(out))\n" % root)
    res = {}
    for flag in ("0", "1"):
        path = os.path.join(root, "tests", "emu", "_tile_%s.npy" % flag)
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, AFB_LINE_TILE=flag), cwd=root)
        res[flag] = np.load(path)
        os.remove(path)
    assert np.array_equal(res["0"], res["1"])
