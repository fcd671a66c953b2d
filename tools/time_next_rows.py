"""Times the "next" rows (not a benchmark of record): Padua transform at several degrees, the general three-term Clenshaw,
mixed tensor spaces, piecewise bisection, the unfused 2-D sampler."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

lib = af.api.lowlevel().lib
st = torch.cuda.current_stream().cuda_stream
vp = C.c_void_p


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for d in (255, 1023, 2047, 4000):
    N = (d + 1) * (d + 2) // 2
    v = torch.randn(N, dtype=torch.float64, device="cuda")
    for kind, nm in ((L.PADUA_FWD, "fwd"), (L.PADUA_INV, "inv")):
        p = af.api.DevicePlan(kind, N, 0, 1, N)
        t = timeit(lambda: p.execute(v.data_ptr(), v.data_ptr(), st))
        print(f"padua {nm} degree {d} (N = {N}): {t:.3f} ms, {p.launches} launches, {N / t / 1e3:.3e} coefficients/s", flush=True)
        p.close()
# general three-term Clenshaw (Legendre), N = 10^4
N, P = 10000, 1 << 26
A, B, Cc = af.Legendre().rec(N)
dc = torch.tensor(np.random.default_rng(0).standard_normal(N) / (1 + np.arange(N)), device="cuda")
dA, dB, dC = (torch.tensor(a, device="cuda") for a in (A, B, Cc))
x = torch.rand(P, dtype=torch.float64, device="cuda") * 2 - 1
y = torch.empty_like(x)
t = timeit(lambda: L.check(lib, lib.afb_clenshaw_rec_device(vp(dc.data_ptr()), N, vp(dA.data_ptr()), vp(dB.data_ptr()), vp(dC.data_ptr()),
                                                           vp(x.data_ptr()), vp(y.data_ptr()), P, vp(st))), 3, 1)
print(f"clenshaw_rec (Legendre) N={N} P={P}: {t:.1f} ms  {P / t / 1e3:.3e} evals/s  {6.0 * N * P / t / 1e9:.2f} TFLOP/s (3 FMA per step)", flush=True)
# mixed tensor spaces
for n in (4096, 16384):
    V = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for kind, nm in ((L.FOURIER_CHEB1_2D_FWD, "Fourier x Chebyshev fwd"), (L.FOURIER_CHEB1_2D_INV, "Fourier x Chebyshev inv")):
        p = af.api.DevicePlan(kind, n, n, 1, n)
        t = timeit(lambda: p.execute(V.data_ptr(), V.data_ptr(), st))
        print(f"2D {nm} {n}^2: {t:.3f} ms  {32.0 * n * n / t / 1e6 / 6459.3:.3f} of the 32 B/elem roofline, {p.launches} launches", flush=True)
        p.close()
    del V
# piecewise bisection (generic bisectioninv on |sin| over -5..5)
f = af.api.abs_fun(af.Fun(np.sin, af.Chebyshev(-5, 5)))
cf = af.normalizedcumsum(f)
u = np.random.default_rng(1).random(1 << 22)
import time
af.api.bisectioninv_generic(cf, u[:1000])
t0 = time.perf_counter()
af.api.bisectioninv_generic(cf, u)
print(f"piecewise generic bisection, 4 pieces, {u.size} samples through the host call: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
