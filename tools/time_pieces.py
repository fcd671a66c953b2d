"""Times individual pieces (not a benchmark of record): usage  python tools/time_pieces.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

st = torch.cuda.current_stream().cuda_stream


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


peak = 6459.3
for n, B in [(16384, 16384), (8192, 32768), (4096, 65536), (2048, 131072), (1024, 262144), (512, 1 << 19), (256, 1 << 20), (128, 1 << 21), (64, 1 << 22)]:
    v = torch.randn(B, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(v)
    for kind, nm in [(L.CHEB1_FWD, "cheb_fwd"), (L.CHEB1_INV, "cheb_inv"), (L.FOURIER_FWD, "four_fwd"), (L.FOURIER_INV, "four_inv")]:
        pl = af.api.DevicePlan(kind, n, 0, B, n)
        t = timeit(lambda: pl.execute(v.data_ptr(), c.data_ptr(), st))
        print(f"{nm} n={n} B={B}: {t:.3f} ms  {16.0 * n * B / (t * 1e-3) / 1e9 / peak:.3f} of HBM")
        pl.close()
    del v, c
n = 4096
B = 32768
z = torch.randn(B, n, dtype=torch.complex128, device="cuda")
o = torch.empty_like(z)
for kind, nm in [(L.LAURENT_FWD, "laur_fwd"), (L.LAURENT_INV, "laur_inv")]:
    pl = af.api.DevicePlan(kind, n, 0, B, n)
    t = timeit(lambda: pl.execute(z.data_ptr(), o.data_ptr(), st))
    print(f"{nm} n={n} B={B}: {t:.3f} ms  {32.0 * n * B / (t * 1e-3) / 1e9 / peak:.3f} of HBM")
del z, o
# 2-D pieces
n = 16384
V = torch.randn(n, n, dtype=torch.float64, device="cuda")
W = torch.empty_like(V)
t = timeit(lambda: W.copy_(V))
print(f"torch copy 16384^2: {t:.3f} ms {16.0 * n * n / (t * 1e-3) / 1e9:.0f} GB/s")
t = timeit(lambda: W.copy_(V.t()))
print(f"torch transpose-copy 16384^2: {t:.3f} ms {16.0 * n * n / (t * 1e-3) / 1e9:.0f} GB/s")
del V, W
for n in (16384, 8192, 4096, 2048):
    V = torch.randn(n, n, dtype=torch.float64, device="cuda")
    W = torch.empty_like(V)
    for kind, nm in [(L.CHEB1_2D_FWD, "fwd"), (L.CHEB1_2D_INV, "inv")]:
        for flags, route in [(0, "row pairs four-step"), (L.PLAN_2D_TRANSPOSE, "transposes")]:
            p2 = af.api.DevicePlan(kind, n, n, 1, n, flags)
            t = timeit(lambda: p2.execute(V.data_ptr(), W.data_ptr(), st))
            print(f"2D {nm} {n}^2 [{route}]: {t:.3f} ms  {32.0 * n * n / (t * 1e-3) / 1e9 / peak:.3f} of the 32 B/elem roofline, {p2.launches} launches")
            p2.close()
    del V, W
# single long transforms
for lg in (15, 17, 20, 22):
    n1 = 1 << lg
    x = torch.randn(n1, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    pl = af.api.DevicePlan(L.CHEB1_FWD, n1, 0, 1, n1)
    t = timeit(lambda: pl.execute(x.data_ptr(), y.data_ptr(), st), reps=20, warm=3)
    print(f"cheb_fwd single n=2^{lg}: {t * 1e3:.1f} us, launches {pl.launches}")
