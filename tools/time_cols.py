"""Times the matrix-coefficient kernels and the fused 2-D sampler (development aid)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

lib = L.load()
af.api.lowlevel()
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


N, P = 66, 1 << 22
Cm = torch.rand(P, N, dtype=torch.float64, device="cuda") + 0.1       # column-major N x P
u = torch.rand(P, dtype=torch.float64, device="cuda")
out = torch.empty_like(u)
t = timeit(lambda: L.check(lib, lib.afb_cheb_normcumsum_device(vp(Cm.data_ptr()), N, P, N, 0, vp(st))))
print(f"normcumsum  N={N} P={P}: {t:.3f} ms  {16.0 * N * P / (t * 1e-3) / 1e9:.0f} GB/s")
t = timeit(lambda: L.check(lib, lib.afb_cheb_bisect_cols_device(vp(Cm.data_ptr()), N, N, vp(u.data_ptr()), vp(out.data_ptr()), P, 47, vp(st))))
print(f"bisect_cols N={N} P={P}: {t:.3f} ms  {47 * 3.0 * N * P / (t * 1e-3) / 1e12:.2f} TF  {P / (t * 1e-3):.3e} samples/s")
x = torch.rand(P, dtype=torch.float64, device="cuda") * 2 - 1
t = timeit(lambda: L.check(lib, lib.afb_clenshaw_cheb_cols_device(vp(Cm.data_ptr()), N, N, vp(x.data_ptr()), vp(out.data_ptr()), P, vp(st))))
print(f"clenshaw_cols N={N} P={P}: {t:.3f} ms  {8.0 * N * P / (t * 1e-3) / 1e9:.0f} GB/s")
R, NA = 8, 66
A = torch.randn(R, NA, dtype=torch.float64, device="cuda") / (1 + torch.arange(NA, device="cuda", dtype=torch.float64))
A[:, 0] = A[:, 0].abs() + 2
CB = torch.rand(R, N, dtype=torch.float64, device="cuda") / (1 + torch.arange(N, device="cuda", dtype=torch.float64)) ** 2
CB[:, 0] += 3
ry = torch.rand(P, dtype=torch.float64, device="cuda") * 8 - 4
t = timeit(lambda: L.check(lib, lib.afb_sample2d_cheb_lowrank_device(vp(A.data_ptr()), NA, vp(CB.data_ptr()), N, R, -4.0, 4.0, -4.0, 4.0,
                                                                      vp(ry.data_ptr()), vp(u.data_ptr()), vp(out.data_ptr()), P, vp(st))))
print(f"sample2d fused N={N} R={R} P={P}: {t:.3f} ms  {P / (t * 1e-3):.3e} samples/s  {(47 * 3.0 * N + 3.0 * R * NA + 2.0 * R * N) * P / (t * 1e-3) / 1e12:.2f} TF")
n1 = n2 = 64
Cc = torch.randn(n2, n1, dtype=torch.float64, device="cuda")
y = torch.rand(P, dtype=torch.float64, device="cuda") * 2 - 1
t = timeit(lambda: L.check(lib, lib.afb_clenshaw_cheb2d_device(vp(Cc.data_ptr()), n1, n2, vp(x.data_ptr()), vp(y.data_ptr()), vp(out.data_ptr()), P, vp(st))))
print(f"clenshaw2d {n1}x{n2} P={P}: {t:.3f} ms  {3.0 * n1 * n2 * P / (t * 1e-3) / 1e12:.2f} TF")
