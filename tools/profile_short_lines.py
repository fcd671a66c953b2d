"""One launch of each real kind at n = 64, 128, 256 (2 GiB batches) -- the command profiled by ncu for the short-line kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
for n in (64, 128, 256):
    B = (1 << 28) // n
    v = torch.randn(B, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(v)
    for kind in (L.CHEB1_FWD, L.CHEB1_INV, L.FOURIER_FWD, L.FOURIER_INV):
        pl = af.api.DevicePlan(kind, n, 0, B, n)
        pl.execute(v.data_ptr(), c.data_ptr(), st)
        pl.execute(v.data_ptr(), c.data_ptr(), st)
        torch.cuda.synchronize()
        pl.close()
    del v, c
print("ok")
