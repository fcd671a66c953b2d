"""Runs one n = 2^20 Chebyshev transform + inverse a few times -- the command profiled by ncu for the cfg-1 row."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

n = 1 << 20
st = torch.cuda.current_stream().cuda_stream
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
f = af.api.DevicePlan(L.CHEB1_FWD, n, 0, 1, n)
i = af.api.DevicePlan(L.CHEB1_INV, n, 0, 1, n)
for _ in range(3):
    f.execute(x.data_ptr(), y.data_ptr(), st)
    i.execute(y.data_ptr(), x.data_ptr(), st)
torch.cuda.synchronize()
print("profile run ok")
