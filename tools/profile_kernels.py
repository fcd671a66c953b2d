"""Runs each hot kernel a few times at its BASELINE.json shape -- the command profiled by ncu
(see profiles/README.md).  Not a benchmark: numbers printed under a profiler are never bench values."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, B = 4096, 65536
st = torch.cuda.current_stream().cuda_stream
v = torch.randn(B, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(v)
plans = {k: af.api.DevicePlan(k, n, 0, B, n) for k in (L.CHEB1_FWD, L.CHEB1_INV, L.FOURIER_FWD, L.FOURIER_INV)}
for _ in range(reps):
    for k, p in plans.items():
        p.execute(v.data_ptr(), c.data_ptr(), st)
torch.cuda.synchronize()
N, P = 10000, 1 << 24
cc = torch.randn(N, dtype=torch.float64, device="cuda") / (1 + torch.arange(N, device="cuda", dtype=torch.float64))
x = torch.rand(P, dtype=torch.float64, device="cuda") * 2 - 1
y = torch.empty_like(x)
f = af.Fun(lambda t: np.exp(-t * t), af.Chebyshev(-5, 5))
cdf = torch.tensor(af.normalizedcumsum(f).coefficients, device="cuda")
for _ in range(reps):
    af.api.clenshaw_device(cc.data_ptr(), N, x.data_ptr(), y.data_ptr(), P, st)
    af.api.sample_device(cdf.data_ptr(), cdf.numel(), -5.0, 5.0, 1, 0, y.data_ptr(), P, st)
torch.cuda.synchronize()
print("profile run ok")
