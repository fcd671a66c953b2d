"""Runs the long-line (n = 16384) Chebyshev kernels a few times -- the command profiled by ncu for the
n = 16384 row of profiles/README.md.  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = (1 << 28) // n
st = torch.cuda.current_stream().cuda_stream
v = torch.randn(B, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(v)
for k in (L.CHEB1_FWD, L.CHEB1_INV):
    p = af.api.DevicePlan(k, n, 0, B, n)
    for _ in range(2):
        p.execute(v.data_ptr(), c.data_ptr(), st)
torch.cuda.synchronize()
print("profile run ok")
