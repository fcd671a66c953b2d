"""Runs the 2-D Chebyshev transform (16384^2, forward and inverse) twice -- the command profiled by ncu for the cfg-4
rows of profiles/README.md.  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
st = torch.cuda.current_stream().cuda_stream
V = torch.randn(n, n, dtype=torch.float64, device="cuda")
W = torch.empty_like(V)
for kind in (L.CHEB1_2D_FWD, L.CHEB1_2D_INV):
    p = af.api.DevicePlan(kind, n, n, 1, n)
    for _ in range(2):
        p.execute(V.data_ptr(), W.data_ptr(), st)
torch.cuda.synchronize()
print("profile run ok")
