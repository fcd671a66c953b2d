import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import approxfun_b200 as af
from approxfun_b200 import _lib as L
from oracle import approxfun_oracle as orc
st = torch.cuda.current_stream().cuda_stream
n = 16384
for B in (2, 3, 149, 301, 1000):
    v = torch.randn(B, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(v)
    pl = af.api.DevicePlan(L.CHEB1_FWD, n, 0, B, n)
    pl.execute(v.data_ptr(), c.data_ptr(), st)
    torch.cuda.synchronize()
    want = orc.cheb_transform(v.cpu().numpy(), axis=1)
    err = float(np.max(np.abs(c.cpu().numpy() - want)))
    np.save(f"/tmp/cl_{os.environ.get('AFB_LINE_CLUSTER','0')}_{B}.npy", c.cpu().numpy()) if B == 301 else None
    print("cluster" if os.environ.get("AFB_LINE_CLUSTER") == "1" else "staged ", B, "max err vs oracle", err, flush=True)
