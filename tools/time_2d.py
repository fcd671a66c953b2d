"""Times the 2-D Chebyshev transform at 16384^2, forward and inverse (not a benchmark of record)."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
import approxfun_b200 as af
from approxfun_b200 import _lib as L
n=16384; st=torch.cuda.current_stream().cuda_stream
V=torch.randn(n,n,dtype=torch.float64,device="cuda"); W=torch.empty_like(V)
for kind,nm in [(L.CHEB1_2D_FWD,"fwd"),(L.CHEB1_2D_INV,"inv")]:
    p=af.api.DevicePlan(kind,n,n,1,n)
    for _ in range(3): p.execute(V.data_ptr(),W.data_ptr(),st)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): p.execute(V.data_ptr(),W.data_ptr(),st)
    e1.record(); torch.cuda.synchronize()
    print(nm, os.environ.get("AFB_ROWS_DEBUG"), e0.elapsed_time(e1)/5)
