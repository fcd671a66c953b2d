"""Times the Padua transform at a few degrees (not a benchmark of record)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import approxfun_b200 as af
from approxfun_b200 import _lib as L
st=torch.cuda.current_stream().cuda_stream
for d in (255, 1023, 2047, 4000):
    N=(d+1)*(d+2)//2
    v=torch.randn(N,dtype=torch.float64,device="cuda"); c=torch.empty_like(v)
    for kind,nm in ((L.PADUA_FWD,"fwd"),(L.PADUA_INV,"inv")):
        p=af.api.DevicePlan(kind,N,0,1,N)
        for _ in range(2): p.execute(v.data_ptr(),c.data_ptr(),st)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): p.execute(v.data_ptr(),c.data_ptr(),st)
        e1.record(); torch.cuda.synchronize()
        t=e0.elapsed_time(e1)/5
        print(f"padua {nm} d={d} N={N}: {t:.3f} ms, {N/t/1e6:.2f} Gcoef/s, launches {p.launches}")
        p.close()
