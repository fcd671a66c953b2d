"""Times the arbitrary-length (Bluestein) transforms at a few non-power-of-two lengths (not a benchmark of record)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import approxfun_b200 as af
from approxfun_b200 import _lib as L
st=torch.cuda.current_stream().cuda_stream
peak=6459.3
for n in (100, 1000, 1025, 3000, 4097, 10000, 12345):
    B=max(1,(1<<26)//n)
    v=torch.randn(B,n,dtype=torch.float64,device="cuda"); c=torch.empty_like(v)
    for kind,nm in ((L.CHEB1_FWD,"cheb_fwd"),(L.FOURIER_FWD,"four_fwd"),(L.CHEB2_FWD,"cheb2_fwd")):
        p=af.api.DevicePlan(kind,n,0,B,n)
        for _ in range(2): p.execute(v.data_ptr(),c.data_ptr(),st)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): p.execute(v.data_ptr(),c.data_ptr(),st)
        e1.record(); torch.cuda.synchronize()
        t=e0.elapsed_time(e1)/3
        print(f"{nm} n={n} B={B}: {t:.3f} ms {16.0*n*B/(t*1e-3)/1e9/peak:.3f} of HBM, launches {p.launches}")
        p.close()
