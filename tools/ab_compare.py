"""A/B harness: times the four real kinds at a few lengths with the library given on the command line (build variants
as approxfun.jl_b200/libafb_<name>.so), so that two builds are compared on the SAME box -- boxes of the pool differ by
several per cent.  usage: python tools/ab_compare.py path/to/lib.so"""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from approxfun_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = os.path.abspath(sys.argv[1])
import approxfun_b200 as af
L=_lib
st=torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
out=[]
for n,B in [(4096,65536),(2048,1<<17),(1024,1<<18),(256,1<<20),(64,1<<22)]:
    v=torch.randn(B,n,dtype=torch.float64,device="cuda"); c=torch.empty_like(v)
    for kind in (L.CHEB1_FWD,L.CHEB1_INV,L.FOURIER_FWD,L.FOURIER_INV):
        pl=af.api.DevicePlan(kind,n,0,B,n)
        t=timeit(lambda: pl.execute(v.data_ptr(),c.data_ptr(),st))
        out.append(f"{16.0*n*B/(t*1e-3)/1e9/6459.3:.3f}")
        pl.close()
print(os.path.basename(_lib.LIB_PATH), " ".join(out))
