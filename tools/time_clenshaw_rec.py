"""Times the general three-term Clenshaw kernel at N = 10^4 (not a benchmark of record)."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import approxfun_b200 as af
from approxfun_b200 import _lib as L
import ctypes as C
from oracle import approxfun_oracle as orc
lib=af.api.lowlevel().lib
st=torch.cuda.current_stream().cuda_stream
N=10000; P=1<<26
A,B,Cc=orc.jacobi_rec(0.5,1.5,N)
c=torch.randn(N,dtype=torch.float64,device="cuda"); dA=torch.tensor(A,device="cuda"); dB=torch.tensor(B,device="cuda"); dC=torch.tensor(Cc,device="cuda")
x=torch.rand(P,dtype=torch.float64,device="cuda")*2-1; y=torch.empty_like(x)
def run(): L.check(lib, lib.afb_clenshaw_rec_device(C.c_void_p(c.data_ptr()),N,C.c_void_p(dA.data_ptr()),C.c_void_p(dB.data_ptr()),C.c_void_p(dC.data_ptr()),C.c_void_p(x.data_ptr()),C.c_void_p(y.data_ptr()),P,C.c_void_p(st)))
run(); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
t=e0.elapsed_time(e1)/2
print(f"clenshaw_rec N={N} P={P}: {t:.2f} ms, {P/t/1e6:.3f} Gevals/s, {6.0*N*P/t/1e9:.2f} TFLOP/s (3 FMA per step)")
