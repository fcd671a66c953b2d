import ctypes as C, sys, torch
sys.path.insert(0, '/root/repo')
import approxfun_b200 as af
from approxfun_b200 import _lib as L
lib = L.load(); af.api.lowlevel()
blocks, iters = 148*8, 20000
buf = torch.empty(blocks*256, dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for mode, fl in ((0,2.0),(1,2.0),(2,3.0),(3,3.0)):
    for _ in range(2): L.check(lib, lib.afb_fp64_probe_device(mode, iters, blocks, C.c_void_p(buf.data_ptr()), C.c_void_p(st)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): L.check(lib, lib.afb_fp64_probe_device(mode, iters, blocks, C.c_void_p(buf.data_ptr()), C.c_void_p(st)))
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)/3
    print(mode, fl*blocks*256*8*iters/(t*1e-3)/1e12, "TF")
