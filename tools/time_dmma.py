"""DMMA (FP64 tensor core) measurements for DESIGN.md 4b (not a benchmark of record):
   (1) issue-rate probes: vector DFMA variants against mma.sync m8n8k4.f64
   (2) the one dense contraction of the path, AB = CB * fA (src/Extras/sample.jl:210): FMA kernel against the DMMA kernel."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402
from approxfun_b200 import _lib as L  # noqa: E402

lib = af.api.lowlevel().lib
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


blocks, iters = 148 * 8, 20000
buf = torch.empty(blocks * 256, dtype=torch.float64, device="cuda")
for mode, name, flop in ((0, "DFMA 3 register operands", 2.0 * 8 * 256), (1, "DFMA repeated operand", 2.0 * 8 * 256),
                         (2, "Clenshaw step DADD+DFMA", 3.0 * 8 * 256), (4, "DMMA m8n8k4", 512.0 * 8 * 8)):
    t = timeit(lambda: L.check(lib, lib.afb_fp64_probe_device(mode, iters, blocks, C.c_void_p(buf.data_ptr()), C.c_void_p(st))), 3, 2)
    print(f"probe {name}: {flop * blocks * iters / (t * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)

g = torch.Generator(device="cuda").manual_seed(1)
for N, R, P in ((66, 8, 1 << 22), (128, 16, 1 << 21), (198, 32, 1 << 20), (198, 50, 1 << 20), (64, 64, 1 << 21)):
    CB = torch.randn(R, N, dtype=torch.float64, device="cuda", generator=g)     # column-major N x R
    fA = torch.randn(P, R, dtype=torch.float64, device="cuda", generator=g)     # column-major R x P
    AB0 = torch.empty(P, N, dtype=torch.float64, device="cuda")
    AB1 = torch.empty(P, N, dtype=torch.float64, device="cuda")
    args = lambda o: (C.c_void_p(CB.data_ptr()), C.c_void_p(fA.data_ptr()), C.c_void_p(o.data_ptr()), N, R, P, C.c_void_p(st))
    t0 = timeit(lambda: L.check(lib, lib.afb_dgemm_nn_device(*args(AB0))))
    t1 = timeit(lambda: L.check(lib, lib.afb_dgemm_nn_dmma_device(*args(AB1))))
    err = float((AB0 - AB1).abs().max())
    ref = float((AB0 - fA @ CB).abs().max())
    fl, by = 2.0 * N * R * P, 8.0 * (N * P + R * P)
    print(f"AB = CB*fA N={N} R={R} P={P}: FMA {t0:.3f} ms ({fl / t0 / 1e9:.1f} TFLOP/s, {by / t0 / 1e6:.0f} GB/s) | "
          f"DMMA {t1:.3f} ms ({fl / t1 / 1e9:.1f} TFLOP/s, {by / t1 / 1e6:.0f} GB/s) | max|FMA-DMMA| {err:.1e}, max|FMA-torch| {ref:.1e}",
          flush=True)
