"""Clenshaw / sampler timing for build variants (points per thread etc.); usage: python tools/time_clenshaw.py [--lib lib.so]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxfun_b200 import _lib  # noqa: E402

args = sys.argv[1:]
if args and args[0] == "--lib":
    _lib.LIB_PATH = os.path.abspath(args[1])
import approxfun_b200 as af  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
N, P = 10000, 1 << 28
c = torch.randn(N, dtype=torch.float64, device="cuda") / (1 + torch.arange(N, device="cuda", dtype=torch.float64))
x = torch.rand(P, dtype=torch.float64, device="cuda") * 2 - 1
y = torch.empty_like(x)


def timeit(fn, reps=3, warm=1):
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


t = timeit(lambda: af.api.clenshaw_device(c.data_ptr(), N, x.data_ptr(), y.data_ptr(), P, st))
print(os.path.basename(_lib.LIB_PATH), f"clenshaw N={N} P={P}: {t:.1f} ms  {3.0 * N * P / t / 1e9:.2f} TFLOP/s  {P / t / 1e3:.3e} evals/s")
f = af.Fun(lambda t_: np.exp(-t_ * t_), af.Chebyshev(-5, 5))
cdf = torch.tensor(af.normalizedcumsum(f).coefficients, device="cuda")
t = timeit(lambda: af.api.sample_device(cdf.data_ptr(), cdf.numel(), -5.0, 5.0, 42, 0, y.data_ptr(), P, st))
print(os.path.basename(_lib.LIB_PATH), f"sample N={cdf.numel()} P={P}: {t:.1f} ms  {47 * 3.0 * cdf.numel() * P / t / 1e9:.2f} TFLOP/s  {P / t / 1e3:.3e} samples/s")
