"""Times the four real kinds (and Laurent) at lengths given on the command line (not a benchmark of record).
usage: python tools/time_sizes.py [--lib path/to/lib.so] n[:batch] ...      (batch defaults to 2 GiB / (8 n))"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxfun_b200 import _lib  # noqa: E402

args = sys.argv[1:]
if args and args[0] == "--lib":
    _lib.LIB_PATH = os.path.abspath(args[1])
    args = args[2:]
import approxfun_b200 as af  # noqa: E402

L = _lib
st = torch.cuda.current_stream().cuda_stream
peak = 6459.3


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


for a in args:
    n, _, b = a.partition(":")
    n = int(n)
    B = int(b) if b else max(1, (1 << 28) // n)
    v = torch.randn(B, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(v)
    row = [f"n={n} B={B}:"]
    for kind, nm in [(L.CHEB1_FWD, "cheb_fwd"), (L.CHEB1_INV, "cheb_inv"), (L.FOURIER_FWD, "four_fwd"), (L.FOURIER_INV, "four_inv")]:
        pl = af.api.DevicePlan(kind, n, 0, B, n)
        t = timeit(lambda: pl.execute(v.data_ptr(), c.data_ptr(), st))
        row.append(f"{nm} {t:.3f} ms {16.0 * n * B / (t * 1e-3) / 1e9 / peak:.3f}")
        pl.close()
    del v, c
    print(os.path.basename(_lib.LIB_PATH), "  ".join(row), flush=True)
