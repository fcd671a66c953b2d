"""Host-array calls on PAGEABLE NumPy arrays (what a Julia Array is): 2-D transform, Padua-sized transforms, sample(f, n),
Clenshaw with a short series.  Run twice for the A/B: AFB_NO_RING=1 hands the pageable pointers to the driver as round 1 did.
Not a benchmark of record."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxfun_b200 import _lib as L  # noqa: E402

if len(sys.argv) > 2 and sys.argv[1] == "--lib":
    L.LIB_PATH = os.path.abspath(sys.argv[2])
import approxfun_b200 as af  # noqa: E402

tag = "driver-staged (AFB_NO_RING=1)" if os.environ.get("AFB_NO_RING") == "1" else "pinned ring"


def best(fn, reps=3):
    fn()
    t = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t = min(t, time.perf_counter() - t0)
    return t * 1e3


rng = np.random.default_rng(0)
V = rng.standard_normal((65536, 4096))
C = np.empty_like(V)
pl = af.api.DevicePlan(L.CHEB1_FWD, 4096, 0, 65536, 4096)
print(f"[{tag}] {os.path.basename(L.LIB_PATH)} batched 65536 x 4096 host call: {best(lambda: pl.execute_host(V, C)):.1f} ms", flush=True)
pl.close()
del V, C
if "--only-1d" in sys.argv:
    sys.exit(0)
for m in (8192, 16384):
    M = rng.standard_normal((m, m))
    out = np.empty_like(M)
    pl = af.api.DevicePlan(L.CHEB1_2D_FWD, m, m, 1, m)
    print(f"[{tag}] 2-D transform {m}^2 host call: {best(lambda: pl.execute_host(M, out)):.1f} ms", flush=True)
    pl.close()
    del M, out
f = af.Fun(lambda t: np.exp(-t * t), af.Chebyshev(-5, 5))
P = 1 << 28
print(f"[{tag}] sample(f, 2^28) (2 GiB out): {best(lambda: af.sample(f, P), 2):.1f} ms", flush=True)
x = rng.random(P) * 2 - 1
c = f.coefficients
ll = af.api.lowlevel()
print(f"[{tag}] clenshaw N = {c.size}, 2^28 points (2 GiB in, 2 GiB out): {best(lambda: ll.clenshaw(c, x), 2):.1f} ms", flush=True)
