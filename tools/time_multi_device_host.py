import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import approxfun_b200 as af
from approxfun_b200 import _lib as L
G = torch.cuda.device_count()
af.api.use_devices(list(range(G)))
m = 16384
rng = np.random.default_rng(0)
M = rng.standard_normal((m, m)); out = np.empty_like(M)
pl = af.api.DevicePlan(L.CHEB1_2D_FWD, m, m, 1, m)
def best(fn, reps=3):
    fn(); t = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t = min(t, time.perf_counter() - t0)
    return t * 1e3
print(f"{G} GPUs one process, 2-D 16384^2 host call, pageable: {best(lambda: pl.execute_host(M, out)):.1f} ms")
with af.api.pinned(M, out):
    print(f"registered: {best(lambda: pl.execute_host(M, out)):.1f} ms")
V = rng.standard_normal((65536, 4096)); C = np.empty_like(V)
p1 = af.api.DevicePlan(L.CHEB1_FWD, 4096, 0, 65536, 4096)
print(f"batched 65536 x 4096 host call, pageable: {best(lambda: p1.execute_host(V, C)):.1f} ms")
with af.api.pinned(V, C):
    print(f"registered: {best(lambda: p1.execute_host(V, C)):.1f} ms")
