"""Host-call latency of small transforms through the Python mirror (plan creation vs plan reuse)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402

S = af.Chebyshev()
for n in (16, 256, 1024, 4096, 65536, 1 << 20):
    v = np.random.default_rng(0).standard_normal(n)
    af.transform(S, v)
    t0 = time.perf_counter()
    for _ in range(20):
        af.transform(S, v)
    t_new = (time.perf_counter() - t0) / 20
    p = af.plan_transform(S, v)
    p * v
    t0 = time.perf_counter()
    for _ in range(50):
        p * v
    t_reuse = (time.perf_counter() - t0) / 50
    print(f"n={n}: transform(S,v) {t_new * 1e6:.0f} us   plan*v (reused plan) {t_reuse * 1e6:.0f} us")
f = af.Fun(lambda x: np.sin(x * x), af.Chebyshev(0, 10))
t0 = time.perf_counter()
f = af.Fun(lambda x: np.sin(x * x), af.Chebyshev(0, 10))
print(f"adaptive Fun(sin(x^2), 0..10): {1e3 * (time.perf_counter() - t0):.1f} ms, {len(f)} coefficients")
x = np.linspace(0, 10, 1000)
f(x)
t0 = time.perf_counter()
for _ in range(20):
    f(x)
print(f"f(x) at 1000 points: {(time.perf_counter() - t0) / 20 * 1e6:.0f} us")
t0 = time.perf_counter()
for _ in range(50):
    f(0.1)
print(f"f(0.1) (scalar): {(time.perf_counter() - t0) / 50 * 1e6:.0f} us")
g = af.Fun(lambda t: np.exp(-t * t), af.Chebyshev(-5, 5))
af.sample(g, 1000)
t0 = time.perf_counter()
for _ in range(20):
    af.sample(g, 1000)
print(f"sample(g, 1000): {(time.perf_counter() - t0) / 20 * 1e6:.0f} us")
