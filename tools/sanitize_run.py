"""Small-shape run of every kernel family, written as a workload for a memory / race checker.  The GPU pool of this project
refuses compute-sanitizer, so it is not run under it there: the bounds checks that exist are the AddressSanitizer build of the
emulator (tests/emu/Makefile, tests/test_emulator_asan.py); on a machine where the tool is allowed this script is the command."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import approxfun_b200 as af  # noqa: E402

rng = np.random.default_rng(0)
for n in (8, 33, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 100, 1000):
    nb = 5 if n <= 4096 else 2
    v = rng.standard_normal((n, nb))
    for S in (af.Chebyshev(), af.Fourier()):
        c = af.transform(S, v)
        w = af.itransform(S, c)
        assert np.max(np.abs(w - v)) < 1e-11, (n, type(S).__name__)
    if n <= 8192:
        z = v + 1j * rng.standard_normal((n, nb))
        assert np.max(np.abs(af.itransform(af.Laurent(), af.transform(af.Laurent(), z)) - z)) < 1e-11
V = rng.standard_normal((96, 64))
C2 = af.transform(af.Chebyshev() ** 2, V)
assert np.max(np.abs(af.itransform(af.Chebyshev() ** 2, C2) - V)) < 1e-11
c = rng.standard_normal(300)
x = rng.uniform(-1, 1, 5000)
af.clenshaw(af.Chebyshev(), c, x)
af.clenshaw(af.Chebyshev(), rng.standard_normal(13000), x[:700])
M = np.abs(rng.standard_normal((20, 300))) + 1
gm = af.chebnormalizedcumsum(M)
af.chebbisectioninv(gm, rng.random(300))
af.clenshaw(M, rng.uniform(-1, 1, 300))
af.api.lowlevel().clenshaw_multi(M[:, :7], x[:100])
f = af.Fun(lambda t: np.exp(-t * t), af.Chebyshev(-5, 5))
af.sample(f, 3000)
af.points(af.Chebyshev(0, 10), 1000)
print("sanitize run ok")
