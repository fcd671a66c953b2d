"""Generates tests/golden/golden_v1.npz from the CPU oracle.

The reference is Julia (absent from this image), so these vectors cannot come from the reference
binary; they are produced by the oracle after it has been pinned to the reference's own golden values
(tests/test_oracle_goldens.py) and cross-checked against O(n^2) extended-precision sums for n <= 100.
They freeze today's oracle output so that oracle regressions and GPU regressions are told apart."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import approxfun_oracle as orc  # noqa: E402
from oracle import oracle_c as orcc  # noqa: E402

rng = np.random.default_rng(20261019)
out = {}
for n in (8, 64, 100, 4096):
    v = rng.standard_normal(n)
    out[f"v_{n}"] = v
    out[f"cheb_fwd_{n}"] = orc.cheb_transform(v)
    out[f"cheb_inv_{n}"] = orc.cheb_itransform(v)
    out[f"fourier_fwd_{n}"] = orc.fourier_transform(v)
    out[f"fourier_inv_{n}"] = orc.fourier_itransform(v)
    if n <= 100:
        assert np.max(np.abs(out[f"cheb_fwd_{n}"] - orc.cheb_transform_direct(v))) < 1e-14
c = rng.standard_normal(200) / (1 + np.arange(200)) ** 1.5
x = rng.uniform(-1, 1, 512)
out["cl_c"], out["cl_x"], out["cl_y"] = c, x, orcc.clenshaw(c, x)
pdf = orc.adaptive_fun_coefficients(lambda t: np.exp(-t * t), -5.0, 5.0)
cdf = orc.normalizedcumsum(pdf, -5.0, 5.0)
u = rng.random(512)
out["bis_c"], out["bis_u"], out["bis_x"] = cdf, u, orcc.cheb_bisect(cdf, u)
out["ncs_in"] = pdf
out["ncs_out"] = orcc.cheb_normcumsum(pdf)
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
print("wrote golden_v1.npz", {k: v.shape for k, v in out.items()})
