"""Golden vectors for the two-sample statistics (row N4): runs the reference's own `mmd`, `MMDu2`, `MMDb`
(src/utils/Statistics.py:13-84) in the build container and stores inputs + outputs in tests/golden/stats.npz.

    python tests/golden/make_stats_golden.py        # needs /root/reference
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402  (stands in for TransportMaps / matplotlib / dynesty)

ref_shim.install()
from utils.Statistics import MMDb, MMDu2, mmd  # noqa: E402

rng = np.random.default_rng(7)
out = {}
cases = [("a", 40, 55, 3, 1.0), ("b", 64, 64, 7, 2.5), ("c", 130, 90, 22, np.sqrt(22.0)), ("d", 33, 35, 1, 0.3)]
for name, m, n, d, sigma in cases:
    x = rng.normal(size=(m, d)) * rng.uniform(0.5, 3.0, size=d)
    y = rng.normal(size=(n, d)) * rng.uniform(0.5, 3.0, size=d) + 0.4
    out[f"{name}_x"], out[f"{name}_y"], out[f"{name}_sigma"] = x, y, sigma
    out[f"{name}_mmdb"] = MMDb(x, y, sigma)
    out[f"{name}_mmdu2"] = MMDu2(x, y, sigma)
    if m <= 64:      # the reference's mmd is a Python double loop
        out[f"{name}_mmd"] = float(np.squeeze(mmd(x, y, k_sigma2=sigma ** 2)))
# identical sets: the biased estimate is exactly zero up to rounding
x = rng.normal(size=(50, 4))
out["same_x"] = x
out["same_mmdb"] = MMDb(x, x.copy(), 1.5)
out["same_mmdu2"] = MMDu2(x, x.copy(), 1.5)
# sample_mean (src/utils/Statistics.py:151-171) on a posterior-like matrix: X0 (SE2), L1 (R2), X1 (SE2) with angles straddling +-pi
from slam.Variables import R2Variable, SE2Variable  # noqa: E402
from utils.Statistics import sample_mean  # noqa: E402

order = [SE2Variable("X0"), R2Variable("L1"), SE2Variable("X1")]
xs = rng.normal(size=(700, 8)) * np.array([2.0, 3.0, 0.4, 5.0, 1.0, 0.5, 0.7, 1.2]) + np.array([10.0, -4.0, 3.0, 1.0, 2.0, -7.0, 8.0, -3.1])
xs[:, [2, 7]] = (xs[:, [2, 7]] + np.pi) % (2 * np.pi) - np.pi
out["sm_x"] = xs
out["sm_mean"], _ = sample_mean(xs, order)
np.savez(os.path.join(HERE, "stats.npz"), **out)
print({k: float(v) for k, v in out.items() if np.ndim(v) == 0})
