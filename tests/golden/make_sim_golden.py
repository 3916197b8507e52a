"""Generates tests/golden/sim.npz: outputs of the REFERENCE's own factor `.sample` methods
(/root/reference/src/factors/Factors.py, imported unmodified through ref_shim.py) with their random draws replayed
from recorded arrays, so that each case is a deterministic (noise -> samples) vector the oracle can be pinned to.

Run in the build container only:   python tests/golden/make_sim_golden.py

Replay: the shimmed Gaussian's `rvs(m)` is `standard_normal((m, dim)) @ chol.T` (ref_shim.py), the range factors draw
their bearing with `np.random.uniform(-pi, pi, m)` (Factors.py:2585, 2599) and mixtures split the rows with
`np.random.multinomial(n, weights)` (Factors.py:3148, 3262, 3343, 3355, 3367); the three numpy entry points are
swapped for queues of pre-drawn arrays while a reference method runs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from factors.Factors import (AmbiguousDataAssociationFactor, BinaryFactorWithNullHypo,  # noqa: E402
                             R2RangeGaussianLikelihoodFactor, SE2R2RangeGaussianLikelihoodFactor,
                             SE2RelativeGaussianLikelihoodFactor, UnaryR2GaussianPriorFactor,
                             UnarySE2ApproximateGaussianPriorFactor)
from geometry.TwoDimension import SE2Pose  # noqa: E402
from slam.Variables import R2Variable, SE2Variable, VariableType  # noqa: E402


from make_sim_golden_replay import Replay  # noqa: E402

rng = np.random.default_rng(7)
out = {}
n = 96
X0, X1 = SE2Variable("X0"), SE2Variable("X1")
L1, L2, L3 = (R2Variable(k, variable_type=VariableType.Landmark) for k in ("L1", "L2", "L3"))
full = np.array([[4e-2, 1e-3, 2e-4], [1e-3, 2e-3, 1e-4], [2e-4, 1e-4, 5e-4]])

# ---- SE2 prior
eps = rng.standard_normal((n, 3))
eps[0, 2] = 0.0                                   # |w| < 1e-10 branch of the exp map
prior = UnarySE2ApproximateGaussianPriorFactor(X0, SE2Pose(1.0, -2.0, 3.0), full)
with Replay(normals=[eps]):
    out["se2_prior_eps"], out["se2_prior_cov"], out["se2_prior_out"] = eps, full, prior.sample(n)

# ---- R2 Gaussian prior
eps = rng.standard_normal((n, 2))
cov2 = np.array([[0.5, 0.1], [0.1, 0.2]])
gp = UnaryR2GaussianPriorFactor(L1, np.array([3.0, -4.0]), cov2)
with Replay(normals=[eps]):
    out["r2_prior_eps"], out["r2_prior_cov"], out["r2_prior_out"] = eps, cov2, gp.sample(n)

# ---- SE2 relative pose: var2 from var1, var1 from var2, observation
btw = SE2RelativeGaussianLikelihoodFactor(X0, X1, SE2Pose(30.0, 1.0, -1.2), full)
poses = rng.standard_normal((n, 3)) * np.array([20.0, 20.0, 2.5])
poses[1, 2] = 3.1415
poses2 = rng.standard_normal((n, 3)) * np.array([20.0, 20.0, 2.5])
for name, kw in (("fwd", dict(var1=poses)), ("bwd", dict(var2=poses)), ("obs", dict(var1=poses, var2=poses2))):
    eps = rng.standard_normal((n, 3))
    with Replay(normals=[eps]):
        out[f"se2_{name}_eps"], out[f"se2_{name}_out"] = eps, btw.sample(**kw)
out["se2_rel_cov"], out["se2_rel_obs"], out["se2_poses"], out["se2_poses2"] = full, btw.observation, poses, poses2

# ---- range: ring around a pose / a point, observation
lm = rng.standard_normal((n, 2)) * 15.0
for cls, tag, a_s, va, vb in ((SE2R2RangeGaussianLikelihoodFactor, "se2r2", poses, X0, L1),
                              (R2RangeGaussianLikelihoodFactor, "r2r2", lm, L1, L2)):
    f = cls(va, vb, 12.5, 0.4)
    eps, u = rng.standard_normal((n, 1)), rng.random(n)
    with Replay(normals=[eps], uniforms=[u]):
        out[f"range_{tag}_gen_eps"], out[f"range_{tag}_gen_u"], out[f"range_{tag}_gen_out"] = eps, u, f.sample(var1=a_s)
    eps = rng.standard_normal((n, 1))
    other = rng.standard_normal((n, 2)) * 15.0
    with Replay(normals=[eps]):
        out[f"range_{tag}_obs_eps"], out[f"range_{tag}_obs_b"], out[f"range_{tag}_obs_out"] = eps, other, f.sample(var1=a_s, var2=other)
out["range_lm"] = lm

# ---- ambiguous data association (3 candidates): simulated observations and observer sampling
w = np.array([0.5, 0.3, 0.2])
counts = np.array([50, 27, 19])
ada = AmbiguousDataAssociationFactor(X0, [L1, L2, L3], w, SE2R2RangeGaussianLikelihoodFactor, 9.0, 0.3)
lms = [rng.standard_normal((n, 2)) * 15.0 for _ in range(3)]
eps = [rng.standard_normal((int(c), 1)) for c in counts]
with Replay(normals=eps, multinomials=[counts]):
    res = ada.sample_observations({X0: poses, L1: lms[0], L2: lms[1], L3: lms[2]})
out["ada_counts"], out["ada_lms"], out["ada_obs_eps"], out["ada_obs_out"] = counts, np.stack(lms), np.concatenate(eps), res

# ---- null-hypothesis mixture: var2 from var1 (ring with sigma / 10 sigma), observation
nh = BinaryFactorWithNullHypo(X0, L1, np.array([0.8, 0.2]), SE2R2RangeGaussianLikelihoodFactor, 7.0, 0.25, 10.0)
counts2 = np.array([70, 26])
eps = [rng.standard_normal((int(c), 1)) for c in counts2]
us = [rng.random(int(c)) for c in counts2]
with Replay(normals=eps, uniforms=us, multinomials=[counts2]):
    res = nh.sample(var1=poses)
out["nh_counts"], out["nh_gen_eps"], out["nh_gen_u"], out["nh_gen_out"] = counts2, np.concatenate(eps), np.concatenate(us), res
eps = [rng.standard_normal((int(c), 1)) for c in counts2]
with Replay(normals=eps, multinomials=[counts2]):
    res = nh.sample(var1=poses, var2=lm)
out["nh_obs_eps"], out["nh_obs_out"] = np.concatenate(eps), res

np.savez_compressed(os.path.join(HERE, "sim.npz"), **out)
print("wrote sim.npz:", {k: v.shape for k, v in out.items()})
