"""Generates tests/golden/factors.npz by evaluating the REFERENCE's own factor classes
(/root/reference/src/factors/Factors.py, imported unmodified through ref_shim.py) on seeded inputs.

Run in the build container only:   python tests/golden/make_factor_golden.py

Cases: SE(2) prior, SE(2) relative pose, SE2-R2 / R2-R2 range, 2- and 3-way ambiguous data
association, null-hypothesis mixture, the fused 14-factor joint of the small range graph (D = 22)
and of its data-association variant, mixture posterior weights.  For the range factors the
shimmed Gaussian is cross-checked against the reference's own `evaluate_loglike` arithmetic."""
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
from sampler.sampler_utils import JointFactor  # noqa: E402
from slam.FactorGraphSimulator import read_factor_graph_from_file  # noqa: E402
from slam.Variables import R2Variable, SE2Variable, VariableType  # noqa: E402

rng = np.random.default_rng(0)
out = {}

X0, X1 = SE2Variable("X0"), SE2Variable("X1")
L1, L2, L3 = (R2Variable(n, variable_type=VariableType.Landmark) for n in ("L1", "L2", "L3"))

# ---- SE2 prior (tight covariance: values reach -1e5)
prior = UnarySE2ApproximateGaussianPriorFactor(X0, SE2Pose(1.0, -2.0, 3.0), np.diag([4e-4, 1.6e-5, 4e-6]))
x = np.array([1.0, -2.0, 3.0]) + rng.standard_normal((64, 3)) * np.array([0.05, 0.01, 0.01])
x[0] = [1.0, -2.0, 3.0]                       # |theta| < 1e-5 branch of det_grad_x_logmap
x[1] = [1.3, -2.2, 3.0 + 2e-11]               # |w| < 1e-10 branch of log_map
x[2] = [1.0, -2.0, -3.2]                      # wraps across +-pi
x[3] = [0.5, -1.0, 3.0 + 7 * np.pi]           # unwrapped input angle
out["se2_prior_x"], out["se2_prior_lp"] = x, prior.log_pdf(x)
full = np.array([[4e-2, 1e-3, 2e-4], [1e-3, 2e-3, 1e-4], [2e-4, 1e-4, 5e-4]])
prior2 = UnarySE2ApproximateGaussianPriorFactor(X0, SE2Pose(-5.0, 7.0, -1.0), full)
x = np.array([-5.0, 7.0, -1.0]) + rng.standard_normal((64, 3)) * np.array([0.3, 0.1, 0.5])
out["se2_prior2_cov"], out["se2_prior2_x"], out["se2_prior2_lp"] = full, x, prior2.log_pdf(x)

# ---- SE2 relative pose (odometry)
cov = np.diag([.04, .0016, .0004])
btw = SE2RelativeGaussianLikelihoodFactor(X0, X1, SE2Pose(30.0, 0.0, 0.0), cov)
xi = rng.standard_normal((64, 3)) * np.array([5.0, 5.0, 2.0])
noise = rng.standard_normal((64, 3)) * np.array([0.3, 0.06, 0.03])
xj = np.array([(SE2Pose(*a) * SE2Pose(30.0, 0.0, 0.0) * SE2Pose.by_exp_map(nz)).array for a, nz in zip(xi, noise)])
x = np.hstack([xi, xj])
x[0, 3:] = (SE2Pose(*x[0, :3]) * SE2Pose(30.0, 0.0, 0.0)).array      # exact: w == 0
x[1, 5] += 2 * np.pi
x[2, 2] = 3.1415
out["se2_between_x"], out["se2_between_lp"] = x, btw.log_pdf(x)
btw2 = SE2RelativeGaussianLikelihoodFactor(X0, X1, SE2Pose(0.0, -30.0, -1.57079633), full)
x = np.hstack([rng.standard_normal((64, 3)) * 3.0, rng.standard_normal((64, 3)) * 3.0 + np.array([20.0, -20.0, -1.5])])
out["se2_between2_x"], out["se2_between2_lp"] = x, btw2.log_pdf(x)

# ---- range
rf = SE2R2RangeGaussianLikelihoodFactor(X0, L1, 42.42640687119285, 2.0)
x = np.hstack([rng.standard_normal((64, 3)) * 3.0, np.array([30.0, -30.0]) + rng.standard_normal((64, 2)) * 4.0])
out["range_x"], out["range_lp"] = x, rf.log_pdf(x)
assert np.allclose(out["range_lp"], [rf.evaluate_loglike(r) for r in x], rtol=0, atol=1e-12)   # shim vs reference arithmetic
r2 = R2RangeGaussianLikelihoodFactor(L1, L2, 30.0, 0.5)
x = np.hstack([rng.standard_normal((64, 2)) * 2.0, np.array([30.0, 0.0]) + rng.standard_normal((64, 2)) * 1.0])
out["r2range_x"], out["r2range_lp"] = x, r2.log_pdf(x)
assert np.allclose(out["r2range_lp"], [r2.evaluate_loglike(r) for r in x], rtol=0, atol=1e-12)

# ---- Gaussian landmark prior
gp = UnaryR2GaussianPriorFactor(L1, np.array([3.0, -4.0]), np.array([[0.5, 0.1], [0.1, 0.3]]))
x = np.array([3.0, -4.0]) + rng.standard_normal((64, 2))
out["gauss_x"], out["gauss_lp"] = x, gp.log_pdf(x)

# ---- mixtures: vars = [observer, observed...] (Factors.py:3070-3088)
ada2 = AmbiguousDataAssociationFactor(X0, [L1, L2], np.array([0.5, 0.5]), SE2R2RangeGaussianLikelihoodFactor, 60.0, 2.0)
x = np.hstack([rng.standard_normal((256, 3)) * 3.0, np.array([30.0, -50.0]) + rng.standard_normal((256, 2)) * 6.0,
               np.array([-40.0, 45.0]) + rng.standard_normal((256, 2)) * 6.0])
out["ada2_x"], out["ada2_lp"] = x, ada2.log_pdf(x)
var2x = {X0: x[:, :3], L1: x[:, 3:5], L2: x[:, 5:7]}
out["ada2_post_w"] = ada2.posterior_weights(var2x)
ada3 = AmbiguousDataAssociationFactor(X0, [L1, L2, L3], np.array([0.2, 0.5, 0.3]), SE2R2RangeGaussianLikelihoodFactor, 25.0, 1.5)
x = np.hstack([rng.standard_normal((256, 3)) * 2.0, np.array([20.0, 15.0]) + rng.standard_normal((256, 2)) * 3.0,
               np.array([-24.0, 5.0]) + rng.standard_normal((256, 2)) * 3.0,
               np.array([0.0, 300.0]) + rng.standard_normal((256, 2)) * 3.0])
x[:4, 3:9] = 1e4            # every component underflows: log(0) = -inf, responsibilities 0.5 each
with np.errstate(divide="ignore"):
    out["ada3_x"], out["ada3_lp"] = x, ada3.log_pdf(x)
out["ada3_post_w"] = ada3.posterior_weights({X0: x[:, :3], L1: x[:, 3:5], L2: x[:, 5:7], L3: x[:, 7:9]})
nh = BinaryFactorWithNullHypo(X0, L1, np.array([0.8, 0.2]), SE2R2RangeGaussianLikelihoodFactor, 42.4, 2.0, null_sigma_scale=10.0)
x = np.hstack([rng.standard_normal((128, 3)) * 3.0, np.array([30.0, -30.0]) + rng.standard_normal((128, 2)) * 15.0])
out["nullhypo_x"], out["nullhypo_lp"] = x, nh.log_pdf(x)

# ---- fused joints of the small graphs (D = 22, variable order X0..X5 L1 L2)
for tag, path in (("joint", "small_case1.fg"), ("joint_da", "small_case1_da.fg")):
    nodes, truth, factors = read_factor_graph_from_file(os.path.join(HERE, "..", "data", path))
    jf = JointFactor(factors, nodes)
    center = np.concatenate([truth[v] for v in nodes])
    x = center + rng.standard_normal((128, center.size)) * np.tile([0.3, 0.3, 0.02], 8)[:center.size]
    out[tag + "_x"], out[tag + "_lp"] = x, jf.log_pdf(x)
    out[tag + "_per_factor"] = np.array([f.log_pdf(x[:, jf.factor_to_indices[f]]) for f in factors])

np.savez_compressed(os.path.join(HERE, "factors.npz"), **out)
for k, v in out.items():
    if k.endswith("_lp"):
        print(k, v.shape, float(np.min(v[np.isfinite(v)])), float(np.max(v)))
