"""Generates tests/golden/factors_r2.npz from the REFERENCE's own R2 factor classes (imported unmodified through
ref_shim.py): R2RelativeGaussianLikelihoodFactor (src/factors/Factors.py:912-1092) and UnaryR2RangeGaussianPriorFactor
(:2226-2298), the two classes the toy_examples/R2* scripts and the graph simulator build.

Run in the build container only:   python tests/golden/make_r2_factor_golden.py

* Densities: R2RelativeGaussianLikelihoodFactor.log_pdf delegates to TransportMaps' AdditiveLinearGaussianLogLikelihood
  (third party, TransportMaps==2.0b3, requirements.txt:18: absent here), so the golden values are the reference's own numpy
  restatement, `evaluate_loglike` (Factors.py:1070-1074), row by row.  UnaryR2RangeGaussianPriorFactor has no evaluable
  density in the reference (its distribution defines none and its evaluate_loglike subtracts the scalar range from the
  position vector, :2291-2293): only its sampler is pinned.
* Samplers: `.sample` with the numpy draws replayed from recorded arrays (make_sim_golden.Replay)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from make_sim_golden_replay import Replay  # noqa: E402
from factors.Factors import R2RelativeGaussianLikelihoodFactor, UnaryR2RangeGaussianPriorFactor  # noqa: E402
from slam.Variables import R2Variable  # noqa: E402

rng = np.random.default_rng(11)
out = {}
n = 96
A, B = R2Variable("x0"), R2Variable("x1")
cov = np.array([[0.3, 0.05], [0.05, 0.1]])
prec = np.array([[10.0, 0.0], [0.0, 10.0]])            # the toy examples' setting
for tag, kw in (("cov", dict(covariance=cov)), ("prec", dict(precision=prec))):
    f = R2RelativeGaussianLikelihoodFactor(A, B, np.array([5.0, -5.0]), **kw)
    x = np.hstack([rng.standard_normal((n, 2)) * 3.0, np.array([5.0, -5.0]) + rng.standard_normal((n, 2)) * 3.0])
    x[0, 2:] = x[0, :2] + np.array([5.0, -5.0])          # delta = 0
    out[f"r2rel_{tag}_x"] = x
    out[f"r2rel_{tag}_lp"] = np.array([f.evaluate_loglike(r) for r in x])
    out[f"r2rel_{tag}_covariance"] = np.asarray(f.covariance)
    pts, pts2 = rng.standard_normal((n, 2)) * 10.0, rng.standard_normal((n, 2)) * 10.0
    for name, args in (("fwd", dict(var1=pts)), ("bwd", dict(var2=pts)), ("obs", dict(var1=pts, var2=pts2))):
        eps = rng.standard_normal((n, 2))
        with Replay(normals=[eps]):
            out[f"r2rel_{tag}_{name}_eps"], out[f"r2rel_{tag}_{name}_out"] = eps, f.sample(**args)
    out[f"r2rel_{tag}_pts"], out[f"r2rel_{tag}_pts2"] = pts, pts2
out["r2rel_obs"] = np.array([5.0, -5.0])

rp = UnaryR2RangeGaussianPriorFactor(A, center=np.array([3.0, -1.0]), mu=7.5, sigma=0.4)
eps, u = rng.standard_normal((n, 1)), rng.random(n)
with Replay(normals=[eps], uniforms=[u]):
    out["range_prior_eps"], out["range_prior_u"], out["range_prior_out"] = eps, u, rp.sample(n)
out["range_prior_params"] = np.array([3.0, -1.0, 7.5, 0.4])

np.savez_compressed(os.path.join(HERE, "factors_r2.npz"), **out)
print("wrote factors_r2.npz:", {k: v.shape for k, v in out.items()})
