"""GPU parity of the float64 factor kernels (through nfisam_factor_logpdf /
nfisam_mixture_posterior_weights) against the golden vectors of the reference's own factor classes and
against the numpy oracle on larger seeded inputs.  Tolerance: |diff| <= 1e-6 (north_star), written below."""
import os

import numpy as np
import pytest

from oracle import factor_oracle as fo
from tests.test_oracle_factors import build_factors

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-6


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(HERE, "golden", "factors.npz")))


def test_single_factors_golden(g):
    for key, f in build_factors(g).items():
        got = f.log_pdf(g[key + "_x"])
        # values reach 2e6 in magnitude: 1e-6 absolute there is 5e-13 relative
        ref = g[key + "_lp"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(got), fin), key
        assert np.max(np.abs(got[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]) * 1e-3)) <= TOL, key
        pdf = f.pdf(g[key + "_x"][:8])
        assert np.allclose(pdf, np.exp(ref[:8]), rtol=1e-9, atol=0)


def test_posterior_weights_golden(g):
    fs = build_factors(g)
    from nfisam_b200.slam import R2Variable, SE2Variable

    X0, L1, L2, L3 = SE2Variable("X0"), R2Variable("L1"), R2Variable("L2"), R2Variable("L3")
    x = g["ada2_x"]
    w = fs["ada2"].posterior_weights({X0: x[:, :3], L1: x[:, 3:5], L2: x[:, 5:7]})
    assert np.allclose(w, g["ada2_post_w"], atol=1e-9)
    x = g["ada3_x"]
    w = fs["ada3"].posterior_weights({X0: x[:, :3], L1: x[:, 3:5], L2: x[:, 5:7], L3: x[:, 7:9]})
    assert np.allclose(w, g["ada3_post_w"], atol=1e-9)


@pytest.mark.parametrize("tag,path", [("joint", "small_case1.fg"), ("joint_da", "small_case1_da.fg")])
def test_fused_joint_golden(g, tag, path):
    from nfisam_b200.factors import JointFactor
    from nfisam_b200.slam.graph_io import read_factor_graph_from_file

    nodes, truth, factors = read_factor_graph_from_file(os.path.join(HERE, "data", path))
    jf = JointFactor(factors, nodes)
    total, per = jf.log_pdf(g[tag + "_x"], per_factor=True)
    assert np.max(np.abs(total - g[tag + "_lp"]) / np.maximum(1.0, np.abs(g[tag + "_lp"]) * 1e-3)) <= TOL
    assert np.max(np.abs(per - g[tag + "_per_factor"]) / np.maximum(1.0, np.abs(g[tag + "_per_factor"]) * 1e-3)) <= TOL


@pytest.mark.parametrize("n", [1, 127, 100_000])
def test_joint_vs_oracle_sizes(n):
    from nfisam_b200.factors import JointFactor, oracle_descriptor
    from nfisam_b200.slam.graph_io import read_factor_graph_from_file

    nodes, truth, factors = read_factor_graph_from_file(os.path.join(HERE, "data", "small_case1_da.fg"))
    jf = JointFactor(factors, nodes)
    rng = np.random.default_rng(n)
    center = np.concatenate([truth[v] for v in nodes])
    x = center + rng.standard_normal((n, center.size)) * np.tile([0.5, 0.5, 0.05], 8)[:center.size]
    x[:, 2] += 2 * np.pi * rng.integers(-2, 3, n)          # unwrapped angles must not matter
    got = jf.log_pdf(x)
    exp = fo.joint_logpdf([oracle_descriptor(f, jf._col_of) for f in factors], x)
    assert np.max(np.abs(got - exp) / np.maximum(1.0, np.abs(exp) * 1e-3)) <= TOL


def test_device_tensor_input_and_bad_descriptor():
    import torch

    from nfisam_b200 import _lib
    from nfisam_b200.factors import SE2R2RangeGaussianLikelihoodFactor, _gpu
    from nfisam_b200.slam import R2Variable, SE2Variable

    f = SE2R2RangeGaussianLikelihoodFactor(SE2Variable("X"), R2Variable("L"), 5.0, 1.0)
    x = torch.randn(1000, 5, dtype=torch.float64, device="cuda")
    out = _gpu.logpdf([f.components({f.vars[0]: 0, f.vars[1]: 3})], x)
    assert out.is_cuda and out.shape == (1000,)
    r = (x[:, :2] - x[:, 3:5]).norm(dim=1)
    assert torch.allclose(out, -0.5 * (r - 5.0) ** 2 - 0.5 * np.log(2 * np.pi), atol=1e-12)
    with pytest.raises(_lib.NfisamError):
        _gpu.logpdf([[dict(type="range", cols=[0, 1, 2, 9], obs=[1.0], info=[1.0], lnorm=0.0)]], x)


def test_r2_factor_classes_golden():
    """R2RelativeGaussianLikelihoodFactor / UnaryR2RangeGaussianPriorFactor (the toy_examples/R2* factor classes): kernel
    densities against the reference's evaluate_loglike values (factors_r2.npz) and the oracle, |diff| <= 1e-6."""
    from nfisam_b200.factors import JointFactor, oracle_descriptor
    from tests.test_oracle_factors import r2_factors

    g2 = dict(np.load(os.path.join(HERE, "golden", "factors_r2.npz")))
    fs = r2_factors(g2)
    for tag in ("cov", "prec"):
        got = fs["r2rel_" + tag].log_pdf(g2[f"r2rel_{tag}_x"])
        assert np.max(np.abs(got - g2[f"r2rel_{tag}_lp"])) <= TOL
    rng = np.random.default_rng(5)
    x = np.array([3.0, -1.0]) + rng.standard_normal((1000, 2)) * 6.0
    got = fs["range_prior"].log_pdf(x)
    assert np.max(np.abs(got - fo.factor_logpdf(oracle_descriptor(fs["range_prior"]), x))) <= TOL
    # fused joint over a row holding both variables
    A, B = fs["r2rel_cov"].vars
    jf = JointFactor([fs["range_prior"], fs["r2rel_cov"]], [A, B])
    x = np.hstack([x, x + np.array([5.0, -5.0]) + rng.standard_normal((1000, 2))])
    exp = fo.joint_logpdf([oracle_descriptor(f, jf._col_of) for f in jf.factors], x)
    assert np.max(np.abs(jf.log_pdf(x) - exp)) <= TOL


def test_batched_posterior_weights_equal_per_factor_calls(g):
    """nfisam_mixture_posterior_weights_batch (row N2): all mixtures of a step in one launch == one call per factor (the reference's
    loop, FactorGraphSolver.py:913-922), and both equal the reference's stored weights."""
    from nfisam_b200.factors import posterior_weights_batch
    from nfisam_b200.slam import R2Variable, SE2Variable

    fs = build_factors(g)
    X0, L1, L2, L3 = SE2Variable("X0"), R2Variable("L1"), R2Variable("L2"), R2Variable("L3")
    x = g["ada3_x"]
    var2x = {X0: x[:, :3], L1: x[:, 3:5], L2: x[:, 5:7], L3: x[:, 7:9]}
    mixtures = [fs["ada3"], fs["ada2"], fs["nullhypo"], fs["ada3"]]
    got = posterior_weights_batch(mixtures, var2x)
    for f, w in zip(mixtures, got):
        assert np.allclose(w, f.posterior_weights(var2x), rtol=0, atol=1e-12)
    assert np.allclose(got[0], g["ada3_post_w"], atol=1e-9) and abs(sum(got[2]) - 1.0) < 1e-12
    assert posterior_weights_batch([], var2x) == []
