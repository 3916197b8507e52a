"""Pins the numpy factor oracle (oracle/factor_oracle.py) -- and the host-side descriptor builders of
nfisam_b200.factors -- to golden values produced by the reference's own factor classes
(tests/golden/make_factor_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import factor_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(HERE, "golden", "factors.npz")))


def build_factors(g):
    from nfisam_b200.factors import (AmbiguousDataAssociationFactor, BinaryFactorWithNullHypo,
                                     R2RangeGaussianLikelihoodFactor, SE2Pose, SE2R2RangeGaussianLikelihoodFactor,
                                     SE2RelativeGaussianLikelihoodFactor, UnaryR2GaussianPriorFactor,
                                     UnarySE2ApproximateGaussianPriorFactor)
    from nfisam_b200.slam import R2Variable, SE2Variable, VariableType

    X0, X1 = SE2Variable("X0"), SE2Variable("X1")
    L1, L2, L3 = (R2Variable(n, VariableType.Landmark) for n in ("L1", "L2", "L3"))
    rng_f = SE2R2RangeGaussianLikelihoodFactor
    return {
        "se2_prior": UnarySE2ApproximateGaussianPriorFactor(X0, SE2Pose(1.0, -2.0, 3.0), np.diag([4e-4, 1.6e-5, 4e-6])),
        "se2_prior2": UnarySE2ApproximateGaussianPriorFactor(X0, SE2Pose(-5.0, 7.0, -1.0), g["se2_prior2_cov"]),
        "se2_between": SE2RelativeGaussianLikelihoodFactor(X0, X1, SE2Pose(30.0, 0.0, 0.0), np.diag([.04, .0016, .0004])),
        "se2_between2": SE2RelativeGaussianLikelihoodFactor(X0, X1, SE2Pose(0.0, -30.0, -1.57079633), g["se2_prior2_cov"]),
        "range": rng_f(X0, L1, 42.42640687119285, 2.0),
        "r2range": R2RangeGaussianLikelihoodFactor(L1, L2, 30.0, 0.5),
        "gauss": UnaryR2GaussianPriorFactor(L1, np.array([3.0, -4.0]), np.array([[0.5, 0.1], [0.1, 0.3]])),
        "ada2": AmbiguousDataAssociationFactor(X0, [L1, L2], np.array([0.5, 0.5]), rng_f, 60.0, 2.0),
        "ada3": AmbiguousDataAssociationFactor(X0, [L1, L2, L3], np.array([0.2, 0.5, 0.3]), rng_f, 25.0, 1.5),
        "nullhypo": BinaryFactorWithNullHypo(X0, L1, np.array([0.8, 0.2]), rng_f, 42.4, 2.0, null_sigma_scale=10.0),
    }


def check(got, ref, tol=1e-6):
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.array_equal(got[~fin], ref[~fin])          # -inf where every component underflows
    assert np.max(np.abs(got[fin] - ref[fin])) <= tol    # north_star: factor log-likelihoods within 1e-6


def test_single_factors(g):
    from nfisam_b200.factors import oracle_descriptor

    for key, f in build_factors(g).items():
        with np.errstate(divide="ignore"):
            got = fo.factor_logpdf(oracle_descriptor(f), g[key + "_x"])
        check(got, g[key + "_lp"])


def test_posterior_weights(g):
    from nfisam_b200.factors import oracle_descriptor

    fs = build_factors(g)
    assert np.allclose(fo.posterior_weights(oracle_descriptor(fs["ada2"]), g["ada2_x"]), g["ada2_post_w"], atol=1e-12)
    assert np.allclose(fo.posterior_weights(oracle_descriptor(fs["ada3"]), g["ada3_x"]), g["ada3_post_w"], atol=1e-12)


@pytest.mark.parametrize("tag,path", [("joint", "small_case1.fg"), ("joint_da", "small_case1_da.fg")])
def test_joint_of_small_graph(g, tag, path):
    from nfisam_b200.factors import JointFactor, oracle_descriptor
    from nfisam_b200.slam.graph_io import factor_graph_to_string, read_factor_graph_from_file

    nodes, truth, factors = read_factor_graph_from_file(os.path.join(HERE, "data", path))
    assert [v.name for v in nodes] == ["X0", "X1", "X2", "X3", "X4", "X5", "L1", "L2"] and len(factors) == 14
    jf = JointFactor(factors, nodes)
    x = g[tag + "_x"]
    descs = [oracle_descriptor(f, jf._col_of) for f in factors]
    check(fo.joint_logpdf(descs, x), g[tag + "_lp"], tol=1e-6)
    for f, d, ref in zip(factors, descs, g[tag + "_per_factor"]):
        check(fo.factor_logpdf(d, x), ref)
    # text round trip of the graph
    text = factor_graph_to_string(nodes, factors, truth)
    tmp = os.path.join(HERE, "golden", "_roundtrip.fg")
    try:
        open(tmp, "w").write(text)
        n2, t2, f2 = read_factor_graph_from_file(tmp)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    assert [str(a) for a in f2] == [str(a) for a in factors]


def test_geometry_exp_log_closed_form():
    """SE(2) exp/log against the closed-form V-matrix the reference's geometry test uses
    (tests/test_2d_geometry_classes.py:108-139)."""
    from nfisam_b200.factors.geometry import se2_compose, se2_exp, se2_inverse, se2_log

    rng = np.random.default_rng(0)
    v = rng.standard_normal((50, 3)) * np.array([3.0, 3.0, 1.0])
    T = se2_exp(v)
    w = v[:, 2]
    V = np.array([[np.sin(w) / w, -(1 - np.cos(w)) / w], [(1 - np.cos(w)) / w, np.sin(w) / w]]).transpose(2, 0, 1)
    assert np.allclose(T[:, :2], np.einsum("nij,nj->ni", V, v[:, :2]), atol=1e-12)
    assert np.allclose(se2_log(T), v, atol=1e-10)
    I = se2_compose(T, se2_inverse(T))
    assert np.allclose(I, 0.0, atol=1e-12)


def test_sampling_statistics():
    """Forward simulators: sample -> log-map residual statistics match the noise model."""
    from nfisam_b200.factors import SE2Pose, SE2R2RangeGaussianLikelihoodFactor, SE2RelativeGaussianLikelihoodFactor
    from nfisam_b200.factors.geometry import se2_compose, se2_inverse, se2_log
    from nfisam_b200.slam import R2Variable, SE2Variable

    np.random.seed(0)
    a, b, l = SE2Variable("A"), SE2Variable("B"), R2Variable("L")
    cov = np.diag([.04, .0016, .0004])
    f = SE2RelativeGaussianLikelihoodFactor(a, b, SE2Pose(30.0, 0.0, 0.3), cov)
    xi = np.random.standard_normal((20000, 3))
    xj = f.sample(var1=xi)
    res = se2_log(se2_compose(se2_inverse(np.array([30.0, 0.0, 0.3])), se2_compose(se2_inverse(xi), xj)))
    assert np.allclose(np.cov(res.T), cov, atol=2e-3)
    back = f.sample(var2=xj)
    res2 = se2_log(se2_compose(se2_inverse(np.array([30.0, 0.0, 0.3])), se2_compose(se2_inverse(back), xj)))
    assert np.allclose(np.cov(res2.T), cov, atol=2e-3)
    obs = f.sample(var1=xi, var2=xj)
    assert obs.shape == (20000, 3) and abs(obs[:, 0].mean() - 30.0) < 0.05
    r = SE2R2RangeGaussianLikelihoodFactor(a, l, 10.0, 0.5)
    lm = r.sample(var1=xi)
    d = np.linalg.norm(lm - xi[:, :2], axis=1)
    assert abs(d.mean() - 10.0) < 0.02 and abs(d.std() - 0.5) < 0.02
    assert r.sample(var1=xi, var2=lm).shape == (20000, 1)


# ---- R2 factor classes of the reference's toy examples (tests/golden/make_r2_factor_golden.py) -----------------------
def r2_factors(g2):
    from nfisam_b200.factors import R2RelativeGaussianLikelihoodFactor, UnaryR2RangeGaussianPriorFactor
    from nfisam_b200.slam import R2Variable

    A, B = R2Variable("x0"), R2Variable("x1")
    cx, cy, mu, sigma = g2["range_prior_params"]
    return {"r2rel_cov": R2RelativeGaussianLikelihoodFactor(A, B, g2["r2rel_obs"], covariance=np.array([[0.3, 0.05], [0.05, 0.1]])),
            "r2rel_prec": R2RelativeGaussianLikelihoodFactor(A, B, g2["r2rel_obs"], precision=np.array([[10.0, 0.0], [0.0, 10.0]])),
            "range_prior": UnaryR2RangeGaussianPriorFactor(A, np.array([cx, cy]), mu, sigma)}


@pytest.fixture(scope="module")
def g2():
    return dict(np.load(os.path.join(HERE, "golden", "factors_r2.npz")))


def test_r2_relative_factor_matches_reference(g2):
    """Density (the reference's evaluate_loglike, Factors.py:1070-1074) and the three sampling directions with replayed
    draws (Factors.py:995-1036), through the host classes and the numpy oracles."""
    from nfisam_b200.factors import oracle_descriptor
    from oracle import sim_oracle as so

    fs = r2_factors(g2)
    for tag in ("cov", "prec"):
        f = fs["r2rel_" + tag]
        np.testing.assert_allclose(f.covariance, g2[f"r2rel_{tag}_covariance"], rtol=1e-12)
        check(fo.factor_logpdf(oracle_descriptor(f), g2[f"r2rel_{tag}_x"]), g2[f"r2rel_{tag}_lp"], tol=1e-9)
        chol = np.linalg.cholesky(f.covariance)
        pts, pts2 = g2[f"r2rel_{tag}_pts"], g2[f"r2rel_{tag}_pts2"]
        tol = dict(rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(so.r2_gen_fwd(pts, g2["r2rel_obs"], g2[f"r2rel_{tag}_fwd_eps"] @ chol.T), g2[f"r2rel_{tag}_fwd_out"], **tol)
        np.testing.assert_allclose(so.r2_gen_bwd(pts, g2["r2rel_obs"], g2[f"r2rel_{tag}_bwd_eps"] @ chol.T), g2[f"r2rel_{tag}_bwd_out"], **tol)
        np.testing.assert_allclose(so.r2_obs(pts, pts2, g2[f"r2rel_{tag}_obs_eps"] @ chol.T), g2[f"r2rel_{tag}_obs_out"], **tol)
    text = str(fs["r2rel_cov"])
    from nfisam_b200.factors import Factor

    back = Factor.construct_from_text(text, fs["r2rel_cov"].vars)
    np.testing.assert_allclose(back.covariance, fs["r2rel_cov"].covariance)
    assert [v.name for v in back.vars] == ["x0", "x1"] and back.observation_var.dim == 2


def test_r2_range_prior_sampler_matches_reference(g2):
    from nfisam_b200.factors import Factor
    from oracle import sim_oracle as so

    cx, cy, mu, sigma = g2["range_prior_params"]
    ang = -np.pi + 2 * np.pi * g2["range_prior_u"]
    got = so.range_prior([cx, cy], mu, sigma * g2["range_prior_eps"][:, 0], ang)
    np.testing.assert_allclose(got, g2["range_prior_out"], rtol=1e-12, atol=1e-12)
    f = r2_factors(g2)["range_prior"]
    back = Factor.construct_from_text(str(f), f.vars)
    assert back.mu == mu and abs(back.covariance - sigma ** 2) < 1e-15 and np.allclose(back.center, [cx, cy])
