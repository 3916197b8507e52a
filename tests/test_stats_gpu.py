"""GPU parity of the two-sample statistics (row N4, nfisam_mmd through nfisam_b200.utils) against the reference's own
outputs (tests/golden/stats.npz) and the numpy oracle on larger seeded inputs.  Tolerance: 1e-9 relative (float64;
the reference's sklearn distances use the |x|^2 + |y|^2 - 2 x.y expansion, the kernel sums squared differences)."""
import os

import numpy as np
import pytest

from oracle import stats_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "stats.npz"))


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_matches_reference_golden(case):
    from nfisam_b200.utils import MMDb, MMDu2, mmd

    x, y, sigma = G[f"{case}_x"], G[f"{case}_y"], float(G[f"{case}_sigma"])
    assert MMDb(x, y, sigma) == pytest.approx(float(G[f"{case}_mmdb"]), rel=1e-9)
    assert MMDu2(x, y, sigma) == pytest.approx(float(G[f"{case}_mmdu2"]), rel=1e-9, abs=1e-13)
    if f"{case}_mmd" in G:
        assert mmd(x, y, sigma ** 2) == pytest.approx(float(G[f"{case}_mmd"]), rel=1e-9)


@pytest.mark.parametrize("m,n,d", [(1, 1, 2), (2, 3, 5), (129, 64, 3), (1000, 777, 22), (3000, 2500, 12), (400, 400, 64),
                                   (400, 333, 65), (300, 500, 158), (257, 129, 308)])
def test_kernel_sums_match_oracle(m, n, d):
    from nfisam_b200.utils.statistics import _mmd

    rng = np.random.default_rng(m + n + d)
    x = rng.normal(size=(m, d)) * 2.0
    y = rng.normal(size=(n, d)) * 1.5 + 0.3
    sigma = float(np.sqrt(d))
    for kind, skip in ((0, False), (1, True)):
        if skip and min(m, n) < 2:
            continue
        val, sums = _mmd(x, y, sigma, kind, want_sums=True)
        ref = so.kernel_sums(x, y, sigma, skip)
        np.testing.assert_allclose(sums, ref, rtol=1e-11)      # wide rows: column-chunked summation order
        want = so.MMDu2(x, y, sigma) if skip else so.MMDb(x, y, sigma)
        assert val == pytest.approx(want, rel=1e-9, abs=1e-13)


def test_deterministic_and_symmetric():
    from nfisam_b200.utils import MMDb, MMDu2

    rng = np.random.default_rng(5)
    x, y = rng.normal(size=(700, 6)), rng.normal(size=(650, 6)) + 0.2
    a = MMDu2(x, y, 2.0)
    assert a == MMDu2(x, y, 2.0)                       # fixed-order reductions: bitwise reproducible
    assert MMDu2(y, x, 2.0) == pytest.approx(a, rel=1e-10)
    assert MMDb(x, x.copy(), 2.0) == pytest.approx(0.0, abs=1e-7)


def test_bad_arguments_fail_loudly():
    from nfisam_b200 import _lib
    from nfisam_b200.utils import MMDb, MMDu2

    x = np.zeros((4, 3))
    with pytest.raises(_lib.NfisamError):
        MMDb(x, x, 0.0)
    with pytest.raises(_lib.NfisamError):
        MMDu2(x[:1], x, 1.0)
    with pytest.raises(ValueError):
        MMDb(x, np.zeros((4, 2)), 1.0)


def test_marginal_statistics_kernel_matches_reference_and_oracle():
    """nfisam_marginal_stats (row N2): means equal the reference's sample_mean (stats.npz), covariance blocks equal the oracle's;
    float32 device matrix, float64 accumulation."""
    from nfisam_b200.slam import R2Variable, SE2Variable
    from nfisam_b200.utils import marginal_mean_cov, sample_mean

    g = np.load(os.path.join(HERE, "golden", "stats.npz"))
    order = [SE2Variable("X0"), R2Variable("L1"), SE2Variable("X1")]
    x = g["sm_x"]
    means, var2mean = sample_mean(x, order)
    assert np.max(np.abs(means - g["sm_mean"])) < 2e-6          # the kernel reads float32 samples
    x32 = x.astype(np.float32).astype(np.float64)
    mean_o, cov_o = so.sample_mean_cov(x32, [0, 0, 1, 0, 0, 0, 0, 1])
    means, var2mean, var2cov = marginal_mean_cov(x, order)
    assert np.max(np.abs(means - mean_o)) < 1e-12
    off = 0
    for v in order:
        assert np.max(np.abs(var2cov[v] - cov_o[off:off + v.dim, off:off + v.dim])) < 1e-10, v.name
        off += v.dim
    # device-resident float32 input, many variables
    import torch
    rng = np.random.default_rng(0)
    many = [SE2Variable(f"X{k}") for k in range(200)]
    xs = (rng.standard_normal((1000, 600)) * 2.0).astype(np.float32)
    means, _, var2cov = marginal_mean_cov(torch.from_numpy(xs).cuda(), many)
    circ = np.tile([0, 0, 1], 200)
    mo, co = so.sample_mean_cov(xs.astype(np.float64), circ)
    assert np.max(np.abs(means - mo)) < 1e-10 and np.max(np.abs(var2cov[many[57]] - co[171:174, 171:174])) < 1e-10
