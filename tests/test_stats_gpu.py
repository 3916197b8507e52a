"""GPU parity of the two-sample statistics (row N4, nfisam_mmd through nfisam_b200.utils) against the reference's own
outputs (tests/golden/stats.npz) and the numpy oracle on larger seeded inputs.  Tolerance: 1e-9 relative (float64;
the reference's sklearn distances use the |x|^2 + |y|^2 - 2 x.y expansion, the kernel sums squared differences)."""
import os

import numpy as np
import pytest

from oracle import stats_oracle as so

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
