"""Pins oracle/sim_oracle.py (CPU restatement of the clique training-set simulation, row N1) to
tests/golden/sim.npz -- outputs of the reference's own factor `.sample` methods with their random draws replayed
(tests/golden/make_sim_golden.py) -- and its Philox4x32-10 generator to the Random123 known-answer vectors."""
import os

import numpy as np
import pytest

from oracle import sim_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = dict(rtol=1e-12, atol=1e-12)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "sim.npz"))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    cases = [([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
             ([0xffffffff] * 4, [0xffffffff] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
             ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for ctr, key, want in cases:
        got = so.philox4x32(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert " ".join("%08x" % v for v in got) == want


def test_noise_streams_are_standard():
    rows = np.arange(200_000)
    u0, u1 = so.uniform2(11, rows, 4)
    n0, n1 = so.normal2(11, rows, 5)
    assert 0.0 <= u0.min() and u0.max() < 1.0 and abs(u0.mean() - 0.5) < 5e-3 and abs(u1.var() - 1 / 12) < 2e-3
    for v in (n0, n1):
        assert abs(v.mean()) < 1e-2 and abs(v.std() - 1.0) < 1e-2 and abs(np.mean(v ** 4) - 3.0) < 0.1
    assert abs(np.corrcoef(n0, n1)[0, 1]) < 1e-2
    # different slots / seeds / rows decorrelate
    assert abs(np.corrcoef(n0, so.normal2(11, rows, 6)[0])[0, 1]) < 1e-2
    assert abs(np.corrcoef(n0, so.normal2(12, rows, 5)[0])[0, 1]) < 1e-2


def test_se2_prior(g):
    chol = np.linalg.cholesky(g["se2_prior_cov"])
    got = so.se2_prior([1.0, -2.0, 3.0], g["se2_prior_eps"] @ chol.T)
    np.testing.assert_allclose(got, g["se2_prior_out"], **TOL)


def test_r2_prior(g):
    chol = np.linalg.cholesky(g["r2_prior_cov"])
    np.testing.assert_allclose(np.array([3.0, -4.0]) + g["r2_prior_eps"] @ chol.T, g["r2_prior_out"], **TOL)


def test_se2_relative(g):
    chol = np.linalg.cholesky(g["se2_rel_cov"])
    obs, p1, p2 = g["se2_rel_obs"], g["se2_poses"], g["se2_poses2"]
    np.testing.assert_allclose(so.se2_gen_fwd(p1, obs, g["se2_fwd_eps"] @ chol.T), g["se2_fwd_out"], **TOL)
    np.testing.assert_allclose(so.se2_gen_bwd(p1, obs, g["se2_bwd_eps"] @ chol.T), g["se2_bwd_out"], **TOL)
    np.testing.assert_allclose(so.se2_obs(p1, p2, g["se2_obs_eps"] @ chol.T), g["se2_obs_out"], **TOL)


@pytest.mark.parametrize("tag,centers", [("se2r2", "se2_poses"), ("r2r2", "range_lm")])
def test_range(g, tag, centers):
    c = g[centers]
    ang = -np.pi + 2 * np.pi * g[f"range_{tag}_gen_u"]
    np.testing.assert_allclose(so.range_gen(c, 12.5, 0.4 * g[f"range_{tag}_gen_eps"][:, 0], ang), g[f"range_{tag}_gen_out"], **TOL)
    got = so.range_obs(c, g[f"range_{tag}_obs_b"], 0.4 * g[f"range_{tag}_obs_eps"][:, 0])
    np.testing.assert_allclose(got, g[f"range_{tag}_obs_out"][:, 0], **TOL)


def test_mixture_row_ranges(g):
    # ambiguous data association: component c simulates the observation for its contiguous block of rows
    edges = np.concatenate([[0], np.cumsum(g["ada_counts"])])
    got = np.zeros(len(g["se2_poses"]))
    for c in range(3):
        lo, hi = edges[c], edges[c + 1]
        got[lo:hi] = so.range_obs(g["se2_poses"][lo:hi], g["ada_lms"][c][lo:hi], 0.3 * g["ada_obs_eps"][lo:hi, 0])
    np.testing.assert_allclose(got, g["ada_obs_out"][:, 0], **TOL)
    # null hypothesis: second component is the 10x wider model
    edges = np.concatenate([[0], np.cumsum(g["nh_counts"])])
    gen = np.zeros((len(got), 2))
    obs = np.zeros(len(got))
    for c, sigma in enumerate((0.25, 2.5)):
        lo, hi = edges[c], edges[c + 1]
        ang = -np.pi + 2 * np.pi * g["nh_gen_u"][lo:hi]
        gen[lo:hi] = so.range_gen(g["se2_poses"][lo:hi], 7.0, sigma * g["nh_gen_eps"][lo:hi, 0], ang)
        obs[lo:hi] = so.range_obs(g["se2_poses"][lo:hi], g["range_lm"][lo:hi], sigma * g["nh_obs_eps"][lo:hi, 0])
    np.testing.assert_allclose(gen, g["nh_gen_out"], **TOL)
    np.testing.assert_allclose(obs, g["nh_obs_out"][:, 0], **TOL)


def test_normalize_training_matches_reference():
    m = np.load(os.path.join(HERE, "golden", "model.npz"))
    data, means, stds = so.normalize_training(m["norm_raw"], m["norm_circ"])
    np.testing.assert_allclose(means, m["norm_means"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(stds, m["norm_stds"], rtol=1e-6)
    np.testing.assert_allclose(data, m["norm_data"], rtol=1e-5, atol=1e-6)


def test_interpreter_matches_the_transforms():
    """simulate() = the pinned transforms fed with the Philox noise of (seed, row, slot)."""
    n, seed = 64, 99
    chol = so.pack_chol(np.linalg.cholesky(np.array([[4e-2, 1e-3, 2e-4], [1e-3, 2e-3, 1e-4], [2e-4, 1e-4, 5e-4]])))
    ops = [dict(type=so.SE2_PRIOR, row_lo=0, row_hi=n, in_a=-1, in_b=-1, out=0, n_out=3, slot=0, obs=[1.0, 2.0, 0.5], chol=chol),
           dict(type=so.SE2_GEN_FWD, row_lo=0, row_hi=n, in_a=0, in_b=-1, out=3, n_out=3, slot=2, obs=[5.0, 0.0, 0.3], chol=chol),
           dict(type=so.RANGE_GEN, row_lo=0, row_hi=40, in_a=3, in_b=-1, out=6, n_out=2, slot=4, obs=[4.0, 0, 0], chol=[0.2, 0, 0, 0, 0, 0]),
           dict(type=so.RANGE_GEN, row_lo=40, row_hi=n, in_a=0, in_b=-1, out=6, n_out=2, slot=4, obs=[4.0, 0, 0], chol=[2.0, 0, 0, 0, 0, 0]),
           dict(type=so.RANGE_OBS, row_lo=0, row_hi=n, in_a=0, in_b=6, out=8, n_out=1, slot=6, obs=[0, 0, 0], chol=[0.2, 0, 0, 0, 0, 0])]
    s = so.simulate(ops, seed, n, 9)
    rows = np.arange(n)
    e0, e1 = so.normal2(seed, rows, 0)
    e2, _ = so.normal2(seed, rows, 1)
    lie = np.column_stack([e0, e1, e2]) @ so.chol_matrix(chol).T
    np.testing.assert_allclose(s[:, :3], so.se2_prior([1.0, 2.0, 0.5], lie), rtol=0, atol=1e-13)
    d = np.hypot(s[:, 6] - s[:, 0], s[:, 7] - s[:, 1])
    np.testing.assert_allclose(s[:, 8], d + 0.2 * so.normal2(seed, rows, 6)[0], rtol=0, atol=1e-13)
    assert np.std(np.hypot(s[:40, 6] - s[:40, 3], s[:40, 7] - s[:40, 4]) - 4.0) < 0.4
