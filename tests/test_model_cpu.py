"""Host-side parity of the solver-facing wrapper against the reference's own outputs (tests/golden/model.npz):
parameter initialisation draw-for-draw, training-sample normalisation, and -- through the oracle backend --
the RNG consumption / column slicing of conditional sampling.  CPU only."""
import os

import numpy as np
import pytest
import torch

from tests.oracle_backend import oracle_backend

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(HERE, "golden", "model.npz")))


def test_initialisation_matches_reference_draw_for_draw(g):
    from nfisam_b200.flows import NSF_AR

    for (d, K, H, seed) in ((5, 9, 8, 11), (14, 12, 8, 12)):
        torch.manual_seed(seed)
        f = NSF_AR(dim=d, K=K, hidden_dim=H)
        assert np.array_equal(f.flat_parameters(), g[f"init_d{d}_K{K}_H{H}_s{seed}"])


def test_normalize_training_samples(g):
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs

    data, means, stds = NFiSAM(NFiSAMArgs()).normalize_training_samples(g["norm_raw"].copy(), list(g["norm_circ"]), "NSF_AR")
    assert np.array_equal(means.numpy(), g["norm_means"]) and np.array_equal(stds.numpy(), g["norm_stds"])
    assert np.array_equal(data.numpy(), g["norm_data"])
    assert stds.numpy()[6] == np.float32(1e-5)                  # clipped std of the collapsed column


def build_model(g):
    from nfisam_b200.flows import NSF_AR, CustomMultivariateNormal
    from nfisam_b200.slam.nfisam import NormalizingFlowModelWithSeparator

    d = 9
    flow = NSF_AR(dim=d, K=9, hidden_dim=8)
    flow.load_flat_parameters(g["wrap_theta"])
    return NormalizingFlowModelWithSeparator([flow], CustomMultivariateNormal(dim=d), CustomMultivariateNormal(dim=6),
                                             list(g["norm_circ"]), torch.tensor(g["norm_means"]), torch.tensor(g["norm_stds"]))


def check_wrapper(g, rtol):
    model = build_model(g)
    torch.manual_seed(21)
    got = model.conditional_sample_given_observation(conditional_dim=3, obs_samples=g["wrap_obs"].copy())
    assert got.dtype == np.float32 and np.allclose(got, g["wrap_cond"], rtol=rtol, atol=2e-4)
    torch.manual_seed(22)
    got = model.conditional_sample_given_observation(conditional_dim=2, obs_samples=g["wrap_obs"][:, :4].copy())
    assert np.allclose(got, g["wrap_cond_prefix"], rtol=rtol, atol=2e-4)
    torch.manual_seed(23)
    got = model.conditional_sample_given_observation(conditional_dim=5, sample_number=32)
    assert np.allclose(got, g["wrap_uncond"], rtol=rtol, atol=2e-4)
    z, plp, ld = model.separator_forward(torch.tensor(np.float32(g["wrap_obs"])))
    assert np.allclose(z.numpy(), g["wrap_sepfwd_z"], rtol=rtol, atol=1e-4)
    assert np.allclose(plp.numpy(), g["wrap_sepfwd_plp"], rtol=rtol, atol=1e-3)
    assert np.allclose(ld.numpy(), g["wrap_sepfwd_ld"], rtol=rtol, atol=1e-3)


def test_wrapper_host_logic_with_oracle_backend(g):
    with oracle_backend():
        check_wrapper(g, rtol=1e-4)


def test_host_shortcuts_are_bit_identical_to_the_reference_calls():
    """np.random.permutation indexing == in-place np.random.shuffle (same rows, same RNG state afterwards); the
    restated circular mean == scipy.stats.circmean."""
    from scipy.stats import circmean as scipy_circmean

    from nfisam_b200.slam.nfisam import circmean

    np.random.seed(7)
    a = np.random.standard_normal((500, 7))
    b = a.copy()
    np.random.seed(11)
    np.random.shuffle(a)
    after_a = np.random.random()
    np.random.seed(11)
    b = b[np.random.permutation(b.shape[0])]
    after_b = np.random.random()
    assert np.array_equal(a, b) and after_a == after_b
    rng = np.random.default_rng(0)
    for _ in range(30):
        x = rng.standard_normal((int(rng.integers(5, 2000)), int(rng.integers(1, 5)))) * rng.uniform(0.1, 3) + rng.uniform(-4, 4)
        assert np.array_equal(circmean(x, high=np.pi, low=-np.pi, axis=0), scipy_circmean(x, high=np.pi, low=-np.pi, axis=0))


def test_conditioner_modules_are_inspectable():
    """flow.layers[i](x) works like in the reference (src/flows/flows.py:26-41, 77-83): the lazily materialised FCNN mirrors hold the
    flow's parameters, so a conditioner evaluated through torch equals W3 tanh(W2 tanh(W1 x + b1) + b2) + b3 of the flat vector."""
    from nfisam_b200.flows import NSF_AR

    torch.manual_seed(5)
    flow = NSF_AR(dim=4, K=5, hidden_dim=8)
    theta = flow.flat_parameters()
    x = torch.randn(7, 2)
    out = flow.layers[1](x)                      # conditioner of dim 2: two inputs
    P, H = 14, 8
    off = P + (H * 1 + H + H * H + H + P * H + P)
    W1 = theta[off:off + 2 * H].reshape(H, 2); off += 2 * H
    b1 = theta[off:off + H]; off += H
    W2 = theta[off:off + H * H].reshape(H, H); off += H * H
    b2 = theta[off:off + H]; off += H
    W3 = theta[off:off + P * H].reshape(P, H); off += P * H
    b3 = theta[off:off + P]
    want = np.tanh(np.tanh(x.numpy() @ W1.T + b1) @ W2.T + b2) @ W3.T + b3
    assert out.shape == (7, P) and np.allclose(out.detach().numpy(), want, atol=1e-5)
