"""GPU parity tests of the flow kernels (through the C ABI via the Python drop-ins) against
 (a) golden vectors produced by the reference's own PyTorch flow (tests/golden/flow_*.npz),
 (b) the CPU oracle (oracle/nsf_oracle.c) on larger seeded inputs,
 (c) size-independent properties (forward/inverse round trip, layout permutation, loss invariance).

Tolerances (float32 flow, north_star: 1e-5 relative): max-norm relative error <= 1e-5 for z / x against the reference's float32
results and the oracle; per-sample log-determinants (sums of d terms of either sign) by allclose(rtol=1e-5, atol=5e-5 .. 1e-4)
here and by the 2e-5 max-norm bar of the error-table tests below, which also print and record the achieved numbers
(profiles/r2_flow_error_table.md).  Scalar losses: 2e-5 relative (float32 mean over n d terms)."""
import numpy as np
import pytest
import torch

from oracle import nsf_oracle as orc

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def make_flow(c, **kw):
    from nfisam_b200.flows import NSF_AR

    d, K, H, B = int(c["d"]), int(c["K"]), int(c["H"]), float(c["B"])
    f = NSF_AR(dim=d, K=K, B=B, hidden_dim=H, **kw)
    f.load_flat_parameters(c["theta"])
    return f


def test_library_loads_and_sees_device():
    from nfisam_b200 import _lib

    assert _lib.require_device() >= 1


def test_param_roundtrip(flow_cases):
    for name, c in flow_cases.items():
        f = make_flow(c)
        f.handle()
        f.load_flat_parameters(np.zeros_like(c["theta"]))
        f._synced = f._param_version()  # pretend in sync, then pull the device copy
        f.pull_parameters()
        assert np.array_equal(f.flat_parameters(), c["theta"]), name


def test_forward_reference_layout_golden(flow_cases):
    for name, c in flow_cases.items():
        f = make_flow(c)
        z, ld = f.forward(torch.tensor(c["x"]))
        assert z.shape == c["z_ref"].shape and not z.is_cuda
        assert _relmax(z.numpy(), c["z_ref"]) < 1e-5, name
        assert np.allclose(ld.numpy(), c["ld_ref"], rtol=1e-5, atol=1e-4), name


def test_forward_per_sample_golden(flow_cases):
    for name, c in flow_cases.items():
        f = make_flow(c, reference_layout=False)
        z, ld = f.forward(torch.tensor(c["x"]).cuda())
        assert z.is_cuda
        assert _relmax(z.cpu().numpy(), c["z_col"]) < 1e-5, name
        assert np.allclose(ld.cpu().numpy(), c["ld_col"], rtol=1e-5, atol=5e-5), name


def test_model_forward_golden(flow_cases):
    from nfisam_b200.flows import CustomMultivariateNormal, NormalizingFlowModel

    for name, c in flow_cases.items():
        f = make_flow(c)
        m = NormalizingFlowModel(CustomMultivariateNormal(dim=int(c["d"])), [f])
        z, plp, ld = m(torch.tensor(c["x"]))
        assert np.allclose(plp.numpy(), c["prior_lp"], rtol=1e-5, atol=1e-4), name
        loss = -(plp + ld).double().mean().item()
        assert abs(loss - float(c["loss"])) < 2e-5 * abs(float(c["loss"])), name
        lp = m.log_prob(torch.tensor(c["x"]))
        assert abs(-lp.double().mean().item() - float(c["loss"])) < 2e-5 * abs(float(c["loss"])), name


def test_prefix_forward_matches_oracle(flow_cases):
    c = flow_cases["d11_K9_H8"]
    f = make_flow(c, reference_layout=False)
    d, K, H, B = 11, 9, 8, 5.0
    for d_in in (1, 4, 10):
        x = c["x"][:, :d_in].copy()
        z, ld = f.forward(torch.tensor(x))
        zo, ldo = orc.forward(c["theta"], d, K, H, B, x)
        assert _relmax(z.numpy(), zo) < 1e-5
        assert np.allclose(ld.numpy(), ldo, rtol=1e-5, atol=5e-5)
        lp = f.log_prob(torch.tensor(x))
        assert np.allclose(lp.numpy(), orc.log_prob(c["theta"], d, K, H, B, x), rtol=1e-5, atol=1e-4)


def test_inverse_golden(flow_cases):
    for name, c in flow_cases.items():
        f = make_flow(c)
        x, ld = f.inverse(torch.tensor(c["zin"]))
        assert _relmax(x.numpy(), c["x_inv"]) < 1e-5, name
        assert np.allclose(ld.numpy(), c["ld_inv"], rtol=1e-5, atol=1e-4), name
        sep = int(c["sep"])
        xc = f.inverse_given_separator(torch.tensor(c["zin_f"]), torch.tensor(c["x_sep"]) if sep else None)
        assert _relmax(xc.numpy(), c["x_cond"]) < 1e-5, name


def test_inverse_fused_normalisation(flow_cases):
    c = flow_cases["d11_K9_H8"]
    f = make_flow(c)
    d, sep = 11, int(c["sep"])
    rng = np.random.default_rng(3)
    mean = rng.normal(size=d).astype(np.float32)
    std = (0.5 + rng.random(d)).astype(np.float32)
    circ = np.zeros(d, np.uint8)
    circ[[2, 7]] = 1
    raw_sep = (c["x_sep"] * std[:sep] + mean[:sep]).astype(np.float32)
    raw_sep[:, 2] += 2 * np.pi  # wrapped away by the normalisation
    out = f.inverse_given_separator(torch.tensor(c["zin_f"]), torch.tensor(raw_sep), norm=(mean, std, circ)).numpy()
    # host restatement of NFiSAM.py:96-118 around the un-normalised kernel
    def wrap(t):
        return (t + np.float32(np.pi)) % np.float32(2 * np.pi) - np.float32(np.pi)
    xs = raw_sep - mean[:sep]
    xs[:, 2] = wrap(xs[:, 2])
    xs = xs / std[:sep]
    plain = f.inverse_given_separator(torch.tensor(c["zin_f"]), torch.tensor(xs)).numpy()
    exp = plain * std[sep:] + mean[sep:]
    exp[:, 7 - sep] = wrap(exp[:, 7 - sep])
    assert np.allclose(out, exp, rtol=1e-5, atol=2e-5)


def test_loss_and_gradient_golden(flow_cases):
    for name, c in flow_cases.items():
        f = make_flow(c)
        loss, g = f.loss_and_grad(torch.tensor(c["x"]))
        assert abs(loss - float(c["loss"])) < 2e-5 * abs(float(c["loss"])), name
        gr = c["grad"]
        assert _relmax(g, gr) < 2e-4, (name, _relmax(g, gr))
        assert np.allclose(g, gr, rtol=1e-3, atol=1e-5 * np.max(np.abs(gr))), name


def test_adam_trajectory_golden(flow_cases):
    for name, c in flow_cases.items():
        if "adam_loss" not in c:
            continue
        f = make_flow(c)
        steps = len(c["adam_loss"])
        hist, ran = f.fit(torch.tensor(c["x"]), steps, float(c["adam_lr"]), average_window=0)
        assert ran == steps
        assert np.allclose(hist[:10], c["adam_loss"][:10], rtol=2e-5, atol=1e-5), (name, hist, c["adam_loss"])
        assert np.allclose(hist, c["adam_loss"], rtol=3e-2), (name, hist, c["adam_loss"])
        assert _relmax(f.flat_parameters(), c["adam_theta"]) < 5e-2, name


@pytest.mark.parametrize("n", [1, 7, 33, 2000, 100003, 200001])
def test_ragged_sizes_vs_oracle(flow_cases, n):
    c = flow_cases["d6_K9_H8"]
    d, K, H, B = 6, 9, 8, 5.0
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((n, d)) * 1.5).astype(np.float32)
    x[rng.random((n, d)) < 0.005] *= 8.0
    f = make_flow(c, reference_layout=False)
    z, ld = f.forward(torch.tensor(x))
    zo, ldo = orc.forward(c["theta"], d, K, H, B, x)
    assert _relmax(z.numpy(), zo) < 1e-5
    assert np.allclose(ld.numpy(), ldo, rtol=1e-5, atol=5e-5)
    zin = rng.standard_normal((n, d)).astype(np.float32)
    xi, ldi = f.inverse(torch.tensor(zin))
    xo, ldo2, bad = orc.inverse(c["theta"], d, K, H, B, zin)
    assert bad == 0
    assert _relmax(xi.numpy(), xo) < 1e-5
    assert np.allclose(ldi.numpy(), ldo2, rtol=1e-5, atol=1e-4)
    loss, g = f.loss_and_grad(torch.tensor(x))
    lo, go = orc.loss_grad(c["theta"], d, K, H, B, x, dtype=np.float64)
    assert abs(loss - lo) < 2e-5 * abs(lo)
    assert _relmax(g, go) < 5e-4


def test_empty_batch(flow_cases):
    c = flow_cases["d4_K5_H8"]
    f = make_flow(c)
    z, ld = f.forward(torch.zeros((0, 4)))
    assert z.shape == (0, 4) and ld.shape == (0,)


def test_error_vs_f64_oracle_not_worse_than_reference(flow_cases):
    """kernel-vs-fp64 error <= 2 x (reference fp32-vs-fp64 error) + eps, on the golden inputs."""
    for name, c in flow_cases.items():
        d, K, H, B = int(c["d"]), int(c["K"]), int(c["H"]), float(c["B"])
        z64, ld64 = orc.forward(c["theta"], d, K, H, B, c["x"], dtype=np.float64)
        f = make_flow(c, reference_layout=False)
        z, ld = f.forward(torch.tensor(c["x"]))
        ref_err = max(np.max(np.abs(c["z_col"] - z64)), 1e-7)
        assert np.max(np.abs(z.numpy() - z64)) <= 2 * ref_err + 2e-6, name
        ref_err = max(np.max(np.abs(c["ld_col"] - ld64)), 1e-7)
        assert np.max(np.abs(ld.numpy() - ld64)) <= 2 * ref_err + 1e-5, name


def test_round_trip_large():
    """1e6 samples, d=12: inverse(forward(x)) == x and log-dets cancel (size-independent property)."""
    from nfisam_b200.flows import NSF_AR

    torch.manual_seed(0)
    f = NSF_AR(dim=12, K=9, hidden_dim=8, reference_layout=False)
    n = 1_000_000
    x = torch.randn(n, 12, device="cuda") * 1.2
    z, ld = f.forward(x)
    xb, ldi = f.inverse(z)
    inside = (x.abs() < 4.99).all(1) & (z.abs() < 4.99).all(1)
    assert inside.float().mean() > 0.9
    assert (xb - x)[inside].abs().max().item() < 2e-4
    assert (ld + ldi)[inside].abs().max().item() < 2e-3
    # layout permutation property: reference layout is the transpose buffer of the per-sample result
    zr, ldr = f.forward(x[:1000], reference_layout=True)
    assert torch.equal(zr, z[:1000].t().contiguous().view(1000, 12))
    assert abs(ldr.double().sum().item() - ld[:1000].double().sum().item()) < 1e-2


def test_host_buffer_api_matches_device_api(flow_cases):
    import ctypes

    from nfisam_b200 import _lib

    c = flow_cases["d11_K9_H8"]
    f = make_flow(c, reference_layout=False)
    n = 300_001
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 11)).astype(np.float32)
    lp_dev = f.log_prob(torch.tensor(x)).numpy()
    xh = torch.tensor(x).pin_memory()
    out = torch.empty(n).pin_memory()
    _lib.check(_lib.load().nfisam_flow_log_prob_host(f.handle(), xh.data_ptr(), n, 11, out.data_ptr()))
    assert np.array_equal(out.numpy(), lp_dev)
    z = rng.standard_normal((n, 6)).astype(np.float32)
    xs = x[:, :5].copy()
    ref = f.inverse_given_separator(torch.tensor(z), torch.tensor(xs)).numpy()
    o2 = np.empty((n, 6), np.float32)
    _lib.check(_lib.load().nfisam_flow_inverse_host(f.handle(), z.ctypes.data_as(ctypes.c_void_p),
                                                    xs.ctypes.data_as(ctypes.c_void_p), n, 5, 6,
                                                    o2.ctypes.data_as(ctypes.c_void_p), None, None, None))
    assert np.array_equal(o2, ref)


def test_training_matches_oracle_and_early_stop():
    """200 Adam iterations on a 2000-sample banana: loss curve tracks the CPU oracle; the windowed
    early stop fires at the same iteration for a plateaued run."""
    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(5)
    d, K, H = 5, 9, 8
    x = rng.standard_normal((2000, d)).astype(np.float32)
    x[:, 1] = 0.5 * x[:, 1] + x[:, 0] ** 2 - 1.0
    x = (x - x.mean(0)) / x.std(0)
    torch.manual_seed(1)
    f = NSF_AR(dim=d, K=K, hidden_dim=H)
    theta0 = f.flat_parameters()
    hist, ran = f.fit(torch.tensor(x), 200, 0.01, average_window=0)
    th_o, hist_o, it_o = orc.train(theta0, d, K, H, 5.0, x, 200, 0.01, average_window=0)
    assert ran == 200 and it_o == 200
    # float32 round-off is amplified along the trajectory: the kernel must stay as close to the
    # float64 oracle as the float32 oracle (= the reference's arithmetic) does, within a small factor
    _, hist64, _ = orc.train(theta0, d, K, H, 5.0, x, 200, 0.01, average_window=0, dtype=np.float64)
    assert np.allclose(hist[:10], hist_o[:10], rtol=1e-5)
    err_gpu = np.abs(hist - hist64)
    err_f32 = np.abs(hist_o - hist64)
    assert err_gpu.max() <= 5 * err_f32.max() + 1e-4, (err_gpu.max(), err_f32.max())
    assert np.allclose(hist, hist_o, rtol=2e-2), np.max(np.abs(hist - hist_o))
    assert hist[-1] < hist[0] - 0.1
    # continuing training keeps the Adam state (bias correction continues)
    f2 = NSF_AR(dim=d, K=K, hidden_dim=H)
    f2.load_flat_parameters(theta0)
    h1, _ = f2.fit(torch.tensor(x), 100, 0.01, average_window=0)
    h2, _ = f2.fit(torch.tensor(x), 100, 0.01, average_window=0, reset_optimizer=False)
    assert np.array_equal(h1, hist[:100])
    assert np.allclose(h2, hist[100:], rtol=2e-3)             # bias-correction powers are rebuilt per launch
    # early stop: tiny lr -> second window mean within tolerance of the first
    f3 = NSF_AR(dim=d, K=K, hidden_dim=H)
    f3.load_flat_parameters(theta0)
    h3, ran3 = f3.fit(torch.tensor(x), 400, 1e-6, average_window=10, loss_delta_tol=1e-2)
    _, h3o, ran3o = orc.train(theta0, d, K, H, 5.0, x, 400, 1e-6, average_window=10, loss_delta_tol=1e-2)
    assert ran3 == ran3o == 20
    assert np.all(h3[20:] == 0)
    # and a real run with the default window stops where the oracle stops
    f4 = NSF_AR(dim=d, K=K, hidden_dim=H)
    f4.load_flat_parameters(theta0)
    h4, ran4 = f4.fit(torch.tensor(x), 2000, 0.02, average_window=50, loss_delta_tol=1e-2)
    _, h4o, ran4o = orc.train(theta0, d, K, H, 5.0, x, 2000, 0.02, average_window=50, loss_delta_tol=1e-2)
    assert ran4 % 50 == 0 and ran4 < 2000
    assert abs(ran4 - ran4o) <= 50, (ran4, ran4o)


def test_training_is_bitwise_reproducible():
    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(9)
    x = torch.tensor(rng.standard_normal((1500, 7)).astype(np.float32))
    outs = []
    for _ in range(2):
        torch.manual_seed(3)
        f = NSF_AR(dim=7, K=9, hidden_dim=8)
        hist, ran = f.fit(x, 60, 0.02, average_window=0)
        outs.append((hist.copy(), f.flat_parameters()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


def test_co_resident_training_is_bit_identical_and_overlaps():
    """nf_train_cfg.concurrency >= 2 launches the two-blocks-per-SM build of the cluster kernel (several cliques of a
    tree level in flight on one GPU): same arithmetic, so loss curve and parameters are bit-identical to the default
    build; eight runs on eight streams are timed both ways and reported."""
    import time

    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(4)
    d, n, iters = 15, 2000, 300
    xs = [torch.tensor(rng.standard_normal((n, d)).astype(np.float32)).cuda() for _ in range(8)]
    streams = [torch.cuda.Stream() for _ in range(8)]
    results, times = {}, {}
    for conc in (1, 8):
        flows = []
        for k in range(8):
            torch.manual_seed(20 + k)
            flows.append(NSF_AR(dim=d, K=9, hidden_dim=8))
            flows[-1].handle()
        for rep in range(2):                               # second repetition is the timed one
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for f, x, st in zip(flows, xs, streams):
                f.fit_launch(x, iters, 0.02, average_window=0, stream=st, concurrency=conc)
            outs = [f.fit_finish(pull=True) for f in flows]
            times[conc] = (time.perf_counter() - t0) * 1e3
        results[conc] = [(h.copy(), f.flat_parameters()) for (h, _), f in zip(outs, flows)]
    print(f"\n8 cliques x {iters} Adam iterations (n={n}, d={d}) on 8 streams: default build {times[1]:.2f} ms, "
          f"co-resident build {times[8]:.2f} ms")
    for (h1, p1), (h8, p8) in zip(results[1], results[8]):
        assert np.array_equal(h1, h8)
        assert np.array_equal(p1, p8)


def test_unsupported_configuration_fails_loudly():
    from nfisam_b200 import _lib
    from nfisam_b200.flows import NSF_AR

    # outside the generic kernels' range (K <= 64, hidden <= 64) and beyond NFISAM_MAX_DIM: a named error, never a fallback
    for kw, d in ((dict(K=65, hidden_dim=8), 3), (dict(K=9, hidden_dim=65), 3), (dict(K=9, hidden_dim=8), 33)):
        f = NSF_AR(dim=d, **kw)
        with pytest.raises(_lib.NfisamError):
            f.forward(torch.zeros(4, d))


def test_large_batch_training_mode_matches_oracle():
    """n >= 16384 switches the training loop to the plain-grid mode (per-block partial gradients + Adam kernel):
    same loss curve as the oracle, same early-stop iteration, continuation keeps the Adam state."""
    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(11)
    d, K, H, n = 5, 9, 8, 20_000
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[:, 2] = 0.4 * x[:, 2] + np.tanh(x[:, 0]) - 0.5 * x[:, 1] ** 2 + 0.5
    x = (x - x.mean(0)) / x.std(0)
    torch.manual_seed(2)
    f = NSF_AR(dim=d, K=K, hidden_dim=H)
    theta0 = f.flat_parameters()
    hist, ran = f.fit(torch.tensor(x), 40, 0.01, average_window=0)
    _, hist_o, _ = orc.train(theta0, d, K, H, 5.0, x, 40, 0.01, average_window=0)
    _, hist64, _ = orc.train(theta0, d, K, H, 5.0, x, 40, 0.01, average_window=0, dtype=np.float64)
    assert ran == 40
    assert np.allclose(hist[:10], hist_o[:10], rtol=1e-5)
    assert np.abs(hist - hist64).max() <= 5 * np.abs(hist_o - hist64).max() + 1e-4
    f2 = NSF_AR(dim=d, K=K, hidden_dim=H)
    f2.load_flat_parameters(theta0)
    h1, _ = f2.fit(torch.tensor(x), 20, 0.01, average_window=0)
    h2, _ = f2.fit(torch.tensor(x), 20, 0.01, average_window=0, reset_optimizer=False)
    assert np.array_equal(h1, hist[:20]) and np.allclose(h2, hist[20:], rtol=1e-4)
    f3 = NSF_AR(dim=d, K=K, hidden_dim=H)
    f3.load_flat_parameters(theta0)
    h3, ran3 = f3.fit(torch.tensor(x), 100, 1e-6, average_window=10, loss_delta_tol=1e-2)
    assert ran3 == 20 and np.all(h3[20:] == 0) and np.all(h3[:20] != 0)
    loss, g = f3.loss_and_grad(torch.tensor(x))
    lo, go = orc.loss_grad(f3.flat_parameters(), d, K, H, 5.0, x, dtype=np.float64)
    assert abs(loss - lo) < 2e-5 * abs(lo) and _relmax(g, go) < 5e-4


def test_large_batch_training_at_multi_robot_clique_dimension():
    """Flow dimension 17 (the cliques of the multi-robot graphs) in the large-batch mode: the block slots are split over 17 dims by
    cost, the blocks past a dim's share only write a zero partial, and the per-warp gradient partials alias the staging regions so
    that two blocks fit per SM.  Loss curve and gradient against the oracle, and bitwise reproducibility of the run."""
    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(17)
    d, K, H, n = 17, 9, 8, 16_500
    x = rng.standard_normal((n, d)).astype(np.float32)
    for i in range(1, d):
        x[:, i] = 0.6 * x[:, i] + 0.5 * np.tanh(x[:, i - 1]) ** 2
    x = (x - x.mean(0)) / x.std(0)
    torch.manual_seed(4)
    f = NSF_AR(dim=d, K=K, hidden_dim=H)
    theta0 = f.flat_parameters()
    loss, g = f.loss_and_grad(torch.tensor(x))
    lo, go = orc.loss_grad(theta0, d, K, H, 5.0, x, dtype=np.float64)
    assert abs(loss - lo) < 2e-5 * abs(lo) and _relmax(g, go) < 5e-4
    hist, ran = f.fit(torch.tensor(x), 8, 0.01, average_window=0)
    _, hist_o, _ = orc.train(theta0, d, K, H, 5.0, x, 8, 0.01, average_window=0)
    assert ran == 8 and np.allclose(hist, hist_o, rtol=2e-5)
    f2 = NSF_AR(dim=d, K=K, hidden_dim=H)
    f2.load_flat_parameters(theta0)
    hist2, _ = f2.fit(torch.tensor(x), 8, 0.01, average_window=0)
    assert np.array_equal(hist, hist2) and np.array_equal(f.flat_parameters(), f2.flat_parameters())


def test_validation_set_slower_stop_matches_reference_loop():
    """training_set_frac < 1 path: validation loss every `validation_interval` iterations, first increase fixes
    slower_stop_iter = int(rate * (i + 1)) (src/slam/NFiSAM.py:452-468).  Golden: the reference's loop body run with
    the reference flow (tests/golden/train_val.npz, produced next to oracle.train_val's cross-check)."""
    import os

    from nfisam_b200.flows import NSF_AR

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_val.npz")))
    f = NSF_AR(dim=4, K=9, hidden_dim=8)
    f.load_flat_parameters(g["theta"])
    hist, ran = f.fit(torch.tensor(g["x"]), int(g["iters"]), float(g["lr"]), val=torch.tensor(g["xv"]),
                      validation_interval=int(g["vi"]), slower_stop_rate=float(g["rate"]))
    assert ran == int(g["ran"]) == 39
    # lr = .05 on 40 samples is a deliberately unstable regime (the loss jumps up at iteration 1): float32
    # round-off is amplified quickly, so the curve is compared tightly only over the first iterations
    assert np.allclose(hist[:9], g["hist"][:9], rtol=2e-5)
    assert np.allclose(hist[:ran], g["hist"][:ran], rtol=5e-2)
    assert np.all(hist[ran:] == 0)
    # the validation passes do not disturb the training trajectory: same curve as a run without validation
    f1 = NSF_AR(dim=4, K=9, hidden_dim=8)
    f1.load_flat_parameters(g["theta"])
    hist1, _ = f1.fit(torch.tensor(g["x"]), 39, float(g["lr"]), average_window=0)
    assert np.array_equal(hist1, hist[:39])
    # no increase of the validation loss within the budget: runs to the end
    f2 = NSF_AR(dim=4, K=9, hidden_dim=8)
    f2.load_flat_parameters(g["theta"])
    x_big = torch.tensor(np.random.default_rng(1).standard_normal((4000, 4)).astype(np.float32))
    hist2, ran2 = f2.fit(x_big[:3000], 60, 0.005, val=x_big[3000:], validation_interval=10, slower_stop_rate=2.0)
    _, ho, rano, _ = orc.train_val(g["theta"], 4, 9, 8, 5.0, x_big[:3000].numpy(), x_big[3000:].numpy(), 60, 0.005,
                                   validation_interval=10, slower_stop_rate=2.0)
    assert ran2 == rano
    assert np.allclose(hist2[:10], ho[:10], rtol=2e-5)


def _random_clique_tree(seed, trunk, branches, depth, K=9, cross=False):
    """A Bayes-tree-shaped list of posterior-pass items with randomly initialised flows: `trunk` cliques in a chain, then
    `branches` chains of `depth` cliques below the last trunk clique.  Every clique has one constant observation column,
    the 3 frontal columns of its parent (and 2 of its grandparent) as given columns, 3 frontal columns of its own."""
    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    items, cols, total, zoff = [], {}, 0, 0

    def add(name, parent, grand):
        nonlocal total, zoff
        sep_cols, sep_const = [-1], [float(rng.normal())]
        if parent is not None:
            sep_cols += cols[parent]
            sep_const += [0.0] * 3
        if grand is not None:
            sep_cols += cols[grand][:2]
            sep_const += [0.0] * 2
        d = len(sep_cols) + 3
        flow = NSF_AR(dim=d, K=K, hidden_dim=8)
        mean = rng.normal(size=d).astype(np.float32)
        std = rng.uniform(0.5, 2.0, size=d).astype(np.float32)
        circ = (rng.uniform(size=d) < 0.3).astype(np.uint8)
        cols[name] = [total, total + 1, total + 2]
        total += 3
        items.append((flow, zoff, sep_cols, sep_const, cols[name], (mean, std, circ)))
        zoff += 3

    prev, prev2 = None, None
    for t in range(trunk):
        add(("t", t), prev, prev2)
        prev, prev2 = ("t", t), prev
    for b in range(branches):
        p, g = prev, prev2
        for k in range(depth):
            add((b, k), p, g)
            p, g = (b, k), p
    if cross and branches >= 2:
        # not a forest: the last clique of branch 1 also reads a column generated inside branch 0
        flow, z0, sc, sk, oc, norm = items[-1]
        sc = list(sc)
        sc[-1] = cols[(0, depth - 1)][0]
        items[-1] = (flow, z0, sc, sk, oc, norm)
    return items, total, zoff


@pytest.mark.parametrize("shape", [(3, 0, 0, False), (2, 3, 4, False), (0, 2, 3, False), (1, 3, 2, True)])
@pytest.mark.parametrize("n", [1, 1000])
def test_posterior_pass_equals_per_clique_launches(shape, n):
    """nfisam_posterior_pass (trunk launch + concurrent subtree launch, warp = 32 rows walking its cliques) is
    bit-identical to one nfisam_flow_inverse_gather per clique, for chains, branching trees, forests and for column
    dependencies that are not a forest (per-item fallback)."""
    from nfisam_b200.flows import posterior_pass

    trunk, branches, depth, cross = shape
    items, total, zw = _random_clique_tree(11, trunk, branches, depth, cross=cross)
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    z = torch.randn((n, zw), generator=g).to(dev)
    z[0, 0] = 7.5                                        # linear tail
    a = torch.zeros((n, total), device=dev)
    b = torch.zeros((n, total), device=dev)
    ca = torch.zeros(1, dtype=torch.int64, device=dev)
    cb = torch.zeros(1, dtype=torch.int64, device=dev)
    for flow, z0, sc, sk, oc, norm in items:
        flow.inverse_gather(z, z0, a, sc, sk, oc, norm=norm, counter=ca)
    posterior_pass(items, z, b, counter=cb)
    torch.cuda.synchronize()
    assert int(ca.item()) == 0 and int(cb.item()) == 0
    assert torch.isfinite(a).all()
    assert float(a.abs().max()) > 0
    assert torch.equal(a, b)


def test_posterior_pass_edge_cases():
    """No rows / no cliques are no-ops; bad column lists fail loudly with a status code, nothing is launched."""
    from nfisam_b200 import _lib
    from nfisam_b200.flows import posterior_pass

    items, total, zw = _random_clique_tree(5, 2, 0, 0)
    dev = torch.device("cuda")
    z = torch.randn((0, zw), device=dev)
    S = torch.zeros((0, total), device=dev)
    posterior_pass(items, z, S)                                    # n = 0
    z = torch.randn((8, zw), device=dev)
    S = torch.zeros((8, total), device=dev)
    posterior_pass([], z, S)                                       # no cliques
    assert float(S.abs().sum()) == 0.0
    flow, z0, sc, sk, oc, norm = items[1]
    with pytest.raises(_lib.NfisamError):
        posterior_pass([items[0], (flow, z0, sc, sk, [total + 3, 1, 2], norm)], z, S)      # output column out of range
    with pytest.raises(_lib.NfisamError):
        posterior_pass([items[0], (flow, zw, sc, sk, oc, norm)], z, S)                     # latent columns out of range


def test_large_batch_log_prob_pair_kernel_matches_single_sample_kernel():
    """Batches of >= 5e5 rows take nf_log_prob_pair_kernel (two samples per thread, shared weight loads); smaller ones
    nf_forward_kernel.  Same arithmetic per sample: a large batch must reproduce, bit for bit, what the same rows give in
    small batches; both agree with the CPU oracle on a subset."""
    from nfisam_b200.flows import NSF_AR

    torch.manual_seed(1)
    d, n = 12, 600_011                                     # odd size: ragged last tile, unpaired last row
    f = NSF_AR(dim=d, K=9, hidden_dim=8, reference_layout=False)
    x = torch.randn(n, d) * 2.0
    x[::101] *= 4.0                                        # linear tails
    xd = x.cuda()
    big = f.log_prob(xd)
    small = torch.cat([f.log_prob(xd[a:a + 100_000].contiguous()) for a in range(0, n, 100_000)])
    assert torch.equal(big, small)
    ref = orc.log_prob(f.flat_parameters(), d, 9, 8, 5.0, x[-3000:].numpy())
    got = big[-3000:].cpu().numpy()
    assert np.allclose(got, ref, rtol=1e-5, atol=2e-4)
    assert _relmax(got, ref) < 1e-5



# ---------------------------------------------------------------------------------------------------------------------
# Achieved error, recorded.  North star: "flow forward / inverse / log-prob within 1e-5 relative (fp32)".
# Relative = max-norm: ||kernel - reference||_inf / ||reference||_inf (element-wise relative error is meaningless next to zero
# crossings: the reference's own float32 result differs from its float64 evaluation by 1e-3 element-wise there).
# Every number is also measured against the float64 oracle next to the reference's own float32 round-off.
# ---------------------------------------------------------------------------------------------------------------------
BAR = 1e-5
# Per-sample log-det alone is held to 2e-5: its max-norm is ~8 where the log-prob's is ~150, and the two float32 evaluations
# (the reference's and ours) EACH sit ~1e-6 x |log-prob| = 5e-5..1e-4 (absolute) from the float64 value in narrow spline bins
# -- the reference's own float32-vs-float64 distance is recorded in the same table.  Measured worst case 1.13e-5 (n = 2000,
# d = 12) / 1.31e-5 (n = 1e5); with -DNF_ACCURATE_MATH=1 (libm instead of the MUFU approximations) the same fixtures give
# 0.99e-5 for the log-det while z / log-prob / inverse get WORSE (1.5e-6 / 9.2e-6 / 3.3e-6 against 1.3e-6 / 5.0e-6 / 0.7e-6):
# the distance is float32 summation order, not the approximations.  Everything else (z, log-prob, inverse, conditional
# inverse) meets 1e-5 with a margin of 2x - 15x (profiles/r2_flow_error.md).
BAR_LOGDET = 2e-5
_error_rows = []


def _record(case, what, got, ref, f64=None, ref_vs_f64=None):
    row = {"case": case, "quantity": what, "n": int(np.asarray(ref).shape[0]), "relmax_vs_reference_f32": _relmax(got, ref),
           "maxabs_vs_reference_f32": float(np.max(np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64))))}
    if f64 is not None:
        row["relmax_vs_f64_oracle"] = _relmax(got, f64)
        row["reference_f32_relmax_vs_f64"] = ref_vs_f64
    _error_rows.append(row)
    return row


def _dump_error_table():
    import json
    import os

    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "flow_error_table.json"), "w") as fh:
            json.dump(_error_rows, fh, indent=1)
    except OSError:
        pass                                   # read-only checkout: the table is still printed below
    print()
    for r in _error_rows:
        print("  %-22s %-14s n=%-7d vs reference f32: %.2e (abs %.2e)%s" % (
            r["case"], r["quantity"], r["n"], r["relmax_vs_reference_f32"], r["maxabs_vs_reference_f32"],
            "   vs f64 oracle: %.2e   (reference's own f32 vs f64: %.2e)" % (r["relmax_vs_f64_oracle"], r["reference_f32_relmax_vs_f64"])
            if "relmax_vs_f64_oracle" in r else ""))


def test_achieved_error_small_fixtures_meet_1e5(flow_cases):
    """Every reference-made small fixture (n = 16-64): forward z / log-det in both layouts, prior log-prob, inverse, conditional
    inverse -- max-norm relative error <= 1e-5 against the reference's float32 outputs."""
    for name, c in flow_cases.items():
        d, K, H, B = int(c["d"]), int(c["K"]), int(c["H"]), float(c["B"])
        x = torch.tensor(c["x"])
        f = make_flow(c)
        z, ld = f.forward(x)
        rows = [_record(name, "z (ref layout)", z.numpy(), c["z_ref"]), _record(name, "logdet (ref)", ld.numpy(), c["ld_ref"])]
        g = make_flow(c, reference_layout=False)
        z64, ld64 = orc.forward(c["theta"].astype(np.float64), d, K, H, B, c["x"].astype(np.float64), dtype=np.float64)
        z, ld = g.forward(x)
        rows.append(_record(name, "z", z.numpy(), c["z_col"], z64, _relmax(c["z_col"], z64)))
        rows.append(_record(name, "logdet", ld.numpy(), c["ld_col"], ld64, _relmax(c["ld_col"], ld64)))
        xi, ldi = f.inverse(torch.tensor(c["zin"]))
        rows.append(_record(name, "inverse x", xi.numpy(), c["x_inv"]))
        rows.append(_record(name, "inverse logdet", ldi.numpy(), c["ld_inv"]))
        sep = int(c["sep"])
        xc = f.inverse_given_separator(torch.tensor(c["zin_f"]), torch.tensor(c["x_sep"]) if sep else None)
        rows.append(_record(name, "conditional x", xc.numpy(), c["x_cond"]))
        for r in rows:
            assert r["relmax_vs_reference_f32"] <= (BAR_LOGDET if r["quantity"].startswith("logdet") else BAR), r


@pytest.mark.parametrize("name", ["n2000_d12_K9_H8", "n2000_d15_K12_H8", "n100000_d12_K9_H8", "n100000_d6_K9_H8"])
def test_achieved_error_large_reference_fixtures_meet_1e5(name):
    """Reference-made fixtures at n = 2000 and n = 1e5 (tests/golden/make_flow_golden_large.py; inputs regenerated from their
    seed and verified by checksum): per-sample log-prob of all n rows, z / log-det / inverse / conditional inverse on every
    64th row -- max-norm relative error <= 1e-5 against the reference's float32 outputs, and the error against the
    reference's own float64 evaluation no larger than 2 x the reference's float32 round-off (+ 1e-6)."""
    import os

    from tests.golden.flow_inputs import checksum, large_inputs

    c = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"flowL_{name}.npz")))
    n, d, K, H, stride = int(c["n"]), int(c["d"]), int(c["K"]), int(c["H"]), int(c["stride"])
    x, zin = large_inputs(n, d, int(c["seed"]))
    assert np.allclose(checksum(x), c["x_sum"], rtol=1e-12) and np.allclose(checksum(zin), c["zin_sum"], rtol=1e-12)
    f = make_flow(c, reference_layout=False)
    lp = f.log_prob(torch.tensor(x)).numpy()
    r = _record(name, "log_prob", lp, c["logp_col"], None)
    assert r["relmax_vs_reference_f32"] <= BAR, r
    sub = slice(None, None, stride)
    r64 = _record(name, "log_prob[::64]", lp[sub], c["logp_col"][sub], c["logp64_sub"], _relmax(c["logp_col"][sub], c["logp64_sub"]))
    assert r64["relmax_vs_f64_oracle"] <= 2.0 * r64["reference_f32_relmax_vs_f64"] + 1e-6, r64
    z, ld = f.forward(torch.tensor(x))
    for what, got, ref, bar in (("z[::64]", z.numpy()[sub], c["z_col_sub"], BAR), ("logdet[::64]", ld.numpy()[sub], c["ld_col_sub"], BAR_LOGDET)):
        r = _record(name, what, got, ref)
        assert r["relmax_vs_reference_f32"] <= bar, r
    fr = make_flow(c)                                   # reference output layout
    zr, ldr = fr.forward(torch.tensor(x))
    r = _record(name, "z (ref layout)[::64]", zr.numpy()[sub], c["z_ref_sub"])
    assert r["relmax_vs_reference_f32"] <= BAR, r
    total = float((fr.log_prob(torch.tensor(x)).double()).sum())     # invariant under the layout permutation
    assert abs(total - float(c["logp_ref_sum"])) <= 1e-6 * abs(float(c["logp_ref_sum"]))
    xi, _ = f.inverse(torch.tensor(zin))
    r = _record(name, "inverse x[::64]", xi.numpy()[sub], c["x_inv_sub"])
    assert r["relmax_vs_reference_f32"] <= BAR, r
    sep = int(c["sep"])
    xc = f.inverse_given_separator(torch.tensor(zin[:, :d - sep].copy()), torch.tensor(x[:, :sep].copy()))
    r = _record(name, "conditional x[::64]", xc.numpy()[sub], c["x_cond_sub"])
    assert r["relmax_vs_reference_f32"] <= BAR, r


def test_zz_error_table_is_written():
    """Runs last in this module: dumps the table of achieved errors (gpurun_out/flow_error_table.json -> profiles/r2_flow_error.md)."""
    assert _error_rows
    _dump_error_table()


# ---------------------------------------------------------------------------------------------------------------------
# Row-sharded training (nfisam_flow_train_launch_sharded): gradient exchange over peer memory fused into the Adam kernel
# ---------------------------------------------------------------------------------------------------------------------
def _banana32(n, d, seed):
    from tests.golden.flow_inputs import banana

    x = banana(n, d, np.random.default_rng(seed))
    return ((x - x.mean(0)) / x.std(0)).astype(np.float32)


def test_sharded_training_with_one_rank_equals_plain_training():
    """A shard group of ONE rank pushes into its own receive area and waits on its own flag: same reduction order as the plain
    large-batch path, so loss curve and parameters are bitwise equal (exercises the fused exchange kernel on a single GPU)."""
    import torch.distributed as dist

    from nfisam_b200.flows import NSF_AR
    from nfisam_b200.flows.flows import ShardGroup

    d, n, iters = 6, 20000, 60
    x = torch.from_numpy(_banana32(n, d, 5)).cuda()
    torch.manual_seed(3)
    a = NSF_AR(dim=d, K=9, hidden_dim=8)
    theta0 = a.flat_parameters()
    hist_a, ran_a = a.fit(x, iters, 0.02, average_window=20, loss_delta_tol=1e-3)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29747", rank=0, world_size=1)
    try:
        group = ShardGroup(dist.group.WORLD, torch.cuda.current_device(), NSF_AR.packed_size(d, 9, 8) + d)
        assert group.rows(n) == (0, n)
        b = NSF_AR(dim=d, K=9, hidden_dim=8)
        b.load_flat_parameters(theta0)
        b.fit_launch(x, iters, 0.02, average_window=20, loss_delta_tol=1e-3, shard=group, n_total=n)
        hist_b, ran_b = b.fit_finish()
        assert not group.timed_out()
        # a second run through the same group: the stamps keep advancing
        b.load_flat_parameters(theta0)
        b.fit_launch(x, iters, 0.02, average_window=20, loss_delta_tol=1e-3, shard=group, n_total=n)
        hist_c, ran_c = b.fit_finish()
    finally:
        dist.destroy_process_group()
    assert ran_a == ran_b == ran_c and np.array_equal(hist_a, hist_b) and np.array_equal(hist_a, hist_c)
    assert np.array_equal(a.flat_parameters(), b.flat_parameters())
    assert hist_a[ran_a - 1] < hist_a[0]


SHARD_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from nfisam_b200.flows import NSF_AR
from nfisam_b200.flows.flows import ShardGroup
from tests.test_flow_gpu import _banana32
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
d, n, iters = 12, 50001, 80
x = torch.from_numpy(_banana32(n, d, 9)).cuda()
torch.manual_seed(4)
flow = NSF_AR(dim=d, K=9, hidden_dim=8, device=rank)
theta0 = flow.flat_parameters()
group = ShardGroup(dist.group.WORLD, rank, NSF_AR.packed_size(d, 9, 8) + d)
r0, r1 = group.rows(n)
out = {{}}
for rep in range(2):
    flow.load_flat_parameters(theta0)
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    flow.fit_launch(x[r0:r1], iters, 0.02, average_window=20, loss_delta_tol=1e-4, shard=group, n_total=n)
    hist, ran = flow.fit_finish()
    ev1.record(); torch.cuda.synchronize()
    out["sharded_ms"] = ev0.elapsed_time(ev1)
assert not group.timed_out()
theta = torch.from_numpy(flow.flat_parameters()).cuda()
both = [torch.empty_like(theta) for _ in range(world)]
dist.all_gather(both, theta)
same = all(bool(torch.equal(both[0], t)) for t in both)
if rank == 0:
    ref = NSF_AR(dim=d, K=9, hidden_dim=8, device=0)
    for rep in range(2):
        ref.load_flat_parameters(theta0)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        hist1, ran1 = ref.fit(x, iters, 0.02, average_window=20, loss_delta_tol=1e-4)
        ev1.record(); torch.cuda.synchronize()
        out["single_ms"] = ev0.elapsed_time(ev1)
    np.savez({out!r}, hist=hist, ran=ran, hist1=hist1, ran1=ran1, same=same, theta=flow.flat_parameters(), theta1=ref.flat_parameters(),
             sharded_ms=out["sharded_ms"], single_ms=out["single_ms"], world=world)
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_training_on_two_gpus_matches_single_gpu(tmp_path):
    """Two ranks, 50 001 rows split 25 001 / 25 000: both ranks end with bitwise identical parameters, and loss curve / parameters
    agree with the one-GPU run up to float32 summation order (the shards' gradients are added in rank order)."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs with peer access")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script, out = tmp_path / "shard_worker.py", str(tmp_path / "res.npz")
    script.write_text(SHARD_WORKER.format(root=root, out=out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", "29749", str(script)], timeout=600)
    r = np.load(out)
    assert bool(r["same"]) and int(r["ran"]) == int(r["ran1"])
    assert np.allclose(r["hist"][:int(r["ran"])], r["hist1"][:int(r["ran"])], rtol=2e-4)
    assert np.max(np.abs(r["theta"] - r["theta1"])) < 5e-3
    print(f"\nrow-sharded training, {int(r['world'])} GPUs: {float(r['sharded_ms']):.2f} ms against {float(r['single_ms']):.2f} ms on one GPU (80 iterations, 50001 x 12)")


# ---------------------------------------------------------------------------------------------------------------------
# (K, hidden) combinations outside the template instantiations: runtime-K / runtime-hidden kernels (nf_generic_kernels.cu)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["d5_K7_H12", "d9_K20_H8", "d6_K3_H5"])
def test_generic_knots_and_hidden_width_match_reference(name):
    """The reference takes any num_knots / hidden_dim (src/flows/flows.py:51); combinations that are not compiled as templates
    run on the generic kernels.  Same fixtures as the compiled combinations, produced by the reference's own flow
    (make_flow_golden.py generic): forward in both layouts, prior log-prob, inverse, conditional inverse at 1e-5 (log-det 2e-5),
    gradient against autograd, Adam trajectory against the reference's optimiser."""
    import os

    c = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"flowG_{name}.npz")))
    d = int(c["d"])
    x = torch.tensor(c["x"])
    f = make_flow(c)
    z, ld = f.forward(x)
    assert _relmax(z.numpy(), c["z_ref"]) <= BAR and _relmax(ld.numpy(), c["ld_ref"]) <= BAR_LOGDET
    g = make_flow(c, reference_layout=False)
    z, ld = g.forward(x)
    assert _relmax(z.numpy(), c["z_col"]) <= BAR and _relmax(ld.numpy(), c["ld_col"]) <= BAR_LOGDET
    lp = g.log_prob(x).numpy()
    want = c["ld_col"] - 0.5 * (c["z_col"].astype(np.float64) ** 2).sum(1) - 0.5 * d * np.log(2 * np.pi)
    assert _relmax(lp, want) <= BAR
    xi, ldi = f.inverse(torch.tensor(c["zin"]))
    assert _relmax(xi.numpy(), c["x_inv"]) <= BAR and _relmax(ldi.numpy(), c["ld_inv"]) <= BAR_LOGDET
    sep = int(c["sep"])
    xc = f.inverse_given_separator(torch.tensor(c["zin_f"]), torch.tensor(c["x_sep"]) if sep else None)
    assert _relmax(xc.numpy(), c["x_cond"]) <= BAR
    loss, grad = g.loss_and_grad(x)
    assert abs(loss - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    assert np.max(np.abs(grad - c["grad"])) <= 2e-4 * np.max(np.abs(c["grad"]))
    steps = len(c["adam_loss"])
    hist, ran = g.fit(x, steps, float(c["adam_lr"]), average_window=0)
    # Against the reference's float32 optimiser the curve is compared tightly over the first steps only: at K = 20 (narrow bins)
    # the reference's own float32 run leaves the float64 trajectory at step 4 (12.9157 against 12.9234 in float64 AND in this
    # kernel), so the whole curve is pinned to the float64 oracle and only loosely to the reference.
    assert ran == steps and np.allclose(hist[:4], c["adam_loss"][:4], rtol=2e-5, atol=1e-5), (hist, c["adam_loss"])
    assert np.allclose(hist, c["adam_loss"], rtol=3e-2)           # same bound as test_adam_trajectory_golden
    _, hist64, _ = orc.train(c["theta"], d, int(c["K"]), int(c["H"]), float(c["B"]), c["x"], steps, float(c["adam_lr"]), average_window=0,
                             dtype=np.float64)
    assert np.allclose(hist, np.asarray(hist64)[:steps], rtol=1e-4), (hist, hist64)
    # the early-stop window works on this path too (same rule as the compiled kernels)
    g.load_flat_parameters(c["theta"])
    hist2, ran2 = g.fit(x, 200, 0.01, average_window=10, loss_delta_tol=0.5)
    assert ran2 == 20 and np.all(hist2[ran2:] == 0.0)
