"""GPU tests of the solver-facing layer: wrapper parity with the reference's outputs (same weights, same
torch seed => same draws), and full incremental solves compared with posterior samples of the REFERENCE
solver (tests/golden/solve_*.npz, tests/golden/make_solve_golden.py).

Posterior tolerance (Monte-Carlo + model-fitting variance: two runs of the REFERENCE with different seeds
differ by up to 0.52 sigma on pose means, 1.3 sigma on landmark means and 1.8x on stds,
tests/golden/solve_small_case1_seed1.npz): per-variable mean within 0.75 sigma_ref + 0.5 (poses) /
1.5 sigma_ref + 0.5 (landmarks) of the reference's mean, std ratio within [0.33, 3] for poses and [1/6, 6] for
range-only landmarks; the biased MMD (RBF kernel, sigma = sqrt(dim), the reference's post-processing metric) between our
samples and the reference's is reported and bounded by 0.45.  The landmark-std and MMD bounds were calibrated on 16 seeds
of this solver with the host and with the device simulator (profiles/r1_posterior_calibration.md): on the last step of the
ambiguous-association graph a small mirror mode of a landmark survives in some runs, the landmark y std ranges over
1.8-13.7 for BOTH pipelines (the reference's two stored runs: 2.0-4.9) and the joint MMD_b over 0.18-0.45."""
import os

import numpy as np
import pytest
import torch

from tests.test_model_cpu import check_wrapper

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def mmd_b(x, y, sigma):
    """Biased MMD estimate with an RBF kernel (reference: src/utils/Statistics.py:68-84), on the device (nfisam_mmd)."""
    from nfisam_b200.utils import MMDb

    v = MMDb(x, y, sigma)
    return 0.0 if np.isnan(v) else float(v)


def test_wrapper_matches_reference_outputs():
    g = dict(np.load(os.path.join(HERE, "golden", "model.npz")))
    check_wrapper(g, rtol=1e-5)


def solve(case, _reseed=True, **kw):
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import graph_file_parser, group_nodes_factors_incrementally

    nodes, truth, factors = graph_file_parser(os.path.join(HERE, "data", case + ".fg"))
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
    if _reseed:
        np.random.seed(0)
        torch.manual_seed(0)
    args = dict(num_knots=9, flow_iterations=600, local_sample_num=2000, learning_rate=.025, hidden_dim=8,
                elimination_method="pose_first", loss_delta_tol=.01, posterior_sample_num=1000)
    args.update(kw)
    solver = NFiSAM(NFiSAMArgs(**args))
    per_step = []
    for sn, sf in steps:
        for v in sn:
            solver.add_node(v)
        for f in sf:
            solver.add_factor(f)
        timer = []
        solver.update_physical_and_working_graphs(timer=timer)
        cur = solver.incremental_inference(timer=timer)
        order = solver.elimination_ordering
        per_step.append(([v.name for v in order], np.hstack([cur[v] for v in order]), timer, solver))
    return per_step


# Runs of this solver per statistical comparison.  A solve of the small graphs takes well under a second on the GPU, and the
# median of 15 seeds is stable where the median of 5 (round 1) flipped with any change of summation order in the kernels.
N_SEEDS = 15


def solve_seeded(case, seed, **kw):
    np.random.seed(seed)
    torch.manual_seed(seed)
    return solve(case, _reseed=False, **kw)


@pytest.mark.parametrize("case", ["small_case1", "small_case1_da", "manhattan_r1_p10", "manhattan_r2_p5", "manhattan_r2_p8_ada"])
def test_incremental_solve_matches_reference_posterior(case):
    """N_SEEDS independently seeded runs of this solver against the reference's stored posterior(s).  NF-iSAM's
    run-to-run spread is large (the reference itself, seeds 0 vs 1: pose means up to 0.52 sigma apart, landmark
    means up to 1.3 sigma, stds up to 1.8x, joint MMD_b up to 0.30), so every statistic is the MEDIAN over our
    runs of the distance to the CLOSEST reference run."""
    path = os.path.join(HERE, "golden", f"solve_{case}.npz")
    if not os.path.exists(path):
        pytest.skip("golden posterior not generated")
    refs = [dict(np.load(path))]
    alt = os.path.join(HERE, "golden", f"solve_{case}_seed1.npz")
    if os.path.exists(alt):
        refs.append(dict(np.load(alt)))
    runs = [solve_seeded(case, seed) for seed in range(N_SEEDS)]
    n_steps = len(runs[0])
    report = []
    for i in range(n_steps):
        names = runs[0][i][0]
        assert names == list(refs[0][f"step{i}_order"])
        if i == n_steps - 1:
            solver = runs[0][i][3]
            tree = sorted("".join(sorted(v.name for v in c.frontal)) + "|" + "".join(sorted(v.name for v in c.separator))
                          for c in solver.physical_bayes_tree.clique_nodes)
            assert tree == list(refs[0][f"step{i}_tree"]), (tree, list(refs[0][f"step{i}_tree"]))
        mean_excess, std_bad, mmds = [], [], []
        for run in runs:
            x = run[i][1]
            best_excess, best_std, best_mmd = np.inf, np.inf, np.inf
            for g in refs:
                ref = g[f"step{i}_samples"]
                assert x.shape == ref.shape
                m, mr, s_, sr = x.mean(0), ref.mean(0), x.std(0), ref.std(0)
                col, excess, stdr, stdr_lm = 0, 0.0, 1.0, 1.0
                for nm in names:
                    w = 2 if nm.startswith("L") else 3
                    sl = slice(col, col + w)
                    col += w
                    # range-only landmark marginals are multi-modal until enough poses have seen them: means / stds
                    # are compared only where the reference's marginal is concentrated; the rest is covered by the MMD
                    if np.all(sr[sl] < 5.0):
                        tol = (1.5 if nm.startswith("L") else 0.75) * sr[sl] + 0.5
                        excess = max(excess, float(np.max(np.abs(m[sl] - mr[sl]) / tol)))
                        r = s_[sl] / np.maximum(sr[sl], 1e-9)
                        worst = float(np.max(np.maximum(r, 1.0 / np.maximum(r, 1e-9))))
                        if nm.startswith("L"):
                            stdr_lm = max(stdr_lm, worst)
                        else:
                            stdr = max(stdr, worst)
                # landmarks are held to 6x, poses to 3x: fold both into one number with the bound 3
                best_excess, best_std = min(best_excess, excess), min(best_std, max(stdr, stdr_lm / 2.0))
                best_mmd = min(best_mmd, mmd_b(x[:500].astype(np.float64), ref[:500].astype(np.float64), np.sqrt(x.shape[1])))
            mean_excess.append(best_excess)
            std_bad.append(best_std)
            mmds.append(best_mmd)
        assert np.median(mean_excess) <= 1.0, (case, i, mean_excess)
        assert np.median(std_bad) <= 3.0, (case, i, std_bad)
        report.append(float(np.median(mmds)))
    print(f"\n[{case}] joint MMD_b vs reference per step (median over the seeds):", np.round(report, 4))
    assert max(report) < 0.45, report
    g = refs[0]
    key = f"step{n_steps - 1}_hypo"
    if key in g and len(g[key]):
        from nfisam_b200.factors import BinaryFactorMixture

        ws = []
        for solver in [r[-1][3] for r in runs]:
            mix = [f for f in solver.physical_factors if isinstance(f, BinaryFactorMixture)]
            ws.append(np.array([f.posterior_weights(solver._samples) for f in mix]))
        w = np.median(np.array(ws), axis=0)
        print("hypothesis weights:", np.round(w, 3).tolist(), "reference:", np.round(g[key], 3).tolist())
        assert np.all(np.abs(w - g[key]) < 0.2)


def test_small_graph_posterior_vs_nested_sampling():
    """The reference checkout ships an independent ground truth for its small range graph: nested-sampling ("dynesty") posteriors
    of steps 0-3 (journal_paper/case1/dyn1/step{i}.sample) next to its own stored NF-iSAM run (run1/step{i}); fixture
    tests/golden/small_case1_nested.npz (make_nested_golden.py).  With the reference's settings for this case
    (run_nfisam.py:12-27: K 9, hidden 8, 2000 training samples, <= 2000 iterations, lr .025, tol .01, 1000 posterior samples) and
    its evaluation protocol (mmd_rmse_time_da_plot_grid.py:180-254: translation columns only, MMDb with sigma = sqrt(#columns),
    500 rows) this solver must be as close to the nested-sampling posterior as the reference's own runs are.  Three runs of
    the reference exist for this graph: the stored journal run (run1) and the two runs of tests/golden/solve_small_case1*.npz
    (make_solve_golden.py, iteration cap 600).  Their MMDb against the nested-sampling posterior is 0.063 / 0.064 / 0.063-0.072
    at steps 0-2 (the noise floor of 500 against 500 samples) and 0.146 / 0.259 / 0.274 at step 3.  Bound, at every step:
        median over seeds of MMDb(ours, nested)  <=  median over the reference's runs of MMDb(reference, nested) + 0.03."""
    g = dict(np.load(os.path.join(HERE, "golden", "small_case1_nested.npz")))
    extra = [dict(np.load(os.path.join(HERE, "golden", f))) for f in ("solve_small_case1.npz", "solve_small_case1_seed1.npz")]
    runs = [solve_seeded("small_case1", seed, flow_iterations=2000) for seed in range(N_SEEDS)]

    def xy_of(names, x, order):
        cols, off = {}, 0
        for nm in names:
            cols[nm] = [off, off + 1]
            off += 2 if nm.startswith("L") else 3
        return x[:500][:, np.concatenate([cols[nm] for nm in order])].astype(np.float64)

    ours, theirs = [], []
    for i in range(4):
        order = [str(n) for n in g[f"order{i}"]]
        dyn = g[f"dyn{i}"][:500].astype(np.float64)
        sigma = np.sqrt(dyn.shape[1])
        ref_vals = [mmd_b(g[f"nf{i}"][:500].astype(np.float64), dyn, sigma)]
        ref_vals += [mmd_b(xy_of([str(n) for n in e[f"step{i}_order"]], e[f"step{i}_samples"], order), dyn, sigma) for e in extra]
        theirs.append(ref_vals)
        ours.append([mmd_b(xy_of(run[i][0], run[i][1], order), dyn, sigma) for run in runs])
    med = [float(np.median(v)) for v in ours]
    ref_med = [float(np.median(v)) for v in theirs]
    print("\nMMDb against the nested-sampling posterior, steps 0-3: ours (median of %d seeds)" % N_SEEDS, np.round(med, 4),
          "min", np.round([min(v) for v in ours], 4), "max", np.round([max(v) for v in ours], 4),
          "| reference runs (journal run1, 600-iteration seeds 0 / 1)", [np.round(v, 4).tolist() for v in theirs])
    for i in range(4):
        assert med[i] <= ref_med[i] + 0.03, (i, med, theirs)


def _step_distance(m, s_, xs, names, g, i, rows=400):
    """(mean excess, std ratio, MMD_b) of a posterior with mean m, std s_ and sample rows xs at step i against one stored
    reference run g: the statistics of test_incremental_solve_matches_reference_posterior (means / stds only where the
    reference's marginal is concentrated)."""
    ref = g[f"step{i}_samples"]
    mr = g[f"step{i}_mean"] if f"step{i}_mean" in g else ref.mean(0)
    sr = g[f"step{i}_std"] if f"step{i}_std" in g else ref.std(0)
    col, excess, stdr, stdr_lm = 0, 0.0, 1.0, 1.0
    for nm in names:
        w = 2 if nm.startswith("L") else 3
        sl = slice(col, col + w)
        col += w
        if np.all(sr[sl] < 5.0):
            tol = (1.5 if nm.startswith("L") else 0.75) * sr[sl] + 0.5
            excess = max(excess, float(np.max(np.abs(m[sl] - mr[sl]) / tol)))
            r = s_[sl] / np.maximum(sr[sl], 1e-9)
            worst = float(np.max(np.maximum(r, 1.0 / np.maximum(r, 1e-9))))
            if nm.startswith("L"):
                stdr_lm = max(stdr_lm, worst)
            else:
                stdr = max(stdr, worst)
    k = min(rows, len(ref), len(xs))
    return excess, max(stdr, stdr_lm / 2.0), mmd_b(xs[:k].astype(np.float64), ref[:k].astype(np.float64), np.sqrt(xs.shape[1]))


def _pose_error(mean, order, names_all, truth):
    """Mean translation error of the pose means against the ground truth."""
    off, c = {}, 0
    for n in names_all:
        off[n] = c
        c += 2 if n.startswith("L") else 3
    errs, col = [], 0
    for nm in order:
        if not nm.startswith("L"):
            errs.append(np.linalg.norm(mean[col:col + 2] - truth[off[nm]:off[nm] + 2]))
        col += 2 if nm.startswith("L") else 3
    return float(np.mean(errs))


# reference settings of example/slam/manhattan_world_with_range/manhattan_plaza/run_nfisam.py:5-10, 42-46
PLAZA_ARGS = dict(num_knots=9, flow_iterations=500, local_sample_num=2000, learning_rate=.01, hidden_dim=8, loss_delta_tol=1e-9,
                  average_window=50, posterior_sample_num=500)


@pytest.mark.parametrize("case,n_steps,kwargs", [("manhattan_r1_p100", 100, dict(flow_iterations=500)),
                                                 ("manhattan_plaza_ada", 136, PLAZA_ARGS)])
def test_100_pose_solve_matches_reference_posterior(case, n_steps, kwargs):
    """manhattan_plaza_ada: the REFERENCE'S OWN large graph, example/slam/manhattan_world_with_range/manhattan_plaza/res/seed0/
    pada0.4_r2_odom0.01_mada3/factor_graph.fg (136 poses, 4 landmarks, 59 ambiguous 2- / 3-way data associations), with the
    settings of its run_nfisam.py, two stored runs of the reference (~65 min of CPU each), compared after 34 / 68 / 102 / 136 steps.
    manhattan_r1_p100: BASELINE configs[3] stand-in of round 1: a 100-pose Manhattan-world range-SLAM graph (4 landmarks, 302 factors, 100 incremental steps)
    solved by the reference (tests/golden/make_solve_golden.py, 500 iterations per clique, ~25 min of CPU per run) and by
    this solver, compared after 25, 50, 75 and 100 steps.

    What "matches" can mean here is set by the reference itself: on this graph its own posterior is far from the ground truth
    and over-confident (stored run, seed 0: mean pose error 3.1 / 5.8 / 7.3 / 8.9 after 25 / 50 / 75 / 100 steps with pose stds
    below 1; landmark L3 collapses onto a mirror mode 186 units from the truth with std 0.09), and two of its runs differ
    from each other by more than the bounds used for the small graphs.  So the test (a) bounds this solver's error against the
    ground truth by the reference's own, and (b) when a second stored reference run exists, bounds the distance of our posterior
    to the closest reference run by the distance between the two reference runs (2x + 1 for the mean excess and the std ratio,
    1.25x + 0.1 for MMD_b, which saturates near 1); with a single stored run the distances are only reported.
    Stored reference runs, seed 0 vs seed 1 (tests/golden/slim_solve_golden.py): mean pose error 3.1 / 4.2, 5.8 / 6.3, 7.3 / 8.9,
    8.9 / 12.6 after 25 / 50 / 75 / 100 steps; between the two runs MMD_b 0.70, 0.87, 0.92, 1.05 and mean excess 2.9, 196, 236,
    228 (a landmark collapsed onto different modes)."""
    path = os.path.join(HERE, "golden", f"solve_{case}.npz")
    if not os.path.exists(path):
        pytest.skip("golden posterior not generated")
    refs = [dict(np.load(path))]
    alt = os.path.join(HERE, "golden", f"solve_{case}_seed1.npz")
    if os.path.exists(alt):
        refs.append(dict(np.load(alt)))
    kept = [int(i) for i in refs[0]["kept_steps"]]
    truth = refs[0]["truth"]
    names_all = [str(n) for n in refs[0]["names"]]
    runs = [solve_seeded(case, seed, **kwargs) for seed in (0, 1, 2, 3, 4)]
    assert len(runs[0]) == n_steps
    report = []
    for i in kept:
        names = runs[0][i][0]
        assert names == list(refs[0][f"step{i}_order"])
        if i == kept[-1]:
            solver = runs[0][i][3]
            tree = sorted("".join(sorted(v.name for v in c.frontal)) + "|" + "".join(sorted(v.name for v in c.separator))
                          for c in solver.physical_bayes_tree.clique_nodes)
            assert tree == list(refs[0][f"step{i}_tree"])
        ours_err = float(np.median([_pose_error(run[i][1].mean(0), names, names_all, truth) for run in runs]))
        ref_err = float(np.mean([_pose_error(g[f"step{i}_mean"], names, names_all, truth) for g in refs]))
        assert ours_err <= 1.5 * ref_err + 1.0, (i, ours_err, ref_err)
        stats = np.array([[min(_step_distance(run[i][1].mean(0), run[i][1].std(0), run[i][1], names, g, i)[j] for g in refs)
                           for j in range(3)] for run in runs])
        med = np.median(stats, axis=0)
        row = {"step": i + 1, "pose_err_ours": round(ours_err, 2), "pose_err_reference": round(ref_err, 2),
               "to_closest_reference(mean_excess,std_ratio,mmd_b)": [round(float(v), 3) for v in med]}
        if len(refs) > 1:
            a, b = refs
            spread = np.maximum(_step_distance(b[f"step{i}_mean"], b[f"step{i}_std"], b[f"step{i}_samples"], names, a, i),
                                _step_distance(a[f"step{i}_mean"], a[f"step{i}_std"], a[f"step{i}_samples"], names, b, i))
            row["reference_seed0_vs_seed1"] = [round(float(v), 3) for v in spread]
            assert med[0] <= 2.0 * spread[0] + 1.0, row
            assert med[1] <= 2.0 * spread[1] + 1.0, row
            assert med[2] <= 1.25 * spread[2] + 0.1, row
        report.append(row)
    print(f"\n[{case}]")
    for row in report:
        print("  ", row)
    key = f"step{kept[-1]}_hypo"
    if key in refs[0]:
        # data-association probabilities of every ambiguous factor after the last step (FactorGraphSolver.py:913-922), all
        # mixtures in one launch.  Two runs of the reference agree on the most likely association of 95 % of the 59 factors
        # (mean absolute difference of the weights 0.011); every run of this solver must agree with the closest reference run
        # at least as well as that, minus 10 points / plus 0.03.
        from nfisam_b200.factors import BinaryFactorMixture, posterior_weights_batch

        ref_w = [np.nan_to_num(g[key]) for g in refs]
        agree, diff = [], []
        for run in runs:
            solver = run[kept[-1]][3]
            mix = [f for f in solver.physical_factors if isinstance(f, BinaryFactorMixture)]
            w = np.zeros_like(ref_w[0])
            for k, wk in enumerate(posterior_weights_batch(mix, solver._samples)):
                w[k, :len(wk)] = wk
            agree.append(max(float(np.mean(w.argmax(1) == r.argmax(1))) for r in ref_w))
            diff.append(min(float(np.abs(w - r).mean()) for r in ref_w))
        ref_agree = float(np.mean(ref_w[0].argmax(1) == ref_w[-1].argmax(1))) if len(ref_w) > 1 else 1.0
        ref_diff = float(np.abs(ref_w[0] - ref_w[-1]).mean()) if len(ref_w) > 1 else 0.0
        print(f"   association weights of {len(ref_w[0])} mixtures: argmax agreement with the closest reference run", np.round(agree, 3),
              "mean |dw|", np.round(diff, 4), "| reference seed 0 vs 1:", round(ref_agree, 3), round(ref_diff, 4))
        assert np.median(agree) >= ref_agree - 0.10 and np.median(diff) <= ref_diff + 0.03


def test_clique_parallel_equals_serial_loop_statistically():
    """Level-synchronous schedule on streams (device simulator) vs the serial reference-order loop (host simulator): same
    posterior up to the run-to-run spread (median of the per-variable means over three seeds each; the reference's own
    seeds differ by up to 0.52 sigma on pose means and 1.3 sigma on landmark means)."""
    a = np.median([solve_seeded("small_case1", s, flow_iterations=300, clique_parallel=True)[-1][1].mean(0) for s in (0, 1, 2)], axis=0)
    runs_b = [solve_seeded("small_case1", s, flow_iterations=300, clique_parallel=False)[-1][1] for s in (0, 1, 2)]
    b = np.median([x.mean(0) for x in runs_b], axis=0)
    sd = np.median([x.std(0) for x in runs_b], axis=0)
    assert np.all(np.abs(a - b) < 0.75 * sd + 0.5), (a, b, sd)


def test_multi_robot_graph_runs_cliques_concurrently():
    from nfisam_b200 import _lib
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import group_nodes_factors_incrementally
    from nfisam_b200.slam.synthetic import make_manhattan_range_graph

    nodes, truth, factors = make_manhattan_range_graph(robots=4, poses=4, landmarks=3, seed=2)
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
    np.random.seed(1)
    torch.manual_seed(1)
    solver = NFiSAM(NFiSAMArgs(num_knots=9, flow_iterations=200, local_sample_num=1000, learning_rate=.02,
                               posterior_sample_num=500, deterministic_cliques=True))
    widths = []
    for sn, sf in steps:
        for v in sn:
            solver.add_node(v)
        for f in sf:
            solver.add_factor(f)
        solver.update_physical_and_working_graphs()
        widths.append(max(len(l) for l in solver.working_bayes_tree.levels()))
        cur = solver.incremental_inference()
    assert max(widths) >= 4
    for var, val in truth.items():
        if var.type.value == "Pose":
            assert np.linalg.norm(cur[var].mean(0)[:2] - val[:2]) < 5.0, (var.name, cur[var].mean(0), val)
    assert _lib.launch_count() > 0


def test_device_resident_posterior_equals_host_loop():
    """The device-resident down-pass (one upload of all latent draws, separator samples stay on the GPU) returns
    bit-identical samples to the reference-shaped host loop under the same torch seed."""
    from nfisam_b200.slam.solver import FactorGraphSolver

    solver = solve("small_case1_da", flow_iterations=100, device_latents=False)[-1][3]
    torch.manual_seed(5)
    a = FactorGraphSolver.sample_posterior(solver)
    torch.manual_seed(5)
    b = solver._scheduler.sample_posterior_device()
    assert set(a.keys()) == set(b.keys())
    for v in a:
        assert np.array_equal(a[v], b[v]), v.name


def test_r2_toy_example_solves_incrementally():
    """example/slam/toy_examples/R2RangeGaussian_example/five_node_range_gaussian_incremental.py through the drop-in API: R2
    variables, GaussianPriorFactor, R2RangeGaussianLikelihoodFactor and R2RelativeGaussianLikelihoodFactor (the factor class
    round 1 lacked), three incremental steps.  Posterior means land on the example's geometry: l1 = (5, 5), x0 at range
    5 sqrt(2) of l1, x1 = x0 + (5, -5), x2 = x1 + (5, 5), l2 = (10, 5)."""
    import torch

    from nfisam_b200.factors import GaussianPriorFactor, R2RangeGaussianLikelihoodFactor, R2RelativeGaussianLikelihoodFactor
    from nfisam_b200.slam import R2Variable
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs

    x0, x1, x2, l1, l2 = (R2Variable(n) for n in ("x0", "x1", "x2", "l1", "l2"))
    sigma = 0.5
    prior_l1 = GaussianPriorFactor(var=l1, mean=np.array([5.0, 5.0]), covariance=np.identity(2) * 0.5)
    prior_l2 = GaussianPriorFactor(var=l2, mean=np.array([10.0, 5.0]), covariance=np.identity(2) * 0.5)
    f_x0_l1 = R2RangeGaussianLikelihoodFactor(var1=x0, var2=l1, observation=5 * np.sqrt(2), sigma=sigma)
    f_l1_x1 = R2RangeGaussianLikelihoodFactor(var1=l1, var2=x1, observation=10, sigma=sigma)
    f_x0_x1 = R2RelativeGaussianLikelihoodFactor(var1=x0, var2=x1, observation=np.array([5, -5]), precision=np.array([[10, 0.0], [0.0, 10]]))
    f_x1_x2 = R2RelativeGaussianLikelihoodFactor(var1=x1, var2=x2, observation=np.array([5, 5]), precision=np.array([[10, 0.0], [0.0, 10]]))
    f_l2_x2 = R2RangeGaussianLikelihoodFactor(var1=l2, var2=x2, observation=5, sigma=sigma)
    np.random.seed(0)
    torch.manual_seed(0)
    model = NFiSAM(NFiSAMArgs(posterior_sample_num=500, flow_number=1, flow_type="NSF_AR", flow_iterations=800, local_sample_num=1000,
                              cuda_training=True, store_clique_samples=False, num_knots=15))
    for nodes, factors in (([l1, x0], [prior_l1, f_x0_l1]), ([x1], [f_x0_x1, f_l1_x1]), ([x2, l2], [prior_l2, f_x1_x2, f_l2_x2])):
        for v in nodes:
            model.add_node(v)
        for f in factors:
            model.add_factor(f)
        model.update_physical_and_working_graphs()
        samples = model.incremental_inference()
    assert set(samples) == {x0, x1, x2, l1, l2} and all(s.shape == (500, 2) for s in samples.values())
    m = {v.name: samples[v].mean(0) for v in samples}
    assert np.linalg.norm(m["l1"] - [5.0, 5.0]) < 1.0 and np.linalg.norm(m["l2"] - [10.0, 5.0]) < 1.0
    d = samples[x1] - samples[x0]
    assert np.linalg.norm(d.mean(0) - [5.0, -5.0]) < 0.5 and np.all(d.std(0) < 1.0)          # the displacement factor binds x0 -> x1
    d = samples[x2] - samples[x1]
    assert np.linalg.norm(d.mean(0) - [5.0, 5.0]) < 0.5
    r = np.linalg.norm(samples[x0] - samples[l1], axis=1)
    assert abs(np.median(r) - 5 * np.sqrt(2)) < 1.0
    r = np.linalg.norm(samples[x2] - samples[l2], axis=1)
    assert abs(np.median(r) - 5.0) < 1.0
