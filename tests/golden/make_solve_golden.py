"""Runs the REFERENCE solver (/root/reference/src, unmodified, through ref_shim.py) on the small
range-SLAM graphs and stores its per-step posterior samples as golden fixtures
(tests/golden/solve_<case>.npz).  Build container only (several minutes of CPU time):

    PYTHONHASHSEED=0 python tests/golden/make_solve_golden.py

Settings follow example/slam/small_range_gaussian_problem/run_nfisam.py:12-27 (K=9, hidden 8,
2000 training samples, lr .025, tol .01, window 50, 1000 posterior samples) with the iteration cap
lowered to 600 to keep the CPU run short.  RNG: random / numpy / torch seeded with 0."""
import os
import random
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from slam.NFiSAM import NFiSAM, NFiSAMArgs  # noqa: E402
from slam.RunBatch import graph_file_parser, group_nodes_factors_incrementally  # noqa: E402
from factors.Factors import BinaryFactorMixture  # noqa: E402

ITERS = int(os.environ.get("GOLDEN_ITERS", "600"))
# settings of example/slam/manhattan_world_with_range/manhattan_plaza/run_nfisam.py:5-10, 42-46 (lr .01, 500 iterations,
# loss_delta_tol 1e-9 = no early stop) for the reference's own 136-pose ambiguous-association graph
PLAZA = dict(num_knots=9, flow_iterations=500, local_sample_num=2000, learning_rate=.01, hidden_dim=8, cuda_training=False,
             elimination_method="pose_first", data_parallel=False, training_set_frac=1.0, loss_delta_tol=1e-9, average_window=50,
             posterior_sample_num=500)


def run(case, seed=0):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    nodes, truth, factors = graph_file_parser(os.path.join(HERE, "..", "data", case + ".fg"), "fg", 0.1)
    steps = group_nodes_factors_incrementally(nodes=nodes, factors=factors, incremental_step=1)
    if case.startswith("manhattan_plaza"):
        args = NFiSAMArgs(**PLAZA)
    else:
        args = NFiSAMArgs(num_knots=9, flow_iterations=ITERS, local_sample_num=2000, learning_rate=.025, hidden_dim=8,
                          cuda_training=False, elimination_method="pose_first", training_set_frac=1.0, loss_delta_tol=.01,
                          posterior_sample_num=1000)
    solver = NFiSAM(args)
    out = {"truth": np.concatenate([truth[v] for v in nodes]), "names": np.array([v.name for v in nodes])}
    mixtures = []
    for i, (sn, sf) in enumerate(steps):
        for v in sn:
            solver.add_node(v)
        for f in sf:
            solver.add_factor(f)
            if isinstance(f, BinaryFactorMixture):
                mixtures.append(f)
        timer = []
        t0 = time.time()
        solver.update_physical_and_working_graphs(timer=timer)
        cur = solver.incremental_inference(timer=timer)
        order = solver.elimination_ordering
        out[f"step{i}_order"] = np.array([v.name for v in order])
        out[f"step{i}_samples"] = np.hstack([cur[v] for v in order]).astype(np.float32)
        out[f"step{i}_timer"] = np.array(timer)
        out[f"step{i}_clique_dims"] = np.array(sorted(len(v) for v in solver._temp_training_loss.values()))
        out[f"step{i}_tree"] = np.array(sorted("".join(sorted(x.name for x in c.frontal)) + "|" + "".join(sorted(x.name for x in c.separator))
                                               for c in solver.physical_bayes_tree.clique_nodes))
        if mixtures:
            ws = [np.asarray(f.posterior_weights(cur), float) for f in mixtures if set(f.vars).issubset(cur.keys())]
            k = max(len(w) for w in ws)             # 2- and 3-way associations: rows padded with NaN
            out[f"step{i}_hypo"] = np.array([np.concatenate([w, np.full(k - len(w), np.nan)]) for w in ws])
        print(case, "step", i, "%.1f s" % (time.time() - t0), [round(t, 2) for t in timer], flush=True)
    np.savez_compressed(os.path.join(HERE, f"solve_{case}.npz" if seed == 0 else f"solve_{case}_seed{seed}.npz"), **out)


if __name__ == "__main__":
    # `case` or `case:seed`; a second seed of the same case calibrates the reference's own run-to-run spread
    for spec in sys.argv[1:] or ["small_case1", "small_case1_da", "small_case1:1"]:
        case, _, seed = spec.partition(":")
        run(case, int(seed or 0))
