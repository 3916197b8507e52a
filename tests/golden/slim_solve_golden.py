"""Shrinks a full per-step golden of make_solve_golden.py (every step, 1000 rows: tens of MB for a 100-pose graph) to a fixture
small enough to commit: the listed steps only, mean / std over all rows plus the first `rows` posterior samples of each.

    python tests/golden/slim_solve_golden.py manhattan_r1_p100 24 49 74 99            # seed 0
    python tests/golden/slim_solve_golden.py manhattan_r1_p100:1 24 49 74 99          # seed 1
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(spec, steps, rows=400):
    case, _, seed = spec.partition(":")
    name = f"solve_{case}.npz" if not seed or seed == "0" else f"solve_{case}_seed{seed}.npz"
    path = os.path.join(HERE, name)
    g = np.load(path)
    n_steps = len([k for k in g.files if k.endswith("_order")])
    # per step [graph update, sampling + training of every clique of the step, posterior sampling] (a step that re-trains two
    # cliques has two [sampler, train] pairs: FactorGraphSolver.py:437-468)
    timers = np.array([[g[f"step{i}_timer"][0], float(np.sum(g[f"step{i}_timer"][1:-1])), g[f"step{i}_timer"][-1]] for i in range(n_steps)])
    out = {"truth": g["truth"], "names": g["names"], "kept_steps": np.array(steps), "timers": timers}
    for i in steps:
        x = g[f"step{i}_samples"]
        out[f"step{i}_order"] = g[f"step{i}_order"]
        out[f"step{i}_mean"] = x.mean(0)
        out[f"step{i}_std"] = x.std(0)
        out[f"step{i}_samples"] = x[:rows].astype(np.float32)
        out[f"step{i}_tree"] = g[f"step{i}_tree"]
        out[f"step{i}_clique_dims"] = g[f"step{i}_clique_dims"]
        if f"step{i}_hypo" in g.files:
            out[f"step{i}_hypo"] = g[f"step{i}_hypo"]
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main(sys.argv[1], [int(a) for a in sys.argv[2:]])
