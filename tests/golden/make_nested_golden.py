"""Stores the reference checkout's own posterior artefacts for the small range graph
(example/slam/small_range_gaussian_problem/journal_paper/case1) as tests/golden/small_case1_nested.npz:

  dyn{i}   nested-sampling ("dynesty") reference posterior of step i, dyn1/step{i}.sample (steps 0-3; 4-5 are absent from the
           checkout, .MISSING_LARGE_BLOBS), translation columns only, variables in the dyn1 ordering, 1000 rows drawn with a fixed seed
  nf{i}    the reference's stored NF-iSAM posterior of the same step, run1/step{i} (1000 rows), same columns

The evaluation protocol is the reference's (mmd_rmse_time_da_plot_grid.py:180-254): translation dims only, MMDb with
sigma = sqrt(number of columns) against the nested-sampling posterior.  Build container only:

    python tests/golden/make_nested_golden.py"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = "/root/reference/example/slam/small_range_gaussian_problem/journal_paper/case1"


def xy_columns(order):
    """translation columns of a sample file whose variables are `order` (poses: 3 columns, landmarks: 2)."""
    cols, off = {}, 0
    for name in order:
        cols[name] = [off, off + 1]
        off += 2 if name.startswith("L") else 3
    return cols, off


def main():
    rng = np.random.default_rng(0)
    out = {}
    for i in range(4):
        dyn_order = open(f"{CASE}/dyn1/step{i}_ordering").read().split()
        dyn = np.loadtxt(f"{CASE}/dyn1/step{i}.sample")
        cols, width = xy_columns(dyn_order)
        assert dyn.shape[1] == width
        keep = np.concatenate([cols[n] for n in dyn_order])
        rows = rng.choice(dyn.shape[0], size=min(1000, dyn.shape[0]), replace=False)
        out[f"dyn{i}"] = dyn[rows][:, keep].astype(np.float32)
        nf_order = open(f"{CASE}/run1/step{i}_ordering").read().split()
        nf = np.loadtxt(f"{CASE}/run1/step{i}")
        ncols, nwidth = xy_columns(nf_order)
        assert nf.shape[1] == nwidth
        out[f"nf{i}"] = nf[:, np.concatenate([ncols[n] for n in dyn_order])].astype(np.float32)
        out[f"order{i}"] = np.array(dyn_order)
    np.savez_compressed(os.path.join(HERE, "small_case1_nested.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
