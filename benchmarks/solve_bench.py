#!/usr/bin/env python
"""Incremental-solve benchmark: seconds per incremental step of the drop-in NFiSAM
solver on synthetic Manhattan-world range-SLAM graphs.

  python benchmarks/solve_bench.py --robots 1 --poses 100                      # configs[3]: 100+ poses, chain tree
  torchrun --nnodes=1 --nproc-per-node 8 benchmarks/solve_bench.py --robots 8 --poses 64   # configs[4]: clique-parallel

One JSON line on rank 0.  Under torchrun (the process group is handed to the solver explicitly: NFiSAMArgs.process_group) cliques of a tree level are dealt
round-robin to the GPUs (nfisam_b200/slam/scheduler.py), NCCL moves trained parameters up and separator samples
down.  Results are independent of the GPU count (per-clique RNG seeding)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_solve(robots=1, poses=100, landmarks=4, ada_prob=0.0, iters=500, samples=2000, posterior=1000, lr=0.02, knots=9,
              max_steps=0, seed=0, process_group=None, device=None, detail=False):
    """One incremental solve of a synthetic graph through the drop-in NFiSAM API; every rank of `process_group` calls this
    (None = single process).  Returns the result dict on every rank (times are the max over ranks)."""
    import hashlib

    import torch
    import torch.distributed as dist

    from nfisam_b200 import _lib
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import group_nodes_factors_incrementally
    from nfisam_b200.slam.synthetic import make_manhattan_range_graph

    world = dist.get_world_size(process_group) if process_group is not None else 1
    nodes, truth, factors = make_manhattan_range_graph(robots=robots, poses=poses, landmarks=landmarks, ada_prob=ada_prob, seed=seed)
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
    if max_steps:
        steps = steps[:max_steps]
    solver = NFiSAM(NFiSAMArgs(num_knots=knots, flow_iterations=iters, local_sample_num=samples, learning_rate=lr, hidden_dim=8,
                               posterior_sample_num=posterior, elimination_method="pose_first", deterministic_cliques=True,
                               seed=seed, process_group=process_group, device=device))
    per_step, splits, widths, trained = [], [], [], []
    launches0 = _lib.launch_count()
    cur = None
    # objects of whatever ran before in this process (bench.py runs several solves back to back) go to the permanent generation:
    # full collections inside the timed steps then only scan this solve's own objects (they showed up as 0.1 s outlier steps)
    import gc
    gc.collect()
    gc.freeze()
    gc_s, gc_step = [0.0, 0.0], []

    def gc_cb(phase, info):
        if phase == "start":
            gc_s[1] = time.perf_counter()
        else:
            gc_s[0] += time.perf_counter() - gc_s[1]

    if detail:
        gc.callbacks.append(gc_cb)
    for sn, sf in steps:
        for v in sn:
            solver.add_node(v)
        for f in sf:
            solver.add_factor(f)
        timer = []
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=process_group)
        gc_s[0] = 0.0
        t0 = time.perf_counter()
        solver.update_physical_and_working_graphs(timer=timer)
        levels = solver.working_bayes_tree.levels()
        widths.append(max(len(lv) for lv in levels))
        cur = solver.incremental_inference(timer=timer)
        torch.cuda.synchronize()
        per_step.append(time.perf_counter() - t0)
        gc_step.append(gc_s[0])
        splits.append(timer)
        trained.append(len(solver._temp_training_loss))
    if detail:
        gc.callbacks.remove(gc_cb)
    gc.unfreeze()
    pose_err = [float(np.linalg.norm(cur[v].mean(0)[:2] - truth[v][:2])) for v in cur if v.type.value == "Pose"]
    lmk_err = [float(np.linalg.norm(cur[v].mean(0)[:2] - truth[v][:2])) for v in cur if v.type.value == "Landmark"]
    order = solver.elimination_ordering
    digest = hashlib.sha1(np.ascontiguousarray(np.hstack([cur[v] for v in order]), dtype=np.float32).tobytes()).digest()[:8]
    same = True
    t = torch.tensor(per_step, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=process_group)
        h = torch.tensor(list(digest), dtype=torch.int64, device="cuda")
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h, group=process_group)
        same = all(bool(torch.equal(hs[0], x)) for x in hs)
    per_step = t.cpu().numpy()
    sp = np.array(splits)
    extra = {"per_step": [float(v) for v in per_step], "splits": [[float(x) for x in row] for row in sp],
             "gc_s": [float(v) for v in gc_step]} if detail else {}
    return {
        **extra,
        "bench": "incremental_solve", "n_gpus": world, "robots": robots, "poses_per_robot": poses,
        "landmarks": landmarks, "steps": len(steps), "variables": len(cur),
        "config": {"K": knots, "hidden": 8, "train_samples": samples, "max_iters": iters, "lr": lr,
                   "posterior_samples": posterior, "ada_prob": ada_prob},
        "s_per_incr_step_mean": float(per_step.mean()), "s_per_incr_step_median": float(np.median(per_step)),
        "s_per_incr_step_last10_mean": float(per_step[-10:].mean()), "total_s": float(per_step.sum()),
        "s_per_incr_step_first": float(per_step[0]), "s_per_incr_step_max": float(per_step.max()),
        "s_per_incr_step_p90": float(np.percentile(per_step, 90)),
        "slowest_steps": [[int(i), round(float(per_step[i]), 4)] for i in np.argsort(per_step)[::-1][:5]],
        "split_mean_graph_sim_train_posterior": [float(x) for x in sp.mean(0)],
        "cliques_trained_per_step_mean": float(np.mean(trained)), "max_level_width": int(max(widths)),
        "pose_mean_error": float(np.mean(pose_err)), "pose_max_error": float(np.max(pose_err)),
        "landmark_mean_error": float(np.mean(lmk_err)) if lmk_err else None,
        "landmark_errors": [round(e, 3) for e in lmk_err],
        "posterior_sha1_8": digest.hex(), "posterior_identical_on_all_ranks": bool(same),
        "gpu_launches": int(_lib.launch_count() - launches0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", type=int, default=1)
    ap.add_argument("--poses", type=int, default=100)
    ap.add_argument("--landmarks", type=int, default=4)
    ap.add_argument("--ada-prob", type=float, default=0.0)
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--posterior", type=int, default=1000)
    ap.add_argument("--lr", type=float, default=0.02)
    ap.add_argument("--knots", type=int, default=9)
    ap.add_argument("--max-steps", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--reps", type=int, default=1, help="repeat the solve; the last repetition is reported (the first warms up)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    res = None
    for _ in range(max(args.reps, 1)):
        res = run_solve(robots=args.robots, poses=args.poses, landmarks=args.landmarks, ada_prob=args.ada_prob, iters=args.iters,
                        samples=args.samples, posterior=args.posterior, lr=args.lr, knots=args.knots, max_steps=args.max_steps,
                        seed=args.seed, process_group=group)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
