#!/usr/bin/env python
"""Incremental-solve benchmark: seconds per incremental step of the drop-in NFiSAM
solver on synthetic Manhattan-world range-SLAM graphs.

  python benchmarks/solve_bench.py --robots 1 --poses 100                      # configs[3]: 100+ poses, chain tree
  torchrun --nnodes=1 --nproc-per-node 8 benchmarks/solve_bench.py --robots 8 --poses 64   # configs[4]: clique-parallel

One JSON line on rank 0.  Under torchrun every rank runs the same host logic; cliques of a tree level are dealt
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
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from nfisam_b200 import _lib
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import group_nodes_factors_incrementally
    from nfisam_b200.slam.synthetic import make_manhattan_range_graph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nodes, truth, factors = make_manhattan_range_graph(robots=args.robots, poses=args.poses, landmarks=args.landmarks,
                                                       ada_prob=args.ada_prob, seed=args.seed)
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
    if args.max_steps:
        steps = steps[:args.max_steps]
    solver = NFiSAM(NFiSAMArgs(num_knots=args.knots, flow_iterations=args.iters, local_sample_num=args.samples,
                               learning_rate=args.lr, hidden_dim=8, posterior_sample_num=args.posterior,
                               elimination_method="pose_first", deterministic_cliques=True, seed=args.seed))
    per_step, splits, widths, trained = [], [], [], []
    launches0 = _lib.launch_count()
    for sn, sf in steps:
        for v in sn:
            solver.add_node(v)
        for f in sf:
            solver.add_factor(f)
        timer = []
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        solver.update_physical_and_working_graphs(timer=timer)
        levels = solver.working_bayes_tree.levels()
        widths.append(max(len(l) for l in levels))
        before = len(solver._clique_density_model)
        cur = solver.incremental_inference(timer=timer)
        torch.cuda.synchronize()
        per_step.append(time.perf_counter() - t0)
        splits.append(timer)
        trained.append(len(solver._temp_training_loss))
    pose_err = [float(np.linalg.norm(cur[v].mean(0)[:2] - truth[v][:2])) for v in cur if v.type.value == "Pose"]
    lmk_err = [float(np.linalg.norm(cur[v].mean(0)[:2] - truth[v][:2])) for v in cur if v.type.value == "Landmark"]
    t = torch.tensor(per_step, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_step = t.cpu().numpy()
    if rank == 0:
        sp = np.array(splits)
        print(json.dumps({
            "bench": "incremental_solve", "n_gpus": world, "robots": args.robots, "poses_per_robot": args.poses,
            "landmarks": args.landmarks, "steps": len(steps), "variables": len(cur),
            "config": {"K": args.knots, "hidden": 8, "train_samples": args.samples, "max_iters": args.iters, "lr": args.lr,
                       "posterior_samples": args.posterior, "ada_prob": args.ada_prob},
            "s_per_incr_step_mean": float(per_step.mean()), "s_per_incr_step_median": float(np.median(per_step)),
            "s_per_incr_step_last10_mean": float(per_step[-10:].mean()), "total_s": float(per_step.sum()),
            "s_per_incr_step_first": float(per_step[0]), "s_per_incr_step_max": float(per_step.max()),
            "s_per_incr_step_p90": float(np.percentile(per_step, 90)),
            "slowest_steps": [[int(i), round(float(per_step[i]), 4)] for i in np.argsort(per_step)[::-1][:5]],
            "split_mean_graph_sim_train_posterior": [float(x) for x in sp.mean(0)],
            "cliques_trained_per_step_mean": float(np.mean(trained)), "max_level_width": int(max(widths)),
            "pose_mean_error": float(np.mean(pose_err)), "pose_max_error": float(np.max(pose_err)),
            "landmark_mean_error": float(np.mean(lmk_err)) if lmk_err else None,
            "gpu_launches": int(_lib.launch_count() - launches0)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
