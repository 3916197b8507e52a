#!/usr/bin/env python
"""Turns the JSON lines of bench.py / benchmarks/*.py (gpurun_out/) into profiles/<round>_results.md."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def load(path):
    rows = []
    if os.path.exists(path):
        for line in open(path):
            line = line.strip()
            if line.startswith("{"):
                try:
                    rows.append(json.loads(line))
                except json.JSONDecodeError:
                    pass
    return rows


def main(tag="r1"):
    md = [f"# Results, round {tag[1:]} (B200, measured through gpurun; CUDA-event timings, never under a profiler)", ""]
    b = load(os.path.join(OUT, f"bench_{tag}_final.json"))
    if b:
        j = b[0]
        r = j["roofline"]
        md += ["## bench.py headline (N = 1)", "",
               f"* workload: {j['config']['workload']}",
               f"* **{j['value'] / 1e9:.2f} G samples/s** resident ({j['ms_per_step']:.3f} ms per 1e7-sample pass), "
               f"e2e through the C ABI with pinned host buffers **{j['e2e']['value'] / 1e9:.2f} G samples/s** "
               f"(H2D {j['e2e']['h2d_bytes_per_step'] / 1e6:.0f} MB + D2H {j['e2e']['d2h_bytes_per_step'] / 1e6:.0f} MB per step: PCIe-bound)",
               f"* roofline: {r['achieved']:.1f} TFLOP/s algorithmic of {r['peak']:.1f} TFLOP/s (FFMA probe, constant operands) = **{r['frac']:.3f}**; "
               f"register-operand FFMA / FFMA2 probes: {r['fp32_probe_tflops']['ffma_reg']:.1f} / {r['fp32_probe_tflops']['ffma2']:.1f} TFLOP/s; "
               f"MUFU {r['mufu']['achieved_gops']:.0f} of {r['mufu']['peak_gops']:.0f} Gop/s; HBM {r['hbm']['achieved']:.0f} of {r['hbm']['peak']:.0f} GB/s ({r['hbm']['frac']:.3f})",
               f"* clocks during the timed region: {j['clocks']}",
               f"* cpu_baseline: {j['cpu_baseline']['value'] / 1e6:.2f} M samples/s ({j['cpu_baseline']['sample']})" if "cpu_baseline" in j else "",
               ""]
        inc = j.get("incr_step", {})
        if inc:
            sv = inc.get("solve_small_case1", {})
            md += ["## clique-flow train + sample, seconds per incremental step", "",
                   f"* synthetic clique (n=2000, dim 11, K 9, 2000 Adam iterations, no early stop): {inc['train_2000_iters_s'] * 1e3:.1f} ms "
                   f"({inc['train_us_per_iter']:.2f} us/iteration) + {inc['sample_1000_s'] * 1e3:.2f} ms for 1000 posterior draws",
                   f"* small range graph (reference settings, 6 steps): {[round(x * 1e3, 1) for x in sv.get('s_per_incr_step', [])]} ms per step, "
                   f"mean position error {sv.get('mean_abs_position_error', float('nan')):.2f}; reference stored run: {sv.get('reference_stored_s_per_step')} s per step",
                   f"* one Adam step over 1e6 x 12 samples: {inc['train_step_1e6_samples_per_s'] / 1e6:.0f} M samples/s"]
            ch = inc.get("solve_manhattan_r1_p100")
            if ch:
                md += [f"* 100-pose range-SLAM graph (configs[3], settings of the stored reference runs, {ch['steps']} steps): "
                       f"{ch['s_per_incr_step_mean'] * 1e3:.1f} ms per step (median {ch['s_per_incr_step_median'] * 1e3:.1f}, p90 {ch['s_per_incr_step_p90'] * 1e3:.1f}, "
                       f"first {ch['s_per_incr_step_first'] * 1e3:.0f}), mean pose error {ch['mean_pose_error']:.2f}; reference: "
                       f"{ch['reference_s_per_step']} s per step, mean pose error {ch['reference_mean_pose_error']} (two stored runs)"]
            md += [""]
    for n in (2, 8):
        g = load(os.path.join(OUT, f"bench_{tag}_g{n}.json"))
        if g:
            md += [f"* bench.py at N = {n} (torchrun, replicas, no collective): {g[0]['value'] / 1e9:.2f} G samples/s, e2e {g[0]['e2e']['value'] / 1e9:.2f} G samples/s"]
    md += [""]
    rows = load(os.path.join(OUT, f"micro_{tag}.jsonl"))
    if rows:
        md += ["## Flow micro-benchmark (M1): K = 9, hidden 8, one GPU", "",
               "| op | d | n | ms | M samples/s | TFLOP/s (algorithmic) | frac of FFMA probe |", "|---|---|---|---|---|---|---|"]
        for r in rows:
            if r["bench"] == "flow" and r["op"] in ("forward", "log_prob", "inverse", "cond_inverse", "train_step"):
                md.append(f"| {r['op']} | {r['d']} | {r['n']:.0e} | {r['ms']:.3f} | {r['samples_per_s'] / 1e6:.0f} | {r.get('tflops', 0):.1f} | {r.get('frac_fp32_peak', float('nan')):.3f} |")
        md += ["", "| 200-iteration training run | d | n | ms | us / iteration |", "|---|---|---|---|---|"]
        for r in rows:
            if r["bench"] == "flow" and r["op"] == "train_200_iters":
                md.append(f"| train_200_iters | {r['d']} | {r['n']} | {r['ms']:.2f} | {r['us_per_iter']:.2f} |")
        md += ["", "## Factor micro-benchmark (M2): float64", "",
               "| case | D | n | ms | M evals/s | GB/s (8 (D+1) n / t) | frac of measured HBM |", "|---|---|---|---|---|---|---|"]
        for r in rows:
            if r["bench"] == "factor":
                md.append(f"| {r['case']} | {r['D']} | {r['n']:.0e} | {r['ms']:.3f} | {r['evals_per_s'] / 1e6:.0f} | {r['hbm_gbs']:.0f} | {r['frac_hbm_peak']:.3f} |")
        pp = [r for r in rows if r["bench"] == "posterior_pass"]
        if pp:
            md += ["", "## Posterior down-pass (S2): one nfisam_posterior_pass call, clique dim 9 (6 given + 3 generated columns)", "",
                   "| cliques | shape | rows | ms | us / clique | M row-cliques / s |", "|---|---|---|---|---|---|"]
            for r in pp:
                shape = "chain" if r["branches"] == 0 else f"{r['branches']} subtrees below a trunk of {r['trunk']}"
                md.append(f"| {r['cliques']} | {shape} | {r['rows']} | {r['ms']:.3f} | {r['us_per_clique']:.2f} | {r['rows_x_cliques_per_s'] / 1e6:.1f} |")
        mm = [r for r in rows if r["bench"] == "mmd"]
        if mm:
            md += ["", "## Two-sample statistics (N4): MMDb, float64", "", "| m = n | d | ms | G kernel evaluations / s | FP64 GFLOP/s (3 d + 30 per pair) |",
                   "|---|---|---|---|---|"]
            for r in mm:
                md.append(f"| {r['m']} | {r['d']} | {r['ms']:.3f} | {r['pairs_per_s'] / 1e9:.1f} | {r['fp64_gflops']:.0f} |")
        md += [""]
    solves = sorted(glob.glob(os.path.join(OUT, "solve_*.json")))
    if solves:
        md += ["## Incremental solves (M3): synthetic Manhattan-world range SLAM", "",
               "| graph | training rows per clique | GPUs | steps | s / incr step (mean) | median | split graph / simulate / train / posterior (ms) | cliques per step | max level width | pose mean error | file |",
               "|---|---|---|---|---|---|---|---|---|---|---|"]
        for p in solves:
            for j in load(p):
                sp = [round(1e3 * x, 1) for x in j["split_mean_graph_sim_train_posterior"]]
                md.append(f"| {j['robots']} robot(s) x {j['poses_per_robot']} poses, {j['landmarks']} landmarks | {j['config']['train_samples']} | {j['n_gpus']} | {j['steps']} | "
                          f"{j['s_per_incr_step_mean']:.4f} | {j['s_per_incr_step_median']:.4f} | {sp} | {j['cliques_trained_per_step_mean']:.0f} | {j['max_level_width']} | "
                          f"{j['pose_mean_error']:.2f} | {os.path.basename(p)} |")
        md += [""]
    open(os.path.join(ROOT, "profiles", f"{tag}_results.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main(*(sys.argv[1:2]))
