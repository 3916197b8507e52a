#!/usr/bin/env python
"""Flow and factor micro-benchmarks on one GPU; one JSON line per row.

Flow rows: forward (z + logdet), log_prob, inverse, conditional inverse (sep = d/2), one Adam step, for
d in {6, 8, 10, 12}, n in {1e5, 1e6, 1e7}, K = 9, hidden 8, and a 200-iteration training run at n in {2000, 1e5}.
Factor rows: SE2 prior / SE2 between / range / 2- and 3-way mixture / fused 14-factor joint (D = 22) at
n in {2000, 1e5, 1e7} float64 rows, with achieved HBM bandwidth = 8 (D + 1) n / t against MEASURED_PEAKS.json.
Timing: CUDA events, 3 warm-up + best-of-5 (inputs > L2 for n = 1e7; smaller n are L2-resident and say so)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=5, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch

    sys.path.insert(0, ROOT)
    from bench import flops_fwd, make_inputs
    from nfisam_b200 import _lib
    from nfisam_b200.factors import (AmbiguousDataAssociationFactor, JointFactor, SE2Pose, SE2R2RangeGaussianLikelihoodFactor,
                                     SE2RelativeGaussianLikelihoodFactor, UnarySE2ApproximateGaussianPriorFactor, _gpu)
    from nfisam_b200.flows import NSF_AR
    from nfisam_b200.slam import R2Variable, SE2Variable, VariableType
    from nfisam_b200.slam.graph_io import read_factor_graph_from_file

    lib = _lib.load()
    dev = torch.device("cuda", 0)
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    peaks = (ctypes.c_double * 4)()
    _lib.check(lib.nfisam_probe_pipe_peaks(0, peaks))
    fp32_peak = max(peaks[0], peaks[1], peaks[2])
    K, H = 9, 8
    ns = [100_000, 1_000_000] if args.quick else [100_000, 1_000_000, 10_000_000]
    for d in (6, 8, 10, 12):
        torch.manual_seed(0)
        flow = NSF_AR(dim=d, K=K, hidden_dim=H, reference_layout=False)
        h = flow.handle()
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for n in ns:
            x = torch.from_numpy(make_inputs(n, d, 1)).to(dev)
            z = torch.empty_like(x)
            ld = torch.empty(n, device=dev)
            zin = torch.randn(n, d, device=dev)
            sep = d // 2
            xs = x[:, :sep].contiguous()
            zf = zin[:, :d - sep].contiguous()
            of = torch.empty(n, d - sep, device=dev)
            rows = {
                "forward": lambda: _lib.check(lib.nfisam_flow_forward(h, x.data_ptr(), n, d, z.data_ptr(), ld.data_ptr(), 0, None, st)),
                "log_prob": lambda: _lib.check(lib.nfisam_flow_log_prob(h, x.data_ptr(), n, d, ld.data_ptr(), st)),
                "inverse": lambda: _lib.check(lib.nfisam_flow_inverse(h, zin.data_ptr(), None, n, 0, d, z.data_ptr(), ld.data_ptr(), None, st)),
                "cond_inverse": lambda: _lib.check(lib.nfisam_flow_inverse(h, zf.data_ptr(), xs.data_ptr(), n, sep, d - sep, of.data_ptr(), None, None, st)),
            }
            for name, fn in rows.items():
                t = timed(fn)
                fl = flops_fwd(d, H, K) if name != "cond_inverse" else flops_fwd(d, H, K) - flops_fwd(sep, H, K)
                print(json.dumps({"bench": "flow", "op": name, "d": d, "K": K, "H": H, "n": n, "ms": t * 1e3,
                                  "samples_per_s": n / t, "tflops": fl * n / t * 1e-12, "frac_fp32_peak": fl * n / t * 1e-12 / fp32_peak,
                                  "l2_resident": n * d * 4 < 100e6}), flush=True)
            if n <= 1_000_000:
                t = timed(lambda: flow.fit(x, 1, 0.01, average_window=0, reset_optimizer=False, pull=False), reps=3, warm=2)
                print(json.dumps({"bench": "flow", "op": "train_step", "d": d, "K": K, "H": H, "n": n, "ms": t * 1e3,
                                  "samples_per_s": n / t, "tflops": 3 * flops_fwd(d, H, K) * n / t * 1e-12}), flush=True)
        for n in (2000, 100_000):
            x = torch.from_numpy(make_inputs(n, d, 2)).to(dev)
            t = timed(lambda: flow.fit(x, 200, 0.01, average_window=0, pull=False), reps=3, warm=1)
            print(json.dumps({"bench": "flow", "op": "train_200_iters", "d": d, "K": K, "H": H, "n": n, "ms": t * 1e3,
                              "us_per_iter": t / 200 * 1e6}), flush=True)
    # ---------------- factors
    X0, X1 = SE2Variable("X0"), SE2Variable("X1")
    L1, L2, L3 = (R2Variable(s, VariableType.Landmark) for s in ("L1", "L2", "L3"))
    rngf = SE2R2RangeGaussianLikelihoodFactor
    nodes, truth, fs = read_factor_graph_from_file(os.path.join(ROOT, "tests", "data", "small_case1.fg"))
    jf = JointFactor(fs, nodes)
    cases = {
        "se2_prior": ([X0], [UnarySE2ApproximateGaussianPriorFactor(X0, SE2Pose(0, 0, 1.57), np.diag([4e-4, 1.6e-5, 4e-6]))]),
        "se2_between": ([X0, X1], [SE2RelativeGaussianLikelihoodFactor(X0, X1, SE2Pose(30, 0, 0), np.diag([.04, .0016, .0004]))]),
        "range": ([X0, L1], [rngf(X0, L1, 42.43, 2.0)]),
        "ada2": ([X0, L1, L2], [AmbiguousDataAssociationFactor(X0, [L1, L2], np.array([.5, .5]), rngf, 42.43, 2.0)]),
        "ada3": ([X0, L1, L2, L3], [AmbiguousDataAssociationFactor(X0, [L1, L2, L3], np.array([.2, .5, .3]), rngf, 42.43, 2.0)]),
        "joint14_D22": (nodes, fs),
    }
    for name, (vs, facs) in cases.items():
        j = JointFactor(facs, vs)
        groups = j.groups()
        D = sum(v.dim for v in vs)
        for n in ([2000, 100_000] if args.quick else [2000, 100_000, 10_000_000]):
            x = torch.randn(n, D, dtype=torch.float64, device=dev) * 0.3
            if name == "joint14_D22":
                x += torch.tensor(np.concatenate([truth[v] for v in nodes]), device=dev)
            out = torch.empty(n, dtype=torch.float64, device=dev)
            arr, nd = _gpu.pack_descs(groups)
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            t = timed(lambda: _lib.check(lib.nfisam_factor_logpdf(arr, nd, x.data_ptr(), n, D, out.data_ptr(), None, 0, st)))
            gbs = 8 * (D + 1) * n / t * 1e-9
            print(json.dumps({"bench": "factor", "case": name, "D": D, "n": n, "ms": t * 1e3, "evals_per_s": n / t,
                              "hbm_gbs": gbs, "frac_hbm_peak": gbs / hbm, "l2_resident": n * D * 8 < 100e6}), flush=True)

    # ---------------- posterior down-pass (S2): one nfisam_posterior_pass call over a clique tree
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_flow_gpu import _random_clique_tree

    from nfisam_b200.flows import posterior_pass

    for trunk, branches, depth in ((100, 0, 0), (1, 8, 63)):
        items, total, zw = _random_clique_tree(11, trunk, branches, depth)
        for n in (1000, 100_000):
            z = torch.randn((n, zw), device=dev)
            S = torch.zeros((n, total), device=dev)
            t = timed(lambda: posterior_pass(items, z, S))
            print(json.dumps({"bench": "posterior_pass", "cliques": len(items), "trunk": trunk, "branches": branches, "rows": n,
                              "clique_dim": items[-1][0].dim, "ms": t * 1e3, "us_per_clique": t / len(items) * 1e6,
                              "rows_x_cliques_per_s": n * len(items) / t}), flush=True)
    # ---------------- two-sample statistics (N4): float64 Gaussian-kernel sums
    from nfisam_b200.utils import MMDb

    for m, d in ((1000, 22), (10_000, 22), (30_000, 12)):
        xa = torch.randn((m, d), dtype=torch.float64, device=dev)
        xb = torch.randn((m, d), dtype=torch.float64, device=dev) + 0.1
        t = timed(lambda: MMDb(xa, xb, float(np.sqrt(d))))
        print(json.dumps({"bench": "mmd", "m": m, "n": m, "d": d, "ms": t * 1e3, "pairs_per_s": 3.0 * m * m / t,
                          "fp64_gflops": 3.0 * m * m * (3 * d + 30) / t * 1e-9}), flush=True)


if __name__ == "__main__":
    main()
