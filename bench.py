#!/usr/bin/env python
"""bench.py -- headline benchmark of the NF-iSAM clique-flow hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on the host cores)

Metric (BASELINE.json): "RQS flow log-prob samples/sec" on the synthetic microbench
(configs[2]: dim 12, K 9, hidden 8, 1e7 samples per GPU); the companion "clique-flow train+sample
s/incr-step" is reported in the same JSON line under "incr_step".

A step = one log-prob pass over one batch of n samples.  `value` is measured with the batch
resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` goes through the
C ABI's host-buffer call with pinned host memory, H2D/D2H inside the timed region.
The batch (n*d*4 B = 480 MB) is larger than the 126 MB L2, so no explicit L2 flush is needed.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D, K_BINS, HID, TAIL = 12, 9, 8, 5.0
N_PER_GPU = 10_000_000


def flops_fwd(d, H, K):
    """algorithmic FLOPs of one forward / log-prob / inverse per sample."""
    macs = H * d * (d - 1) // 2 + (d - 1) * (H * H + H * (3 * K - 1))
    return 2 * macs + (d - 1) * (2 * H + 3 * K - 1) + d * (15 * K + 45)


def flops_fwd_executed(d, H, K):
    """FLOPs the kernel actually issues: only the two derivative columns of the bin that holds the input are
    evaluated (nf_common.cuh, lazy derivative rows); the 2K width / height columns are evaluated in pairs (packed FFMA2),
    the half float4 past them is skipped."""
    ppw = (2 * K + 1) // 2 * 2
    macs = H * d * (d - 1) // 2 + (d - 1) * (H * H + H * (ppw + 2))
    return 2 * macs + (d - 1) * (2 * H + ppw + 2) + d * (13 * K + 55)


def sfu_fwd(d, H, K):
    return 2 * H * (d - 1) + d * (4 * K + 4)


def make_inputs(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    mask = rng.random((n, d), dtype=np.float32) < 0.005   # 0.5 % of entries in the |x| > B tails
    x[mask] *= 8.0
    return x


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from a thread
    every ~2 ms (the timed region can be tens of milliseconds, too short for `nvidia-smi -lms`);
    falls back to the nvidia-smi recipe of B200_PROFILING.md when pynvml is unavailable."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.mode = None

    def _poll(self, nv, handle):
        bits = {}
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            v = getattr(nv, attr, None)
            if v is None:
                v = getattr(nv, attr.replace("ClocksEventReason", "ClocksThrottleReason"), None)
            if v is not None:
                bits[name] = v
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    r = int(get_reasons(handle))
                    for name, bit in bits.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                ent = vis.split(",")[self.gpu].strip()
                if ent.isdigit():
                    idx = int(ent)
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self.thread.start()
            self.mode = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.sm:
                # the poller never got a sample in (slow NVML calls on a busy host): one synchronous reading, GPU still warm
                try:
                    import pynvml as nv

                    h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                except Exception:
                    pass
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 2 ms polling"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def trained_like_theta(d, K, H, seed=0):
    """PyTorch default init of the reference flow (torch.manual_seed(seed)), state_dict order."""
    import torch

    from nfisam_b200.flows import NSF_AR

    torch.manual_seed(seed)
    flow = NSF_AR(dim=d, K=K, B=TAIL, hidden_dim=H)
    return flow.flat_parameters()


# ------------------------------------------------------------------------------------------------
# CPU arms.  (a) the reference's OWN PyTorch implementation of the path, staged verbatim into oracle/_ref/flows by
# oracle/make_ref.py at build() time (kind "reference"): NormalizingFlowModel.forward, src/flows/models.py:11-24, with
# torch.set_num_threads(host cores);  (b) the OpenMP C port of oracle/ (kind "port"), with the thread count set explicitly
# (torchrun exports OMP_NUM_THREADS=1).  Only these legs may touch oracle/.
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_log_prob_rate(theta, n_sample, seed=123, threads=None):
    from oracle import nsf_oracle as orc

    threads = threads or host_threads()
    used = orc.set_num_threads(threads)
    x = make_inputs(n_sample, D, seed)
    t0 = time.perf_counter()
    orc.log_prob(theta, D, K_BINS, HID, TAIL, x)
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, used


def cpu_baseline_block(theta, target_s=12.0):
    rate, _, used = cpu_log_prob_rate(theta, 100_000)
    n = int(min(max(rate * target_s, 100_000), N_PER_GPU))
    rate, dt, used = cpu_log_prob_rate(theta, n)
    return {"value": rate, "unit": "samples/s", "cores": used, "kind": "port",
            "sample": f"oracle/nsf_oracle.c log_prob (OpenMP, {used} threads) on {n} of the "
                      f"{N_PER_GPU} samples of the workload, {dt:.1f} s"}


def reference_torch_model(theta):
    """The reference's NormalizingFlowModel (oracle/_ref/flows, staged by oracle/make_ref.py) carrying `theta` (state_dict
    order).  Returns None when the staged copy is absent."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "flows", "flows.py")):
        return None
    import torch

    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from flows.flows import NSF_AR as RefNSF
    from flows.models import NormalizingFlowModel as RefModel
    from flows.prior_dist import CustomMultivariateNormal as RefPrior

    flow = RefNSF(dim=D, K=K_BINS, B=TAIL, hidden_dim=HID)
    params = [flow.init_param]
    for layer in flow.layers:
        for j in (0, 2, 4):
            params += [layer.network[j].weight, layer.network[j].bias]
    off = 0
    with torch.no_grad():
        for p in params:
            k = p.numel()
            p.copy_(torch.from_numpy(np.asarray(theta[off:off + k], np.float32).reshape(tuple(p.shape)).copy()))
            off += k
    assert off == len(theta)
    return RefModel(RefPrior(dim=D), [flow])       # the prior NFiSAM.fit_clique_density_model builds (src/slam/NFiSAM.py:389-431)


def torch_reference_rate(model, n_sample, seed=123):
    import torch

    x = torch.from_numpy(make_inputs(n_sample, D, seed))
    with torch.no_grad():
        t0 = time.perf_counter()
        z, prior_logprob, log_det = model.forward(x)
        lp = prior_logprob + log_det           # what the training loss of the reference averages (src/slam/NFiSAM.py:470-472)
        dt = time.perf_counter() - t0
    return n_sample / dt, dt, float(lp.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    theta = trained_like_theta(D, K_BINS, HID)
    threads = host_threads()
    model = reference_torch_model(theta)
    port_rate, _, port_threads = cpu_log_prob_rate(theta, 200_000, threads=threads)
    if model is not None:
        import torch

        torch.set_num_threads(threads)
        kind = "reference"
        rate0, _, _ = torch_reference_rate(model, 50_000)
        n = int(min(max(rate0 * 4.0, 50_000), 2_000_000))   # ~4 s per step; the reference materialises (n*d, K) tensors

        def one():
            return torch_reference_rate(model, n)[0]
        used = torch.get_num_threads()
        what = (f"{n} samples per step through the reference's own NormalizingFlowModel.forward (oracle/_ref/flows, unmodified "
                f"src/flows/*.py, PyTorch CPU, torch.set_num_threads({used}))")
    else:
        kind = "port"
        rate0 = port_rate
        n = int(min(max(rate0 * 4.0, 50_000), N_PER_GPU))

        def one():
            return cpu_log_prob_rate(theta, n, threads=threads)[0]
        used = port_threads
        what = f"{n} samples per step, oracle/nsf_oracle.c with OpenMP on {used} threads (oracle/_ref absent)"
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    rate = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "rqs_flow_log_prob_samples_per_sec", "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[2] synthetic RQS flow microbench: log_prob, dim {D}, K {K_BINS}, hidden {HID}, "
                               f"{N_PER_GPU} samples per GPU (CPU arm: bounded sample of {n} per step)"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": used, "kind": kind, "sample": what,
                         "c_port_samples_per_s": port_rate, "c_port_threads": port_threads,
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_near_gpu(index):
    """Pin this process to the host cores NVML reports as local to GPU `index`, so that the pinned staging buffers
    of the end-to-end leg are allocated on the GPU's own NUMA node (matters when 8 ranks stream H2D at once).
    Returns the number of cores bound to (0 = left unchanged)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus = sorted(local & os.sched_getaffinity(0))
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from nfisam_b200 import _lib
    from nfisam_b200.flows import NSF_AR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    _lib.require_device()
    bound_cores = bind_near_gpu(local_rank)

    n = args.samples
    theta = trained_like_theta(D, K_BINS, HID)
    flow = NSF_AR(dim=D, K=K_BINS, B=TAIL, hidden_dim=HID, device=local_rank, reference_layout=False)
    flow.load_flat_parameters(theta)
    h = flow.handle()
    x_host = torch.from_numpy(make_inputs(n, D, 1000 + rank)).pin_memory()
    x_dev = x_host.to(dev)
    logp = torch.empty(n, dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def step():
        _lib.check(lib.nfisam_flow_log_prob(h, x_dev.data_ptr(), n, D, logp.data_ptr(), stream))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = _lib.launch_count()
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = float(np.mean([ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    out_host = torch.empty(n, dtype=torch.float32).pin_memory()

    def e2e_step():
        _lib.check(lib.nfisam_flow_log_prob_host(h, x_host.data_ptr(), n, D, out_host.data_ptr()))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (nf_log_prob_pair_kernel<9,8>), live CUDA-event timing
        peaks = (ctypes.c_double * 4)()
        _lib.check(lib.nfisam_probe_pipe_peaks(local_rank, peaks))
        fp32_peak = ctypes.c_double(max(peaks[0], peaks[1], peaks[2]))
        mufu_peak = ctypes.c_double(peaks[3])
        fl = flops_fwd(D, HID, K_BINS)
        achieved_tflops = fl * n / (per_launch_ms * 1e-3) * 1e-12
        alg_bytes = (4 * D + 4) * n
        hbm_peak, hbm_src = measured_peaks()
        hbm_ach = alg_bytes / (per_launch_ms * 1e-3) * 1e-9
        roofline = {
            "bound": "fp32_fma", "kernel": "nf_log_prob_pair_kernel<K=9,H=8> (log-prob, two samples per thread, cp.async tile prefetch; batches below 5e5 rows use "
                                          "nf_forward_kernel)",
            "achieved": achieved_tflops, "peak": fp32_peak.value, "unit": "TFLOP/s",
            "frac": achieved_tflops / fp32_peak.value if fp32_peak.value > 0 else None,
            "peak_source": "measured live: best of the FFMA / FFMA2 probe kernels (nfisam_probe_pipe_peaks); "
                           "MEASURED_PEAKS.json holds no FP32 figure. The kernel is FMA/MUFU-bound (95 flop/B), not HBM- or "
                           "tensor-bound (SURVEY.md 8d)",
            "fp32_probe_tflops": {"ffma_reg": peaks[0], "ffma2": peaks[1], "ffma_const": peaks[2]},
            "flops_per_sample": fl, "executed_flops_per_sample": flops_fwd_executed(D, HID, K_BINS),
            "sfu_ops_per_sample": sfu_fwd(D, HID, K_BINS),
            "mufu": {"achieved_gops": sfu_fwd(D, HID, K_BINS) * n / (per_launch_ms * 1e-3) * 1e-9, "peak_gops": mufu_peak.value},
            "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                    "algorithmic_bytes_per_sample": 4 * D + 4, "peak_source": hbm_src},
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, from the `ncu --set full` capture of this
            # same command (profiles/r2_forward_kernel.md, capture r2_prof_fwd_b: 480.2 + 37.5 MB); algorithmic bytes: (4 d + 4) n = 520 MB
            "traffic": 517.7e6 * (n / 1.0e7), "traffic_unit": "bytes per launch",
            "traffic_source": "ncu --set full capture r2_prof_fwd_b of this round's kernel (profiles/r2_forward_kernel.md), scaled by n / 1e7",
            "algorithmic_bytes_per_launch": alg_bytes,
            "launch_ms": per_launch_ms,
        }
        line = {
            "metric": "rqs_flow_log_prob_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2] synthetic RQS flow microbench: log_prob, dim {D}, K {K_BINS}, hidden {HID}, "
                                   f"{n} samples per GPU, 0.5% tail entries, PyTorch-default-init weights (seed 0)",
                       "l2": "inputs (480 MB per pass) larger than the 126 MB L2; no flush needed",
                       "parallelism": "replicas: samples sharded over ranks, no collective on the data path"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": 4 * D * n, "d2h_bytes_per_step": 4 * n,
                    "api": "nfisam_flow_log_prob_host (C ABI, pinned host buffers, 2-stream chunked pipeline)",
                    # what bounds it: the host link.  Per-GPU H2D rate the timed region sustained (the D2H of the results overlaps)
                    "h2d_gbs_per_gpu": 4 * D * e2e_value / world * 1e-9,
                    "host_cores_bound_near_gpu": bound_cores, "host_cores_visible": os.cpu_count()},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
    # ---- companion metric "clique-flow train+sample s/incr-step".  Everything below that involves the solver runs on
    # ALL ranks (round 1 ran it on rank 0 only while the other ranks sat in a barrier: mismatched collectives, NCCL abort).
    extra = None
    if not args.no_extra:
        extra = {}
        if world == 1:
            extra = secondary_measurements(lib, _lib, dev, local_rank)
        group = dist.group.WORLD if world > 1 else None
        extra.update(clique_parallel_solves(group, local_rank))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if extra is not None:
            line["incr_step"] = extra
        if world == 1 and not args.no_cpu:
            # LAST: the OpenMP oracle runs on every host core and raises the thread count of the libgomp instance torch shares;
            # run before the solves it left them with 0.1 s outlier steps (host threads spinning next to the Python thread)
            line["cpu_baseline"] = cpu_baseline_block(theta)
        print(json.dumps(line))


def clique_parallel_solves(group, local_rank):
    """configs[4]: synthetic multi-robot range-SLAM graphs (8 robots, shared landmarks, ambiguous associations with
    probability .4) solved incrementally through the drop-in NFiSAM API; with N > 1 the cliques of a
    tree level are dealt to the N GPUs (NFiSAMArgs.process_group).  Called by every rank."""
    from benchmarks.solve_bench import run_solve

    out = {}
    # warm-up: module loads, kernel attribute set-up, NCCL channels
    run_solve(robots=8, poses=4, ada_prob=0.4, iters=500, samples=2000, process_group=group, device=local_rank)
    detail = bool(os.environ.get("NFISAM_BENCH_DETAIL"))          # diagnostics: per-step times and phase splits on stderr
    out["solve_mr8x64"] = run_solve(robots=8, poses=64, ada_prob=0.4, iters=500, samples=2000, process_group=group, device=local_rank,
                                    detail=detail)
    if detail:
        ps, sp = np.array(out["solve_mr8x64"].pop("per_step")), np.array(out["solve_mr8x64"].pop("splits"))
        gcs = np.array(out["solve_mr8x64"].pop("gc_s"))
        for i in np.argsort(ps)[::-1][:8]:
            print(f"[detail] step {i}: {ps[i]:.4f} s, graph/sim/train/posterior {np.round(sp[i], 4)}, unaccounted {ps[i] - sp[i].sum():.4f}, "
                  f"gc {gcs[i]:.4f}", file=sys.stderr)
    out["solve_mr8x16_n50k"] = run_solve(robots=8, poses=16, ada_prob=0.4, iters=500, samples=50000, process_group=group,
                                        device=local_rank)
    return out


def secondary_measurements(lib, _lib, dev, local_rank):
    """Companion metric: one synthetic clique fit (2000 samples, dim 11, K 9, 2000 Adam iterations,
    the reference's small-graph setting) + 1000-sample posterior draw = seconds per incremental
    step on a chain-shaped Bayes tree (one clique retrained per step)."""
    import torch

    from nfisam_b200.flows import NSF_AR

    rng = np.random.default_rng(7)
    d = 11
    x = rng.standard_normal((2000, d)).astype(np.float32)
    for i in range(1, d):
        x[:, i] = 0.6 * x[:, i] + 0.5 * np.tanh(x[:, i - 1]) ** 2
    x = (x - x.mean(0)) / x.std(0)
    xd = torch.from_numpy(x).to(dev)
    torch.manual_seed(0)
    flow = NSF_AR(dim=d, K=9, hidden_dim=8, device=local_rank)
    theta0 = flow.flat_parameters()
    iters = 2000
    flow.fit(xd, 50, 0.025, average_window=0)          # warm-up (module load, attribute set-up)
    res = {}
    times = []
    for _ in range(3):
        flow.load_flat_parameters(theta0)
        flow.handle()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        hist, ran = flow.fit(xd, iters, 0.025, average_window=0, pull=False)
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
    res["train_2000_iters_s"] = float(np.median(times))
    res["train_us_per_iter"] = 1e6 * res["train_2000_iters_s"] / iters
    res["loss_first_last"] = [float(hist[0]), float(hist[ran - 1])]
    z = torch.randn(1000, d - 5, device=dev)
    xs = xd[:1000, :5].contiguous()
    flow.inverse_given_separator(z, xs)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(10):
        flow.inverse_given_separator(z, xs)
    torch.cuda.synchronize(dev)
    res["sample_1000_s"] = (time.perf_counter() - t0) / 10
    res["s_per_incr_step"] = res["train_2000_iters_s"] + res["sample_1000_s"]
    res["config"] = "1 clique/step: n=2000, dim=11, K=9, hidden=8, 2000 full-batch Adam iterations (no early stop), lr .025; 1000 posterior draws"
    # large-batch training throughput (one Adam step over 1e6 samples, dim 12)
    xl = torch.randn(1_000_000, 12, device=dev)
    fl = NSF_AR(dim=12, K=9, hidden_dim=8, device=local_rank)
    fl.fit(xl, 2, 0.01, average_window=0, pull=False)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    fl.fit(xl, 10, 0.01, average_window=0, pull=False)
    torch.cuda.synchronize(dev)
    res["train_step_1e6_samples_per_s"] = 10 * 1_000_000 / (time.perf_counter() - t0)
    res["solve_small_case1"] = solve_small_graph()
    res["solve_manhattan_plaza_ada"] = solve_chain_graph()
    return res


def solve_chain_graph():
    """configs[3]: the reference's own 136-pose Manhattan-world range-SLAM graph with ambiguous data association
    (example/slam/manhattan_world_with_range/manhattan_plaza/res/seed0/pada0.4_r2_odom0.01_mada3/factor_graph.fg: 4 landmarks,
    59 two- / three-way AmbiguousDataAssociationFactors), 136 incremental steps through the drop-in NFiSAM API with the
    settings of its run_nfisam.py:5-10, 42-46 (K=9, hidden 8, 2000 training samples, 500 Adam iterations without early stop,
    lr .01, 500 posterior samples)."""
    import torch

    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import graph_file_parser, group_nodes_factors_incrementally

    nodes, truth, factors = graph_file_parser(os.path.join(ROOT, "tests", "data", "manhattan_plaza_ada.fg"))
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
    out = None
    for rep in range(2):                  # the first repetition warms up module loads / kernel attributes
        np.random.seed(0)
        torch.manual_seed(0)
        solver = NFiSAM(NFiSAMArgs(num_knots=9, flow_iterations=500, local_sample_num=2000, learning_rate=.01, hidden_dim=8,
                                   elimination_method="pose_first", loss_delta_tol=1e-9, average_window=50, posterior_sample_num=500))
        per_step, splits, cur = [], [], None
        for sn, sf in steps:
            for v in sn:
                solver.add_node(v)
            for f in sf:
                solver.add_factor(f)
            timer = []
            t0 = time.perf_counter()
            solver.update_physical_and_working_graphs(timer=timer)
            cur = solver.incremental_inference(timer=timer)
            per_step.append(time.perf_counter() - t0)
            splits.append(timer)
        per_step = np.asarray(per_step)
        err = float(np.mean([np.linalg.norm(cur[v].mean(0)[:2] - truth[v][:2]) for v in truth if v.name.startswith("X")]))
        ref = {}
        gp = os.path.join(ROOT, "tests", "golden", "solve_manhattan_plaza_ada.npz")
        if os.path.exists(gp):
            g = np.load(gp)
            t = g["timers"]
            ref = {"reference_s_per_step": float(np.mean(t.sum(1))), "reference_train_s_per_step": float(np.mean(t[:, 1:-1].sum(1))),
                   "reference_note": "stored CPU run of the reference on this graph and these settings (tests/golden/solve_manhattan_plaza_ada.npz, "
                                     "2 host threads of the build container)"}
        out = {"graph": "manhattan_plaza pada0.4_r2_odom0.01_mada3 (reference's own, 136 poses, 59 ADA factors)", "steps": len(per_step),
               "s_per_incr_step_mean": float(per_step.mean()), "s_per_incr_step_median": float(np.median(per_step)),
               "s_per_incr_step_p90": float(np.percentile(per_step, 90)), "s_per_incr_step_first": float(per_step[0]),
               "split_mean_graph_sim_train_posterior": [float(x) for x in np.array(splits).mean(0)],
               "mean_pose_error": err, **ref}
    return out


def solve_small_graph():
    """The reference's own CPU-runnable case (configs[0], example/slam/small_range_gaussian_problem/run_nfisam.py:
    K=9, hidden 8, 2000 training samples, <= 2000 Adam iterations with early stop, lr .025, 1000 posterior samples),
    solved incrementally through the drop-in NFiSAM API.  Reports seconds per incremental step and the split
    [graph update, training-set simulation, flow training, posterior sampling]."""
    import torch

    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import graph_file_parser, group_nodes_factors_incrementally

    nodes, truth, factors = graph_file_parser(os.path.join(ROOT, "tests", "data", "small_case1.fg"))
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
    out = None
    for rep in range(2):                  # first repetition warms up module loads / attribute set-up
        np.random.seed(0)
        torch.manual_seed(0)
        solver = NFiSAM(NFiSAMArgs(num_knots=9, flow_iterations=2000, local_sample_num=2000, learning_rate=.025, hidden_dim=8,
                                   elimination_method="pose_first", loss_delta_tol=.01, posterior_sample_num=1000))
        per_step, splits = [], []
        for sn, sf in steps:
            for v in sn:
                solver.add_node(v)
            for f in sf:
                solver.add_factor(f)
            timer = []
            t0 = time.perf_counter()
            solver.update_physical_and_working_graphs(timer=timer)
            cur = solver.incremental_inference(timer=timer)
            per_step.append(time.perf_counter() - t0)
            splits.append([round(t, 5) for t in timer])
        err = float(np.mean([np.linalg.norm(cur[v].mean(0)[:2] - truth[v][:2]) for v in truth]))
        out = {"s_per_incr_step": per_step, "split_graph_sim_train_posterior": splits, "mean_abs_position_error": err,
               "reference_stored_s_per_step": [4.65, 5.59, 4.99, 4.31, 6.40, 6.29],
               "reference_note": "BASELINE.md: stored run of the reference (unknown GPU, cuda_training) on the same graph and settings"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the incr_step companion measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
