#!/usr/bin/env python
"""bench.py — Riccati solves/sec (n=12, m=4, N=100, fp64) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl noc|reference]

One "step" = one pass of the hot path over one batch: the finite-horizon LTV Riccati sweep of
BASELINE.json configs[1] (batch 65,536 per GPU, distinct A_k,B_k per instance and step), every gain
K_k written.  `value` is timed with the A_k,B_k stream resident in HBM (12.6 GB >> 126 MB L2, so no
L2 flush is needed between steps).  `e2e` goes through the C-ABI with HOST buffers (pinned x0 in; K0,
cost, status out) — nominal rollout, linearisation and sweep all inside the timed region — as mean
and median of >= 50 steps; `e2e_same_contract` / `e2e_all_gains_to_host` are the `value` contract
(all K_k, and A_k,B_k from host memory) through host buffers.  `extra` holds one block per other
BASELINE.json config (1, 3, 4, 5, control-limited 3b / 5b; tools/bench_configs.py), each with its own
roofline and CPU sample.
N > 1 (torchrun, one rank per GPU): every rank sweeps its own 65,536 instances (weak scaling, no
data-path collective), then the library's own NCCL plane (noc_comm) picks the best rollout across the
GPUs without a host round trip; `extra.config4_sharded_1M` is configs[3]'s strong-scaling batch.

`--impl reference`: the reference publishes no solver source, so this arm times the
in-repo CPU oracle (kind "port") on all host cores.  bench.py is one of the three places allowed to
execute oracle/ (tests/, smoke(), and the CPU legs here) — never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "riccati_solves_per_sec_n12_m4_N100_fp64"
UNIT = "solves/s"
N_X, N_U, HORIZON = 12, 4, 100
BATCH_PER_GPU = 65536
BYTES_PER_STEP = 8 * (N_X * N_X + N_X * N_U) + 8 * N_U * N_X       # read A_k,B_k + write K_k = 1920 B
FLOPS_PER_STEP_SYM = 9045                                          # SURVEY.md §8d, symmetry-aware count


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def fp64_peak_tflops():
    """DFMA peak measured by tools/ubench on this pool's B200 (profiles/ubench_r01.json)."""
    p = os.path.join(ROOT, "profiles", "ubench_r01.json")
    try:
        return float(json.load(open(p))["dfma_tflops_sustained"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/noc_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                 "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(power)}
        return out


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_leg(sample: int, threads: int = 0, repeats: int = 1):
    """Oracle LTV sweep over `sample` synthetic instances of the same workload; returns solves/s."""
    from noc_b200 import problems as pb
    from oracle import pyoracle as oracle
    oracle.build()
    if threads <= 0:
        threads = len(os.sched_getaffinity(0))     # not OMP_NUM_THREADS (torchrun sets it to 1)
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(sample, seed=0x4E4F43)
    AB = oracle.make_ab_batch(0, HORIZON, pb.DT, x0)          # untimed input generation
    oracle.lqr_batch(N_X, N_U, HORIZON, AB[:min(sample, 256)], qr, x0dev=x0[:min(sample, 256)],
                     threads=threads)                         # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.lqr_batch(N_X, N_U, HORIZON, AB, qr, x0dev=x0, threads=threads)
        best = min(best, time.perf_counter() - t0)
    return sample / best, (threads if threads > 0 else oracle.max_threads()), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as oracle
    from noc_b200 import problems as pb
    oracle.build()
    sample = args.ref_sample
    # all host threads, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which would make the reference
    # arm single-threaded at N > 1
    threads = len(os.sched_getaffinity(0))
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(sample, seed=0x4E4F43)
    AB = oracle.make_ab_batch(0, HORIZON, pb.DT, x0)
    for _ in range(args.warmup):
        oracle.lqr_batch(N_X, N_U, HORIZON, AB, qr, x0dev=x0, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.lqr_batch(N_X, N_U, HORIZON, AB, qr, x0dev=x0, threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    # the GPU arm's e2e contract on the CPU: x0 -> nominal rollout -> linearise -> sweep -> K0, cost (no A_k,B_k given)
    oracle.lqr_from_x0_batch(0, HORIZON, pb.DT, x0[:256], qr, threads=threads, want_P0=False)
    t0 = time.perf_counter()
    fx_steps = max(1, min(args.steps, 3))
    for _ in range(fx_steps):
        oracle.lqr_from_x0_batch(0, HORIZON, pb.DT, x0, qr, threads=threads, want_P0=False)
    fx = sample * fx_steps / (time.perf_counter() - t0)
    sample_txt = (f"{sample} of the {BATCH_PER_GPU} config-2 instances per step (LTV sweep n=12 m=4 N=100, all K_k "
                  f"written), OpenMP over instances")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "contracts": {
            "value": "A_k,B_k given in host memory -> all N gains K_k, K0, cost (the GPU arm's `value` and "
                     "`e2e_same_contract`)",
            "from_x0": {"value": fx, "unit": UNIT,
                        "what": "x0 given -> nominal rollout -> linearisation -> sweep -> K0, cost (the GPU arm's `e2e`)"}},
        "gpu_launches": 0,
        "note": "KahnShi/NOC ships no solver source (SURVEY.md §0): this is the in-repo CPU oracle, not the reference",
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(gpus, batch=BATCH_PER_GPU):
    return {
        "workload": "configs[1]: finite-horizon LTV LQR Riccati sweep, quadrotor n=12 m=4 N=100, "
                    f"batch {batch} random initial states per GPU, distinct A_k,B_k per instance and step",
        "n": N_X, "m": N_U, "N": HORIZON, "batch_per_gpu": batch, "global_batch": batch * gpus,
        "sharding": f"dp{gpus} (independent instances; in-library NCCL plane: all-gather of the ranks' 16-byte "
                    "(cost, index) pairs + final pick, results left on the device)",
        "l2": "inputs larger than L2 (A_k,B_k stream 12.6 GB per step vs 126 MB L2); no flush needed",
    }


def bind_to_gpu_numa(torch, local):
    """Pin this process to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE the pinned
    host buffers are allocated: first touch then places them on the GPU's NUMA node, and the e2e copies of the eight
    ranks do not all cross one socket's memory controller.  Returns a description for the JSON line."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": txt, "bound": True}
        return {"cpus": txt, "bound": False}
    except Exception as ex:
        return {"bound": False, "why": str(ex)[:80]}


def current_kernel_traffic():
    """dram bytes per launch of the headline kernel from the committed ncu capture — only if that capture was taken on
    the kernel source this library was built from (sha256 of riccati12_mma.cuh); otherwise None, never a stale number."""
    import hashlib
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        rec = json.load(open(tp))
        sha = hashlib.sha256(open(os.path.join(ROOT, "noc_b200", "csrc", "riccati12_mma.cuh"), "rb").read()).hexdigest()
        if rec.get("riccati12_mma_cuh_sha256") == sha:
            return rec.get("k_ric12_mma_dram_bytes_per_launch_65536"), rec.get("source")
    except Exception:
        pass
    return None, None


def host_timed(fn, steps, sync):
    """Wall-clock per-step times of a synchronous host-buffer call (results are in host memory on return)."""
    ts = []
    t0 = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t1)
    sync()
    return time.perf_counter() - t0, ts


# ------------------------------------------------------------------------------------------ GPU arm
def run_noc(args):
    with StdoutToStderr():
        line = _run_noc(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def _run_noc(args):
    import torch
    from noc_b200 import capi, problems as pb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libnoc_b200 has no CPU fallback (use --impl reference "
                         "for the CPU oracle)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(torch, local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, N = args.batch, HORIZON
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(B, seed=0x4E4F43 + rank)

    solver = capi.Solver(local)
    stream = torch.cuda.Stream(device=dev)
    solver.set_stream(stream.cuda_stream)
    # the multi-GPU plane lives in the library (noc_comm: NCCL loaded by libnoc_b200 itself); torch.distributed only
    # carries the 128-byte id, the barriers and the max-over-ranks of the timings
    comm = capi.Comm.from_torch_distributed(solver, dist, dev) if world > 1 else None
    f64 = dict(dtype=torch.float64, device=dev)

    # ---- inputs resident in HBM: x0 -> hover-thrust nominal -> A_k,B_k stream (our own kernels, untimed)
    d_x0 = torch.from_numpy(x0).to(dev)
    d_qr = torch.from_numpy(qr).to(dev)
    d_xbar = torch.empty(B, N + 1, N_X, **f64)
    d_AB = torch.empty(B, N, N_X * N_X + N_X * N_U, **f64)
    d_K = torch.empty(B, N, N_U, N_X, **f64)
    d_K0 = torch.empty(B, N_U, N_X, **f64)
    d_cost = torch.empty(B, **f64)
    d_status = torch.full((B,), -1, dtype=torch.int32, device=dev)
    d_best = torch.zeros(2, **f64)                 # {double cost; int64 index}, written by the library on the device
    torch.cuda.synchronize()
    p_model = capi.problem(capi.MODEL_QUAD12, N, flags=capi.FLAG_ASYNC)
    p_ltv = capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_ASYNC)
    with torch.cuda.stream(stream):
        solver.rollout_open_loop(p_model, B, d_x0, None, d_xbar)
        solver.linearize_batch(p_model, B, d_xbar, None, d_AB)
    stream.synchronize()
    del d_xbar

    def step():
        solver.lqr_solve_batch(p_ltv, B, d_AB, d_qr, x0dev=d_x0, K=d_K, K0=d_K0, cost=d_cost, status=d_status)
        if comm is not None:   # pick the best rollout over all GPUs: local multi-CTA argmin, 16-byte pairs over NVLink, final pick
            comm.best_rollout(d_cost, B, rank * B, d_best)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], **f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = solver.launch_count
    solver.profile_reset()
    solver.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    solver.profile_enable(False)
    ric_ms, ric_n = solver.profile_read(capi.KERNEL_RICCATI)
    launches = solver.launch_count - launches0
    ms = max_over_ranks(ms)
    n_bad = int((d_status != 0).sum().item())
    value = world * B * args.steps / (ms * 1e-3)
    best_pair = None
    if comm is not None:
        bp = d_best.cpu()
        best_pair = {"cost": float(bp[0].item()), "index": int(bp.view(torch.int64)[1].item())}

    # ---- e2e: the same solves through the C-ABI with HOST buffers (pinned, pre-faulted by the warm-up calls), copies
    # inside the timed region.  Contract: x0 in -> K0 (the MPC gain), cost, status out.
    e2e_steps = max(args.steps, 50)
    h_x0 = torch.from_numpy(x0).pin_memory()
    h_K0 = torch.zeros(B, N_U, N_X, dtype=torch.float64).pin_memory()
    h_P0 = torch.zeros(B, N_X, N_X, dtype=torch.float64).pin_memory()
    h_cost = torch.zeros(B, dtype=torch.float64).pin_memory()
    h_status = torch.zeros(B, dtype=torch.int32).pin_memory()
    p_sync = capi.problem(capi.MODEL_QUAD12, N)
    h2d = h_x0.numel() * 8 + qr.size * 8
    d2h = (h_K0.numel() + h_cost.numel()) * 8 + h_status.numel() * 4

    def e2e_step():
        solver.lqr_solve_from_x0(p_sync, B, h_x0, qr, K0=h_K0, cost=h_cost, status=h_status)

    for _ in range(max(args.warmup, 5)):
        e2e_step()
    barrier()
    e2e_s, step_s = host_timed(e2e_step, e2e_steps, torch.cuda.synchronize)
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * B * e2e_steps / e2e_s
    e2e_bad = int((h_status != 0).sum().item())
    e2e_median = max_over_ranks(float(np.median(step_s)))

    # round 1's e2e contract (P0, 1152 B per instance, as well) for continuity
    def e2e_p0_step():
        solver.lqr_solve_from_x0(p_sync, B, h_x0, qr, K0=h_K0, P0=h_P0, cost=h_cost, status=h_status)

    for _ in range(3):
        e2e_p0_step()
    barrier()
    p0_s, p0_steps = host_timed(e2e_p0_step, 20, torch.cuda.synchronize)
    p0_s = max_over_ranks(p0_s)

    # ---- same-contract extras (N = 1): what `value` computes, through host buffers
    allk = same = None
    if world == 1 and not args.no_allk:
        try:   # x0 in, ALL gains out (2.5 GB per step over PCIe: the copy dominates)
            h_K = torch.zeros(B, N, N_U, N_X, dtype=torch.float64).pin_memory()
            solver.lqr_solve_from_x0(p_sync, B, h_x0, qr, K=h_K, K0=h_K0, cost=h_cost, status=h_status)
            reps = 3
            dt_allk, _ = host_timed(lambda: solver.lqr_solve_from_x0(p_sync, B, h_x0, qr, K=h_K, K0=h_K0, cost=h_cost,
                                                                     status=h_status), reps, torch.cuda.synchronize)
            allk = {"value": B * reps / dt_allk, "unit": UNIT, "ms_per_step": 1e3 * dt_allk / reps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + h_K.numel() * 8,
                    "what": "noc_lqr_solve_from_x0: pinned host x0 in; ALL N gains K_k, K0, cost, status to pinned host memory"}
            # A_k,B_k in HOST memory in, all gains out: literally the reference arm's contract.  Bounded batch (the
            # 12.6 GB stream of the full batch would need as much pinned memory); two-buffer pipeline in the library.
            Bs = min(B, 8192)
            h_AB = torch.empty(Bs, N, N_X * N_X + N_X * N_U, dtype=torch.float64).pin_memory()
            h_AB.copy_(d_AB[:Bs])
            p_ltv_sync = capi.problem(capi.MODEL_LTV_USER, N)
            fn = lambda: solver.lqr_solve_batch(p_ltv_sync, Bs, h_AB, qr, x0dev=h_x0[:Bs], K=h_K[:Bs], K0=h_K0[:Bs],
                                                cost=h_cost[:Bs], status=h_status[:Bs])
            fn()
            dt_same, _ = host_timed(fn, reps, torch.cuda.synchronize)
            same = {"value": Bs * reps / dt_same, "unit": UNIT, "ms_per_step": 1e3 * dt_same / reps, "batch": Bs,
                    "h2d_bytes_per_step": h_AB.numel() * 8 + Bs * N_X * 8 + qr.size * 8,
                    "d2h_bytes_per_step": Bs * (N * 48 + 48 + 1) * 8 + Bs * 4,
                    "what": "noc_lqr_solve_batch with A_k,B_k in pinned HOST memory (1536 B per instance-step over PCIe) and "
                            "all N gains K_k, K0, cost, status back to the host: the contract of `value` and of the "
                            "reference arm, PCIe-bound"}
            del h_K, h_AB
        except Exception as ex:   # pinned allocation refused: report, do not fail the bench
            allk = allk or {"unavailable": str(ex)[:120]}
            same = same or {"unavailable": str(ex)[:120]}

    # ---- extra: the from-x0 pipeline with x0 and the outputs resident in HBM (no PCIe), device-timed
    p_dev = capi.problem(capi.MODEL_QUAD12, N, flags=capi.FLAG_ASYNC)

    def pipe_step():
        solver.lqr_solve_from_x0(p_dev, B, d_x0, d_qr, K0=d_K0, cost=d_cost, status=d_status)

    for _ in range(args.warmup):
        pipe_step()
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        pipe_step()
    p1.record(stream)
    barrier()
    pipe_ms = max_over_ranks(p0.elapsed_time(p1) / args.steps)

    # ---- SURVEY 8d's literal input generator (mt19937_64 per instance, body rates U(-0.5,0.5)) on the headline sweep
    survey = None
    if world == 1 and not args.no_extra:
        xs = pb.sample_x0_survey(B)
        d_xs = torch.from_numpy(xs).to(dev)
        d_xb2 = torch.empty(B, N + 1, N_X, **f64)
        with torch.cuda.stream(stream):
            solver.rollout_open_loop(p_model, B, d_xs, None, d_xb2)
            solver.linearize_batch(p_model, B, d_xb2, None, d_AB)
        del d_xb2
        fn = lambda: solver.lqr_solve_batch(p_ltv, B, d_AB, d_qr, x0dev=d_xs, K=d_K, K0=d_K0, cost=d_cost, status=d_status)
        fn()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(5):
            fn()
        s1.record(stream)
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1) / 5
        survey = {"value": B / sms * 1e3, "unit": UNIT, "ms_per_step": sms, "status_not_ok": int((d_status != 0).sum().item()),
                  "what": "the headline sweep on SURVEY.md 8d's literal inputs: std::mt19937_64(0x4E4F43 + instance), body rates "
                          "U(-0.5,0.5) (the default sampler uses U(-0.2,0.2))"}
    # ---- the sweep with its 12x12x16 products on the FP64 tensor cores (DMMA), timed on the headline workload: the
    # warp-per-instance recursion of riccati_scan.cuh, one chunk (no scan), A_k,B_k and K_k in HBM as for `value`
    dmma = None
    if world == 1 and not args.no_extra:
        dmma = {}
        p_tp = capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_ASYNC | capi.FLAG_TIME_PARALLEL)
        solver.set_option(capi.OPT_TIME_PARALLEL_CHUNKS, 1)
        for name, flag in (("dmma", 0), ("dfma", 1)):
            solver.set_option(capi.OPT_TIME_PARALLEL_DFMA, flag)
            fn = lambda: solver.lqr_solve_batch(p_tp, B, d_AB, d_qr, x0dev=d_x0, K=d_K, K0=d_K0, cost=d_cost, status=d_status)
            fn()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            for _ in range(3):
                fn()
            s1.record(stream)
            torch.cuda.synchronize()
            tms = s0.elapsed_time(s1) / 3
            dmma[name] = {"ms_per_step": tms, "value": B / tms * 1e3, "unit": UNIT,
                          "hbm_frac": B * N * BYTES_PER_STEP / (tms * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"]}
        solver.set_option(capi.OPT_TIME_PARALLEL_DFMA, 0)
        solver.set_option(capi.OPT_TIME_PARALLEL_CHUNKS, 0)
        dmma["what"] = ("k_ric12_tp<DMMA> vs <DFMA>, chunks = 1: one warp per instance, G = P F and H = W + F'G as "
                        "mma.sync.m8n8k4.f64 tiles (24 per step) vs DFMA chains; same records and outputs as `value`; the "
                        "headline kernel (`value`) is k_ric12_mma: eight instances per warp, the same products as 21 DMMAs "
                        "per instance-step, bit-identical to the oracle")
        # A/B of the headline kernel: the same call with the products as DFMA chains (k_ric12, round 1's kernel)
        solver.set_option(capi.OPT_NO_TENSOR_SWEEP, 1)
        fn = lambda: solver.lqr_solve_batch(p_ltv, B, d_AB, d_qr, x0dev=d_x0, K=d_K, K0=d_K0, cost=d_cost, status=d_status)
        fn()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(5):
            fn()
        s1.record(stream)
        torch.cuda.synchronize()
        solver.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
        tms = s0.elapsed_time(s1) / 5
        dmma["dfma_chains_k_ric12"] = {"ms_per_step": tms, "value": B / tms * 1e3, "unit": UNIT,
                                       "hbm_frac": B * N * BYTES_PER_STEP / (tms * 1e-3) / 1e9 / measured_peaks()[0]["hbm_gbs"],
                                       "what": "NOC_OPT_NO_TENSOR_SWEEP: the headline call on k_ric12 (DFMA chains), bit-identical outputs"}
    del d_AB, d_K
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs (tools/bench_configs.py): device-timed, each with its roofline and CPU sample
    extra = {}
    from tools import bench_configs as bc
    bb = bc.Bench(torch, solver, stream, dev, quick=args.quick_extra, cpu=not args.no_cpu)
    if not args.no_extra:
        r4 = bb.config4_sharded(comm, world, rank)          # every rank: configs[3] is the sharded 1M-solve batch
        r4["ms"] = max_over_ranks(r4["ms"])
        r4.update({"config": 4, "n_gpus": world, "scaling": "strong", "solves_per_s": r4["total"] / r4["ms"] * 1e3,
                   "what": "configs[3]: 1,048,576 MPC warm-start LQR solves (N=50, K0 / cost) sharded over the GPUs through "
                           "noc_comm_lqr_solve_from_x0 (cost all-gather under the sweep + cross-GPU pick); max over ranks",
                   "roofline": bc.fp64_roofline(float(r4["total"]) / world * r4["N"] * bc.FLOPS_STEP[12], r4["ms"],
                                                "whole step (rollout + k_ric12<RIC_LQR,FUSED> + gather + pick)",
                                                "per-GPU solves x N x 9,045 over the whole step time (dense symmetric count)")})
        extra["config4_sharded_1M"] = r4
        if world == 1:
            for key, fn in (("config1", bb.config1), ("config4_share", bb.config4_share), ("config3", lambda: bb.slq(3)),
                            ("config5", lambda: bb.slq(5)), ("config3b_box", lambda: bb.slq(3, box=(0.0, 4.0))),
                            ("config5b_box", lambda: bb.slq(5, box=(0.0, 4.0))), ("n24_sweeps", bb.ric24), ("generic_kernel", bb.generic)):
                try:
                    extra[key] = fn()
                except Exception as ex:
                    extra[key] = {"failed": f"{type(ex).__name__}: {str(ex)[:200]}"}
                log(f"extra {key}: done")

    line = None
    if rank == 0:
        peaks, peaks_kind = measured_peaks()
        ric_avg_ms = ric_ms / max(ric_n, 1)
        achieved = B * N * BYTES_PER_STEP / (ric_avg_ms * 1e-3) / 1e9
        traffic, traffic_src = current_kernel_traffic()
        fp64_peak = fp64_peak_tflops()
        fp64_ach = B * N * FLOPS_PER_STEP_SYM / (ric_avg_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world, B),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "noc_lqr_solve_from_x0: pinned host x0 -> nominal rollout -> fused linearise+Riccati sweep -> "
                           "pinned host K0, cost, status (device->host copies of finished chunks overlap the next chunk)",
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "median_ms_per_step": 1e3 * e2e_median,
                    "max_ms_per_step": 1e3 * float(np.max(step_s)),
                    "value_at_median": world * B / e2e_median, "host_numa_binding": numa,
                    "contract_note": "outputs are the first-step gain K0 (what an MPC loop applies), cost and status; the "
                                     "reference arm's `contracts.from_x0` times the same contract on the CPU; the same-"
                                     "contract figures for `value` are e2e_same_contract / e2e_all_gains_to_host"},
            "e2e_with_P0": {"value": world * B * 20 / p0_s, "unit": UNIT, "ms_per_step": 1e3 * p0_s / 20,
                            "median_ms_per_step": 1e3 * float(np.median(p0_steps)),
                            "d2h_bytes_per_step": d2h + h_P0.numel() * 8, "what": "round 1's e2e contract: K0, P0, cost, status"},
            "e2e_same_contract": same,
            "e2e_all_gains_to_host": allk,
            "pipeline_device_resident": {"value": world * B / (pipe_ms * 1e-3), "unit": UNIT, "ms_per_step": pipe_ms,
                                         "what": "x0 -> rollout -> fused linearise+sweep -> K0,cost,status, all in HBM (max over ranks)"},
            "config2_survey_inputs": survey,
            "dmma_sweep": dmma,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peaks_kind,
                         "kernel": "k_ric12_mma (n=12 LTV sweep, products on the FP64 tensor cores, A_THEN_B records)",
                         "kernel_ms": ric_avg_ms, "kernel_launches": ric_n,
                         "algorithmic_bytes_per_launch": B * N * BYTES_PER_STEP,
                         "kernel_share_of_step": ric_ms / ms if world == 1 else None,
                         "fp64": {"achieved_tflops_symmetric_count": fp64_ach, "peak_tflops_measured_dfma": fp64_peak,
                                  "frac": (fp64_ach / fp64_peak) if fp64_peak else None}},
            "clocks": clocks,
            "status_not_ok": n_bad + e2e_bad,
            "extra": extra,
        }
        if best_pair is not None:
            line["best_rollout"] = best_pair
        if world == 1 and not args.no_cpu:
            v, cores, secs = cpu_leg(args.cpu_sample)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{args.cpu_sample} of the {B} instances, one pass, {secs:.2f} s wall on {cores} threads "
                          "(in-repo C oracle, OpenMP over instances; the reference has no solver source)"}
    if comm is not None:
        comm.close()
    solver.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


class StdoutToStderr:
    """Route fd 1 to stderr while libraries initialise (NCCL prints its version banner on stdout); rank 0's JSON
    line is the only thing this program writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="noc", choices=["noc", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="instances per GPU (default: config 2)")
    ap.add_argument("--cpu-sample", type=int, default=32768)
    ap.add_argument("--ref-sample", type=int, default=8192)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-allk", action="store_true", help="skip the all-gains-to-host / same-contract extras")
    ap.add_argument("--no-extra", action="store_true", help="skip the blocks of the other BASELINE.json configs")
    ap.add_argument("--quick-extra", action="store_true", help="small batches in the extra blocks (smoke runs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "noc":
        log("note: fewer than 3 warm-up steps requested")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "noc" and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), __file__] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference(args)
    return run_noc(args)


if __name__ == "__main__":
    sys.exit(main())
