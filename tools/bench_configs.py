#!/usr/bin/env python
"""Measurements of BASELINE.json configs 1, 3, 4, 5 (+ the control-limited 3b / 5b and the thread-per-instance
kernel's rates) — config 2 is bench.py's headline.  bench.py imports this module and puts every block into its JSON
line (`extra`); run alone it prints one JSON object per block:

    python tools/bench_configs.py [--quick] [--only 1 3 ...] [--no-cpu]

Every block: device-resident inputs and outputs, CUDA events on the launching stream, >= 1 untimed warm-up call, and
    {"ms", "solves_per_s", ..., "roofline": {"bound", "achieved", "peak", "frac", "unit", "kernel", "kernel_ms"},
     "cpu_baseline": {...}}                     (cpu_baseline: the in-repo CPU oracle on a bounded sample, kind "port")
Roofline convention: FP64-bound work is counted with the SYMMETRY-AWARE DENSE flop count per Riccati
step (9,045 at n=12/m=4, 71,595 at n=24/m=8; + the affine iLQR terms for SLQ) against the measured DFMA peak
(profiles/ubench_r01.json).  The fused kernels skip the structural zeros of the built-in Jacobians, so they execute
fewer flops than they are credited with — `flops_model` says so in every block.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_STEP = {12: 9045, 24: 71595}          # symmetry-aware dense count per Riccati step
FLOPS_AFFINE = {12: 600, 24: 2400}          # affine iLQR extras per step (same table)
BYTES_STEP = {12: 1920, 24: 7680}           # compulsory HBM bytes per step of the UNFUSED sweep (A_k,B_k in, K_k out)


def fp64_peak_tflops():
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ubench_r01.json")))["dfma_tflops_sustained"])
    except Exception:
        return 34.2


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def ev_time(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fp64_roofline(flops, kernel_ms, kernel, model):
    peak = fp64_peak_tflops()
    ach = flops / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
    return {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "kernel": kernel,
            "kernel_ms": kernel_ms, "flops_model": model, "peak_source": "measured DFMA stream, profiles/ubench_r01.json"}


class Bench:
    """Shared state of the config blocks: one solver on one stream of one GPU."""

    def __init__(self, torch, solver, stream, dev, quick=False, cpu=True):
        from noc_b200 import capi, problems as pb
        self.torch, self.s, self.stream, self.dev, self.quick, self.cpu = torch, solver, stream, dev, quick, cpu
        self.capi, self.pb = capi, pb
        self.f64 = dict(dtype=torch.float64, device=dev)
        self.i32 = dict(dtype=torch.int32, device=dev)
        self._oracle = None

    @property
    def oracle(self):
        if self._oracle is None:
            from oracle import pyoracle
            pyoracle.build()
            self._oracle = pyoracle
        return self._oracle

    def T(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def profiled(self, fn, reps):
        """Run fn `reps` times with the library's per-kernel event timing on; returns (ms per call, {kernel id: ms per call})."""
        capi = self.capi
        self.s.profile_reset(); self.s.profile_enable(True)
        ms = ev_time(self.torch, self.stream, fn, reps)
        self.s.profile_enable(False)
        names = (("setup", capi.KERNEL_SETUP), ("rollout", capi.KERNEL_ROLLOUT), ("linearize", capi.KERNEL_LINEARIZE),
                 ("riccati", capi.KERNEL_RICCATI), ("forward", capi.KERNEL_FORWARD))
        prof = {}
        for name, kid in names:
            t, n = self.s.profile_read(kid)
            if n:
                prof[name] = {"ms": t / reps, "launches": n // reps}
        return ms, prof

    # ------------------------------------------------------------------------------------------- config 1
    def config1(self):
        """Infinite-horizon LQR (DARE fixed point) — configs[0] is the reference's single-instance CPU case; the GPU
        number is a batch of operating points (hover at random attitudes / yaw)."""
        torch, s, capi, pb = self.torch, self.s, self.capi, self.pb
        Q, R, _ = pb.default_weights()
        nb = 4096 if self.quick else 65536
        rng = np.random.default_rng(1)
        xop = np.zeros((nb, 12))
        xop[:, 6:8] = rng.uniform(-0.2, 0.2, (nb, 2))
        xop[:, 8] = rng.uniform(-np.pi, np.pi, nb)
        d_x = self.T(xop); d_qr = self.T(np.concatenate([Q.ravel(), R.ravel()]))
        dP = torch.empty(nb, 144, **self.f64); dK = torch.empty(nb, 48, **self.f64)
        dit = torch.empty(nb, **self.i32); dst = torch.empty(nb, **self.i32)
        pm = capi.problem(capi.MODEL_QUAD12, 1, tol=1e-12, max_iter=10000, flags=capi.FLAG_ASYNC)
        fn = lambda: s.dare_solve_at(pm, nb, d_x, None, d_qr, dP, dK, dit, dst)
        fn()
        ms, prof = self.profiled(fn, 3)
        its = dit.cpu().numpy().astype(np.int64)
        kms = prof["riccati"]["ms"]
        # a warp runs until its slowest instance converges: credit the iterations the instances needed, not the warps'
        out = {"config": 1, "what": "infinite-horizon LQR (DARE fixed point, tol 1e-12) of the quadrotor linearised at "
                                    f"{nb} hover operating points (noc_dare_solve_at: record built in the kernel)",
               "batch": nb, "ms": ms, "solves_per_s": nb / ms * 1e3, "iterations_per_s": float(its.sum()) / ms * 1e3,
               "iters_min_max": [int(its.min()), int(its.max())], "status_not_ok": int((dst != 0).sum().item()),
               "kernel_ms": prof,
               "roofline": fp64_roofline(float(its.sum()) * FLOPS_STEP[12], kms, "k_ric12<RIC_DARE,FUSED>",
                                         "iterations x 9,045 (dense symmetric count; the kernel skips the Jacobian's zeros)")}
        if self.cpu:
            o = self.oracle
            A, B = o.linearize(0, np.zeros(12), pb.hover_input(), pb.DT)
            t0 = time.perf_counter(); reps = 20
            for _ in range(reps):
                P, K, it = o.dare(A, B, Q, R, tol=1e-12, max_iter=10000)
            us = (time.perf_counter() - t0) / reps * 1e6
            out["cpu_baseline"] = {"value": 1e6 / us, "unit": "solves/s", "cores": 1, "kind": "port", "iterations": int(it),
                                   "sample": f"configs[0] itself: one hover DARE, {us:.0f} us per solve, single thread"}
        return out

    # ------------------------------------------------------------------------------------------- config 4
    def config4_share(self):
        """One GPU's share of the 1M-solve MPC batch (131,072 x N=50, K0 + cost only, from x0)."""
        torch, s, capi, pb = self.torch, self.s, self.capi, self.pb
        nb, N = (8192 if self.quick else 131072), 50
        qr = pb.pack_qrqf(*pb.default_weights()); d_qr = self.T(qr)
        x0 = pb.sample_x0(nb, seed=4); d_x0 = self.T(x0)
        dK0 = torch.empty(nb, 48, **self.f64); dc = torch.empty(nb, **self.f64); dst = torch.empty(nb, **self.i32)
        p = capi.problem(0, N, flags=capi.FLAG_ASYNC)
        fn = lambda: s.lqr_solve_from_x0(p, nb, d_x0, d_qr, K0=dK0, cost=dc, status=dst)
        fn(); fn()
        ms, prof = self.profiled(fn, 5)
        out = {"config": 4, "what": f"MPC warm-start LQR, N=50, x0 -> nominal rollout -> fused linearise+sweep -> K0 + cost; "
                                    f"one GPU's share ({nb}) of the 1,048,576-solve batch",
               "batch": nb, "N": N, "ms": ms, "solves_per_s": nb / ms * 1e3, "status_not_ok": int((dst != 0).sum().item()),
               "kernel_ms": prof,
               "roofline": fp64_roofline(float(nb) * N * FLOPS_STEP[12], prof["riccati"]["ms"], "k_ric12<RIC_LQR,FUSED>",
                                         "batch x N x 9,045 (dense symmetric count; the kernel skips the Jacobian's zeros; "
                                         "no A_k,B_k stream exists on this path, so there is no HBM roofline)")}
        if self.cpu:
            o = self.oracle
            smp = 1024 if self.quick else 16384
            t0 = time.perf_counter(); o.lqr_from_x0_batch(0, N, pb.DT, x0[:smp], qr, want_P0=False); dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": smp / dt, "unit": "solves/s", "cores": o.max_threads(), "kind": "port",
                                   "sample": f"{smp} instances, one pass, {dt:.2f} s (rollout + linearise + sweep)"}
        return out

    def config4_sharded(self, comm, world, rank):
        """configs[3] proper: 1,048,576 solves sharded over the ranks (STRONG scaling), through the in-library plane:
        each rank's slice -> sweep in wave-sized chunks, costs gathered over NVLink under the next chunk, best rollout
        picked across GPUs, no host round trip inside a step.  Returns the rank-local ms (bench.py takes the max)."""
        torch, s, capi, pb = self.torch, self.s, self.capi, self.pb
        TOTAL, N = ((1 << 16) if self.quick else (1 << 20)), 50
        cnt = TOTAL // world
        qr = pb.pack_qrqf(*pb.default_weights()); d_qr = self.T(qr)
        d_x0 = self.T(pb.sample_x0_survey(cnt, first_instance=rank * cnt))
        cost_all = torch.empty(world * cnt, **self.f64)
        dst = torch.empty(cnt, **self.i32)
        best = torch.zeros(2, **self.f64)
        p = capi.problem(0, N, flags=capi.FLAG_ASYNC)
        if comm is not None:
            fn = lambda: comm.lqr_solve_from_x0(p, cnt, d_x0, d_qr, best, cost_all=cost_all, status=dst)
        else:
            def fn():
                s.lqr_solve_from_x0(p, cnt, d_x0, d_qr, cost=cost_all, status=dst)
                s.argmin_cost_async(cost_all, cnt, best)
        for _ in range(3):
            fn()
        ms = ev_time(torch, self.stream, fn, 5)
        b = best.cpu()
        return {"ms": ms, "total": TOTAL, "N": N, "per_gpu": cnt, "status_not_ok": int((dst != 0).sum().item()),
                "best_cost": float(b[0].item()), "best_index": int(b.view(torch.int64)[1].item()),
                "inputs": "SURVEY.md 8d sampler (std::mt19937_64 per instance, body rates U(-0.5,0.5))"}

    # ------------------------------------------------------------------------------------------- configs 3 / 5
    def slq(self, cfg, box=None):
        torch, s, capi, pb = self.torch, self.s, self.capi, self.pb
        model, N, nb_full, smp = {3: (0, 200, 16384, 256), 5: (1, 100, 8192, 64)}[cfg]
        nb = 512 if self.quick else nb_full
        n, m = pb.dims(model)
        qr = pb.pack_qrqf(*pb.default_weights(model)); d_q = self.T(qr)
        xs = pb.hover_state(nb, model=model); xg = pb.sample_goals(nb, seed=30 + cfg, model=model)
        d_xs, d_xg = self.T(xs), self.T(xg)
        dx = torch.empty(nb, N + 1, n, **self.f64); du = torch.empty(nb, N, m, **self.f64); dc = torch.empty(nb, **self.f64)
        dit = torch.empty(nb, **self.i32); dst = torch.empty(nb, **self.i32)
        kw = dict(u_min=box[0], u_max=box[1]) if box else {}
        p = capi.problem(model, N, **kw)
        fn = lambda: s.slq_solve_batch(p, nb, d_xs, d_xg, d_q, x=dx, u=du, cost=dc, iters=dit, status=dst)
        fn()
        ms, prof = self.profiled(fn, 2)
        its = dit.cpu().numpy().astype(np.int64); st = dst.cpu().numpy()
        kms = prof["riccati"]["ms"]
        kern = (f"k_ric12<RIC_AFFINE,FUSED" + (",BOX>" if box else "> (work lists of at most 12 instances per SM: k_prep12 + k_ric12_wpi, "
                                                     "one instance per warp on the FP64 tensor cores)")) if n == 12 else \
               ("k_ric24_mma<RIC_AFFINE,FUSED" + (",BOX" if box else "") + ",PREP> (FP64 tensor cores: 125 DMMAs per step; step records by k_prep24)")
        skips = ("the fused sweep skips the Jacobian's zeros" if n == 12 else
                 "the tensor sweep issues 125 of the 144 dense 8x8x4 tiles: four tiles of [B|A] are structurally zero")
        out = {"config": f"{cfg}b" if box else cfg,
               "what": f"SLQ/iLQR n={n} m={m} N={N}, hover start, waypoint goals, tol 1e-8, max 50 iterations, 11 step lengths" +
                       (f", input box [{box[0]}, {box[1]}] (control-limited DDP)" if box else ""),
               "batch": nb, "ms": ms, "solves_per_s": nb / ms * 1e3, "iterations_per_s": float(its.sum()) / ms * 1e3,
               "iters_mean": float(its.mean()), "iters_max": int(its.max()),
               "status_counts": {int(k): int((st == k).sum()) for k in np.unique(st)}, "kernel_ms": prof,
               "roofline": fp64_roofline(float(its.sum()) * N * (FLOPS_STEP[n] + FLOPS_AFFINE[n]), kms, kern,
                                         f"SLQ iterations x N x ({FLOPS_STEP[n]} + {FLOPS_AFFINE[n]} affine) for the backward "
                                         f"pass (dense symmetric count; {skips}; "
                                         "regularisation-retry sweeps and the forward pass are not credited)"),
               "backward_share_of_solve": kms / ms}
        if self.cpu:
            o = self.oracle
            smp = min(smp, nb)
            t0 = time.perf_counter()
            ref = o.slq_batch(o.slq_opts(model, N, **kw), xs[:smp], xg[:smp], qr, want_traj=False)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": smp / dt, "unit": "solves/s", "iterations_per_s": float(ref["iters"].sum()) / dt,
                                   "cores": o.max_threads(), "kind": "port", "sample": f"the first {smp} instances, {dt:.2f} s"}
            out["sample_parity_bit_exact"] = bool(np.array_equal(its[:smp], ref["iters"]) and
                                                  np.array_equal(dc.cpu().numpy()[:smp], ref["cost"]))
        return out

    # ------------------------------------------------------------------------------------------- n=24 sweeps, tensor vs DFMA
    def ric24(self):
        """The n=24 / m=8 sweeps alone (8192 x N=100): LTV sweep over A_k,B_k records in HBM and the fused from-x0 sweep, on
        k_ric24_mma (FP64 tensor cores, default) and on k_ric24 (DFMA chains, NOC_OPT_NO_TENSOR_SWEEP) — same bits."""
        torch, s, capi, pb = self.torch, self.s, self.capi, self.pb
        nb, N = (1024 if self.quick else 8192), 100
        qr = self.T(pb.pack_qrqf(*pb.default_weights(1)))
        x0 = self.T(pb.sample_x0(nb, seed=9, model=1))
        xbar = torch.empty(nb, N + 1, 24, **self.f64); AB = torch.empty(nb, N, 768, **self.f64)
        K = torch.empty(nb, N, 8, 24, **self.f64); K0 = torch.empty(nb, 8, 24, **self.f64); cost = torch.empty(nb, **self.f64)
        st = torch.empty(nb, **self.i32)
        pm = capi.problem(1, N, flags=capi.FLAG_ASYNC)
        pl = capi.problem(capi.MODEL_LTV_USER, N, n=24, m=8, flags=capi.FLAG_ASYNC)
        s.rollout_open_loop(pm, nb, x0, None, xbar); s.linearize_batch(pm, nb, xbar, None, AB)
        out = {"batch": nb, "N": N, "what": "n=24 m=8 sweeps: LTV over HBM records (all K_k written) and fused from-x0 (K0 + cost)"}
        ref = None
        for name, off in (("tensor", 0), ("dfma", 1)):
            s.set_option(capi.OPT_NO_TENSOR_SWEEP, off)
            f1 = lambda: s.lqr_solve_batch(pl, nb, AB, qr, x0dev=x0, K=K, K0=K0, cost=cost, status=st)
            f2 = lambda: s.lqr_solve_from_x0(pm, nb, x0, qr, K0=K0, cost=cost, status=st)
            f1(); _, p1 = self.profiled(f1, 3)
            kk = K.clone()
            f2(); _, p2 = self.profiled(f2, 3)
            ms = p1["riccati"]["ms"]
            blk = {"ltv_kernel_ms": ms, "ltv_solves_per_s": nb / ms * 1e3,
                   "ltv_roofline": fp64_roofline(float(nb) * N * FLOPS_STEP[24], ms,
                                                 "k_ric24_mma<RIC_LQR> (144 DMMAs per step)" if not off else "k_ric24<RIC_LQR>",
                                                 "batch x N x 71,595 (dense symmetric count)"),
                   "ltv_hbm_frac": nb * N * BYTES_STEP[24] / (ms * 1e-3) / 1e9 / hbm_peak_gbs(),
                   "fused_kernel_ms": p2["riccati"]["ms"], "fused_solves_per_s": nb / p2["riccati"]["ms"] * 1e3,
                   "status_not_ok": int((st != 0).sum().item())}
            if ref is None: ref = kk
            else: blk["gains_bit_identical_to_tensor"] = bool(torch.equal(ref, kk))
            out[name] = blk
        s.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
        # per-instance weights: k_ric24_mma<PERW> (each warp stages its instance's W' tiles); the thread-per-instance
        # kernel, which ran this call before, on a small batch beside it
        qrb = (qr.reshape(1, -1) * torch.linspace(0.5, 2.0, nb, **self.f64).reshape(-1, 1)).contiguous()
        f3 = lambda: s.lqr_solve_batch(pl, nb, AB, qrb, x0dev=x0, K=K, K0=K0, cost=cost, status=st, qr_per_instance=True)
        f3(); _, p3 = self.profiled(f3, 3)
        ms = p3["riccati"]["ms"]
        nbg = min(nb, 512)
        f4 = lambda: s.lqr_solve_batch(pl, nbg, AB, qrb, x0dev=x0, K=K, K0=K0, cost=cost, status=st, qr_per_instance=True)
        s.set_option(capi.OPT_NO_TENSOR_SWEEP, 1)
        try:
            f4(); msg = ev_time(torch, self.stream, f4, 1)
        finally:
            s.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
        out["per_instance_weights"] = {
            "ltv_kernel_ms": ms, "ltv_solves_per_s": nb / ms * 1e3, "status_not_ok": int((st[:nb] != 0).sum().item()),
            "ltv_roofline": fp64_roofline(float(nb) * N * FLOPS_STEP[24], ms, "k_ric24_mma<RIC_LQR, PERW>",
                                          "batch x N x 71,595 (dense symmetric count)"),
            "thread_per_instance_kernel": {"batch": nbg, "ms": msg, "solves_per_s": nbg / msg * 1e3}}
        return out

    # ------------------------------------------------------------------------------------------- thread-per-instance kernel
    def generic(self):
        """The two calls that ran on the thread-per-instance kernel k_ric_generic in round 1, on the kernels
        that took them over — the LTV sweep with per-instance weights (k_ric12_mma<PERW>) and DARE at n=24/m=8
        (k_ric24_mma<RIC_DARE>) — each with k_ric_generic's rate on a smaller batch beside it.  (The block keeps its name.)"""
        torch, s, capi, pb = self.torch, self.s, self.capi, self.pb
        out = {}
        # per-instance weights, (12,4), N=100: k_ric12_mma<PERW> (round 1/2: the thread-per-instance kernel, 0.28 M solves/s),
        # and the same call forced onto the thread-per-instance kernel (NOC_OPT_NO_TENSOR_SWEEP) on a smaller batch
        nb, N = (2048 if self.quick else 65536), 100
        qr = pb.pack_qrqf(*pb.default_weights())
        x0 = pb.sample_x0(nb, seed=9); d_x0 = self.T(x0)
        d_xbar = torch.empty(nb, N + 1, 12, **self.f64); d_AB = torch.empty(nb, N, 192, **self.f64)
        pm = capi.problem(capi.MODEL_QUAD12, N, flags=capi.FLAG_ASYNC)
        s.rollout_open_loop(pm, nb, d_x0, None, d_xbar); s.linearize_batch(pm, nb, d_xbar, None, d_AB)
        scale = np.linspace(0.5, 2.0, nb)[:, None]                      # every instance its own weights
        d_qrb = self.T(np.tile(qr, (nb, 1)) * scale)
        dK = torch.empty(nb, N, 48, **self.f64)
        dK0 = torch.empty(nb, 48, **self.f64); dst = torch.empty(nb, **self.i32)
        pl = capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_ASYNC)
        fn = lambda: s.lqr_solve_batch(pl, nb, d_AB, d_qrb, K=dK, K0=dK0, status=dst, qr_per_instance=True)
        fn()
        ms = ev_time(torch, self.stream, fn, 10)
        nbg = min(nb, 8192)
        fng = lambda: s.lqr_solve_batch(pl, nbg, d_AB, d_qrb, K=dK, K0=dK0, status=dst, qr_per_instance=True)
        s.set_option(capi.OPT_NO_TENSOR_SWEEP, 1)
        try:
            fng()
            msg = ev_time(torch, self.stream, fng, 2)
        finally:
            s.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
        gbs = nb * (N * BYTES_STEP[12] + 304 * 8) / (ms * 1e-3) / 1e9
        out["per_instance_weights_n12"] = {
            "batch": nb, "N": N, "ms": ms, "solves_per_s": nb / ms * 1e3, "status_not_ok": int((dst != 0).sum().item()),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak_gbs(), "unit": "GB/s", "frac": gbs / hbm_peak_gbs(),
                         "kernel": "k_ric12_mma<PERW>", "bytes_model": "batch x (N x 1920 + 2432)"},
            "fp64": fp64_roofline(float(nb) * N * FLOPS_STEP[12], ms, "k_ric12_mma<PERW>", "batch x N x 9,045"),
            "thread_per_instance_kernel": {"batch": nbg, "ms": msg, "solves_per_s": nbg / msg * 1e3,
                                           "roofline": fp64_roofline(float(nbg) * N * FLOPS_STEP[12], msg, "k_ric_generic", "batch x N x 9,045")},
            "what": "noc_lqr_solve_batch(qr_per_instance=1), A_k,B_k and the [batch][304] weights resident in HBM, all K_k written"}
        del dK
        # DARE n=24: the dual model linearised at its hover rest configuration
        nb2 = 512 if self.quick else 4096
        o_rec = None
        try:
            xh = pb.hover_state(1, model=1)
            d_xh = self.T(np.repeat(xh, 2, axis=0).reshape(1, 2, 24))
            d_ab = torch.empty(1, 1, 24 * 24 + 24 * 8, **self.f64)
            p24 = capi.problem(capi.MODEL_DUAL24, 1, flags=capi.FLAG_ASYNC)
            s.linearize_batch(p24, 1, d_xh, None, d_ab)
            o_rec = d_ab.reshape(1, -1).repeat(nb2, 1).contiguous()
        except Exception as ex:  # pragma: no cover
            out["dare_n24"] = {"unavailable": str(ex)[:160]}
        if o_rec is not None:
            Q, R, _ = pb.default_weights(1)
            d_qr = self.T(np.concatenate([Q.ravel(), R.ravel()]))
            dP = torch.empty(nb2, 576, **self.f64); dK = torch.empty(nb2, 192, **self.f64)
            dit = torch.empty(nb2, **self.i32); dst2 = torch.empty(nb2, **self.i32)
            pd = capi.problem(capi.MODEL_LTV_USER, 1, n=24, m=8, tol=1e-12, max_iter=10000, flags=capi.FLAG_ASYNC)
            fn2 = lambda: s.dare_solve_batch(pd, nb2, o_rec, d_qr, dP, dK, dit, dst2)
            fn2()
            ms2 = ev_time(torch, self.stream, fn2, 2)
            its = dit.cpu().numpy().astype(np.int64)
            out["dare_n24"] = {"batch": nb2, "ms": ms2, "solves_per_s": nb2 / ms2 * 1e3, "iterations": int(its.max()),
                               "status_not_ok": int((dst2 != 0).sum().item()),
                               "roofline": fp64_roofline(float(its.sum()) * FLOPS_STEP[24], ms2,
                                                         "k_ric24_mma<RIC_DARE> (FP64 tensor cores, record fragments resident in registers)",
                                                         "iterations x 71,595")}
            # the thread-per-instance kernel it replaced (still the path for ROWS_AB records and other shapes)
            s.set_option(capi.OPT_NO_TENSOR_SWEEP, 1)
            nb3 = min(nb2, 512)
            fn3 = lambda: s.dare_solve_batch(pd, nb3, o_rec, d_qr, dP, dK, dit, dst2)
            fn3()
            ms3 = ev_time(torch, self.stream, fn3, 1)
            s.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
            out["dare_n24"]["k_ric_generic"] = {"batch": nb3, "ms": ms3, "solves_per_s": nb3 / ms3 * 1e3}
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", nargs="*", default=[], help="blocks to run: 1 3 3b 4 5 5b ric24 generic")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--box", type=float, nargs=2, default=[0.0, 4.0], metavar=("U_MIN", "U_MAX"))
    args = ap.parse_args()
    import torch
    from noc_b200 import capi
    dev = torch.device("cuda", 0)
    s = capi.Solver(0)
    stream = torch.cuda.Stream(device=dev)
    s.set_stream(stream.cuda_stream)
    b = Bench(torch, s, stream, dev, quick=args.quick, cpu=not args.no_cpu)
    want = lambda k: not args.only or k in args.only
    blocks = (("1", b.config1), ("4", b.config4_share), ("3", lambda: b.slq(3)), ("5", lambda: b.slq(5)),
              ("3b", lambda: b.slq(3, box=tuple(args.box))), ("5b", lambda: b.slq(5, box=tuple(args.box))), ("ric24", b.ric24), ("generic", b.generic))
    for key, fn in blocks:
        if want(key):
            print(json.dumps(fn()), flush=True)
    s.close()


if __name__ == "__main__":
    main()
