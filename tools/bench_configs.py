#!/usr/bin/env python
"""Secondary measurements for BASELINE.md §4: BASELINE.json configs 1, 3, 4, 5 (config 2 is bench.py's line).
One GPU; CUDA-event timing through the C-ABI with device-resident inputs; the CPU oracle (all host threads) on a
bounded sample beside each number.  Writes one JSON object per config to stdout.
    python tools/bench_configs.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ev_time(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", type=int, default=0, help="run a single config (1, 3, 4 or 5)")
    ap.add_argument("--box", type=float, nargs=2, default=None, metavar=("U_MIN", "U_MAX"),
                    help="configs 3 and 5 with the input box u_min <= u <= u_max (control-limited SLQ, DESIGN.md 2.4b)")
    args = ap.parse_args()
    import torch
    from noc_b200 import capi, problems as pb
    from oracle import pyoracle as oracle
    oracle.build()
    dev = torch.device("cuda", 0)
    s = capi.Solver(0)
    stream = torch.cuda.Stream(device=dev)
    s.set_stream(stream.cuda_stream)
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    threads = oracle.max_threads()
    out = []

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    # ---- config 1: infinite-horizon LQR, hover linearisation
    Q, R, Qf = pb.default_weights()
    if args.only in (0, 1):
        A, B = oracle.linearize(0, np.zeros(12), pb.hover_input(), pb.DT)
        t0 = time.perf_counter()
        for _ in range(20):
            P, K, it = oracle.dare(A, B, Q, R, tol=1e-12, max_iter=10000)
        cpu_us = (time.perf_counter() - t0) / 20 * 1e6
        nb = 65536 if not args.quick else 4096
        rng = np.random.default_rng(1)
        recs = np.empty((nb, 192)); xop = np.zeros((nb, 12))
        for b in range(min(nb, 512)):
            x = np.zeros(12); x[6:8] = rng.uniform(-0.2, 0.2, 2); x[8] = rng.uniform(-np.pi, np.pi)
            Ab, Bb = oracle.linearize(0, x, pb.hover_input(), pb.DT)
            recs[b] = np.concatenate([Ab.ravel(), Bb.ravel()]); xop[b] = x
        recs[512:] = recs[np.arange(512, nb) % 512] if nb > 512 else recs[:0]
        xop[512:] = xop[np.arange(512, nb) % 512] if nb > 512 else xop[:0]
        d_rec = T(recs); d_qr = T(np.concatenate([Q.ravel(), R.ravel()]))
        dP = torch.empty(nb, 144, **f64); dK = torch.empty(nb, 48, **f64); dit = torch.empty(nb, **i32); dst = torch.empty(nb, **i32)
        p = capi.problem(capi.MODEL_LTV_USER, 1, tol=1e-12, max_iter=10000, flags=capi.FLAG_ASYNC)
        fn = lambda: s.dare_solve_batch(p, nb, d_rec, d_qr, dP, dK, dit, dst)
        fn(); ms = ev_time(torch, stream, fn, 3)
        its = dit.cpu().numpy()
        # the same operating points through noc_dare_solve_at: linearised in the kernel, structural zeros skipped
        d_xop = T(xop); dP2 = torch.empty_like(dP); dK2 = torch.empty_like(dK); dit2 = torch.empty_like(dit); dst2 = torch.empty_like(dst)
        pm = capi.problem(capi.MODEL_QUAD12, 1, tol=1e-12, max_iter=10000, flags=capi.FLAG_ASYNC)
        fn2 = lambda: s.dare_solve_at(pm, nb, d_xop, None, d_qr, dP2, dK2, dit2, dst2)
        fn2(); ms2 = ev_time(torch, stream, fn2, 3)
        same = bool(torch.equal(dP, dP2) and torch.equal(dK, dK2) and torch.equal(dit, dit2))
        out.append({"config": 1, "what": "DARE fixed point, quadrotor hover, tol 1e-12", "cpu_oracle_us_per_solve_1thread": cpu_us,
                    "gpu_at_operating_points_ms": ms2, "gpu_at_operating_points_solves_per_s": nb / ms2 * 1e3,
                    "at_operating_points_bit_identical": same,
                    "cpu_iterations": int(it), "gpu_batch": nb, "gpu_ms": ms, "gpu_solves_per_s": nb / ms * 1e3,
                    "gpu_riccati_steps_per_s": float(its.sum()) / ms * 1e3, "gpu_iters_min_max": [int(its.min()), int(its.max())],
                    "status_ok": int((dst.cpu().numpy() == 0).sum())})
        print(json.dumps(out[-1]), flush=True)

    # ---- config 4: MPC warm starts, N=50, K0 / cost only (per-GPU share 131,072)
    if args.only in (0, 4):
        qr = pb.pack_qrqf(Q, R, Qf); d_qrqf = T(qr)
        nb = 131072 if not args.quick else 8192
        x0 = pb.sample_x0(nb, seed=4)
        d_x0 = T(x0)
        dK0 = torch.empty(nb, 48, **f64); dc = torch.empty(nb, **f64); dst = torch.empty(nb, **i32)
        p = capi.problem(0, 50, flags=capi.FLAG_ASYNC)
        fn = lambda: s.lqr_solve_from_x0(p, nb, d_x0, d_qrqf, K0=dK0, cost=dc, status=dst)
        fn(); fn()
        s.profile_reset(); s.profile_enable(True)
        ms = ev_time(torch, stream, fn, 5)
        s.profile_enable(False)
        prof4 = {name: s.profile_read(kid)[0] / 5 for name, kid in (("rollout", capi.KERNEL_ROLLOUT), ("linearize", capi.KERNEL_LINEARIZE),
                                                                   ("riccati", capi.KERNEL_RICCATI), ("setup", capi.KERNEL_SETUP))}
        smp = 8192 if not args.quick else 1024
        t0 = time.perf_counter(); oracle.lqr_from_x0_batch(0, 50, pb.DT, x0[:smp], qr, want_P0=False); cpu = smp / (time.perf_counter() - t0)
        out.append({"config": 4, "what": "MPC warm-start LQR N=50 from x0 (rollout+linearise+sweep), K0+cost only, one GPU's share",
                    "gpu_batch": nb, "gpu_ms": ms, "gpu_solves_per_s": nb / ms * 1e3, "cpu_oracle_solves_per_s": cpu,
                    "cpu_threads": threads, "cpu_sample": smp, "status_ok": int((dst.cpu().numpy() == 0).sum()),
                    "kernel_ms": prof4})
        print(json.dumps(out[-1]), flush=True)

    # ---- configs 3 and 5: SLQ
    for cfg, model, N, nb_full, smp in ((3, 0, 200, 16384, 256), (5, 1, 100, 8192, 64)):
        if args.only not in (0, cfg):
            continue
        nb = nb_full if not args.quick else 512
        n, m = pb.dims(model)
        Qm, Rm, Qfm = pb.default_weights(model); qrm = pb.pack_qrqf(Qm, Rm, Qfm); d_q = T(qrm)
        xs = pb.hover_state(nb, model=model); xg = pb.sample_goals(nb, seed=30 + cfg, model=model)
        d_xs, d_xg = T(xs), T(xg)
        dx = torch.empty(nb, N + 1, n, **f64); du = torch.empty(nb, N, m, **f64); dc = torch.empty(nb, **f64)
        dit = torch.empty(nb, **i32); dst = torch.empty(nb, **i32)
        box = dict(u_min=args.box[0], u_max=args.box[1]) if args.box else {}
        p = capi.problem(model, N, **box)
        fn = lambda: s.slq_solve_batch(p, nb, d_xs, d_xg, d_q, x=dx, u=du, cost=dc, iters=dit, status=dst)
        fn()
        s.profile_reset(); s.profile_enable(True)
        ms = ev_time(torch, stream, fn, 2)
        s.profile_enable(False)
        prof = {name: s.profile_read(kid) for name, kid in (("linearize", capi.KERNEL_LINEARIZE), ("riccati", capi.KERNEL_RICCATI),
                                                           ("forward", capi.KERNEL_FORWARD), ("rollout", capi.KERNEL_ROLLOUT))}
        its = dit.cpu().numpy(); st = dst.cpu().numpy()
        t0 = time.perf_counter()
        ref = oracle.slq_batch(oracle.slq_opts(model, N, **box), xs[:smp], xg[:smp], qrm, want_traj=False)
        cpu_s = time.perf_counter() - t0
        out.append({"config": cfg, "what": f"SLQ/iLQR n={n} m={m} N={N}, hover start, waypoint goals, tol 1e-8, max 50 it" +
                    (f", input box [{args.box[0]}, {args.box[1]}]" if box else ""),
                    "gpu_batch": nb, "gpu_ms": ms, "gpu_solves_per_s": nb / ms * 1e3, "gpu_slq_iterations_per_s": float(its.sum()) / ms * 1e3,
                    "iters_mean": float(its.mean()), "iters_max": int(its.max()), "status_counts": {int(k): int((st == k).sum()) for k in np.unique(st)},
                    "kernel_ms_per_solve_call": {k: v[0] / 2 for k, v in prof.items()},
                    "cpu_oracle_solves_per_s": smp / cpu_s, "cpu_oracle_iterations_per_s": float(ref["iters"].sum()) / cpu_s,
                    "cpu_threads": threads, "cpu_sample": smp,
                    "sample_parity": bool(np.array_equal(its[:smp], ref["iters"]) and np.array_equal(dc.cpu().numpy()[:smp], ref["cost"]))})
        print(json.dumps(out[-1]), flush=True)
    s.close()


if __name__ == "__main__":
    main()
