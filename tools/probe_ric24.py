"""n=24 sweeps, k_ric24_mma (FP64 tensor cores) vs k_ric24 (DFMA chains, NOC_OPT_NO_TENSOR_SWEEP): device time of the sweep
kernel (library event pairs) of the LTV sweep over HBM records, of the fused from-x0 sweep, and of ONE SLQ backward pass
(max_iter = 1), over batch sizes from one warp per SM to several waves.  One JSON line per (batch, variant)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noc_b200 import capi, problems as pb

N = int(os.environ.get("N24", 100))
batches = [int(b) for b in os.environ.get("B24", "148,592,1776,3552,8192,16384").split(",")]
dev = torch.device("cuda", 0)
s = capi.Solver(0)
stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = torch.from_numpy(pb.pack_qrqf(*pb.default_weights(1))).to(dev)


def prof(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s.profile_reset(); s.profile_enable(True)
    for _ in range(reps): fn()
    torch.cuda.synchronize(); s.profile_enable(False)
    ms, cnt = s.profile_read(capi.KERNEL_RICCATI)
    return ms / max(cnt, 1), cnt // reps


for B in batches:
    x0 = torch.from_numpy(pb.sample_x0(B, seed=9, model=1)).to(dev)
    xg = torch.from_numpy(pb.sample_goals(B, seed=35, model=1)).to(dev)
    xs = torch.from_numpy(pb.hover_state(B, model=1)).to(dev)
    xbar = torch.empty(B, N + 1, 24, **f64); AB = torch.empty(B, N, 768, **f64)
    K = torch.empty(B, N, 8, 24, **f64); K0 = torch.empty(B, 8, 24, **f64); cost = torch.empty(B, **f64)
    st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    dx = torch.empty(B, N + 1, 24, **f64); du = torch.empty(B, N, 8, **f64)
    pm = capi.problem(1, N, flags=capi.FLAG_ASYNC)
    pl = capi.problem(capi.MODEL_LTV_USER, N, n=24, m=8, flags=capi.FLAG_ASYNC)
    with torch.cuda.stream(stream):
        s.rollout_open_loop(pm, B, x0, None, xbar); s.linearize_batch(pm, B, xbar, None, AB)
    stream.synchronize()
    for name, off in (("tensor", 0), ("dfma", 1)):
        s.set_option(capi.OPT_NO_TENSOR_SWEEP, off)
        row = {"B": B, "N": N, "variant": name}
        row["ltv_ms"], _ = prof(lambda: s.lqr_solve_batch(pl, B, AB, qr, x0dev=x0, K=K, K0=K0, cost=cost, status=st))
        row["fused_ms"], _ = prof(lambda: s.lqr_solve_from_x0(pm, B, x0, qr, K=K, K0=K0, cost=cost, status=st))
        for tag, kw in (("slq1_ms", {}), ("slq1_box_ms", dict(u_min=0.0, u_max=4.0))):
            p1 = capi.problem(1, N, max_iter=1, **kw)
            row[tag], row[tag + "_launches"] = prof(lambda: s.slq_solve_batch(p1, B, xs, xg, qr, x=dx, u=du, cost=cost, iters=it, status=st))
        row["ltv_tflops_sym"] = B * N * 71595 / row["ltv_ms"] * 1e-9
        row["slq1_us_per_step"] = row["slq1_ms"] / N * 1e3
        print(json.dumps(row), flush=True)
    del AB, K, xbar
    s.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
