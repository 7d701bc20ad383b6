"""Timing experiments on the LTV sweep (value kernel): which outputs / inputs cost what."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noc_b200 import capi, problems as pb
B, N = 65536, int(os.environ.get("NH", 100))
dev = torch.device("cuda", 0)
s = capi.Solver(0)
stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = torch.from_numpy(pb.pack_qrqf(*pb.default_weights())).to(dev)
x0 = torch.from_numpy(pb.sample_x0(B, seed=9)).to(dev)
xbar = torch.empty(B, N + 1, 12, **f64); AB = torch.empty(B, N, 192, **f64)
K = torch.empty(B, N, 4, 12, **f64); K0 = torch.empty(B, 4, 12, **f64); cost = torch.empty(B, **f64); P0 = torch.empty(B, 144, **f64)
st = torch.empty(B, dtype=torch.int32, device=dev)
pm = capi.problem(0, N, flags=capi.FLAG_ASYNC)
pl = capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_ASYNC)
with torch.cuda.stream(stream):
    s.rollout_open_loop(pm, B, x0, None, xbar); s.linearize_batch(pm, B, xbar, None, AB)
stream.synchronize()
def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
out["all_K"] = timeit(lambda: s.lqr_solve_batch(pl, B, AB, qr, x0dev=x0, K=K, K0=K0, cost=cost, status=st))
out["K0_only"] = timeit(lambda: s.lqr_solve_batch(pl, B, AB, qr, x0dev=x0, K0=K0, cost=cost, status=st))
out["K_only_no_cost"] = timeit(lambda: s.lqr_solve_batch(pl, B, AB, qr, K=K))
out["status_only"] = timeit(lambda: s.lqr_solve_batch(pl, B, AB, qr, status=st))
out["fused_K0"] = timeit(lambda: s.lqr_solve_from_x0(pm, B, x0, qr, K0=K0, cost=cost, status=st))
out["fused_allK"] = timeit(lambda: s.lqr_solve_from_x0(pm, B, x0, qr, K=K, K0=K0, cost=cost, status=st))
for b in (1184, 2368, 4736, 9472, 8192, 16384, 32768):
    out[f"all_K_B{b}"] = timeit(lambda: s.lqr_solve_batch(pl, b, AB, qr, x0dev=x0, K=K, K0=K0, cost=cost, status=st))
print(json.dumps({k: round(v, 3) for k, v in out.items()}))
