"""n=24 / m=8 LTV sweep (k_ric24) throughput with the A_k,B_k stream resident in HBM, and the fused from-x0 path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noc_b200 import capi, problems as pb
B, N = int(os.environ.get("B24", 8192)), 100
dev = torch.device("cuda", 0)
s = capi.Solver(0)
stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = torch.from_numpy(pb.pack_qrqf(*pb.default_weights(1))).to(dev)
x0 = torch.from_numpy(pb.sample_x0(B, seed=9, model=1)).to(dev)
xbar = torch.empty(B, N + 1, 24, **f64); AB = torch.empty(B, N, 768, **f64)
K = torch.empty(B, N, 8, 24, **f64); K0 = torch.empty(B, 8, 24, **f64); cost = torch.empty(B, **f64)
st = torch.empty(B, dtype=torch.int32, device=dev)
pm = capi.problem(1, N, flags=capi.FLAG_ASYNC)
pl = capi.problem(capi.MODEL_LTV_USER, N, n=24, m=8, flags=capi.FLAG_ASYNC)
with torch.cuda.stream(stream):
    s.rollout_open_loop(pm, B, x0, None, xbar); s.linearize_batch(pm, B, xbar, None, AB)
stream.synchronize()
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms_ltv = timeit(lambda: s.lqr_solve_batch(pl, B, AB, qr, x0dev=x0, K=K, K0=K0, cost=cost, status=st))
ms_fused = timeit(lambda: s.lqr_solve_from_x0(pm, B, x0, qr, K=K, K0=K0, cost=cost, status=st))
flops = B * N * 71595
print(json.dumps({"B": B, "N": N, "ltv_ms": ms_ltv, "ltv_solves_per_s": B / ms_ltv * 1e3, "ltv_tflops_sym": flops / ms_ltv * 1e-9,
                  "ltv_hbm_gbs": B * N * 7680 / ms_ltv * 1e-6, "fused_from_x0_ms": ms_fused, "fused_solves_per_s": B / ms_fused * 1e3,
                  "status_bad": int((st != 0).sum().item())}))
