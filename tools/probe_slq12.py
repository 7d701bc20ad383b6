"""One SLQ backward pass (max_iter = 1) at n=12, N=200 over work-list sizes: k_ric12_wpi (+ k_prep12) vs k_ric12<RIC_AFFINE>
(NOC_OPT_NO_TENSOR_SWEEP).  Device time of the sweep kernel and of the step-record kernel (library event pairs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noc_b200 import capi, problems as pb
N = int(os.environ.get("N12", 200))
batches = [int(b) for b in os.environ.get("B12", "148,296,592,1184,1776,2368,3552,4736,5920").split(",")]
dev = torch.device("cuda", 0)
s = capi.Solver(0); stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = torch.from_numpy(pb.pack_qrqf(*pb.default_weights(0))).to(dev)
for B in batches:
    xg = torch.from_numpy(pb.sample_goals(B, seed=33, model=0)).to(dev); xs = torch.from_numpy(pb.hover_state(B, model=0)).to(dev)
    cost = torch.empty(B, **f64); st = torch.empty(B, dtype=torch.int32, device=dev); it = torch.empty(B, dtype=torch.int32, device=dev)
    dx = torch.empty(B, N + 1, 12, **f64); du = torch.empty(B, N, 4, **f64)
    p1 = capi.problem(0, N, max_iter=1)
    row = {"B": B, "N": N}
    for name, off in (("wpi", 0), ("k_ric12", 1)):
        s.set_option(capi.OPT_NO_TENSOR_SWEEP, off)
        fn = lambda: s.slq_solve_batch(p1, B, xs, xg, qr, x=dx, u=du, cost=cost, iters=it, status=st)
        fn(); torch.cuda.synchronize()
        s.profile_reset(); s.profile_enable(True)
        for _ in range(3): fn()
        torch.cuda.synchronize(); s.profile_enable(False)
        row[name + "_sweep_ms"] = round(s.profile_read(capi.KERNEL_RICCATI)[0] / 3, 4)
        row[name + "_prep_ms"] = round(s.profile_read(capi.KERNEL_LINEARIZE)[0] / 3, 4)
    s.set_option(capi.OPT_NO_TENSOR_SWEEP, 0)
    print(json.dumps(row), flush=True)
