#!/usr/bin/env python
"""Host-side cost of the C-ABI calls: per-call wall time of asynchronous device-pointer calls, and the per-step
distribution of the synchronous host-buffer (e2e) call.  Diagnostic; prints JSON lines."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from noc_b200 import capi, problems as pb

dev = torch.device("cuda", 0)
s = capi.Solver(0)
stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = pb.pack_qrqf(*pb.default_weights()); d_qr = torch.from_numpy(qr).to(dev)

def dev_case(nb, N, prof):
    x0 = pb.sample_x0(nb, seed=4); d_x0 = torch.from_numpy(x0).to(dev)
    K0 = torch.empty(nb, 48, **f64); c = torch.empty(nb, **f64); st = torch.empty(nb, dtype=torch.int32, device=dev)
    p = capi.problem(0, N, flags=capi.FLAG_ASYNC)
    fn = lambda: s.lqr_solve_from_x0(p, nb, d_x0, d_qr, K0=K0, cost=c, status=st)
    fn(); fn(); torch.cuda.synchronize()
    s.profile_reset(); s.profile_enable(prof)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    e0.record(stream)
    for _ in range(10):
        t = time.perf_counter(); fn(); host.append(time.perf_counter() - t)
    e1.record(stream); torch.cuda.synchronize()
    s.profile_enable(False)
    print(json.dumps({"case": "dev", "nb": nb, "N": N, "profiling": prof, "gpu_ms_per_call": e0.elapsed_time(e1) / 10,
                      "host_ms_per_call": [round(1e3 * h, 3) for h in host]}), flush=True)

for prof in (False, True, False):
    dev_case(131072, 50, prof)
dev_case(65536, 100, True)

B, N = 65536, 100
x0 = pb.sample_x0(B)
h_x0 = torch.from_numpy(x0).pin_memory()
h_K0 = torch.zeros(B, 4, 12, dtype=torch.float64).pin_memory()
h_c = torch.zeros(B, dtype=torch.float64).pin_memory(); h_s = torch.zeros(B, dtype=torch.int32).pin_memory()
p = capi.problem(0, N)
fn = lambda: s.lqr_solve_from_x0(p, B, h_x0, qr, K0=h_K0, cost=h_c, status=h_s)
for rnd in range(3):
    for _ in range(5): fn()
    ts = []
    for _ in range(100):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    ts = np.array(ts) * 1e3
    big = [(int(i), round(float(ts[i]), 2)) for i in np.argsort(-ts)[:6]]
    print(json.dumps({"case": "e2e", "round": rnd, "mean_ms": float(ts.mean()), "median_ms": float(np.median(ts)), "largest": big}), flush=True)
s.close()
