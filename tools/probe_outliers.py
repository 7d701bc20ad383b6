#!/usr/bin/env python
"""Where do the e2e outliers come from?  Same from-x0 call in four I/O variants + cudaMemGetInfo alone."""
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
B, N = 65536, 100
x0 = pb.sample_x0(B)
h_x0 = torch.from_numpy(x0).pin_memory(); d_x0 = h_x0.to(dev)
h_K0 = torch.zeros(B, 4, 12, dtype=torch.float64).pin_memory(); d_K0 = torch.zeros(B, 4, 12, **f64)
h_c = torch.zeros(B, dtype=torch.float64).pin_memory(); d_c = torch.zeros(B, **f64)
h_s = torch.zeros(B, dtype=torch.int32).pin_memory(); d_s = torch.zeros(B, dtype=torch.int32, device=dev)
p = capi.problem(0, N)

def stats(name, fn, n=150):
    for _ in range(5): fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    ts = np.array(ts) * 1e3
    print(json.dumps({"case": name, "mean_ms": round(float(ts.mean()), 3), "median_ms": round(float(np.median(ts)), 3),
                      "n_over_2x_median": int((ts > 2 * np.median(ts)).sum()), "max_ms": round(float(ts.max()), 2)}), flush=True)

def memget():
    torch.cuda.mem_get_info()
stats("cudaMemGetInfo_alone", memget, 2000)
stats("all_device_sync", lambda: s.lqr_solve_from_x0(p, B, d_x0, d_qr, K0=d_K0, cost=d_c, status=d_s))
stats("host_x0_device_out", lambda: s.lqr_solve_from_x0(p, B, h_x0, d_qr, K0=d_K0, cost=d_c, status=d_s))
stats("device_x0_host_out", lambda: s.lqr_solve_from_x0(p, B, d_x0, d_qr, K0=h_K0, cost=h_c, status=h_s))
stats("device_x0_host_cost_only", lambda: s.lqr_solve_from_x0(p, B, d_x0, d_qr, cost=h_c, status=h_s))
stats("all_host", lambda: s.lqr_solve_from_x0(p, B, h_x0, qr, K0=h_K0, cost=h_c, status=h_s))
s.set_option(capi.OPT_SINGLE_STREAM, 1) if hasattr(capi, "OPT_SINGLE_STREAM") else None
stats("all_host_single_stream", lambda: s.lqr_solve_from_x0(p, B, h_x0, qr, K0=h_K0, cost=h_c, status=h_s))
def plain_copy():
    d_K0.copy_(h_K0, non_blocking=True); h_K0.copy_(d_K0, non_blocking=True); torch.cuda.synchronize()
stats("torch_h2d_d2h_25MB", plain_copy)
s.close()
