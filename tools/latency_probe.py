"""Single-call latency of the C-ABI for small batches (host buffers), next to the CPU oracle."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noc_b200 import capi, problems as pb
from oracle import pyoracle as oracle
qr = pb.pack_qrqf(*pb.default_weights())
s = capi.Solver(0)
out = []
for B in (1, 8, 64, 512, 4096):
    x0 = pb.sample_x0(B, seed=3)
    K0 = np.empty((B, 4, 12)); cost = np.empty(B); st = np.empty(B, dtype=np.int32)
    p = capi.problem(0, 100)
    for _ in range(5): s.lqr_solve_from_x0(p, B, x0, qr, K0=K0, cost=cost, status=st)
    ts = []
    for _ in range(200):
        t = time.perf_counter(); s.lqr_solve_from_x0(p, B, x0, qr, K0=K0, cost=cost, status=st); ts.append(time.perf_counter() - t)
    gpu_us = float(np.median(ts)) * 1e6   # median: the VM shows occasional multi-ms stalls that a mean over 50 calls absorbs
    p90_us = float(np.percentile(ts, 90)) * 1e6
    t = time.perf_counter(); oracle.lqr_from_x0_batch(0, 100, 0.02, x0, qr, want_P0=False); cpu_us = (time.perf_counter() - t) * 1e6
    out.append({"batch": B, "gpu_call_us_median": round(gpu_us, 1), "gpu_call_us_p90": round(p90_us, 1), "cpu_oracle_us_all_threads": round(cpu_us, 1)})
print(json.dumps(out))
