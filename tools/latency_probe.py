"""Single-call latency of the C-ABI for small batches (host buffers), sequential vs time-parallel sweep, next to the
CPU oracle; and the device time of the sweep kernels alone (library event pairs)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noc_b200 import capi, problems as pb
from oracle import pyoracle as oracle
qr = pb.pack_qrqf(*pb.default_weights())
s = capi.Solver(0)
out = []
for N in (50, 100, 200):
    for B in (1, 8, 64, 512, 2048):
        x0 = pb.sample_x0(B, seed=3)
        K0 = np.empty((B, 4, 12)); cost = np.empty(B); st = np.empty(B, dtype=np.int32)
        row = {"N": N, "batch": B}
        for name, flags in (("sequential", 0), ("time_parallel", capi.FLAG_TIME_PARALLEL)):
            p = capi.problem(0, N, flags=flags)
            for _ in range(5): s.lqr_solve_from_x0(p, B, x0, qr, K0=K0, cost=cost, status=st)
            ts = []
            for _ in range(200):
                t = time.perf_counter(); s.lqr_solve_from_x0(p, B, x0, qr, K0=K0, cost=cost, status=st); ts.append(time.perf_counter() - t)
            s.profile_reset(); s.profile_enable(True)
            for _ in range(20): s.lqr_solve_from_x0(p, B, x0, qr, K0=K0, cost=cost, status=st)
            s.profile_enable(False)
            row[name] = {"call_us_median": round(float(np.median(ts)) * 1e6, 1), "call_us_p90": round(float(np.percentile(ts, 90)) * 1e6, 1),
                         "sweep_kernel_us": round(s.profile_read(capi.KERNEL_RICCATI)[0] / 20 * 1e3, 1),
                         "rollout_us": round(s.profile_read(capi.KERNEL_ROLLOUT)[0] / 20 * 1e3, 1),
                         "linearize_us": round(s.profile_read(capi.KERNEL_LINEARIZE)[0] / 20 * 1e3, 1)}
        t = time.perf_counter(); oracle.lqr_from_x0_batch(0, N, 0.02, x0, qr, want_P0=False); row["cpu_oracle_us_all_threads"] = round((time.perf_counter() - t) * 1e6, 1)
        out.append(row)
        print(json.dumps(row), flush=True)
