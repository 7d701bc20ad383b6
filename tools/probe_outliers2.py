#!/usr/bin/env python
"""Are the early-process stalls environmental?  (1) a pure-torch kernel+sync loop first thing in a fresh process;
(2) our async call split into enqueue / synchronize wall time with the library's per-kernel GPU time beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
T0 = time.perf_counter()
dev = torch.device("cuda", 0)
a = torch.zeros(64 << 20, dtype=torch.float64, device=dev)   # 512 MB: ~0.2 ms per pass
torch.cuda.synchronize()
ts = []; at = []
t_end = time.perf_counter() + 2.5
while time.perf_counter() < t_end:
    t = time.perf_counter(); a.add_(1.0); a.add_(1.0); a.add_(1.0); a.add_(1.0); torch.cuda.synchronize(); ts.append(time.perf_counter() - t); at.append(t - T0)
ts = np.array(ts) * 1e3
idx = np.where(ts > 3 * np.median(ts))[0]
print(json.dumps({"case": "torch_only_first_2.5s", "n": len(ts), "median_ms": float(np.median(ts)), "stalls": [(round(at[i], 3), round(float(ts[i]), 2)) for i in idx][:40]}), flush=True)

from noc_b200 import capi, problems as pb
s = capi.Solver(0)
stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = pb.pack_qrqf(*pb.default_weights()); d_qr = torch.from_numpy(qr).to(dev)
B, N = 65536, 100
d_x0 = torch.from_numpy(pb.sample_x0(B)).to(dev)
d_K0 = torch.zeros(B, 4, 12, **f64); d_c = torch.zeros(B, **f64); d_s = torch.zeros(B, dtype=torch.int32, device=dev)
p = capi.problem(0, N, flags=capi.FLAG_ASYNC)
rows = []
s.profile_enable(True)
for i in range(400):
    s.profile_reset()
    t0 = time.perf_counter(); s.lqr_solve_from_x0(p, B, d_x0, d_qr, K0=d_K0, cost=d_c, status=d_s)
    t1 = time.perf_counter(); stream.synchronize(); t2 = time.perf_counter()
    g = s.profile_read(capi.KERNEL_RICCATI)[0] + s.profile_read(capi.KERNEL_ROLLOUT)[0]
    rows.append((t0 - T0, 1e3 * (t1 - t0), 1e3 * (t2 - t1), g))
r = np.array(rows)
med = np.median(r[:, 1] + r[:, 2])
bad = r[(r[:, 1] + r[:, 2]) > 2 * med]
print(json.dumps({"case": "noc_async_split", "median_total_ms": float(med), "n_stalls": int(len(bad)),
                  "stalls_t_enq_sync_gpukernel_ms": [[round(float(v), 2) for v in row] for row in bad[:30]]}), flush=True)
s.close()
