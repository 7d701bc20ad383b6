"""Per-step timing of the e2e call (host buffers through the C-ABI) to see the spread."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from noc_b200 import capi, problems as pb
B, N = 65536, 100
Q, R, Qf = pb.default_weights(); qr = pb.pack_qrqf(Q, R, Qf)
x0 = pb.sample_x0(B, seed=1)
s = capi.Solver(0)
h_x0 = torch.from_numpy(x0).pin_memory()
h_K0 = torch.empty(B, 4, 12, dtype=torch.float64).pin_memory()
h_P0 = torch.empty(B, 12, 12, dtype=torch.float64).pin_memory()
h_cost = torch.empty(B, dtype=torch.float64).pin_memory()
h_status = torch.empty(B, dtype=torch.int32).pin_memory()
p = capi.problem(capi.MODEL_QUAD12, N)
ts = []
for i in range(int(os.environ.get('CALLS', 40))):
    t = time.perf_counter()
    s.lqr_solve_from_x0(p, B, h_x0, qr, K0=h_K0, P0=h_P0, cost=h_cost, status=h_status)
    ts.append((time.perf_counter() - t) * 1e3)
print("ms per call:", " ".join(f"{t:.2f}" for t in ts))
s.profile_reset(); s.profile_enable(True)
for i in range(5):
    s.lqr_solve_from_x0(p, B, h_x0, qr, K0=h_K0, P0=h_P0, cost=h_cost, status=h_status)
s.profile_enable(False)
print({k: s.profile_read(v)[0] / 5 for k, v in (("rollout", 1), ("riccati", 3), ("setup", 0))})

a = np.array(ts[2:])
print("stats ms: median %.3f mean %.3f p90 %.3f max %.3f  slow(>1.15x median) %d of %d" % (np.median(a), a.mean(), np.percentile(a, 90), a.max(), int((a > 1.15 * np.median(a)).sum()), a.size))
