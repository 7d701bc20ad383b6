"""A few one-iteration SLQ solves of the dual-UAV model (n=24) — the target of an ncu capture of k_ric24_mma<RIC_AFFINE>.
usage: python tools/run_ric24_once.py [batch] [N] [box]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noc_b200 import capi, problems as pb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
box = dict(u_min=0.0, u_max=4.0) if len(sys.argv) > 3 and sys.argv[3] == "box" else {}
s = capi.Solver(0); qr = pb.pack_qrqf(*pb.default_weights(1))
xs = pb.hover_state(B, model=1); xg = pb.sample_goals(B, seed=35, model=1)
p = capi.problem(1, N, max_iter=1, **box)
cost = np.empty(B); it = np.empty(B, dtype=np.int32); st = np.empty(B, dtype=np.int32)
for _ in range(3):
    s.slq_solve_batch(p, B, xs, xg, qr, cost=cost, iters=it, status=st)
s.close()
