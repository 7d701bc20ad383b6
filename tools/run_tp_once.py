"""One time-parallel from-x0 call (batch 64, N=100) — the target of an ncu capture of k_ric12_tp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from noc_b200 import capi, problems as pb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
s = capi.Solver(0); qr = pb.pack_qrqf(*pb.default_weights())
x0 = pb.sample_x0(B, seed=3); K0 = np.empty((B, 4, 12))
p = capi.problem(0, N, flags=capi.FLAG_TIME_PARALLEL)
for _ in range(3):
    s.lqr_solve_from_x0(p, B, x0, qr, K0=K0)
s.close()
