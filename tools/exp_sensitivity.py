"""Which resource binds the sweep?  Times the LTV sweep with deliberately crippled builds (libnoc_b200_exp<N>.so,
compiled with -DNOC_EXP=N: results are WRONG, only the time is of interest):
  exp1: half of the P-row shared-memory loads (-9 % L1TEX wavefronts, same DFMAs)
  exp2: half of the phase-1 DFMAs (-22 % FP64 work, same loads)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from noc_b200 import capi, problems as pb
lib = sys.argv[1] if len(sys.argv) > 1 else None
if lib:
    capi.LIB_PATH = os.path.abspath(lib)
B, N = int(os.environ.get('EXP_B', 65536)), 100
dev = torch.device("cuda", 0)
s = capi.Solver(0)
stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
f64 = dict(dtype=torch.float64, device=dev)
qr = torch.from_numpy(pb.pack_qrqf(*pb.default_weights())).to(dev)
x0 = torch.from_numpy(pb.sample_x0(B, seed=9)).to(dev)
xbar = torch.empty(B, N + 1, 12, **f64); AB = torch.empty(B, N, 192, **f64)
K = torch.empty(B, N, 4, 12, **f64); st = torch.empty(B, dtype=torch.int32, device=dev)
pm = capi.problem(0, N, flags=capi.FLAG_ASYNC); pl = capi.problem(capi.MODEL_LTV_USER, N, flags=capi.FLAG_ASYNC)
s.rollout_open_loop(pm, B, x0, None, xbar); s.linearize_batch(pm, B, xbar, None, AB); s.synchronize()
fn = lambda: s.lqr_solve_batch(pl, B, AB, qr, K=K, status=st)
fn(); fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10): fn()
e1.record(stream); torch.cuda.synchronize()
print(json.dumps({"lib": lib or "product", "ms": e0.elapsed_time(e1) / 10}))
