#!/usr/bin/env python
"""A/B of the headline sweep (config 2: n=12, N=100, records in HBM, every K_k written): k_ric12_mma (FP64 tensor cores)
vs k_ric12 (DFMA chains, NOC_OPT_NO_TENSOR_SWEEP), same inputs, outputs compared bit for bit.
usage: python tools/ab_sweep.py [batch=65536] [N=100] [reps=10]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from noc_b200 import capi, problems as pb

def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda:0")
    s = capi.Solver(0)
    Q, R, Qf = pb.default_weights()
    qr = torch.from_numpy(pb.pack_qrqf(Q, R, Qf)).to(dev)
    x0 = torch.from_numpy(pb.sample_x0(batch, seed=7)).to(dev)
    prob = capi.problem(capi.MODEL_QUAD12, N)
    xbar = torch.empty(batch, N + 1, 12, dtype=torch.float64, device=dev)
    ubar = torch.from_numpy(np.tile(pb.hover_input(), (batch, N, 1))).to(dev)
    AB = torch.empty(batch, N, 192, dtype=torch.float64, device=dev)
    s.rollout_open_loop(prob, batch, x0, ubar, xbar)
    s.linearize_batch(prob, batch, xbar, ubar, AB)
    out = {}
    res = {}
    for name, flag, variant in (("k_ric12_mma", 0, 0), ("k_ric12_mma_interlock", 0, 1), ("k_ric12_mma_direct", 0, 2), ("k_ric12", 1, 0)):
        s.set_option(capi.OPT_NO_TENSOR_SWEEP, flag)
        s.set_option(capi.OPT_TENSOR_SWEEP_VARIANT, variant)
        K = torch.empty(batch, N, 4, 12, dtype=torch.float64, device=dev)
        P0 = torch.empty(batch, 12, 12, dtype=torch.float64, device=dev)
        cost = torch.empty(batch, dtype=torch.float64, device=dev)
        status = torch.empty(batch, dtype=torch.int32, device=dev)
        lp = capi.problem(capi.MODEL_LTV_USER, N)
        call = lambda: s.lqr_solve_batch(lp, batch, AB, qr, x0dev=x0, K=K, P0=P0, cost=cost, status=status)
        for _ in range(3): call()
        s.synchronize(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream()
        s.set_stream(st.cuda_stream)
        call(); torch.cuda.synchronize()
        e0.record(st)
        for _ in range(reps): call()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"ms": ms, "solves_per_s": batch / ms * 1e3, "hbm_frac": batch * N * 1920 / (ms * 1e-3) / 1e9 / 6550.4}
        res[name] = (K, P0, cost, status)
    b = res["k_ric12"]
    out["bit_identical"] = {n: bool(all(torch.equal(x.view(torch.int64) if x.dtype == torch.float64 else x,
                                                     y.view(torch.int64) if y.dtype == torch.float64 else y) for x, y in zip(a, b)))
                            for n, a in res.items() if n != "k_ric12"}
    a = res["k_ric12_mma"]
    out["status_ok"] = bool((a[3] == 0).all().item())
    out["batch"], out["N"] = batch, N
    print(json.dumps(out))

if __name__ == "__main__":
    main()
