#!/usr/bin/env python
"""BASELINE.json configs[3]: 1,048,576 receding-horizon LQR solves (N=50, K0 / u0 / cost only) sharded across the
GPUs of one box (strong scaling: the total is fixed).  Launch: python tools/bench_mpc_sharded.py (1 GPU) or
python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/bench_mpc_sharded.py
Each rank solves its contiguous slice with its own noc_ctx (no data-path collective); costs are all-gathered over
NCCL and the best rollout picked (noc_argmin_cost).  Device-resident x0, CUDA-event timing, max over ranks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from noc_b200 import capi, problems as pb, sharding

TOTAL, N = 1 << 20, 50


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    b0, b1 = sharding.shard_slice(TOTAL, rank, world)
    B = b1 - b0
    x0 = pb.sample_x0(TOTAL, seed=4)[b0:b1]
    qr = pb.pack_qrqf(*pb.default_weights())
    s = capi.Solver(local)
    stream = torch.cuda.Stream(device=dev); s.set_stream(stream.cuda_stream)
    f64 = dict(dtype=torch.float64, device=dev)
    d_x0 = torch.from_numpy(np.ascontiguousarray(x0)).to(dev); d_qr = torch.from_numpy(qr).to(dev)
    K0 = torch.empty(B, 4, 12, **f64); cost = torch.empty(B, **f64); st = torch.empty(B, dtype=torch.int32, device=dev)
    p = capi.problem(0, N, flags=capi.FLAG_ASYNC)
    best = {}

    def step():
        s.lqr_solve_from_x0(p, B, d_x0, d_qr, K0=K0, cost=cost, status=st)
        with torch.cuda.stream(stream):
            full = sharding.gather_costs(cost, TOTAL, dist)
        best["i"], best["c"] = s.argmin_cost(full, TOTAL)

    for _ in range(3):
        step()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if dist:
        t = torch.tensor([ms], **f64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    bad = int((st != 0).sum().item())
    if rank == 0:
        print(json.dumps({"config": 4, "what": "1,048,576 MPC warm-start LQR solves N=50 (K0+cost), sharded, cost all-gather + argmin",
                          "n_gpus": world, "ms": ms, "solves_per_s": TOTAL / ms * 1e3, "best": best, "status_bad_rank0": bad}))
    s.close()
    if dist:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
