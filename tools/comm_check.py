#!/usr/bin/env python
"""Multi-rank check of the in-library multi-GPU plane (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/comm_check.py
Every rank solves its slice through noc_comm_lqr_solve_from_x0 (costs and K0 of ALL ranks gathered device-resident over
NVLink by the library's own NCCL communicator, best rollout picked on the device) and compares the gathered arrays
and the winner with a single-GPU solve of the whole batch.  torch.distributed only distributes the 128-byte NCCL id."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from noc_b200 import capi, problems as pb
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    s = capi.Solver(local)
    comm = capi.Comm.from_torch_distributed(s, dist, dev)
    cnt, N = 20000, 10
    qr = pb.pack_qrqf(*pb.default_weights())
    x0 = pb.sample_x0(world * cnt, seed=5)
    d_qr = torch.from_numpy(qr).to(dev)
    d_x0 = torch.from_numpy(x0[rank * cnt:(rank + 1) * cnt]).to(dev)
    cost_all = torch.empty(world * cnt, dtype=torch.float64, device=dev)
    K0_all = torch.empty(world * cnt, 48, dtype=torch.float64, device=dev)
    st = torch.empty(cnt, dtype=torch.int32, device=dev)
    best = torch.zeros(2, dtype=torch.float64, device=dev)
    idx, val = comm.lqr_solve_from_x0(capi.problem(0, N), cnt, d_x0, d_qr, best, cost_all=cost_all, K0_all=K0_all,
                                      status=st, want_host=True)
    # reference: the whole batch on this GPU alone
    full_x0 = torch.from_numpy(x0).to(dev)
    rc = torch.empty(world * cnt, dtype=torch.float64, device=dev)
    rK = torch.empty(world * cnt, 48, dtype=torch.float64, device=dev)
    s.lqr_solve_from_x0(capi.problem(0, N), world * cnt, full_x0, d_qr, K0=rK, cost=rc)
    ok = torch.equal(rc, cost_all) and torch.equal(rK, K0_all) and int((st != 0).sum()) == 0
    ri = int(torch.argmin(rc).item())
    ok = ok and idx == ri and val == float(rc[ri].item())
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"COMM CHECK {'OK' if int(t.item()) else 'FAILED'}: {world} ranks x {cnt} instances, best = ({idx}, {val})", flush=True)
    comm.close(); s.close()
    dist.destroy_process_group()
    return 0 if int(t.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
