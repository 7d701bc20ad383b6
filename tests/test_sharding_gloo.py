"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): slices, ragged gather, best-rollout pick.
The per-rank "device" is the CPU oracle (tests may use it); the GPU version of the same flow is bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from noc_b200 import problems as pb
from noc_b200 import sharding


def test_shard_slices_cover_batch():
    for batch in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            sl = [sharding.shard_slice(batch, g, world) for g in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == batch
            assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in sl]
            assert max(sizes) - min(sizes) <= 1


def test_best_rollout_skips_nan():
    c = np.array([3.0, np.nan, 1.5, 1.5, np.inf])
    assert sharding.best_rollout(c) == (2, 1.5)
    assert sharding.best_rollout(np.array([np.nan, np.nan]))[0] == -1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, N, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as oracle
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    x0 = pb.sample_x0(batch, seed=99)

    def solve_slice(b, e):
        return oracle.lqr_from_x0_batch(0, N, 0.02, x0[b:e], qr, threads=1)["cost"] if e > b else np.empty(0)

    full, idx, val = sharding.solve_sharded(batch, solve_slice, dist)
    if rank == 0:
        q.put((full.numpy().copy(), idx, val))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 37), (3, 20)])
def test_sharded_solve_matches_single_rank(oracle, world, batch):
    N = 12
    Q, R, Qf = pb.default_weights()
    qr = pb.pack_qrqf(Q, R, Qf)
    ref = oracle.lqr_from_x0_batch(0, N, 0.02, pb.sample_x0(batch, seed=99), qr)["cost"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, idx, val = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(full, ref)                       # sharded == single-rank, bitwise, in instance order
    assert idx == int(np.argmin(ref)) and val == ref.min()
