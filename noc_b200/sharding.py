"""Batch sharding across the GPUs of one box.

Instances are independent, so there is no data-path collective: rank g of G owns the contiguous slice
[g*B/G, (g+1)*B/G) and solves it with its own noc_ctx.  The only exchange is the gather of per-instance costs
(8 B each) — NCCL over NVLink on GPUs, gloo in the CPU tests — followed by "pick the best rollout".
`solve_slice` is whatever runs one rank's slice (capi.Solver on a GPU; tests plug the CPU oracle in).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_slice(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of rank `rank`; the first batch % world ranks get one extra instance.
    Same rule as lqr_control::shardSlice (include/lqr_control/batched_lqr.hpp)."""
    base, rem = divmod(batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_costs(local_cost, batch: int, dist=None):
    """All-gather the per-rank cost slices into the full [batch] vector (ragged slices are padded with +inf).
    `local_cost` is a torch tensor on the rank's device (CUDA for nccl, CPU for gloo)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_cost
    world, rank = dist.get_world_size(), dist.get_rank()
    width = -(-batch // world)
    pad = torch.full((width,), float("inf"), dtype=local_cost.dtype, device=local_cost.device)
    pad[: local_cost.numel()] = local_cost
    out = torch.empty(world * width, dtype=local_cost.dtype, device=local_cost.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for g in range(world):
        b, e = shard_slice(batch, g, world)
        parts.append(out[g * width: g * width + (e - b)])
    return torch.cat(parts)


def best_rollout(full_cost) -> Tuple[int, float]:
    """Index and value of the smallest finite cost (NaN = failed instance); ties -> smallest index.
    CPU restatement of k_argmin_cost, used to cross-check it."""
    c = np.asarray(full_cost.detach().cpu().numpy() if hasattr(full_cost, "detach") else full_cost, dtype=np.float64)
    c = np.where(np.isnan(c), np.inf, c)
    i = int(np.argmin(c))
    return (i, float(c[i])) if np.isfinite(c[i]) else (-1, float("nan"))


def solve_sharded(batch: int, solve_slice: Callable[[int, int], "np.ndarray"], dist=None, device="cpu"):
    """Run `solve_slice(begin, end) -> cost[end-begin]` on this rank's slice, gather all costs, pick the best.
    Returns (full_cost tensor, best_index, best_cost)."""
    import torch
    world = dist.get_world_size() if (dist is not None and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    b, e = shard_slice(batch, rank, world)
    local = solve_slice(b, e)
    local_t = local if hasattr(local, "detach") else torch.from_numpy(np.ascontiguousarray(local)).to(device)
    full = gather_costs(local_t, batch, dist)
    idx, val = best_rollout(full)
    return full, idx, val
