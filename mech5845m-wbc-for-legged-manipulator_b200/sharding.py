"""Multi-GPU partition of a batch of robot states.

States are independent, so N states are split into contiguous shards, one per rank (one process per GPU); the tree
table and the configuration are replicated, and nothing is exchanged while a tick runs.  The single collective is an
all-gather of the solutions and solver reports AFTER the timed region, used for the cross-rank correctness check
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference has no distributed code at all.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_info():
    """(rank, world, local_rank) from the torchrun environment (1 process when not launched by torchrun)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_states, rank, world):
    """Contiguous shard [lo, hi) of `n_states` for `rank`: the first n % world ranks hold one extra state."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(int(n_states), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_states, world):
    return [shard_range(n_states, r, world)[1] - shard_range(n_states, r, world)[0] for r in range(world)]


def gather_states(local, n_states, group=None):
    """All-gather per-rank results [n_local, ...] into the global [n_states, ...] tensor (ragged shards allowed)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_states, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("local shard does not match shard_range()")
    local = local.contiguous()
    if len(set(sizes)) == 1:
        out = torch.empty((n_states,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def max_over_ranks(value, device, group=None):
    """Device-timed durations are reported as the maximum over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
