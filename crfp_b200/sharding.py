"""Clip sharding across the GPUs of one box (BASELINE.json configs[2]).

The recurrence is sequential in time but clips never interact (every op of the forward is per-sample), so the unit
of multi-GPU work is the clip: rank r owns a contiguous block of clip indices, runs the whole forward on them with no
collective on the data path, and the only communication is the optional result gather.  One process per GPU
(`torchrun`), `torch.distributed` for the plumbing: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def clip_range(n_clips: int, world_size: int, rank: int):
    """Contiguous balanced partition: the first `n_clips % world_size` ranks get one extra clip."""
    if not (0 <= rank < world_size) or n_clips < 0:
        raise ValueError("bad rank / world_size / n_clips")
    base, extra = divmod(n_clips, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_clips(tensors, world_size: int, rank: int):
    """Slice the leading (clip) dimension of every tensor for this rank."""
    n = tensors[0].shape[0]
    lo, hi = clip_range(n, world_size, rank)
    return [t[lo:hi] for t in tensors]


def gather_clips(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """All-gather per-rank result slabs (possibly of different clip counts) back into clip order on every rank."""
    world = dist.get_world_size(group)
    counts = [clip_range(n_clips, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in counts)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, counts)], dim=0)


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-timed numbers are reported as the max over ranks (bench contract)."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
