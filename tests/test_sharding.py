"""CPU tests (-m "not gpu"): the N>1 clip-sharding plumbing with world_size 2 over gloo."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crfp_b200.sharding import clip_range, gather_clips, max_over_ranks, shard_clips


def test_clip_range_partitions_exactly():
    for n in (0, 1, 5, 8, 64, 67):
        for world in (1, 2, 4, 8):
            spans = [clip_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_clips):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_clips * 6, dtype=torch.float32).view(n_clips, 2, 3)
        (mine,) = shard_clips([full], world, rank)
        lo, hi = clip_range(n_clips, world, rank)
        assert torch.equal(mine, full[lo:hi])
        result = mine * 2.0 + 1.0                      # stand-in for the per-clip forward (no collective inside)
        gathered = gather_clips(result, n_clips)
        assert torch.equal(gathered, full * 2.0 + 1.0)
        assert max_over_ranks(float(rank + 1), torch.device("cpu")) == float(world)
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2_gloo():
    mp.spawn(_worker, args=(2, 29533, 5), nprocs=2, join=True)
