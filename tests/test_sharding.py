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


def test_tile_plan_partitions_frame():
    from crfp_b200.tiling import tile_plan
    for h, w, gy, gx, halo in [(180, 320, 2, 4, 32), (64, 96, 2, 2, 24), (45, 80, 3, 3, 0), (16, 16, 1, 1, 8)]:
        plan = tile_plan(h, w, gy, gx, halo)
        cover = torch.zeros(h, w, dtype=torch.int32)
        for (y0, y1, x0, x1), (ey0, ey1, ex0, ex1) in plan:
            cover[y0:y1, x0:x1] += 1
            assert 0 <= ey0 <= y0 < y1 <= ey1 <= h and 0 <= ex0 <= x0 < x1 <= ex1 <= w
            assert y0 - ey0 == min(halo, y0) and ey1 - y1 == min(halo, h - y1)
            assert x0 - ex0 == min(halo, x0) and ex1 - x1 == min(halo, w - x1)
        assert torch.all(cover == 1)


def _tile_worker(rank, world, port):
    """The per-frame exchange step of the tiled runner: neighbour strips only.  Every rank starts with the truth inside
    the INTERIORS of its own tiles and stale halos; after the exchange every extended tile equals the truth crop."""
    from crfp_b200.tiling import exchange_halos, halo_pairs, tile_plan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = 12, 20
        plan = tile_plan(h, w, 2, 3, 4)
        pairs = halo_pairs(plan)
        truth_hr = torch.arange(8 * h * 8 * w * 4, dtype=torch.float32).view(1, 8 * h, 8 * w, 4)
        truth_l1 = torch.arange(2 * h * 2 * w * 24, dtype=torch.float32).view(1, 2 * h, 2 * w, 24) * 0.5
        states = {}
        for k, ((y0, y1, x0, x1), (ey0, ey1, ex0, ex1)) in enumerate(plan):
            if k % world != rank:
                continue
            hr = torch.full((1, 8 * (ey1 - ey0), 8 * (ex1 - ex0), 4), -1.0)             # stale everywhere ...
            l1 = torch.full((1, 2 * (ey1 - ey0), 2 * (ex1 - ex0), 24), -1.0)
            hr[:, 8 * (y0 - ey0):8 * (y1 - ey0), 8 * (x0 - ex0):8 * (x1 - ex0)] = truth_hr[:, 8 * y0:8 * y1, 8 * x0:8 * x1]
            l1[:, 2 * (y0 - ey0):2 * (y1 - ey0), 2 * (x0 - ex0):2 * (x1 - ex0)] = truth_l1[:, 2 * y0:2 * y1, 2 * x0:2 * x1]
            states[k] = (hr, l1)                                                           # ... except my own interiors
        got = exchange_halos(states, plan, pairs, world, rank)
        for k, (hr, l1) in states.items():
            (_, (ey0, ey1, ex0, ex1)) = plan[k]
            assert torch.equal(hr, truth_hr[:, 8 * ey0:8 * ey1, 8 * ex0:8 * ex1]), k
            assert torch.equal(l1, truth_l1[:, 2 * ey0:2 * ey1, 2 * ex0:2 * ex1]), k
        # only neighbour strips crossed the wire: far less than the all-gather of every interior
        everything = (truth_hr.numel() + truth_l1.numel()) * 4
        assert 0 < got < everything
    finally:
        dist.destroy_process_group()


def test_halo_pairs_cover_exactly_the_halos():
    from crfp_b200.tiling import halo_pairs, tile_plan
    for h, w, gy, gx, halo in [(24, 40, 2, 4, 5), (30, 30, 3, 3, 8), (16, 48, 1, 3, 4)]:
        plan = tile_plan(h, w, gy, gx, halo)
        pairs = halo_pairs(plan)
        for j, ((y0, y1, x0, x1), (ey0, ey1, ex0, ex1)) in enumerate(plan):
            cover = torch.zeros(h, w, dtype=torch.int32)
            cover[y0:y1, x0:x1] += 1
            for (k, jj, (a0, a1, b0, b1)) in pairs:
                if jj == j:
                    assert k != j
                    cover[a0:a1, b0:b1] += 1
            assert torch.all(cover[ey0:ey1, ex0:ex1] == 1)          # halo fully covered, exactly once
            cover[ey0:ey1, ex0:ex1] = 0
            assert torch.all(cover == 0)                            # and nothing outside the extended tile


def test_tile_state_exchange_world2_gloo():
    mp.spawn(_tile_worker, args=(2, 29537), nprocs=2, join=True)
