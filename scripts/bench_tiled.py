"""BASELINE.json configs[3]: ONE long clip spatially tiled with neighbour halo exchange over N GPUs
(default LR 270x480 -> 2160x3840; --h 135 --w 240 is the 1080p-output clip).
Launch: python scripts/bench_tiled.py            (1 GPU: all tiles back to back, plus the untiled forward for reference)
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
               scripts/bench_tiled.py [--grid 2x4] [--halo auto] [--frames 20]
Prints one JSON line (rank 0): frames/s of the whole clip, device-timed, max over ranks; with N > 1 rank 0 also runs the
untiled forward and the single-process tiled forward on its own GPU and reports
  tiled_vs_untiled_max_abs   (<= 1e-4: fp32 rounding of tile-local sampling coordinates)
  multi_rank_bitmatch        (the N-rank result equals the single-process tiled result bit for bit).
Not the headline bench (bench.py is); this is the measurement DESIGN.md quotes for the tiling row."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import CRFP_DSV                                  # noqa: E402
from crfp_b200.synthetic import make_clip, make_state_dict      # noqa: E402
from crfp_b200.tiling import TiledClipRunner                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="2x4")
    ap.add_argument("--halo", default="auto")
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--h", type=int, default=270)
    ap.add_argument("--w", type=int, default=480)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--gather", action="store_true", help="gather the output frames on every rank at the end of the clip "
                                                          "(default: they stay sharded by tile on their GPUs)")
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gy, gx = (int(v) for v in a.grid.split("x"))
    halo = a.halo if a.halo == "auto" else int(a.halo)
    model = CRFP_DSV("cuda", mid_channels=32).eval()
    model.load_state_dict(make_state_dict(seed=1), strict=True)
    model = model.cuda()
    lrs, fvs, mks, _ = make_clip(seed=3, n=1, t=a.frames, h=a.h, w=a.w, fv_size=256)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    runner = TiledClipRunner(model, grid=(gy, gx), halo=halo, gather_output=a.gather)

    def timed(fn, steps, warmup, sync_ranks=True):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1 and sync_ranks:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
        if world > 1 and sync_ranks:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    ms_t, out_t = timed(lambda: runner(lrs, fvs, mks), a.steps, a.warmup)
    line = {"metric": "output frames/sec, one clip spatially tiled", "unit": "frames/s", "n_gpus": world,
            "value": a.frames / ms_t * 1e3, "ms_per_clip": ms_t, "ms_per_frame": ms_t / a.frames,
            "config": {"workload": f"{a.h}x{a.w} LR -> {8 * a.h}x{8 * a.w}, t={a.frames}, grid {a.grid}",
                       "gather_output": a.gather},
            "tiling": dict(runner.last)}
    if not a.no_check:
        # gather the sharded result for the comparisons (outside the timed region)
        full = out_t.clone()
        if world > 1 and not a.gather:
            dist.all_reduce(full, op=dist.ReduceOp.SUM)       # interiors are disjoint, everything else is zero
        if rank == 0:
            ms_u, out_u = timed(lambda: model(lrs, fvs, mks), a.steps, a.warmup, sync_ranks=False)
            line["untiled_1gpu_fps"] = a.frames / ms_u * 1e3
            line["tiled_vs_untiled_max_abs"] = (full - out_u).abs().max().item()
            line["speedup_vs_untiled_1gpu"] = line["value"] / line["untiled_1gpu_fps"]
            if world > 1:
                single = TiledClipRunner(model, grid=(gy, gx), halo=runner.last["halo"], distributed=False)
                ms_s, out_s = timed(lambda: single(lrs, fvs, mks), 1, 1, sync_ranks=False)
                line["tiled_1gpu_fps"] = a.frames / ms_s * 1e3
                line["multi_rank_bitmatch"] = bool(torch.equal(full, out_s))
        if world > 1:
            dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
