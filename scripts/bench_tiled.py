"""BASELINE.json configs[3]: ONE 270x480-LR clip (8x -> 2160x3840) spatially tiled with halo exchange over N GPUs.
Launch: python scripts/bench_tiled.py            (1 GPU: all tiles back to back, plus the untiled forward for reference)
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
               scripts/bench_tiled.py [--grid 2x4] [--halo 32] [--frames 20]
Prints one JSON line (rank 0): frames/s of the whole clip, device-timed, max over ranks.  Not the headline bench
(bench.py is); this is the measurement DESIGN.md quotes for the tiling row."""
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
    ap.add_argument("--halo", type=int, default=32)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--h", type=int, default=270)
    ap.add_argument("--w", type=int, default=480)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-gather", action="store_true", help="leave the output frames sharded by tile on their ranks")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gy, gx = (int(v) for v in a.grid.split("x"))
    model = CRFP_DSV("cuda", mid_channels=32).eval()
    model.load_state_dict(make_state_dict(seed=1), strict=True)
    model = model.cuda()
    lrs, fvs, mks, _ = make_clip(seed=3, n=1, t=a.frames, h=a.h, w=a.w, fv_size=256)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    runner = TiledClipRunner(model, grid=(gy, gx), halo=a.halo, gather_output=not a.no_gather)

    def timed(fn):
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), out

    ms_t, out_t = timed(lambda: runner(lrs, fvs, mks))
    line = {"metric": "output frames/sec, one clip spatially tiled", "unit": "frames/s", "n_gpus": world,
            "value": a.frames / ms_t * 1e3, "ms_per_clip": ms_t,
            "config": {"workload": f"{a.h}x{a.w} LR -> {8 * a.h}x{8 * a.w}, t={a.frames}, grid {a.grid}, halo {a.halo}",
                       "gather_output": not a.no_gather}}
    if world == 1:
        ms_u, out_u = timed(lambda: model(lrs, fvs, mks))
        line["untiled_fps"] = a.frames / ms_u * 1e3
        line["max_abs_tiled_vs_untiled"] = (out_t - out_u).abs().max().item()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
