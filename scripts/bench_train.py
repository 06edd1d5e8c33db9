"""Training-step benchmark (BASELINE.json configs[4]): forward + backward + fused gradient all-reduce + Adam.

  python scripts/bench_train.py [--shape v7|crop] [--steps K] [--warmup W]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         scripts/bench_train.py ...          (one rank per GPU, NCCL all-reduce of the 9.14 MB gradient bucket)

Shapes: v7 = one Vimeo-7 clip per GPU (t=7, LR 64x112 -> 512x896); crop = the reference's training crop per GPU
(n=8, t=15, LR 32x32 -> 256x256, FV 128: /root/reference/train.sh:18-21).  Synthetic data, random-init weights.
Timed with CUDA events over K steps after W warm-ups, max over ranks; prints one JSON line on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="v7", choices=["v7", "crop", "tiny"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graphs", action="store_true", help="replay forward+backward from one captured CUDA graph per step")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl")
    from crfp_b200 import CRFP_DSV, _lib
    from crfp_b200.synthetic import make_clip, make_state_dict
    from crfp_b200.trainer import Trainer

    n, t, h, w, fv = {"v7": (1, 7, 64, 112, 128), "crop": (8, 15, 32, 32, 128), "tiny": (1, 3, 16, 16, 48)}[args.shape]
    model = CRFP_DSV("cuda", mid_channels=32)
    model.load_state_dict(make_state_dict(seed=1), strict=True)
    model.cuda()
    tr = Trainer(model, freeze_flow_iters=0, use_graphs=args.graphs)   # FNet trains from step 0: the full backward is timed
    if args.graphs:
        args.warmup = max(args.warmup, 4)             # two eager steps, the capture, one replay
    lrs, fvs, mks, _ = make_clip(seed=2 + rank, n=n, t=t, h=h, w=w, fv_size=fv)
    hr = torch.rand(n, t, 3, 8 * h, 8 * w, generator=torch.Generator().manual_seed(3 + rank))
    batch = (lrs.cuda(), fvs.cuda(), mks.cuda(), hr.cuda())
    losses = []
    for _ in range(args.warmup):
        losses.append(tr.step(*batch).item())
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tr.time_comm, tr.comm_events = world > 1, []
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(*batch)
    e1.record()
    torch.cuda.synchronize()
    comm_ms = [a.elapsed_time(b) for a, b in tr.comm_events]
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    losses.append(loss.item())
    if rank == 0:
        sec = ms.item() / 1e3
        print(json.dumps({
            "metric": "training frames/sec (forward + backward + all-reduce + Adam)",
            "value": world * n * t * args.steps / sec, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms.item() / args.steps, "higher_is_better": True, "scaling": "weak",
            "dtype": "fp32", "data": "synthetic", "cuda_graph": bool(args.graphs and tr.use_graphs and tr._graphs),
            # (the library's launch counter is per host thread and the backward runs on autograd's thread: the kernel
            # count of a step, 4 190 incl. copies at V7, comes from scripts/train_kernel_times.py instead)
            "config": {"workload": f"{args.shape}: n={n} clips/GPU, t={t}, LR {h}x{w} -> {8 * h}x{8 * w}, FV {fv}",
                       "params": int(tr.flat_p.numel()), "grad_bucket_bytes": int(tr.flat_g.numel() * 4)},
            "allreduce": ({"ms_per_step_mean": sum(comm_ms) / len(comm_ms), "ms_per_step_max": max(comm_ms), "backend": "nccl",
                           "what": "ONE all_reduce of the flat fp32 gradient bucket + the 1/world scale, CUDA events on rank 0 "
                                   "(includes waiting for the slowest rank's backward)"} if comm_ms else None),
            "loss_first_last": [losses[0], losses[-1]],
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
