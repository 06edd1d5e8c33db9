#!/usr/bin/env python
"""bench.py — output frames/s of the CRFP hot path (CRFP_DSV.forward) on B200, per the driver contract.

A "step" is one forward of one synthetic clip (per rank) through the drop-in module:
  workload "R-lit"  : LR 180x320 (REDS *_sharp_BI frames, BASELINE.json configs[1]) through the x8 network
                      -> 1440x2560, t frames (default 100)
  workload "R-nat"  : LR 90x160 -> 720x1280 (the x8 network's native route to 1280x720)
  workload "V7"     : LR 64x112, t=7 (configs[0], the reference's CPU-runnable case)
`value`  = frames/s with the clip already resident in HBM (CUDA events, max over ranks).
`e2e`    = frames/s through the public API from HOST buffers: pinned lrs + fovea patches + coords H2D, forward,
           output frames D2H, all inside the timed region.
`roofline` = the align kernel (DCNv2 @L1, the kernel BASELINE.json's metric names) timed alone through the C ABI.
`cpu_baseline` / `--impl reference` = the oracle port of the reference's PyTorch path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"R-lit": (180, 320, 100), "R-nat": (90, 160, 100), "V7": (64, 112, 7)}
METRIC = "output frames/sec"
ALIGN_BYTES_PER_L1_PX = 1120  # (32 in + 144 offset + 72 mask + 32 out) fp32, SURVEY.md 8(d)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crfp_b200", choices=["crfp_b200", "reference"])
    ap.add_argument("--workload", default="R-lit", choices=list(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override frames per clip")
    ap.add_argument("--clips", type=int, default=1, help="clips per GPU per step")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"],
                    help="tc: fp32 storage, tcgen05 3 x bf16 split contractions (fp32-grade, default); fp32: all-SIMT FFMA")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Samples before this point (warm-up) are dropped."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows[getattr(self, 'first', 0):]:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_fps(h, w, frames, threads):
    """The reference's PyTorch path (oracle port, bit-identical to the reference on CPU) on the host cores."""
    import torch
    from crfp_b200.synthetic import make_clip, make_state_dict
    from oracle import crfp_oracle as O
    torch.set_num_threads(threads)
    sd = make_state_dict(seed=1)
    lrs, fvs, mks, _ = make_clip(seed=2, n=1, t=frames, h=h, w=w, fv_size=96)
    t0 = time.perf_counter()
    out = O.crfp_dsv_forward(sd, lrs, fvs, mks)
    dt = time.perf_counter() - t0
    assert out.shape[1] == frames
    return frames / dt, dt


def run_reference(args):
    """`--impl reference`: rank 0 times the CPU path on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    h, w, _ = WORKLOADS[args.workload]
    frames = 2
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_fps(h, w, frames, cores)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_reference_fps(h, w, frames, cores)
        times.append(dt)
    total = sum(times)
    value = frames * args.steps / total
    sample = f"{frames}-frame clip of {args.workload} (LR {h}x{w} -> {8 * h}x{8 * w}) per step, oracle port, torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: LR {h}x{w} -> {8 * h}x{8 * w} (x8 network), {frames} frames/step",
                   "clips_per_gpu": 1},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _events_avg_ms(torch, launch, reps, warm=3):
    for i in range(warm):
        launch(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (a, b) in enumerate(evs):
        a.record()
        launch(i)
        b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    return sum(ms) / len(ms)


def time_align_kernel(torch, h1, w1, precision, reps=12):
    """Average device time of the DCNv2 @L1 align kernel alone (C ABI), 3 rotating input sets > L2 (cold launches)."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_dcn, pack_dcn_tc3
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    nbuf = 3  # 3 x (29 + 199) MB > 126 MB L2
    xs = [torch.randn(1, h1, w1, 32, generator=g).to(dev) for _ in range(nbuf)]
    oms, flows = [], []
    for _ in range(nbuf):
        fl = torch.randn(1, h1, w1, 2, generator=g) * 2.0
        om = torch.empty(1, h1, w1, 216)
        om[..., :144] = torch.tanh(torch.randn(1, h1, w1, 144, generator=g)) * 10.0 * 0.35 + fl.flip(-1).repeat(1, 1, 1, 72)
        om[..., 144:] = torch.rand(1, h1, w1, 72, generator=g)
        oms.append(om.to(dev)); flows.append(fl.to(dev))
    wt = (torch.randn(32, 32, 3, 3, generator=g) * 0.05).to(dev)
    out = torch.empty(1, h1, w1, 32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if precision == "tc":
        hi, lo, bp = pack_dcn_tc3(wt, torch.zeros(32, device=dev), 8)
        wptr = hi.data_ptr()
    else:
        wp, bp = pack_dcn(wt, torch.zeros(32, device=dev), 8)
        wptr = wp.data_ptr()

    def launch(i):
        k = i % nbuf
        d = L.DcnDesc(n=1, h=h1, w=w1, c=32, cout=32, dg=8, shared_taps=0, x=xs[k].data_ptr(), x_cstride=32,
                      x_coffset=0, offset=oms[k].data_ptr(), off_cstride=216, off_coffset=0,
                      mask=oms[k].data_ptr(), mask_cstride=216, mask_coffset=144, weight=wptr,
                      bias=bp.data_ptr(), out=out.data_ptr(), out_cstride=32, out_coffset=0)
        if precision == "tc":
            L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), flows[k].data_ptr(), st), "dcn_v2_tc3")
        else:
            L.check(L.lib().crfp_dcn_v2_fwd(C.byref(d), st), "dcn_v2")

    return _events_avg_ms(torch, launch, reps)


def time_conv_mix(torch, h1, w1, reps=4):
    """The per-frame mix of tensor-core conv launches (conv_tc3_ws_kernel, the dominant kernel of the step): every L1
    layer shape of one steady-state frame through crfp_conv3x3_tc3_fwd.  Returns (avg ms per launch, launches,
    algorithmic bytes per launch = unique fp32 operands in + out, useful fp32 FLOP per launch)."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_conv_tc3
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # (source channels, extra flow channels, cout, count per frame, output kind)
    layers = [([32, 32], 2, 32, 3, "nhwc"), ([32], 0, 32, 9, "nhwc"), ([32, 32], 0, 32, 5, "nhwc"),
              ([32], 0, 216, 3, "nhwc"), ([24], 0, 64, 1, "shuffle4"), ([32], 0, 64, 1, "shuffle4")]
    bufs = [torch.randn(1, h1, w1, 32, device=dev) for _ in range(5)]   # 5 x 29.5 MB planes, rotated: > L2 together
    buf24 = torch.randn(1, h1, w1, 24, device=dev)
    flow = torch.randn(1, h1, w1, 2, device=dev)
    out216 = torch.empty(1, h1, w1, 216, device=dev)
    out_hr = torch.empty(1, 4 * h1, 4 * w1, 4, device=dev)
    descs, keep = [], []
    px = h1 * w1
    tot_bytes = tot_flop = 0
    rot = 0
    for srcs, extra, cout, count, kind in layers:
        cin = sum(srcs) + extra
        w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
        hi, lo, bp, wx = pack_conv_tc3(w, torch.zeros(cout, device=dev), srcs, extra=extra)
        keep.append((hi, lo, bp, wx))
        for _ in range(count):
            d = L.ConvTc3Desc()
            d.n, d.h, d.w, d.nsrc = 1, h1, w1, len(srcs)
            for i, c in enumerate(srcs):
                t = buf24 if c == 24 else bufs[(rot + i) % 5]
                d.src[i] = L.TcSrc(ptr=t.data_ptr(), c=c, cstride=t.shape[-1], coffset=0)
            d.cout, d.act = cout, 1
            if cout == 216:   # the fused offset + mask heads: 10 * tanh + flow on 144 channels, sigmoid on 72
                d.act, d.flow, d.head_split, d.head_mag = L.ACT_DCN_HEAD, flow.data_ptr(), 144, 10.0
            d.weight_hi, d.weight_lo, d.bias = hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
            if extra:
                d.extra, d.w_extra = flow.data_ptr(), wx.data_ptr()
            d.post_scale, d.ndst = 1.0, 1
            if kind == "shuffle4":
                d.out_kind, d.shuffle_r = L.TC_OUT_SHUFFLE_F32, 4
                d.dst[0] = L.TcSrc(ptr=out_hr.data_ptr(), c=4, cstride=4, coffset=0)
            else:
                o = out216 if cout == 216 else bufs[(rot + 3) % 5]
                d.out_kind = L.TC_OUT_F32
                d.dst[0] = L.TcSrc(ptr=o.data_ptr(), c=cout, cstride=o.shape[-1], coffset=0)
            descs.append(d)
            rot += 1
            tot_bytes += (cin + cout) * 4 * px
            tot_flop += 2 * 9 * cin * cout * px

    def launch(_):
        for d in descs:
            L.check(L.lib().crfp_conv3x3_tc3_fwd(C.byref(d), st), "conv_tc3")

    ms = _events_avg_ms(torch, launch, reps, warm=2)
    return ms / len(descs), len(descs), tot_bytes / len(descs), tot_flop / len(descs)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from crfp_b200 import CRFP_DSV, _lib
    from crfp_b200.synthetic import make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    h, w, t = WORKLOADS[args.workload]
    if args.frames:
        t = args.frames
    n = args.clips
    H, W_ = 8 * h, 8 * w
    fv = 96

    model = CRFP_DSV("cuda", mid_channels=32, precision=args.precision).eval()
    model.load_state_dict(make_state_dict(seed=1), strict=True)
    model.to(dev)

    # synthetic clip: smooth-ish LR frames, Gaussian gaze, random fovea patch; seeded per rank
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    coarse = torch.rand(n, t, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
    lrs_h = torch.nn.functional.interpolate(coarse.view(n * t, 3, *coarse.shape[-2:]), size=(h, w), mode="bicubic",
                                            align_corners=False).view(n, t, 3, h, w)
    lrs_h = (lrs_h + 0.05 * torch.rand(n, t, 3, h, w, generator=g)).clamp_(0, 1).contiguous().pin_memory()
    patch_h = torch.rand(n, t, 3, fv, fv, generator=g).pin_memory()
    gy = (torch.randn(n, t, generator=g) * 50 + H / 2).floor().long() - fv // 2
    gx = (torch.randn(n, t, generator=g) * 50 + W_ / 2).floor().long() - fv // 2
    coords = torch.stack([gy.clamp(0, H - fv), gx.clamp(0, W_ - fv)], -1)

    def build_inputs():
        lrs = lrs_h.to(dev, non_blocking=True)
        patch = patch_h.to(dev, non_blocking=True)
        fvs = torch.zeros(n, t, 3, H, W_, device=dev)
        mks = torch.zeros(n, t, 1, H, W_, device=dev, dtype=torch.bool)
        for b in range(n):
            for i in range(t):
                y, x = int(coords[b, i, 0]), int(coords[b, i, 1])
                fvs[b, i, :, y:y + fv, x:x + fv] = patch[b, i]
                mks[b, i, :, y:y + fv, x:x + fv] = True
        return lrs, fvs, mks

    lrs, fvs, mks = build_inputs()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_decouple():
        """A device spin kernel ahead of the start event: the host enqueues the whole timed region (K graph launches)
        while the GPU is still spinning, so a host stall (this pool's VMs page memory in lazily: 50-150 ms hiccups
        were measured) cannot starve the GPU inside the timed region.  The spin ends before the start event fires."""
        torch.cuda._sleep(int(min(1.0, 0.2 + 0.01 * args.steps) * 1.9e9))

    # ---------------- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi -lms 100 starts sampling during the warm-up, keeps going through the timed region
    for _ in range(args.warmup):
        out = model(lrs, fvs, mks)
    barrier()
    _lib.lib().crfp_launch_count_reset()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_decouple()
    e0.record()
    for _ in range(args.steps):
        out = model(lrs, fvs, mks)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(_lib.lib().crfp_launch_count())
    clocks = sampler.stop() if rank == 0 else None
    assert torch.isfinite(out[:, -1]).all()
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    frames_total = world * n * t * args.steps
    value = frames_total / (ms_max * 1e-3)

    # ---------------- end-to-end through the public API from host buffers
    e2e = None
    if not args.no_e2e:
        out_h = torch.empty(n, t, 3, H, W_, dtype=torch.float32).pin_memory()
        lrs_d, patch_d = torch.empty_like(lrs_h, device=dev), torch.empty_like(patch_h, device=dev)

        def e2e_step():
            lrs_d.copy_(lrs_h, non_blocking=True)       # H2D from pinned memory, every step
            patch_d.copy_(patch_h, non_blocking=True)
            model.forward_patch(lrs_d, patch_d, coords, out_host=out_h)   # coords: host integers, copied inside
            # every frame is D2H-copied to pinned memory on a side stream while later frames compute

        del fvs, mks, out
        model._graphs.clear()
        for _ in range(max(2, min(args.warmup, 3))):
            e2e_step()
        barrier()
        host_decouple()          # as above: the K steps (H2D copies, graph launches, D2H copies) are enqueued during the spin
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        enq = (time.perf_counter() - t0) * 1e3
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": frames_total / (float(ems.item()) * 1e-3), "unit": "frames/s",
               "ms_per_step": float(ems.item()) / args.steps, "host_enqueue_ms_per_step": enq / args.steps,
               "h2d_bytes_per_step": int(lrs_h.numel() * 4 + patch_h.numel() * 4 + coords.numel() * 4),
               "d2h_bytes_per_step": int(out_h.numel() * 4),
               "api": "CRFP_DSV.forward_patch(lrs, fovea_patch, coords, out_host=pinned): H2D of lrs + patches + coords, "
                      "device-side fovea paste, forward, every frame D2H-copied on a side stream while the recurrence "
                      "continues; device-timed (events around the K steps, copies included), host enqueue time reported beside it"}
        del out_h

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the align kernel (timed alone, cold inputs)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    a_ms = time_align_kernel(torch, 2 * h, 2 * w, args.precision)
    alg_bytes = ALIGN_BYTES_PER_L1_PX * (2 * h) * (2 * w)
    a_ach = alg_bytes / (a_ms * 1e-3) / 1e9
    align = {"kernel": ("dcn_tc3_kernel" if args.precision == "tc" else "dcn_l1_kernel") + " (DCNv2 align @L1, C=32 dg=8)",
             "bound": "hbm", "achieved": a_ach, "peak": peak, "unit": "GB/s", "frac": a_ach / peak,
             "traffic": (247.1e6 if (args.workload == "R-lit" and args.precision == "tc") else None),
             "traffic_source": "dram__bytes_read+write per launch, ncu, profiles/r01/v3_dram_R-lit.csv (R-lit only)",
             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": a_ms,
             "timing": "kernel timed alone through the C ABI, CUDA events, 3 rotating input sets > L2"}
    if args.precision == "tc":
        c_ms, c_n, c_bytes, c_flop = time_conv_mix(torch, 2 * h, 2 * w)
        c_ach = c_bytes / (c_ms * 1e-3) / 1e9
        roofline = {"kernel": "conv_tc3_ws_kernel (tcgen05 3x3 implicit-GEMM conv, the per-frame mix of its %d L1 launches)" % c_n,
                    "bound": "hbm", "achieved": c_ach, "peak": peak, "unit": "GB/s", "frac": c_ach / peak,
                    "traffic": (64.3e6 if args.workload == "R-lit" else None),
                    "traffic_source": "dram__bytes_read+write, mean over the 23 L1 conv launches of one steady-state frame, "
                                      "ncu, profiles/r01/v3_dram_R-lit.csv (R-lit only; below the algorithmic bytes: L2 reuse)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": c_bytes, "avg_launch_ms": c_ms,
                    "tensor": {"useful_fp32_tflops": c_flop / (c_ms * 1e-3) / 1e12,
                               "issued_bf16_tflops": 3 * c_flop / (c_ms * 1e-3) / 1e12, "bf16_peak_tflops": bf16_peak,
                               "frac_of_bf16_peak": 3 * c_flop / (c_ms * 1e-3) / 1e12 / bf16_peak},
                    "timing": "all L1 conv launches of one steady-state frame replayed back to back through "
                              "crfp_conv3x3_tc3_fwd, CUDA events, rotating 29.5 MB planes + 199 MB heads output (> L2)",
                    "align_kernel": align}
    else:
        roofline = dict(align)
        roofline["peak_source"] = peak_src

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        fr = 2
        fps, dt = cpu_reference_fps(h, w, fr, cores)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{fr}-frame clip of {args.workload} (LR {h}x{w}), oracle port of the reference's PyTorch path, "
                         f"torch CPU fp32, {dt:.1f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: LR {h}x{w} -> {H}x{W_} (x8 network; BASELINE.json configs[1]), "
                               f"{t}-frame clip, fovea 96x96, CRFP_DSV mid_channels=32",
                   "precision": ("fp32 storage; dense contractions as 3 x bf16 split products on tcgen05 with fp32 TMEM "
                                 "accumulation (parity <= 1e-3 vs the fp32 reference)") if args.precision == "tc"
                   else "fp32 SIMT FFMA everywhere",
                   "clips_per_gpu": n, "frames_per_clip": t, "parallelism": f"clip-sharded x{world}, no collective",
                   "l2": "per-step working set (>= 4 GB of HR planes) exceeds the 126 MB L2",
                   "launch": ("whole-clip CUDA graph replay (gpu_launches counts the kernels inside the graphs)"
                              if model.use_graphs else "eager launches") +
                             "; a 30 ms device spin ahead of the start event lets the host enqueue the timed region early"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
