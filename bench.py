#!/usr/bin/env python
"""bench.py — output frames/s of the CRFP hot path (CRFP_DSV.forward) on B200, per the driver contract.

A "step" is one forward of this rank's synthetic clip(s) through the drop-in module:
  workload "R-lit"  : LR 180x320 (REDS *_sharp_BI frames, BASELINE.json configs[1]) through the x8 network
                      -> 1440x2560, t frames (default 100)
  workload "R-nat"  : LR 90x160 -> 720x1280 (the x8 network's native route to 1280x720)
  workload "V7"     : LR 64x112, t=7 (configs[0], the reference's CPU-runnable case)
`value`    = frames/s with the clip already resident in HBM (CUDA events, max over ranks).
`e2e`      = frames/s through the public API from HOST buffers: pinned lrs + fovea patches + coords H2D, forward,
             every output frame D2H into pinned memory as 8-bit RGB (quantised on the device with the reference's own
             (sr*255).clip(0,255).round()), all inside the timed region; event-timed AND wall-clock-timed.
`e2e_f32`  = the same with fp32 frames (4x the D2H bytes; on an 8-GPU box bound by the host side of the D2H stream).
`roofline` = the dominant kernel (conv_tc3_ws, per-frame launch mix) and the align kernel (DCNv2 @L1, the kernel
             BASELINE.json's metric names), each timed alone through the C ABI with CUDA events; `traffic` is read from the
             committed ncu DRAM-byte capture under profiles/ (null when there is none for the workload).
`cpu_baseline` / `--impl reference` = the oracle port of the reference's PyTorch path on the host cores, steady-state
             frames only.   `gpu_stock_baseline` = the same PyTorch path moved to the B200 (cuDNN,
             ATen grid_sample, torchvision deform_conv2d; TF32 off and on) — the GPU baseline the kernels must beat.
`extra`    = the other shapes (R-nat, V7) measured the same way, shorter; the reduced-precision tier on the headline
             workload; the streaming shell's per-frame latency.
--total-clips N : BASELINE.json configs[2] — N clips strong-sharded across the ranks (N / world per GPU, sequentially).
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"R-lit": (180, 320, 100), "R-nat": (90, 160, 100), "V7": (64, 112, 7)}
METRIC = "output frames/sec"
ALIGN_BYTES_PER_L1_PX = 1120  # (32 in + 144 offset + 72 mask + 32 out) fp32, SURVEY.md 8(d)
FV = 96


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crfp_b200", choices=["crfp_b200", "reference"])
    ap.add_argument("--workload", default="R-lit", choices=list(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override frames per clip")
    ap.add_argument("--clips", type=int, default=1, help="clips per forward call (batch dimension n)")
    ap.add_argument("--total-clips", type=int, default=0,
                    help="configs[2]: this many clips in total, strong-sharded across the ranks (total/world per GPU per step)")
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32", "half"],
                    help="tc: fp32 storage, tcgen05 3 x bf16 split contractions (fp32-grade, default); fp32: all-SIMT FFMA; "
                         "half: reduced-precision tier (north_star's bf16 tier)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the R-nat / V7 extra lines and the stock-PyTorch GPU baseline")
    return ap.parse_args()


def workload_config(workload, h, w, t, clips_per_step, total_clips=0):
    """The WORKLOAD description both arms print (identical dicts -> same_config)."""
    tag = {"R-lit": "BASELINE.json configs[1]", "R-nat": "the x8 network's route to 1280x720", "V7": "BASELINE.json configs[0]"}[workload]
    cfg = {"workload": f"{workload}: LR {h}x{w} -> {8 * h}x{8 * w} (x8 network; {tag}), "
                       f"{t}-frame clip, fovea {FV}x{FV}, CRFP_DSV mid_channels=32",
           "clips_per_gpu": clips_per_step, "frames_per_clip": t,
           "l2": "per-step working set (>= 4 GB of HR planes at R-lit) exceeds the 126 MB L2"}
    if total_clips:
        cfg["total_clips"] = total_clips
    return cfg


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


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_steady(h, w, frames, threads):
    """The reference's PyTorch path (oracle port, bit-identical to the reference on CPU) on the host cores, one clip of
    `frames` frames.  Returns (steady-state frames, seconds attributed to them): the clip-level work (FNet flows,
    encoders) is prorated per frame and the first frame — which has no alignment path at all (CRFP.py:1634-1670) — is
    excluded, i.e. frames 2..t as BASELINE.md 3.5 prescribes."""
    import torch
    from crfp_b200.synthetic import make_clip, make_state_dict
    from oracle import crfp_oracle as O
    torch.set_num_threads(threads)
    sd = make_state_dict(seed=1)
    lrs, fvs, mks, _ = make_clip(seed=2, n=1, t=frames, h=h, w=w, fv_size=FV)
    with torch.no_grad():
        t0 = time.perf_counter()
        flows = O.compute_flow(sd, lrs)
        x_lr, x_hr = O.encoders(sd, lrs, fvs, mks)
        t_clip = time.perf_counter() - t0
        state, t_frames = None, []
        for i in range(frames):
            t1 = time.perf_counter()
            out, state = O.frame_step(sd, 32, state, x_lr[:, i], x_hr[:, i], mks[:, i], lrs[:, i],
                                      flows[:, i - 1] if i else None)
            t_frames.append(time.perf_counter() - t1)
        assert out.shape[-2:] == (8 * h, 8 * w)
    steady = frames - 1
    return steady, t_clip * steady / frames + sum(t_frames[1:]), t_clip + sum(t_frames)


def run_reference(args):
    """`--impl reference`: rank 0 times the CPU path on a bounded sample of the same workload (3-frame clips; the rate
    counts the 2 steady-state frames of each)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    h, w, t = WORKLOADS[args.workload]
    if args.frames:
        t = args.frames
    frames = 3
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 2)):          # the CPU path has no lazy state worth more than two warm-ups
        cpu_reference_steady(h, w, frames, cores)
    nf, total, wall = 0, 0.0, 0.0
    for _ in range(args.steps):
        k, dt, full = cpu_reference_steady(h, w, frames, cores)
        nf, total, wall = nf + k, total + dt, wall + full
    value = nf / total
    sample = (f"{frames}-frame clips of {args.workload} (LR {h}x{w} -> {8 * h}x{8 * w}), one per step; rate = steady-state "
              f"frames 2..{frames} (clip-level FNet / encoder work prorated, the alignment-free first frame excluded), "
              f"oracle port of the reference's PyTorch path, torch CPU fp32, {wall / args.steps:.1f} s wall per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, h, w, t, args.clips, args.total_clips),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ stock PyTorch on the GPU
def gpu_stock_baseline(torch, h, w, t=5):
    """The reference's PyTorch path (oracle port) moved to the B200 as is: cuDNN convs, ATen grid_sample, torchvision
    deform_conv2d.  Protocol of /root/reference/test_runtime.py:142-186: 30 runs, the first 9 are warm-up, CUDA events
    around each of the 21 timed runs.  TF32 off and on.  A BASELINE leg (like cpu_baseline), not a product path."""
    from crfp_b200.synthetic import make_clip, make_state_dict
    from oracle import crfp_oracle as O
    sd = {k: v.cuda() for k, v in make_state_dict(seed=1).items()}
    lrs, fvs, mks, _ = make_clip(seed=2, n=1, t=t, h=h, w=w, fv_size=FV)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    res = {"frames_per_run": t, "runs": 21, "warmup_runs": 9,
           "what": "oracle port of the reference's PyTorch path on the same B200 (cuDNN / ATen / torchvision deform_conv2d)"}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            total = 0.0
            with torch.no_grad():
                for i in range(30):
                    torch.cuda.synchronize()
                    a.record()
                    O.crfp_dsv_forward(sd, lrs, fvs, mks)
                    b.record()
                    torch.cuda.synchronize()
                    if i >= 9:
                        total += a.elapsed_time(b) * 1e-3
            res[name + "_fps"] = t * 21 / total
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return res


# ------------------------------------------------------------------------------------------------ kernels timed alone
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
    """Average device time of the DCNv2 @L1 align kernel alone (C ABI), 3 rotating input sets > L2 (cold launches).
    Tensor-core precisions run it the way the frame does: RAW head-conv outputs in, tanh / sigmoid / + flow applied by
    the sampler (crfp_dcn_desc.head_raw) unless CRFP_HEAD_EPI=1 restores the two-pass form."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_dcn, pack_dcn_tc3
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    raw = precision != "fp32" and os.environ.get("CRFP_HEAD_EPI") is None
    nbuf = 3  # 3 x (29 + 199) MB > 126 MB L2
    xs = [torch.randn(1, h1, w1, 32, generator=g).to(dev) for _ in range(nbuf)]
    oms, flows = [], []
    for _ in range(nbuf):
        fl = torch.randn(1, h1, w1, 2, generator=g) * 2.0
        om = torch.empty(1, h1, w1, 216)
        pre_off = torch.randn(1, h1, w1, 144, generator=g) * 0.37          # 10 * tanh(.) ~ a few pixels
        pre_msk = torch.randn(1, h1, w1, 72, generator=g)
        if raw:
            om[..., :144], om[..., 144:] = pre_off, pre_msk
        else:
            om[..., :144] = torch.tanh(pre_off) * 10.0 + fl.flip(-1).repeat(1, 1, 1, 72)
            om[..., 144:] = torch.sigmoid(pre_msk)
        oms.append(om.to(dev)); flows.append(fl.to(dev))
    wt = (torch.randn(32, 32, 3, 3, generator=g) * 0.05).to(dev)
    out = torch.empty(1, h1, w1, 32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if precision != "fp32":
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
        if raw:
            d.head_raw, d.head_flow, d.head_mag = 1, flows[k].data_ptr(), 10.0
        if precision != "fp32":
            L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), flows[k].data_ptr(), st), "dcn_v2_tc3")
        else:
            L.check(L.lib().crfp_dcn_v2_fwd(C.byref(d), st), "dcn_v2")

    return _events_avg_ms(torch, launch, reps), raw


def time_align_fused(torch, h1, w1, precision, reps=12):
    """Average device time of crfp_dcn_align_fused (offset / mask heads + activations + DCNv2 @L1 in ONE kernel: what the
    frame runs in the tensor-core precisions), 3 rotating input sets, cold launches."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_align_heads, pack_dcn_tc3
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    wdt = torch.float16 if precision == "half" else torch.bfloat16
    nbuf = 3
    zs = [torch.randn(1, h1, w1, 32, generator=g).to(dev) for _ in range(nbuf)]
    xs = [torch.randn(1, h1, w1, 32, generator=g).to(dev) for _ in range(nbuf)]
    fls = [(torch.randn(1, h1, w1, 2, generator=g) * 2.0).to(dev) for _ in range(nbuf)]
    w_off = (torch.randn(144, 32, 3, 3, generator=g) * 0.02).to(dev)     # 10 * tanh(.) ~ a few pixels
    w_msk = (torch.randn(72, 32, 3, 3, generator=g) * 0.05).to(dev)
    wf, bf = pack_align_heads(w_off, torch.zeros(144, device=dev), w_msk, torch.zeros(72, device=dev), wdt)
    hi, lo, bp = pack_dcn_tc3((torch.randn(32, 32, 3, 3, generator=g) * 0.05).to(dev), torch.zeros(32, device=dev), 8, wdt)
    out = torch.empty(1, h1, w1, 32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launch(i):
        k = i % nbuf
        d = L.AlignFusedDesc(n=1, h=h1, w=w1, z=zs[k].data_ptr(), z_cstride=32, z_coffset=0, flow=fls[k].data_ptr(),
                             x=xs[k].data_ptr(), x_cstride=32, x_coffset=0, heads_w=wf.data_ptr(), heads_b=bf.data_ptr(),
                             dcn_w_hi=hi.data_ptr(), dcn_w_lo=lo.data_ptr(), dcn_b=bp.data_ptr(), out=out.data_ptr(),
                             out_cstride=32, out_coffset=0, head_mag=10.0, half=int(precision == "half"))
        L.check(L.lib().crfp_dcn_align_fused(C.byref(d), st), "dcn_align_fused")

    return _events_avg_ms(torch, launch, reps)


def time_conv_mix(torch, h1, w1, reps=4, precision="tc", with_heads=True):
    """The per-frame mix of tensor-core conv launches (conv_tc3_ws_kernel, the dominant kernel of the step): every L1
    layer shape of one steady-state frame through crfp_conv3x3_tc3_fwd.  Returns (avg ms per launch, launches,
    algorithmic bytes per launch = unique fp32 operands in + out, useful fp32 FLOP per launch)."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_conv_tc3
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    raw_heads = os.environ.get("CRFP_HEAD_EPI") is None
    # (source channels, extra flow channels, cout, count per frame, output kind)
    layers = [([32, 32], 2, 32, 3, "nhwc"), ([32], 0, 32, 9, "nhwc"), ([32, 32], 0, 32, 5, "nhwc"),
              ([32], 0, 216, 3, "nhwc"), ([24], 0, 64, 1, "shuffle4"), ([32], 0, 64, 1, "shuffle4")]
    if not with_heads:   # the fused align kernel computes the 216-channel heads itself: no such conv launch in the frame
        layers = [l for l in layers if l[2] != 216]
    wdt = torch.float16 if precision == "half" else torch.bfloat16
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
        hi, lo, bp, wx = pack_conv_tc3(w, torch.zeros(cout, device=dev), srcs, extra=extra, dtype=wdt)
        keep.append((hi, lo, bp, wx))
        for _ in range(count):
            d = L.ConvTc3Desc()
            d.n, d.h, d.w, d.nsrc = 1, h1, w1, len(srcs)
            for i, c in enumerate(srcs):
                t = buf24 if c == 24 else bufs[(rot + i) % 5]
                d.src[i] = L.TcSrc(ptr=t.data_ptr(), c=c, cstride=t.shape[-1], coffset=0)
            d.cout, d.act = cout, 1
            if cout == 216:   # the fused offset + mask heads: raw outputs (the align kernel's sampler activates them)
                d.act = 0     # ... or the two-pass form: 10 * tanh + flow on 144 channels, sigmoid on 72, in this epilogue
                if not raw_heads:
                    d.act, d.flow, d.head_split, d.head_mag = L.ACT_DCN_HEAD, flow.data_ptr(), 144, 10.0
            d.weight_hi, d.weight_lo, d.bias = hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
            if extra:
                d.extra, d.w_extra = flow.data_ptr(), wx.data_ptr()
            d.post_scale, d.ndst = 1.0, 1
            d.half = int(precision == "half")
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


def dram_traffic_from_profiles(workload, kernel_regex, grid_regex=None):
    """Mean dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels matching `kernel_regex`, from the newest
    committed ncu capture profiles/rNN/*dram*<workload>*.csv.  Returns (bytes or None, path or None, launches)."""
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*", f"*dram*{workload}*.csv")),
                   key=lambda p: (re.findall(r"profiles/r(\d+)", p.replace(os.sep, "/"))[-1], os.path.basename(p)))
    if not paths:
        return None, None, 0
    path = paths[-1]
    per = {}
    try:
        with open(path, newline="") as f:
            rows = [r for r in csv.reader(f) if len(r) >= 15 and r[0].isdigit()]
        for r in rows:
            if r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and re.search(kernel_regex, r[4]) and \
                    (grid_regex is None or re.search(grid_regex, r[8])):
                v = float(r[14].replace(",", ""))
                v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1.0)
                per[r[0]] = per.get(r[0], 0.0) + v
    except Exception:
        return None, None, 0
    if not per:
        return None, os.path.relpath(path, ROOT), 0
    return sum(per.values()) / len(per), os.path.relpath(path, ROOT), len(per)


# ------------------------------------------------------------------------------------------------ one workload
def make_host_inputs(torch, n, t, h, w, seed):
    """Synthetic clip: smooth-ish LR frames, Gaussian gaze walk, random fovea patches; pinned."""
    H, W_ = 8 * h, 8 * w
    g = torch.Generator(device="cpu").manual_seed(seed)
    coarse = torch.rand(n, t, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
    lrs_h = torch.nn.functional.interpolate(coarse.view(n * t, 3, *coarse.shape[-2:]), size=(h, w), mode="bicubic",
                                            align_corners=False).view(n, t, 3, h, w)
    lrs_h = (lrs_h + 0.05 * torch.rand(n, t, 3, h, w, generator=g)).clamp_(0, 1).contiguous().pin_memory()
    patch_h = torch.rand(n, t, 3, FV, FV, generator=g).pin_memory()
    gy = (torch.randn(n, t, generator=g) * 50 + H / 2).floor().long() - FV // 2
    gx = (torch.randn(n, t, generator=g) * 50 + W_ / 2).floor().long() - FV // 2
    coords = torch.stack([gy.clamp(0, H - FV), gx.clamp(0, W_ - FV)], -1)
    return lrs_h, patch_h, coords


_wc_keep = []


def pinned_output(torch, shape):
    """Pinned host buffer for the streamed output frames.  CRFP_PIN_WC=1: write-combined pinned memory (cudaHostAlloc
    with cudaHostAllocWriteCombined) — device writes over PCIe skip the CPU cache snoop, which matters when 8 GPUs
    stream into one socket; the host then reads it with streaming loads only (A/B switch, default: torch's pin_memory)."""
    if os.environ.get("CRFP_PIN_WC"):
        try:
            import ctypes
            rt = ctypes.CDLL("libcudart.so.12")
            n = 1
            for s_ in shape:
                n *= s_
            p = ctypes.c_void_p()
            rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n * 4), ctypes.c_uint(0x04 | 0x01))   # WC | portable
            if rc == 0 and p.value:
                buf = (ctypes.c_float * n).from_address(p.value)
                tns = torch.frombuffer(buf, dtype=torch.float32).view(*shape)
                if tns.is_pinned():
                    _wc_keep.append(buf)
                    return tns
        except Exception:
            pass
    return torch.empty(*shape, dtype=torch.float32).pin_memory()


def measure(torch, dist, model, workload, t, n, calls, steps, warmup, world, rank, dev, do_e2e, sampler=None):
    """Device-resident and end-to-end throughput of `calls` forward calls of n clips each per step."""
    from crfp_b200 import _lib
    h, w, _ = WORKLOADS[workload]
    H, W_ = 8 * h, 8 * w
    lrs_h, patch_h, coords = make_host_inputs(torch, n, t, h, w, 100 + rank)

    lrs = lrs_h.to(dev, non_blocking=True)
    patch = patch_h.to(dev, non_blocking=True)
    fvs = torch.zeros(n, t, 3, H, W_, device=dev)
    mks = torch.zeros(n, t, 1, H, W_, device=dev, dtype=torch.bool)
    for b in range(n):
        for i in range(t):
            y, x = int(coords[b, i, 0]), int(coords[b, i, 1])
            fvs[b, i, :, y:y + FV, x:x + FV] = patch[b, i]
            mks[b, i, :, y:y + FV, x:x + FV] = True
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_decouple():
        """A device spin kernel ahead of the start event: the host enqueues the timed region (graph launches) while the
        GPU is still spinning, so a host stall (this pool's VMs page memory in lazily: 50-150 ms hiccups were measured)
        cannot starve the GPU inside the device-timed region.  The spin ends before the start event fires."""
        torch.cuda._sleep(int(min(1.0, 0.2 + 0.01 * steps * calls) * 1.9e9))

    def step():
        for _ in range(calls):             # the clips of this rank, one forward call each (same resident buffers)
            o = model(lrs, fvs, mks)
        return o

    for _ in range(warmup):
        out = step()
    barrier()
    _lib.lib().crfp_launch_count_reset()
    if sampler is not None:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_decouple()
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(_lib.lib().crfp_launch_count())
    clocks = sampler.stop() if sampler is not None else None
    assert torch.isfinite(out[:, -1]).all()
    tms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    frames_total = world * n * t * calls * steps
    res = {"value": frames_total / (ms_max * 1e-3), "ms_per_step": ms_max / steps, "gpu_launches": launches, "clocks": clocks}

    if do_e2e:
        lrs_d, patch_d = torch.empty_like(lrs_h, device=dev), torch.empty_like(patch_h, device=dev)
        del fvs, mks, out
        h2d_bytes = int(lrs_h.numel() * 4 + patch_h.numel() * 4 + coords.numel() * 4) * calls

        def run_e2e(out_h, pinned_kind):
            """K end-to-end steps delivering every frame into the pinned host tensor `out_h` (fp32 or uint8): event-timed after
            the spin, then once more under the wall clock."""
            def e2e_step():
                for _ in range(calls):
                    lrs_d.copy_(lrs_h, non_blocking=True)       # H2D from pinned memory, every call
                    patch_d.copy_(patch_h, non_blocking=True)
                    model.forward_patch(lrs_d, patch_d, coords, out_host=out_h)   # coords: host integers, copied inside
                    # every frame is D2H-copied to pinned memory on a side stream while later frames compute

            model._graphs.clear()
            for _ in range(max(2, min(warmup, 3))):
                e2e_step()
            barrier()
            host_decouple()      # as above: the K steps (H2D copies, graph launches, D2H copies) are enqueued during the spin
            t0 = time.perf_counter()
            e0.record()
            for _ in range(steps):
                e2e_step()
            e1.record()
            enq = (time.perf_counter() - t0) * 1e3
            barrier()
            ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
            # the same K steps once more under the WALL clock, without the spin: barrier + sync on both sides
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            barrier()
            wall = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
                dist.all_reduce(wall, op=dist.ReduceOp.MAX)
            d2h = int(out_h.numel() * out_h.element_size()) * calls
            return {"value": frames_total / (float(ems.item()) * 1e-3), "unit": "frames/s",
                    "ms_per_step": float(ems.item()) / steps,
                    "wall_value": frames_total / (float(wall.item()) * 1e-3), "wall_ms_per_step": float(wall.item()) / steps,
                    "host_enqueue_ms_per_step": enq / steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                    "d2h_gbs_per_rank": d2h / (float(ems.item()) / steps * 1e-3) / 1e9, "pinned": pinned_kind}

        # Headline end-to-end number: the frames reach the host as the reference STORES them — 8-bit RGB, quantised on the device
        # with the reference's own expression (sr * 255).clip(0, 255).round() (trainer.py:446-474) — so a frame is 11 MB over PCIe.
        out_u8 = torch.empty(n, t, 3, H, W_, dtype=torch.uint8).pin_memory()
        res["e2e"] = run_e2e(out_u8, "torch pin_memory")
        res["e2e"]["frame_format"] = "uint8 RGB planes, (sr*255).clip(0,255).round() computed on the device"
        res["e2e"]["api"] = ("CRFP_DSV.forward_patch(lrs, fovea_patch, coords, out_host=pinned uint8): H2D of lrs + patches + coords, "
                             "device-side fovea paste, forward, every finished frame quantised and D2H-copied on a side stream "
                             "while the recurrence continues. `value`: CUDA events around the K steps (copies included, host "
                             "enqueue time beside it); `wall_value`: the same K steps under time.perf_counter between two "
                             "barrier+synchronize.  e2e_f32 = the same with fp32 frames (4x the D2H bytes: what the reference's "
                             "own eval loop moves with .round().cpu(); on this 8-GPU box its aggregate is bound by the host side "
                             "of the D2H stream, ~91 GB/s over all ranks, DESIGN.md 7)")
        del out_u8
        out_h = pinned_output(torch, (n, t, 3, H, W_))
        res["e2e_f32"] = run_e2e(out_h, "write-combined (cudaHostAllocWriteCombined)" if _wc_keep else "torch pin_memory")
        res["e2e_f32"]["frame_format"] = "fp32 RGB planes"
        del out_h
    return res


def stream_latency(torch, precision, h=134, w=240, frames=40):
    """Median wall-clock latency of one streaming call (MRCF_simple_v18, batch 1, one frame, device synchronised after the
    call): what a display loop sees per frame."""
    import statistics
    from crfp_b200 import MRCF_simple_v18
    from crfp_b200.synthetic import make_clip, make_state_dict
    lrs, fvs, mks, fv_sp = make_clip(seed=3, n=1, t=frames, h=h, w=w, fv_size=96)
    fgs = torch.zeros(1, frames, 1, 8 * h, 8 * w)
    for i in range(frames):
        cy, cx = int(fv_sp[0, i, 0]) + 48, int(fv_sp[0, i, 1]) + 48
        fgs[0, i, 0, max(cy - 270, 0):cy + 270, max(cx - 480, 0):cx + 480] = 1
    lrs, fvs, mks, fgs = lrs.cuda(), fvs.cuda(), mks.cuda(), fgs.cuda()
    m = MRCF_simple_v18("cuda", mid_channels=32, precision=precision).eval()
    m.load_state_dict(make_state_dict(seed=1), strict=True)
    m.cuda()
    m.alias_output = True
    lat = []
    for rep in range(2):                     # first pass warms up and captures the per-frame graph
        m.clear_states()
        lat = []
        for i in range(frames):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
    steady = sorted(lat[3:])
    return {"workload": f"streaming: MRCF_simple_v18, one frame per call at LR {h}x{w} -> {8 * h}x{8 * w}, regional fg mask, "
                        "device synchronised after every call (test_video.py:316-374)",
            "metric": "latency per frame", "unit": "ms", "value": statistics.median(steady), "p95": steady[int(0.95 * len(steady))],
            "frames_per_s": 1e3 / statistics.median(steady), "higher_is_better": False, "precision": precision,
            "launch": "per-frame CUDA-graph replay"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from crfp_b200 import CRFP_DSV
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
    calls, scaling = 1, "weak"
    if args.total_clips:
        if args.total_clips % (world * n):
            raise SystemExit("--total-clips must be a multiple of gpus x clips")
        calls, scaling = args.total_clips // (world * n), "strong"

    model = CRFP_DSV("cuda", mid_channels=32, precision=args.precision).eval()
    model.load_state_dict(make_state_dict(seed=1), strict=True)
    model.to(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()          # nvidia-smi -lms 100 starts sampling during the warm-up, keeps going through the timed region
    main_res = measure(torch, dist, model, args.workload, t, n, calls, args.steps, args.warmup, world, rank, dev,
                       not args.no_e2e, sampler)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- rooflines (kernels timed alone, cold inputs)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    px1 = (2 * h) * (2 * w)
    fused = args.precision != "fp32" and os.environ.get("CRFP_ALIGN_UNFUSED") is None and os.environ.get("CRFP_HEAD_EPI") is None
    if fused:
        # the align kernel of the tensor-core precisions: DCN_module's heads + activations + DCNv2 in one launch
        a_ms = time_align_fused(torch, 2 * h, 2 * w, args.precision)
        fused_bytes = (32 + 2 + 32 + 32) * 4 * px1           # z + flow + x in, aligned out: 392 B per L1 pixel
        unfused_bytes = ALIGN_BYTES_PER_L1_PX * px1            # what the DCNv2 op alone moves when offsets / masks live in HBM
        flop = (3 if args.precision == "tc" else 2) * 2 * 288 * (224 + 32) * px1   # issued 16-bit MMA FLOP (split products)
        a_traffic, a_src, a_cnt = dram_traffic_from_profiles(args.workload, r"dcn_align_fused_kernel")
        align = {"kernel": "dcn_align_fused_kernel (dcn_offset + dcn_mask convs, 10*tanh + flow / sigmoid, DCNv2 @L1 C=32 dg=8: "
                           "the 216-channel offset / mask tensor stays in TMEM)",
                 "bound": "tensor", "achieved": flop / (a_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                 "frac": flop / (a_ms * 1e-3) / 1e12 / bf16_peak,
                 "note": "bound by the shared-memory pipe feeding small-N tcgen05 MMAs and the bilinear gather, not by HBM; "
                         "the HBM view is kept below for continuity with the DCNv2-op figure of SURVEY.md 8(d)",
                 "hbm": {"algorithmic_bytes_per_launch": fused_bytes, "achieved_gbs": fused_bytes / (a_ms * 1e-3) / 1e9,
                         "frac_of_hbm_peak": fused_bytes / (a_ms * 1e-3) / 1e9 / peak,
                         "dcnv2_op_equivalent_bytes": unfused_bytes,
                         "dcnv2_op_equivalent_frac": unfused_bytes / (a_ms * 1e-3) / 1e9 / peak},
                 "traffic": a_traffic,
                 "traffic_source": (f"dram__bytes_read+write per launch, mean of {a_cnt} launches, ncu, {a_src}" if a_traffic else
                                    "no committed ncu DRAM capture for this workload"),
                 "issued_flop_per_launch": flop, "avg_launch_ms": a_ms,
                 "timing": "kernel timed alone through the C ABI, CUDA events, 3 rotating input sets > L2"}
    else:
        a_ms, a_raw = time_align_kernel(torch, 2 * h, 2 * w, args.precision)
        alg_bytes = ALIGN_BYTES_PER_L1_PX * px1
        a_ach = alg_bytes / (a_ms * 1e-3) / 1e9
        a_traffic, a_src, a_cnt = dram_traffic_from_profiles(args.workload, r"dcn_tc3_ws_kernel|dcn_l1_kernel")
        align = {"kernel": ("dcn_tc3_ws_kernel" if args.precision != "fp32" else "dcn_l1_kernel") + " (DCNv2 align @L1, C=32 dg=8" +
                           (", head activations fused into the sampler)" if a_raw else ")"),
                 "bound": "hbm", "achieved": a_ach, "peak": peak, "unit": "GB/s", "frac": a_ach / peak,
                 "traffic": a_traffic,
                 "traffic_source": (f"dram__bytes_read+write per launch, mean of {a_cnt} launches, ncu, {a_src}" if a_traffic else
                                    "no committed ncu DRAM capture for this workload"),
                 "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": a_ms,
                 "timing": "kernel timed alone through the C ABI, CUDA events, 3 rotating input sets > L2"}
    if args.precision != "fp32":
        c_ms, c_n, c_bytes, c_flop = time_conv_mix(torch, 2 * h, 2 * w, precision=args.precision, with_heads=not fused)
        c_ach = c_bytes / (c_ms * 1e-3) / 1e9
        c_traffic, c_src, c_cnt = dram_traffic_from_profiles(args.workload, r"conv_tc3_ws_kernel", r"\((5|10), ")
        roofline = {"kernel": "conv_tc3_ws_kernel (tcgen05 3x3 implicit-GEMM conv, the per-frame mix of its %d L1 launches)" % c_n,
                    "bound": "hbm", "achieved": c_ach, "peak": peak, "unit": "GB/s", "frac": c_ach / peak,
                    "traffic": c_traffic,
                    "traffic_source": (f"dram__bytes_read+write, mean over the {c_cnt} L1 conv launches of the captured frames, ncu, "
                                       f"{c_src} (below the algorithmic bytes: L2 reuse)" if c_traffic else
                                       "no committed ncu DRAM capture for this workload"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": c_bytes, "avg_launch_ms": c_ms,
                    "tensor": {"useful_fp32_tflops": c_flop / (c_ms * 1e-3) / 1e12,
                               "issued_bf16_tflops": (3 if args.precision == "tc" else 2) * c_flop / (c_ms * 1e-3) / 1e12,
                               "bf16_peak_tflops": bf16_peak,
                               "frac_of_bf16_peak": (3 if args.precision == "tc" else 2) * c_flop / (c_ms * 1e-3) / 1e12 / bf16_peak},
                    "timing": "all L1 conv launches of one steady-state frame replayed back to back through "
                              "crfp_conv3x3_tc3_fwd, CUDA events, rotating 29.5 MB planes (5 x 29.5 + HR outputs > L2)",
                    "align_kernel": align}
    else:
        roofline = dict(align)
        roofline["peak_source"] = peak_src

    # ---------------- baselines measured in the same run (rank 0, N = 1 only)
    cpu = stock = None
    extra = []
    if world == 1:
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fr = 5
            k, dt, wall = cpu_reference_steady(h, w, fr, cores)
            cpu = {"value": k / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"one {fr}-frame clip of {args.workload} (LR {h}x{w}); rate = steady-state frames 2..{fr} "
                             f"(BASELINE.md 3.5; clip-level work prorated), oracle port of the reference's PyTorch path, "
                             f"torch CPU fp32, {wall:.1f} s"}
        if not args.no_extras:
            try:
                stock = gpu_stock_baseline(torch, h, w)
                stock["workload"] = args.workload
            except Exception as e:  # noqa: BLE001  (a baseline leg must never take the bench line down)
                stock = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            for wl in WORKLOADS:
                if wl == args.workload or args.total_clips:
                    continue
                model._graphs.clear()
                model._ws.clear()
                torch.cuda.empty_cache()
                r = measure(torch, dist, model, wl, WORKLOADS[wl][2], 1, 1, max(2, min(args.steps, 5)), 3, 1, 0, dev,
                            not args.no_e2e)
                extra.append({"workload": workload_config(wl, *WORKLOADS[wl][:2], WORKLOADS[wl][2], 1)["workload"],
                              "value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"],
                              "e2e": {k: r["e2e"][k] for k in ("value", "wall_value", "ms_per_step")} if "e2e" in r else None,
                              "precision": args.precision, "dtype": "f32" if args.precision != "half" else "f16"})

        if not args.no_extras and args.precision == "tc" and not args.total_clips:
            # the reduced-precision tier (north_star's "bf16" tier, <= 5e-3: tests/test_gpu_half.py) on the headline workload
            try:
                del model
                torch.cuda.empty_cache()
                mh = CRFP_DSV("cuda", mid_channels=32, precision="half").eval()
                mh.load_state_dict(make_state_dict(seed=1), strict=True)
                mh.to(dev)
                r = measure(torch, dist, mh, args.workload, t, 1, 1, max(2, min(args.steps, 5)), 3, 1, 0, dev, not args.no_e2e)
                extra.append({"workload": workload_config(args.workload, h, w, t, 1)["workload"], "precision": "half",
                              "dtype": "f16 operands (one fp16 activation product x fp16 hi/lo split weights), fp32 storage and accumulation",
                              "value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"],
                              "e2e": {k: r["e2e"][k] for k in ("value", "wall_value", "ms_per_step")} if "e2e" in r else None,
                              "parity": "max-abs <= 5e-3 vs the reference goldens over 100 frames (tests/test_gpu_half.py)"})
                model = mh
            except Exception as e:  # noqa: BLE001
                extra.append({"precision": "half", "unavailable": f"{type(e).__name__}: {e}"[:200]})

        if not args.no_extras and not args.total_clips:
            # the reference's real-time protocol: one frame per call through the streaming shell, a device synchronisation
            # after every call (test_video.py:316-374), LR 134x240 -> 1072x1920 (test_video.py:234-240), regional fg mask
            try:
                extra.append(stream_latency(torch, args.precision))
            except Exception as e:  # noqa: BLE001
                extra.append({"workload": "streaming", "unavailable": f"{type(e).__name__}: {e}"[:200]})

    cfg = workload_config(args.workload, h, w, t, n * calls, args.total_clips)
    prec = {"tc": "fp32 storage; dense contractions as 3 x bf16 split products on tcgen05 with fp32 TMEM accumulation "
                  "(parity <= 1e-3 vs the fp32 reference)",
            "fp32": "fp32 SIMT FFMA everywhere",
            "half": "reduced-precision tier: fp32 storage, 16-bit tensor-core operands (see DESIGN.md)"}[args.precision]
    line = {
        "metric": METRIC, "value": main_res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32" if args.precision != "half" else "f16", "data": "synthetic",
        "config": cfg,
        "impl_config": {"precision": prec, "clips_per_call": n, "calls_per_step": calls,
                        "parallelism": f"clip-sharded x{world}, no collective",
                        "launch": ("whole-clip CUDA graph replay (gpu_launches counts the kernels inside the graphs)"
                                   if model.use_graphs else "eager launches") +
                                  "; a >= 200 ms device spin ahead of the start event lets the host enqueue the timed region early"},
        "e2e": main_res.get("e2e"), "e2e_f32": main_res.get("e2e_f32"), "gpu_launches": main_res["gpu_launches"],
        "clocks": main_res["clocks"],
        "roofline": roofline, "cpu_baseline": cpu, "gpu_stock_baseline": stock, "extra": extra,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
