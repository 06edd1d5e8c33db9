"""Per-frame latency of the streaming protocol (1 frame per call, batch 1; /root/reference/test_video.py:316-374 with
model/CRFP_test.py:2250-2478): LR 134x240 -> 1072x1920 by default (test_video.py:234-240), fovea 96x96, regional fg mask.
Every call is followed by a device synchronisation (the caller displays the frame), so the number is wall-clock latency
per frame: host launch work + device time.  Eager launches vs the per-frame CUDA-graph replay of MRCF_simple_v18.
usage: python scripts/bench_stream.py [--h 134 --w 240 --frames 60]"""
import argparse, json, os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import MRCF_simple_v18
from crfp_b200.synthetic import make_clip, make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=134)
ap.add_argument("--w", type=int, default=240)
ap.add_argument("--frames", type=int, default=60)
ap.add_argument("--modes", default="eager,graph")
a = ap.parse_args()
h, w, t = a.h, a.w, a.frames
lrs, fvs, mks, fv_sp = make_clip(seed=3, n=1, t=t, h=h, w=w, fv_size=96)
fgs = torch.zeros(1, t, 1, 8 * h, 8 * w)
for i in range(t):                       # regional DCN window around the gaze (test_video.py:352-358), 540x960 at 1080p
    cy, cx = int(fv_sp[0, i, 0]) + 48, int(fv_sp[0, i, 1]) + 48
    fgs[0, i, 0, max(cy - 270, 0):cy + 270, max(cx - 480, 0):cx + 480] = 1
lrs, fvs, mks, fgs = lrs.cuda(), fvs.cuda(), mks.cuda(), fgs.cuda()
res = {"shape": f"LR {h}x{w} -> {8 * h}x{8 * w}", "frames": t}
outs = {}
for mode in a.modes.split(","):
    m = MRCF_simple_v18("cuda", mid_channels=32).eval()
    m.load_state_dict(make_state_dict(seed=1), strict=True)
    m.cuda()
    m.use_graphs = mode == "graph"
    m.alias_output = True                # the displayed frame is consumed before the next call
    lat = []
    for rep in range(2):                 # first pass warms up (and captures); second pass is timed
        m.clear_states()
        lat, frames = [], []
        for i in range(t):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            o = m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
            frames.append(o.clone())
    outs[mode] = torch.cat(frames, 1)
    steady = lat[3:]
    res[mode] = {"ms_per_frame_median": statistics.median(steady), "ms_per_frame_p95": sorted(steady)[int(0.95 * len(steady))],
                 "ms_first_frame": lat[0], "fps": 1e3 / statistics.median(steady)}
if "eager" in outs and "graph" in outs:
    res["graph_equals_eager"] = bool(torch.equal(outs["eager"], outs["graph"]))
print(json.dumps(res))
