"""Where a streaming call's time goes (MRCF_simple_v18, one frame per call): per-call wall latency with and without the
device sync, device span per call (CUDA events), and the CUPTI kernel table of 10 steady-state calls.
usage: python scripts/stream_kernel_times.py [--h 134 --w 240] [--no-fg]"""
import argparse, collections, os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import MRCF_simple_v18
from crfp_b200.synthetic import make_clip, make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=134)
ap.add_argument("--w", type=int, default=240)
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--no-fg", action="store_true", help="fgs = all ones (no regional DCN window)")
a = ap.parse_args()
h, w, t = a.h, a.w, a.frames
lrs, fvs, mks, fv_sp = make_clip(seed=3, n=1, t=t, h=h, w=w, fv_size=96)
fgs = torch.ones(1, t, 1, 8 * h, 8 * w) if a.no_fg else torch.zeros(1, t, 1, 8 * h, 8 * w)
if not a.no_fg:
    for i in range(t):
        cy, cx = int(fv_sp[0, i, 0]) + 48, int(fv_sp[0, i, 1]) + 48
        fgs[0, i, 0, max(cy - 270, 0):cy + 270, max(cx - 480, 0):cx + 480] = 1
lrs, fvs, mks, fgs = lrs.cuda(), fvs.cuda(), mks.cuda(), fgs.cuda()
m = MRCF_simple_v18("cuda", mid_channels=32).eval()
m.load_state_dict(make_state_dict(seed=1), strict=True)
m.cuda()
m.use_graphs = False
m.alias_output = True


def one_pass(record=None):
    m.clear_states()
    for i in range(t):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = time.perf_counter()
        e0.record()
        m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
        e1.record()
        c1 = time.perf_counter()
        torch.cuda.synchronize()
        c2 = time.perf_counter()
        if record is not None:
            record.append(((c1 - c0) * 1e3, (c2 - c0) * 1e3, e0.elapsed_time(e1)))


one_pass()
rec = []
one_pass(rec)
st = rec[3:]
print(f"steady-state call at LR {h}x{w}: host enqueue {statistics.median(r[0] for r in st):.2f} ms, wall incl. sync "
      f"{statistics.median(r[1] for r in st):.2f} ms, device span {statistics.median(r[2] for r in st):.2f} ms; first call "
      f"{rec[0][1]:.2f} ms wall / {rec[0][2]:.2f} ms device")
from torch.profiler import profile, ProfilerActivity
m.clear_states()
for i in range(4):
    m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(4, 14):
        m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name.split("(")[0][:80]
        agg[k][0] += 1
        agg[k][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"device busy per call: {tot / 10 / 1e3:.3f} ms over {sum(v[0] for v in agg.values()) / 10:.0f} kernels / copies")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {v[1] / 10:9.1f} us/call {100 * v[1] / tot:5.1f} %  x{v[0] / 10:5.1f}  avg {v[1] / v[0]:7.1f} us  {k}")
