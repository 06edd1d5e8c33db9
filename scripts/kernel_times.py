"""In-pipeline kernel times (CUPTI via torch.profiler; kernels keep overlapping under PDL, unlike an ncu launch list)
plus per-step min / median frame time.  usage: python scripts/kernel_times.py [--workload R-lit] [--frames 20] [--steps 8]"""
import argparse, collections, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_state_dict
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="R-lit")
ap.add_argument("--frames", type=int, default=20)
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--no-profile", action="store_true")
a = ap.parse_args()
h, w, _ = bench.WORKLOADS[a.workload]
t, n, fv = a.frames, 1, 96
H, W = 8 * h, 8 * w
dev = torch.device("cuda")
model = CRFP_DSV("cuda", mid_channels=32).eval()
model.load_state_dict(make_state_dict(seed=1), strict=True)
model.to(dev)
g = torch.Generator(device="cpu").manual_seed(100)
coarse = torch.rand(n, t, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
lrs = torch.nn.functional.interpolate(coarse.view(n * t, 3, *coarse.shape[-2:]), size=(h, w), mode="bicubic", align_corners=False).view(n, t, 3, h, w)
lrs = (lrs + 0.05 * torch.rand(n, t, 3, h, w, generator=g)).clamp_(0, 1).to(dev)
fvs = torch.zeros(n, t, 3, H, W, device=dev)
mks = torch.zeros(n, t, 1, H, W, device=dev, dtype=torch.bool)
gy = (torch.randn(n, t, generator=g) * 50 + H / 2).floor().long().clamp(0, H - fv)
gx = (torch.randn(n, t, generator=g) * 50 + W / 2).floor().long().clamp(0, W - fv)
for i in range(t):
    y, x = int(gy[0, i]), int(gx[0, i])
    fvs[0, i, :, y:y + fv, x:x + fv] = torch.rand(3, fv, fv, generator=g).to(dev)
    mks[0, i, :, y:y + fv, x:x + fv] = True
for _ in range(3):
    model(lrs, fvs, mks)
torch.cuda.synchronize()
import time, gc
if os.environ.get('NOGC'): gc.collect(); gc.freeze(); gc.disable()
times, cpu = [], []
for _ in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.perf_counter()
    e0.record(); model(lrs, fvs, mks); e1.record()
    cpu.append((time.perf_counter() - c0) * 1e3 / t)
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / t)
print("per-step ms/frame:", " ".join(f"{x:.2f}" for x in times), "| cpu enqueue ms/frame:", " ".join(f"{x:.2f}" for x in cpu))
print(f"ms/frame: min {min(times):.3f} median {statistics.median(times):.3f} max {max(times):.3f}  -> {1e3 / statistics.median(times):.1f} fps (median)")
if not a.no_profile:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(lrs, fvs, mks)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name.split("(")[0][:70]][0] += 1
            agg[ev.name.split("(")[0][:70]][1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"{k:70s} n/frame={v[0] / t:6.2f} us/frame={v[1] / t:8.1f} avg={v[1] / v[0]:8.1f} us share={100 * v[1] / tot:5.1f}%")
    print(f"sum of kernel times {tot / t:.1f} us/frame")
