"""Where a training step's time goes: wall vs device time per step, and the per-kernel device-time table of one step
(CUPTI through torch.profiler; includes the ATen glue kernels).  usage: python scripts/train_kernel_times.py [v7|crop]"""
import collections, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_clip, make_state_dict
from crfp_b200.trainer import Trainer

shape = sys.argv[1] if len(sys.argv) > 1 else "v7"
graphs = len(sys.argv) > 2 and sys.argv[2] == "graphs"
n, t, h, w, fv = {"v7": (1, 7, 64, 112, 128), "crop": (8, 15, 32, 32, 128)}[shape]
model = CRFP_DSV("cuda", mid_channels=32)
model.load_state_dict(make_state_dict(seed=1), strict=True)
model.cuda()
tr = Trainer(model, freeze_flow_iters=0, use_graphs=graphs)
lrs, fvs, mks, _ = make_clip(seed=2, n=n, t=t, h=h, w=w, fv_size=fv)
hr = torch.rand(n, t, 3, 8 * h, 8 * w, generator=torch.Generator().manual_seed(3))
batch = (lrs.cuda(), fvs.cuda(), mks.cuda(), hr.cuda())
for _ in range(4 if graphs else 2):
    tr.step(*batch)
torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.perf_counter()
    e0.record(); tr.step(*batch); e1.record()
    c1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"step: host enqueue {1e3 * (c1 - c0):.1f} ms, device span {e0.elapsed_time(e1):.1f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:80]
        agg[name][0] += 1
        agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"device busy time of one step: {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} kernels / copies")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{us / 1e3:9.2f} ms {100 * us / tot:5.1f} %  x{cnt:<5d} {name}")
