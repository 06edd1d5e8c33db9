"""Host-side time of one streaming call, split by C-ABI entry point (ctypes wrappers timed with perf_counter) — finds what
the host spends its time on between the launches.  usage: python scripts/stream_host_times.py [--h 134 --w 240]"""
import argparse, collections, os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import MRCF_simple_v18, _lib as L
from crfp_b200.synthetic import make_clip, make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=134)
ap.add_argument("--w", type=int, default=240)
ap.add_argument("--frames", type=int, default=16)
a = ap.parse_args()
h, w, t = a.h, a.w, a.frames
lrs, fvs, mks, fv_sp = make_clip(seed=3, n=1, t=t, h=h, w=w, fv_size=96)
fgs = torch.zeros(1, t, 1, 8 * h, 8 * w)
for i in range(t):
    cy, cx = int(fv_sp[0, i, 0]) + 48, int(fv_sp[0, i, 1]) + 48
    fgs[0, i, 0, max(cy - 270, 0):cy + 270, max(cx - 480, 0):cx + 480] = 1
lrs, fvs, mks, fgs = lrs.cuda(), fvs.cuda(), mks.cuda(), fgs.cuda()
m = MRCF_simple_v18("cuda", mid_channels=32).eval()
m.load_state_dict(make_state_dict(seed=1), strict=True)
m.cuda()
m.use_graphs = False
m.alias_output = True

lib = L.lib()
acc = collections.defaultdict(list)


class Timed:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *args):
        t0 = time.perf_counter()
        r = self.fn(*args)
        acc[self.name].append((time.perf_counter() - t0) * 1e3)
        return r


class LibProxy:
    def __getattr__(self, name):
        return Timed(name, getattr(lib, name))


proxy = LibProxy()
L.lib = lambda: proxy
for rep in range(2):
    m.clear_states()
    acc.clear()
    tot = []
    for i in range(t):
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
        tot.append((time.perf_counter() - c0) * 1e3)
        torch.cuda.synchronize()
print(f"LR {h}x{w}: host time of one call {statistics.median(tot[3:]):.3f} ms; by C entry point (median of calls, ms):")
for k, v in acc.items():
    print(f"  {k:32s} x{len(v) / t:4.1f}/call  median {statistics.median(v):.3f}  max {max(v):.3f}")
# the same call with the device check and property query taken out of the loop
