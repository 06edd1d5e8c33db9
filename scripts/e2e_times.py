"""e2e (host buffers in, pinned host frames out) per-step wall times for different graph chunk sizes."""
import os, statistics, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_state_dict
t, h, w, fv = int(os.environ.get("FRAMES", "100")), 180, 320, 96
H, W = 8 * h, 8 * w
dev = torch.device("cuda")
model = CRFP_DSV("cuda", mid_channels=32).eval()
model.load_state_dict(make_state_dict(seed=1), strict=True)
model.to(dev)
g = torch.Generator().manual_seed(0)
lrs_h = torch.rand(1, t, 3, h, w, generator=g).pin_memory()
patch_h = torch.rand(1, t, 3, fv, fv, generator=g).pin_memory()
coords = torch.stack([torch.randint(0, H - fv, (1, t), generator=g), torch.randint(0, W - fv, (1, t), generator=g)], -1)
out_h = torch.empty(1, t, 3, H, W).pin_memory()
lrs_d, patch_d = torch.empty_like(lrs_h, device=dev), torch.empty_like(patch_h, device=dev)
for gf in [int(v) for v in os.environ.get("GF", "20,100,1000").split(",")]:
    model.graph_frames = gf
    model._graphs.clear(); model._seen_key = None
    ts = []
    for it in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lrs_d.copy_(lrs_h, non_blocking=True); patch_d.copy_(patch_h, non_blocking=True)
        model.forward_patch(lrs_d, patch_d, coords, out_host=out_h)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ts.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3))
    tot = [a for a, _ in ts[3:]]; enq = [b for _, b in ts[3:]]
    print(f"graph_frames {gf:5d}: step ms min {min(tot):7.1f} median {statistics.median(tot):7.1f} max {max(tot):7.1f} | enqueue ms median {statistics.median(enq):6.1f} -> {t / statistics.median(tot) * 1e3:6.1f} fps (median) {t / min(tot) * 1e3:6.1f} (best)")
model.use_graphs = False
ts = []
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lrs_d.copy_(lrs_h, non_blocking=True); patch_d.copy_(patch_h, non_blocking=True)
    model.forward_patch(lrs_d, patch_d, coords, out_host=out_h)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(f"eager: step ms min {min(ts[2:]):.1f} median {statistics.median(ts[2:]):.1f} -> {t / statistics.median(ts[2:]) * 1e3:.1f} fps")
