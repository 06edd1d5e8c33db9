"""Long-recurrence and full-size goldens from the REAL reference (run in the build container only).

TEST INFRASTRUCTURE.  Same shims as oracle/make_golden.py (dcn_v2 -> torchvision, stub fnet.pth).  Runs the unmodified
reference `CRFP_DSV` (/root/reference/model/CRFP.py:1510-1686) on
  long_t100_32x48      100-frame recurrence at a small LR size  (the risk SURVEY.md 7 names: error growth through
                       10*tanh offsets over the BASELINE clip length)
  rnat_t12_90x160      REDS-native shape (LR 90x160 -> 720x1280), 12 frames
  rlit_t3_180x320      the headline shape (LR 180x320 -> 1440x2560, BASELINE.json configs[1]), 3 frames
asserts that the oracle restatement is bit-identical on the FULL outputs, and stores compact fixtures (full outputs
would be 0.1-1 GB): for every frame a strided sub-grid of the output (every `stride`-th pixel from a per-frame
offset), a crop around the fovea rectangle, and float64 checksums (sum, sum of squares) of the whole frame.

Usage:  python oracle/make_golden_long.py            (writes tests/golden/long_*.pt)
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from crfp_b200.synthetic import make_clip, make_state_dict  # noqa: E402
from oracle import crfp_oracle as O  # noqa: E402
from oracle.make_golden import build_ref_model, load_reference  # noqa: E402

CASES = [
    # name, t, h, w, fv, seed, stride, crop
    ("long_t100_32x48", 100, 32, 48, 96, 41, 8, 32),
    ("rnat_t12_90x160", 12, 90, 160, 96, 42, 16, 48),
    ("rlit_t3_180x320", 3, 180, 320, 96, 43, 16, 64),
]


def compact(out, fv_sp, fv, stride, crop):
    """out (1,t,3,H,W) -> dict of per-frame sub-grids, fovea-centred crops and float64 checksums."""
    _, t, _, H, W = out.shape
    grids, crops, sums, sqs, org = [], [], [], [], []
    for i in range(t):
        oy, ox = (3 * i) % stride, (5 * i) % stride
        grids.append(out[0, i, :, oy::stride, ox::stride].contiguous())
        cy = min(max(int(fv_sp[0, i, 0]) + fv // 2 - crop // 2, 0), H - crop)
        cx = min(max(int(fv_sp[0, i, 1]) + fv // 2 - crop // 2, 0), W - crop)
        crops.append(out[0, i, :, cy:cy + crop, cx:cx + crop].contiguous())
        org.append((oy, ox, cy, cx))
        sums.append(float(out[0, i].double().sum()))
        sqs.append(float((out[0, i].double() ** 2).sum()))
    return {"grids": grids, "crops": torch.stack(crops), "origins": org, "sum": sums, "sumsq": sqs, "stride": stride,
            "crop": crop}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    CRFP, _ = load_reference()
    sd = make_state_dict(seed=1)
    ref = build_ref_model(CRFP.CRFP_DSV, sd)
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, t, h, w, fv, seed, stride, crop in CASES:
        lrs, fvs, mks, fv_sp = make_clip(seed=seed, n=1, t=t, h=h, w=w, fv_size=fv)
        t0 = time.time()
        with torch.no_grad():
            y_ref = ref(lrs.clone(), fvs.clone(), mks.clone())
        t1 = time.time()
        y_or = O.crfp_dsv_forward(sd, lrs, fvs, mks)
        d = (y_ref - y_or).abs().max().item()
        print(f"{name}: reference {t1 - t0:.1f} s, oracle {time.time() - t1:.1f} s; ref vs oracle max-abs {d:.3e}; "
              f"out range [{y_ref.min():.3f},{y_ref.max():.3f}]")
        assert d == 0.0, "oracle restatement deviates from the reference"
        fix = {"case": dict(n=1, t=t, h=h, w=w, fv=fv, seed=seed, weight_seed=1), "fv_sp": fv_sp,
               "lrs_sum": float(lrs.double().sum()), "fvs_sum": float(fvs.double().sum())}
        fix.update(compact(y_ref, fv_sp, fv, stride, crop))
        path = os.path.join(out_dir, name + ".pt")
        torch.save(fix, path)
        print(f"   wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
