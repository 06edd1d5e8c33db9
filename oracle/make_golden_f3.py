"""Goldens for SURVEY.md 8(f) rank 3 from the REAL reference (run in the build container only; TEST INFRASTRUCTURE):
  * gaze / scan trajectories of `dataset.reds.fovea_generator` (/root/reference/dataset/reds.py:17-226) for every
    scan method the shipped scripts use, at several (frames, H, W, FV, step) combinations; Rscan with np.random.seed;
  * FV images / masks of that function on a small random clip (tensor branch);
  * `utils.calc_psnr_and_ssim_cuda` (/root/reference/utils.py:165-254) on random image pairs, with masks, both
    reduction modes.
Usage: python oracle/make_golden_f3.py   (writes tests/golden/f3_*.pt)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def main():
    sys.path.insert(0, REF)
    from dataset.reds import fovea_generator  # type: ignore
    import utils as ref_utils  # type: ignore
    out_dir = os.path.join(ROOT, "tests", "golden")
    traj = []
    for method in ("Hscan", "Vscan", "Cscan", "Zscan", "Rscan", "Evenscan", "DemoHscan", "Dscan"):
        for (t, H, W, fv, step) in ((7, 256, 448, 96, 0.1), (15, 256, 256, 128, 0.1), (100, 1440, 2560, 96, 0.01),
                                    (30, 720, 1280, 96, 0.05), (5, 128, 192, 32, 0.2)):
            gt = torch.zeros(t, 3, H, W)
            np.random.seed(1234 + t)
            try:
                _, _, fv_sp = fovea_generator(gt, method=method, step=step, FV_HW=(fv, fv))
            except Exception as e:  # some scans fail on some sizes in the reference itself: record that too
                traj.append(dict(method=method, t=t, H=H, W=W, fv=fv, step=step, error=type(e).__name__))
                continue
            traj.append(dict(method=method, t=t, H=H, W=W, fv=fv, step=step, seed=1234 + t, fv_sp=fv_sp.clone()))
    print(f"{len(traj)} trajectories, {sum('error' in x for x in traj)} reference errors")
    # the tensor branch on a small clip: FV images and masks
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(6, 3, 64, 96, generator=g)
    clips = []
    for method in ("Hscan", "Rscan", "DemoHscan", "Cscan"):
        np.random.seed(77)
        fvs, sps, fv_sp = fovea_generator(gt, method=method, step=0.1, FV_HW=(24, 24))
        clips.append(dict(method=method, fv=24, step=0.1, seed=77, fvs=torch.stack(fvs), sps=torch.stack(sps), fv_sp=fv_sp))
    # metrics
    mets = []
    for (B, C, H, W, seed) in ((1, 3, 64, 96, 1), (2, 3, 40, 56, 2), (1, 1, 33, 47, 3)):
        g = torch.Generator().manual_seed(seed)
        hr = torch.rand(B, C, H, W, generator=g)
        sr = (hr + 0.1 * torch.randn(B, C, H, W, generator=g)).clamp(0, 1)
        mask = torch.zeros(B, 1, H, W)
        mask[:, :, H // 4: H // 2 + 5, W // 3: W // 3 + 17] = 1
        one = torch.ones(B, 1, H, W)
        for name, m in (("rect", mask), ("ones", one)):
            p, s = ref_utils.calc_psnr_and_ssim_cuda(sr, hr, m)
            mets.append(dict(B=B, C=C, H=H, W=W, seed=seed, mask=name, psnr=float(p), ssim=float(s)))
        p, s = ref_utils.calc_psnr_and_ssim_cuda(sr, hr, one, batch_avg=True)
        mets.append(dict(B=B, C=C, H=H, W=W, seed=seed, mask="batch_avg", psnr=p.clone(), ssim=s.clone()))
        # identical images: the reference's finite PSNR
        p, s = ref_utils.calc_psnr_and_ssim_cuda(hr, hr, one)
        mets.append(dict(B=B, C=C, H=H, W=W, seed=seed, mask="identical", psnr=float(p), ssim=float(s)))
    torch.save({"trajectories": traj, "clips": clips, "gt_seed": 3, "metrics": mets}, os.path.join(out_dir, "f3_fovea_metrics.pt"))
    print("wrote", os.path.join(out_dir, "f3_fovea_metrics.pt"), os.path.getsize(os.path.join(out_dir, "f3_fovea_metrics.pt")) / 1e6, "MB")


if __name__ == "__main__":
    main()
