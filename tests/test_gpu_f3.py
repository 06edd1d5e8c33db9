"""GPU tests (-m gpu) of SURVEY.md 8(f) rank 3: the device-side `fovea_generator` and the fused PSNR / SSIM kernel against
values produced by the REAL reference (tests/golden/f3_fovea_metrics.pt, oracle/make_golden_f3.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fix(golden_dir):
    return torch.load(os.path.join(golden_dir, "f3_fovea_metrics.pt"))


def test_device_fovea_generator_equals_the_reference(fix):
    from crfp_b200.fovea import fovea_generator
    gt = torch.rand(6, 3, 64, 96, generator=torch.Generator().manual_seed(fix["gt_seed"]))
    for c in fix["clips"]:
        np.random.seed(c["seed"])
        fvs, mks, fv_sp = fovea_generator(gt.cuda(), method=c["method"], step=c["step"], fv_hw=(c["fv"], c["fv"]))
        torch.cuda.synchronize()
        assert torch.equal(fv_sp, c["fv_sp"]), c["method"]                # exact integer fovea coordinates
        assert torch.equal(fvs.cpu(), c["fvs"]), c["method"]              # GT * mask, bit for bit
        assert mks.dtype == torch.bool and torch.equal(mks.cpu().float(), c["sps"][:, :1]), c["method"]
    with pytest.raises(Exception):
        fovea_generator(gt, method="Hscan")                               # CPU tensors: no fallback


def test_generated_clip_feeds_the_model(fix):
    """fovea_generator output -> CRFP_DSV.forward, and the same fovea through forward_patch (patch + coords)."""
    from crfp_b200 import CRFP_DSV
    from crfp_b200.fovea import fovea_generator
    from crfp_b200.synthetic import make_state_dict
    g = torch.Generator().manual_seed(4)
    t, h, w, fv = 4, 16, 24, 48
    gt = torch.rand(t, 3, 8 * h, 8 * w, generator=g).cuda()
    lrs = torch.rand(1, t, 3, h, w, generator=g).cuda()
    fvs, mks, fv_sp = fovea_generator(gt, method="Cscan", step=0.1, fv_hw=(fv, fv))
    m = CRFP_DSV("cuda", mid_channels=32).eval()
    m.load_state_dict(make_state_dict(seed=1), strict=True)
    m.cuda()
    a = m(lrs, fvs[None], mks[None])
    patch = torch.stack([gt[i, :, fv_sp[i, 0]:fv_sp[i, 0] + fv, fv_sp[i, 1]:fv_sp[i, 1] + fv] for i in range(t)])[None]
    b = m.forward_patch(lrs, patch.contiguous(), fv_sp[None])
    assert torch.equal(a, b)


def test_fused_psnr_ssim_matches_the_reference(fix):
    from crfp_b200.metrics import calc_psnr_and_ssim_cuda
    for met in fix["metrics"]:
        B, C, H, W = met["B"], met["C"], met["H"], met["W"]
        g = torch.Generator().manual_seed(met["seed"])
        hr = torch.rand(B, C, H, W, generator=g)
        sr = (hr + 0.1 * torch.randn(B, C, H, W, generator=g)).clamp(0, 1)
        mask = torch.zeros(B, 1, H, W)
        mask[:, :, H // 4: H // 2 + 5, W // 3: W // 3 + 17] = 1
        one = torch.ones(B, 1, H, W)
        kind = met["mask"]
        if kind == "batch_avg":
            p, s = calc_psnr_and_ssim_cuda(sr.cuda(), hr.cuda(), one.cuda(), batch_avg=True)
            assert (p.cpu() - met["psnr"]).abs().max().item() < 1e-3 and (s.cpu() - met["ssim"]).abs().max().item() < 1e-4
            continue
        if kind == "identical":
            p, s = calc_psnr_and_ssim_cuda(hr.cuda(), hr.cuda(), one.cuda())
        else:
            m = mask if kind == "rect" else one
            p, s = calc_psnr_and_ssim_cuda(sr.cuda(), hr.cuda(), m.cuda())
            pb, sb = calc_psnr_and_ssim_cuda(sr.cuda(), hr.cuda(), m.bool().cuda())     # bool masks take the byte path
            assert abs(float(pb) - float(p)) < 1e-6 and abs(float(sb) - float(s)) < 1e-6
        assert abs(float(p) - met["psnr"]) < 1e-3, (met, float(p))
        assert abs(float(s) - met["ssim"]) < 1e-4, (met, float(s))
    # the other input ranges the reference auto-detects: [0, 255] and [-1, 1]
    g = torch.Generator().manual_seed(9)
    hr = torch.rand(1, 3, 48, 64, generator=g)
    sr = (hr + 0.05 * torch.randn(1, 3, 48, 64, generator=g)).clamp(0, 1)
    p0, s0 = calc_psnr_and_ssim_cuda(sr.cuda(), hr.cuda())
    p1, s1 = calc_psnr_and_ssim_cuda((sr * 255).cuda(), (hr * 255).cuda())
    p2, s2 = calc_psnr_and_ssim_cuda((sr * 2 - 1).cuda(), (hr * 2 - 1).cuda())
    assert abs(float(p0) - float(p1)) < 1e-3 and abs(float(p0) - float(p2)) < 1e-3
    assert abs(float(s0) - float(s1)) < 1e-4 and abs(float(s0) - float(s2)) < 1e-4
