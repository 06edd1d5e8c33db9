"""PSNR / SSIM exactly as the reference evaluates them (`utils.py:165-254`: `psnr_cuda`, `ssim_cuda`,
`calc_psnr_and_ssim_cuda`): images in [0, 1], PSNR = `-20 * log10(sqrt(mse))` with optional mask weighting and the
reference's finite value for identical images; SSIM with an 11-tap Gaussian window (sigma 1.5, zero padding),
C1 = 0.01^2, C2 = 0.03^2, the same mask weighting.

`psnr` is a handful of torch ops (plumbing, used for the BASELINE.json "within 0.05 dB" criterion on any device).
`calc_psnr_and_ssim_cuda` is the evaluation loop's per-frame call (test_video.py:361-374 issues it four times per frame
at 1080p): ONE fused kernel pass over both images (`crfp_psnr_ssim`) instead of five depthwise convolutions and ~15
pointwise kernels, reduced deterministically from per-CTA partial sums."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L


def psnr(img1: torch.Tensor, img2: torch.Tensor, mask: torch.Tensor | None = None, batch_avg: bool = False):
    """img1, img2: (B, C, H, W) in [0, 1].  `batch_avg=True`: per-image PSNR (B,); otherwise one value over the
    pixels selected by `mask` ((B, 1, H, W) or broadcastable; all pixels when None)."""
    b, c, h, w = img1.shape
    d2 = (img1.to(torch.float32) - img2.to(torch.float32)) ** 2
    if batch_avg:
        mse = d2.reshape(b, -1).mean(1)
        floor = -20.0 * torch.log10(torch.sqrt(torch.tensor((1 / 255.0) ** 2 / (c * h * w), device=mse.device)))
        return torch.where(mse == 0, floor, -20.0 * torch.log10(torch.sqrt(mse)))
    if mask is None:
        mask = torch.ones(b, 1, h, w, device=img1.device)
    m = mask.to(torch.float32)
    mse = (d2 * m).sum() / (m.sum() * c)
    if float(mse) == 0.0:
        return -20.0 * torch.log10(torch.sqrt(torch.tensor((1 / 255.0) ** 2 / (b * c * h * w))))
    return -20.0 * torch.log10(torch.sqrt(mse))


def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """The reference's 1-D window (utils.py:186-188): float64 exponentials -> fp32 tensor -> normalised in fp32."""
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def _to_unit_range(sr, hr):
    """Range detection of calc_psnr_and_ssim_cuda (utils.py:243-250): [0,255] and [-1,1] inputs are mapped to [0,1]."""
    span = float(hr.max() - hr.min())
    if span > 2:
        return sr / 255.0, hr / 255.0
    if span > 1:
        return (sr + 1.0) / 2.0, (hr + 1.0) / 2.0
    return sr, hr


def calc_psnr_and_ssim_cuda(sr: torch.Tensor, hr: torch.Tensor, mask: torch.Tensor | None = None, batch_avg: bool = False):
    """Drop-in for `utils.calc_psnr_and_ssim_cuda(sr, hr, mask, batch_avg=...)`: CUDA tensors (B, C, H, W); `mask`
    (B, 1, H, W) float / bool / uint8 or None (all ones).  Returns (psnr, ssim) as the reference does: CPU scalars in the
    masked mode, (B,) device tensors with batch_avg=True (which ignores the mask, like the reference)."""
    if not (sr.is_cuda and hr.is_cuda):
        raise L.CrfpError("calc_psnr_and_ssim_cuda needs CUDA tensors: crfp_b200 has no CPU fallback")
    if sr.shape != hr.shape or sr.dim() != 4:
        raise ValueError("sr and hr must be (B, C, H, W) tensors of the same shape")
    sr, hr = _to_unit_range(sr.to(torch.float32), hr.to(torch.float32))
    sr, hr = sr.contiguous(), hr.contiguous()
    b, c, h, w = sr.shape
    mf = mu = None
    if mask is not None and not batch_avg:
        m = mask.expand(b, 1, h, w) if tuple(mask.shape) != (b, 1, h, w) else mask
        if m.dtype in (torch.bool, torch.uint8):
            mu = m.contiguous().view(torch.uint8)
        else:
            mf = m.to(torch.float32).contiguous()
    lib = L.lib()
    tx, ty = C.c_int32(), C.c_int32()
    L.check(lib.crfp_psnr_ssim_tiles(h, w, C.byref(tx), C.byref(ty)), "psnr_ssim_tiles")
    tiles = tx.value * ty.value
    partial = torch.empty(b * c, tiles, 3, device=sr.device, dtype=torch.float32)
    win = (C.c_float * 11)(*gaussian_window().tolist())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.crfp_psnr_ssim(b, c, h, w, sr.data_ptr(), hr.data_ptr(), None if mf is None else mf.data_ptr(),
                               None if mu is None else mu.data_ptr(), win, partial.data_ptr(), st), "psnr_ssim")
    p64 = partial.to(torch.float64)
    if batch_avg:
        per = p64.view(b, c * tiles, 3).sum(1)                      # mask is all ones here: sum m = C*H*W per image
        mse = (per[:, 1] / per[:, 2]).to(torch.float32)
        floor = -20.0 * torch.log10(torch.sqrt(torch.tensor((1 / 255.0) ** 2 / (c * h * w), device=mse.device)))
        ps = torch.where(mse == 0, floor, -20.0 * torch.log10(torch.sqrt(mse)))
        return ps, (per[:, 0] / per[:, 2]).to(torch.float32)
    tot = p64.sum((0, 1)).cpu()
    # sum m was accumulated once per channel plane: tot[2] == mask.sum() * C, the reference's denominator
    mse = float(tot[1] / tot[2])
    if mse == 0.0:
        ps = -20.0 * torch.log10(torch.sqrt(torch.tensor((1 / 255.0) ** 2 / (b * c * h * w))))
    else:
        ps = -20.0 * torch.log10(torch.sqrt(torch.tensor(mse, dtype=torch.float32)))
    return ps, torch.tensor(float(tot[0] / tot[2]), dtype=torch.float32)
