"""PSNR exactly as the reference evaluates it (`utils.py:165-184`, `psnr_cuda`): images in [0, 1],
`-20 * log10(sqrt(mse))`, optional foveal mask weighting, the reference's finite value for identical images.
Plain torch ops on whatever device the tensors live on (plumbing, not a hot path); used for the BASELINE.json
"within 0.05 dB" criterion between precisions."""
from __future__ import annotations

import torch


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
