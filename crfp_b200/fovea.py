"""Gaze / scan trajectories and the foveated inputs they produce — the data loader's `fovea_generator`
(/root/reference/dataset/reds.py:17-226) with the per-pixel work on the GPU.

The reference builds, per frame, a full-size mask and `GT * mask` with Python slicing on the CPU.  Here the trajectory
(a handful of integers per frame: host arithmetic, restated below scan by scan) is separated from the pixels: the
rectangles go to the device as int32 and ONE kernel (`crfp_fovea_from_gt`) writes `fvs = GT * mask` and `mks` for the
whole clip — or nothing is materialised at all and `CRFP_DSV.forward_patch(lrs, patch, coords)` pastes the FV x FV
patches into persistent buffers.

`scan_positions` returns exactly the `fv_sp` of the reference for the scans its scripts use:
  Hscan / Vscan     a line through the centre, `step` of the frame per frame            (reds.py:45-56, 69-72)
  Cscan             boustrophedon rows on a sqrt(t) x sqrt(t) grid                       (reds.py:40-44, 77-89)
  Zscan             diagonal zig-zag over the same grid                                  (reds.py:90-113)
  Rscan             Gaussian gaze around the centre, sigma = 5 % of the frame            (reds.py:114-118)
  Evenscan          raster over the FV-sized cells starting at cell 20                   (reds.py:159-169)
  DemoHscan         a vertical edge sweeping left and right in 8-pixel steps; the mask is the open quadrant below /
                    right of the point, not an FV rectangle                              (reds.py:170-186, 197-198)
  anything else     the main diagonal                                                    (reds.py:57-63, 187-188)
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

LINE_SCANS = ("Hscan", "Vscan")
GRID_SCANS = ("Cscan", "Zscan")


def _percent_range(n_frames, step, limit_a, limit_b, n_cells):
    """Start / stop / step of the scan in whole percents of the frame, after the reference's step shrinking:
    the scan starts at 10 % and must stay below the given limits after `n_cells` steps."""
    start = 0.1
    if start + n_cells * step > limit_a or start + n_cells * step > limit_b:
        step = min((limit_a - start) / n_cells, (limit_b - start) / n_cells)
    return int(start * 100), int(step * 100)


def scan_positions(n_frames: int, gt_h: int, gt_w: int, method: str = "Rscan", step: float = 0.1, fv_hw=(32, 32),
                   rng=None) -> torch.Tensor:
    """Top-left corners `[y, x]` of the fovea for every frame, int64 (n_frames, 2) — `fv_sp` of the reference.
    `rng`: numpy RandomState / Generator-like with `.normal` for Rscan (default: the global numpy state, as the
    reference uses)."""
    fv_h, fv_w = fv_hw
    centre_h = (gt_h * 0.5 - fv_h // 2) / gt_h
    centre_w = (gt_w * 0.5 - fv_w // 2) / gt_w
    end_h = (gt_h * 0.9 - fv_h) / gt_h
    end_w = (gt_w * 0.9 - fv_w) / gt_w
    side = math.ceil(math.sqrt(n_frames))
    pos = []
    if method in GRID_SCANS:
        sp, st = _percent_range(n_frames, step, end_h, end_w, side)
        ep = int(sp + math.ceil(math.sqrt(n_frames) - 1) * st)
        v = h = sp
        dv = dh = st
        for _ in range(n_frames):
            pos.append([int((v / 100) * gt_h), int((h / 100) * gt_w)])
            if method == "Cscan":
                if (h == ep and dh > 0) or (h == sp and dh < 0):
                    dh = -dh
                    v += dv
                else:
                    h += dh
            else:  # Zscan: bounce off the four sides of the grid
                if h == ep and dv < 0:
                    dv = -dv
                    v += dv
                    dh = -abs(dh)
                elif v == sp and dh > 0:
                    h += dh
                    dh = -dh
                    dv = abs(dv)
                elif v == ep and dh < 0:
                    dh = -dh
                    h += dh
                    dv = -abs(dv)
                elif h == sp and dv > 0:
                    v += dv
                    dv = -dv
                    dh = abs(dh)
                else:
                    h += dh
                    v += dv
    elif method == "Rscan":
        r = np.random if rng is None else rng
        ys = r.normal(centre_h, 0.05, n_frames).clip(0, end_h)
        xs = r.normal(centre_w, 0.05, n_frames).clip(0, end_w)
        pos = [[int(a * gt_h), int(b * gt_w)] for a, b in zip(ys, xs)]
    elif method == "Evenscan":
        cells_h, cells_w = gt_h // fv_h, gt_w // fv_w
        pitch_h, pitch_w = gt_h / cells_h, gt_w / cells_w
        for i in range(20, 20 + n_frames):
            cx, cy = i % cells_w, (i // cells_w) % cells_h
            pos.append([int((1 + cy) * pitch_h - (pitch_h + fv_h) / 2), int((1 + cx) * pitch_w - (pitch_w + fv_w) / 2)])
    elif method == "DemoHscan":
        x, d = gt_w - 8, -1
        for _ in range(n_frames):
            pos.append([0, x])
            x += d * 8
            if x < 0 or x >= gt_w:
                d = -d
                x += d * 8
    else:
        if method == "Hscan":
            sp, st = _percent_range(n_frames, step, end_w, end_w, n_frames)
        elif method == "Vscan":
            sp, st = _percent_range(n_frames, step, end_h, end_h, n_frames)
        else:
            sp, st = _percent_range(n_frames, step, end_h, end_w, n_frames)
        ep = int(sp + n_frames * st)
        if st == 0:
            raise ValueError("arg 3 must not be zero")   # range(sp, ep, 0): the reference raises here as well
        for v in range(sp, ep, st):
            if method == "Hscan":
                pos.append([int(centre_h * gt_h), int((v / 100) * gt_w)])
            elif method == "Vscan":
                pos.append([int((v / 100) * gt_h), int(centre_w * gt_w)])
            else:
                pos.append([int((v / 100) * gt_h), int((v / 100) * gt_w)])
    return torch.tensor(pos)


def rects_from_positions(fv_sp: torch.Tensor, gt_h: int, gt_w: int, fv_hw, method: str) -> torch.Tensor:
    """int32 (frames, 4) = [y0, x0, y1, x1) clipped to the frame; DemoHscan marks everything below / right of the point
    (reds.py:197-198), every other scan the FV_H x FV_W rectangle (reds.py:199-200; Python slicing clips at the border
    and treats negative starts as offsets from the end, which only Evenscan can produce — rejected here)."""
    fv_h, fv_w = fv_hw
    p = fv_sp.to(torch.int64)
    if int(p.min()) < 0:
        raise ValueError("negative fovea position (the reference's slicing would wrap around)")
    y0, x0 = p[:, 0].clamp(max=gt_h), p[:, 1].clamp(max=gt_w)
    if method == "DemoHscan":
        y1, x1 = torch.full_like(y0, gt_h), torch.full_like(x0, gt_w)
    else:
        y1, x1 = (y0 + fv_h).clamp(max=gt_h), (x0 + fv_w).clamp(max=gt_w)
    return torch.stack([y0, x0, y1, x1], -1).to(torch.int32)


def fovea_generator(gt_imgs: torch.Tensor, method: str = "Rscan", step: float = 0.1, fv_hw=(32, 32), rng=None):
    """Device-side `fovea_generator` for a CUDA clip `gt_imgs` (t, 3, H, W) fp32: returns (fvs (t,3,H,W), mks (t,1,H,W)
    bool, fv_sp (t,2) int64 on the CPU) with `fvs = GT * mask` — one kernel for the whole clip.  The reference's tensor
    branch returns the mask replicated over the 3 channels; only one plane is kept here (the model takes (n,t,1,H,W))."""
    if not (isinstance(gt_imgs, torch.Tensor) and gt_imgs.is_cuda and gt_imgs.dim() == 4):
        raise L.CrfpError("gt_imgs must be a CUDA tensor (t, C, H, W): crfp_b200 has no CPU fallback")
    t, c, h, w = gt_imgs.shape
    fv_sp = scan_positions(t, h, w, method, step, fv_hw, rng)
    if fv_sp.shape[0] < t:
        raise IndexError(f"the {method} scan yields {fv_sp.shape[0]} positions for {t} frames (the reference fails here too)")
    rects = rects_from_positions(fv_sp[:t], h, w, fv_hw, method).to(gt_imgs.device)
    gt = gt_imgs.to(torch.float32).contiguous()
    fvs = torch.empty_like(gt)
    mks = torch.empty(t, 1, h, w, device=gt.device, dtype=torch.uint8)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().crfp_fovea_from_gt(gt.data_ptr(), rects.data_ptr(), t, c, h, w, fvs.data_ptr(), mks.data_ptr(), st),
            "fovea_from_gt")
    return fvs, mks.view(torch.bool), fv_sp
