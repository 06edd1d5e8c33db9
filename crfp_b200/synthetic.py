"""Seeded synthetic weights and clips for the CRFP hot path.

There are no datasets or checkpoints in this environment, so every test and
benchmark runs on synthetic inputs of the reference's shapes:

* clips follow the data-loader contract of the reference
  (/root/reference/dataset/reds.py:190-226): `lrs (n,t,3,h,w)` in [0,1],
  `fvs (n,t,3,8h,8w)` zero outside an FV x FV rectangle, `mks (n,t,1,8h,8w)`
  bool rectangle mask, `fv_sp (n,t,2)` int64 top-left `[y, x]`; the gaze walk
  is the Gaussian one of /root/reference/test_video.py:149,309-310,337-338.
* weights are NOT the reference's default init: that init zeroes the DCN
  offset/mask heads and makes `dcn.weight` an identity
  (/root/reference/model/CRFP.py:354-370), which would leave the deformable
  path unexercised.  Every tensor is drawn independently of
  module construction order, keyed by its name, so the reference model, the
  oracle and the CUDA path can all be loaded with the very same numbers.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .spec import crfp_param_shapes


def _key_seed(seed: int, key: str) -> int:
    return (seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1)


def make_state_dict(seed: int = 1, mid_channels: int = 32, y_only: bool = False,
                    flow_gain: float = 0.05, variant: str = "dsv") -> "OrderedDict[str, torch.Tensor]":
    """Random fp32 state_dict with the reference's 118 keys and shapes.

    conv weights ~ N(0, (g*sqrt(2/fan_in))^2); residual-block convs use g=0.1
    (as the reference's init does, CRFP.py:460-470); DCN heads and DCN weights
    ~ N(0, 0.05^2); the last FNet conv is scaled by
    `flow_gain` so that tanh(.)*256 gives flows of a few LR pixels.
    """
    sd = OrderedDict()
    for key, shape in crfp_param_shapes(variant, mid_channels, y_only).items():
        g = torch.Generator(device="cpu")
        g.manual_seed(_key_seed(seed, key))
        if key.endswith(".bias"):
            t = torch.randn(shape, generator=g) * 0.02
            if key.endswith("dcn.bias") or "dcn_offset" in key or "dcn_mask" in key:
                t = torch.randn(shape, generator=g) * 0.05
            if key == "spynet.flow.2.bias":
                t = t * flow_gain
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            std = (2.0 / fan_in) ** 0.5
            if ".main.2.0.conv" in key:
                std *= 0.1
            if key.endswith("dcn.weight") or "dcn_offset" in key or "dcn_mask" in key:
                std = 0.05
            t = torch.randn(shape, generator=g) * std
            if key == "spynet.flow.2.weight":
                t = t * flow_gain
        sd[key] = t.contiguous()
    return sd


def fovea_rect(fv_sp: torch.Tensor, fv_size: int, H: int, W: int) -> torch.Tensor:
    """Bool mask (n,t,1,H,W) that is True on [y:y+FV, x:x+FV]
    (/root/reference/dataset/reds.py:196-201)."""
    n, t, _ = fv_sp.shape
    mks = torch.zeros(n, t, 1, H, W, dtype=torch.bool)
    for b in range(n):
        for i in range(t):
            y, x = int(fv_sp[b, i, 0]), int(fv_sp[b, i, 1])
            mks[b, i, 0, y:y + fv_size, x:x + fv_size] = True
    return mks


def make_clip(seed: int, n: int, t: int, h: int, w: int, fv_size: int = 96,
              smooth: bool = True):
    """Seeded synthetic clip: returns (lrs, fvs, mks, fv_sp) on the CPU."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    H, W = 8 * h, 8 * w
    fv = min(fv_size, H, W)
    if smooth:
        ch, cw = max(h // 8, 2), max(w // 8, 2)
        base = torch.rand(n, 1, 3, ch, cw, generator=g)
        drift = torch.rand(n, t, 3, ch, cw, generator=g)
        coarse = 0.7 * base + 0.3 * drift
        lrs = F.interpolate(coarse.reshape(n * t, 3, ch, cw), size=(h, w), mode="bicubic",
                            align_corners=False).reshape(n, t, 3, h, w)
        lrs = (lrs + 0.05 * torch.rand(n, t, 3, h, w, generator=g)).clamp_(0.0, 1.0)
    else:
        lrs = torch.rand(n, t, 3, h, w, generator=g)
    gy = torch.randn(n, t, generator=g) * 50.0 + H / 2
    gx = torch.randn(n, t, generator=g) * 50.0 + W / 2
    y0 = (gy.floor().long() - fv // 2).clamp_(0, H - fv)
    x0 = (gx.floor().long() - fv // 2).clamp_(0, W - fv)
    fv_sp = torch.stack([y0, x0], dim=-1)
    mks = fovea_rect(fv_sp, fv, H, W)
    fvs = torch.rand(n, t, 3, H, W, generator=g) * mks
    return lrs.contiguous(), fvs.contiguous(), mks, fv_sp
