"""`MRCF_simple_v18` of the reference's timing harness — drop-in for `model.CRFP_runtime.MRCF_simple_v18`
(/root/reference/model/CRFP_runtime.py:8364-8682, driven by test_runtime.py:149): same constructor,
`forward(lrs, fvs, warp_size=(H, W))` and state_dict (158 tensors: the DSV modules plus the two-input / bottleneck
residual blocks `forward_resblocks_k{,_}.{conv1,conv2,main.1.0.conv1,main.1.0.conv2}`, CRFP_runtime.py:406-556).

Semantics kept exactly: alignment — flow, warps, DCN, recurrent state — only inside the
top-left `warp_size` (HR pixels) region; `fvs` is an HR image of its own size anchored at the top-left corner and fused
without a mask; after the first frame the residual-block outputs only feed the next frame's `feat_lv*` state (the
level-to-level propagation is commented out in the reference, :8552,8567,8582); the reference's prints, CUDA-event timers
and module-level cached grid are not reproduced.

Round-1 status: node-by-node over the library's operator kernels (the SIMT fp32 conv / DCNv2 / flow_warp / resize entry
points, via the same `_Net` wiring helper as the training forward) — correct, unfused, not tuned; checked against golden
outputs of the real reference class on the CPU through the test kernel set, NOT yet run on a GPU.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import autograd as A
from .model import _build_tree
from .spec import crfp_param_shapes
from .training import _Net, pixel_shuffle_nhwc, pixel_unshuffle_nhwc


def runtime_param_shapes(mid_channels: int = 32):
    """state_dict shapes of CRFP_runtime.MRCF_simple_v18(mid_channels): the DSV shapes for every shared module plus the
    runtime file's residual blocks."""
    C, c = mid_channels, mid_channels // 8
    shapes = OrderedDict((k, v) for k, v in crfp_param_shapes("dsv", mid_channels, False).items()
                         if not k.startswith("forward_resblocks_"))

    def block(name, cin1, cin2, ch):
        for key, shp in ((".conv1", (ch, cin1)), (".conv2", (ch, cin2)), (".main.1.0.conv1", (ch // 2, ch)),
                         (".main.1.0.conv2", (ch, ch // 2))):
            shapes[name + key + ".weight"] = (shp[0], shp[1], 3, 3)
            shapes[name + key + ".bias"] = (shp[0],)

    q3 = 3 * C // 4
    for k in range(3):
        block(f"forward_resblocks_{k}_", q3, q3 // 3, C)      # ResidualBlocksWithInputConv(24, 32): conv2 takes in // 3
        block(f"forward_resblocks_{k}", 2 * C, C, C)           # ..._v2(64, 32): conv2 takes in // 2
    block("forward_resblocks_3_", c, c // 3, c)
    block("forward_resblocks_3", 2 * c, c, c)
    return shapes


def make_runtime_state_dict(seed: int = 21, mid_channels: int = 32, flow_gain: float = 0.05):
    """Seeded random weights under the runtime class's parameter names (same policy as synthetic.make_state_dict)."""
    sd = OrderedDict()
    for key, shape in runtime_param_shapes(mid_channels).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
        if key.endswith(".bias"):
            t = torch.randn(shape, generator=g) * (0.05 if ("dcn.bias" in key or "dcn_offset" in key or "dcn_mask" in key) else 0.02)
        else:
            std = (2.0 / (shape[1] * 9)) ** 0.5
            if ".main.1.0.conv" in key:
                std *= 0.1
            if key.endswith("dcn.weight") or "dcn_offset" in key or "dcn_mask" in key:
                std = 0.05
            t = torch.randn(shape, generator=g) * std
        if key.startswith("spynet.flow.2."):
            t = t * flow_gain
        sd[key] = t.contiguous()
    return sd


class MRCF_simple_v18(nn.Module):
    def __init__(self, device, mid_channels=16, y_only=False, hr_dcn=True, offset_prop=True, split_ratio=3,
                 spynet_pretrained=None, kernels=None):
        super().__init__()
        if mid_channels != 32 or y_only or not hr_dcn or not offset_prop or split_ratio != 3:
            raise L.CrfpError("crfp_b200 implements mid_channels=32, y_only=False, hr_dcn=True, offset_prop=True, "
                              "split_ratio=3 (what test_runtime.py builds)")
        self.device = device
        self.mid_channels, self.last_channels = mid_channels, mid_channels // 8
        self.dg_num, self.dk, self.max_residue_magnitude = 8, 3, 10
        self.y_only, self.hr_dcn, self.offset_prop, self.split_ratio = y_only, hr_dcn, offset_prop, split_ratio
        self.K = kernels or A.CUDA
        _build_tree(self, runtime_param_shapes(mid_channels))

    def init_weights(self, pretrained=None, strict=True):
        if isinstance(pretrained, str):
            saved = {k: v for k, v in torch.load(pretrained, map_location=self.device).items()}
            sd = self.state_dict()
            sd.update(saved)
            self.load_state_dict(sd, strict=strict)
        elif pretrained is not None:
            raise TypeError(f'"pretrained" must be a str or None. But received {type(pretrained)}.')

    # ---- ResidualBlocksWithInputConv{,_v2}.forward (CRFP_runtime.py:464-556)
    @staticmethod
    def _res_blocks(net, name, srcs, feat2=None):
        h, w = srcs[0].shape[1:3]
        if feat2 is None or tuple(feat2.shape[1:3]) == (h, w):   # same size: conv1's output overwrites ALL of conv2's
            x = net.conv(name + ".conv1", srcs, A.ACT_LRELU)
        else:
            feat = net.conv(name + ".conv2", [feat2]).clone()
            feat[:, :h, :w] = net.conv(name + ".conv1", srcs)
            x = F.leaky_relu(feat, 0.1)
        return x + net.conv(name + ".main.1.0.conv2", [net.conv(name + ".main.1.0.conv1", [x], A.ACT_RELU)])

    @torch.no_grad()
    def forward(self, lrs, fvs, warp_size=(1080, 1920)):
        K, C = self.K, self.mid_channels
        net = _Net(self, K)
        WP_h, WP_w = warp_size
        n, t, c, h, w = lrs.shape
        if c != 3 or fvs.dim() != 5 or fvs.shape[:3] != lrs.shape[:3]:
            raise ValueError(f"lrs must be (n,t,3,h,w) and fvs (n,t,3,H,W), got {tuple(lrs.shape)} and {tuple(fvs.shape)}")
        fh, fw = fvs.shape[-2:]
        if fh > 8 * h or fw > 8 * w or WP_h % 8 or WP_w % 8:
            raise ValueError("fvs must fit inside the HR frame and warp_size must be a multiple of 8")
        wh, ww = min(WP_h // 8, h), min(WP_w // 8, w)            # the warp region in LR pixels (slicing clamps)
        lr = K.to_nhwc(lrs.to(torch.float32).reshape(n * t, c, h, w).contiguous())
        fv = K.to_nhwc(fvs.to(torch.float32).reshape(n * t, c, fh, fw).contiguous())
        flows = None
        if t > 1:
            lr5 = lr.view(n, t, h, w, c)[:, :, :wh, :ww]
            flows = net.fnet(lr5[:, 1:].reshape(n * (t - 1), wh, ww, c), lr5[:, :-1].reshape(n * (t - 1), wh, ww, c))
            flows = flows.view(n, t - 1, wh, ww, 2)
        lr_up8 = A.up_bilinear(K, lr, 8).view(n, t, 8 * h, 8 * w, c)
        x_lr = net.conv("encoder_lr.slice1.2", [net.conv("encoder_lr.slice1.0", [lr], A.ACT_LRELU)], A.ACT_LRELU)
        x_lr = x_lr.view(n, t, h, w, -1)
        x_hr = net.conv("encoder_hr.slice1.2", [net.conv("encoder_hr.slice1.0", [torch.cat((fv, fv), dim=-1)], A.ACT_LRELU)],
                        A.ACT_LRELU).view(n, t, fh, fw, -1)
        q3 = 3 * (C // 4)
        outs, S, feats = [], None, [None] * 3
        for i in range(t):
            prop = pixel_shuffle_nhwc(net.conv("upsample.upsample_conv", [x_lr[:, i]]), 2)
            if i > 0:
                flow = flows[:, i - 1]
                flow_lv3 = A.up_bilinear(K, flow, 2, 2.0)
                flow_lv0 = A.up_bilinear(K, flow, 8, 8.0)
                S0 = S
                S0_w = A.flow_warp(K, S0, flow_lv0)
                P_w = net.conv("downsample.downsample_conv", [pixel_unshuffle_nhwc(S0_w, 4)])
                P = net.conv("downsample.downsample_conv", [pixel_unshuffle_nhwc(S0, 4)])
                feats = list(torch.chunk(A.flow_warp(K, torch.cat(feats, dim=-1), flow_lv3), 3, dim=-1))
                offfeat = None
                prop_c = prop[:, :2 * wh, :2 * ww]
                for k in range(3):
                    cur = torch.cat((prop_c, feats[k]), dim=-1)
                    aligned, offfeat = net.dcn_module(f"dcn_{k}", cur, P, P_w, flow_lv3, offfeat, dg=8)
                    y = self._res_blocks(net, f"forward_resblocks_{k}", [cur, aligned], cur)
                    feats[k] = y[..., q3:]
                q = pixel_shuffle_nhwc(net.conv("upsample_post.upsample_conv", [prop], A.ACT_LRELU), 4)
                qc = q[:, :8 * wh, :8 * ww]
                aligned3, _ = net.dcn_module("dcn_3", qc, S0, S0_w, flow_lv0, offfeat, dg=1, repeat=True, pixelshuffle=True)
                S = self._res_blocks(net, "forward_resblocks_3", [qc, aligned3], q)
            else:
                for k in range(3):
                    y = self._res_blocks(net, f"forward_resblocks_{k}_", [prop])
                    feats[k] = y[:, :2 * wh, :2 * ww, q3:]
                    prop = y[..., :q3]
                q = pixel_shuffle_nhwc(net.conv("upsample_post.upsample_conv", [prop], A.ACT_LRELU), 4)
                S = self._res_blocks(net, "forward_resblocks_3_", [q])
            Fz = net.conv("conv_tttf", [S[:, :fh, :fw], x_hr[:, i]])
            S = S.clone()
            S[:, :fh, :fw] = Fz
            S = F.leaky_relu(S, 0.1)
            outs.append((net.conv("conv_last", [S]) + lr_up8[:, i]).permute(0, 3, 1, 2))
            S = S[:, :8 * wh, :8 * ww].contiguous()
        return torch.stack(outs, dim=1)
