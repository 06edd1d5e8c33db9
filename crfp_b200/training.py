"""Differentiable forward of CRFP_DSV for the training step (BASELINE.json configs[4]).

Same computation as `CRFP_DSV.forward` (/root/reference/model/CRFP.py:1510-1686) — which the inference path runs as
fused whole-frame kernels — but expressed node by node over `crfp_b200.autograd` so that torch.autograd can tape the
t-frame recurrence (BPTT).  Every conv / DCNv2 / flow_warp / resize / avg-pool node is a hand-written kernel pair
(forward + backward) behind the C ABI; activations are dense fp32 NHWC tensors.

Round-1 status: the tensor plumbing between nodes (torch.cat / slicing / pixel-(un)shuffle views) and the cheap
pointwise glue (tanh, sigmoid, residual adds, the fovea blend) still run as ATen ops — they are the next fusion
targets; the reference's named training bottlenecks (conv, DCN, flow_warp forward + backward) do not.
"""
from __future__ import annotations

import torch

from . import autograd as A


# ---- NHWC views of the reference's layout ops
def pixel_shuffle_nhwc(x, r):
    """F.pixel_shuffle(r) (model/CRFP.py:187-193) on NHWC: channel o*r*r + dy*r + dx -> pixel (y*r+dy, x*r+dx), ch o."""
    n, h, w, c = x.shape
    o = c // (r * r)
    return x.reshape(n, h, w, o, r, r).permute(0, 1, 4, 2, 5, 3).reshape(n, h * r, w * r, o)


def pixel_unshuffle_nhwc(x, r):
    """pixel_unshuffle (model/CRFP.py:28-42) on NHWC: (n,h*r,w*r,c) -> (n,h,w,c*r*r), channel c*r*r + dy*r + dx."""
    n, hr, wr, c = x.shape
    h, w = hr // r, wr // r
    return x.reshape(n, h, r, w, r, c).permute(0, 1, 3, 5, 2, 4).reshape(n, h, w, c * r * r)


class _Net:
    """Functional view over the model's parameters (reference names, model/CRFP.py state_dict)."""

    def __init__(self, model, K, defer_wgrad=False):
        self.p = dict(model.named_parameters())
        # deferred weight gradients (Trainer): one launch per layer over all frames at the end of the backward pass
        self.defer = A.WgradDeferral(K) if defer_wgrad else None
        self.K = K
        self.C = model.mid_channels
        self.max_mag = float(model.max_residue_magnitude)
        # per-step caches (this object lives for ONE forward/backward, during which the weights are constant): packed
        # weight layouts per layer, and the concatenated head weights (one torch.cat node per step, not per frame)
        self._cache = {}
        self._cat = {}
        import os
        self.fnet_fp32 = os.environ.get("CRFP_TRAIN_TC_FNET", "0") != "1"

    def _layer_cache(self, key, srcs):
        d = self._cache.setdefault((key, tuple(s.shape[-1] for s in srcs)), {})
        # The flow network stays on the fp32 SIMT kernels unless CRFP_TRAIN_TC_FNET=1: flow = 256 * tanh(.) positions every
        # warp and DCN sample and its gradients are the most ill-conditioned of the model.  Measured on the B200 with the
        # flow network on the tensor-core convs as well: worst per-tensor gradient error 5.1e-2 (spynet.encoder3.2.weight,
        # LR 8x16 case) instead of 2.1e-2, for 1.7 ms of a 37.9 ms V7 step.
        if isinstance(key, str) and key.startswith("spynet.") and self.fnet_fp32:
            d["no_tc"] = True
        return d

    def conv(self, name, srcs, act=A.ACT_NONE):
        srcs = list(srcs)
        cache = self._layer_cache(name, srcs)
        wt, b = self.p[name + ".weight"], self.p[name + ".bias"]
        if self.defer is not None and "defer" not in cache:
            cache["defer"], cache["wparams"], cache["bparams"] = self.defer, [(wt, 0, wt.shape[0])], [(b, 0, b.shape[0])]
        return A.conv3x3(self.K, wt, b, srcs, act, cache)

    def conv_cat(self, names, srcs, act=A.ACT_NONE):
        """one conv over the concatenated output channels of several reference convs sharing an input"""
        key = tuple(names)
        if key not in self._cat:
            self._cat[key] = (torch.cat([self.p[n + ".weight"] for n in names], dim=0),
                              torch.cat([self.p[n + ".bias"] for n in names], dim=0))
        w, b = self._cat[key]
        srcs = list(srcs)
        cache = self._layer_cache(key, srcs)
        if self.defer is not None and "defer" not in cache:
            wp, bp, lo = [], [], 0
            for n_ in names:
                co = self.p[n_ + ".weight"].shape[0]
                wp.append((self.p[n_ + ".weight"], lo, lo + co))
                bp.append((self.p[n_ + ".bias"], lo, lo + co))
                lo += co
            cache["defer"], cache["wparams"], cache["bparams"] = self.defer, wp, bp
        return A.conv3x3(self.K, w, b, srcs, act, cache)

    # ---- ResidualBlocksWithInputConv (model/CRFP.py:433-552): conv+LReLU, then x + conv2(relu(conv1(x)))
    def res_blocks(self, name, srcs):
        x = self.conv(name + ".main.0", srcs, A.ACT_LRELU)
        y = self.conv(name + ".main.2.0.conv2", [self.conv(name + ".main.2.0.conv1", [x], A.ACT_RELU)])
        return x + y

    # ---- DCN_module.forward (model/CRFP.py:324-352)
    def dcn_module(self, name, cur_x, pre_x, pre_x_aligned, flow, pre_offset, dg, repeat=False, pixelshuffle=False):
        z = self.conv(name + ".dcn_block.0", [cur_x, pre_x_aligned, flow], A.ACT_LRELU)
        z = self.conv(name + ".dcn_block.2", [z], A.ACT_LRELU)
        if pre_offset is not None:
            if pixelshuffle:
                pre_offset = pixel_shuffle_nhwc(self.conv(name + ".upsample.upsample_conv", [pre_offset]), 4) * 2.0
            z = self.conv(name + ".conv_fuse", [z, pre_offset], A.ACT_LRELU)
        heads = self.conv_cat([name + ".dcn_offset", name + ".dcn_mask"], [z])     # offset ++ mask, one launch
        # 10 * tanh + flow / sigmoid (+ the 9-fold repeat of the HR module's single pair): one kernel forward, one backward
        offset, mask = A.dcn_heads_act(self.K, heads, flow, 9 * dg, repeat, self.max_mag)
        dcache = self._cache.setdefault((name + ".dcn", dg), {})
        dcache["hint"] = flow          # sampling-window hint of the tensor-core align op (values do not depend on it)
        out = A.dcn_v2(self.K, pre_x, offset, mask, self.p[name + ".dcn.weight"], self.p[name + ".dcn.bias"], dg, dcache)
        return out, z

    # ---- FNet.forward (model/CRFP.py:797-814) on pairs (x1, x2) NHWC 3-channel
    def fnet(self, x1, x2):
        K = self.K
        h, w = x1.shape[1:3]
        out = None
        srcs = [x1, x2]
        for name in ("encoder1", "encoder2", "encoder3"):
            out = self.conv(f"spynet.{name}.0", srcs, A.ACT_RELU)
            out = self.conv(f"spynet.{name}.2", [out], A.ACT_RELU)
            out = A.avgpool2(K, out)
            srcs = [out]
        for name in ("decoder1", "decoder2", "decoder3"):
            out = self.conv(f"spynet.{name}.0", [out], A.ACT_RELU)
            out = self.conv(f"spynet.{name}.2", [out], A.ACT_RELU)
            out = A.up_bilinear(K, out, 2)
        out = self.conv("spynet.flow.0", [out], A.ACT_RELU)
        out = torch.tanh(self.conv("spynet.flow.2", [out])) * 256.0
        return A.resize_to(K, out, h, w)

    # ---- one iteration of the t-loop (model/CRFP.py:1555-1684)
    def frame_step(self, state, x_lr_cur, x_hr_cur, mkf, lr_up8, flow):
        K, C = self.K, self.C
        q3 = 3 * (C // 4)
        n, h, w, _ = x_lr_cur.shape
        prop = pixel_shuffle_nhwc(self.conv("upsample.upsample_conv", [x_lr_cur]), 2)          # 24 ch @L1
        if state is not None:
            S0, feat_mix = state
            flow_lv3 = A.up_bilinear(K, flow, 2, 2.0)
            flow_lv0 = A.up_bilinear(K, flow, 8, 8.0)
            P = self.conv("downsample.downsample_conv", [pixel_unshuffle_nhwc(S0, 4)])
            P_w = A.flow_warp(K, P, flow_lv3)
            S0_w = A.flow_warp(K, S0, flow_lv0)
            feats = list(torch.chunk(A.flow_warp(K, feat_mix, flow_lv3), 3, dim=-1))
            offfeat = None
            for k in range(3):
                cur = torch.cat((prop, feats[k]), dim=-1)
                aligned, offfeat = self.dcn_module(f"dcn_{k}", cur, P, P_w, flow_lv3, offfeat, dg=8)
                y = self.res_blocks(f"forward_resblocks_{k}", [cur, aligned])
                prop, feats[k] = y[..., :q3], y[..., q3:]
            q = pixel_shuffle_nhwc(self.conv("upsample_post.upsample_conv", [prop], A.ACT_LRELU), 4)   # lrelu o shuffle
            aligned3, _ = self.dcn_module("dcn_3", q, S0, S0_w, flow_lv0, offfeat, dg=1, repeat=True, pixelshuffle=True)
            S = self.res_blocks("forward_resblocks_3", [q, aligned3])
        else:
            z_l1 = x_lr_cur.new_zeros(n, 2 * h, 2 * w, C)
            z_f = x_lr_cur.new_zeros(n, 2 * h, 2 * w, C // 4)
            z_hr = x_lr_cur.new_zeros(n, 8 * h, 8 * w, C // 8)
            feats = [None] * 3
            for k in range(3):
                y = self.res_blocks(f"forward_resblocks_{k}", [prop, z_l1, z_f])
                prop, feats[k] = y[..., :q3], y[..., q3:]
            q = pixel_shuffle_nhwc(self.conv("upsample_post.upsample_conv", [prop], A.ACT_LRELU), 4)
            S = self.res_blocks("forward_resblocks_3", [q, z_hr])
        Fz = self.conv("conv_tttf", [S, x_hr_cur])
        S = A.fovea_blend(self.K, Fz, S, mkf)          # lrelu(m * Fz + (1 - m) * S, 0.1), one kernel each way
        out = self.conv("conv_last", [S]) + lr_up8
        return out, (S, torch.cat(feats, dim=-1))


def forward_train(model, lrs, fvs, mks, K=None, defer_wgrad=False):
    """CRFP_DSV.forward(lrs, fvs, mks) with autograd: returns (n,t,3,8h,8w) carrying grad to every parameter that
    requires it.  `K` is the kernel set (default: the CUDA library)."""
    K = K or A.CUDA
    net = _Net(model, K, defer_wgrad)
    n, t, c, h, w = lrs.shape
    lrs = lrs.to(torch.float32)
    lr = K.to_nhwc(lrs.reshape(n * t, c, h, w).contiguous())                                 # (n*t, h, w, 3)
    fv = K.to_nhwc(fvs.to(torch.float32).reshape(n * t, c, 8 * h, 8 * w).contiguous())
    mkf = mks.reshape(n * t, 8 * h, 8 * w, 1).to(torch.float32)
    # clip-level stage (CRFP.py:1536-1553): flows, LR features, fovea compositing, HR features
    flows = None
    if t > 1:
        lr5 = lr.view(n, t, h, w, c)
        x1 = lr5[:, 1:].reshape(n * (t - 1), h, w, c)      # frame i
        x2 = lr5[:, :-1].reshape(n * (t - 1), h, w, c)     # frame i-1
        flows = net.fnet(x1, x2).view(n, t - 1, h, w, 2)
    lr_up8 = A.up_bilinear(K, lr, 8)
    x_lr = net.conv("encoder_lr.slice1.2", [net.conv("encoder_lr.slice1.0", [lr], A.ACT_LRELU)], A.ACT_LRELU)
    fv = fv * mkf + lr_up8 * (1.0 - mkf)
    x_hr = net.conv("encoder_hr.slice1.2",
                    [net.conv("encoder_hr.slice1.0", [torch.cat((fv, lr_up8), dim=-1)], A.ACT_LRELU)], A.ACT_LRELU)
    x_lr = x_lr.view(n, t, h, w, -1)
    x_hr = x_hr.view(n, t, 8 * h, 8 * w, -1)
    mk5 = mkf.view(n, t, 8 * h, 8 * w, 1)
    up5 = lr_up8.view(n, t, 8 * h, 8 * w, c)
    state, outs = None, []
    for i in range(t):
        out, state = net.frame_step(state, x_lr[:, i], x_hr[:, i], mk5[:, i], up5[:, i],
                                    flows[:, i - 1] if i > 0 else None)
        outs.append(out.permute(0, 3, 1, 2))
    return torch.stack(outs, dim=1)
