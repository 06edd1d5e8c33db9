"""ORACLE — CPU restatement of the CRFP_DSV hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file; the product path
(`crfp_b200/`) never does and fails loudly when its CUDA library is missing.

What it restates (op for op, fp32, plain PyTorch CPU ops, functional over a
state_dict so that it needs neither `/root/reference` nor `dcn_v2`):

  flow_warp                   /root/reference/model/CRFP.py:90-130
  PixelShufflePack            /root/reference/model/CRFP.py:154-193
  PixelUnShufflePack_v2       /root/reference/model/CRFP.py:239-279, 28-42
  DCN_module.forward          /root/reference/model/CRFP.py:324-352
  ResidualBlocksWithInputConv /root/reference/model/CRFP.py:433-552
  FNet.forward                /root/reference/model/CRFP.py:797-814
  CRFP_DSV.forward            /root/reference/model/CRFP.py:1510-1686
  CRFP.forward (v15)          /root/reference/model/CRFP.py:1223-1368
  CRFP_simple.forward (v13)   /root/reference/model/CRFP.py:938-1080
  LTE_simple_lr / _hr_single  /root/reference/model/LTE.py:34-51, 100-117
  MRCF_simple_v18 (streaming) /root/reference/model/CRFP_test.py:2214-2478

Third-party arithmetic: the modulated deformable convolution lives in
`jinfagang/DCNv2_latest` (README.md:26 of the reference; un-vendored, no
version pinned, not installed here).  Its published algorithm is restated in
`dcn_v2_naive` below and the fast path uses `torchvision.ops.deform_conv2d`
(same algorithm lineage, same channel layout); the two are checked against
each other in tests/test_oracle.py.  The reference holds no test or golden
vector at that boundary, so the DCNv2 boundary itself is "parity unpinned";
the oracle as a whole IS pinned: `oracle/make_golden.py` imports the real
reference classes in the build container (with a `dcn_v2` -> torchvision shim)
and the committed fixtures under tests/golden/ are the reference's own outputs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

try:  # fast CPU path for the deformable conv; same maths as dcn_v2_naive
    from torchvision.ops import deform_conv2d as _tv_deform_conv2d
except Exception:  # pragma: no cover
    _tv_deform_conv2d = None


# ----------------------------------------------------------------------------- basic ops
def conv3x3(x, sd, name):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)


def lrelu(x):
    return F.leaky_relu(x, 0.1)


def up_bilinear(x, scale):
    """nn.Upsample(scale_factor=s, mode='bilinear', align_corners=False) (CRFP.py:1471-1478)."""
    return F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False)


def flow_warp(x, flow_nchw, padding_mode="zeros"):
    """Backward warp of x (n,c,h,w) by flow (n,2,h,w): ch0 = x displacement, ch1 = y (CRFP.py:90-130)."""
    n, _, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    grid = torch.stack((gx, gy), 2).type_as(x)
    gf = grid + flow_nchw.permute(0, 2, 3, 1)
    gxn = 2.0 * gf[:, :, :, 0] / max(w - 1, 1) - 1.0
    gyn = 2.0 * gf[:, :, :, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(x, torch.stack((gxn, gyn), dim=3), mode="bilinear",
                         padding_mode=padding_mode, align_corners=True)


def flow_warp_indices(flow_nchw):
    """Integer sampling indices (x0, y0 = floor of the un-normalised grid position) that
    grid_sample uses for `flow_warp`, replicating the fp32 op sequence exactly
    (normalise CRFP.py:118-121, un-normalise ATen grid_sampler align_corners=True)."""
    n, _, h, w = flow_nchw.shape
    gy, gx = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")
    fx = gx.float() + flow_nchw[:, 0]
    fy = gy.float() + flow_nchw[:, 1]
    gxn = 2.0 * fx / max(w - 1, 1) - 1.0
    gyn = 2.0 * fy / max(h - 1, 1) - 1.0
    ix = ((gxn + 1.0) / 2.0) * (w - 1)
    iy = ((gyn + 1.0) / 2.0) * (h - 1)
    return ix.floor().to(torch.int32), iy.floor().to(torch.int32)


def dcn_v2_naive(x, offset, mask, weight, bias, dg):
    """Modulated deformable 3x3 conv, stride 1, pad 1, dil 1 — restated from the DCNv2 paper /
    jinfagang/DCNv2_latest `modulated_deformable_im2col`.

    out[n,o,y,x] = b[o] + sum_{c,i,j} W[o,c,i,j] * m[n,g*9+t,y,x] * bilinear(in[n,c], py, px)
    with t=i*3+j, g=c//(C/dg), py = y-1+i+off[n,(g*9+t)*2], px = x-1+j+off[n,(g*9+t)*2+1];
    a sample is 0 when py<=-1 or py>=H or px<=-1 or px>=W, else a 4-corner bilinear with
    out-of-range corners contributing 0.
    """
    n, c, h, w = x.shape
    co = weight.shape[0]
    cpg = c // dg
    ys = torch.arange(h, dtype=x.dtype).view(1, h, 1)
    xs = torch.arange(w, dtype=x.dtype).view(1, 1, w)
    cols = x.new_zeros(n, c, 9, h, w)
    xf = x.reshape(n, c, h * w)
    for g in range(dg):
        xg = xf[:, g * cpg:(g + 1) * cpg]
        for t in range(9):
            i, j = divmod(t, 3)
            py = ys - 1 + i + offset[:, (g * 9 + t) * 2]
            px = xs - 1 + j + offset[:, (g * 9 + t) * 2 + 1]
            inside = (py > -1) & (py < h) & (px > -1) & (px < w)
            y0 = torch.floor(py)
            x0 = torch.floor(px)
            ly, lx = py - y0, px - x0
            hy, hx = 1 - ly, 1 - lx
            val = x.new_zeros(n, cpg, h, w)
            for (yy, xx, wt) in ((y0, x0, hy * hx), (y0, x0 + 1, hy * lx),
                                 (y0 + 1, x0, ly * hx), (y0 + 1, x0 + 1, ly * lx)):
                ok = inside & (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
                idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).long().view(n, 1, h * w)
                v = torch.gather(xg, 2, idx.expand(n, cpg, h * w)).view(n, cpg, h, w)
                val = val + v * (wt * ok).unsqueeze(1)
            cols[:, g * cpg:(g + 1) * cpg, t] = val * mask[:, g * 9 + t].unsqueeze(1)
    out = torch.einsum("nckp,ock->nop", cols.reshape(n, c, 9, h * w), weight.reshape(co, c, 9))
    return out.view(n, co, h, w) + bias.view(1, co, 1, 1)


def dcn_v2(x, offset, mask, weight, bias, dg, naive=False):
    """`dcn_v2.DCNv2.forward(input, offset, mask)` (call site CRFP.py:350)."""
    if naive or _tv_deform_conv2d is None:
        return dcn_v2_naive(x, offset, mask, weight, bias, dg)
    return _tv_deform_conv2d(x, offset, weight, bias, stride=1, padding=1, dilation=1, mask=mask)


def pixel_shuffle_pack(x, sd, name, r):
    return F.pixel_shuffle(conv3x3(x, sd, name + ".upsample_conv"), r)


def res_blocks_with_input_conv(x, sd, name):
    """conv+LReLU then one ResidualBlockNoBN (CRFP.py:433-552, num_blocks=1, res_scale=1)."""
    x = lrelu(conv3x3(x, sd, name + ".main.0"))
    y = conv3x3(F.relu(conv3x3(x, sd, name + ".main.2.0.conv1")), sd, name + ".main.2.0.conv2")
    return x + y


# ----------------------------------------------------------------------------- DCN_module
def dcn_module(sd, name, cur_x, pre_x, pre_x_aligned, flow, pre_offset=None, *, dg, repeat=False,
               pixelshuffle=False, max_mag=10.0, naive=False, return_offsets=False):
    """DCN_module.forward (CRFP.py:324-352)."""
    z = torch.cat([cur_x, pre_x_aligned, flow], dim=1)
    z = lrelu(conv3x3(z, sd, name + ".dcn_block.0"))
    z = lrelu(conv3x3(z, sd, name + ".dcn_block.2"))
    if pre_offset is not None:
        if pixelshuffle:
            pre_offset = pixel_shuffle_pack(pre_offset, sd, name + ".upsample", 4) * 2.0
        z = lrelu(conv3x3(torch.cat([z, pre_offset], dim=1), sd, name + ".conv_fuse"))
    offset = max_mag * torch.tanh(conv3x3(z, sd, name + ".dcn_offset"))
    mask = torch.sigmoid(conv3x3(z, sd, name + ".dcn_mask"))
    if repeat:
        B, C, H, W = offset.shape
        offset = offset.view(B, 2, C // 2, H, W)
        offset = offset + flow.flip(1).unsqueeze(2).repeat(1, 1, C // 2, 1, 1)
        offset = offset.repeat(1, 9, 1, 1, 1).view(B, C * 9, H, W)
        mask = mask.repeat(1, 9, 1, 1)
    else:
        offset = offset + flow.flip(1).repeat(1, offset.size(1) // 2, 1, 1)
    out = dcn_v2(pre_x, offset, mask, sd[name + ".dcn.weight"], sd[name + ".dcn.bias"], dg, naive=naive)
    if return_offsets:
        return out, z, offset, mask
    return out, z


# ----------------------------------------------------------------------------- FNet
def fnet(sd, x1, x2, prefix="spynet."):
    """FNet.forward(x1, x2) (CRFP.py:797-814)."""
    _, _, h, w = x1.shape
    out = torch.cat([x1, x2], dim=1)
    for name in ("encoder1", "encoder2", "encoder3"):
        out = F.relu(conv3x3(out, sd, prefix + name + ".0"))
        out = F.relu(conv3x3(out, sd, prefix + name + ".2"))
        out = F.avg_pool2d(out, 2, 2)
    for name in ("decoder1", "decoder2", "decoder3"):
        out = F.relu(conv3x3(out, sd, prefix + name + ".0"))
        out = F.relu(conv3x3(out, sd, prefix + name + ".2"))
        out = up_bilinear(out, 2)
    out = F.relu(conv3x3(out, sd, prefix + "flow.0"))
    out = torch.tanh(conv3x3(out, sd, prefix + "flow.2")) * 256
    return F.interpolate(out, size=(h, w), mode="bilinear", align_corners=False)


def compute_flow(sd, lrs):
    """CRFP_DSV.compute_flow (CRFP.py:1483-1508): flow from frame i to frame i-1."""
    n, t, c, h, w = lrs.shape
    lrs_1 = lrs[:, :-1].reshape(-1, c, h, w)
    lrs_2 = lrs[:, 1:].reshape(-1, c, h, w)
    return fnet(sd, lrs_2, lrs_1).view(n, t - 1, 2, h, w)


# ----------------------------------------------------------------------------- the recurrence
def _split(y, C):
    """DSV split: chunk(4) -> first 3 chunks propagate, last chunk is next frame's state (CRFP.py:1592-1596)."""
    q = C // 4
    return y[:, :3 * q], y[:, 3 * q:]


def frame_step(sd, C, state, x_lr_cur, x_hr_cur, mk_cur, lr_cur, flow, *, fg_lv0=None, fg_lv3=None,
               naive_dcn=False, taps=None):
    """One iteration of the `for i in range(t)` loop of CRFP_DSV.forward (CRFP.py:1555-1684).

    state = None for the first frame, else (S, feat_lv0, feat_lv1, feat_lv2) with S the HR 4-ch plane.
    fg_* are the streaming model's regional masks (CRFP_test.py:2347-2389); None in clip mode.
    `taps` (dict) collects intermediates for per-op parity tests.
    """
    feat_prop_lv0 = pixel_shuffle_pack(x_lr_cur, sd, "upsample", 2)
    n = lr_cur.shape[0]
    h, w = lr_cur.shape[-2:]
    c = C // 8
    if state is not None:
        S0, feat_lv0, feat_lv1, feat_lv2 = state
        flow_lv3 = up_bilinear(flow, 2) * 2.0
        flow_lv0 = up_bilinear(flow, 8) * 8.0
        P = conv3x3(F.pixel_unshuffle(S0, 4), sd, "downsample.downsample_conv")
        P_w = flow_warp(P, flow_lv3)
        S0_w = flow_warp(S0, flow_lv0)
        feat_mix = flow_warp(torch.cat((feat_lv0, feat_lv1, feat_lv2), dim=1), flow_lv3)
        feat_lv0, feat_lv1, feat_lv2 = torch.chunk(feat_mix, 3, dim=1)
        feats = [feat_lv0, feat_lv1, feat_lv2]
        prop = feat_prop_lv0
        offfeat = None
        if taps is not None:
            taps.update(flow_lv3=flow_lv3, flow_lv0=flow_lv0, P=P, P_w=P_w, S0_w=S0_w, feat_mix=feat_mix)
        for k in range(3):
            cur = torch.cat((prop, feats[k]), dim=1)
            if taps is not None:
                A, offfeat, off, msk = dcn_module(sd, f"dcn_{k}", cur, P, P_w, flow_lv3, offfeat, dg=8,
                                                  naive=naive_dcn, return_offsets=True)
                taps[f"dcn{k}_cur"], taps[f"dcn{k}_offset"], taps[f"dcn{k}_mask"] = cur, off, msk
                taps[f"dcn{k}_out"], taps[f"dcn{k}_offfeat"] = A, offfeat
            else:
                A, offfeat = dcn_module(sd, f"dcn_{k}", cur, P, P_w, flow_lv3, offfeat, dg=8, naive=naive_dcn)
            y = torch.cat([cur, A], dim=1)
            if fg_lv0 is not None and k > 0:   # CRFP_test.py:2347 is a dead store for k == 0
                y = y * fg_lv0
            y = res_blocks_with_input_conv(y, sd, f"forward_resblocks_{k}")
            prop, feats[k] = _split(y, C)
            if taps is not None:
                taps[f"res{k}_out"] = y
        feat_lv0, feat_lv1, feat_lv2 = feats
        q = lrelu(pixel_shuffle_pack(prop, sd, "upsample_post", 4))
        if taps is not None:
            A3, _, off3, msk3 = dcn_module(sd, "dcn_3", q, S0, S0_w, flow_lv0, offfeat, dg=1, repeat=True,
                                           pixelshuffle=True, naive=naive_dcn, return_offsets=True)
            taps.update(q=q, dcn3_offset=off3, dcn3_mask=msk3, dcn3_out=A3)
        else:
            A3, _ = dcn_module(sd, "dcn_3", q, S0, S0_w, flow_lv0, offfeat, dg=1, repeat=True,
                               pixelshuffle=True, naive=naive_dcn)
        y = torch.cat([q, A3], dim=1)
        if fg_lv3 is not None:
            y = y * fg_lv3
        S = res_blocks_with_input_conv(y, sd, "forward_resblocks_3")
    else:
        zeros_l1 = lr_cur.new_zeros(n, C, 2 * h, 2 * w)
        zeros_hr = lr_cur.new_zeros(n, c, 8 * h, 8 * w)
        feats = [lr_cur.new_zeros(n, C // 4, 2 * h, 2 * w) for _ in range(3)]
        prop = feat_prop_lv0
        for k in range(3):
            y = res_blocks_with_input_conv(torch.cat([prop, zeros_l1, feats[k]], dim=1), sd,
                                           f"forward_resblocks_{k}")
            prop, feats[k] = _split(y, C)
        feat_lv0, feat_lv1, feat_lv2 = feats
        q = lrelu(pixel_shuffle_pack(prop, sd, "upsample_post", 4))
        S = res_blocks_with_input_conv(torch.cat([q, zeros_hr], dim=1), sd, "forward_resblocks_3")
    Fz = conv3x3(torch.cat([S, x_hr_cur], dim=1), sd, "conv_tttf")
    mkf = mk_cur.float()
    S = lrelu(mkf * Fz + (1 - mkf) * S)
    out = conv3x3(S, sd, "conv_last")
    out = out + up_bilinear(lr_cur, 8)
    if taps is not None:
        taps.update(S=S, out=out)
    return out, (S, feat_lv0, feat_lv1, feat_lv2)


def frame_step_v1x(sd, C, variant, state, x_lr_cur, x_hr_cur, mk_cur, lr_cur, flow, *, naive_dcn=False):
    """One loop iteration of CRFP.forward (v15, CRFP.py:1260-1368) / CRFP_simple.forward (v13, CRFP.py:968-1080).
    No DSV split; the HR state is warped first and both the state and its warp are downsampled; v15 additionally
    feeds the warped planes into every residual block (3-way concat)."""
    three = (variant == "v15")
    prop = pixel_shuffle_pack(x_lr_cur, sd, "upsample", 2)          # 32 ch @L1
    n = lr_cur.shape[0]
    h, w = lr_cur.shape[-2:]
    c = C // 8
    if state is not None:
        S0 = state
        flow_lv3 = up_bilinear(flow, 2) * 2.0
        flow_lv0 = up_bilinear(flow, 8) * 8.0
        S0_w = flow_warp(S0, flow_lv0)
        P_w = conv3x3(F.pixel_unshuffle(S0_w, 4), sd, "downsample.downsample_conv")
        P = conv3x3(F.pixel_unshuffle(S0, 4), sd, "downsample.downsample_conv")
        offfeat = None
        cur = prop
        for k in range(3):
            A, offfeat = dcn_module(sd, f"dcn_{k}", cur, P, P_w, flow_lv3, offfeat, dg=8, naive=naive_dcn)
            y = torch.cat([cur, A, P_w], dim=1) if three else torch.cat([cur, A], dim=1)
            cur = res_blocks_with_input_conv(y, sd, f"forward_resblocks_{k}")
        q = lrelu(pixel_shuffle_pack(cur, sd, "upsample_post", 4))
        A3, _ = dcn_module(sd, "dcn_3", q, S0, S0_w, flow_lv0, offfeat, dg=1, repeat=True, pixelshuffle=True,
                           naive=naive_dcn)
        y = torch.cat([q, A3, S0_w], dim=1) if three else torch.cat([q, A3], dim=1)
        S = res_blocks_with_input_conv(y, sd, "forward_resblocks_3")
    else:
        z1 = lr_cur.new_zeros(n, C, 2 * h, 2 * w)
        z3 = lr_cur.new_zeros(n, c, 8 * h, 8 * w)
        cur = prop
        for k in range(3):
            y = torch.cat([cur, z1, z1], dim=1) if three else torch.cat([cur, z1], dim=1)
            cur = res_blocks_with_input_conv(y, sd, f"forward_resblocks_{k}")
        q = lrelu(pixel_shuffle_pack(cur, sd, "upsample_post", 4))
        y = torch.cat([q, z3, z3], dim=1) if three else torch.cat([q, z3], dim=1)
        S = res_blocks_with_input_conv(y, sd, "forward_resblocks_3")
    Fz = conv3x3(torch.cat([S, x_hr_cur], dim=1), sd, "conv_tttf")
    mkf = mk_cur.float()
    S = lrelu(mkf * Fz + (1 - mkf) * S)
    out = conv3x3(S, sd, "conv_last") + up_bilinear(lr_cur, 8)
    return out, S


@torch.no_grad()
def crfp_forward(sd, lrs, fvs, mks, variant="dsv", mid_channels=32, naive_dcn=False):
    """forward(lrs, fvs, mks) of CRFP_DSV ("dsv"), CRFP ("v15") or CRFP_simple ("v13")."""
    if variant == "dsv":
        return crfp_dsv_forward(sd, lrs, fvs, mks, mid_channels, naive_dcn)
    n, t, c, h, w = lrs.shape
    flows = compute_flow(sd, lrs) if t > 1 else None
    x_lr, x_hr = encoders(sd, lrs, fvs, mks)
    state, outs = None, []
    for i in range(t):
        out, state = frame_step_v1x(sd, mid_channels, variant, state, x_lr[:, i], x_hr[:, i], mks[:, i], lrs[:, i],
                                    flows[:, i - 1] if i > 0 else None, naive_dcn=naive_dcn)
        outs.append(out)
    return torch.stack(outs, dim=1)


def encoders(sd, lrs, fvs, mks):
    """Clip-level part of CRFP_DSV.forward (CRFP.py:1536-1553): LR features, fovea compositing, HR features."""
    B, N, Cc, H, W = lrs.shape
    lrs_lv0 = lrs.reshape(B * N, Cc, H, W)
    lrs_lv3 = up_bilinear(lrs_lv0, 8)
    x_lr = lrelu(conv3x3(lrelu(conv3x3(lrs_lv0, sd, "encoder_lr.slice1.0")), sd, "encoder_lr.slice1.2"))
    mkf = mks.float()
    fvs = fvs * mkf + lrs_lv3.view(B, N, Cc, H * 8, W * 8) * (1 - mkf)
    x = torch.cat((fvs.view(B * N, Cc, H * 8, W * 8), lrs_lv3), dim=1)
    x_hr = lrelu(conv3x3(lrelu(conv3x3(x, sd, "encoder_hr.slice1.0")), sd, "encoder_hr.slice1.2"))
    return x_lr.view(B, N, -1, H, W), x_hr.view(B, N, -1, H * 8, W * 8)


@torch.no_grad()
def crfp_dsv_forward(sd, lrs, fvs, mks, mid_channels=32, naive_dcn=False, taps=None):
    """CRFP_DSV.forward(lrs, fvs, mks) -> (n,t,3,8h,8w) (CRFP.py:1510-1686).

    `taps`, if a list, receives one dict of intermediates per frame.
    """
    n, t, c, h, w = lrs.shape
    flows = compute_flow(sd, lrs) if t > 1 else None
    x_lr, x_hr = encoders(sd, lrs, fvs, mks)
    state = None
    outs = []
    for i in range(t):
        tp = {} if taps is not None else None
        flow = flows[:, i - 1] if i > 0 else None
        if tp is not None and flow is not None:
            tp["flow"] = flow
        out, state = frame_step(sd, mid_channels, state, x_lr[:, i], x_hr[:, i], mks[:, i], lrs[:, i], flow,
                                naive_dcn=naive_dcn, taps=tp)
        if taps is not None:
            taps.append(tp)
        outs.append(out)
    return torch.stack(outs, dim=1)


class StreamingOracle:
    """CRFP_test.MRCF_simple_v18 (CRFP_test.py:2114-2478): stateful, `forward(lrs, fvs, mks, fgs)` +
    `clear_states()`.  The first frame of a stream is paired with itself for the flow
    (CRFP_test.py:2232-2239) but takes the no-alignment branch (CRFP_test.py:2306,2396)."""

    def __init__(self, sd, mid_channels=32):
        self.sd, self.C = sd, mid_channels
        self.clear_states()

    def clear_states(self):
        self.pre_lr = None
        self.state = None

    @torch.no_grad()
    def forward(self, lrs, fvs, mks, fgs):
        sd = self.sd
        n, t, c, h, w = lrs.shape
        if self.pre_lr is not None:
            ext = torch.cat((self.pre_lr.unsqueeze(1), lrs), dim=1)
        else:
            ext = torch.cat((lrs[:, -1].unsqueeze(1), lrs), dim=1)
        self.pre_lr = lrs[:, -1].clone()
        flows = compute_flow(sd, ext)
        x_lr, x_hr = encoders(sd, lrs, fvs, mks)
        B, N, _, H, W = fgs.shape
        fg_lv3 = fgs.float()
        fg_lv0 = up_bilinear(fg_lv3.view(B * N, 1, H, W), 0.25).view(B, N, 1, H // 4, W // 4)
        outs = []
        for i in range(t):
            out, self.state = frame_step(sd, self.C, self.state, x_lr[:, i], x_hr[:, i], mks[:, i], lrs[:, i],
                                         flows[:, i], fg_lv0=fg_lv0[:, i], fg_lv3=fg_lv3[:, i])
            outs.append(out)
        return torch.stack(outs, dim=1)

    __call__ = forward


# ----------------------------------------------------------------------------- SPyNet
SPYNET_MEAN = (0.485, 0.456, 0.406)
SPYNET_STD = (0.229, 0.224, 0.225)
SPYNET_WIDTHS = ((8, 32), (32, 64), (64, 32), (32, 16), (16, 2))


def spynet_basic_module(sd, level, x, prefix=""):
    """SPyNetBasicModule.forward (CRFP.py:687-741): five `conv(act(x))` 7x7 layers — ReLU is applied to the INPUT of
    every conv including the first (`conv.forward`, CRFP.py:145-152), there is no activation on the output."""
    for i in range(5):
        name = f"{prefix}basic_module.{level}.basic_module.{i}.conv"
        x = F.conv2d(F.relu(x), sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=3)
    return x


def spynet_compute_flow(sd, ref, supp, prefix=""):
    """SPyNet.compute_flow (CRFP.py:593-650): inputs already multiples of 32."""
    n, _, h, w = ref.shape
    mean = ref.new_tensor(SPYNET_MEAN).view(1, 3, 1, 1)
    std = ref.new_tensor(SPYNET_STD).view(1, 3, 1, 1)
    ref = [(ref - mean) / std]
    supp = [(supp - mean) / std]
    for _ in range(5):
        ref.append(F.avg_pool2d(ref[-1], kernel_size=2, stride=2, count_include_pad=False))
        supp.append(F.avg_pool2d(supp[-1], kernel_size=2, stride=2, count_include_pad=False))
    ref, supp = ref[::-1], supp[::-1]
    flow = ref[0].new_zeros(n, 2, h // 32, w // 32)
    for level in range(len(ref)):
        if level == 0:
            flow_up = flow
        else:
            flow_up = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
        warped = flow_warp(supp[level], flow_up, padding_mode="border")
        flow = flow_up + spynet_basic_module(sd, level, torch.cat([ref[level], warped, flow_up], 1), prefix)
    return flow


def spynet(sd, ref, supp, prefix=""):
    """SPyNet.forward(ref, supp) (CRFP.py:652-685): resize to multiples of 32, pyramid, resize back, rescale the flow."""
    h, w = ref.shape[2:4]
    w_up = w if (w % 32) == 0 else 32 * (w // 32 + 1)
    h_up = h if (h % 32) == 0 else 32 * (h // 32 + 1)
    ref = F.interpolate(ref, size=(h_up, w_up), mode="bilinear", align_corners=False)
    supp = F.interpolate(supp, size=(h_up, w_up), mode="bilinear", align_corners=False)
    flow = F.interpolate(spynet_compute_flow(sd, ref, supp, prefix), size=(h, w), mode="bilinear", align_corners=False)
    scale = flow.new_tensor([float(w) / float(w_up), float(h) / float(h_up)]).view(1, 2, 1, 1)
    return flow * scale


# ----------------------------------------------------------------------------- CRFP_runtime.MRCF_simple_v18
def _rt_res_blocks(sd, name, feat1, feat2=None):
    """ResidualBlocksWithInputConv / _v2 of the runtime file (CRFP_runtime.py:464-556): `conv1(feat1)` pasted into the
    top-left corner of `conv2(feat2)` when a second (larger) input is given, LeakyReLU, then ONE bottleneck
    ResidualBlockNoBN (C -> C/2 -> C, CRFP_runtime.py:406-462; `main` = [LeakyReLU, Sequential(block)])."""
    o1 = conv3x3(feat1, sd, name + ".conv1")
    if feat2 is not None:
        feat = conv3x3(feat2, sd, name + ".conv2").clone()
        feat[:, :, :o1.shape[2], :o1.shape[3]] = o1
    else:
        feat = o1
    x = lrelu(feat)
    return x + conv3x3(F.relu(conv3x3(x, sd, name + ".main.1.0.conv1")), sd, name + ".main.1.0.conv2")


@torch.no_grad()
def runtime_v18_forward(sd, lrs, fvs, warp_size=(1080, 1920), mid_channels=32):
    """CRFP_runtime.MRCF_simple_v18.forward(lrs, fvs, warp_size) (CRFP_runtime.py:8472-8664), the timing harness of
    test_runtime.py: alignment (flow, warps, DCN, state) only inside the top-left `warp_size` (HR pixels) region, `fvs`
    an HR image of its own size anchored at the top-left corner and fused WITHOUT a mask, no level-to-level propagation
    of the residual-block outputs after the first frame (CRFP_runtime.py:8552,8567,8582: commented out)."""
    C = mid_channels
    WP_h, WP_w = warp_size
    n, t, c, h, w = lrs.shape
    flows = compute_flow(sd, lrs[:, :, :, :WP_h // 8, :WP_w // 8]) if t > 1 else None
    lrs0 = lrs.reshape(n * t, c, h, w)
    x_lr = lrelu(conv3x3(lrelu(conv3x3(lrs0, sd, "encoder_lr.slice1.0")), sd, "encoder_lr.slice1.2")).view(n, t, -1, h, w)
    fh, fw = fvs.shape[-2:]
    f0 = fvs.reshape(n * t, c, fh, fw)
    x_hr = lrelu(conv3x3(lrelu(conv3x3(torch.cat((f0, f0), 1), sd, "encoder_hr.slice1.0")), sd, "encoder_hr.slice1.2"))
    x_hr = x_hr.view(n, t, -1, fh, fw)
    q4 = C // 4
    outs, S, feats = [], None, [None] * 3
    for i in range(t):
        prop = pixel_shuffle_pack(x_lr[:, i], sd, "upsample", 2)
        if i > 0:
            flow = flows[:, i - 1]
            flow_lv3 = up_bilinear(flow, 2) * 2.0
            flow_lv0 = up_bilinear(flow, 8) * 8.0
            S0 = S
            S0_w = flow_warp(S0, flow_lv0)
            P_w = conv3x3(F.pixel_unshuffle(S0_w, 4), sd, "downsample.downsample_conv")
            P = conv3x3(F.pixel_unshuffle(S0, 4), sd, "downsample.downsample_conv")
            feats = list(torch.chunk(flow_warp(torch.cat(feats, dim=1), flow_lv3), 3, dim=1))
            offfeat = None
            for k in range(3):
                cur = torch.cat((prop[:, :, :WP_h // 4, :WP_w // 4], feats[k]), dim=1)
                A, offfeat = dcn_module(sd, f"dcn_{k}", cur, P, P_w, flow_lv3, offfeat, dg=8)
                y = _rt_res_blocks(sd, f"forward_resblocks_{k}", torch.cat([cur, A], dim=1), cur)
                feats[k] = y[:, 3 * q4:][:, :, :WP_h // 4, :WP_w // 4]
            q = lrelu(pixel_shuffle_pack(prop, sd, "upsample_post", 4))
            qc = q[:, :, :WP_h, :WP_w]
            A3, _ = dcn_module(sd, "dcn_3", qc, S0, S0_w, flow_lv0, offfeat, dg=1, repeat=True, pixelshuffle=True)
            S = _rt_res_blocks(sd, "forward_resblocks_3", torch.cat([qc, A3], dim=1), q)
        else:
            for k in range(3):
                y = _rt_res_blocks(sd, f"forward_resblocks_{k}_", prop)
                feats[k] = y[:, 3 * q4:][:, :, :WP_h // 4, :WP_w // 4]
                prop = y[:, :3 * q4]
            q = lrelu(pixel_shuffle_pack(prop, sd, "upsample_post", 4))
            S = _rt_res_blocks(sd, "forward_resblocks_3_", q)
        Fz = conv3x3(torch.cat([S[:, :, :fh, :fw], x_hr[:, i]], dim=1), sd, "conv_tttf")
        S = S.clone()
        S[:, :, :fh, :fw] = Fz
        S = lrelu(S)
        outs.append(conv3x3(S, sd, "conv_last") + up_bilinear(lrs[:, i], 8))
        S = S[:, :, :WP_h, :WP_w]
    return torch.stack(outs, dim=1)
