"""Operator-level host wrappers over the C ABI (torch tensors in, torch tensors out).

These mirror the reference's operator interfaces for the hot path so that the parity tests read like the
reference's own call sites:

  flow_warp(x, flow)                 /root/reference/model/CRFP.py:90-130   (x NCHW, flow (n,h,w,2))
  DCNv2(...).forward(input, offset, mask)   dcn_v2 extension, call site CRFP.py:350 (all NCHW)
  conv3x3(...)                       nn.Conv2d(3x3) + fused epilogues, NHWC building block

torch is used for device memory and the current stream only; all arithmetic runs in libcrfp_b200.so.
Tensors must be CUDA fp32; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from .packing import cin_packed, cout_packed, pack_conv, pack_conv_tc, pack_conv_tc3, pack_dcn


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise L.CrfpError(f"{name} must be a CUDA tensor (libcrfp_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise L.CrfpError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def to_nhwc(x: torch.Tensor, cpad: int | None = None) -> torch.Tensor:
    """NCHW -> NHWC (optionally zero-padded to `cpad` channels) with the library's layout kernel."""
    x = _req(x, "x")
    n, c, h, w = x.shape
    cpad = cpad or c
    out = torch.empty(n, h, w, cpad, device=x.device, dtype=torch.float32)
    L.check(L.lib().crfp_nchw_to_nhwc(n, c, h, w, x.data_ptr(), c * h * w, cpad, out.data_ptr(), _stream()), "nchw_to_nhwc")
    return out


def to_nchw(x: torch.Tensor, c: int | None = None, coffset: int = 0) -> torch.Tensor:
    x = _req(x, "x")
    n, h, w, cs = x.shape
    c = c or cs
    out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    L.check(L.lib().crfp_nhwc_to_nchw(n, c, h, w, x.data_ptr(), cs, coffset, out.data_ptr(), c * h * w, _stream()),
            "nhwc_to_nchw")
    return out


def flow_warp_nhwc(x: torch.Tensor, flow: torch.Tensor, border: bool = False) -> torch.Tensor:
    """x (n,h,w,c) NHWC with c % 4 == 0, flow (n,h,w,2) -> (n,h,w,c)."""
    x, flow = _req(x, "x"), _req(flow, "flow")
    n, h, w, c = x.shape
    if tuple(flow.shape) != (n, h, w, 2):
        raise ValueError(f"The spatial sizes of input ({(h, w)}) and flow ({tuple(flow.shape[1:3])}) are not the same.")
    out = torch.empty_like(x)
    d = L.WarpDesc(n=n, h=h, w=w, c=c, x=x.data_ptr(), x_cstride=c, x_coffset=0, flow=flow.data_ptr(),
                   out=out.data_ptr(), out_cstride=c, out_coffset=0, border=int(border))
    L.check(L.lib().crfp_flow_warp_fwd(C.byref(d), _stream()), "flow_warp")
    return out


def flow_warp(x: torch.Tensor, flow: torch.Tensor, interpolation="bilinear", padding_mode="zeros",
              align_corners=True) -> torch.Tensor:
    """Drop-in for the reference's flow_warp (CRFP.py:90): x (n,c,h,w), flow (n,h,w,2) -> (n,c,h,w)."""
    if interpolation != "bilinear" or not align_corners or padding_mode not in ("zeros", "border"):
        raise L.CrfpError("flow_warp: only bilinear / align_corners=True / zeros|border is implemented")
    if x.size()[-2:] != flow.size()[1:3]:
        raise ValueError(f"The spatial sizes of input ({x.size()[-2:]}) and flow ({flow.size()[1:3]}) are not the same.")
    c = x.shape[1]
    cp = (c + 3) // 4 * 4
    y = flow_warp_nhwc(to_nhwc(x, cp), flow, border=(padding_mode == "border"))
    return to_nchw(y, c)


def flow_warp_indices(flow: torch.Tensor):
    """Integer corner indices (x0, y0) used by flow_warp for flow (n,h,w,2)."""
    flow = _req(flow, "flow")
    n, h, w, _ = flow.shape
    x0 = torch.empty(n, h, w, device=flow.device, dtype=torch.int32)
    y0 = torch.empty_like(x0)
    L.check(L.lib().crfp_flow_warp_indices(n, h, w, flow.data_ptr(), x0.data_ptr(), y0.data_ptr(), _stream()),
            "flow_warp_indices")
    return x0, y0


def dcn_v2_nhwc(x, offset, mask, w_packed, b_packed, dg, cout, shared_taps=False, indices=False):
    """x (n,h,w,C), offset (n,h,w,dg*18) [or 2*dg shared], mask (n,h,w,dg*9) [or dg] -> (n,h,w,cout)."""
    x, offset, mask = _req(x, "input"), _req(offset, "offset"), _req(mask, "mask")
    n, h, w, c = x.shape
    out = torch.empty(n, h, w, cout, device=x.device, dtype=torch.float32)
    d = L.DcnDesc(n=n, h=h, w=w, c=c, cout=cout, dg=dg, shared_taps=int(shared_taps),
                  x=x.data_ptr(), x_cstride=c, x_coffset=0,
                  offset=offset.data_ptr(), off_cstride=offset.shape[-1], off_coffset=0,
                  mask=mask.data_ptr(), mask_cstride=mask.shape[-1], mask_coffset=0,
                  weight=w_packed.data_ptr(), bias=b_packed.data_ptr(),
                  out=out.data_ptr(), out_cstride=cout, out_coffset=0)
    if indices:
        y0 = torch.empty(n, h, w, dg * 9, device=x.device, dtype=torch.int32)
        x0 = torch.empty_like(y0)
        L.check(L.lib().crfp_dcn_v2_indices(C.byref(d), y0.data_ptr(), x0.data_ptr(), _stream()), "dcn_v2_indices")
        return y0, x0
    L.check(L.lib().crfp_dcn_v2_fwd(C.byref(d), _stream()), "dcn_v2")
    return out


class DCNv2(nn.Module):
    """Drop-in for `dcn_v2.DCNv2` as the reference constructs and calls it (CRFP.py:318-320, 350):
    DCNv2(in, out, 3, stride=1, padding=1, dilation=1, deformable_groups=dg); forward(input, offset, mask), NCHW."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1, deformable_groups=1):
        super().__init__()
        if (kernel_size, stride, padding, dilation) != (3, 1, 1, 1):
            raise L.CrfpError("DCNv2: only 3x3 / stride 1 / pad 1 / dilation 1 (the CRFP configuration) is implemented")
        self.in_channels, self.out_channels, self.deformable_groups = in_channels, out_channels, deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        stdv = 1.0 / (in_channels * 9) ** 0.5
        self.weight.data.uniform_(-stdv, stdv)

    @torch.no_grad()
    def forward(self, input, offset, mask):
        dg = self.deformable_groups
        if offset.shape[1] != dg * 18 or mask.shape[1] != dg * 9:
            raise L.CrfpError("DCNv2: offset/mask channel count does not match deformable_groups")
        wp, bp = pack_dcn(self.weight, self.bias, dg)
        y = dcn_v2_nhwc(to_nhwc(input), to_nhwc(offset), to_nhwc(mask), wp, bp, dg, self.out_channels)
        return to_nchw(y)


def conv3x3_nhwc(srcs, weight, bias, act=L.ACT_NONE, residual=None, shuffle_r=0, post_scale=1.0, split=None,
                 flow=None, head_split=0, head_mag=10.0, modes=None, packed=None):
    """3x3 conv over the channel concat of NHWC `srcs` with an OIHW weight; returns NHWC output(s).

    split=(c0, c1): route the first c0 output channels to one tensor and the next c1 to another.
    shuffle_r>0: F.pixel_shuffle(r) fused into the store.  modes[i]=1 reads source i through pixel_unshuffle(4).
    """
    srcs = [_req(s, "src") for s in srcs]
    modes = modes or [0] * len(srcs)
    n = srcs[0].shape[0]
    c_list = []
    for s, m in zip(srcs, modes):
        c_list.append(s.shape[-1] * 16 if m == L.SRC_UNSHUFFLE4 else s.shape[-1])
    h, w = (srcs[0].shape[1], srcs[0].shape[2]) if modes[0] == 0 else (srcs[0].shape[1] // 4, srcs[0].shape[2] // 4)
    cout = weight.shape[0]
    wp, bp = packed if packed is not None else pack_conv(weight, bias, c_list, modes)   # `packed`: a caller-side cache
    assert wp.shape[1] == cin_packed(c_list) and wp.shape[2] == cout_packed(cout)
    d = L.ConvDesc()
    d.n, d.h, d.w, d.nsrc = n, h, w, len(srcs)
    for i, (s, m) in enumerate(zip(srcs, modes)):
        d.src[i] = L.Src(ptr=s.data_ptr(), c=c_list[i], cstride=s.shape[-1], coffset=0, mode=m)
    d.cout, d.act = cout, act
    d.weight, d.bias = wp.data_ptr(), bp.data_ptr()
    d.post_scale, d.head_mag, d.head_split = post_scale, head_mag, head_split
    if flow is not None:
        flow = _req(flow, "flow")
        d.flow = flow.data_ptr()
    if residual is not None:
        residual = _req(residual, "residual")
        d.residual, d.res_cstride, d.res_coffset = residual.data_ptr(), residual.shape[-1], 0
    dev = srcs[0].device
    if shuffle_r:
        r = shuffle_r
        outs = [torch.empty(n, h * r, w * r, cout // (r * r), device=dev, dtype=torch.float32)]
        d.out_mode, d.shuffle_r, d.ndst = L.OUT_SHUFFLE, r, 1
        d.dst[0] = L.Dst(ptr=outs[0].data_ptr(), c=cout // (r * r), cstride=cout // (r * r), coffset=0)
    else:
        parts = list(split) if split else [cout]
        outs = [torch.empty(n, h, w, c, device=dev, dtype=torch.float32) for c in parts]
        d.out_mode, d.ndst = L.OUT_NHWC, len(parts)
        for i, o in enumerate(outs):
            d.dst[i] = L.Dst(ptr=o.data_ptr(), c=parts[i], cstride=parts[i], coffset=0)
    L.check(L.lib().crfp_conv3x3_fwd(C.byref(d), _stream()), "conv3x3")
    return outs[0] if len(outs) == 1 else tuple(outs)


def resize_bilinear_nhwc(x, hout, wout, rscale_h, rscale_w, mul=1.0):
    x = _req(x, "x")
    n, h, w, c = x.shape
    out = torch.empty(n, hout, wout, c, device=x.device, dtype=torch.float32)
    L.check(L.lib().crfp_resize_bilinear(n, h, w, c, x.data_ptr(), hout, wout, rscale_h, rscale_w, mul, out.data_ptr(),
                                         _stream()), "resize_bilinear")
    return out


def avgpool2_nhwc(x):
    x = _req(x, "x")
    n, h, w, c = x.shape
    out = torch.empty(n, h // 2, w // 2, c, device=x.device, dtype=torch.float32)
    L.check(L.lib().crfp_avgpool2(n, h, w, c, x.data_ptr(), out.data_ptr(), _stream()), "avgpool2")
    return out


def conv3x3_tc_nhwc(srcs, weight, bias, act=L.ACT_NONE, residual=None, out_kind=L.TC_OUT_BF16, shuffle_r=0,
                    post_scale=1.0, split=None, flow=None, head_split=0, head_mag=10.0, c_real=None):
    """Tensor-core (tcgen05) 3x3 conv over the channel concat of bf16 NHWC `srcs` (channel counts multiples of 8)
    with an OIHW fp32 weight.  `c_real[i]` = how many channels of source i the reference weight really has
    (the remaining channels of that source must be zero)."""
    for s in srcs:
        if not (s.is_cuda and s.dtype == torch.bfloat16):
            raise L.CrfpError("conv3x3_tc_nhwc: sources must be CUDA bfloat16 NHWC")
    srcs = [s.contiguous() for s in srcs]
    n, h, w, _ = srcs[0].shape
    c_real = c_real or [s.shape[-1] for s in srcs]
    cout = weight.shape[0]
    wp, bp = pack_conv_tc(weight, bias, c_real)
    d = L.ConvTcDesc()
    d.n, d.h, d.w, d.nsrc = n, h, w, len(srcs)
    for i, s in enumerate(srcs):
        d.src[i] = L.TcSrc(ptr=s.data_ptr(), c=s.shape[-1], cstride=s.shape[-1], coffset=0)
    d.cout, d.act = cout, act
    d.weight, d.bias = wp.data_ptr(), bp.data_ptr()
    d.post_scale, d.head_mag, d.head_split = post_scale, head_mag, head_split
    d.out_kind, d.shuffle_r = out_kind, shuffle_r
    if flow is not None:
        flow = _req(flow, "flow")
        d.flow = flow.data_ptr()
    if residual is not None:
        residual = residual.contiguous()
        d.residual, d.res_cstride, d.res_coffset = residual.data_ptr(), residual.shape[-1], 0
    dev = srcs[0].device
    if out_kind == L.TC_OUT_SHUFFLE_F32:
        r = shuffle_r
        outs = [torch.zeros(n, h * r, w * r, cout // (r * r), device=dev, dtype=torch.float32)]
    elif out_kind == L.TC_OUT_F32:
        outs = [torch.zeros(n, h, w, cout, device=dev, dtype=torch.float32)]
    else:
        parts = list(split) if split else [cout]
        outs = [torch.zeros(n, h, w, c, device=dev, dtype=torch.bfloat16) for c in parts]
    d.ndst = len(outs)
    for i, o in enumerate(outs):
        d.dst[i] = L.TcSrc(ptr=o.data_ptr(), c=o.shape[-1], cstride=o.shape[-1], coffset=0)
    L.check(L.lib().crfp_conv3x3_tc_fwd(C.byref(d), _stream()), "conv3x3_tc")
    return outs[0] if len(outs) == 1 else tuple(outs)


def conv3x3_tc3_nhwc(srcs, weight, bias, act=L.ACT_NONE, residual=None, shuffle_r=0, post_scale=1.0, split=None,
                     flow=None, head_split=0, head_mag=10.0, extra=None, packed=None, cout=None, src_slice=None, res_pre=False,
                     out=None):
    """fp32-accurate tensor-core conv (3 x bf16 split) over the channel concat of fp32 NHWC `srcs` (+ optional
    2-channel fp32 `extra` source whose weights are the trailing input channels of `weight`).  `packed` = the
    (hi, lo, bias, w_extra) operands of a caller-side cache (then `weight` / `bias` are not read and `cout` is required).
    K-split passes of a conv over more than 64 input channels: `src_slice=(offset, count)` reads a channel slice of the single
    source, `res_pre=True` adds `residual` before the activation, `out=` writes into an existing tensor (in place over the
    residual)."""
    srcs = [_req(s, "src") for s in srcs]
    n, h, w, _ = srcs[0].shape
    c_list = [s.shape[-1] for s in srcs] if src_slice is None else [src_slice[1]]
    cout = weight.shape[0] if cout is None else cout
    hi, lo, bp, wx = packed if packed is not None else pack_conv_tc3(weight, bias, c_list, extra=0 if extra is None else extra.shape[-1])
    d = L.ConvTc3Desc()
    d.n, d.h, d.w, d.nsrc = n, h, w, len(srcs)
    for i, s in enumerate(srcs):
        if src_slice is None:
            d.src[i] = L.TcSrc(ptr=s.data_ptr(), c=s.shape[-1], cstride=s.shape[-1], coffset=0)
        else:
            d.src[i] = L.TcSrc(ptr=s.data_ptr(), c=src_slice[1], cstride=s.shape[-1], coffset=src_slice[0])
    d.cout, d.act = cout, act
    d.res_pre = int(res_pre)
    d.weight_hi, d.weight_lo, d.bias = hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
    if extra is not None:
        extra = _req(extra, "extra")
        d.extra, d.w_extra = extra.data_ptr(), wx.data_ptr()
    d.post_scale, d.head_mag, d.head_split = post_scale, head_mag, head_split
    if flow is not None:
        flow = _req(flow, "flow")
        d.flow = flow.data_ptr()
    if residual is not None:
        residual = _req(residual, "residual")
        d.residual, d.res_cstride, d.res_coffset = residual.data_ptr(), residual.shape[-1], 0
    dev = srcs[0].device
    if shuffle_r:
        r = shuffle_r
        outs = [torch.zeros(n, h * r, w * r, cout // (r * r), device=dev, dtype=torch.float32)]
        d.out_kind, d.shuffle_r = L.TC_OUT_SHUFFLE_F32, r
    else:
        parts = list(split) if split else [cout]
        # every channel of every pixel is written when the segments are whole 4-channel groups (the kernel's store unit)
        alloc = torch.empty if all(c % 4 == 0 for c in parts) else torch.zeros
        outs = [out] if out is not None else [alloc(n, h, w, c, device=dev, dtype=torch.float32) for c in parts]
        d.out_kind = L.TC_OUT_F32
    d.ndst = len(outs)
    for i, o in enumerate(outs):
        d.dst[i] = L.TcSrc(ptr=o.data_ptr(), c=o.shape[-1], cstride=o.shape[-1], coffset=0)
    L.check(L.lib().crfp_conv3x3_tc3_fwd(C.byref(d), _stream()), "conv3x3_tc3")
    return outs[0] if len(outs) == 1 else tuple(outs)


def dcn_align_fused_nhwc(z, flow, x, w_off, b_off, w_msk, b_msk, w_dcn, b_dcn, head_mag=10.0, indices=False):
    """The tail of DCN_module.forward (CRFP.py:337-350) as ONE kernel (`crfp_dcn_align_fused`): z, x (n,h,w,32), flow
    (n,h,w,2) fp32 NHWC; dcn_offset / dcn_mask / dcn weights in the reference's OIHW layout.  Returns the aligned tensor
    (n,h,w,32) — and, with indices=True, the int32 (n,h,w,72) floor(py), floor(px) the kernel sampled at."""
    from .packing import pack_align_heads, pack_dcn_tc3
    z, flow, x = _req(z, "z"), _req(flow, "flow"), _req(x, "x")
    n, h, w, _ = z.shape
    wf, bf = pack_align_heads(w_off, b_off, w_msk, b_msk)
    hi, lo, bp = pack_dcn_tc3(w_dcn, b_dcn, 8)
    out = torch.empty(n, h, w, 32, device=z.device, dtype=torch.float32)
    d = L.AlignFusedDesc(n=n, h=h, w=w, z=z.data_ptr(), z_cstride=32, z_coffset=0, flow=flow.data_ptr(), x=x.data_ptr(),
                         x_cstride=32, x_coffset=0, heads_w=wf.data_ptr(), heads_b=bf.data_ptr(), dcn_w_hi=hi.data_ptr(),
                         dcn_w_lo=lo.data_ptr(), dcn_b=bp.data_ptr(), out=out.data_ptr(), out_cstride=32, out_coffset=0,
                         head_mag=head_mag)
    if indices:
        y0 = torch.full((n, h, w, 72), -12345, device=z.device, dtype=torch.int32)
        x0 = torch.full_like(y0, -12345)
        d.dbg_y0, d.dbg_x0 = y0.data_ptr(), x0.data_ptr()
    L.check(L.lib().crfp_dcn_align_fused(C.byref(d), _stream()), "dcn_align_fused")
    return (out, y0, x0) if indices else out
