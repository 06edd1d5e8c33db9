"""GPU tests (-m gpu) of the tcgen05 tensor-core kernels (bf16 storage, fp32 accumulation) against a PyTorch
fp32 reference evaluated on the SAME bf16-rounded operands: differences are accumulation order + the final
bf16 rounding of the output (<= 2^-8 relative)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def bf(x):
    return x.to(torch.bfloat16).float()


def nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous().cpu()


def close_bf16(got, ref, extra=0.0):
    tol = ref.abs() * 2.0 ** -7 + 2e-3 + extra
    bad = (got - ref).abs() > tol
    assert not bad.any(), f"max err {(got - ref).abs().max().item():.4e}, {int(bad.sum())} elements out of tolerance"


@pytest.mark.parametrize("c_list,cout,hw,act", [
    ([32], 32, (9, 128), 1), ([32], 32, (21, 300), 1), ([32, 32], 32, (12, 130), 1), ([32, 32, 8], 32, (10, 70), 1),
    ([24], 64, (16, 140), 0), ([32], 216, (7, 200), 0), ([64], 96, (5, 64), 2)])
def test_conv_tc(c_list, cout, hw, act):
    from crfp_b200 import ops
    g = _g(1)
    h, w = hw
    n = 2
    srcs = [bf(torch.randn(n, c, h, w, generator=g)) for c in c_list]
    wt = bf(torch.randn(cout, sum(c_list), 3, 3, generator=g) * (2.0 / (9 * sum(c_list))) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(torch.cat(srcs, 1), wt, b, padding=1)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else F.relu(ref) if act == 2 else ref
    got = nchw(ops.conv3x3_tc_nhwc([nhwc_bf16(s) for s in srcs], wt.cuda(), b.cuda(), act=act))
    close_bf16(got, ref)


def test_conv_tc_epilogues():
    from crfp_b200 import _lib as L
    from crfp_b200 import ops
    g = _g(2)
    n, h, w = 1, 11, 150
    x = bf(torch.randn(n, 32, h, w, generator=g))
    r = bf(torch.randn(n, 32, h, w, generator=g))
    wt = bf(torch.randn(32, 32, 3, 3, generator=g) * 0.08)
    b = torch.randn(32, generator=g) * 0.1
    ref = F.conv2d(x, wt, b, padding=1) + r
    a, c = ops.conv3x3_tc_nhwc([nhwc_bf16(x)], wt.cuda(), b.cuda(), residual=nhwc_bf16(r), split=(24, 8))
    close_bf16(nchw(a), ref[:, :24])
    close_bf16(nchw(c), ref[:, 24:])
    # fp32 output + DCN heads epilogue
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    w216 = bf(torch.randn(216, 32, 3, 3, generator=g) * 0.05)
    b216 = torch.randn(216, generator=g) * 0.05
    raw = F.conv2d(x, w216, b216, padding=1)
    off = 10 * torch.tanh(raw[:, :144]) + flow.flip(1).repeat(1, 72, 1, 1)
    msk = torch.sigmoid(raw[:, 144:])
    got = nchw(ops.conv3x3_tc_nhwc([nhwc_bf16(x)], w216.cuda(), b216.cuda(), act=L.ACT_DCN_HEAD, out_kind=L.TC_OUT_F32,
                                   flow=flow.permute(0, 2, 3, 1).contiguous().cuda(), head_split=144, head_mag=10.0))
    assert (got[:, :144] - off).abs().max().item() < 2e-3
    assert (got[:, 144:] - msk).abs().max().item() < 1e-4
    # pixel shuffle x4 to fp32 with LeakyReLU and x2 scale (upsample_post / dcn_3.upsample)
    x24 = bf(torch.randn(n, 24, h, w, generator=g))
    w64 = bf(torch.randn(64, 24, 3, 3, generator=g) * 0.1)
    b64 = torch.randn(64, generator=g) * 0.1
    ref4 = F.leaky_relu(F.pixel_shuffle(F.conv2d(x24, w64, b64, padding=1), 4), 0.1) * 2.0
    got4 = nchw(ops.conv3x3_tc_nhwc([nhwc_bf16(x24)], w64.cuda(), b64.cuda(), act=1, out_kind=L.TC_OUT_SHUFFLE_F32,
                                    shuffle_r=4, post_scale=2.0))
    assert (got4 - ref4).abs().max().item() < 1e-4


def test_flow_warp_bf16():
    import ctypes as C
    from crfp_b200 import _lib as L
    from oracle import crfp_oracle as O
    g = _g(3)
    n, c, h, w = 2, 32, 20, 36
    x = bf(torch.randn(n, c, h, w, generator=g))
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    ref = O.flow_warp(x, flow)
    xd = nhwc_bf16(x)
    fd = flow.permute(0, 2, 3, 1).contiguous().cuda()
    out = torch.zeros_like(xd)
    d = L.WarpDesc(n=n, h=h, w=w, c=c, x=xd.data_ptr(), x_cstride=c, x_coffset=0, flow=fd.data_ptr(),
                   out=out.data_ptr(), out_cstride=c, out_coffset=0, border=0)
    L.check(L.lib().crfp_flow_warp_bf16_fwd(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "warp bf16")
    close_bf16(nchw(out), ref)


def test_dcn_v2_tc():
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_dcn_tc
    from oracle import crfp_oracle as O
    g = _g(4)
    n, h, w = 2, 21, 37
    x = bf(torch.randn(n, 32, h, w, generator=g))
    off = torch.randn(n, 144, h, w, generator=g) * 4
    off[0, :, :2] = 40.0
    off[0, :, 2:4] = 0.0
    msk = torch.rand(n, 72, h, w, generator=g)
    wt = bf(torch.randn(32, 32, 3, 3, generator=g) * 0.05)
    b = torch.randn(32, generator=g) * 0.05
    ref = O.dcn_v2(x, off, msk, wt, b, 8)
    xd = nhwc_bf16(x)
    om = torch.cat([off, msk], 1).permute(0, 2, 3, 1).contiguous().cuda()       # fused heads tensor (216 ch fp32)
    wp, bp = pack_dcn_tc(wt.cuda(), b.cuda(), 8)
    out = torch.zeros(n, h, w, 32, device="cuda", dtype=torch.bfloat16)
    d = L.DcnDesc(n=n, h=h, w=w, c=32, cout=32, dg=8, shared_taps=0, x=xd.data_ptr(), x_cstride=32, x_coffset=0,
                  offset=om.data_ptr(), off_cstride=216, off_coffset=0, mask=om.data_ptr(), mask_cstride=216,
                  mask_coffset=144, weight=wp.data_ptr(), bias=bp.data_ptr(), out=out.data_ptr(), out_cstride=32,
                  out_coffset=0)
    L.check(L.lib().crfp_dcn_v2_tc_fwd(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dcn tc")
    # the modulated columns are rounded to bf16 before the contraction: ~2^-9 relative per term
    close_bf16(nchw(out), ref, extra=4e-3)


def test_dcn_v2_tc3_fp32_accuracy():
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_dcn_tc3
    from oracle import crfp_oracle as O
    g = _g(5)
    n, h, w = 2, 21, 37
    x = torch.randn(n, 32, h, w, generator=g)
    off = torch.randn(n, 144, h, w, generator=g) * 4
    off[0, :, :2] = 40.0
    off[0, :, 2:4] = 0.0
    msk = torch.rand(n, 72, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(32, generator=g) * 0.05
    ref = O.dcn_v2(x, off, msk, wt, b, 8)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    om = torch.cat([off, msk], 1).permute(0, 2, 3, 1).contiguous().cuda()
    hi, lo, bp = pack_dcn_tc3(wt.cuda(), b.cuda(), 8)
    out = torch.zeros(n, h, w, 32, device="cuda")
    d = L.DcnDesc(n=n, h=h, w=w, c=32, cout=32, dg=8, shared_taps=0, x=xd.data_ptr(), x_cstride=32, x_coffset=0,
                  offset=om.data_ptr(), off_cstride=216, off_coffset=0, mask=om.data_ptr(), mask_cstride=216,
                  mask_coffset=144, weight=hi.data_ptr(), bias=bp.data_ptr(), out=out.data_ptr(), out_cstride=32,
                  out_coffset=0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), None, st), "dcn tc3")
    err = (nchw(out) - ref).abs().max().item()
    print(f"dcn tc3 max-abs {err:.3e}")
    assert err < 1e-4
    # the flow hint only moves the shared-memory sampling window: any hint (good, zero, wild) gives the same result
    for scale in (0.0, 3.0, 60.0):
        hint = (torch.randn(n, h, w, 2, generator=g) * scale).cuda()
        out2 = torch.zeros_like(out)
        d.out = out2.data_ptr()
        L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), hint.data_ptr(), st), "dcn tc3 hint")
        assert (nchw(out2) - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("c_list,cout,hw,act,extra", [
    ([32], 32, (9, 128), 1, 0), ([32], 32, (21, 300), 1, 0), ([32, 32], 32, (12, 130), 1, 2), ([24], 64, (16, 140), 0, 0),
    ([32], 216, (7, 200), 0, 0), ([64], 96, (5, 64), 2, 0), ([32, 32], 32, (6, 70), 1, 0), ([64], 128, (6, 20), 2, 0)])
def test_conv_tc3_fp32_accuracy(c_list, cout, hw, act, extra):
    """3 x bf16 split tensor-core conv on fp32 data: must be fp32-grade (<= 2e-4 abs on O(1) outputs)."""
    from crfp_b200 import ops
    g = _g(11)
    h, w = hw
    n = 2
    srcs = [torch.randn(n, c, h, w, generator=g) for c in c_list]
    ex = torch.randn(n, extra, h, w, generator=g) * 3 if extra else None
    cin = sum(c_list) + extra
    wt = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    allsrc = srcs + ([ex] if extra else [])
    ref = F.conv2d(torch.cat(allsrc, 1), wt, b, padding=1)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else F.relu(ref) if act == 2 else ref
    got = nchw(ops.conv3x3_tc3_nhwc([s.permute(0, 2, 3, 1).contiguous().cuda() for s in srcs], wt.cuda(), b.cuda(), act=act,
                                    extra=None if ex is None else ex.permute(0, 2, 3, 1).contiguous().cuda()))
    err = (got - ref).abs().max().item()
    print(f"tc3 {c_list}->{cout}: max-abs {err:.3e}")
    assert err < 2e-4


def test_conv_tc3_epilogues():
    from crfp_b200 import _lib as L
    from crfp_b200 import ops
    g = _g(12)
    n, h, w = 1, 11, 150
    f32 = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    x = torch.randn(n, 32, h, w, generator=g)
    r = torch.randn(n, 32, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.08
    b = torch.randn(32, generator=g) * 0.1
    ref = F.conv2d(x, wt, b, padding=1) + r
    a, c = ops.conv3x3_tc3_nhwc([f32(x)], wt.cuda(), b.cuda(), residual=f32(r), split=(24, 8))
    assert (nchw(a) - ref[:, :24]).abs().max().item() < 2e-4 and (nchw(c) - ref[:, 24:]).abs().max().item() < 2e-4
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    w216 = torch.randn(216, 32, 3, 3, generator=g) * 0.05
    b216 = torch.randn(216, generator=g) * 0.05
    raw = F.conv2d(x, w216, b216, padding=1)
    off = 10 * torch.tanh(raw[:, :144]) + flow.flip(1).repeat(1, 72, 1, 1)
    msk = torch.sigmoid(raw[:, 144:])
    got = nchw(ops.conv3x3_tc3_nhwc([f32(x)], w216.cuda(), b216.cuda(), act=L.ACT_DCN_HEAD, flow=f32(flow), head_split=144))
    assert (got[:, :144] - off).abs().max().item() < 1e-3 and (got[:, 144:] - msk).abs().max().item() < 1e-4
    x24 = torch.randn(n, 24, h, w, generator=g)
    w64 = torch.randn(64, 24, 3, 3, generator=g) * 0.1
    b64 = torch.randn(64, generator=g) * 0.1
    ref4 = F.leaky_relu(F.pixel_shuffle(F.conv2d(x24, w64, b64, padding=1), 4), 0.1) * 2.0
    got4 = nchw(ops.conv3x3_tc3_nhwc([f32(x24)], w64.cuda(), b64.cuda(), act=1, shuffle_r=4, post_scale=2.0))
    assert (got4 - ref4).abs().max().item() < 2e-4
    # pixel shuffle x2 (upsample: 32 -> 96 -> 24 channels): the vectorised sub-pixel store
    w96 = torch.randn(96, 32, 3, 3, generator=g) * 0.08
    b96 = torch.randn(96, generator=g) * 0.1
    ref2 = F.pixel_shuffle(F.conv2d(x, w96, b96, padding=1), 2)
    got2 = nchw(ops.conv3x3_tc3_nhwc([f32(x)], w96.cuda(), b96.cuda(), shuffle_r=2))
    assert got2.shape == ref2.shape and (got2 - ref2).abs().max().item() < 2e-4


def test_conv_tc3_unshuffle_source():
    """PixelUnShufflePack_v2 (pixel_unshuffle(4) + conv 64->32, CRFP.py:239-279) read straight from the 4-channel HR plane."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_conv_tc3
    g = _g(21)
    n, h, w = 2, 13, 70
    s_hr = torch.randn(n, 4, 4 * h, 4 * w, generator=g)
    wt = torch.randn(32, 64, 3, 3, generator=g) * 0.06
    b = torch.randn(32, generator=g) * 0.1
    ref = F.conv2d(F.pixel_unshuffle(s_hr, 4), wt, b, padding=1)
    x = s_hr.permute(0, 2, 3, 1).contiguous().cuda()
    hi, lo, bp, _ = pack_conv_tc3(wt.cuda(), b.cuda(), [64], modes=[1])
    out = torch.zeros(n, h, w, 32, device="cuda")
    d = L.ConvTc3Desc()
    d.n, d.h, d.w, d.nsrc = n, h, w, 1
    d.src[0] = L.TcSrc(ptr=x.data_ptr(), c=64, cstride=4, coffset=0, _pad=1)
    d.cout, d.act = 32, 0
    d.weight_hi, d.weight_lo, d.bias = hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
    d.post_scale, d.out_kind, d.ndst = 1.0, L.TC_OUT_F32, 1
    d.dst[0] = L.TcSrc(ptr=out.data_ptr(), c=32, cstride=32, coffset=0)
    L.check(L.lib().crfp_conv3x3_tc3_fwd(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc3 unshuffle")
    assert (nchw(out) - ref).abs().max().item() < 2e-4


def _dcn_desc(L, n, h, w, xd, om, wptr, bptr, out):
    return L.DcnDesc(n=n, h=h, w=w, c=32, cout=32, dg=8, shared_taps=0, x=xd.data_ptr(), x_cstride=32, x_coffset=0,
                     offset=om.data_ptr(), off_cstride=216, off_coffset=0, mask=om.data_ptr(), mask_cstride=216,
                     mask_coffset=144, weight=wptr, bias=bptr, out=out.data_ptr(), out_cstride=32, out_coffset=0)


def test_production_align_kernel_dumps_its_own_indices():
    """Exact integer sampling indices FROM THE KERNEL THAT PRODUCES THE OUTPUT (dcn_tc3_ws_kernel, the persistent
    TMA + tcgen05 align kernel): floor(py), floor(px) of every (pixel, group, tap) it samples equal floor() of the
    reference positions computed in torch fp32 (y-1+i + dy as one fp32 add), incl. out-of-range and on-integer offsets."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_dcn_tc3
    g = _g(21)
    n, h, w = 2, 37, 53
    x = torch.randn(n, 32, h, w, generator=g)
    off = torch.randn(n, 144, h, w, generator=g) * 4
    off[0, :, :2] = 40.0       # far outside
    off[0, :, 2:4] = 0.0       # exactly on integer positions
    off[1, :, :, :3] = -1.0    # exactly on the -1 border
    off[1, :, 5:7] = torch.round(off[1, :, 5:7]) + 0.9999999
    msk = torch.rand(n, 72, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(32, generator=g) * 0.05
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    om = torch.cat([off, msk], 1).permute(0, 2, 3, 1).contiguous().cuda()
    hi, lo, bp = pack_dcn_tc3(wt.cuda(), b.cuda(), 8)
    out = torch.zeros(n, h, w, 32, device="cuda")
    y0 = torch.full((n, h, w, 72), -12345, device="cuda", dtype=torch.int32)
    x0 = torch.full_like(y0, -12345)
    d = _dcn_desc(L, n, h, w, xd, om, hi.data_ptr(), bp.data_ptr(), out)
    d.dbg_y0, d.dbg_x0 = y0.data_ptr(), x0.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    hint = (torch.randn(n, h, w, 2, generator=g) * 3).cuda()
    L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), hint.data_ptr(), st), "dcn tc3 + index dump")
    ys = torch.arange(h).view(1, h, 1, 1).float()
    xs = torch.arange(w).view(1, 1, w, 1).float()
    t = torch.arange(72) % 9
    offp = off.permute(0, 2, 3, 1)
    ry0 = torch.floor((ys - 1 + (t // 3).float()) + offp[..., 0::2]).int()
    rx0 = torch.floor((xs - 1 + (t % 3).float()) + offp[..., 1::2]).int()
    assert torch.equal(y0.cpu(), ry0) and torch.equal(x0.cpu(), rx0)
    # and the side kernel (crfp_dcn_v2_indices) agrees with the production kernel
    y1, x1 = torch.empty_like(y0), torch.empty_like(x0)
    L.check(L.lib().crfp_dcn_v2_indices(C.byref(d), y1.data_ptr(), x1.data_ptr(), st), "indices")
    assert torch.equal(y0, y1) and torch.equal(x0, x1)
    from oracle import crfp_oracle as O
    assert (nchw(out) - O.dcn_v2(x, off, msk, wt, b, 8)).abs().max().item() < 1e-4


def test_hr_align_kernel_dumps_its_own_indices():
    """dcn_hr_kernel (C=4, dg=1, one (dy,dx) per pixel shared by the 9 taps): indices dumped by the kernel itself."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200.packing import pack_dcn
    g = _g(22)
    n, h, w = 1, 40, 56
    x = torch.randn(n, 4, h, w, generator=g)
    om = torch.randn(n, 2, h, w, generator=g) * 5
    om[0, :, :2] = 0.0
    m1 = torch.rand(n, 1, h, w, generator=g)
    wt = torch.randn(4, 4, 3, 3, generator=g) * 0.2
    b = torch.randn(4, generator=g) * 0.1
    wp, bp = pack_dcn(wt.cuda(), b.cuda(), 1)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    omd = torch.cat([om, m1], 1).permute(0, 2, 3, 1).contiguous().cuda()      # (dy, dx, m) per pixel
    out = torch.zeros(n, h, w, 4, device="cuda")
    y0 = torch.full((n, h, w, 9), -12345, device="cuda", dtype=torch.int32)
    x0 = torch.full_like(y0, -12345)
    d = L.DcnDesc(n=n, h=h, w=w, c=4, cout=4, dg=1, shared_taps=1, x=xd.data_ptr(), x_cstride=4, x_coffset=0,
                  offset=omd.data_ptr(), off_cstride=3, off_coffset=0, mask=omd.data_ptr(), mask_cstride=3, mask_coffset=2,
                  weight=wp.data_ptr(), bias=bp.data_ptr(), out=out.data_ptr(), out_cstride=4, out_coffset=0)
    d.dbg_y0, d.dbg_x0 = y0.data_ptr(), x0.data_ptr()
    L.check(L.lib().crfp_dcn_v2_fwd(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dcn hr + dump")
    ys = torch.arange(h).view(1, h, 1, 1).float()
    xs = torch.arange(w).view(1, 1, w, 1).float()
    t = torch.arange(9)
    omp = om.permute(0, 2, 3, 1)
    ry0 = torch.floor((ys - 1 + (t // 3).float()) + omp[..., 0:1]).int()
    rx0 = torch.floor((xs - 1 + (t % 3).float()) + omp[..., 1:2]).int()
    assert torch.equal(y0.cpu(), ry0) and torch.equal(x0.cpu(), rx0)
    from oracle import crfp_oracle as O
    ref = O.dcn_v2(x, om.repeat(1, 9, 1, 1), m1.repeat(1, 9, 1, 1), wt, b, 1)
    assert (nchw(out) - ref).abs().max().item() < 1e-4


def test_align_kernel_raw_heads_equals_epilogue_heads():
    """crfp_dcn_desc.head_raw: the sampler applies 10*tanh + flow / sigmoid to the RAW head-conv outputs
    (DCN_module.forward, CRFP.py:337-349) with the conv epilogue's own arithmetic -> bit-identical to the two-pass
    form, sampled indices included; and within 1e-4 of the oracle's DCN_module tail."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200 import ops
    from crfp_b200.packing import pack_dcn_tc3
    from oracle import crfp_oracle as O
    g = _g(23)
    n, h, w = 1, 26, 150
    z = torch.randn(n, 32, h, w, generator=g)
    x = torch.randn(n, 32, h, w, generator=g)
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    w_off = torch.randn(144, 32, 3, 3, generator=g) * 0.05
    w_msk = torch.randn(72, 32, 3, 3, generator=g) * 0.05
    b_off = torch.randn(144, generator=g) * 0.05
    b_msk = torch.randn(72, generator=g) * 0.05
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(32, generator=g) * 0.05
    zd, xd = z.permute(0, 2, 3, 1).contiguous().cuda(), x.permute(0, 2, 3, 1).contiguous().cuda()
    fd = flow.permute(0, 2, 3, 1).contiguous().cuda()
    wh = torch.cat([w_off, w_msk], 0).cuda()
    bh = torch.cat([b_off, b_msk], 0).cuda()
    om_act = ops.conv3x3_tc3_nhwc([zd], wh, bh, act=L.ACT_DCN_HEAD, flow=fd, head_split=144, head_mag=10.0)
    om_raw = ops.conv3x3_tc3_nhwc([zd], wh, bh, act=L.ACT_NONE)
    hi, lo, bp = pack_dcn_tc3(wt.cuda(), b.cuda(), 8)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs, idx = [], []
    for om, raw in ((om_act, 0), (om_raw, 1)):
        out = torch.zeros(n, h, w, 32, device="cuda")
        y0 = torch.zeros(n, h, w, 72, device="cuda", dtype=torch.int32)
        x0 = torch.zeros_like(y0)
        d = _dcn_desc(L, n, h, w, xd, om, hi.data_ptr(), bp.data_ptr(), out)
        d.dbg_y0, d.dbg_x0 = y0.data_ptr(), x0.data_ptr()
        if raw:
            d.head_raw, d.head_flow, d.head_mag = 1, fd.data_ptr(), 10.0
        L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), fd.data_ptr(), st), "dcn tc3")
        outs.append(out); idx.append((y0, x0))
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(idx[0][0], idx[1][0]) and torch.equal(idx[0][1], idx[1][1])
    # oracle tail of DCN_module: offset = 10 tanh(conv) + flow.flip(1).repeat, mask = sigmoid(conv), DCNv2
    off = 10.0 * torch.tanh(F.conv2d(z, w_off, b_off, padding=1)) + flow.flip(1).repeat(1, 72, 1, 1)
    msk = torch.sigmoid(F.conv2d(z, w_msk, b_msk, padding=1))
    ref = O.dcn_v2(x, off, msk, wt, b, 8)
    err = (nchw(outs[1]) - ref).abs().max().item()
    print(f"fused-activation align kernel vs oracle DCN_module tail: max-abs {err:.3e}")
    assert err < 2e-4
    # the raw mode needs the flow; the SIMT kernels do not implement it
    d.head_flow = None
    assert L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), fd.data_ptr(), st) < 0
    d.head_flow = fd.data_ptr()
    assert L.lib().crfp_dcn_v2_fwd(C.byref(d), st) < 0


@pytest.mark.parametrize("n,h,w", [(1, 16, 8), (1, 26, 150), (2, 37, 53), (1, 48, 64)])
def test_dcn_align_fused(n, h, w):
    """crfp_dcn_align_fused (heads conv + activations + DCNv2 in one kernel, offsets never in HBM) against the oracle's
    DCN_module tail (CRFP.py:337-350) and against the two-kernel path (same arithmetic -> same sampled integer indices)."""
    import ctypes as C
    from crfp_b200 import _lib as L
    from crfp_b200 import ops
    from crfp_b200.packing import pack_dcn_tc3
    from oracle import crfp_oracle as O
    g = _g(31 + h)
    z = torch.randn(n, 32, h, w, generator=g)
    x = torch.randn(n, 32, h, w, generator=g)
    flow = torch.randn(n, 2, h, w, generator=g) * 3
    flow[0, :, : h // 3] *= 6.0            # large flows: window centred by the flow, global fallback beyond it
    w_off = torch.randn(144, 32, 3, 3, generator=g) * 0.05
    w_msk = torch.randn(72, 32, 3, 3, generator=g) * 0.05
    b_off = torch.randn(144, generator=g) * 0.05
    b_msk = torch.randn(72, generator=g) * 0.05
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(32, generator=g) * 0.05
    f32 = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    zd, xd, fd = f32(z), f32(x), f32(flow)
    out, y0, x0 = ops.dcn_align_fused_nhwc(zd, fd, xd, w_off.cuda(), b_off.cuda(), w_msk.cuda(), b_msk.cuda(), wt.cuda(), b.cuda(),
                                           indices=True)
    torch.cuda.synchronize()
    off = 10.0 * torch.tanh(F.conv2d(z, w_off, b_off, padding=1)) + flow.flip(1).repeat(1, 72, 1, 1)
    msk = torch.sigmoid(F.conv2d(z, w_msk, b_msk, padding=1))
    ref = O.dcn_v2(x, off, msk, wt, b, 8)
    err = (nchw(out) - ref).abs().max().item()
    print(f"dcn_align_fused {n}x{h}x{w}: max-abs vs oracle DCN_module tail {err:.3e}")
    assert err < 2e-4
    # the two-kernel tensor-core path: raw heads conv + align kernel with the activations in its sampler
    om_raw = ops.conv3x3_tc3_nhwc([zd], torch.cat([w_off, w_msk], 0).cuda(), torch.cat([b_off, b_msk], 0).cuda(), act=L.ACT_NONE)
    hi, lo, bp = pack_dcn_tc3(wt.cuda(), b.cuda(), 8)
    out2 = torch.zeros(n, h, w, 32, device="cuda")
    y1 = torch.zeros(n, h, w, 72, device="cuda", dtype=torch.int32)
    x1 = torch.zeros_like(y1)
    d = _dcn_desc(L, n, h, w, xd, om_raw, hi.data_ptr(), bp.data_ptr(), out2)
    d.head_raw, d.head_flow, d.head_mag = 1, fd.data_ptr(), 10.0
    d.dbg_y0, d.dbg_x0 = y1.data_ptr(), x1.data_ptr()
    L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), fd.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "dcn tc3")
    torch.cuda.synchronize()
    same_idx = (y0 == y1).float().mean().item(), (x0 == x1).float().mean().item()
    print(f"   vs two-kernel path: max-abs {(out - out2).abs().max().item():.3e}, identical sampled indices {same_idx}")
    assert (out - out2).abs().max().item() < 1e-4
    assert min(same_idx) > 0.9999          # the head values differ by accumulation order at most: floor() may flip on a tie
    assert int((y0 == -12345).sum()) == 0 and int((x0 == -12345).sum()) == 0   # every sample of every pixel was taken
