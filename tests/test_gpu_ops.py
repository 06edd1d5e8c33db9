"""GPU parity tests (-m gpu) at the operator boundary: every kernel called through the C ABI against the oracle
(plain PyTorch fp32 on the CPU) on the same seeded inputs.  Tolerances: fp32 max-abs <= 1e-4 per op (the
end-to-end budget of BASELINE.json is 1e-3); integer sampling indices must agree exactly."""
import pytest
import torch
import torch.nn.functional as F

from oracle import crfp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from crfp_b200 import ops as _ops
    return _ops


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous().cpu()


def _g(seed):
    return torch.Generator().manual_seed(seed)


def test_layout_roundtrip(ops):
    x = torch.randn(2, 3, 9, 13, generator=_g(0))
    y = ops.to_nhwc(x.cuda(), 4)
    assert torch.equal(y[..., :3].cpu(), x.permute(0, 2, 3, 1)) and y[..., 3].abs().sum() == 0
    assert torch.equal(ops.to_nchw(y, 3).cpu(), x)


@pytest.mark.parametrize("c_list,cout,hw,act", [
    ([32], 32, (40, 70), 1), ([32, 32], 32, (17, 33), 1), ([32, 32, 2], 32, (24, 40), 1), ([3, 3], 32, (16, 24), 2),
    ([24], 32, (20, 20), 0), ([64], 128, (9, 11), 2), ([256], 256, (5, 6), 2), ([32], 96, (12, 36), 0),
    ([128], 64, (33, 65), 2)])
def test_conv_wide(ops, c_list, cout, hw, act):
    g = _g(1)
    h, w = hw
    srcs = [torch.randn(2, c, h, w, generator=g) for c in c_list]
    wt = torch.randn(cout, sum(c_list), 3, 3, generator=g) * (2.0 / (9 * sum(c_list))) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(torch.cat(srcs, 1), wt, b, padding=1)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else F.relu(ref) if act == 2 else ref
    got = nchw(ops.conv3x3_nhwc([nhwc(s) for s in srcs], wt.cuda(), b.cuda(), act=act))
    assert (got - ref).abs().max().item() < 1e-4


def test_conv_wide_residual_split_shuffle_unshuffle(ops):
    g = _g(2)
    x = torch.randn(1, 32, 22, 38, generator=g)
    r = torch.randn(1, 32, 22, 38, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.08
    b = torch.randn(32, generator=g) * 0.1
    ref = F.conv2d(x, wt, b, padding=1) + r
    a, c = ops.conv3x3_nhwc([nhwc(x)], wt.cuda(), b.cuda(), residual=nhwc(r), split=(24, 8))
    assert (nchw(a) - ref[:, :24]).abs().max().item() < 1e-4 and (nchw(c) - ref[:, 24:]).abs().max().item() < 1e-4
    # PixelShufflePack x2 (32->96) and x4 (24->64) with LeakyReLU and post-scale
    w2 = torch.randn(96, 32, 3, 3, generator=g) * 0.08
    b2 = torch.randn(96, generator=g) * 0.1
    ref2 = F.pixel_shuffle(F.conv2d(x, w2, b2, padding=1), 2)
    assert (nchw(ops.conv3x3_nhwc([nhwc(x)], w2.cuda(), b2.cuda(), shuffle_r=2)) - ref2).abs().max().item() < 1e-4
    x24 = torch.randn(1, 24, 10, 12, generator=g)
    w4 = torch.randn(64, 24, 3, 3, generator=g) * 0.1
    b4 = torch.randn(64, generator=g) * 0.1
    ref4 = F.leaky_relu(F.pixel_shuffle(F.conv2d(x24, w4, b4, padding=1), 4), 0.1) * 2.0
    got4 = nchw(ops.conv3x3_nhwc([nhwc(x24)], w4.cuda(), b4.cuda(), act=1, shuffle_r=4, post_scale=2.0))
    assert (got4 - ref4).abs().max().item() < 1e-4
    # PixelUnShufflePack_v2: pixel_unshuffle(4) + conv 64->32 read straight from the HR plane
    s = torch.randn(2, 4, 32, 48, generator=g)
    wd = torch.randn(32, 64, 3, 3, generator=g) * 0.06
    bd = torch.randn(32, generator=g) * 0.1
    refd = F.conv2d(F.pixel_unshuffle(s, 4), wd, bd, padding=1)
    gotd = nchw(ops.conv3x3_nhwc([nhwc(s)], wd.cuda(), bd.cuda(), modes=[1]))
    assert (gotd - refd).abs().max().item() < 1e-4


def test_conv_dcn_heads_epilogue(ops):
    g = _g(3)
    z = torch.randn(1, 32, 18, 26, generator=g)
    flow = torch.randn(1, 2, 18, 26, generator=g) * 3
    wo = torch.randn(144, 32, 3, 3, generator=g) * 0.05
    bo = torch.randn(144, generator=g) * 0.05
    wm = torch.randn(72, 32, 3, 3, generator=g) * 0.05
    bm = torch.randn(72, generator=g) * 0.05
    off = 10 * torch.tanh(F.conv2d(z, wo, bo, padding=1)) + flow.flip(1).repeat(1, 72, 1, 1)
    msk = torch.sigmoid(F.conv2d(z, wm, bm, padding=1))
    got = nchw(ops.conv3x3_nhwc([nhwc(z)], torch.cat([wo, wm]).cuda(), torch.cat([bo, bm]).cuda(), act=3,
                                flow=nhwc(flow), head_split=144, head_mag=10.0))
    assert (got[:, :144] - off).abs().max().item() < 1e-4
    assert (got[:, 144:] - msk).abs().max().item() < 1e-5


@pytest.mark.parametrize("c_list,cout,act", [([4], 4, 1), ([4, 4], 4, 1), ([4, 4, 2], 4, 1), ([6], 4, 1), ([4], 3, 0),
                                             ([32], 2, 4)])
def test_conv_thin(ops, c_list, cout, act):
    g = _g(4)
    h, w = 37, 45
    srcs = [torch.randn(2, c, h, w, generator=g) for c in c_list]
    wt = torch.randn(cout, sum(c_list), 3, 3, generator=g) * 0.15
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(torch.cat(srcs, 1), wt, b, padding=1)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else torch.tanh(ref) * 256 if act == 4 else ref
    res = torch.randn(2, cout, h, w, generator=g) if act == 1 else None
    if res is not None:
        ref = ref + res
    got = nchw(ops.conv3x3_nhwc([nhwc(s) for s in srcs], wt.cuda(), b.cuda(), act=act,
                                residual=None if res is None else nhwc(res)))
    tol = 2e-3 if act == 4 else 1e-4   # tanh*256 amplifies fp32 ulps
    assert (got - ref).abs().max().item() < tol


@pytest.mark.parametrize("c,hw,scale", [(32, (36, 64), 2.0), (4, (64, 96), 6.0), (8, (20, 28), 1.0), (3, (9, 11), 3.0)])
def test_flow_warp_values_and_indices(ops, c, hw, scale):
    g = _g(5)
    h, w = hw
    x = torch.randn(2, c, h, w, generator=g)
    flow = torch.randn(2, 2, h, w, generator=g) * scale
    flow[0, :, : h // 2] = 0.0                      # zero flow rows: floor() sits on the fp32 round-trip edge
    flow[1, :, :, :3] = 1000.0                      # far out of range: zeros
    ref = O.flow_warp(x, flow)
    got = ops.flow_warp(x.cuda(), flow.permute(0, 2, 3, 1).contiguous().cuda()).cpu()
    assert (got - ref).abs().max().item() < 1e-5
    x0, y0 = ops.flow_warp_indices(flow.permute(0, 2, 3, 1).contiguous().cuda())
    rx0, ry0 = O.flow_warp_indices(flow)
    assert torch.equal(x0.cpu(), rx0) and torch.equal(y0.cpu(), ry0)     # exact integer agreement


def test_flow_warp_border_and_size_check(ops):
    g = _g(6)
    x = torch.randn(1, 4, 12, 14, generator=g)
    flow = torch.randn(1, 2, 12, 14, generator=g) * 8
    ref = O.flow_warp(x, flow, padding_mode="border")
    got = ops.flow_warp(x.cuda(), flow.permute(0, 2, 3, 1).contiguous().cuda(), padding_mode="border").cpu()
    assert (got - ref).abs().max().item() < 1e-5
    with pytest.raises(ValueError):
        ops.flow_warp(x.cuda(), torch.zeros(1, 11, 14, 2).cuda())


def test_dcn_v2_l1_values_and_indices(ops):
    g = _g(7)
    n, h, w = 2, 21, 27
    x = torch.randn(n, 32, h, w, generator=g)
    off = torch.randn(n, 144, h, w, generator=g) * 4
    off[0, :, :2] = 40.0     # completely outside
    off[0, :, 2:4] = 0.0     # exactly on integer positions
    msk = torch.rand(n, 72, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) * 0.05
    b = torch.randn(32, generator=g) * 0.05
    ref = O.dcn_v2(x, off, msk, wt, b, 8)
    ref_naive = O.dcn_v2_naive(x, off, msk, wt, b, 8)
    assert (ref - ref_naive).abs().max().item() < 1e-4
    mod = ops.DCNv2(32, 32, 3, stride=1, padding=1, dilation=1, deformable_groups=8).cuda()
    mod.weight.data.copy_(wt)
    mod.bias.data.copy_(b)
    got = mod(x.cuda(), off.cuda(), msk.cuda()).cpu()
    assert (got - ref).abs().max().item() < 1e-4
    from crfp_b200.packing import pack_dcn
    wp, bp = pack_dcn(wt.cuda(), b.cuda(), 8)
    y0, x0 = ops.dcn_v2_nhwc(nhwc(x), nhwc(off), nhwc(msk), wp, bp, 8, 32, indices=True)
    ys = torch.arange(h).view(1, h, 1, 1).float()
    xs = torch.arange(w).view(1, 1, w, 1).float()
    t = torch.arange(72) % 9
    offp = off.permute(0, 2, 3, 1)
    ry0 = torch.floor((ys - 1 + (t // 3).float()) + offp[..., 0::2]).int()
    rx0 = torch.floor((xs - 1 + (t % 3).float()) + offp[..., 1::2]).int()
    assert torch.equal(y0.cpu(), ry0) and torch.equal(x0.cpu(), rx0)


def test_dcn_v2_hr_shared_offsets(ops):
    g = _g(8)
    n, h, w = 1, 40, 56
    x = torch.randn(n, 4, h, w, generator=g)
    om = torch.randn(n, 2, h, w, generator=g) * 5          # (dy, dx) shared by the 9 taps
    m1 = torch.rand(n, 1, h, w, generator=g)
    wt = torch.randn(4, 4, 3, 3, generator=g) * 0.2
    b = torch.randn(4, generator=g) * 0.1
    off18 = om.repeat(1, 9, 1, 1)                          # reference materialisation (CRFP.py:342-347)
    ref = O.dcn_v2(x, off18, m1.repeat(1, 9, 1, 1), wt, b, 1)
    from crfp_b200.packing import pack_dcn
    wp, bp = pack_dcn(wt.cuda(), b.cuda(), 1)
    got = nchw(ops.dcn_v2_nhwc(nhwc(x), nhwc(om), nhwc(m1), wp, bp, 1, 4, shared_taps=True))
    assert (got - ref).abs().max().item() < 1e-4
    got2 = nchw(ops.dcn_v2_nhwc(nhwc(x), nhwc(off18), nhwc(m1.repeat(1, 9, 1, 1)), wp, bp, 1, 4))
    assert (got2 - ref).abs().max().item() < 1e-4


def test_resize_and_avgpool(ops):
    g = _g(9)
    x = torch.randn(2, 5, 11, 13, generator=g)
    for s in (2, 8):
        ref = F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False) * s
        got = nchw(ops.resize_bilinear_nhwc(nhwc(x), 11 * s, 13 * s, 1.0 / s, 1.0 / s, float(s)))
        assert (got - ref).abs().max().item() < 1e-5
    x2 = torch.randn(1, 2, 16, 16, generator=g)
    ref = F.interpolate(x2, size=(18, 20), mode="bilinear", align_corners=False)
    got = nchw(ops.resize_bilinear_nhwc(nhwc(x2), 18, 20, 16 / 18, 16 / 20))
    assert (got - ref).abs().max().item() < 1e-5
    x3 = torch.randn(1, 1, 32, 48, generator=g)
    ref = F.interpolate(x3, scale_factor=0.25, mode="bilinear", align_corners=False)
    assert (nchw(ops.resize_bilinear_nhwc(nhwc(x3), 8, 12, 4.0, 4.0)) - ref).abs().max().item() < 1e-6
    x4 = torch.randn(2, 7, 9, 11, generator=g)
    assert (nchw(ops.avgpool2_nhwc(nhwc(x4))) - F.avg_pool2d(x4, 2, 2)).abs().max().item() < 1e-6


def test_bad_arguments_return_status_not_crash(ops):
    import ctypes as C
    from crfp_b200 import _lib as L
    d = L.ConvDesc()
    assert L.lib().crfp_conv3x3_fwd(C.byref(d), None) < 0
    assert L.lib().crfp_conv3x3_fwd(None, None) == -5
    dd = L.DcnDesc(n=1, h=4, w=4, c=16, cout=16, dg=4)
    assert L.lib().crfp_dcn_v2_fwd(C.byref(dd), None) < 0
