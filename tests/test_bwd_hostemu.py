"""CPU checks of the training kernels (crfp_b200/csrc/bwd.cu) through a host emulation of the SAME source
(tests/tools/hostemu: g++ + a serial CUDA shim, same C-ABI entry points on host pointers).  Gradients are compared
with torch autograd of the oracle's ops (the reference's own training path is ATen autograd + dcn_v2's backward).
The GPU twins of these tests are in tests/test_gpu_zz_training.py."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import hostemu
from crfp_b200 import _lib as L
from crfp_b200 import autograd as A
from oracle import crfp_oracle as O

K = hostemu.HostEmuKernelSet()
K_DIRECT = hostemu.HostEmuKernelSet(dgrad_as_conv=False)


def _g(seed):
    return torch.Generator().manual_seed(seed)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def test_emulation_exports_every_training_symbol():
    h = hostemu.lib()
    for name in L.TRAIN_SYMBOLS:
        assert hasattr(h, name)
    assert len(L.TRAIN_SYMBOLS) == 17
    # argument validation returns a status, never crashes
    assert h.crfp_conv3x3_bwd_data(1, 0, 4, 4, 4, 4, 0, None, None, None, None) == -1
    assert h.crfp_conv3x3_bwd_data(1, 4, 4, 4, 4, 6, 4, None, None, None, None) == -1      # slice outside cin_total
    assert h.crfp_conv3x3_bwd_data(1, 4, 4, 4, 4, 4, 0, None, None, None, None) == -5
    assert h.crfp_act_bwd(4, 0, None, None, None, None) == -2
    assert h.crfp_dcn_v2_bwd(None, None) == -5
    d = L.DcnBwdDesc(n=1, h=4, w=4, c=6, cout=4, dg=4)
    assert h.crfp_dcn_v2_bwd(C.byref(d), None) == -1


@pytest.mark.parametrize("c_list,cout,hw,act", [([32], 32, (9, 13), 1), ([32, 32, 2], 32, (7, 10), 1), ([3, 3], 16, (8, 8), 2),
                                                ([4, 4], 4, (11, 9), 0), ([6], 4, (6, 7), 1), ([4], 3, (5, 5), 0),
                                                ([24, 32, 8], 32, (6, 6), 1), ([8], 12, (1, 5), 0),
                                                ([4, 4], 4, (3, 100), 1), ([6], 4, (3, 70), 0),   # wide rows: column-segmented wgrad
                                                ([64], 64, (40, 8), 0),                          # several rows per wgrad chunk
                                                ([32], 32, (3, 100), 1)])                        # column segments, 4x4-tile mapping
@pytest.mark.parametrize("direct", [False, True])
def test_conv3x3_grads(c_list, cout, hw, act, direct):
    """direct=False: backward-data as a forward conv with the rotated / transposed kernel (the product default);
    direct=True: the gather kernels crfp_conv3x3_bwd_data (A/B path)."""
    K = K_DIRECT if direct else globals()["K"]
    g = _g(1)
    h, w = hw
    srcs = [torch.randn(2, c, h, w, generator=g, requires_grad=True) for c in c_list]
    wt = (torch.randn(cout, sum(c_list), 3, 3, generator=g) * 0.1).requires_grad_()
    b = (torch.randn(cout, generator=g) * 0.1).requires_grad_()
    ref = F.conv2d(torch.cat(srcs, 1), wt, b, padding=1)
    ref = F.leaky_relu(ref, 0.1) if act == 1 else F.relu(ref) if act == 2 else ref
    dy = torch.randn(ref.shape, generator=g)
    ref_grads = torch.autograd.grad(ref, [wt, b, *srcs], dy)
    s2 = [nhwc(s.detach()).requires_grad_() for s in srcs]
    w2, b2 = wt.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    out = A.conv3x3(K, w2, b2, s2, act)
    assert (nchw(out) - ref).abs().max().item() < 1e-5
    got = torch.autograd.grad(out, [w2, b2, *s2], nhwc(dy))
    assert (got[0] - ref_grads[0]).abs().max().item() < 2e-4
    assert (got[1] - ref_grads[1]).abs().max().item() < 2e-4
    for a, r in zip(got[2:], ref_grads[2:]):
        assert (nchw(a) - r).abs().max().item() < 1e-4


@pytest.mark.parametrize("c_list,cout,hw", [([4, 4], 4, (11, 9)), ([4, 4, 2], 4, (3, 100)), ([6], 4, (6, 70)), ([4], 3, (40, 33)),
                                            ([32], 32, (9, 13))])
def test_conv3x3_weight_gradient_two_stage(c_list, cout, hw):
    """CRFP_WGRAD_THIN=2stage: the thin layers' pixel chunks write partial sums to a workspace and a reduce kernel adds
    them up (no atomics on dw); wide layers (workspace query returns 0) keep the atomic path."""
    K2 = hostemu.HostEmuKernelSet(wgrad_two_stage=True)
    g = _g(8)
    h, w = hw
    srcs = [torch.randn(2, c, h, w, generator=g, requires_grad=True) for c in c_list]
    wt = (torch.randn(cout, sum(c_list), 3, 3, generator=g) * 0.1).requires_grad_()
    b = (torch.randn(cout, generator=g) * 0.1).requires_grad_()
    ref = F.leaky_relu(F.conv2d(torch.cat(srcs, 1), wt, b, padding=1), 0.1)
    dy = torch.randn(ref.shape, generator=g)
    rg = torch.autograd.grad(ref, [wt, b], dy)
    w2, b2 = wt.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    out = A.conv3x3(K2, w2, b2, [nhwc(s.detach()) for s in srcs], 1)
    got = torch.autograd.grad(out, [w2, b2], nhwc(dy))
    assert (got[0] - rg[0]).abs().max().item() < 1e-5 * rg[0].abs().max().item() + 2e-4
    assert (got[1] - rg[1]).abs().max().item() < 1e-5 * rg[1].abs().max().item() + 2e-4
    lib = hostemu.lib()
    c0 = c_list[0]      # thin = at most 64 threads' worth of weight elements (4x4 register tiles when both counts are % 4)
    elems = 9 * (c0 // 4) * (cout // 4) if (c0 % 4 == 0 and cout % 4 == 0) else 9 * c0 * cout
    assert (lib.crfp_conv3x3_bwd_weight_workspace(2, h, w, c0, cout) > 0) == (elems <= 64)


def test_conv3x3_partial_requires_grad():
    g = _g(2)
    x = torch.randn(1, 5, 6, 8, generator=g)
    wt = torch.randn(4, 8, 3, 3, generator=g).requires_grad_()
    b = torch.zeros(4)
    out = A.conv3x3(K, wt, b, [x], 0)              # frozen bias, input without grad (FNet-frozen phase, first layer)
    (dw,) = torch.autograd.grad(out.sum(), [wt])
    xr = nchw(x)
    ref = torch.autograd.grad(F.conv2d(xr, wt, b, padding=1).sum(), [wt])[0]
    assert (dw - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("c,dg,cout,hw", [(32, 8, 32, (9, 11)), (4, 1, 4, (12, 10)), (8, 2, 4, (5, 7))])
def test_dcn_v2_grads(c, dg, cout, hw):
    g = _g(3)
    n, (h, w) = 2, hw
    x = torch.randn(n, c, h, w, generator=g, requires_grad=True)
    off = (torch.randn(n, dg * 18, h, w, generator=g) * 2.5).requires_grad_()
    with torch.no_grad():
        off[0, :, :1] += 30.0                 # fully outside: zero value, zero gradients
    msk = torch.rand(n, dg * 9, h, w, generator=g).requires_grad_()
    wt = (torch.randn(cout, c, 3, 3, generator=g) * 0.1).requires_grad_()
    b = (torch.randn(cout, generator=g) * 0.1).requires_grad_()
    ref = O.dcn_v2(x, off, msk, wt, b, dg)
    dy = torch.randn(ref.shape, generator=g)
    rg = torch.autograd.grad(ref, [x, off, msk, wt, b], dy)
    x2, o2, m2 = (nhwc(t.detach()).requires_grad_() for t in (x, off, msk))
    w2, b2 = wt.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    out = A.dcn_v2(K, x2, o2, m2, w2, b2, dg)
    gg = torch.autograd.grad(out, [x2, o2, m2, w2, b2], nhwc(dy))
    for name, a, r in zip("x off mask".split(), gg[:3], rg[:3]):
        assert (nchw(a) - r).abs().max().item() < 2e-4, name
    assert (gg[3] - rg[3]).abs().max().item() < 5e-4
    assert (gg[4] - rg[4]).abs().max().item() < 5e-4


def test_dcn_v2_naive_restating_agrees_on_gradients():
    """the oracle's own naive DCNv2 restatement (published algorithm) gives the same gradients as the kernel"""
    g = _g(4)
    x = torch.randn(1, 8, 6, 7, generator=g, requires_grad=True)
    off = (torch.randn(1, 36, 6, 7, generator=g) * 1.5).requires_grad_()
    msk = torch.rand(1, 18, 6, 7, generator=g).requires_grad_()
    wt = (torch.randn(4, 8, 3, 3, generator=g) * 0.1).requires_grad_()
    b = torch.zeros(4, requires_grad=True)
    ref = O.dcn_v2_naive(x, off, msk, wt, b, 2)
    dy = torch.randn(ref.shape, generator=g)
    rg = torch.autograd.grad(ref, [x, off, msk], dy)
    x2, o2, m2 = (nhwc(t.detach()).requires_grad_() for t in (x, off, msk))
    out = A.dcn_v2(K, x2, o2, m2, wt.detach(), b.detach(), 2)
    gg = torch.autograd.grad(out, [x2, o2, m2], nhwc(dy))
    for a, r in zip(gg, rg):
        assert (nchw(a) - r).abs().max().item() < 2e-4


@pytest.mark.parametrize("c,hw,scale", [(32, (10, 14), 2.0), (4, (16, 12), 5.0), (24, (7, 9), 1.0)])
def test_flow_warp_grads(c, hw, scale):
    g = _g(5)
    h, w = hw
    x = torch.randn(2, c, h, w, generator=g, requires_grad=True)
    flow = (torch.randn(2, 2, h, w, generator=g) * scale).requires_grad_()
    with torch.no_grad():
        flow[1, :, :, :2] = 100.0
    ref = O.flow_warp(x, flow)
    dy = torch.randn(ref.shape, generator=g)
    rg = torch.autograd.grad(ref, [x, flow], dy)
    x2, f2 = nhwc(x.detach()).requires_grad_(), nhwc(flow.detach()).requires_grad_()
    out = A.flow_warp(K, x2, f2)
    gg = torch.autograd.grad(out, [x2, f2], nhwc(dy))
    assert (nchw(gg[0]) - rg[0]).abs().max().item() < 1e-5
    assert (nchw(gg[1]) - rg[1]).abs().max().item() < 1e-4
    (only_flow,) = torch.autograd.grad(A.flow_warp(K, x2.detach(), f2), [f2], nhwc(dy))
    assert torch.equal(only_flow, gg[1])


def test_resize_avgpool_grads():
    g = _g(6)
    x = torch.randn(2, 5, 6, 7, generator=g, requires_grad=True)
    for s, mul in ((2, 2.0), (8, 8.0), (2, 1.0)):
        ref = F.interpolate(x, scale_factor=s, mode="bilinear", align_corners=False) * mul
        dy = torch.randn(ref.shape, generator=g)
        (rg,) = torch.autograd.grad(ref, [x], dy)
        x2 = nhwc(x.detach()).requires_grad_()
        out = A.up_bilinear(K, x2, s, mul)
        assert (nchw(out) - ref).abs().max().item() < 1e-5
        (gg,) = torch.autograd.grad(out, [x2], nhwc(dy))
        assert (nchw(gg) - rg).abs().max().item() < 2e-6 * rg.abs().max().item()   # sums of up to 64*mul terms
    x3 = torch.randn(1, 2, 8, 8, generator=g, requires_grad=True)          # F.interpolate(size=) of FNet, 8 -> 9 x 11
    ref = F.interpolate(x3, size=(9, 11), mode="bilinear", align_corners=False)
    dy = torch.randn(ref.shape, generator=g)
    (rg,) = torch.autograd.grad(ref, [x3], dy)
    x4 = nhwc(x3.detach()).requires_grad_()
    (gg,) = torch.autograd.grad(A.resize_to(K, x4, 9, 11), [x4], nhwc(dy))
    assert (nchw(gg) - rg).abs().max().item() < 1e-5
    assert A.resize_to(K, x4, 8, 8) is x4
    for hw in ((6, 8), (7, 9)):                                              # odd sizes: trailing row / column dropped
        x5 = torch.randn(2, 3, *hw, generator=g, requires_grad=True)
        ref = F.avg_pool2d(x5, 2, 2)
        dy = torch.randn(ref.shape, generator=g)
        (rg,) = torch.autograd.grad(ref, [x5], dy)
        x6 = nhwc(x5.detach()).requires_grad_()
        (gg,) = torch.autograd.grad(A.avgpool2(K, x6), [x6], nhwc(dy))
        assert (nchw(gg) - rg).abs().max().item() < 1e-6


def test_charbonnier_and_adam():
    g = _g(7)
    pred = torch.randn(2, 3, 16, 16, generator=g, requires_grad=True)
    tgt = torch.randn(2, 3, 16, 16, generator=g)
    ref = torch.sqrt((pred - tgt) ** 2 + 1e-12).mean()
    (rg,) = torch.autograd.grad(ref, [pred])
    p2 = pred.detach().clone().requires_grad_()
    loss = A.charbonnier_loss(K, p2, tgt, 1e-12, 1.0)
    assert abs(loss.item() - ref.item()) < 1e-5
    (gg,) = torch.autograd.grad(loss, [p2])
    assert (gg - rg).abs().max().item() < 1e-8
    # Adam: 3 steps against torch.optim.Adam with the reference's settings (trainer.py:149, option.py:70-74)
    p = torch.randn(1000, generator=g)
    q = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([q], lr=2e-4, betas=(0.9, 0.999), eps=1e-12)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    h = hostemu.lib()
    for step in range(1, 4):
        grad = torch.randn(1000, generator=g) * (0.1 if step != 2 else 1e-9)
        q.grad = grad.clone()
        opt.step()
        bc1, bc2 = 1 - 0.9 ** step, 1 - 0.999 ** step
        assert h.crfp_adam_step(1000, p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), 0.9, 0.999, 1e-12,
                                2e-4 / bc1, bc2 ** 0.5, None) == 0
        assert (p - q.detach()).abs().max().item() < 3e-7      # 1 ulp of O(1) parameters


def test_opt_in_wgrad_variants_in_a_fresh_process():
    """CRFP_WGRAD_KX3=1 (sliding-window weight-gradient kernel) is read once per process: run the conv gradient tests
    again in a child process with the switch on."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CRFP_WGRAD_KX3="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-k", "test_conv3x3_grads",
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("nk,repeat", [(72, False), (9, True)])
def test_dcn_heads_activation_matches_the_torch_formulation(nk, repeat):
    """crfp_dcn_heads_act_fwd / _bwd == max_mag * tanh(offset heads) + flow.flip(-1) (repeated), sigmoid(mask heads) and
    their autograd (model/CRFP.py:337-347)."""
    from crfp_b200 import autograd as A
    K = hostemu.HostEmuKernelSet()
    g = _g(41)
    n, h, w = 2, 5, 7
    ch = 3 if repeat else 3 * nk
    heads = torch.randn(n, h, w, ch, generator=g).requires_grad_()
    flow = (torch.randn(n, h, w, 2, generator=g) * 3).requires_grad_()
    noff = 2 if repeat else 2 * nk
    off_ref = 10.0 * torch.tanh(heads[..., :noff])
    msk_ref = torch.sigmoid(heads[..., noff:])
    fyx = flow.flip(-1)
    if repeat:
        off_ref = (off_ref + fyx).repeat(1, 1, 1, nk)
        msk_ref = msk_ref.repeat(1, 1, 1, nk)
    else:
        off_ref = off_ref + fyx.repeat(1, 1, 1, nk)
    do, dm = torch.randn(off_ref.shape, generator=g), torch.randn(msk_ref.shape, generator=g)
    rg = torch.autograd.grad([off_ref, msk_ref], [heads, flow], [do, dm])
    h2, f2 = heads.detach().clone().requires_grad_(), flow.detach().clone().requires_grad_()
    off, msk = A.dcn_heads_act(K, h2, f2, nk, repeat, 10.0)
    assert (off - off_ref).abs().max().item() < 1e-5 and (msk - msk_ref).abs().max().item() < 1e-6
    got = torch.autograd.grad([off, msk], [h2, f2], [do, dm])
    assert (got[0] - rg[0]).abs().max().item() < 2e-5 and (got[1] - rg[1]).abs().max().item() < 1e-4


def test_fovea_blend_matches_the_torch_formulation():
    """crfp_fovea_blend_fwd / _bwd == F.leaky_relu(m * f + (1 - m) * s, 0.1) and its autograd (model/CRFP.py:1672-1675)."""
    from crfp_b200 import autograd as A
    K = hostemu.HostEmuKernelSet()
    g = _g(43)
    n, h, w, c = 2, 9, 11, 4
    f = torch.randn(n, h, w, c, generator=g).requires_grad_()
    s = torch.randn(n, h, w, c, generator=g).requires_grad_()
    m = (torch.rand(n, h, w, 1, generator=g) > 0.6).float()
    ref = F.leaky_relu(m * f + (1.0 - m) * s, 0.1)
    dy = torch.randn(ref.shape, generator=g)
    rg = torch.autograd.grad(ref, [f, s], dy)
    f2, s2 = f.detach().clone().requires_grad_(), s.detach().clone().requires_grad_()
    out = A.fovea_blend(K, f2, s2, m)
    assert (out - ref).abs().max().item() < 1e-6
    got = torch.autograd.grad(out, [f2, s2], dy)
    assert (got[0] - rg[0]).abs().max().item() < 1e-6 and (got[1] - rg[1]).abs().max().item() < 1e-6
