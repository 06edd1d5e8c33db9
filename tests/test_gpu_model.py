"""GPU parity tests (-m gpu) of the drop-in modules against the oracle and the reference's golden outputs.
Bar (BASELINE.json north_star): fp32 max-abs <= 1e-3 on the output tensor."""
import os

import pytest
import torch

from crfp_b200.synthetic import make_clip, make_state_dict
from oracle import crfp_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(seed=1)


@pytest.fixture(scope="module", params=["tc", "fp32"])
def model(request, sd):
    """Both precisions must meet the fp32 bar: 'tc' = tcgen05 3 x bf16 split contractions (default), 'fp32' = SIMT."""
    from crfp_b200 import CRFP_DSV
    m = CRFP_DSV("cuda", mid_channels=32, precision=request.param).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def _run(model, lrs, fvs, mks):
    out = model(lrs.cuda(), fvs.cuda(), mks.cuda())
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("name", ["dsv_n1_t3_16x24", "dsv_n2_t2_18x20", "dsv_n1_t1_8x8"])
def test_against_reference_golden(name, golden_dir, model):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    out = _run(model, lrs, fvs, mks)
    assert out.shape == fix["out"].shape
    err = (out - fix["out"]).abs().max().item()
    print(f"{name} [{model.precision}]: max-abs vs reference golden {err:.3e}")
    assert err <= TOL


@pytest.mark.parametrize("n,t,h,w,fv", [(1, 5, 32, 48, 96), (2, 3, 24, 40, 64), (1, 4, 45, 80, 96), (3, 2, 16, 16, 128)])
def test_against_oracle(model, sd, n, t, h, w, fv):
    lrs, fvs, mks, _ = make_clip(seed=21, n=n, t=t, h=h, w=w, fv_size=fv)
    taps = []
    ref = O.crfp_dsv_forward(sd, lrs, fvs, mks, taps=taps)
    out = _run(model, lrs, fvs, mks)
    errs = [(out[:, i] - ref[:, i]).abs().max().item() for i in range(t)]
    print(f"[{model.precision}] n={n} t={t} {h}x{w}: per-frame max-abs {['%.2e' % e for e in errs]}")
    assert max(errs) <= TOL


def test_mask_variants_and_skip_flag(model, sd):
    """Arbitrary boolean masks (quadrant, empty, full) and the exactness of the outside-fovea skip."""
    n, t, h, w = 1, 3, 16, 24
    lrs, fvs, mks, _ = make_clip(seed=5, n=n, t=t, h=h, w=w, fv_size=32)
    mks[:, 0] = False                                  # no fovea at all in frame 0
    mks[:, 1, :, : 4 * h, : 4 * w] = True              # a whole quadrant (DemoHscan, dataset/reds.py:197-198)
    mks[:, 2] = True                                   # everything is fovea
    fvs = torch.rand(n, t, 3, 8 * h, 8 * w, generator=torch.Generator().manual_seed(1)) * mks
    ref = O.crfp_dsv_forward(sd, lrs, fvs, mks)
    out = _run(model, lrs, fvs, mks)
    assert (out - ref).abs().max().item() <= TOL
    model.skip_outside_fovea = False
    out2 = _run(model, lrs, fvs, mks)
    model.skip_outside_fovea = True
    assert torch.equal(out, out2)


def test_patch_coords_entry_and_determinism(model, sd):
    n, t, h, w, fv = 1, 2, 16, 24, 48
    lrs, fvs, mks, fv_sp = make_clip(seed=9, n=n, t=t, h=h, w=w, fv_size=fv)
    patch = torch.stack([torch.stack([fvs[b, i, :, fv_sp[b, i, 0]:fv_sp[b, i, 0] + fv, fv_sp[b, i, 1]:fv_sp[b, i, 1] + fv]
                                      for i in range(t)]) for b in range(n)])
    a = _run(model, lrs, fvs, mks)
    b = model.forward_patch(lrs.cuda(), patch.cuda(), fv_sp).cpu()
    assert torch.equal(a, b)                           # exact agreement on the integer fovea rectangle
    assert torch.equal(a, _run(model, lrs, fvs, mks))  # bitwise deterministic


def test_out_host_streaming_copy(model):
    lrs, fvs, mks, _ = make_clip(seed=17, n=2, t=3, h=16, w=24, fv_size=48)
    ref = _run(model, lrs, fvs, mks)
    host = torch.empty(ref.shape, dtype=torch.float32).pin_memory()
    out = model(lrs.cuda(), fvs.cuda(), mks.cuda(), out_host=host)
    torch.cuda.synchronize()
    assert torch.equal(host, ref) and torch.equal(out.cpu(), ref)


def test_out_host_uint8_frames(model):
    """out_host may be a pinned uint8 tensor: frames quantised on the device as the reference saves them
    ((sr * 255).clip(0, 255).round(), trainer.py:446-474) — eager and graph replay, n = 1 and n = 2."""
    for n in (1, 2):
        lrs, fvs, mks, _ = make_clip(seed=18, n=n, t=5, h=16, w=24, fv_size=48)
        ref = _run(model, lrs, fvs, mks)
        want = (ref * 255.0).clip(0.0, 255.0).round().to(torch.uint8)
        host = torch.zeros(ref.shape, dtype=torch.uint8).pin_memory()
        a, b, c = lrs.cuda(), fvs.cuda(), mks.cuda()
        for _ in range(3):                       # eager, capture, replay
            host.zero_()
            out = model(a, b, c, out_host=host)
            torch.cuda.synchronize()
            assert torch.equal(host, want) and torch.equal(out.cpu(), ref)


def test_clip_batch_equals_single_clips(model):
    """Clips are independent (the multi-GPU sharding unit): a batch of 3 equals three batch-1 runs bit for bit."""
    lrs, fvs, mks, _ = make_clip(seed=13, n=3, t=3, h=16, w=24, fv_size=48)
    full = _run(model, lrs, fvs, mks)
    for b in range(3):
        one = _run(model, lrs[b:b + 1], fvs[b:b + 1], mks[b:b + 1])
        assert torch.equal(full[b:b + 1], one)


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_streaming_module_matches_reference_golden(golden_dir, sd, precision):
    from crfp_b200 import MRCF_simple_v18
    fix = torch.load(os.path.join(golden_dir, "stream_n1_t3_16x24.pt"))
    c = fix["case"]
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    fgs = torch.ones(c["n"], c["t"], 1, 8 * c["h"], 8 * c["w"])
    fgs[..., : 4 * c["h"], :] = 0.0
    fgs[:, 0] = 1.0
    m = MRCF_simple_v18("cuda", mid_channels=32, precision=precision).eval()
    m.load_state_dict(sd, strict=True)
    m.cuda()
    outs = [m(lrs[:, i:i + 1].cuda(), fvs[:, i:i + 1].cuda(), mks[:, i:i + 1].cuda(), fgs[:, i:i + 1].cuda()).cpu()
            for i in range(c["t"])]
    err = (torch.cat(outs, 1) - fix["out"]).abs().max().item()
    print(f"streaming: max-abs vs reference golden {err:.3e}")
    assert err <= TOL
    m.clear_states()
    again = m(lrs[:, :1].cuda(), fvs[:, :1].cuda(), mks[:, :1].cuda(), fgs[:, :1].cuda()).cpu()
    assert torch.equal(again, outs[0])


def test_streaming_graph_replay_and_state_validation(sd):
    """MRCF_simple_v18: per-frame CUDA-graph replay (from the third steady-state call on) is bit-identical to eager
    launches, the clip protocol clear_states() -> frames works across replays, and a frame of another size while state is
    held raises instead of reading a stale state (the reference fails with a shape error)."""
    from crfp_b200 import MRCF_simple_v18, _lib
    n, t, h, w = 1, 7, 16, 24
    lrs, fvs, mks, _ = make_clip(seed=19, n=n, t=t, h=h, w=w, fv_size=48)
    fgs = torch.ones(n, t, 1, 8 * h, 8 * w)
    fgs[:, 3:, :, : 4 * h] = 0.0
    lrs, fvs, mks, fgs = lrs.cuda(), fvs.cuda(), mks.cuda(), fgs.cuda()

    def stream(m, keep=True):
        m.clear_states()
        outs = []
        for i in range(t):
            o = m(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1])
            outs.append(o if keep else o.clone())
        return torch.cat(outs, 1)

    m = MRCF_simple_v18("cuda", mid_channels=32).eval()
    m.load_state_dict(sd, strict=True)
    m.cuda()
    m.use_graphs = False
    eager = stream(m)
    m.use_graphs = True
    first = stream(m)                      # frames 1.. : eager, capture on the second steady-state call, then replays
    assert any(sb["graphs"] for sb in m._sbuf.values())
    _lib.lib().crfp_launch_count_reset()
    again = stream(m)                      # frame 0 eager, frames 1.. replayed
    assert _lib.lib().crfp_launch_count() > 0
    torch.cuda.synchronize()
    assert torch.equal(first, eager) and torch.equal(again, eager)
    # two frames per call take their own graph and continue the same recurrence
    m.clear_states()
    o01 = m(lrs[:, 0:2], fvs[:, 0:2], mks[:, 0:2], fgs[:, 0:2])
    o2 = m(lrs[:, 2:3], fvs[:, 2:3], mks[:, 2:3], fgs[:, 2:3])     # t changes between calls: the state survives
    assert torch.equal(o01, eager[:, 0:2]) and torch.equal(o2, eager[:, 2:3])
    # another frame size while state is held
    lrs2, fvs2, mks2, _ = make_clip(seed=20, n=1, t=1, h=24, w=24, fv_size=48)
    with pytest.raises(ValueError, match="clear_states"):
        m(lrs2.cuda(), fvs2.cuda(), mks2.cuda(), torch.ones(1, 1, 1, 192, 192).cuda())
    m.clear_states()
    assert torch.isfinite(m(lrs2.cuda(), fvs2.cuda(), mks2.cuda(), torch.ones(1, 1, 1, 192, 192).cuda())).all()


def test_reds_native_shape_long_clip_properties(model, sd):
    """BASELINE config-2 shape (LR 90x160 -> 720x1280), t=12: finite, deterministic, and frame i only depends on
    frames <= i (causality of the recurrence): a 12-frame run and an 8-frame run agree on the first 8 frames."""
    lrs, fvs, mks, _ = make_clip(seed=3, n=1, t=12, h=90, w=160, fv_size=96)
    out = _run(model, lrs, fvs, mks)
    assert torch.isfinite(out).all()
    out8 = _run(model, lrs[:, :8], fvs[:, :8], mks[:, :8])
    assert torch.equal(out[:, :8], out8)
    ref = O.crfp_dsv_forward(sd, lrs[:, :3], fvs[:, :3], mks[:, :3])
    assert (out[:, :3] - ref).abs().max().item() <= TOL


@pytest.mark.parametrize("variant,precision", [("v15", "tc"), ("v15", "fp32"), ("v13", "tc"), ("v13", "fp32")])
def test_sibling_models_against_reference_golden(golden_dir, variant, precision):
    """CRFP (v15) and CRFP_simple (v13): same kernels, different wiring."""
    import crfp_b200
    fix = torch.load(os.path.join(golden_dir, f"{variant}_n1_t3_16x24.pt"))
    c = fix["case"]
    sdv = make_state_dict(seed=1, variant=variant)
    assert abs(float(sum(v.double().sum() for v in sdv.values())) - fix["weights_sum"]) < 1e-6
    cls = crfp_b200.CRFP if variant == "v15" else crfp_b200.CRFP_simple
    m = cls("cuda", mid_channels=32, precision=precision).eval()
    m.load_state_dict(sdv, strict=True)
    m.cuda()
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    out = m(lrs.cuda(), fvs.cuda(), mks.cuda()).cpu()
    err = (out - fix["out"]).abs().max().item()
    print(f"{variant} [{precision}]: max-abs vs reference golden {err:.3e}")
    assert err <= TOL
    # a second, larger clip against the live oracle
    lrs, fvs, mks, _ = make_clip(seed=23, n=2, t=3, h=24, w=40, fv_size=64)
    ref = O.crfp_forward(sdv, lrs, fvs, mks, variant=variant)
    err2 = (m(lrs.cuda(), fvs.cuda(), mks.cuda()).cpu() - ref).abs().max().item()
    print(f"{variant} [{precision}]: max-abs vs oracle, second clip {err2:.3e}")
    assert err2 <= TOL


def test_cuda_graph_replay_matches_eager(sd):
    """The whole-clip CUDA graph (captured on the second call with the same input buffers) must reproduce the eager
    forward bit for bit, count its launches, follow in-place input updates and stream frames to the host."""
    from crfp_b200 import CRFP_DSV, _lib
    m = CRFP_DSV("cuda", mid_channels=32).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    lrs, fvs, mks, _ = make_clip(seed=33, n=1, t=4, h=24, w=40, fv_size=64)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    m.use_graphs = False
    _lib.lib().crfp_launch_count_reset()
    eager = m(lrs, fvs, mks).clone()
    n_eager = _lib.lib().crfp_launch_count()
    m.use_graphs = True
    m.graph_frames = 3                    # 4 frames -> two chained graphs
    o1 = m(lrs, fvs, mks)                 # first sighting: eager
    assert len(m._graphs) == 0
    o2 = m(lrs, fvs, mks)                 # second sighting: capture + replay
    assert len(m._graphs) == 1 and m.use_graphs and len(next(iter(m._graphs.values()))["graphs"]) == 2
    _lib.lib().crfp_launch_count_reset()
    o3 = m(lrs, fvs, mks)                 # replay
    assert _lib.lib().crfp_launch_count() == n_eager > 0
    torch.cuda.synchronize()
    assert torch.equal(o1, eager) and torch.equal(o3, eager)
    # default contract: every call returns its own tensor (results kept by the caller survive later replays) ...
    assert o3 is not o2 and o3.data_ptr() != o2.data_ptr()
    # ... aliasing is opt-in: the graph-owned tensor itself, or the caller's own `out` buffer
    m.alias_output = True
    a1, a2 = m(lrs, fvs, mks), m(lrs, fvs, mks)
    assert a1 is a2 and torch.equal(a1, eager)
    m.alias_output = False
    mine = torch.empty_like(eager)
    for _ in range(3):                    # eager, capture, replay — all into the caller's buffer
        mine.zero_()
        r = m(lrs, fvs, mks, out=mine)
        assert r.data_ptr() == mine.data_ptr() and torch.equal(mine, eager)
    # inputs that need a conversion (temporaries at allocator-chosen addresses) never enter a graph
    ngraphs = len(m._graphs)
    for _ in range(3):
        assert torch.equal(m(lrs.double(), fvs, mks), eager)
    assert len(m._graphs) == ngraphs
    # the graph reads the buffers, not a snapshot of their contents
    lrs2, fvs2, mks2, _ = make_clip(seed=34, n=1, t=4, h=24, w=40, fv_size=64)
    lrs.copy_(lrs2.cuda()); fvs.copy_(fvs2.cuda()); mks.copy_(mks2.cuda())
    o4 = m(lrs, fvs, mks).clone()
    m.use_graphs = False
    assert torch.equal(o4, m(lrs, fvs, mks))
    # streaming device->host copies inside the graph
    m.use_graphs = True
    host = torch.empty(o4.shape, dtype=torch.float32).pin_memory()
    for _ in range(3):
        host.zero_()
        m(lrs, fvs, mks, out_host=host)
        torch.cuda.synchronize()
        assert torch.equal(host, o4.cpu())
    # new weights invalidate the graphs
    m.load_state_dict(make_state_dict(seed=2), strict=True)
    o5 = m(lrs, fvs, mks)
    assert not torch.equal(o5, o4)


def test_forward_patch_device_paste(sd):
    """forward_patch (device-side fovea paste into persistent buffers, previous rectangles cleared, graph replay from
    the second call on) == forward on host-assembled fvs / mks, for changing gaze positions."""
    from crfp_b200 import CRFP_DSV
    from crfp_b200.synthetic import fovea_rect
    m = CRFP_DSV("cuda", mid_channels=32).eval()
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    n, t, h, w, fv = 2, 3, 16, 24, 48
    H, W = 8 * h, 8 * w
    g = torch.Generator().manual_seed(5)
    lrs = torch.rand(n, t, 3, h, w, generator=g).cuda()
    patch = torch.rand(n, t, 3, fv, fv, generator=g).cuda()
    for trial in range(4):
        coords = torch.stack([torch.randint(0, H - fv + 1, (n, t), generator=g), torch.randint(0, W - fv + 1, (n, t), generator=g)], -1)
        if trial == 3:
            patch.copy_(torch.rand(n, t, 3, fv, fv, generator=g))
        fvs = torch.zeros(n, t, 3, H, W)
        for b in range(n):
            for i in range(t):
                y, x = int(coords[b, i, 0]), int(coords[b, i, 1])
                fvs[b, i, :, y:y + fv, x:x + fv] = patch[b, i].cpu()
        mks = fovea_rect(coords, fv, H, W)
        m.use_graphs = False
        ref = m(lrs, fvs.cuda(), mks.cuda()).clone()
        m.use_graphs = True
        out = m.forward_patch(lrs, patch, coords)
        torch.cuda.synchronize()
        assert torch.equal(out, ref), f"trial {trial}"
    assert len(m._graphs) >= 1
    with pytest.raises(ValueError):
        m.forward_patch(lrs, patch, coords + 10_000)


def test_reds_literal_full_size_properties(sd):
    """BASELINE configs[1] at full size (LR 180x320 -> 1440x2560), t=4, through size-independent properties: finite,
    tc and fp32 precisions agree within the fp32 bar, the fovea tile skip is bit-identical to the full computation,
    graph replay is bit-identical to eager execution, and the recurrence is causal."""
    from crfp_b200 import CRFP_DSV
    lrs, fvs, mks, _ = make_clip(seed=11, n=1, t=4, h=180, w=320, fv_size=96)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    outs = {}
    for prec in ("tc", "fp32"):
        m = CRFP_DSV("cuda", mid_channels=32, precision=prec).eval()
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        m.use_graphs = False
        outs[prec] = m(lrs, fvs, mks).clone()
        assert torch.isfinite(outs[prec]).all()
        if prec == "tc":
            m.skip_outside_fovea = False
            assert torch.equal(m(lrs, fvs, mks), outs[prec])            # exact tile skip
            m.skip_outside_fovea = True
            assert torch.equal(m(lrs[:, :2], fvs[:, :2], mks[:, :2]), outs[prec][:, :2])   # causality
            m.use_graphs = True
            m(lrs, fvs, mks)
            assert torch.equal(m(lrs, fvs, mks), outs[prec]) and len(m._graphs) == 1       # graph replay
        del m
    err = (outs["tc"] - outs["fp32"]).abs().max().item()
    # the 0.05 dB criterion of BASELINE.json, PSNR as the reference computes it, against a fixed pseudo ground truth
    from crfp_b200.metrics import psnr
    gt = torch.rand(outs["tc"].shape[1:], generator=torch.Generator().manual_seed(3)).cuda()
    dpsnr = abs(float(psnr(outs["tc"][0].clamp(0, 1), gt)) - float(psnr(outs["fp32"][0].clamp(0, 1), gt)))
    print(f"R-lit full size: tc vs fp32 max-abs {err:.3e}, delta PSNR {dpsnr:.2e} dB")
    assert err <= TOL and dpsnr <= 0.05
