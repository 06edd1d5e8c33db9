"""CPU tests (-m "not gpu"): host logic — weight packing, module shell / state_dict contract, C-ABI exports."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

from crfp_b200 import packing
from crfp_b200.spec import crfp_dsv_param_shapes
from crfp_b200.synthetic import fovea_rect, make_clip, make_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emulate_packed_conv(srcs, modes, wp, bp, cout):
    """Run the PACKED weights through a plain conv on the packed input layout (what the kernel computes)."""
    cols = []
    for s, m in zip(srcs, modes):
        if m == packing.SRC_UNSHUFFLE4:
            n, c, H, W = s.shape                     # packed order (dy*4+dx)*c + ch
            u = s.view(n, c, H // 4, 4, W // 4, 4).permute(0, 3, 5, 1, 2, 4).reshape(n, 16 * c, H // 4, W // 4)
            cols.append(u)
            pad = (-16 * c) % 4
        else:
            cols.append(s)
            pad = (-s.shape[1]) % 4
        if pad:
            cols.append(torch.zeros(s.shape[0], pad, *cols[-1].shape[2:]))
    x = torch.cat(cols, 1)
    cp = wp.shape[1]
    x = F.pad(x, (0, 0, 0, 0, 0, cp - x.shape[1]))
    w = wp.view(3, 3, cp, -1).permute(3, 2, 0, 1)     # (co_packed, ci_packed, ky, kx)
    return F.conv2d(x, w, bp, padding=1)[:, :cout]


@pytest.mark.parametrize("c_list,modes,cout", [([32, 32, 2], [0, 0, 0], 32), ([3, 3], [0, 0], 32), ([24], [0], 64),
                                               ([4, 4, 2], [0, 0, 0], 4), ([6], [0], 4), ([32], [0], 216)])
def test_pack_conv_equals_plain_conv(c_list, modes, cout):
    g = torch.Generator().manual_seed(0)
    cin = sum(c_list)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    b = torch.randn(cout, generator=g)
    srcs = [torch.randn(1, c, 6, 7, generator=g) for c in c_list]
    wp, bp = packing.pack_conv(w, b, c_list, modes)
    assert wp.shape == (9, packing.cin_packed(c_list), packing.cout_packed(cout))
    ref = F.conv2d(torch.cat(srcs, 1), w, b, padding=1)
    got = _emulate_packed_conv(srcs, modes, wp, bp, cout)
    assert (ref - got).abs().max().item() < 1e-4


def test_pack_conv_unshuffle_source_and_input_slice():
    g = torch.Generator().manual_seed(1)
    s_hr = torch.randn(1, 4, 16, 24, generator=g)
    w = torch.randn(32, 64, 3, 3, generator=g)
    b = torch.randn(32, generator=g)
    wp, bp = packing.pack_conv(w, b, [64], [packing.SRC_UNSHUFFLE4])
    ref = F.conv2d(F.pixel_unshuffle(s_hr, 4), w, b, padding=1)
    got = _emulate_packed_conv([s_hr], [packing.SRC_UNSHUFFLE4], wp, bp, 32)
    assert (ref - got).abs().max().item() < 1e-4
    # first-frame variant: only the first 24 input channels of a 64-input conv
    x = torch.randn(1, 24, 5, 5, generator=g)
    wp2, bp2 = packing.pack_conv(w, b, [24], [0], ci_lo=0)
    ref2 = F.conv2d(torch.cat([x, torch.zeros(1, 40, 5, 5)], 1), w, b, padding=1)
    assert (ref2 - _emulate_packed_conv([x], [0], wp2, bp2, 32)).abs().max().item() < 1e-4


def test_pack_dcn_layout():
    g = torch.Generator().manual_seed(2)
    w = torch.randn(32, 32, 3, 3, generator=g)
    b = torch.randn(32, generator=g)
    wp, bp = packing.pack_dcn(w, b, 8)
    assert wp.shape == (288, 32)
    for (gi, t, c, o) in [(0, 0, 0, 0), (3, 5, 2, 17), (7, 8, 3, 31)]:
        assert wp[(gi * 9 + t) * 4 + c, o] == w[o, gi * 4 + c, t // 3, t % 3]
    wp4, _ = packing.pack_dcn(torch.randn(4, 4, 3, 3, generator=g), torch.zeros(4), 1)
    assert wp4.shape == (36, 4)


def test_state_dict_contract():
    shapes = crfp_dsv_param_shapes(32)
    assert len(shapes) == 118
    assert sum(int(torch.tensor(s).prod()) for s in shapes.values()) == 2284352
    from crfp_b200 import CRFP_DSV, MRCF_simple_v18
    sd = make_state_dict(1)
    for cls in (CRFP_DSV, MRCF_simple_v18):
        m = cls("cuda", mid_channels=32)
        own = m.state_dict()
        assert list(own.keys()) == list(shapes.keys())
        assert all(tuple(v.shape) == tuple(shapes[k]) and v.dtype == torch.float32 for k, v in own.items())
        m.load_state_dict(sd, strict=True)
    # reference init policy: DCN heads zero, dcn.weight identity at the centre tap (CRFP.py:354-370)
    m = CRFP_DSV("cuda", mid_channels=32)
    assert m.dcn_0.dcn_offset.weight.abs().sum() == 0 and m.dcn_3.dcn_mask.bias.abs().sum() == 0
    assert torch.equal(m.dcn_1.dcn.weight[:, :, 1, 1], torch.eye(32))


def test_sibling_state_dict_contracts():
    from crfp_b200 import CRFP, CRFP_simple
    from crfp_b200.spec import crfp_param_shapes
    for cls, variant, nparam in ((CRFP, "v15", 2326000), (CRFP_simple, "v13", 2298208)):   # SURVEY.md section 2 [probe]
        shapes = crfp_param_shapes(variant, 32)
        m = cls("cuda", mid_channels=32)
        assert list(m.state_dict().keys()) == list(shapes.keys()) and len(shapes) == 118
        assert sum(p.numel() for p in m.parameters()) == nparam
        m.load_state_dict(make_state_dict(1, variant=variant), strict=True)


def test_module_errors_loudly_without_cuda_inputs():
    from crfp_b200 import CRFP_DSV
    from crfp_b200._lib import CrfpError
    m = CRFP_DSV("cuda", mid_channels=32).eval()
    lrs, fvs, mks, _ = make_clip(0, 1, 2, 8, 8, 16)
    with pytest.raises(CrfpError):
        m(lrs, fvs, mks)           # CPU tensors: no fallback
    with pytest.raises(CrfpError):
        CRFP_DSV("cuda", mid_channels=16)
    with pytest.raises(TypeError):
        m.init_weights(pretrained=123)


def test_fovea_rect_matches_dataset_contract():
    sp = torch.tensor([[[3, 5], [0, 0]]])
    mk = fovea_rect(sp, 4, 16, 16)
    assert mk.dtype == torch.bool and mk.shape == (1, 2, 1, 16, 16)
    assert mk[0, 0, 0, 3:7, 5:9].all() and mk[0, 0].sum() == 16 and mk[0, 1, 0, :4, :4].all()


def test_c_abi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "crfp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(crfp_[a-z0-9_]+)\s*\(", hdr))
    from crfp_b200 import _lib
    assert declared == set(_lib.SYMBOLS.keys()), declared ^ set(_lib.SYMBOLS.keys())
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} not exported"
    assert built_lib.crfp_abi_version() == 1
    assert built_lib.crfp_status_string(-3).decode() == "workspace too small"
    assert built_lib.crfp_conv_cout_packed(216) == 224 and built_lib.crfp_conv_cout_packed(3) == 4
    arr = (ctypes.c_int32 * 3)(32, 32, 2)
    assert built_lib.crfp_conv_cin_packed(3, arr) == 72 == packing.cin_packed([32, 32, 2])


def test_layer_table_matches_state_dict(built_lib):
    from crfp_b200 import _lib
    shapes = crfp_dsv_param_shapes(32)
    for variant in ("v15", "v13"):          # sibling tables pack too
        from crfp_b200.spec import crfp_param_shapes
        sdv = make_state_dict(1, variant=variant)
        for info in _lib.layer_table(variant):
            w, b = packing.pack_layer(info, sdv)
            if info["tc"]:
                packing.pack_layer_tc3(info, sdv)
    table = _lib.layer_table()
    used = set()
    sd = make_state_dict(1)
    for info in table:
        assert info["key"] + ".weight" in shapes
        used.add(info["key"])
        if info["key2"]:
            used.add(info["key2"])
        w, b = packing.pack_layer(info, sd)
        assert w.is_contiguous() and b.is_contiguous() and w.dtype == torch.float32
        if info["kind"] != 1:
            cin = shapes[info["key"] + ".weight"][1]
            assert info["ci_lo"] + sum(info["c"]) <= cin
            assert w.shape == (9, packing.cin_packed(info["c"]), packing.cout_packed(info["cout"]))
    assert used == {k[:-7] for k in shapes if k.endswith(".weight")}   # every parameter is consumed
    # workspace queries answer without a GPU
    shp = _lib.DsvShape(n=1, t=5, h=90, w=160, mid_channels=32)
    assert built_lib.crfp_dsv_frame_workspace(ctypes.byref(shp)) > 0
    assert built_lib.crfp_dsv_prepare_workspace(ctypes.byref(shp)) > 0
    bad = _lib.DsvShape(n=1, t=5, h=4, w=160, mid_channels=32)
    assert built_lib.crfp_dsv_frame_workspace(ctypes.byref(bad)) == 0


def test_psnr_matches_reference_formula():
    """crfp_b200.metrics.psnr == the reference's psnr_cuda (utils.py:165-184) on masked / unmasked / identical inputs."""
    import math
    import torch
    from crfp_b200.metrics import psnr
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(2, 3, 8, 10, generator=g), torch.rand(2, 3, 8, 10, generator=g)
    mse = ((a - b) ** 2).mean().item()
    assert abs(float(psnr(a, b)) - (-20 * math.log10(math.sqrt(mse)))) < 1e-5
    m = (torch.rand(2, 1, 8, 10, generator=g) > 0.5).float()
    mse_m = (((a - b) ** 2) * m).sum().item() / (m.sum().item() * 3)
    assert abs(float(psnr(a, b, m)) - (-20 * math.log10(math.sqrt(mse_m)))) < 1e-5
    per = psnr(a, b, batch_avg=True)
    assert per.shape == (2,) and abs(float(per[0]) - (-20 * math.log10(math.sqrt(((a[0] - b[0]) ** 2).mean().item())))) < 1e-5
    assert math.isfinite(float(psnr(a, a))) and float(psnr(a, a)) > 60


def test_training_and_spynet_refuse_cpu_tensors():
    """No CPU fallback anywhere: the training forward, the loss and SPyNet raise on CPU tensors instead of computing."""
    import pytest
    import torch
    from crfp_b200 import CRFP_DSV, SPyNet, _lib
    from crfp_b200 import autograd as A
    from crfp_b200.synthetic import make_clip
    from crfp_b200.training import forward_train
    lrs, fvs, mks, _ = make_clip(seed=0, n=1, t=2, h=8, w=8, fv_size=16)
    model = CRFP_DSV("cuda", mid_channels=32).train()
    with pytest.raises(_lib.CrfpError):
        forward_train(model, lrs, fvs, mks)                       # default kernel set = the CUDA library
    with pytest.raises(_lib.CrfpError):
        model(lrs, fvs, mks)                                      # module entry, training mode
    with pytest.raises(_lib.CrfpError):
        A.charbonnier_loss(A.CUDA, torch.rand(1, 3, 4, 4), torch.rand(1, 3, 4, 4))
    with pytest.raises(_lib.CrfpError):
        A.conv3x3(A.CUDA, torch.rand(4, 4, 3, 3), torch.rand(4), [torch.rand(1, 5, 5, 4)])
    with pytest.raises(_lib.CrfpError):
        SPyNet(None, "cpu")(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32))
