"""Generate tests/golden/*.pt from the REAL reference (run in the build container only).

TEST INFRASTRUCTURE.  Imports the unmodified reference classes from /root/reference with two
shims: a `dcn_v2` module that forwards to torchvision.ops.deform_conv2d, and a
stub `fnet.pth` (the reference loads one unconditionally, CRFP.py:1407).  It then
  1. checks that the oracle restatement (oracle/crfp_oracle.py) reproduces the reference's output
     on every case (bit-identical or <= 1e-6), and
  2. saves the reference's outputs (+ a few intermediates of the oracle) as small fixtures that
     travel to the GPU box, where /root/reference does not exist.

Usage:  python oracle/make_golden.py            (writes tests/golden/)
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from crfp_b200.synthetic import make_clip, make_state_dict  # noqa: E402
from oracle import crfp_oracle as O  # noqa: E402


def _install_dcn_shim():
    import torchvision.ops as tvo

    class DCNv2(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                     deformable_groups=1):
            super().__init__()
            self.stride, self.padding, self.dilation = stride, padding, dilation
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
            self.bias = nn.Parameter(torch.zeros(out_channels))
            n = in_channels * kernel_size * kernel_size
            self.weight.data.uniform_(-1.0 / n ** 0.5, 1.0 / n ** 0.5)

        def forward(self, input, offset, mask):
            return tvo.deform_conv2d(input, offset, self.weight, self.bias, stride=self.stride,
                                     padding=self.padding, dilation=self.dilation, mask=mask)

    mod = types.ModuleType("dcn_v2")
    mod.DCNv2 = DCNv2
    sys.modules["dcn_v2"] = mod


def load_reference():
    _install_dcn_shim()
    sys.path.insert(0, REF)
    from model import CRFP, CRFP_test  # type: ignore
    return CRFP, CRFP_test


def build_ref_model(cls, sd, **kw):
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "fnet.pth")
        torch.save({k[len("spynet."):]: v for k, v in sd.items() if k.startswith("spynet.")}, path)
        m = cls("cpu", mid_channels=32, spynet_pretrained=path, **kw)
    missing = set(m.state_dict().keys()) ^ set(sd.keys())
    assert not missing, f"state_dict key mismatch: {sorted(missing)[:8]}"
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    m.load_state_dict(sd, strict=True)
    return m.eval()


CASES = [
    # name, n, t, h, w, fv
    ("dsv_n1_t3_16x24", 1, 3, 16, 24, 48),
    ("dsv_n2_t2_18x20", 2, 2, 18, 20, 40),   # h,w not multiples of 8: FNet resamples 16->18, 16->20
    ("dsv_n1_t1_8x8", 1, 1, 8, 8, 32),       # single frame: no flow / warp / DCN at all
]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    CRFP, CRFP_test = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    sd = make_state_dict(seed=1)
    ref = build_ref_model(CRFP.CRFP_DSV, sd)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    for ci, (name, n, t, h, w, fv) in enumerate(CASES):
        lrs, fvs, mks, fv_sp = make_clip(seed=2 + ci, n=n, t=t, h=h, w=w, fv_size=fv)
        with torch.no_grad():
            y_ref = ref(lrs.clone(), fvs.clone(), mks.clone())
        taps = []
        y_or = O.crfp_dsv_forward(sd, lrs, fvs, mks, taps=taps)
        d = (y_ref - y_or).abs().max().item()
        y_nv = O.crfp_dsv_forward(sd, lrs, fvs, mks, naive_dcn=True)
        d2 = (y_ref - y_nv).abs().max().item()
        print(f"{name}: ref vs oracle max-abs {d:.3e}; vs oracle(naive DCN) {d2:.3e}; out range "
              f"[{y_ref.min():.3f},{y_ref.max():.3f}]")
        assert d <= 1e-6, "oracle restatement deviates from the reference"
        assert d2 <= 2e-4, "naive DCN restatement deviates from the reference"
        fix = {"case": dict(n=n, t=t, h=h, w=w, fv=fv, seed=2 + ci, weight_seed=1),
               "weights_sum": wsum, "lrs_sum": float(lrs.double().sum()), "fvs_sum": float(fvs.double().sum()),
               "fv_sp": fv_sp, "out": y_ref.contiguous()}
        if t > 1:
            tp = taps[-1]
            fix["taps_last"] = {k: tp[k].contiguous() for k in ("flow", "dcn0_out", "S")}
            print("   flow abs mean %.3f max %.3f | dcn0 offset abs mean %.3f | mask mean %.3f" % (
                tp["flow"].abs().mean(), tp["flow"].abs().max(), tp["dcn0_offset"].abs().mean(),
                tp["dcn0_mask"].mean()))
        torch.save(fix, os.path.join(out_dir, name + ".pt"))

    # sibling models: CRFP (v15) and CRFP_simple (v13)
    for variant, cls in (("v15", CRFP.CRFP), ("v13", CRFP.CRFP_simple)):
        sdv = make_state_dict(seed=1, variant=variant)
        refv = build_ref_model(cls, sdv)
        name, n, t, h, w, fv = f"{variant}_n1_t3_16x24", 1, 3, 16, 24, 48
        lrs, fvs, mks, fv_sp = make_clip(seed=21, n=n, t=t, h=h, w=w, fv_size=fv)
        with torch.no_grad():
            y_ref = refv(lrs.clone(), fvs.clone(), mks.clone())
        y_or = O.crfp_forward(sdv, lrs, fvs, mks, variant=variant)
        d = (y_ref - y_or).abs().max().item()
        print(f"{name}: ref vs oracle max-abs {d:.3e}; out range [{y_ref.min():.3f},{y_ref.max():.3f}]")
        assert d <= 1e-6, "oracle restatement deviates from the reference"
        torch.save({"case": dict(n=n, t=t, h=h, w=w, fv=fv, seed=21, weight_seed=1, variant=variant),
                    "weights_sum": float(sum(v.double().sum() for v in sdv.values())),
                    "lrs_sum": float(lrs.double().sum()), "out": y_ref.contiguous()}, os.path.join(out_dir, name + ".pt"))

    # streaming model: same state_dict, frame by frame, with a regional fg mask
    name, n, t, h, w, fv = "stream_n1_t3_16x24", 1, 3, 16, 24, 48
    lrs, fvs, mks, fv_sp = make_clip(seed=11, n=n, t=t, h=h, w=w, fv_size=fv)
    fgs = torch.ones(n, t, 1, 8 * h, 8 * w)
    fgs[..., : 4 * h, :] = 0.0
    fgs[:, 0] = 1.0
    sref = build_ref_model(CRFP_test.MRCF_simple_v18, sd)
    sor = O.StreamingOracle(sd)
    outs_ref, outs_or = [], []
    with torch.no_grad():
        for i in range(t):
            outs_ref.append(sref(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1]))
            outs_or.append(sor(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1]))
    y_ref, y_or = torch.cat(outs_ref, 1), torch.cat(outs_or, 1)
    d = (y_ref - y_or).abs().max().item()
    print(f"{name}: streaming ref vs oracle max-abs {d:.3e}")
    assert d <= 1e-6
    torch.save({"case": dict(n=n, t=t, h=h, w=w, fv=fv, seed=11, weight_seed=1), "weights_sum": wsum,
                "lrs_sum": float(lrs.double().sum()), "out": y_ref.contiguous()},
               os.path.join(out_dir, name + ".pt"))

    # operator-level known answers (tiny): flow_warp via the reference function, DCNv2 via the shim
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 5, 9, 11, generator=g)
    fl = torch.randn(2, 2, 9, 11, generator=g) * 3.0
    yw = CRFP.flow_warp(x, fl.permute(0, 2, 3, 1))
    assert (yw - O.flow_warp(x, fl)).abs().max().item() == 0.0
    xd = torch.randn(1, 8, 7, 9, generator=g)
    off = torch.randn(1, 2 * 2 * 9, 7, 9, generator=g) * 4.0
    msk = torch.rand(1, 2 * 9, 7, 9, generator=g)
    wd = torch.randn(6, 8, 3, 3, generator=g) * 0.2
    bd = torch.randn(6, generator=g)
    import torchvision.ops as tvo
    yd = tvo.deform_conv2d(xd, off, wd, bd, stride=1, padding=1, dilation=1, mask=msk)
    assert (yd - O.dcn_v2_naive(xd, off, msk, wd, bd, 2)).abs().max().item() < 1e-4
    torch.save({"warp": dict(x=x, flow=fl, out=yw), "dcn": dict(x=xd, offset=off, mask=msk, weight=wd, bias=bd,
                                                                 dg=2, out=yd)},
               os.path.join(out_dir, "ops_kat.pt"))
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
