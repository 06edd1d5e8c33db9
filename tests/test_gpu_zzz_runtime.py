"""GPU test (-m gpu) of the `CRFP_runtime.MRCF_simple_v18` drop-in against the golden outputs of the real reference class
(fp32 bar: max-abs <= 1e-3).  CPU twin: tests/test_runtime_shell.py.  First ran green on the driver's B200 at the end of
round 1 (GPUTEST_r01.json: XPASS); the xfail marker is gone, a regression turns the suite red."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["runtime_full_n1_t3_16x24", "runtime_region_n2_t3_16x24"])
def test_runtime_shell_matches_reference_golden(name, golden_dir):
    from crfp_b200.runtime import MRCF_simple_v18, make_runtime_state_dict
    from crfp_b200.synthetic import make_clip
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    sd = make_runtime_state_dict(seed=c["wseed"])
    lrs, _, _, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=16)
    fvs = torch.rand(c["n"], c["t"], 3, c["fh"], c["fw"], generator=torch.Generator().manual_seed(c["fv_seed"]))
    model = MRCF_simple_v18("cuda", mid_channels=32)
    model.load_state_dict(sd, strict=True)
    model.cuda()
    out = model(lrs.cuda(), fvs.cuda(), warp_size=tuple(c["warp"])).cpu()
    assert out.shape == fix["out"].shape and (out - fix["out"]).abs().max().item() < 1e-3
