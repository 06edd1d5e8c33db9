"""GPU tests (-m gpu) of the legacy SPyNet flow pyramid on the B200 against the golden outputs of the real reference
class (fp32 bar: max-abs <= 1e-3 on flows of ~1 px).  CPU twin: tests/test_spynet.py.

First ran green on the driver's B200 at the end of round 1 (GPUTEST_r01.json: XPASS); the xfail marker is gone, a
regression turns the suite red."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_spynet_ops_on_gpu():
    from crfp_b200.spynet import CUDA as K
    g = torch.Generator().manual_seed(1)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()      # noqa: E731
    for cin, cout, hw in ((8, 32, (33, 47)), (64, 32, (16, 24)), (16, 2, (20, 31))):
        x = torch.randn(2, cin, *hw, generator=g)
        wt = torch.randn(cout, cin, 7, 7, generator=g) * 0.05
        b = torch.randn(cout, generator=g) * 0.1
        res = torch.randn(2, cout, *hw, generator=g)
        want = F.conv2d(F.relu(x), wt, b, padding=3) + res
        wp = wt.permute(2, 3, 1, 0).reshape(49, cin, cout).contiguous().cuda()
        got = K.conv_kxk(nhwc(x), wp, b.cuda(), 7, True, residual=nhwc(res)).cpu()
        assert (got.permute(0, 3, 1, 2) - want).abs().max().item() < 2e-4
    f = torch.randn(2, 2, 9, 13, generator=g) * 4
    want = F.interpolate(f, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
    assert (K.resize_ac(nhwc(f), 18, 26, 2.0).cpu().permute(0, 3, 1, 2) - want).abs().max().item() < 1e-5


@pytest.mark.parametrize("name", ["spynet_n2_64x96", "spynet_n1_40x72"])
def test_spynet_matches_reference_golden(name, golden_dir):
    from crfp_b200.spynet import SPyNet, make_spynet_pair, make_spynet_state_dict
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    sd = make_spynet_state_dict(seed=c["wseed"])
    ref, supp = make_spynet_pair(seed=c["seed"], n=c["n"], h=c["h"], w=c["w"])
    net = SPyNet(None, "cuda")
    net.load_state_dict({**sd, "mean": net.mean, "std": net.std}, strict=True)
    net.cuda()
    out = net(ref.cuda(), supp.cuda()).cpu()
    assert out.shape == fix["out"].shape and (out - fix["out"]).abs().max().item() < 1e-3
