"""CPU tests (-m "not gpu") of the legacy SPyNet flow pyramid: the oracle restatement against golden
outputs of the REAL reference class (oracle/make_golden_spynet.py), the new kernels of spynet.cu through the host emulation,
and the module's wiring end to end.  GPU twin: tests/test_gpu_zzz_spynet.py."""
import os

import pytest
import torch
import torch.nn.functional as F

import hostemu
from crfp_b200.spynet import SPyNet, make_spynet_pair, make_spynet_state_dict
from oracle import crfp_oracle as O

CASES = ["spynet_n2_64x96", "spynet_n1_40x72"]


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _case(golden_dir, name):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    sd = make_spynet_state_dict(seed=c["wseed"])
    ref, supp = make_spynet_pair(seed=c["seed"], n=c["n"], h=c["h"], w=c["w"])
    assert abs(float(sum(v.double().sum() for v in sd.values())) - fix["weights_sum"]) < 1e-6
    assert abs(float(ref.double().sum()) - fix["ref_sum"]) < 1e-6
    return fix, sd, ref, supp


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name, golden_dir):
    fix, sd, ref, supp = _case(golden_dir, name)
    out = O.spynet(sd, ref, supp)
    assert out.shape == fix["out"].shape and (out - fix["out"]).abs().max().item() <= 1e-6
    assert fix["out"].abs().max().item() > 0.3          # the fixture exercises real (non-zero) flows


def test_spynet_kernels_through_the_emulation():
    K = hostemu.spy_kernels()
    g = _g(1)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()      # noqa: E731
    for cin, cout, hw in ((8, 32, (9, 12)), (32, 16, (7, 7)), (16, 2, (5, 11))):
        x = torch.randn(2, cin, *hw, generator=g)
        wt = torch.randn(cout, cin, 7, 7, generator=g) * 0.05
        b = torch.randn(cout, generator=g) * 0.1
        res = torch.randn(2, cout, *hw, generator=g)
        want = F.conv2d(F.relu(x), wt, b, padding=3) + res
        wp = wt.permute(2, 3, 1, 0).reshape(49, cin, cout).contiguous()
        got = K.conv_kxk(nhwc(x), wp, b, 7, True, residual=nhwc(res))
        assert (got.permute(0, 3, 1, 2) - want).abs().max().item() < 1e-4
        got = K.conv_kxk(nhwc(x), wp, b, 7, False)
        assert (got.permute(0, 3, 1, 2) - F.conv2d(x, wt, b, padding=3)).abs().max().item() < 1e-4
    f = torch.randn(2, 2, 3, 5, generator=g) * 4
    want = F.interpolate(f, scale_factor=2, mode="bilinear", align_corners=True) * 2.0
    assert (K.resize_ac(nhwc(f), 6, 10, 2.0).permute(0, 3, 1, 2) - want).abs().max().item() < 1e-5
    one = torch.randn(1, 2, 1, 1, generator=g)
    want = F.interpolate(one, scale_factor=2, mode="bilinear", align_corners=True)
    assert (K.resize_ac(nhwc(one), 2, 2, 1.0).permute(0, 3, 1, 2) - want).abs().max().item() < 1e-6
    img = torch.rand(2, 3, 4, 6, generator=g)
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    got = K.affine(nhwc(img), mean, std, torch.ones(3), 4)
    assert torch.equal(got[..., :3], nhwc((img - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1))) and got[..., 3].abs().sum() == 0
    h = hostemu.lib()
    assert h.crfp_conv_kxk_fwd(1, 4, 4, 8, 8, 6, 0, None, None, None, None, None, None) == -1     # even kernel size
    assert h.crfp_conv_kxk_fwd(1, 4, 4, 8, 8, 7, 0, None, None, None, None, None, None) == -5


@pytest.mark.parametrize("name", CASES)
def test_module_matches_reference_golden_through_the_emulation(name, golden_dir):
    fix, sd, ref, supp = _case(golden_dir, name)
    net = SPyNet(None, "cpu", kernels=hostemu.spy_kernels())
    assert set(net.state_dict().keys()) == set(sd.keys()) | {"mean", "std"}        # the reference's state_dict keys
    net.load_state_dict({**sd, "mean": net.mean, "std": net.std}, strict=True)
    out = net(ref, supp)
    assert out.shape == fix["out"].shape and (out - fix["out"]).abs().max().item() < 1e-4
    with pytest.raises(ValueError):
        net(ref, supp[:, :2])
    with pytest.raises(TypeError):
        SPyNet(3, "cpu")
