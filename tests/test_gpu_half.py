"""GPU tests (-m gpu) of the reduced-precision tier `precision="half"` (CRFP_PREC_HALF; north_star's "bf16" tier):
max-abs <= 5e-3 and PSNR delta <= 0.05 dB against the reference goldens / the oracle, over the same long-recurrence and
full-size cases as the fp32-grade precisions (tests/test_gpu_long.py)."""
import os

import pytest
import torch

from crfp_b200.synthetic import make_clip, make_state_dict
from oracle import crfp_oracle as O
from fixture_compare import compare_with_fixture

pytestmark = pytest.mark.gpu
TOL = 5e-3


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(seed=1)


def _model(sd, precision="half"):
    from crfp_b200 import CRFP_DSV
    m = CRFP_DSV("cuda", mid_channels=32, precision=precision).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda()


@pytest.mark.parametrize("name", ["long_t100_32x48", "rnat_t12_90x160", "rlit_t3_180x320"])
def test_half_precision_against_reference_golden(golden_dir, sd, name):
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=1, t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    out = _model(sd)(lrs.cuda(), fvs.cuda(), mks.cuda()).cpu()
    errs, mean_errs = compare_with_fixture(out, fix)
    step = max(1, len(errs) // 10)
    print(f"\n{name} [half]: max-abs vs REFERENCE golden per frame (every {step}th): {' '.join('%.1e' % e for e in errs[::step])}; "
          f"worst {max(errs):.3e}, checksum mean-error {max(mean_errs):.2e}")
    assert max(errs) <= TOL


@pytest.mark.parametrize("n,t,h,w,fv", [(1, 5, 32, 48, 96), (2, 3, 24, 40, 64)])
def test_half_precision_against_oracle_and_psnr(sd, n, t, h, w, fv):
    from crfp_b200.metrics import psnr
    lrs, fvs, mks, _ = make_clip(seed=21, n=n, t=t, h=h, w=w, fv_size=fv)
    ref = O.crfp_dsv_forward(sd, lrs, fvs, mks)
    out = _model(sd)(lrs.cuda(), fvs.cuda(), mks.cuda()).cpu()
    err = (out - ref).abs().max().item()
    gt = torch.rand(ref.shape[1:], generator=torch.Generator().manual_seed(3))
    dps = max(abs(float(psnr(out[b].clamp(0, 1), gt[:])) - float(psnr(ref[b].clamp(0, 1), gt[:]))) for b in range(n))
    print(f"[half] n={n} t={t} {h}x{w}: max-abs vs oracle {err:.3e}, delta PSNR {dps:.2e} dB")
    assert err <= TOL and dps <= 0.05


def test_half_precision_is_a_different_arithmetic(sd):
    """Sanity: the tier really changes the arithmetic (it is not silently the fp32-grade path) and stays deterministic."""
    lrs, fvs, mks, _ = make_clip(seed=7, n=1, t=3, h=16, w=24, fv_size=48)
    a = _model(sd)(lrs.cuda(), fvs.cuda(), mks.cuda())
    b = _model(sd, "tc")(lrs.cuda(), fvs.cuda(), mks.cuda())
    assert not torch.equal(a, b) and (a - b).abs().max().item() <= TOL
    assert torch.equal(a, _model(sd)(lrs.cuda(), fvs.cuda(), mks.cuda()))
