"""GPU test (-m gpu): spatial tiling with per-frame halo refresh (crfp_b200/tiling.py, BASELINE.json configs[3]) must
reproduce the untiled forward on every pixel; all tiles run back to back on one GPU ("virtual ranks")."""
import pytest
import torch

from crfp_b200.synthetic import make_clip, make_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from crfp_b200 import CRFP_DSV
    m = CRFP_DSV("cuda", mid_channels=32).eval()
    m.load_state_dict(make_state_dict(seed=1), strict=True)
    return m.cuda()


@pytest.mark.parametrize("h,w,grid,halo", [(64, 96, (2, 2), 24), (48, 128, (1, 4), 24), (72, 72, (3, 2), 32)])
def test_tiled_equals_untiled(model, h, w, grid, halo):
    from crfp_b200.tiling import TiledClipRunner
    lrs, fvs, mks, _ = make_clip(seed=5, n=1, t=4, h=h, w=w, fv_size=128)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    ref = model(lrs, fvs, mks)
    out = TiledClipRunner(model, grid=grid, halo=halo)(lrs, fvs, mks)
    torch.cuda.synchronize()
    errs = [(out[:, i] - ref[:, i]).abs().max().item() for i in range(out.shape[1])]
    print(f"{h}x{w} grid {grid} halo {halo}: per-frame max-abs tiled vs untiled {['%.2e' % e for e in errs]}")
    assert max(errs) <= 1e-4


def test_too_small_halo_is_visible(model):
    """Sanity of the test itself: with no halo the seams must show (otherwise the comparison proves nothing)."""
    from crfp_b200.tiling import TiledClipRunner
    lrs, fvs, mks, _ = make_clip(seed=5, n=1, t=3, h=64, w=96, fv_size=128)
    lrs, fvs, mks = lrs.cuda(), fvs.cuda(), mks.cuda()
    ref = model(lrs, fvs, mks)
    out = TiledClipRunner(model, grid=(2, 2), halo=0)(lrs, fvs, mks)
    assert (out - ref).abs().max().item() > 1e-3
