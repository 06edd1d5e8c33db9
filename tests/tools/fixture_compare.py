"""Shared helper of the long-recurrence / full-size GPU tests: compare a model output with the compact reference fixtures
written by oracle/make_golden_long.py (per-frame strided sub-grid, fovea crop, float64 checksums)."""


def compare_with_fixture(out, fix):
    """out (1,t,3,H,W) CPU tensor vs the compact reference fixture -> per-frame max-abs over sub-grid and crop, and the
    per-frame mean error implied by the float64 checksums."""
    t = out.shape[1]
    stride, crop = fix["stride"], fix["crop"]
    errs, mean_errs = [], []
    for i in range(t):
        oy, ox, cy, cx = fix["origins"][i]
        e1 = (out[0, i, :, oy::stride, ox::stride] - fix["grids"][i]).abs().max().item()
        e2 = (out[0, i, :, cy:cy + crop, cx:cx + crop] - fix["crops"][i]).abs().max().item()
        errs.append(max(e1, e2))
        mean_errs.append(abs(float(out[0, i].double().sum()) - fix["sum"][i]) / out[0, i].numel())
    return errs, mean_errs
