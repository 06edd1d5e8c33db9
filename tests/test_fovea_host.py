"""CPU tests (-m "not gpu") of the gaze / scan trajectory restatement against `fovea_generator` of the REAL reference
(tests/golden/f3_fovea_metrics.pt, oracle/make_golden_f3.py): every scan method x 5 clip geometries, including the
geometries on which the reference itself raises."""
import os

import numpy as np
import pytest
import torch

from crfp_b200.fovea import rects_from_positions, scan_positions


@pytest.fixture(scope="module")
def fix(golden_dir):
    return torch.load(os.path.join(golden_dir, "f3_fovea_metrics.pt"))


def test_scan_positions_match_the_reference(fix):
    seen = set()
    for t in fix["trajectories"]:
        args = (t["t"], t["H"], t["W"], t["method"], t["step"], (t["fv"], t["fv"]))
        if "error" in t:
            with pytest.raises(ValueError):
                scan_positions(*args)
            continue
        np.random.seed(t["seed"])
        p = scan_positions(*args)
        assert torch.equal(p, t["fv_sp"]), (t["method"], t["t"], t["H"], t["W"])
        seen.add(t["method"])
    assert seen == {"Hscan", "Vscan", "Cscan", "Zscan", "Rscan", "Evenscan", "DemoHscan", "Dscan"}


def test_rscan_takes_an_explicit_generator():
    a = scan_positions(20, 256, 448, "Rscan", 0.1, (96, 96), rng=np.random.RandomState(5))
    np.random.seed(5)
    b = scan_positions(20, 256, 448, "Rscan", 0.1, (96, 96))
    assert torch.equal(a, b)


def test_rectangles(fix):
    for c in fix["clips"]:
        t, _, h, w = c["fvs"].shape
        r = rects_from_positions(c["fv_sp"][:t], h, w, (c["fv"], c["fv"]), c["method"])
        for i in range(t):
            m = torch.zeros(h, w)
            m[r[i, 0]:r[i, 2], r[i, 1]:r[i, 3]] = 1
            assert torch.equal(m, c["sps"][i, 0]), (c["method"], i)
    with pytest.raises(ValueError):
        rects_from_positions(torch.tensor([[-3, 4]]), 64, 64, (8, 8), "Evenscan")
