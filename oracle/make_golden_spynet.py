"""Generate tests/golden/spynet_*.pt from the REAL reference class `model.CRFP.SPyNet` (run in the build container only).
TEST INFRASTRUCTURE.  Checks that oracle/crfp_oracle.py::spynet reproduces the reference bit for bit on every case and
saves the reference's outputs as fixtures.  Usage: python oracle/make_golden_spynet.py"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import load_reference  # noqa: E402
from crfp_b200.spynet import make_spynet_state_dict, make_spynet_pair  # noqa: E402
from oracle import crfp_oracle as O  # noqa: E402

CASES = [("spynet_n2_64x96", 2, 64, 96), ("spynet_n1_40x72", 1, 40, 72)]   # second: not multiples of 32 (resize + rescale)


def main():
    CRFP, _ = load_reference()
    sd = make_spynet_state_dict(seed=11)
    net = CRFP.SPyNet(None, "cpu").eval()
    own = net.state_dict()
    assert set(own.keys()) == set(sd.keys()) | {"mean", "std"}, sorted(set(own.keys()) ^ set(sd.keys()))[:6]
    net.load_state_dict({**sd, "mean": own["mean"], "std": own["std"]}, strict=True)
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, n, h, w in CASES:
        ref, supp = make_spynet_pair(seed=12, n=n, h=h, w=w)
        with torch.no_grad():
            want = net(ref, supp)
            got = O.spynet(sd, ref, supp)
        err = (want - got).abs().max().item()
        print(f"{name}: reference vs oracle max-abs {err:.3e}, |flow| max {want.abs().max().item():.3f}")
        assert err == 0.0, "the oracle must reproduce the reference bit for bit"
        torch.save({"case": dict(n=n, h=h, w=w, seed=12, wseed=11), "out": want,
                    "weights_sum": float(sum(v.double().sum() for v in sd.values())),
                    "ref_sum": float(ref.double().sum())}, os.path.join(out_dir, name + ".pt"))


if __name__ == "__main__":
    main()
