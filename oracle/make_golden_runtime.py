"""Generate tests/golden/runtime_*.pt from the REAL `model.CRFP_runtime.MRCF_simple_v18` (build container only).
TEST INFRASTRUCTURE.  The runtime file is a CUDA timing harness: it moves a tensor to cuda:0 at import, imports
`memory_profiler`, and records CUDA events inside forward.  It is made importable / runnable on the CPU by shims that do
not touch its arithmetic: a stub `memory_profiler`, no-op `torch.cuda.Event` / `torch.cuda.synchronize`, and `.to(cuda)`
ignored during the import.  Checks that oracle/crfp_oracle.py::runtime_v18_forward reproduces the class bit for bit.
Usage: python oracle/make_golden_runtime.py"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import REF, _install_dcn_shim  # noqa: E402
from crfp_b200.runtime import make_runtime_state_dict  # noqa: E402
from crfp_b200.synthetic import make_clip  # noqa: E402
from oracle import crfp_oracle as O  # noqa: E402

# name, n, t, h, w, fovea (fh, fw), warp_size
CASES = [("runtime_full_n1_t3_16x24", 1, 3, 16, 24, (48, 48), (1080, 1920)),       # warp region = whole frame
         ("runtime_region_n2_t3_16x24", 2, 3, 16, 24, (40, 56), (64, 96))]          # alignment only in the top-left 64x96


def load_runtime_class():
    _install_dcn_shim()
    mp = types.ModuleType("memory_profiler")
    mp.profile = lambda f=None, **k: (f if f is not None else (lambda g: g))
    sys.modules["memory_profiler"] = mp

    class _Event:
        def __init__(self, *a, **k):
            pass

        def record(self, *a, **k):
            pass

        def elapsed_time(self, other):
            return 0.0

    torch.cuda.Event = _Event
    torch.cuda.synchronize = lambda *a, **k: None
    orig_to = torch.Tensor.to

    def to_no_cuda(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, torch.device) and x.type == "cuda"))
        return self if not a and not k else orig_to(self, *a, **k)

    torch.Tensor.to = to_no_cuda
    try:
        sys.path.insert(0, REF)
        from model import CRFP_runtime  # type: ignore
    finally:
        torch.Tensor.to = orig_to
    return CRFP_runtime.MRCF_simple_v18


def main():
    cls = load_runtime_class()
    sd = make_runtime_state_dict(seed=21)
    model = cls("cpu", mid_channels=32).eval()
    own = model.state_dict()
    assert set(own.keys()) == set(sd.keys()), sorted(set(own.keys()) ^ set(sd.keys()))[:8]
    for k, v in own.items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, tuple(v.shape), tuple(sd[k].shape))
    model.load_state_dict(sd, strict=True)
    for name, n, t, h, w, (fh, fw), warp in CASES:
        lrs, _, _, _ = make_clip(seed=22, n=n, t=t, h=h, w=w, fv_size=16)
        fvs = torch.rand(n, t, 3, fh, fw, generator=torch.Generator().manual_seed(23))
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):      # the class prints its timers
            want = model(lrs.clone(), fvs.clone(), warp_size=warp)
        got = O.runtime_v18_forward(sd, lrs, fvs, warp)
        err = (want - got).abs().max().item()
        print(f"{name}: reference vs oracle max-abs {err:.3e}; out range [{want.min().item():.3f}, {want.max().item():.3f}]")
        assert err == 0.0, "the oracle must reproduce the reference bit for bit"
        torch.save({"case": dict(n=n, t=t, h=h, w=w, fh=fh, fw=fw, warp=warp, seed=22, fv_seed=23, wseed=21), "out": want,
                    "weights_sum": float(sum(v.double().sum() for v in sd.values()))},
                   os.path.join(ROOT, "tests", "golden", name + ".pt"))


if __name__ == "__main__":
    main()
