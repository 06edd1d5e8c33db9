"""crfp_b200 — B200 (sm_100a) implementation of CRFP's recurrent cross-resolution propagation hot path.

Public surface (mirrors the reference's module API, /root/reference/model/CRFP.py and CRFP_test.py):
    CRFP_DSV / CRFP / CRFP_simple (device, mid_channels=32, ...).forward(lrs, fvs, mks)
    MRCF_simple_v18(...).forward(lrs, fvs, mks, fgs) / clear_states()
    flow_warp(x, flow), DCNv2(...)(input, offset, mask)
    SPyNet(pretrained, device).forward(ref, supp)                       (legacy flow pyramid, model/CRFP.py:554-741)
    Trainer(model, ...).step(lrs, fvs, mks, hr)                         (one iteration of trainer.py:206-293)
    crfp_b200.runtime.MRCF_simple_v18(...).forward(lrs, fvs, warp_size) (model/CRFP_runtime.py:8364-8682)
Everything below these signatures runs in libcrfp_b200.so (include/crfp_b200.h); there is no CPU fallback.
"""
from .model import CRFP, CRFP_DSV, CRFP_simple, MRCF_simple_v18  # noqa: F401
from .ops import DCNv2, flow_warp  # noqa: F401
from .spynet import SPyNet  # noqa: F401
from .trainer import Trainer  # noqa: F401

__all__ = ["CRFP_DSV", "CRFP", "CRFP_simple", "MRCF_simple_v18", "DCNv2", "flow_warp", "SPyNet", "Trainer"]
