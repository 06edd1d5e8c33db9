"""Generate tests/golden/train_*.pt from the REAL reference (run in the build container only): one training iteration's
forward + loss + backward exactly as `Trainer.train_basicvsr` runs it (/root/reference/trainer.py:233-249) — the
reference `CRFP_DSV` in train() mode, the reference's own `CharbonnierLoss` (loss/loss.py:126-176), `loss.backward()`.
TEST INFRASTRUCTURE.  Checks that torch autograd over the ORACLE's forward reproduces every reference gradient, and
saves the loss, the norm of each of the 118 gradients and a few whole gradient tensors as a small fixture.
(The DCNv2 backward is torchvision's here, through the same shim as the forward fixtures: SURVEY.md 8(c).)
Usage: python oracle/make_golden_train.py"""
from __future__ import annotations

import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import REF, build_ref_model, load_reference  # noqa: E402
from crfp_b200.synthetic import make_clip, make_state_dict  # noqa: E402
from oracle import crfp_oracle as O  # noqa: E402

FULL = ["conv_last.weight", "conv_tttf.bias", "dcn_1.dcn.bias", "dcn_2.dcn_offset.bias", "dcn_3.dcn_mask.weight",
        "forward_resblocks_3.main.0.weight", "spynet.flow.2.weight", "upsample_post.upsample_conv.bias"]
CASES = [("train_dsv_n1_t3_8x8", 1, 3, 8, 8, 24), ("train_dsv_n2_t2_8x16", 2, 2, 8, 16, 24)]


def reference_charbonnier():
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "loss", "loss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CharbonnierLoss()


def main():
    CRFP, _ = load_reference()
    cb = reference_charbonnier()
    sd = make_state_dict(seed=1)
    for name, n, t, h, w, fv in CASES:
        lrs, fvs, mks, _ = make_clip(seed=2, n=n, t=t, h=h, w=w, fv_size=fv)
        hr = torch.rand(n, t, 3, 8 * h, 8 * w, generator=torch.Generator().manual_seed(3))
        model = build_ref_model(CRFP.CRFP_DSV, sd).train()
        sr = model(lrs=lrs, fvs=fvs, mks=mks)
        loss = 1.0 * cb(sr.view(n * t, 3, 8 * h, 8 * w), hr.view(n * t, 3, 8 * h, 8 * w))       # trainer.py:235-246
        model.zero_grad()
        loss.backward()
        ref_grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        # the oracle's forward under autograd
        sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
        flows = O.compute_flow(sdg, lrs)
        x_lr, x_hr = O.encoders(sdg, lrs, fvs, mks)
        state, outs = None, []
        for i in range(t):
            o, state = O.frame_step(sdg, 32, state, x_lr[:, i], x_hr[:, i], mks[:, i], lrs[:, i], flows[:, i - 1] if i else None)
            outs.append(o)
        oloss = torch.sqrt((torch.stack(outs, 1) - hr) ** 2 + 1e-12).mean()
        ograds = dict(zip(sdg.keys(), torch.autograd.grad(oloss, list(sdg.values()))))
        worst = max(((ograds[k] - ref_grads[k]).norm() / ref_grads[k].norm()).item() for k in sd)
        print(f"{name}: loss reference {loss.item():.8f} oracle {oloss.item():.8f}; worst relative L2 gradient difference "
              f"oracle vs reference over {len(sd)} tensors: {worst:.3e}")
        assert abs(loss.item() - oloss.item()) < 1e-7 and worst < 1e-5
        torch.save({"case": dict(n=n, t=t, h=h, w=w, fv=fv, seed=2, hr_seed=3), "loss": float(loss),
                    "grad_norms": {k: float(v.norm()) for k, v in ref_grads.items()},
                    "grads": {k: ref_grads[k] for k in FULL},
                    "weights_sum": float(sum(v.double().sum() for v in sd.values()))},
                   os.path.join(ROOT, "tests", "golden", name + ".pt"))


if __name__ == "__main__":
    main()
