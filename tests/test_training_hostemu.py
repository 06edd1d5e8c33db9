"""CPU check of the whole training step's wiring: `crfp_b200.training.forward_train` + Charbonnier + backward, with the
backward kernels of bwd.cu running through the host emulation (tests/tools/hostemu), against torch autograd over the
ORACLE's forward (the reference trains through exactly that graph: model/CRFP.py:1510-1686 + trainer.py:233-250).
Also the Trainer step (two Adam groups, cosine schedule, FNet freeze) against torch.optim.Adam on the oracle grads.
The GPU twin (real kernels on the B200) is tests/test_gpu_zz_training.py."""
import math

import pytest
import torch

import hostemu
from crfp_b200 import CRFP_DSV
from crfp_b200.synthetic import make_clip, make_state_dict
from crfp_b200.trainer import Trainer, annealing_cos
from crfp_b200.training import forward_train, pixel_shuffle_nhwc, pixel_unshuffle_nhwc
from oracle import crfp_oracle as O

K = hostemu.HostEmuKernelSet()


def oracle_forward_with_grad(sd, lrs, fvs, mks, C=32):
    """O.crfp_dsv_forward without its no_grad decorator (same calls, same order)."""
    n, t, c, h, w = lrs.shape
    flows = O.compute_flow(sd, lrs) if t > 1 else None
    x_lr, x_hr = O.encoders(sd, lrs, fvs, mks)
    state, outs = None, []
    for i in range(t):
        out, state = O.frame_step(sd, C, state, x_lr[:, i], x_hr[:, i], mks[:, i], lrs[:, i],
                                  flows[:, i - 1] if i > 0 else None)
        outs.append(out)
    return torch.stack(outs, dim=1)


def charbonnier(sr, hr):
    return torch.sqrt((sr - hr) ** 2 + 1e-12).mean()


def _setup(seed=1, n=1, t=3, h=8, w=8):
    sd = make_state_dict(seed=seed)
    lrs, fvs, mks, _ = make_clip(seed=seed + 1, n=n, t=t, h=h, w=w, fv_size=24)
    g = torch.Generator().manual_seed(seed + 2)
    hr = torch.rand(n, t, 3, 8 * h, 8 * w, generator=g)
    model = CRFP_DSV("cuda", mid_channels=32)
    model.load_state_dict(sd, strict=True)
    return sd, model, lrs, fvs, mks, hr


def test_layout_views_match_torch():
    x = torch.randn(2, 32, 5, 7)
    ref = torch.nn.functional.pixel_shuffle(x, 4)
    got = pixel_shuffle_nhwc(x.permute(0, 2, 3, 1).contiguous(), 4).permute(0, 3, 1, 2)
    assert torch.equal(got, ref)
    y = torch.randn(2, 4, 8, 12)
    ref = torch.nn.functional.pixel_unshuffle(y, 4)
    got = pixel_unshuffle_nhwc(y.permute(0, 2, 3, 1).contiguous(), 4).permute(0, 3, 1, 2)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("n,t,h,w", [(1, 3, 8, 8), (2, 2, 8, 16),
                                     (1, 2, 12, 20)])      # h, w not multiples of 8: FNet resamples 8x16 -> 12x20
def test_forward_and_every_parameter_gradient_match_the_oracle(n, t, h, w):
    sd, model, lrs, fvs, mks, hr = _setup(1, n, t, h, w)
    model.train()
    sr = forward_train(model, lrs, fvs, mks, K)
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref = oracle_forward_with_grad(sdg, lrs, fvs, mks)
    assert sr.shape == ref.shape and (sr - ref).abs().max().item() < 1e-4
    loss = charbonnier(sr, hr)
    ref_loss = charbonnier(ref, hr)
    assert abs(loss.item() - ref_loss.item()) < 1e-6
    names = list(sdg.keys())
    ref_grads = torch.autograd.grad(ref_loss, [sdg[k] for k in names])
    loss.backward()
    params = dict(model.named_parameters())
    worst = 0.0
    for k, rg in zip(names, ref_grads):
        g = params[k].grad
        assert g is not None, k
        scale = rg.abs().max().item()
        assert scale > 0, f"{k}: zero reference gradient (the test would not exercise it)"
        # d(bilinear)/d(position) jumps where a DCN / warp sample crosses an integer coordinate, and the two forwards
        # differ by ~1e-5 in the offsets (fp32 summation order), so a handful of samples legitimately take the other
        # one-sided derivative: bound the error in the L2 sense tightly and the single worst element loosely
        rel2 = ((g - rg).norm() / rg.norm()).item()
        relmax = (g - rg).abs().max().item() / scale
        worst = max(worst, rel2)
        # (the ragged case has ONE recurrent frame of 24x40 L1 pixels: a single flipped sample weighs up to 7e-3 in dcn_2's
        # offset path; the FNet gradients, which pass through the resize backward, stay below 2e-3)
        assert rel2 < (1e-2 if h % 8 else 2e-3) and relmax < (6e-2 if h % 8 else 3e-2), (k, rel2, relmax, scale)
    print(f"worst relative L2 gradient error over {len(names)} tensors: {worst:.2e}")


def test_single_frame_clip_trains_everything_but_the_alignment_path():
    """t == 1: no flow, no warp, no DCN (CRFP.py:1637-1668 branch only): their parameters get no gradient at all, like in
    the reference; every other gradient matches."""
    sd, model, lrs, fvs, mks, hr = _setup(7, 2, 1, 8, 8)
    model.train()
    sr = forward_train(model, lrs, fvs, mks, K)
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref = oracle_forward_with_grad(sdg, lrs, fvs, mks)
    assert (sr - ref).abs().max().item() < 1e-4
    names = list(sdg.keys())
    ref_grads = torch.autograd.grad(charbonnier(ref, hr), [sdg[k] for k in names], allow_unused=True)
    charbonnier(sr, hr).backward()
    params = dict(model.named_parameters())
    unused = 0
    for k, rg in zip(names, ref_grads):
        g = params[k].grad
        if rg is None:
            unused += 1
            assert g is None or g.abs().max().item() == 0, k
            assert k.startswith(("spynet.", "dcn_", "downsample.")), k
        else:
            assert ((g - rg).norm() / rg.norm()).item() < 1e-3, k
    assert unused > 40


def test_trainer_step_matches_torch_adam_on_the_oracle():
    """Three Trainer iterations.  Per iteration: the loss equals the oracle's loss AT THE SAME PARAMETERS, the lr follows
    the reference's cosine schedule, and the parameter update equals torch.optim.Adam's (the reference's optimiser,
    trainer.py:149) fed with the same gradients — incl. the FNet freeze (no gradient, no Adam state advance).  The
    gradients themselves are checked against the oracle in the test above; chaining the two optimisers freely instead
    would only measure how Adam (eps 1e-12, a sign-like first step) amplifies fp32 noise in near-zero gradients."""
    sd, model, lrs, fvs, mks, hr = _setup(3, 1, 2, 8, 8)
    ref_p = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    main = [p for k, p in ref_p.items() if "spynet" not in k]
    flow = [p for k, p in ref_p.items() if "spynet" in k]
    opt = torch.optim.Adam([{"params": main, "lr": 2e-4}, {"params": flow, "lr": 2.5e-5}], lr=2e-4, betas=(0.9, 0.999),
                           eps=1e-12)
    tr = Trainer(model, freeze_flow_iters=1, period=10, kernels=K)
    assert tr.flat_p.numel() == 2284352 and tr.ranges[0][1] == tr.ranges[1][0]
    params = dict(model.named_parameters())
    losses = []
    for it in range(3):
        with torch.no_grad():
            ref_loss = charbonnier(oracle_forward_with_grad({k: v.detach() for k, v in ref_p.items()}, lrs, fvs, mks), hr)
        for grp, base in zip(opt.param_groups, (2e-4, 2.5e-5)):          # before_train_iter, trainer.py:604-622
            grp["lr"] = annealing_cos(base, 1e-7, min(it / 10, 1))
        loss = tr.step(lrs, fvs, mks, hr)
        losses.append(loss.item())
        assert abs(loss.item() - ref_loss.item()) < 1e-5, it      # fp32 sum order of 12 288 terms
        assert math.isclose(tr.lr[0], opt.param_groups[0]["lr"]) and math.isclose(tr.lr[1], opt.param_groups[1]["lr"])
        for k, rp in ref_p.items():
            frozen = "spynet" in k and it < 1                            # trainer.py:223-229
            assert params[k].requires_grad == (not frozen)
            rp.grad = None if frozen else params[k].grad.detach().clone()
            if frozen:
                assert params[k].grad.abs().max().item() == 0
        opt.step()
        for k, rp in ref_p.items():
            assert (params[k].detach() - rp.detach()).abs().max().item() < 1e-6, (it, k)
            with torch.no_grad():
                rp.copy_(params[k])                                       # stay on the same trajectory
    for k in sd:
        assert (params[k].detach() - sd[k]).abs().max().item() > 0, f"{k} did not train"
    assert tr.group_steps == [3, 2]
    assert losses[2] < losses[0]


def _ddp_worker(rank, world, port, out_dir):
    """world_size 2 over gloo: each rank trains on its own clip; ONE all-reduce of the flat gradient bucket."""
    import os
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd, model, lrs, fvs, mks, hr = _setup(5, 2, 2, 8, 8)
        tr = Trainer(model, freeze_flow_iters=0, kernels=hostemu.HostEmuKernelSet())
        sl = slice(rank, rank + 1)
        loss = tr.step(lrs[sl], fvs[sl], mks[sl], hr[sl])
        torch.save({"p": tr.flat_p.clone(), "g": tr.flat_g.clone(), "loss": loss}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_step_equals_the_single_process_step_on_both_clips(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_ddp_worker, args=(2, 29541, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in (0, 1))
    assert torch.equal(r0["p"], r1["p"]) and torch.equal(r0["g"], r1["g"])       # replicas stay identical
    sd, model, lrs, fvs, mks, hr = _setup(5, 2, 2, 8, 8)
    tr = Trainer(model, freeze_flow_iters=0, kernels=K)
    loss = tr.step(lrs, fvs, mks, hr)                                              # both clips in one process
    assert abs(loss.item() - 0.5 * (r0["loss"].item() + r1["loss"].item())) < 1e-6
    gn = tr.flat_g.norm().item()
    assert (tr.flat_g - r0["g"]).norm().item() < 1e-4 * gn                         # mean of per-rank grads == batch grad


@pytest.mark.parametrize("name", ["train_dsv_n1_t3_8x8", "train_dsv_n2_t2_8x16"])
def test_loss_and_gradients_match_the_real_reference_fixture(name, golden_dir):
    """tests/golden/train_*.pt = loss and gradients of the REAL reference (`CRFP_DSV.train()`, the reference's own
    CharbonnierLoss, loss.backward(); oracle/make_golden_train.py).  Both the oracle's autograd and the product's
    training step (backward kernels through the host emulation) must reproduce them."""
    import os
    fix = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fix["case"]
    sd = make_state_dict(seed=1)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - fix["weights_sum"]) < 1e-6
    lrs, fvs, mks, _ = make_clip(seed=c["seed"], n=c["n"], t=c["t"], h=c["h"], w=c["w"], fv_size=c["fv"])
    hr = torch.rand(c["n"], c["t"], 3, 8 * c["h"], 8 * c["w"], generator=torch.Generator().manual_seed(c["hr_seed"]))
    # the oracle
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    oloss = charbonnier(oracle_forward_with_grad(sdg, lrs, fvs, mks), hr)
    ograds = dict(zip(sdg.keys(), torch.autograd.grad(oloss, list(sdg.values()))))
    assert abs(oloss.item() - fix["loss"]) < 1e-6
    for k, nrm in fix["grad_norms"].items():
        assert abs(ograds[k].norm().item() - nrm) <= 1e-5 * nrm, k
    for k, g in fix["grads"].items():
        assert (ograds[k] - g).abs().max().item() <= 1e-6 * g.abs().max().item(), k
    # the product path
    model = CRFP_DSV("cuda", mid_channels=32)
    model.load_state_dict(sd, strict=True)
    model.train()
    loss = charbonnier(forward_train(model, lrs, fvs, mks, K), hr)
    loss.backward()
    assert abs(loss.item() - fix["loss"]) < 1e-5
    params = dict(model.named_parameters())
    for k, nrm in fix["grad_norms"].items():
        assert abs(params[k].grad.norm().item() - nrm) <= 5e-3 * nrm, k
    for k, g in fix["grads"].items():
        assert ((params[k].grad - g).norm() / g.norm()).item() < 2e-3, k


def test_graph_capture_failure_falls_back_to_eager_steps():
    """Trainer(use_graphs=True) where capture is impossible (no CUDA here): the third step's capture attempt raises, the
    trainer warns, switches graphs off and keeps training eagerly with consistent state."""
    sd, model, lrs, fvs, mks, hr = _setup(3, 1, 2, 8, 8)
    tr = Trainer(model, freeze_flow_iters=0, kernels=K, use_graphs=True)
    losses = [tr.step(lrs, fvs, mks, hr).item() for _ in range(2)]
    with pytest.warns(UserWarning, match="graph capture failed"):
        losses.append(tr.step(lrs, fvs, mks, hr).item())
    assert tr.use_graphs is False and not tr._graphs
    losses.append(tr.step(lrs, fvs, mks, hr).item())
    assert tr.cur_iter == 4 and tr.group_steps == [4, 4] and losses[-1] < losses[0]


def test_trainer_resume_continues_the_same_trajectory():
    """model.state_dict() + Trainer.state_dict() after 2 steps, loaded into a fresh model / trainer: step 3 is identical."""
    sd, model, lrs, fvs, mks, hr = _setup(3, 1, 2, 8, 8)
    tr = Trainer(model, freeze_flow_iters=1, period=10, kernels=K)
    for _ in range(2):
        tr.step(lrs, fvs, mks, hr)
    ckpt_model = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ckpt_opt = tr.state_dict()
    loss_a = tr.step(lrs, fvs, mks, hr)
    model_b = CRFP_DSV("cuda", mid_channels=32)
    model_b.load_state_dict(ckpt_model, strict=True)
    tr_b = Trainer(model_b, freeze_flow_iters=1, period=10, kernels=K)
    tr_b.load_state_dict(ckpt_opt)
    loss_b = tr_b.step(lrs, fvs, mks, hr)
    assert loss_a.item() == loss_b.item() and tr_b.cur_iter == 3 and tr_b.group_steps == [3, 2]
    assert torch.equal(tr.flat_p, tr_b.flat_p) and torch.equal(tr.flat_m, tr_b.flat_m)


def _ddp_seed_worker(rank, world, port, out_dir):
    """Ranks built from DIFFERENT weights (what independent `kaiming_normal_` inits or different checkpoints give): the
    Trainer must broadcast rank 0's parameters at construction so the replicas start — and stay — identical."""
    import os
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, model, lrs, fvs, mks, hr = _setup(5, 2, 2, 8, 8)
        model.load_state_dict(make_state_dict(seed=11 + rank), strict=True)        # rank-dependent start
        before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
        tr = Trainer(model, freeze_flow_iters=0, kernels=hostemu.HostEmuKernelSet())
        start = tr.flat_p.clone()
        sl = slice(rank, rank + 1)
        tr.step(lrs[sl], fvs[sl], mks[sl], hr[sl])
        st = tr.state_dict()
        st["cur_iter"] += 7 * rank                                                 # rank-dependent resume state
        st["exp_avg"] = st["exp_avg"] + float(rank)
        tr.load_state_dict(st)
        torch.save({"before": before, "start": start, "p": tr.flat_p.clone(), "m": tr.flat_m.clone(), "it": tr.cur_iter},
                   os.path.join(out_dir, f"s{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_replicas_are_synchronised_from_rank_zero(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_ddp_seed_worker, args=(2, 29547, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"s{r}.pt") for r in (0, 1))
    assert not torch.equal(r0["before"], r1["before"])                            # the ranks really started apart
    assert torch.equal(r0["start"], r1["start"])                                   # identical after construction
    # rank 0's parameters win (the flat buffer is ordered [main | spynet], so compare order-free sums)
    assert abs(float(r0["start"].double().sum()) - float(r0["before"].double().sum())) < 1e-6
    assert torch.equal(r0["p"], r1["p"])                                           # still identical after a step
    assert torch.equal(r0["m"], r1["m"]) and r0["it"] == r1["it"] == 1             # load_state_dict re-syncs from rank 0


def test_trainer_state_dict_is_validated():
    sd, model, lrs, fvs, mks, hr = _setup(3, 1, 2, 8, 8)
    tr = Trainer(model, freeze_flow_iters=0, period=10, kernels=K)
    tr.step(lrs, fvs, mks, hr)
    st = tr.state_dict()
    assert st["ranges"] == [list(r) for r in tr.ranges] and st["hyper"]["period"] == 10
    bad = dict(st, ranges=[[0, 5], [5, st["exp_avg"].numel()]])
    with pytest.raises(ValueError, match="parameter groups"):
        tr.load_state_dict(bad)
    model2 = CRFP_DSV("cuda", mid_channels=32)
    model2.load_state_dict(sd, strict=True)
    tr2 = Trainer(model2, freeze_flow_iters=0, period=20, kernels=K)
    with pytest.raises(ValueError, match="hyper-parameters"):
        tr2.load_state_dict(st)
    tr2._graphs["x"], tr2._seen["x"] = 1, 1
    tr2.load_state_dict(st, strict_hyper=False)                                    # the checkpoint's schedule wins
    assert tr2.period == 10 and tr2.cur_iter == 1 and not tr2._graphs and not tr2._seen


def test_deferred_weight_gradients_equal_the_per_frame_ones():
    """forward_train(defer_wgrad=True): the conv weight gradients are collected per layer during the backward pass and
    launched once per layer and source when it ends (autograd.WgradDeferral, torch's queue_callback), incl. the fused
    offset / mask heads whose gradient is routed back to the two reference parameters — same .grad as the per-frame path."""
    grads = {}
    for defer in (False, True):
        sd, model, lrs, fvs, mks, hr = _setup(1, 2, 3, 8, 8)
        model.train()
        sr = forward_train(model, lrs, fvs, mks, K, defer_wgrad=defer)
        charbonnier(sr, hr).backward()
        grads[defer] = {k: p.grad.clone() for k, p in model.named_parameters()}
        assert all(g is not None for g in grads[defer].values())
    for k, g in grads[False].items():
        d = grads[True][k]
        assert d.shape == g.shape
        assert (d - g).abs().max().item() <= 1e-5 * g.abs().max().item() + 1e-9, k      # summation order over frames
    # a second backward through a fresh graph accumulates into .grad like autograd does
    sd, model, lrs, fvs, mks, hr = _setup(1, 1, 2, 8, 8)
    model.train()
    for _ in range(2):
        charbonnier(forward_train(model, lrs, fvs, mks, K, defer_wgrad=True), hr).backward()
    once = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    charbonnier(forward_train(model, lrs, fvs, mks, K, defer_wgrad=True), hr).backward()
    for k, p in model.named_parameters():
        assert (once[k] - 2 * p.grad).abs().max().item() <= 2e-5 * p.grad.abs().max().item() + 1e-9, k
