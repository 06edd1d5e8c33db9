"""One training iteration of the reference's `Trainer.train_basicvsr` loop (/root/reference/trainer.py:206-293) on the
B200 kernels: forward with autograd over `crfp_b200.training`, Charbonnier loss, backward, ONE fused gradient
all-reduce across ranks, Adam with two parameter groups on a cosine schedule.

  groups      main parameters lr_rate (2e-4, train.sh:12) / `spynet.*` (the FNet) lr_rate_flow (2.5e-5, train.sh:13)
              trainer.py:132-149; Adam betas (0.9, 0.999), eps 1e-12 (option.py:70-74)
  schedule    cosine annealing from the base lr to 1e-7 over 600 000 iterations, no restarts (trainer.py:120-128,
              604-622, annealing_cos :70-83), applied BEFORE each iteration (before_train_iter)
  freeze      FNet parameters do not train during the first 5 000 iterations (trainer.py:223-229)
  loss        rec_w * CharbonnierLoss()(sr.view(B*N,C,H,W), hr.view(B*N,C,H,W))   (trainer.py:233-246)

Parameters, gradients and both Adam moments live in FLAT fp32 buffers ([main | spynet], 2 284 352 elements = 9.14 MB
each): every `p.data` / `p.grad` is a view into them, so the multi-GPU gradient exchange is a single NCCL all-reduce
of one bucket and the optimiser is one `crfp_adam_step` launch per group.
"""
from __future__ import annotations

import math
import os

import torch

from . import autograd as A
from .training import forward_train


def annealing_cos(start, end, factor, weight=1.0):
    """trainer.py:70-83."""
    cos_out = math.cos(math.pi * factor) + 1
    return end + 0.5 * weight * (start - end) * cos_out


class Trainer:
    """Build it AFTER the model sits on its CUDA device and keep the model there: every `p.data` / `p.grad` becomes a
    view into the trainer's flat buffers (a later `model.to(...)` / `.half()` would silently detach them).  One instance
    per process (= per GPU); under torch.distributed every rank calls `step` with its own clips."""

    def __init__(self, model, lr_rate=2e-4, lr_rate_flow=2.5e-5, beta1=0.9, beta2=0.999, eps=1e-12, rec_w=1.0,
                 period=600000, min_lr=1e-7, freeze_flow_iters=5000, process_group=None, kernels=None, use_graphs=False):
        self.model = model
        # use_graphs: capture forward + loss + backward of a step into ONE CUDA graph (per input shape and FNet-freeze
        # phase; captured on the third step of a kind, after two eager ones) and replay it afterwards: a step is
        # ~2 000 small launches plus the autograd tape, i.e. host-bound when launched eagerly.  The gradient
        # all-reduce and the two Adam launches stay outside the graph (their scalars change every iteration).
        self.use_graphs = use_graphs
        self._graphs, self._seen = {}, {}
        self.time_comm, self.comm_events = False, []                 # optional CUDA-event timing of the gradient all-reduce
        self.K = kernels or A.CUDA
        # weight gradients of the conv layers: one launch per layer over all frames at the end of the backward pass (the
        # Trainer reads gradients from the flat .grad views, which is what the deferral writes); CRFP_WGRAD_DEFER=0: per frame
        self.defer_wgrad = os.environ.get("CRFP_WGRAD_DEFER", "1") != "0" and self.K is A.CUDA
        self.betas, self.eps, self.rec_w = (beta1, beta2), eps, rec_w
        self.period, self.min_lr, self.freeze_flow_iters = period, min_lr, freeze_flow_iters
        self.base_lr = [lr_rate, lr_rate_flow]
        self.cur_iter = 0
        self.pg = process_group
        named = list(model.named_parameters())
        groups = [[(k, p) for k, p in named if "spynet" not in k], [(k, p) for k, p in named if "spynet" in k]]
        total = sum(p.numel() for _, p in named)
        dev = named[0][1].device
        self.flat_p = torch.empty(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.ranges, self.group_steps = [], [0, 0]
        off = 0
        for grp in groups:
            start = off
            for _, p in grp:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
                p.grad = self.flat_g[off:off + n].view_as(p)
                off += n
            self.ranges.append((start, off))
        self.group_params = [[p for _, p in grp] for grp in groups]
        self.lr = list(self.base_lr)
        self.sync_parameters()

    def _world(self):
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return torch.distributed.get_world_size(self.pg)
        return 1

    def _src_rank(self):
        """Global rank of group rank 0 (torch.distributed.broadcast takes the GLOBAL rank of the source)."""
        if self.pg is None:
            return 0
        return torch.distributed.get_global_rank(self.pg, 0)

    def sync_parameters(self):
        """Replicas must START identical: the reference trains ONE weight copy under DataParallel (main.py:37-38,
        trainer.py:206-293), so under torch.distributed the flat parameter / moment buffers and the step counters of
        group-rank 0 are broadcast to every rank (as DistributedDataParallel does at construction).  Called by
        __init__ and load_state_dict; a no-op in a single process."""
        if self._world() <= 1:
            return
        src = self._src_rank()
        for buf in (self.flat_p, self.flat_m, self.flat_v):
            torch.distributed.broadcast(buf, src=src, group=self.pg)
        counters = torch.tensor([self.cur_iter] + list(self.group_steps), device=self.flat_p.device, dtype=torch.int64)
        torch.distributed.broadcast(counters, src=src, group=self.pg)
        self.cur_iter, self.group_steps = int(counters[0]), [int(counters[1]), int(counters[2])]
        self.model._packed = None                                    # inference-side packed weights are stale now
        if hasattr(self.model, "_graphs"):
            self.model._graphs.clear()

    # ---- trainer.py:604-622
    def get_lr(self, base_lr):
        alpha = min(self.cur_iter / self.period, 1)
        return annealing_cos(base_lr, self.min_lr, alpha, 1.0)

    def _set_flow_trainable(self):
        train_flow = self.cur_iter >= self.freeze_flow_iters
        for p in self.group_params[1]:
            p.requires_grad_(train_flow)
        return train_flow

    def step(self, lrs, fvs, mks, hr):
        """One iteration on this rank's clips: lrs (n,t,3,h,w), fvs (n,t,3,8h,8w), mks (n,t,1,8h,8w) bool,
        hr (n,t,3,8h,8w).  Returns the (detached) loss tensor of this rank."""
        model = self.model
        model.train()
        train_flow = self._set_flow_trainable()
        self.lr = [self.get_lr(b) for b in self.base_lr]            # before_train_iter
        for grp in self.group_params:                                # keep every .grad a view of the flat bucket
            for p in grp:
                if p.grad is None:
                    raise RuntimeError("a parameter lost its flat gradient view")
        loss = self._fwd_bwd_graphed(lrs, fvs, mks, hr, train_flow) if self.use_graphs else self._fwd_bwd(lrs, fvs, mks, hr)
        ws = self._world()
        if ws > 1:                                                   # ONE all-reduce of the 9.14 MB bucket
            if self.time_comm:                                       # device time of the exchange (bench_train.py)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            torch.distributed.all_reduce(self.flat_g, group=self.pg)
            self.flat_g.mul_(1.0 / ws)
            if self.time_comm:
                e1.record()
                self.comm_events.append((e0, e1))
        self._adam(train_flow)
        self.cur_iter += 1
        model._packed = None                                         # inference-side packed weights are stale now
        if hasattr(model, "_graphs"):
            model._graphs.clear()
        return loss

    def _fwd_bwd(self, lrs, fvs, mks, hr):
        K = self.K
        sr = forward_train(self.model, lrs, fvs, mks, K, defer_wgrad=self.defer_wgrad)
        b, n, c, h, w = sr.shape
        loss = A.charbonnier_loss(K, sr.reshape(b * n, c, h, w), hr.reshape(b * n, c, h, w).to(torch.float32), 1e-12,
                                  self.rec_w)
        self.flat_g.zero_()                                          # optimizer.zero_grad()
        loss.backward()
        return loss.detach()

    def _fwd_bwd_graphed(self, lrs, fvs, mks, hr, train_flow):
        batch = (lrs, fvs, mks, hr)
        key = (tuple((tuple(t_.shape), t_.dtype) for t_ in batch), str(lrs.device), train_flow)
        entry = self._graphs.get(key)
        if entry is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] < 3:                                  # eager warm-ups (allocator, packing caches, autotune-free)
                return self._fwd_bwd(*batch)
            try:
                static = [t_.detach().clone() for t_ in batch]
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    loss = self._fwd_bwd(*static)
                entry = self._graphs[key] = dict(graph=g, static=static, loss=loss)
            except Exception as e:  # noqa: BLE001
                import warnings
                warnings.warn(f"crfp_b200.Trainer: CUDA graph capture failed ({e}); continuing without graphs")
                self.use_graphs = False
                return self._fwd_bwd(*batch)
        for dst, src in zip(entry["static"], batch):
            dst.copy_(src, non_blocking=True)
        entry["graph"].replay()
        return entry["loss"].clone()

    # ---- resume (the reference only saves model weights, trainer.py:276-279; the optimiser state is offered on top)
    def _hyper(self):
        return {"base_lr": list(self.base_lr), "betas": list(self.betas), "eps": self.eps, "rec_w": self.rec_w,
                "period": self.period, "min_lr": self.min_lr, "freeze_flow_iters": self.freeze_flow_iters}

    def state_dict(self):
        return {"cur_iter": self.cur_iter, "group_steps": list(self.group_steps), "exp_avg": self.flat_m.clone(),
                "exp_avg_sq": self.flat_v.clone(), "ranges": [list(r) for r in self.ranges], "hyper": self._hyper()}

    def load_state_dict(self, state, strict_hyper=True):
        """Restores the moments and counters.  The [main | spynet] split of the flat buffers must match (a checkpoint
        of a model with the same element count but another split would put moments on the wrong parameters); the
        schedule / Adam hyper-parameters must match too unless strict_hyper=False (then the checkpoint's win)."""
        if state["exp_avg"].numel() != self.flat_m.numel() or state["exp_avg_sq"].numel() != self.flat_v.numel():
            raise ValueError("optimizer state does not match this model")
        if "ranges" in state and [list(r) for r in state["ranges"]] != [list(r) for r in self.ranges]:
            raise ValueError(f"optimizer state was saved for parameter groups {state['ranges']}, this trainer has {self.ranges}")
        if "hyper" in state:
            mine = self._hyper()
            if state["hyper"] != mine:
                if strict_hyper:
                    diff = {k: (state["hyper"].get(k), v) for k, v in mine.items() if state["hyper"].get(k) != v}
                    raise ValueError(f"optimizer hyper-parameters differ from the checkpoint's (saved, current): {diff}")
                hp = state["hyper"]
                self.base_lr, self.betas, self.eps, self.rec_w = list(hp["base_lr"]), tuple(hp["betas"]), hp["eps"], hp["rec_w"]
                self.period, self.min_lr, self.freeze_flow_iters = hp["period"], hp["min_lr"], hp["freeze_flow_iters"]
        self.cur_iter, self.group_steps = int(state["cur_iter"]), [int(x) for x in state["group_steps"]]
        self.flat_m.copy_(state["exp_avg"].to(self.flat_m.device))
        self.flat_v.copy_(state["exp_avg_sq"].to(self.flat_v.device))
        self._graphs.clear()                                         # captured steps belong to the previous phase
        self._seen.clear()
        self.sync_parameters()

    def _adam(self, train_flow):
        lib, st = self.K.lib(), self.K.stream()
        b1, b2 = self.betas
        for gi, (lo, hi) in enumerate(self.ranges):
            if gi == 1 and not train_flow:                           # params without grad are skipped by torch's Adam
                continue
            self.group_steps[gi] += 1
            t = self.group_steps[gi]
            step_size = self.lr[gi] / (1 - b1 ** t)
            bc2_sqrt = math.sqrt(1 - b2 ** t)
            rc = lib.crfp_adam_step(hi - lo, self.flat_p[lo:hi].data_ptr(), self.flat_g[lo:hi].data_ptr(),
                                    self.flat_m[lo:hi].data_ptr(), self.flat_v[lo:hi].data_ptr(), b1, b2, self.eps,
                                    step_size, bc2_sqrt, st)
            if rc != 0:
                raise RuntimeError(f"crfp_adam_step failed with status {rc}")
