"""torch.autograd bindings of the training kernels: every Function's forward AND backward run in libcrfp_b200.so.

The reference trains through ATen/cuDNN autograd and `dcn_v2`'s own backward (`loss.backward()`,
/root/reference/trainer.py:246-250).  Here torch.autograd is only the tape (it records which op consumed which
tensor across the t-frame recurrence, i.e. BPTT); the arithmetic of each node is a hand-written kernel behind the C ABI:

  Conv3x3Fn     crfp_conv3x3_fwd  | crfp_act_bwd; data: crfp_conv3x3_fwd again on the rotated / transposed weights (or the
                                    gather kernels crfp_conv3x3_bwd_data, CRFP_DGRAD=direct); crfp_conv3x3_bwd_weight   (nn.Conv2d + act)
  DCNv2Fn       crfp_dcn_v2_fwd   | crfp_dcn_v2_bwd                        (dcn_v2.DCNv2, model/CRFP.py:350)
  FlowWarpFn    crfp_flow_warp_fwd| crfp_flow_warp_bwd                     (flow_warp, model/CRFP.py:90-130)
  ResizeFn      crfp_resize_bilinear | crfp_resize_bilinear_bwd            (nn.Upsample / F.interpolate, bilinear)
  AvgPool2Fn    crfp_avgpool2     | crfp_avgpool2_bwd                      (nn.AvgPool2d(2,2))
  CharbonnierFn crfp_charbonnier_fwd_bwd                                   (loss/loss.py:116-124)

All tensors are dense fp32 NHWC.  `KernelSet` is the only place that touches the library: the product instance
(`CUDA`) requires CUDA tensors and raises otherwise — there is no CPU implementation in this package.  (The CPU test
suite injects its own kernel set built from a host emulation of bwd.cu, see tests/tools/hostemu.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L

ACT_NONE, ACT_LRELU, ACT_RELU = L.ACT_NONE, L.ACT_LRELU, L.ACT_RELU


class KernelSet:
    """Forward primitives + the C-ABI handle whose crfp_*_bwd entry points the Functions below call."""

    name = "cuda"
    # backward-data through the tiled forward conv kernel (default) or the direct gather kernel crfp_conv3x3_bwd_data
    # (CRFP_DGRAD=direct; kept for A/B and for shapes the forward kernel would not take)
    dgrad_as_conv = os.environ.get("CRFP_DGRAD", "conv") != "direct"
    # The weight-gradient entry points take an optional workspace (partial sums + a reduce kernel instead of same-address
    # atomics).  The library says how much it wants: the round-2 tiled kernel (wgrad.cu) for the wide layers; 0 for the
    # shapes that do not need one.  CRFP_WGRAD_V1=1 (read by the library) switches back to round 1's kernels, for which
    # CRFP_WGRAD_THIN=2stage selects their two-stage variant on the thin layers (measured slower: 81.6 vs 76.5 ms).
    wgrad_two_stage = os.environ.get("CRFP_WGRAD_V1") is None or os.environ.get("CRFP_WGRAD_THIN", "atomic") == "2stage"

    def lib(self):
        return L.lib()

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def req(self, t: torch.Tensor, what: str) -> torch.Tensor:
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise L.CrfpError(f"{what} must be a CUDA tensor (libcrfp_b200 has no CPU path)")
        if t.dtype != torch.float32:
            raise L.CrfpError(f"{what} must be float32, got {t.dtype}")
        return t.contiguous()

    # ---- forward primitives (GPU-verified SIMT fp32 kernels of the inference library)
    # `cache` (a dict owned by the caller, valid while the weights do not change — one training step): the packed
    # weight layouts are built once per layer and step instead of once per frame
    # training forward / backward-data through the tensor-core conv (crfp_conv3x3_tc3_fwd: 3 x bf16 split products, fp32
    # accumulation, ~1e-6 relative) wherever its shape rules hold; CRFP_TRAIN_TC=0: everything on the fp32 SIMT kernels
    train_tc = os.environ.get("CRFP_TRAIN_TC", "1") != "0"

    @staticmethod
    def _tc3_split(c_list, cout):
        """(tensor-core channel list, extra channels) if crfp_conv3x3_tc3_fwd takes this conv, else None."""
        from .packing import tc3_cout_tile
        c_list = list(c_list)
        extra = 0
        if len(c_list) > 1 and c_list[-1] == 2:
            extra, c_list = 2, c_list[:-1]
        k = sum(c_list)
        if not c_list or len(c_list) > 3 or any(c % 8 for c in c_list) or k > 64 or cout % 4 or tc3_cout_tile(cout, k) is None:
            return None
        return c_list, extra

    def conv3x3(self, srcs, weight, bias, act, cache=None):
        from . import ops
        from .packing import pack_conv
        if self.train_tc and act in (ACT_NONE, ACT_LRELU, ACT_RELU) and not (cache is not None and cache.get("no_tc")):
            if len(srcs) == 1 and srcs[0].shape[-1] in (128, 192, 256) and self._tc3_split([64], weight.shape[0]) is not None:
                return self._conv3x3_ksplit(srcs[0], weight, bias, act, cache, transposed=False, lo=0, nout=weight.shape[0])
            sp = self._tc3_split([s.shape[-1] for s in srcs], weight.shape[0])
            if sp is not None:
                from .packing import pack_conv_tc3_device
                tc, extra = sp
                packed = cache.get("fwd_tc") if cache is not None else None
                if packed is None:
                    packed = pack_conv_tc3_device(weight, bias, sum(tc), extra)
                    if cache is not None:
                        cache["fwd_tc"] = packed
                return ops.conv3x3_tc3_nhwc(srcs[:len(tc)], None, None, act=act, extra=srcs[-1] if extra else None,
                                            packed=packed, cout=weight.shape[0])
        packed = None
        if cache is not None:
            packed = cache.get("fwd")
            if packed is None:
                packed = cache["fwd"] = pack_conv(weight, bias, [s.shape[-1] for s in srcs])
        return ops.conv3x3_nhwc(srcs, weight, bias, act=act, packed=packed)

    def _conv3x3_ksplit(self, x, weight, bias, act, cache, transposed, lo, nout):
        """A conv over 128 / 192 / 256 input channels as passes of 64 through the tensor-core kernel: pass 1 adds the bias,
        the later passes add the partial sums (in place, before the activation: crfp_conv_tc3_desc.res_pre), the last one
        applies the activation.  transposed: backward data (x = dy, K = the layer's output channels)."""
        from . import ops
        from .packing import pack_conv_tc3_device
        key = "ksplit_t" if transposed else "ksplit"
        packs = cache.get(key) if cache is not None else None
        passes = x.shape[-1] // 64
        if packs is None:
            packs = []
            for k in range(passes):
                if transposed:
                    packs.append(pack_conv_tc3_device(weight, None, 64, 0, lo=lo, transposed=True, nout=nout, k_lo=64 * k))
                else:
                    packs.append(pack_conv_tc3_device(weight, bias if k == 0 else None, 64, 0, lo=64 * k))
            if cache is not None:
                cache[key] = packs
        out = None
        for k in range(passes):
            out = ops.conv3x3_tc3_nhwc([x], None, None, act=act if k == passes - 1 else ACT_NONE, packed=packs[k], cout=nout,
                                       src_slice=(64 * k, 64), residual=out, res_pre=k > 0, out=out)
        return out

    def conv3x3_dgrad(self, g, weight, off, c, cache):
        """Backward data of input channels [off, off + c) as a tensor-core conv of g with the transposed, rotated kernel
        (packed on the device straight from the OIHW parameter); None when the shape is not the tensor-core kernel's."""
        if not self.train_tc:
            return None
        if g.shape[-1] in (128, 192, 256) and self._tc3_split([64], c) is not None:
            return self._conv3x3_ksplit(g, weight, None, ACT_NONE, cache, transposed=True, lo=off, nout=c)
        if self._tc3_split([g.shape[-1]], c) is None:
            return None
        from . import ops
        from .packing import pack_conv_tc3_device
        packed = cache.get("dgrad_tc")
        if packed is None:
            packed = cache["dgrad_tc"] = pack_conv_tc3_device(weight, None, g.shape[-1], 0, lo=off, transposed=True, nout=c)
        return ops.conv3x3_tc3_nhwc([g], None, None, act=ACT_NONE, packed=packed, cout=c)

    def dcn_v2(self, x, offset, mask, weight, bias, dg, cache=None):
        from . import ops
        from .packing import pack_dcn
        if self.train_tc and dg == 8 and x.shape[-1] == 32 and weight.shape[0] == 32:
            # the L1 align op on the persistent TMA + tcgen05 kernel (crfp_dcn_v2_tc3_fwd: 3 x bf16 split contraction);
            # cache["hint"] = the flow (any hint gives the same result, a good one keeps the samples inside the TMA window)
            from .packing import pack_dcn_tc3
            packed = cache.get("fwd_tc") if cache is not None else None
            if packed is None:
                packed = pack_dcn_tc3(weight, bias, dg)
                if cache is not None:
                    cache["fwd_tc"] = packed
            hi, lo, bp = packed
            n, h, w, c = x.shape
            out = torch.empty(n, h, w, 32, device=x.device, dtype=torch.float32)
            d = L.DcnDesc(n=n, h=h, w=w, c=c, cout=32, dg=dg, shared_taps=0, x=x.data_ptr(), x_cstride=c, x_coffset=0,
                          offset=offset.data_ptr(), off_cstride=offset.shape[-1], off_coffset=0,
                          mask=mask.data_ptr(), mask_cstride=mask.shape[-1], mask_coffset=0,
                          weight=hi.data_ptr(), bias=bp.data_ptr(), out=out.data_ptr(), out_cstride=32, out_coffset=0)
            hint = cache.get("hint") if cache is not None else None
            if hint is not None:
                hint = self.req(hint.detach(), "flow hint")
            L.check(L.lib().crfp_dcn_v2_tc3_fwd(C.byref(d), lo.data_ptr(), hint.data_ptr() if hint is not None else None,
                                                self.stream()), "dcn_v2_tc3")
            return out
        packed = cache.get("fwd") if cache is not None else None
        if packed is None:
            packed = pack_dcn(weight, bias, dg)
            if cache is not None:
                cache["fwd"] = packed
        return ops.dcn_v2_nhwc(x, offset, mask, packed[0], packed[1], dg, weight.shape[0])

    def flow_warp(self, x, flow):
        from . import ops
        return ops.flow_warp_nhwc(x, flow)

    def resize(self, x, hout, wout, rh, rw, mul):
        from . import ops
        return ops.resize_bilinear_nhwc(x, hout, wout, rh, rw, mul)

    def avgpool2(self, x):
        from . import ops
        return ops.avgpool2_nhwc(x)

    def to_nhwc(self, x):
        from . import ops
        return ops.to_nhwc(x)


CUDA = KernelSet()


def _chk(K, status, what):
    if status != 0:
        if K is CUDA:
            L.check(status, what)
        raise L.CrfpError(f"{what} failed with status {status}")


# ----------------------------------------------------------------------------------------------- deferred weight gradients
class WgradDeferral:
    """The weight gradient of a conv layer is a sum over the t frames of the recurrence (and over the sources of a concat).
    Instead of one launch + one zero fill + one accumulation add per layer AND frame inside the backward pass, the
    (x, dy) pairs of a layer are collected and ONE crfp_conv3x3_bwd_weight_batched launch per layer and source runs when
    the pass ends (torch's queue_callback: after the last node, on the thread that called backward(), inside a CUDA-graph
    capture if there is one).  The result is added into `param.grad` directly, so this is only used where the caller reads
    gradients from `.grad` (the Trainer): `torch.autograd.grad` would see None for the deferred weights."""

    def __init__(self, K):
        self.K = K
        self.layers = []
        self.armed = False

    def add(self, cache, srcs, g, c_list, need_w, need_b):
        pend = cache.setdefault("wg_pending", [])
        if not pend:
            self.layers.append((cache, tuple(c_list), need_w, need_b))
        pend.append((srcs, g))
        if not self.armed:
            self.armed = True
            torch.autograd.Variable._execution_engine.queue_callback(self.flush)

    def flush(self):
        K = self.K
        lib, st = K.lib(), K.stream()
        layers, self.layers, self.armed = self.layers, [], False
        if not layers:
            return
        # ONE zero fill for every layer's dw / db accumulator (views of a flat buffer) instead of two fills per layer
        sizes = []
        for cache, c_list, _, _ in layers:
            cout = cache["wg_pending"][0][1].shape[-1]
            sizes.append((9 * sum(c_list) * cout, cout))
        dev0 = layers[0][0]["wg_pending"][0][1].device
        flat = torch.zeros(sum(a + b + (-(a + b)) % 4 for a, b in sizes), device=dev0, dtype=torch.float32)
        pos = 0
        for (cache, c_list, need_w, need_b), (nw, nb) in zip(layers, sizes):
            pend = cache.pop("wg_pending")
            g0 = pend[0][1]
            n, h, w, cout = g0.shape
            cin = sum(c_list)
            dev = g0.device
            dw = flat[pos:pos + nw].view(9, cin, cout)
            dbt = flat[pos + nw:pos + nw + nb]
            pos += nw + nb + (-(nw + nb)) % 4            # keep every accumulator 16-byte aligned
            cnt = len(pend)
            gs = (C.c_void_p * cnt)(*[g.data_ptr() for _, g in pend])
            off = 0
            for i, c in enumerate(c_list):
                xs = (C.c_void_p * cnt)(*[srcs[i].data_ptr() for srcs, _ in pend])
                ws_floats = lib.crfp_conv3x3_bwd_weight_workspace(n * min(cnt, 16), h, w, c, cout)
                ws = torch.empty(ws_floats, device=dev, dtype=torch.float32) if ws_floats else None
                _chk(K, lib.crfp_conv3x3_bwd_weight_batched(cnt, xs, gs, n, h, w, c, cout, cin, off, dw.data_ptr(),
                                                            dbt.data_ptr() if i == 0 else None,
                                                            ws.data_ptr() if ws is not None else None, ws_floats, st),
                     "conv3x3_bwd_weight_batched")
                off += c
            dW = dw.permute(2, 1, 0).reshape(cout, cin, 3, 3)
            for param, lo, hi in cache["wparams"] if need_w else ():
                if param.requires_grad:
                    param.grad = dW[lo:hi].contiguous() if param.grad is None else param.grad.add_(dW[lo:hi])
            for param, lo, hi in cache["bparams"] if need_b else ():
                if param.requires_grad:
                    param.grad = dbt[lo:hi].clone() if param.grad is None else param.grad.add_(dbt[lo:hi])


# ----------------------------------------------------------------------------------------------- conv3x3 (+bias+act)
class Conv3x3Fn(torch.autograd.Function):
    """act(conv3x3(cat(srcs, -1), weight) + bias); weight OIHW (the reference's nn.Conv2d parameter)."""

    @staticmethod
    def forward(ctx, K, act, cache, weight, bias, *srcs):
        srcs = [K.req(s.detach(), "conv source") for s in srcs]
        out = K.conv3x3(srcs, weight.detach(), bias.detach(), act, cache)
        ctx.K, ctx.act, ctx.c_list, ctx.cache = K, act, [s.shape[-1] for s in srcs], cache
        ctx.save_for_backward(weight, out if act != ACT_NONE else None, *srcs)
        return out

    @staticmethod
    def backward(ctx, dy):
        K, act = ctx.K, ctx.act
        weight, out, *srcs = ctx.saved_tensors
        lib, st = K.lib(), K.stream()
        dy = K.req(dy, "grad_output")
        n, h, w, cout = dy.shape
        cin = sum(ctx.c_list)
        if act != ACT_NONE:
            g = torch.empty_like(dy)
            _chk(K, lib.crfp_act_bwd(dy.numel(), act, dy.data_ptr(), out.data_ptr(), g.data_ptr(), st), "act_bwd")
        else:
            g = dy
        need_w, need_b = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        need_x = any(ctx.needs_input_grad[5:])
        dW = db = None
        dxs = [None] * len(srcs)
        if need_x:
            cache = ctx.cache if ctx.cache is not None else {}
            off = 0
            if K.dgrad_as_conv:
                # backward-data of a 3x3/s1/p1 conv IS a 3x3/s1/p1 conv of g with the transposed, 180-degree-rotated
                # kernel: dx[ci] = sum_co conv(g[co], W[co,ci,2-ky,2-kx]).  Running it through the library's tiled
                # forward conv kernel (shared-memory tiles, ~40 % of the FFMA peak) replaced the one-thread-per-
                # (pixel, 4 ci) gather kernel, which sat on the L1 load pipe (ncu: 156 us per 32-channel L1 layer).
                tc_dgrad = getattr(K, "conv3x3_dgrad", None)
                for i, c in enumerate(ctx.c_list):  # one launch per source of the concat: dense per-source gradients
                    if ctx.needs_input_grad[5 + i]:
                        sub = cache.setdefault(("dgrad", i), {})
                        if cache.get("no_tc"):
                            sub["no_tc"] = True
                        elif tc_dgrad is not None:
                            dxs[i] = tc_dgrad(g, weight, off, c, sub)
                            if dxs[i] is not None:
                                off += c
                                continue
                        w2 = cache.get("w2")
                        if w2 is None:
                            w2 = cache["w2"] = weight.detach().transpose(0, 1).flip(2, 3).contiguous()   # (cin, cout, 3, 3)
                        if "zb" not in sub:
                            sub["zb"] = torch.zeros(c, device=dy.device, dtype=torch.float32)
                        dxs[i] = K.conv3x3([g], w2[off:off + c], sub["zb"], ACT_NONE, sub)
                    off += c
            else:
                w_t = cache.get("w_t")
                if w_t is None:
                    w_t = cache["w_t"] = weight.detach().permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()  # [tap][co][ci]
                for i, c in enumerate(ctx.c_list):
                    if ctx.needs_input_grad[5 + i]:
                        dxs[i] = torch.empty(n, h, w, c, device=dy.device, dtype=torch.float32)
                        _chk(K, lib.crfp_conv3x3_bwd_data(n, h, w, c, cout, cin, off, g.data_ptr(), w_t.data_ptr(),
                                                          dxs[i].data_ptr(), st), "conv3x3_bwd_data")
                    off += c
        defer = ctx.cache.get("defer") if ctx.cache is not None else None
        if (need_w or need_b) and defer is not None:
            defer.add(ctx.cache, srcs, g, ctx.c_list, need_w, need_b)   # launched once per layer when the pass ends
        elif need_w or need_b:
            dw = torch.zeros(9, cin, cout, device=dy.device, dtype=torch.float32)
            dbt = torch.zeros(cout, device=dy.device, dtype=torch.float32)
            off = 0
            for i, c in enumerate(ctx.c_list):
                ws, ws_floats = None, 0
                if K.wgrad_two_stage:       # partial sums + a reduce kernel where the library asks for a workspace
                    ws_floats = lib.crfp_conv3x3_bwd_weight_workspace(n, h, w, c, cout)
                    if ws_floats:
                        ws = torch.empty(ws_floats, device=dy.device, dtype=torch.float32)
                _chk(K, lib.crfp_conv3x3_bwd_weight(n, h, w, c, cout, cin, off, srcs[i].data_ptr(), g.data_ptr(),
                                                    dw.data_ptr(), dbt.data_ptr() if i == 0 else None,
                                                    ws.data_ptr() if ws is not None else None, ws_floats, st),
                     "conv3x3_bwd_weight")
                off += c
            if need_w:
                dW = dw.permute(2, 1, 0).reshape(cout, cin, 3, 3)
            if need_b:
                db = dbt
        return (None, None, None, dW, db, *dxs)


def conv3x3(K, weight, bias, srcs, act=ACT_NONE, cache=None):
    """`cache`: optional dict reused for every call with the SAME weight tensor values (packed layouts live there)."""
    return Conv3x3Fn.apply(K, act, cache, weight, bias, *srcs)


# ----------------------------------------------------------------------------------------------- DCNv2
class DCNv2Fn(torch.autograd.Function):
    """dcn_v2.DCNv2.forward(input, offset, mask) in NHWC: offset (n,h,w,dg*18), mask (n,h,w,dg*9); weight OIHW."""

    @staticmethod
    def forward(ctx, K, dg, cache, x, offset, mask, weight, bias):
        x, offset, mask = K.req(x.detach(), "input"), K.req(offset.detach(), "offset"), K.req(mask.detach(), "mask")
        if offset.shape[-1] != dg * 18 or mask.shape[-1] != dg * 9 or x.shape[-1] % dg:
            raise L.CrfpError("DCNv2: offset/mask channel count does not match deformable_groups")
        out = K.dcn_v2(x, offset, mask, weight.detach(), bias.detach(), dg, cache)
        ctx.K, ctx.dg, ctx.cache = K, dg, cache
        ctx.save_for_backward(x, offset, mask, weight)
        return out

    @staticmethod
    def backward(ctx, dout):
        K, dg = ctx.K, ctx.dg
        x, offset, mask, weight = ctx.saved_tensors
        dout = K.req(dout, "grad_output")
        n, h, w, c = x.shape
        cout = weight.shape[0]
        cpg, kk = c // dg, 9 * c
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        wk = ctx.cache.get("wk") if ctx.cache is not None else None
        if wk is None:
            wk = weight.detach().reshape(cout, dg, cpg, 9).permute(1, 3, 2, 0).reshape(kk, cout).contiguous()  # [k][co]
            wk = (wk, wk.t().contiguous())                                                                      # + [co][k]
            if ctx.cache is not None:
                ctx.cache["wk"] = wk
        wk, wk_t = wk
        dx = torch.zeros(n, h, w, c, **f32)
        doff = torch.empty(n, h, w, dg * 18, **f32)
        dmask = torch.empty(n, h, w, dg * 9, **f32)
        dwk = torch.zeros(kk, cout, **f32)
        dbias = torch.zeros(cout, **f32)
        col = torch.empty(n * h * w, kk, **f32)
        ws_floats = K.lib().crfp_dcn_v2_bwd_workspace(n, h, w, c, cout) if K.wgrad_two_stage else 0
        ws = torch.empty(ws_floats, **f32) if ws_floats else None
        d = L.DcnBwdDesc(n=n, h=h, w=w, c=c, cout=cout, dg=dg, x=x.data_ptr(), offset=offset.data_ptr(),
                         mask=mask.data_ptr(), weight=wk.data_ptr(), dout=dout.data_ptr(), dx=dx.data_ptr(),
                         doffset=doff.data_ptr(), dmask=dmask.data_ptr(), dweight=dwk.data_ptr(),
                         dbias=dbias.data_ptr(), col=col.data_ptr(), weight_t=wk_t.data_ptr(),
                         wg_workspace=ws.data_ptr() if ws is not None else None, wg_ws_floats=ws_floats)
        _chk(K, K.lib().crfp_dcn_v2_bwd(C.byref(d), K.stream()), "dcn_v2_bwd")
        dW = dwk.view(dg, 9, cpg, cout).permute(3, 0, 2, 1).reshape(cout, c, 3, 3)
        ng = ctx.needs_input_grad
        return (None, None, None, dx if ng[3] else None, doff if ng[4] else None, dmask if ng[5] else None,
                dW if ng[6] else None, dbias if ng[7] else None)


def dcn_v2(K, x, offset, mask, weight, bias, dg, cache=None):
    return DCNv2Fn.apply(K, dg, cache, x, offset, mask, weight, bias)


# ----------------------------------------------------------------------------------------------- DCN heads activation
class DcnHeadsActFn(torch.autograd.Function):
    """(offset, mask) of DCN_module.forward from the fused heads conv output and the flow (model/CRFP.py:337-347):
    offset = mag * tanh(heads[..., :noff]) + flow.flip(-1) (repeated over the pairs), mask = sigmoid(heads[..., noff:]);
    repeat=True: one pair / one mask per pixel shared by the nk taps.  One kernel forward, one backward."""

    @staticmethod
    def forward(ctx, K, nk, repeat, mag, heads, flow):
        heads, flow = K.req(heads.detach(), "heads"), K.req(flow.detach(), "flow")
        n, h, w, ch = heads.shape
        if ch != (3 if repeat else 3 * nk) or tuple(flow.shape) != (n, h, w, 2):
            raise L.CrfpError("dcn_heads_act: heads / flow shapes do not match")
        offset = torch.empty(n, h, w, 2 * nk, device=heads.device, dtype=torch.float32)
        mask = torch.empty(n, h, w, nk, device=heads.device, dtype=torch.float32)
        _chk(K, K.lib().crfp_dcn_heads_act_fwd(n * h * w, nk, int(repeat), float(mag), heads.data_ptr(), flow.data_ptr(),
                                               offset.data_ptr(), mask.data_ptr(), K.stream()), "dcn_heads_act_fwd")
        ctx.K, ctx.nk, ctx.repeat, ctx.mag = K, nk, repeat, mag
        ctx.save_for_backward(heads)
        return offset, mask

    @staticmethod
    def backward(ctx, doffset, dmask):
        K = ctx.K
        (heads,) = ctx.saved_tensors
        n, h, w, _ = heads.shape
        if doffset is None:
            doffset = torch.zeros(n, h, w, 2 * ctx.nk, device=heads.device, dtype=torch.float32)
        if dmask is None:
            dmask = torch.zeros(n, h, w, ctx.nk, device=heads.device, dtype=torch.float32)
        doffset, dmask = K.req(doffset, "grad_offset"), K.req(dmask, "grad_mask")
        dheads = torch.empty_like(heads)
        dflow = torch.empty(n, h, w, 2, device=heads.device, dtype=torch.float32)
        _chk(K, K.lib().crfp_dcn_heads_act_bwd(n * h * w, ctx.nk, int(ctx.repeat), float(ctx.mag), heads.data_ptr(),
                                               doffset.data_ptr(), dmask.data_ptr(), dheads.data_ptr(), dflow.data_ptr(),
                                               K.stream()), "dcn_heads_act_bwd")
        return None, None, None, None, dheads, dflow


def dcn_heads_act(K, heads, flow, nk, repeat, mag):
    return DcnHeadsActFn.apply(K, nk, repeat, mag, heads, flow)


# ----------------------------------------------------------------------------------------------- fovea blend
class FoveaBlendFn(torch.autograd.Function):
    """lrelu(mask * f + (1 - mask) * s, 0.1) (model/CRFP.py:1672-1675); mask (n,H,W,1) carries no gradient."""

    @staticmethod
    def forward(ctx, K, f, s, mask):
        f, s, mask = K.req(f.detach(), "f"), K.req(s.detach(), "s"), K.req(mask.detach(), "mask")
        n, h, w, c = f.shape
        if tuple(s.shape) != (n, h, w, c) or mask.numel() != n * h * w or c % 4:
            raise L.CrfpError("fovea_blend: shapes do not match")
        out = torch.empty_like(f)
        _chk(K, K.lib().crfp_fovea_blend_fwd(n * h * w, c, f.data_ptr(), s.data_ptr(), mask.data_ptr(), out.data_ptr(), K.stream()),
             "fovea_blend_fwd")
        ctx.K = K
        ctx.save_for_backward(out, mask)
        return out

    @staticmethod
    def backward(ctx, dout):
        K = ctx.K
        out, mask = ctx.saved_tensors
        dout = K.req(dout, "grad_output")
        n, h, w, c = out.shape
        df, ds = torch.empty_like(out), torch.empty_like(out)
        _chk(K, K.lib().crfp_fovea_blend_bwd(n * h * w, c, dout.data_ptr(), out.data_ptr(), mask.data_ptr(), df.data_ptr(),
                                             ds.data_ptr(), K.stream()), "fovea_blend_bwd")
        return None, df, ds, None


def fovea_blend(K, f, s, mask):
    return FoveaBlendFn.apply(K, f, s, mask)


# ----------------------------------------------------------------------------------------------- flow_warp
class FlowWarpFn(torch.autograd.Function):
    """flow_warp(x, flow) (model/CRFP.py:90-130), x (n,h,w,c) with c % 4 == 0, flow (n,h,w,2) = (dx, dy)."""

    @staticmethod
    def forward(ctx, K, x, flow):
        x, flow = K.req(x.detach(), "x"), K.req(flow.detach(), "flow")
        if tuple(flow.shape) != (*x.shape[:3], 2):
            raise ValueError(f"The spatial sizes of input ({tuple(x.shape[1:3])}) and flow ({tuple(flow.shape[1:3])}) "
                             "are not the same.")
        ctx.K = K
        ctx.save_for_backward(x, flow)
        return K.flow_warp(x, flow)

    @staticmethod
    def backward(ctx, dy):
        K = ctx.K
        x, flow = ctx.saved_tensors
        dy = K.req(dy, "grad_output")
        n, h, w, c = x.shape
        need_x, need_f = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_x or need_f):
            return None, None, None
        dx = torch.zeros_like(x) if need_x else None
        df = torch.empty_like(flow) if need_f else None
        _chk(K, K.lib().crfp_flow_warp_bwd(n, h, w, c, x.data_ptr(), flow.data_ptr(), dy.data_ptr(),
                                           dx.data_ptr() if need_x else None, df.data_ptr() if need_f else None,
                                           K.stream()), "flow_warp_bwd")
        return None, dx, df


def flow_warp(K, x, flow):
    return FlowWarpFn.apply(K, x, flow)


# ----------------------------------------------------------------------------------------------- resize / pool
class ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, K, x, hout, wout, rh, rw, mul):
        x = K.req(x.detach(), "x")
        ctx.K, ctx.args, ctx.in_shape = K, (hout, wout, rh, rw, mul), tuple(x.shape)
        return K.resize(x, hout, wout, rh, rw, mul)

    @staticmethod
    def backward(ctx, dy):
        K = ctx.K
        if not ctx.needs_input_grad[1]:
            return (None,) * 7
        dy = K.req(dy, "grad_output")
        n, hin, win, c = ctx.in_shape
        hout, wout, rh, rw, mul = ctx.args
        dx = torch.zeros(n, hin, win, c, device=dy.device, dtype=torch.float32)
        _chk(K, K.lib().crfp_resize_bilinear_bwd(n, hin, win, c, hout, wout, rh, rw, mul, dy.data_ptr(), dx.data_ptr(),
                                                 K.stream()), "resize_bilinear_bwd")
        return (None, dx, None, None, None, None, None)


def up_bilinear(K, x, scale, mul=1.0):
    """nn.Upsample(scale_factor=scale, mode='bilinear', align_corners=False)(x) * mul."""
    n, h, w, c = x.shape
    return ResizeFn.apply(K, x, int(h * scale), int(w * scale), 1.0 / scale, 1.0 / scale, float(mul))


def resize_to(K, x, hout, wout):
    """F.interpolate(x, size=(hout, wout), mode='bilinear', align_corners=False)."""
    n, h, w, c = x.shape
    if (h, w) == (hout, wout):
        return x
    return ResizeFn.apply(K, x, hout, wout, h / hout, w / wout, 1.0)


class AvgPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, K, x):
        x = K.req(x.detach(), "x")
        ctx.K, ctx.in_shape = K, tuple(x.shape)
        return K.avgpool2(x)

    @staticmethod
    def backward(ctx, dy):
        K = ctx.K
        dy = K.req(dy, "grad_output")
        n, hin, win, c = ctx.in_shape
        dx = torch.empty(n, hin, win, c, device=dy.device, dtype=torch.float32)
        _chk(K, K.lib().crfp_avgpool2_bwd(n, hin, win, c, dy.data_ptr(), dx.data_ptr(), K.stream()), "avgpool2_bwd")
        return None, dx


def avgpool2(K, x):
    return AvgPool2Fn.apply(K, x)


# ----------------------------------------------------------------------------------------------- loss
class CharbonnierFn(torch.autograd.Function):
    """CharbonnierLoss(loss_weight, reduction='mean', eps)(pred, target) (loss/loss.py:116-176): one kernel computes the
    loss sum and d loss / d pred."""

    @staticmethod
    def forward(ctx, K, pred, target, eps, loss_weight):
        p, t = K.req(pred.detach(), "pred"), K.req(target.detach(), "target")
        if p.shape != t.shape:
            raise ValueError("pred and target must have the same shape")
        count = p.numel()
        loss_sum = torch.zeros(1, device=p.device, dtype=torch.float32)
        dpred = torch.empty_like(p)
        _chk(K, K.lib().crfp_charbonnier_fwd_bwd(count, p.data_ptr(), t.data_ptr(), eps, loss_weight / count,
                                                 loss_sum.data_ptr(), dpred.data_ptr(), K.stream()), "charbonnier")
        ctx.save_for_backward(dpred)
        return (loss_sum * (loss_weight / count)).reshape(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return None, dpred * g, None, None, None


def charbonnier_loss(K, pred, target, eps=1e-12, loss_weight=1.0):
    return CharbonnierFn.apply(K, pred, target, float(eps), float(loss_weight))
