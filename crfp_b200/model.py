"""nn.Module shells with the reference's API over libcrfp_b200.

  CRFP_DSV          /root/reference/model/CRFP.py:1387-1706   forward(lrs, fvs, mks) -> (n,t,3,8h,8w)
  MRCF_simple_v18   /root/reference/model/CRFP_test.py:2114-2478   stateful forward(lrs, fvs, mks, fgs) +
                    clear_states(); identical state_dict (118 tensors)

The shells own the parameters (same names, shapes and dtypes as the reference so checkpoints load with
strict=True in both directions) and device buffers; every arithmetic op of the forward runs in the CUDA
library through `crfp_dsv_prepare` / `crfp_dsv_frame`.  No CPU fallback: a missing library or a non-CUDA
input raises.
"""
from __future__ import annotations

import collections
import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib as L
from .packing import pack_align_heads, pack_layer, pack_layer_tc3
from .spec import crfp_param_shapes
from .synthetic import fovea_rect


class _Holder(nn.Module):
    """Plain container so that parameter paths equal the reference's (`dcn_0.dcn_block.0.weight`, ...)."""


class _DCNParams(nn.Module):
    """Parameter holder for `dcn_k.dcn` (dcn_v2.DCNv2: weight (Co,Ci,3,3), bias (Co))."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(cout, cin, 3, 3))
        self.bias = nn.Parameter(torch.zeros(cout))


def _build_tree(root: nn.Module, shapes):
    for key, shape in shapes.items():
        if not key.endswith(".weight"):
            continue
        path = key[: -len(".weight")].split(".")
        node = root
        for comp in path[:-1]:
            if comp not in node._modules:
                node.add_module(comp, _Holder())
            node = node._modules[comp]
        cout, cin = shape[0], shape[1]
        leaf = _DCNParams(cin, cout) if path[-1] == "dcn" else nn.Conv2d(cin, cout, 3, 1, 1, bias=True)
        node.add_module(path[-1], leaf)


def _kaiming_fan_in_(conv: nn.Conv2d, scale: float):
    nn.init.kaiming_normal_(conv.weight, a=0, mode="fan_in", nonlinearity="relu")
    conv.weight.data *= scale
    nn.init.constant_(conv.bias, 0)


_DEVICE_OK = set()


def _check_device(dev):
    """sm_100 check, once per device and process (inside `with torch.cuda.device(dev)`)."""
    key = str(dev)
    if key not in _DEVICE_OK:
        L.check(L.lib().crfp_check_device(), "device check (sm_100 required)")
        _DEVICE_OK.add(key)


class _CRFPBase(nn.Module):
    VARIANT = "dsv"

    def __init__(self, device, mid_channels=16, y_only=False, hr_dcn=True, offset_prop=True, spynet_pretrained=None,
                 precision="tc"):
        super().__init__()
        if precision not in ("fp32", "tc", "half"):
            raise L.CrfpError("precision must be 'tc' (fp32 storage, tcgen05 3 x bf16 split contractions, fp32-grade), "
                              "'fp32' (all-SIMT fp32 FFMA) or 'half' (reduced-precision tier: one fp16 activation product "
                              "against fp16 hi / lo split weights, fp32 storage and accumulation; <= 5e-3)")
        self.precision = precision
        if mid_channels != 32:
            raise L.CrfpError("crfp_b200 implements the shipped configuration mid_channels=32 (main.py:34)")
        if y_only or not hr_dcn or not offset_prop:
            raise L.CrfpError("crfp_b200 implements y_only=False, hr_dcn=True, offset_prop=True "
                              "(the defaults every shipped script uses, option.py:58-63)")
        self.device = device
        self.mid_channels = mid_channels
        self.last_channels = mid_channels // 8
        self.dg_num, self.dk, self.max_residue_magnitude = 8, 3, 10
        self.y_only, self.hr_dcn, self.offset_prop, self.split_ratio = y_only, hr_dcn, offset_prop, 3
        _build_tree(self, crfp_param_shapes(self.VARIANT, mid_channels, y_only))
        self._init_like_reference()
        if spynet_pretrained is not None:
            self.spynet.load_state_dict(torch.load(spynet_pretrained, map_location="cpu"))
        self._packed = None       # (version key, device, blob tensors, DsvWeights)
        self._ws = {}             # workspace cache keyed by shape
        self.skip_outside_fovea = True
        # CUDA graphs: a clip forward is ~60 launches per frame; the second call on the same input buffers captures
        # the whole clip into one graph and later calls replay it (one launch per clip: immune to host jitter)
        self.use_graphs = os.environ.get("CRFP_NO_GRAPHS") is None
        self.graph_frames = 20                     # frames per captured graph
        # A graph owns its output buffer.  Default: a replay returns a COPY of it (one device-to-device copy, < 1 % of
        # a clip), so tensors returned by earlier calls stay valid like the reference's fresh tensors do.
        # alias_output = True (or an explicit `out=` buffer) returns the graph-owned tensor itself, overwritten by the
        # next replay on the same inputs — the cudagraph-style contract, for callers that consume each result at once.
        self.alias_output = False
        # second stream for the off-chain work of a frame (crfp_dsv_frame_desc.aux_stream).  Opt-in (CRFP_AUX=1): measured
        # on B200 at R-lit it is 1 % SLOWER than one stream (351.8 vs 354.9 fps) — the tensor-core kernels hold every SM's
        # registers, so forked kernels only run in their launch gaps and lengthen the critical path's tail
        self.use_aux_stream = os.environ.get("CRFP_AUX", "0") == "1"
        self._aux = None
        self._graphs = collections.OrderedDict()   # key -> dict(graphs, out, launches)
        self._seen_key = None

    # ---- init policy of the reference (statistically identical, not RNG-stream identical)
    def _init_like_reference(self):
        for name, m in self.named_modules():
            if isinstance(m, nn.Conv2d):
                if name.endswith("upsample_conv") or name.endswith("downsample_conv"):
                    _kaiming_fan_in_(m, 1.0)           # PixelShufflePack.init_weights, CRFP.py:182-185
                elif ".main.2.0.conv" in name:
                    _kaiming_fan_in_(m, 0.1)           # ResidualBlockNoBN.init_weights, CRFP.py:459-470
                elif name.endswith("dcn_offset") or name.endswith("dcn_mask"):
                    nn.init.zeros_(m.weight)           # DCN_module.init_dcn, CRFP.py:354-359
                    nn.init.zeros_(m.bias)
            elif isinstance(m, _DCNParams):            # conv_identify, CRFP.py:361-370
                with torch.no_grad():
                    m.weight.zero_()
                    m.bias.zero_()
                    for p in range(min(m.weight.shape[0], m.weight.shape[1])):
                        m.weight[p, p, 1, 1] = 1.0

    def init_weights(self, pretrained=None, strict=True):
        """Same contract as the reference (CRFP.py:1688-1706)."""
        if isinstance(pretrained, str):
            saved = {k: v for k, v in torch.load(pretrained, map_location=self.device).items()}
            sd = self.state_dict()
            sd.update(saved)
            self.load_state_dict(sd, strict=strict)
        elif pretrained is not None:
            raise TypeError(f'"pretrained" must be a str or None. But received {type(pretrained)}.')

    # ---- packed weights (refreshed whenever a parameter changes or moves)
    def _weights(self, device):
        params = list(self.parameters())
        key = (str(device), self.precision, tuple((p.data_ptr(), p._version) for p in params))
        if self._packed is not None and self._packed[0] == key:
            return self._packed[2]
        sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in self.state_dict().items()}
        table = L.layer_table(self.VARIANT)
        W = L.DsvWeights()
        W.mid_channels, W.nlayers = self.mid_channels, len(table)
        W.precision = {"tc": L.PREC_TC3, "half": L.PREC_HALF, "fp32": L.PREC_FP32}[self.precision]
        wdt = torch.float16 if self.precision == "half" else torch.bfloat16
        W.variant = L.VARIANTS[self.VARIANT]
        keep = []
        for i, info in enumerate(table):
            w, b = pack_layer(info, sd)
            keep.append((w, b))
            W.layer[i].w, W.layer[i].b = w.data_ptr(), b.data_ptr()
            if info["tc"] and self.precision in ("tc", "half"):
                # the flow network keeps the 3 x bf16 split in every precision (crfp_dsv_frame routes it accordingly)
                hi, lo, bt, wx = pack_layer_tc3(info, sd, torch.bfloat16 if info["key"].startswith("spynet.") else wdt)
                keep.append((hi, lo, bt, wx))
                W.layer_tc[i].w_hi, W.layer_tc[i].w_lo, W.layer_tc[i].b = hi.data_ptr(), lo.data_ptr(), bt.data_ptr()
                if wx is not None:
                    W.layer_tc[i].w_extra = wx.data_ptr()
                if info["kind"] == 2 and info["cout"] == 216:     # L1 offset / mask heads: the fused align kernel's packing
                    wf, bf = pack_align_heads(sd[info["key"] + ".weight"], sd[info["key"] + ".bias"],
                                              sd[info["key2"] + ".weight"], sd[info["key2"] + ".bias"], wdt)
                    keep.append((wf, bf))
                    W.layer_tc[i].w_fused, W.layer_tc[i].b_fused = wf.data_ptr(), bf.data_ptr()
        self._packed = (key, keep, W)
        self._graphs.clear()      # captured graphs point at the previous packed weights
        for sb in getattr(self, "_sbuf", {}).values():
            sb["graphs"].clear()
            sb["seen"].clear()
        return W

    def _clip_buffers(self, n, t, h, w, device):
        key = (n, t, h, w, str(device))
        if key not in self._ws:
            self._ws.clear()
            self._graphs.clear()  # captured graphs point at the previous clip buffers
            if hasattr(self, "_sbuf"):
                self._sbuf.clear()
            skey = (n, h, w, str(device))
            if getattr(self, "_state_bufs", None) is None or self._state_bufs[0] != skey:
                # recurrent state, O(1) in t: its own cache so that a streaming module may change t between calls
                self._state_bufs = (skey, torch.zeros(n * 64 * h * w * 4, device=device, dtype=torch.float32),
                                    torch.zeros(n * 4 * h * w * 24, device=device, dtype=torch.float32))
            lib = L.lib()
            shp = L.DsvShape(n=n, t=t, h=h, w=w, mid_channels=self.mid_channels)
            pw, fw = lib.crfp_dsv_prepare_workspace(C.byref(shp)), lib.crfp_dsv_frame_workspace(C.byref(shp))
            if pw == 0 or fw == 0:
                raise L.CrfpError(f"unsupported clip shape n={n} t={t} h={h} w={w}")
            f32 = dict(device=device, dtype=torch.float32)
            self._ws[key] = dict(
                shape=shp,
                ws=torch.empty(max(pw, fw), device=device, dtype=torch.uint8),
                lr4=torch.empty(n * t * h * w * 4, **f32),
                x_lr=torch.empty(n * t * h * w * self.mid_channels, **f32),
                flows=torch.zeros(n * t * h * w * 2, **f32),
                state_hr=self._state_bufs[1],
                state_l1=self._state_bufs[2],
            )
        return self._ws[key]

    @staticmethod
    def _check_inputs(lrs, fvs, mks):
        for name, t_ in (("lrs", lrs), ("fvs", fvs), ("mks", mks)):
            if not (isinstance(t_, torch.Tensor) and t_.is_cuda):
                raise L.CrfpError(f"{name} must be a CUDA tensor: crfp_b200 has no CPU fallback")
        n, t, c, h, w = lrs.shape
        if c != 3:
            raise ValueError(f"lrs must have 3 channels, got {c}")
        if tuple(fvs.shape) != (n, t, 3, 8 * h, 8 * w):
            raise ValueError(f"fvs must be {(n, t, 3, 8 * h, 8 * w)}, got {tuple(fvs.shape)}")
        if tuple(mks.shape) != (n, t, 1, 8 * h, 8 * w):
            raise ValueError(f"mks must be {(n, t, 1, 8 * h, 8 * w)}, got {tuple(mks.shape)}")
        lrs = lrs.to(torch.float32).contiguous()
        fvs = fvs.to(torch.float32).contiguous()
        mks = (mks != 0).to(torch.uint8).contiguous() if mks.dtype != torch.bool else mks.contiguous().view(torch.uint8)
        return lrs, fvs, mks

    @staticmethod
    def _same_storage(pairs):
        """True when every converted tensor still IS the caller's storage (fp32 / bool, contiguous): only then may a
        CUDA graph keyed on the input addresses be captured — a temporary made by a dtype / layout conversion lives at an
        allocator-chosen address that a later replay must not read."""
        return all(a.data_ptr() == b.data_ptr() for a, b in pairs)

    def _run_frames(self, buf, W, lrs, fvs, mks, fgs, out, first_flags, out_host=None, frames=None):
        """Frame loop.  `out_host` (pinned CPU tensor shaped like `out`): every finished frame is copied to the host on a
        side stream while the next frames are computed (device->host traffic overlaps the recurrence)."""
        lib = L.lib()
        copy_stream = None
        u8 = None
        if out_host is not None:
            if tuple(out_host.shape) != tuple(out.shape) or out_host.dtype not in (out.dtype, torch.uint8) or not out_host.is_pinned():
                raise ValueError("out_host must be a pinned CPU tensor with the output's shape, fp32 or uint8")
            if out_host.dtype == torch.uint8:
                # frames quantised on the device exactly as the reference saves them ((sr * 255).clip(0, 255).round(),
                # trainer.py:446-474): 4x fewer bytes over PCIe; two staging frames so the copy of frame i overlaps the
                # quantisation of frame i + 1
                if getattr(self, "_u8", None) is None or self._u8[0].shape != out[:, 0].shape or self._u8[0].device != out.device:
                    self._u8 = [torch.empty(out[:, 0].shape, device=out.device, dtype=torch.uint8) for _ in range(2)]
                self._u8_ev = [None, None]   # per call: the previous call ended with wait_stream(copy_stream)
                u8 = self._u8
            if getattr(self, "_copy_stream", None) is None or self._copy_stream.device != out.device:
                self._copy_stream = torch.cuda.Stream(device=out.device)
            copy_stream = self._copy_stream
        n, t, _, h, w = lrs.shape
        hw, HW = h * w, 64 * h * w
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        d = L.DsvFrameDesc()
        d.shape = L.DsvShape(n=n, t=1, h=h, w=w, mid_channels=self.mid_channels)
        d.skip_outside_fovea = int(self.skip_outside_fovea)
        d.lr4_clip_stride, d.x_lr_clip_stride, d.flow_clip_stride = t * hw * 4, t * hw * self.mid_channels, t * hw * 2
        d.fvs_clip_stride, d.mks_clip_stride, d.out_clip_stride = t * 3 * HW, t * HW, t * 3 * HW
        d.state_hr, d.state_l1 = buf["state_hr"].data_ptr(), buf["state_l1"].data_ptr()
        if self.use_aux_stream:
            if self._aux is None or self._aux[0] != out.device:
                evs = [torch.cuda.Event() for _ in range(3)]
                aux = torch.cuda.Stream(device=out.device)
                for e in evs:
                    e.record(aux)          # events are created lazily: force the handles into existence
                self._aux = (out.device, aux, evs)
            d.aux_stream = self._aux[1].cuda_stream
            for k, e in enumerate(self._aux[2]):
                d.aux_events[k] = e.cuda_event
        ws = buf["ws"]
        for i in (range(t) if frames is None else frames):
            d.first = int(first_flags[i])
            d.lr4 = buf["lr4"].data_ptr() + i * hw * 4 * 4
            d.x_lr = buf["x_lr"].data_ptr() + i * hw * self.mid_channels * 4
            d.flow = buf["flows"].data_ptr() + i * hw * 2 * 4
            d.fvs = fvs.data_ptr() + i * 3 * HW * 4
            d.mks = mks.data_ptr() + i * HW
            if fgs is not None:
                d.fg, d.fg_clip_stride = fgs.data_ptr() + i * HW * 4, t * HW
            d.out = out.data_ptr() + i * 3 * HW * 4
            L.check(lib.crfp_dsv_frame(C.byref(d), C.byref(W), ws.data_ptr(), ws.numel(), st), f"dsv_frame[{i}]")
            if copy_stream is not None:
                src = out[:, i]
                if u8 is not None:
                    k = i & 1
                    if self._u8_ev[k] is not None:               # the copy that last read this staging frame is done
                        torch.cuda.current_stream().wait_event(self._u8_ev[k])
                    if n == 1:
                        L.check(lib.crfp_quantize_u8(src.data_ptr(), u8[k].data_ptr(), src.numel(), st), "quantize_u8")
                    else:                                        # strided over clips: one launch per clip
                        for b in range(n):
                            L.check(lib.crfp_quantize_u8(out[b, i].data_ptr(), u8[k][b].data_ptr(), out[b, i].numel(), st), "quantize_u8")
                    src = u8[k]
                ev = torch.cuda.Event()
                ev.record()
                copy_stream.wait_event(ev)
                with torch.cuda.stream(copy_stream):
                    out_host[:, i].copy_(src, non_blocking=True)
                    if u8 is not None:
                        self._u8_ev[i & 1] = torch.cuda.Event()
                        self._u8_ev[i & 1].record(copy_stream)
        if copy_stream is not None:
            torch.cuda.current_stream().wait_stream(copy_stream)


class CRFP_DSV(_CRFPBase):
    """Drop-in for `model.CRFP.CRFP_DSV` (the model main.py:34 builds)."""

    def forward(self, lrs, fvs, mks, out_host=None, out=None):
        """Reference signature `forward(lrs, fvs, mks)`; the optional `out_host` (pinned CPU tensor) additionally
        streams every finished frame to the host while the recurrence continues; the optional `out` (CUDA fp32 tensor
        (n,t,3,8h,8w)) receives the result in place of a freshly allocated tensor."""
        if self._wants_grad():
            # training: node-by-node forward over the autograd kernel pairs (crfp_b200/training.py); the fused
            # whole-frame inference kernels keep no intermediates to differentiate through
            if self.VARIANT != "dsv":
                raise NotImplementedError("crfp_b200: the training forward is implemented for CRFP_DSV only")
            if out_host is not None:
                raise ValueError("out_host streaming is an inference feature: the training output carries grad")
            from .training import forward_train
            lrs, fvs, mks = self._check_inputs(lrs, fvs, mks)
            return forward_train(self, lrs, fvs, mks)
        with torch.no_grad():
            user = (lrs, fvs, mks)
            lrs, fvs, mks = self._check_inputs(lrs, fvs, mks)
            key = ("clip", lrs.data_ptr(), fvs.data_ptr(), mks.data_ptr())
            direct = self._same_storage(zip(user, (lrs, fvs, mks)))
            return self._forward_clip(key, lrs, fvs, mks, out_host, None, out=out, graphable=direct)

    def _wants_grad(self):
        return torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())

    def _check_mode(self):
        if self._wants_grad():
            raise NotImplementedError("crfp_b200: this entry point is inference-only; call under torch.no_grad() / "
                                      "model.eval(), or use forward(lrs, fvs, mks) for training")

    def _forward_clip(self, key, lrs, fvs, mks, out_host, pre, out=None, graphable=True):
        """prepare + frame loop, eagerly or as a replay of the captured whole-clip graph.  `pre(stream)` = extra work
        at the head of the clip (the fovea paste of forward_patch)."""
        n, t, _, h, w = lrs.shape
        dev = lrs.device
        if out is not None and (tuple(out.shape) != (n, t, 3, 8 * h, 8 * w) or out.dtype != torch.float32 or
                                out.device != dev or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous CUDA fp32 tensor of shape {(n, t, 3, 8 * h, 8 * w)} on {dev}")
        with torch.cuda.device(dev):
            _check_device(dev)
            W = self._weights(dev)
            buf = self._clip_buffers(n, t, h, w, dev)

            def run(out, part=None):
                """part None: the whole clip; part k: frames [k*G, (k+1)*G) (+ the clip-level work when k == 0)."""
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                if part is None or part == 0:
                    if pre is not None:
                        pre(st)
                    shp = L.DsvShape(n=n, t=t, h=h, w=w, mid_channels=self.mid_channels)
                    L.check(L.lib().crfp_dsv_prepare(C.byref(shp), C.byref(W), lrs.data_ptr(), None, buf["lr4"].data_ptr(),
                                                     buf["x_lr"].data_ptr(), buf["flows"].data_ptr(), buf["ws"].data_ptr(),
                                                     buf["ws"].numel(), st), "dsv_prepare")
                frames = None if part is None else range(part * self.graph_frames, min(t, (part + 1) * self.graph_frames))
                self._run_frames(buf, W, lrs, fvs, mks, None, out, [i == 0 for i in range(t)], out_host, frames)

            key = key + (n, t, h, w, str(dev), 0 if out_host is None else out_host.data_ptr(), bool(self.skip_outside_fovea),
                         0 if out is None else out.data_ptr())
            if self.use_graphs and graphable and not torch.cuda.is_current_stream_capturing():
                entry = self._graphs.get(key)
                if entry is None and self._seen_key == key:
                    entry = self._capture(key, run, (n, t, 3, 8 * h, 8 * w), dev, (t + self.graph_frames - 1) // self.graph_frames,
                                          out)
                self._seen_key = key
                if entry is not None:
                    self._graphs.move_to_end(key)
                    for g in entry["graphs"]:
                        g.replay()
                    L.lib().crfp_launch_count_add(entry["launches"])
                    if out is not None or self.alias_output:
                        return entry["out"]
                    return entry["out"].clone()
            if out is None:
                out = torch.empty(n, t, 3, 8 * h, 8 * w, device=dev, dtype=torch.float32)
            run(out)
        return out

    def _capture(self, key, run, out_shape, dev, parts, out=None):
        """Capture one whole-clip forward (prepare + every frame, incl. the streaming device->host copies) into CUDA
        graphs of `graph_frames` frames each: the next graph is launched while the previous one executes, so only
        the first graph's launch latency is exposed.  The entry owns its output tensor (or writes the caller's `out`);
        see `alias_output` for what a replay returns.  Returns None (and switches graphs off) if the capture fails."""
        lib = L.lib()
        if out is None:
            out = torch.empty(*out_shape, device=dev, dtype=torch.float32)
        graphs = []
        try:
            torch.cuda.synchronize(dev)
            before = lib.crfp_launch_count()
            for k in range(parts):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    run(out, k)
                graphs.append(g)
            launches = lib.crfp_launch_count() - before
            lib.crfp_launch_count_add(-launches)        # captured, not launched
        except Exception as e:  # noqa: BLE001
            import warnings
            warnings.warn(f"crfp_b200: CUDA graph capture failed ({e}); continuing without graphs")
            self.use_graphs = False
            return None
        while len(self._graphs) >= 3:
            self._graphs.popitem(last=False)
        entry = dict(graphs=graphs, out=out, launches=launches)
        self._graphs[key] = entry
        return entry

    def forward_patch(self, lrs, fovea_patch, coords, out_host=None, out=None):
        """Convenience entry named by BASELINE.json: `fovea_patch` (n,t,3,FV,FV) pasted at integer top-left
        `coords` (n,t,2) = [y, x] exactly as the data loader does (dataset/reds.py:196-201) — on the device, into
        persistent full-frame fvs / mks buffers (only the previous call's rectangles are cleared)."""
        self._check_mode()
        with torch.no_grad():
            if not (isinstance(lrs, torch.Tensor) and lrs.is_cuda and isinstance(fovea_patch, torch.Tensor) and fovea_patch.is_cuda):
                raise L.CrfpError("lrs and fovea_patch must be CUDA tensors: crfp_b200 has no CPU fallback")
            n, t, c, h, w = lrs.shape
            fv = fovea_patch.shape[-1]
            H, Wd = 8 * h, 8 * w
            if c != 3 or tuple(fovea_patch.shape) != (n, t, 3, fv, fv) or fv > min(H, Wd):
                raise ValueError(f"lrs must be (n,t,3,h,w) and fovea_patch (n,t,3,fv,fv) with fv <= {min(H, Wd)}")
            if tuple(coords.shape) != (n, t, 2):
                raise ValueError(f"coords must be {(n, t, 2)}")
            cc = coords.detach().to("cpu", torch.int64)
            if int(cc.min()) < 0 or int(cc[..., 0].max()) > H - fv or int(cc[..., 1].max()) > Wd - fv:
                raise ValueError("fovea patch outside the frame")
            dev = lrs.device
            user = (lrs, fovea_patch)
            lrs = lrs.to(torch.float32).contiguous()
            patch = fovea_patch.to(torch.float32).contiguous()
            direct = self._same_storage(zip(user, (lrs, patch)))
            pk = (n, t, h, w, fv, str(dev))
            if getattr(self, "_patch_buf", None) is None or self._patch_buf[0] != pk:
                self._graphs.clear()
                self._patch_buf = (pk, dict(fvs=torch.zeros(n, t, 3, H, Wd, device=dev),
                                            mks=torch.zeros(n, t, 1, H, Wd, device=dev, dtype=torch.uint8),
                                            coords=torch.zeros(n, t, 2, device=dev, dtype=torch.int32),
                                            prev=torch.zeros(n, t, 2, device=dev, dtype=torch.int32)))
            pb = self._patch_buf[1]
            # outside the graph: host -> device copy of n*t*8 bytes through a small ring of pinned staging buffers, so
            # the call never blocks on the stream (the host may run several steps ahead of the device)
            if "stage" not in pb:
                pb["stage"] = [torch.empty(n, t, 2, dtype=torch.int32).pin_memory() for _ in range(4)]
                pb["stage_ev"] = [None] * 4
                pb["stage_i"] = 0
            k = pb["stage_i"]
            pb["stage_i"] = (k + 1) % 4
            if pb["stage_ev"][k] is not None:
                pb["stage_ev"][k].synchronize()
            pb["stage"][k].copy_(cc)
            pb["coords"].copy_(pb["stage"][k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pb["stage_ev"][k] = ev
            lib = L.lib()

            def pre(st):
                args = (n * t, fv, H, Wd, pb["fvs"].data_ptr(), pb["mks"].data_ptr())
                L.check(lib.crfp_fovea_paste(None, pb["prev"].data_ptr(), *args, 1, st), "fovea clear")
                L.check(lib.crfp_fovea_paste(patch.data_ptr(), pb["coords"].data_ptr(), *args, 0, st), "fovea paste")
                pb["prev"].copy_(pb["coords"])

            key = ("patch", lrs.data_ptr(), patch.data_ptr(), fv)
            return self._forward_clip(key, lrs, pb["fvs"], pb["mks"], out_host, pre, out=out, graphable=direct)


class CRFP(CRFP_DSV):
    """Drop-in for `model.CRFP.CRFP` ("v15", CRFP.py:1101-1385): 3-way concat, HR state warped before down-sampling."""
    VARIANT = "v15"


class CRFP_simple(CRFP_DSV):
    """Drop-in for `model.CRFP.CRFP_simple` ("v13", CRFP.py:816-1099)."""
    VARIANT = "v13"


class MRCF_simple_v18(_CRFPBase):
    """Drop-in for the streaming `model.CRFP_test.MRCF_simple_v18`: one (or a few) frames per call, recurrent
    state kept on the module, `clear_states()` between clips.

    Latency path: the frame protocol of test_video.py:316-374 is one call per frame at batch 1, i.e. ~60 kernel
    launches of host work per call.  From the third call of a kind (same n, t, h, w; first-of-stream or not) the call is
    a CUDA-graph replay: the inputs are copied into persistent staging buffers (4 device copies), ONE graph launch runs
    FNet + encoder + the frame step(s) + the previous-LR-frame update, and the result is returned as a copy of the
    graph-owned output (`alias_output = True` returns the graph-owned tensor itself)."""

    def __init__(self, device, mid_channels=16, y_only=False, hr_dcn=True, offset_prop=True, split_ratio=3,
                 spynet_pretrained=None, precision="tc"):
        super().__init__(device, mid_channels, y_only, hr_dcn, offset_prop, spynet_pretrained, precision)
        if split_ratio != 3:
            raise L.CrfpError("split_ratio=3 only")
        self.pre_lr = None
        self._has_state = False
        self._state_key = None      # (n, h, w, device) the recurrent state was built for
        self._sbuf = {}             # (n, t, h, w, device) -> staging buffers + graphs

    def clear_states(self):
        self.pre_lr = None
        self._has_state = False
        self._state_key = None

    def _stream_buffers(self, n, t, h, w, dev):
        key = (n, t, h, w, str(dev))
        if key not in self._sbuf:
            self._sbuf.clear()
            H, Wd = 8 * h, 8 * w
            f32 = dict(device=dev, dtype=torch.float32)
            self._sbuf[key] = dict(lrs=torch.empty(n, t, 3, h, w, **f32), fvs=torch.empty(n, t, 3, H, Wd, **f32),
                                   mks=torch.empty(n, t, 1, H, Wd, device=dev, dtype=torch.uint8),
                                   fgs=torch.empty(n, t, 1, H, Wd, **f32), out=torch.empty(n, t, 3, H, Wd, **f32),
                                   graphs={}, seen={})
        return self._sbuf[key]

    @torch.no_grad()
    def forward(self, lrs, fvs, mks, fgs):
        lrs, fvs, mks = self._check_inputs(lrs, fvs, mks)
        n, t, _, h, w = lrs.shape
        dev = lrs.device
        fgs = fgs.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(fgs.shape) != (n, t, 1, 8 * h, 8 * w):
            raise ValueError(f"fgs must be {(n, t, 1, 8 * h, 8 * w)}")
        skey = (n, h, w, str(dev))
        if self._has_state and self._state_key != skey:
            # the reference would fail with a torch shape error when the warped state meets the new frame size
            raise ValueError(f"the recurrent state belongs to frames of (n, h, w, device) = {self._state_key}, got {skey}: "
                             "call clear_states() before changing the stream's shape")
        if self.pre_lr is not None and tuple(self.pre_lr.shape) != (n, 3, h, w):
            raise ValueError(f"pre_lr has shape {tuple(self.pre_lr.shape)}, expected {(n, 3, h, w)}: call clear_states()")
        with torch.cuda.device(dev):
            _check_device(dev)
            W = self._weights(dev)
            buf = self._clip_buffers(n, t, h, w, dev)
            if "prev" not in buf:
                buf["prev"] = torch.empty(n, 3, h, w, device=dev, dtype=torch.float32)
            first0 = not self._has_state
            have_prev = self.pre_lr is not None
            if have_prev and self.pre_lr.data_ptr() != buf["prev"].data_ptr():
                buf["prev"].copy_(self.pre_lr)      # state carried over from a call with another t (new workspace)

            def run(lrs_, fvs_, mks_, fgs_, out_):
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                shp = L.DsvShape(n=n, t=t, h=h, w=w, mid_channels=self.mid_channels)
                # the first frame of a stream is paired with the last frame of this call (CRFP_test.py:2232-2239); its
                # flow is never used because that frame takes the no-alignment branch
                if not have_prev:
                    buf["prev"].copy_(lrs_[:, -1])
                L.check(L.lib().crfp_dsv_prepare(C.byref(shp), C.byref(W), lrs_.data_ptr(), buf["prev"].data_ptr(),
                                                 buf["lr4"].data_ptr(), buf["x_lr"].data_ptr(), buf["flows"].data_ptr(),
                                                 buf["ws"].data_ptr(), buf["ws"].numel(), st), "dsv_prepare")
                buf["prev"].copy_(lrs_[:, -1])
                self._run_frames(buf, W, lrs_, fvs_, mks_, fgs_, out_, [first0 and i == 0 for i in range(t)])

            out = None
            if self.use_graphs and not torch.cuda.is_current_stream_capturing():
                sb = self._stream_buffers(n, t, h, w, dev)
                gkey = (first0, have_prev, bool(self.skip_outside_fovea))
                g = sb["graphs"].get(gkey)
                if g is None:
                    sb["seen"][gkey] = sb["seen"].get(gkey, 0) + 1
                    if sb["seen"][gkey] >= 2 and not first0:       # steady-state calls only: capture on the second one
                        g = self._capture_stream(sb, gkey, run)
                if g is not None:
                    sb["lrs"].copy_(lrs); sb["fvs"].copy_(fvs); sb["mks"].copy_(mks); sb["fgs"].copy_(fgs)
                    g["graph"].replay()
                    L.lib().crfp_launch_count_add(g["launches"])
                    out = sb["out"] if self.alias_output else sb["out"].clone()
            if out is None:
                out = torch.empty(n, t, 3, 8 * h, 8 * w, device=dev, dtype=torch.float32)
                run(lrs, fvs, mks, fgs, out)
            self.pre_lr = buf["prev"]
            self._has_state = True
            self._state_key = skey
        return out

    def _capture_stream(self, sb, gkey, run):
        lib = L.lib()
        try:
            torch.cuda.synchronize()
            before = lib.crfp_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run(sb["lrs"], sb["fvs"], sb["mks"], sb["fgs"], sb["out"])
            launches = lib.crfp_launch_count() - before
            lib.crfp_launch_count_add(-launches)        # captured, not launched
        except Exception as e:  # noqa: BLE001
            import warnings
            warnings.warn(f"crfp_b200: CUDA graph capture of the streaming step failed ({e}); continuing without graphs")
            self.use_graphs = False
            return None
        sb["graphs"][gkey] = dict(graph=g, launches=launches)
        return sb["graphs"][gkey]
