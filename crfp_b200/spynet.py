"""SPyNet — the legacy 6-level flow pyramid, drop-in for `model.CRFP.SPyNet`
(/root/reference/model/CRFP.py:554-685; SPyNetBasicModule :687-741; `conv` :145-152).

No shipped model uses it (CRFP_DSV estimates flow with FNet); only the legacy `CRFP_runtime` classes do
(model/CRFP_runtime.py:831).  Same constructor, `forward(ref, supp) -> flow (n,2,h,w)` and state_dict
(`basic_module.{level}.basic_module.{i}.conv.{weight,bias}` + the `mean` / `std` buffers).  Every arithmetic step runs in
libcrfp_b200.so: the 7x7 `conv(act(x))` layers, the align_corners=True x2 of the flow and the normalisation / rescale
(spynet.cu), and the avg-pool, align_corners=False resize and border-mode flow_warp kernels of the hot path.  CUDA tensors
only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import zlib

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L

WIDTHS = ((8, 32), (32, 64), (64, 32), (32, 16), (16, 2))


class SpyKernels:
    """The library calls SPyNet needs (dense fp32 NHWC tensors in, fresh tensors out)."""

    def lib(self):
        return L.lib()

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def req(self, t, what):
        if not (isinstance(t, torch.Tensor) and t.is_cuda):
            raise L.CrfpError(f"{what} must be a CUDA tensor (libcrfp_b200 has no CPU path)")
        return t.to(torch.float32).contiguous()

    def _chk(self, status, what):
        if status != 0:
            raise L.CrfpError(f"libcrfp_b200 {what} failed with status {status}")

    # ---- spynet.cu
    def conv_kxk(self, x, w_packed, bias, k, relu_in, residual=None):
        n, h, w, cin = x.shape
        cout = bias.numel()
        out = torch.empty(n, h, w, cout, device=x.device, dtype=torch.float32)
        self._chk(self.lib().crfp_conv_kxk_fwd(n, h, w, cin, cout, k, int(relu_in), x.data_ptr(), w_packed.data_ptr(),
                                               bias.data_ptr(), None if residual is None else residual.data_ptr(),
                                               out.data_ptr(), self.stream()), "conv_kxk")
        return out

    def resize_ac(self, x, hout, wout, mul):
        n, h, w, c = x.shape
        out = torch.empty(n, hout, wout, c, device=x.device, dtype=torch.float32)
        self._chk(self.lib().crfp_resize_bilinear_ac(n, h, w, c, x.data_ptr(), hout, wout, float(mul), out.data_ptr(),
                                                     self.stream()), "resize_bilinear_ac")
        return out

    def affine(self, x, sub, div, mul, c_out):
        n, h, w, c = x.shape
        out = torch.empty(n, h, w, c_out, device=x.device, dtype=torch.float32)
        self._chk(self.lib().crfp_channel_affine(n * h * w, c, c_out, x.data_ptr(), sub.data_ptr(), div.data_ptr(),
                                                 mul.data_ptr(), out.data_ptr(), self.stream()), "channel_affine")
        return out

    # ---- hot-path kernels reused
    def to_nhwc(self, x):
        from . import ops
        return ops.to_nhwc(x)

    def to_nchw(self, x):
        from . import ops
        return ops.to_nchw(x)

    def avgpool2(self, x):
        from . import ops
        return ops.avgpool2_nhwc(x)

    def resize(self, x, hout, wout):
        from . import ops
        return ops.resize_bilinear_nhwc(x, hout, wout, x.shape[1] / hout, x.shape[2] / wout, 1.0)

    def flow_warp_border(self, x, flow):
        from . import ops
        return ops.flow_warp_nhwc(x, flow, border=True)


CUDA = SpyKernels()


class conv(nn.Module):
    """`conv` of the reference (model/CRFP.py:145-152): forward = conv(ReLU(x)); parameter holder here."""

    def __init__(self, in_channels, out_channels, kernel_size=7, stride=1, padding=3):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=True)


class SPyNetBasicModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.basic_module = nn.Sequential(*[conv(ci, co, 7, 1, 3) for ci, co in WIDTHS])


class SPyNet(nn.Module):
    def __init__(self, pretrained, device, kernels=None):
        super().__init__()
        self.device = device
        self.K = kernels or CUDA
        self.basic_module = nn.ModuleList([SPyNetBasicModule() for _ in range(6)])
        if isinstance(pretrained, str):
            saved = {k: v for k, v in torch.load(pretrained, map_location=self.device).items()}
            sd = self.state_dict()
            sd.update(saved)
            self.load_state_dict(sd, strict=True)
        elif pretrained is not None:
            raise TypeError(f"[pretrained] should be str or None, but got {type(pretrained)}.")
        self.register_buffer("mean", torch.Tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer("std", torch.Tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))
        self._packed = None

    def _weights(self, device):
        params = list(self.parameters())
        key = (str(device), tuple((p.data_ptr(), p._version) for p in params))
        if self._packed is None or self._packed[0] != key:
            layers = []
            for level in range(6):
                row = []
                for i in range(5):
                    cv = self.basic_module[level].basic_module[i].conv
                    wt = cv.weight.detach().to(device=device, dtype=torch.float32)               # (cout, cin, 7, 7)
                    wp = wt.permute(2, 3, 1, 0).reshape(49, wt.shape[1], wt.shape[0]).contiguous()  # [tap][cin][cout]
                    row.append((wp, cv.bias.detach().to(device=device, dtype=torch.float32).contiguous()))
                layers.append(row)
            self._packed = (key, layers)
        return self._packed[1]

    @torch.no_grad()
    def compute_flow(self, ref, supp):
        """ref / supp: NHWC 3-channel planes whose sizes are multiples of 32 -> NHWC 2-channel flow (model/CRFP.py:593-650)."""
        K = self.K
        n, h, w, _ = ref.shape
        dev = ref.device
        layers = self._weights(dev)
        mean = self.mean.reshape(3).to(device=dev, dtype=torch.float32)
        std = self.std.reshape(3).to(device=dev, dtype=torch.float32)
        one = torch.ones(3, device=dev, dtype=torch.float32)
        # normalised images as 4-channel planes (4th channel zero: the warp kernel moves float4 pixels)
        refs = [K.affine(ref, mean, std, one, 4)]
        supps = [K.affine(supp, mean, std, one, 4)]
        for _ in range(5):
            refs.append(K.avgpool2(refs[-1]))
            supps.append(K.avgpool2(supps[-1]))
        refs, supps = refs[::-1], supps[::-1]
        flow = torch.zeros(n, h // 32, w // 32, 2, device=dev, dtype=torch.float32)
        for level in range(6):
            if level == 0:
                flow_up = flow
            else:
                flow_up = K.resize_ac(flow, flow.shape[1] * 2, flow.shape[2] * 2, 2.0)
            warped = K.flow_warp_border(supps[level], flow_up)
            x = torch.cat([refs[level][..., :3], warped[..., :3], flow_up], dim=-1).contiguous()   # 8 channels
            for i in range(5):
                wp, b = layers[level][i]
                x = K.conv_kxk(x, wp, b, 7, True, residual=flow_up if i == 4 else None)
            flow = x
        return flow

    @torch.no_grad()
    def forward(self, ref, supp):
        """flow from ref to supp, NCHW (n,3,h,w) x 2 -> (n,2,h,w) (model/CRFP.py:652-685)."""
        K = self.K
        ref, supp = K.req(ref, "ref"), K.req(supp, "supp")
        if ref.shape != supp.shape or ref.dim() != 4 or ref.shape[1] != 3:
            raise ValueError(f"ref and supp must both be (n,3,h,w), got {tuple(ref.shape)} and {tuple(supp.shape)}")
        h, w = ref.shape[2:4]
        w_up = w if (w % 32) == 0 else 32 * (w // 32 + 1)
        h_up = h if (h % 32) == 0 else 32 * (h // 32 + 1)
        r, s = K.to_nhwc(ref), K.to_nhwc(supp)
        if (h_up, w_up) != (h, w):
            r, s = K.resize(r, h_up, w_up), K.resize(s, h_up, w_up)
        flow = self.compute_flow(r, s)
        if (h_up, w_up) != (h, w):
            flow = K.resize(flow, h, w)
        dev = flow.device
        zero2 = torch.zeros(2, device=dev, dtype=torch.float32)
        one2 = torch.ones(2, device=dev, dtype=torch.float32)
        scale = torch.tensor([float(w) / float(w_up), float(h) / float(h_up)], dtype=torch.float32).to(dev)
        return K.to_nchw(K.affine(flow, zero2, one2, scale, 2))


# ---- seeded synthetic weights / inputs (no checkpoints or datasets in this environment)
def make_spynet_state_dict(seed: int = 11, gain: float = 0.6):
    """Random fp32 weights under the reference's SPyNet parameter names: N(0, (gain*sqrt(2/fan_in))^2), small biases."""
    sd = {}
    for level in range(6):
        for i, (ci, co) in enumerate(WIDTHS):
            for kind, shape in (("weight", (co, ci, 7, 7)), ("bias", (co,))):
                key = f"basic_module.{level}.basic_module.{i}.conv.{kind}"
                g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
                std = gain * (2.0 / (ci * 49)) ** 0.5 if kind == "weight" else 0.02
                sd[key] = (torch.randn(shape, generator=g) * std).contiguous()
    return sd


def make_spynet_pair(seed: int, n: int, h: int, w: int):
    """Two smooth random frames in [0,1] (the second a perturbed copy of the first)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(n, 3, max(h // 8, 2), max(w // 8, 2), generator=g)
    ref = F.interpolate(coarse, size=(h, w), mode="bicubic", align_corners=False)
    drift = F.interpolate(torch.rand(n, 3, max(h // 8, 2), max(w // 8, 2), generator=g), size=(h, w), mode="bicubic",
                          align_corners=False)
    ref = (ref + 0.05 * torch.rand(n, 3, h, w, generator=g)).clamp_(0, 1)
    supp = (0.8 * ref + 0.2 * drift + 0.05 * torch.rand(n, 3, h, w, generator=g)).clamp_(0, 1)
    return ref.contiguous(), supp.contiguous()
