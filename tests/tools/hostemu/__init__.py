"""TEST INFRASTRUCTURE ONLY: g++ build + ctypes handle of the host emulation of crfp_b200/csrc/bwd.cu and spynet.cu.

The backward kernels are sync-free one-thread-per-element SIMT kernels, so the very same source compiles as plain
C++ against `cuda_shim.h` (a serial loop over the launch grid) and exports the same C-ABI entry points, taking HOST
pointers.  The CPU suite uses it to check the kernels' arithmetic and indexing against torch autograd in a container
without a GPU.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", "..", ".."))
_SRCS = [os.path.join(_ROOT, "crfp_b200", "csrc", f) for f in ("bwd.cu", "spynet.cu")]   # the sync-free sources
_OUT = os.path.join(_HERE, "_build", "libcrfp_bwd_hostemu.so")
_lib = None


def build() -> str:
    deps = _SRCS + [os.path.join(_HERE, "cuda_shim.h"), os.path.join(_ROOT, "include", "crfp_b200.h")]
    if os.path.exists(_OUT) and all(os.path.getmtime(_OUT) >= os.path.getmtime(d) for d in deps):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DCRFP_HOST_EMU",
           "-I", _HERE, *_SRCS, "-o", _OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host-emulation build of bwd.cu failed:\n" + res.stderr)
    return _OUT


def lib():
    """ctypes handle with the prototypes of crfp_b200._lib.TRAIN_SYMBOLS (host pointers instead of device pointers)."""
    global _lib
    if _lib is None:
        from crfp_b200 import _lib as L
        h = C.CDLL(build())
        for name in L.TRAIN_SYMBOLS + L.SPYNET_SYMBOLS:
            res, args = L.SYMBOLS[name]
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        assert h.crfp_sizeof_dcn_bwd_desc() == C.sizeof(L.DcnBwdDesc)
        _lib = h
    return _lib


class HostEmuKernelSet:
    """Kernel set for crfp_b200.autograd in the CPU suite: the BACKWARD entry points are the host-emulated kernels of
    bwd.cu (the code under test); the forward primitives — GPU-only kernels of the inference library, verified on the
    B200 by tests/test_gpu_ops.py — are stood in for by the oracle's plain-PyTorch ops on CPU tensors."""

    name = "hostemu"

    def __init__(self, dgrad_as_conv=True, wgrad_two_stage=False):
        self.dgrad_as_conv = dgrad_as_conv      # False: exercise the direct crfp_conv3x3_bwd_data kernels
        self.wgrad_two_stage = wgrad_two_stage  # True: partial sums + reduce kernel for the thin layers

    def lib(self):
        return lib()

    def stream(self):
        return None

    def req(self, t, what):
        import torch
        assert isinstance(t, torch.Tensor) and not t.is_cuda and t.dtype == torch.float32, what
        return t.contiguous()

    @staticmethod
    def _nchw(x):
        return x.permute(0, 3, 1, 2)

    @staticmethod
    def _nhwc(x):
        return x.permute(0, 2, 3, 1).contiguous()

    def conv3x3(self, srcs, weight, bias, act, cache=None):
        import torch
        import torch.nn.functional as F
        y = F.conv2d(self._nchw(torch.cat(list(srcs), dim=-1)), weight, bias, padding=1)
        y = F.leaky_relu(y, 0.1) if act == 1 else F.relu(y) if act == 2 else y
        return self._nhwc(y)

    def dcn_v2(self, x, offset, mask, weight, bias, dg, cache=None):
        from oracle import crfp_oracle as O
        return self._nhwc(O.dcn_v2(self._nchw(x), self._nchw(offset), self._nchw(mask), weight, bias, dg))

    def flow_warp(self, x, flow):
        from oracle import crfp_oracle as O
        return self._nhwc(O.flow_warp(self._nchw(x), self._nchw(flow)))

    def resize(self, x, hout, wout, rh, rw, mul):
        import torch.nn.functional as F
        y = F.interpolate(self._nchw(x), size=(hout, wout), mode="bilinear", align_corners=False)
        # F.interpolate(size=) uses in/out as the scale, which equals rh / rw for every call site of the model
        assert abs(rh - x.shape[1] / hout) < 1e-6 and abs(rw - x.shape[2] / wout) < 1e-6
        return self._nhwc(y * mul)

    def avgpool2(self, x):
        import torch.nn.functional as F
        return self._nhwc(F.avg_pool2d(self._nchw(x), 2, 2))

    def to_nhwc(self, x):
        return self._nhwc(x)


def spy_kernels():
    """SPyNet kernel set for the CPU suite: conv_kxk / resize_ac / channel_affine are the host-emulated kernels of
    spynet.cu (the code under test); the hot-path kernels SPyNet reuses (GPU-only, verified on the B200 by
    tests/test_gpu_ops.py) are stood in for by plain PyTorch / oracle ops."""
    import torch
    import torch.nn.functional as F
    from crfp_b200.spynet import SpyKernels
    from oracle import crfp_oracle as O

    class HostEmuSpyKernels(SpyKernels):
        def lib(self):
            return lib()

        def stream(self):
            return None

        def req(self, t, what):
            assert isinstance(t, torch.Tensor) and not t.is_cuda, what
            return t.to(torch.float32).contiguous()

        def to_nhwc(self, x):
            return x.permute(0, 2, 3, 1).contiguous()

        def to_nchw(self, x):
            return x.permute(0, 3, 1, 2).contiguous()

        def avgpool2(self, x):
            return self.to_nhwc(F.avg_pool2d(self.to_nchw(x), 2, 2, count_include_pad=False))

        def resize(self, x, hout, wout):
            return self.to_nhwc(F.interpolate(self.to_nchw(x), size=(hout, wout), mode="bilinear", align_corners=False))

        def flow_warp_border(self, x, flow):
            return self.to_nhwc(O.flow_warp(self.to_nchw(x), self.to_nchw(flow), padding_mode="border"))

    return HostEmuSpyKernels()
