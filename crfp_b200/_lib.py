"""ctypes binding of libcrfp_b200.so (the C ABI declared in include/crfp_b200.h).

The shared library is built in-tree by `crfp_b200.build.build()` (nvcc, sm_100a).  There is NO CPU
fallback: importing this module without the built library, or calling an op without a CUDA device,
raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcrfp_b200.so")

CRFP_OK = 0
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_DCN_HEAD, ACT_TANH256 = 0, 1, 2, 3, 4
SRC_PLAIN, SRC_UNSHUFFLE4 = 0, 1
OUT_NHWC, OUT_SHUFFLE = 0, 1
TC_OUT_BF16, TC_OUT_F32, TC_OUT_SHUFFLE_F32 = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_TC3, PREC_HALF = 0, 1, 2, 3
MAX_LAYERS = 72

c_float_p = C.POINTER(C.c_float)


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("c", C.c_int32), ("cstride", C.c_int32), ("coffset", C.c_int32),
                ("mode", C.c_int32)]


class Dst(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("c", C.c_int32), ("cstride", C.c_int32), ("coffset", C.c_int32),
                ("_pad", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("nsrc", C.c_int32),
                ("src", Src * 3),
                ("cout", C.c_int32), ("act", C.c_int32),
                ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("out_mode", C.c_int32), ("shuffle_r", C.c_int32), ("ndst", C.c_int32), ("head_split", C.c_int32),
                ("dst", Dst * 2),
                ("residual", C.c_void_p), ("res_cstride", C.c_int32), ("res_coffset", C.c_int32),
                ("flow", C.c_void_p), ("post_scale", C.c_float), ("head_mag", C.c_float)]


class TcSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("c", C.c_int32), ("cstride", C.c_int32), ("coffset", C.c_int32), ("_pad", C.c_int32)]


class ConvTcDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("nsrc", C.c_int32),
                ("src", TcSrc * 3),
                ("cout", C.c_int32), ("act", C.c_int32),
                ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("out_kind", C.c_int32), ("shuffle_r", C.c_int32), ("ndst", C.c_int32), ("head_split", C.c_int32),
                ("dst", TcSrc * 2),
                ("residual", C.c_void_p), ("res_cstride", C.c_int32), ("res_coffset", C.c_int32),
                ("flow", C.c_void_p), ("post_scale", C.c_float), ("head_mag", C.c_float)]


class ConvTc3Desc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("nsrc", C.c_int32),
                ("src", TcSrc * 3),
                ("cout", C.c_int32), ("act", C.c_int32),
                ("weight_hi", C.c_void_p), ("weight_lo", C.c_void_p), ("bias", C.c_void_p),
                ("extra", C.c_void_p), ("w_extra", C.c_void_p),
                ("out_kind", C.c_int32), ("shuffle_r", C.c_int32), ("ndst", C.c_int32), ("head_split", C.c_int32),
                ("dst", TcSrc * 2),
                ("residual", C.c_void_p), ("res_cstride", C.c_int32), ("res_coffset", C.c_int32),
                ("flow", C.c_void_p), ("post_scale", C.c_float), ("head_mag", C.c_float),
                ("half", C.c_int32), ("res_pre", C.c_int32)]


class WarpDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("x", C.c_void_p), ("x_cstride", C.c_int32), ("x_coffset", C.c_int32),
                ("flow", C.c_void_p),
                ("out", C.c_void_p), ("out_cstride", C.c_int32), ("out_coffset", C.c_int32),
                ("border", C.c_int32), ("_pad", C.c_int32)]


class DcnDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("cout", C.c_int32), ("dg", C.c_int32),
                ("shared_taps", C.c_int32), ("head_raw", C.c_int32),
                ("x", C.c_void_p), ("x_cstride", C.c_int32), ("x_coffset", C.c_int32),
                ("offset", C.c_void_p), ("off_cstride", C.c_int32), ("off_coffset", C.c_int32),
                ("mask", C.c_void_p), ("mask_cstride", C.c_int32), ("mask_coffset", C.c_int32),
                ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("out", C.c_void_p), ("out_cstride", C.c_int32), ("out_coffset", C.c_int32),
                ("head_flow", C.c_void_p), ("head_mag", C.c_float), ("_pad", C.c_int32),
                ("dbg_y0", C.c_void_p), ("dbg_x0", C.c_void_p)]


class AlignFusedDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("_pad", C.c_int32),
                ("z", C.c_void_p), ("z_cstride", C.c_int32), ("z_coffset", C.c_int32),
                ("flow", C.c_void_p),
                ("x", C.c_void_p), ("x_cstride", C.c_int32), ("x_coffset", C.c_int32),
                ("heads_w", C.c_void_p), ("heads_b", C.c_void_p),
                ("dcn_w_hi", C.c_void_p), ("dcn_w_lo", C.c_void_p), ("dcn_b", C.c_void_p),
                ("out", C.c_void_p), ("out_cstride", C.c_int32), ("out_coffset", C.c_int32),
                ("head_mag", C.c_float), ("half", C.c_int32),
                ("dbg_y0", C.c_void_p), ("dbg_x0", C.c_void_p)]


class DcnBwdDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32), ("cout", C.c_int32), ("dg", C.c_int32),
                ("x", C.c_void_p), ("offset", C.c_void_p), ("mask", C.c_void_p), ("weight", C.c_void_p),
                ("dout", C.c_void_p), ("dx", C.c_void_p), ("doffset", C.c_void_p), ("dmask", C.c_void_p),
                ("dweight", C.c_void_p), ("dbias", C.c_void_p), ("col", C.c_void_p), ("weight_t", C.c_void_p),
                ("wg_workspace", C.c_void_p), ("wg_ws_floats", C.c_size_t)]


class Layer(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p)]


class LayerTc(C.Structure):
    _fields_ = [("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("b", C.c_void_p), ("w_extra", C.c_void_p),
                ("w_fused", C.c_void_p), ("b_fused", C.c_void_p)]


class DsvWeights(C.Structure):
    _fields_ = [("mid_channels", C.c_int32), ("nlayers", C.c_int32), ("precision", C.c_int32), ("variant", C.c_int32),
                ("layer", Layer * MAX_LAYERS), ("layer_tc", LayerTc * MAX_LAYERS)]


class LayerInfo(C.Structure):
    _fields_ = [("key", C.c_char_p), ("key2", C.c_char_p), ("kind", C.c_int32), ("nsrc", C.c_int32),
                ("c", C.c_int32 * 3), ("mode", C.c_int32 * 3), ("cout", C.c_int32), ("ci_lo", C.c_int32),
                ("dg", C.c_int32), ("thin", C.c_int32), ("tc", C.c_int32), ("_pad", C.c_int32)]


class DsvShape(C.Structure):
    _fields_ = [("n", C.c_int32), ("t", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("mid_channels", C.c_int32), ("_pad", C.c_int32)]


class DsvFrameDesc(C.Structure):
    _fields_ = [("shape", DsvShape), ("first", C.c_int32), ("skip_outside_fovea", C.c_int32),
                ("lr4", C.c_void_p), ("lr4_clip_stride", C.c_longlong),
                ("x_lr", C.c_void_p), ("x_lr_clip_stride", C.c_longlong),
                ("flow", C.c_void_p), ("flow_clip_stride", C.c_longlong),
                ("fvs", C.c_void_p), ("fvs_clip_stride", C.c_longlong),
                ("mks", C.c_void_p), ("mks_clip_stride", C.c_longlong),
                ("fg", C.c_void_p), ("fg_clip_stride", C.c_longlong),
                ("state_hr", C.c_void_p), ("state_l1", C.c_void_p),
                ("out", C.c_void_p), ("out_clip_stride", C.c_longlong),
                ("aux_stream", C.c_void_p), ("aux_events", C.c_void_p * 3)]


# every symbol include/crfp_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "crfp_abi_version": (C.c_int, []),
    "crfp_status_string": (C.c_char_p, [C.c_int]),
    "crfp_last_cuda_error": (C.c_char_p, []),
    "crfp_launch_count": (C.c_longlong, []),
    "crfp_launch_count_reset": (None, []),
    "crfp_launch_count_add": (None, [C.c_longlong]),
    "crfp_check_device": (C.c_int, []),
    "crfp_selftest_umma": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_selftest_umma_sbo": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_selftest_umma_rate": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "crfp_conv3x3_fwd": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "crfp_conv3x3_tc_fwd": (C.c_int, [C.POINTER(ConvTcDesc), C.c_void_p]),
    "crfp_tc_cout_tile": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "crfp_sizeof_conv_tc_desc": (C.c_size_t, []),
    "crfp_conv3x3_tc3_fwd": (C.c_int, [C.POINTER(ConvTc3Desc), C.c_void_p]),
    "crfp_conv3x3_tc3_trace": (C.c_int, [C.POINTER(ConvTc3Desc), C.c_void_p, C.c_void_p]),
    "crfp_tc3_cout_tile": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "crfp_sizeof_conv_tc3_desc": (C.c_size_t, []),
    "crfp_conv_cin_packed": (C.c_int, [C.c_int, C.POINTER(C.c_int32)]),
    "crfp_conv_cout_packed": (C.c_int, [C.c_int]),
    "crfp_sizeof_conv_desc": (C.c_size_t, []),
    "crfp_flow_warp_fwd": (C.c_int, [C.POINTER(WarpDesc), C.c_void_p]),
    "crfp_flow_warp_bf16_fwd": (C.c_int, [C.POINTER(WarpDesc), C.c_void_p]),
    "crfp_flow_warp_indices": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_sizeof_warp_desc": (C.c_size_t, []),
    "crfp_dcn_v2_fwd": (C.c_int, [C.POINTER(DcnDesc), C.c_void_p]),
    "crfp_dcn_v2_tc_fwd": (C.c_int, [C.POINTER(DcnDesc), C.c_void_p]),
    "crfp_dcn_v2_tc3_fwd": (C.c_int, [C.POINTER(DcnDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_dcn_align_fused": (C.c_int, [C.POINTER(AlignFusedDesc), C.c_void_p]),
    "crfp_dcn_align_fused_trace": (C.c_int, [C.POINTER(AlignFusedDesc), C.c_void_p, C.c_void_p]),
    "crfp_sizeof_align_fused_desc": (C.c_size_t, []),
    "crfp_dcn_v2_indices": (C.c_int, [C.POINTER(DcnDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_sizeof_dcn_desc": (C.c_size_t, []),
    "crfp_resize_bilinear": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                       C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "crfp_avgpool2": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_fovea_paste": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p]),
    "crfp_fovea_from_gt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "crfp_quantize_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "crfp_psnr_ssim_tiles": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "crfp_psnr_ssim": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "crfp_nchw_to_nhwc": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "crfp_nhwc_to_nchw": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                    C.c_longlong, C.c_void_p]),
    "crfp_dsv_num_layers": (C.c_int, []),
    "crfp_dsv_layer_info": (C.c_int, [C.c_int, C.POINTER(LayerInfo)]),
    "crfp_layer_info_variant": (C.c_int, [C.c_int, C.c_int, C.POINTER(LayerInfo)]),
    "crfp_dsv_prepare_workspace": (C.c_size_t, [C.POINTER(DsvShape)]),
    "crfp_dsv_prepare": (C.c_int, [C.POINTER(DsvShape), C.POINTER(DsvWeights), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "crfp_dsv_frame_workspace": (C.c_size_t, [C.POINTER(DsvShape)]),
    "crfp_dsv_frame": (C.c_int, [C.POINTER(DsvFrameDesc), C.POINTER(DsvWeights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "crfp_sizeof_dsv_weights": (C.c_size_t, []),
    "crfp_sizeof_dsv_frame_desc": (C.c_size_t, []),
    # training: backward kernels, loss, optimiser (bwd.cu)
    "crfp_act_bwd": (C.c_int, [C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "crfp_conv3x3_bwd_data": (C.c_int, [C.c_int] * 7 + [C.c_void_p] * 4),
    "crfp_conv3x3_bwd_weight": (C.c_int, [C.c_int] * 7 + [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]),
    "crfp_pack_conv_tc3": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p] * 5),
    "crfp_conv3x3_bwd_weight_workspace": (C.c_size_t, [C.c_int] * 5),
    "crfp_conv3x3_bwd_weight_batched": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)] + [C.c_int] * 7 +
                                        [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p]),
    "crfp_dcn_v2_bwd_workspace": (C.c_size_t, [C.c_int] * 5),
    "crfp_fovea_blend_fwd": (C.c_int, [C.c_longlong, C.c_int] + [C.c_void_p] * 5),
    "crfp_fovea_blend_bwd": (C.c_int, [C.c_longlong, C.c_int] + [C.c_void_p] * 6),
    "crfp_dcn_heads_act_fwd": (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 5),
    "crfp_dcn_heads_act_bwd": (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 6),
    "crfp_dcn_v2_bwd": (C.c_int, [C.POINTER(DcnBwdDesc), C.c_void_p]),
    "crfp_sizeof_dcn_bwd_desc": (C.c_size_t, []),
    "crfp_flow_warp_bwd": (C.c_int, [C.c_int] * 4 + [C.c_void_p] * 6),
    "crfp_resize_bilinear_bwd": (C.c_int, [C.c_int] * 6 + [C.c_float] * 3 + [C.c_void_p] * 3),
    "crfp_avgpool2_bwd": (C.c_int, [C.c_int] * 4 + [C.c_void_p] * 3),
    "crfp_charbonnier_fwd_bwd": (C.c_int, [C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "crfp_adam_step": (C.c_int, [C.c_longlong] + [C.c_void_p] * 4 + [C.c_float] * 5 + [C.c_void_p]),
    # SPyNet operators (spynet.cu)
    "crfp_conv_kxk_fwd": (C.c_int, [C.c_int] * 7 + [C.c_void_p] * 6),
    "crfp_resize_bilinear_ac": (C.c_int, [C.c_int] * 4 + [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "crfp_channel_affine": (C.c_int, [C.c_longlong, C.c_int, C.c_int] + [C.c_void_p] * 6),
}

# the training entry points alone: also exported by the host-emulation build the CPU tests use (tests/tools/hostemu)
TRAIN_SYMBOLS = [k for k in SYMBOLS if "_bwd" in k or k in ("crfp_adam_step", "crfp_dcn_heads_act_fwd", "crfp_fovea_blend_fwd")]
SPYNET_SYMBOLS = ["crfp_conv_kxk_fwd", "crfp_resize_bilinear_ac", "crfp_channel_affine"]

_lib = None


class CrfpError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CrfpError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        for struct, fn in ((ConvDesc, h.crfp_sizeof_conv_desc), (ConvTcDesc, h.crfp_sizeof_conv_tc_desc), (ConvTc3Desc, h.crfp_sizeof_conv_tc3_desc), (WarpDesc, h.crfp_sizeof_warp_desc),
                           (DcnDesc, h.crfp_sizeof_dcn_desc), (DcnBwdDesc, h.crfp_sizeof_dcn_bwd_desc), (DsvWeights, h.crfp_sizeof_dsv_weights),
                           (DsvFrameDesc, h.crfp_sizeof_dsv_frame_desc)):
            if C.sizeof(struct) != fn():
                raise CrfpError(f"ABI mismatch: sizeof({struct.__name__}) python {C.sizeof(struct)} != C {fn()}")
        _lib = h
    return _lib


def check(status: int, what: str = "") -> None:
    if status != CRFP_OK:
        h = lib()
        msg = h.crfp_status_string(status).decode()
        if status == -4:
            msg += ": " + h.crfp_last_cuda_error().decode()
        raise CrfpError(f"libcrfp_b200 {what} failed: {msg} ({status})")


VARIANTS = {"dsv": 0, "v15": 1, "v13": 2}


def layer_table(variant: str = "dsv"):
    """[(key, key2, kind, [c...], [mode...], cout, ci_lo, dg, thin, tc)] of a model variant, as exported by the library."""
    h = lib()
    out = []
    for i in range(h.crfp_dsv_num_layers()):
        info = LayerInfo()
        check(h.crfp_layer_info_variant(VARIANTS[variant], i, C.byref(info)), "layer_info")
        out.append(dict(key=info.key.decode(), key2=info.key2.decode() if info.key2 else None, kind=info.kind,
                        c=[info.c[j] for j in range(info.nsrc)], mode=[info.mode[j] for j in range(info.nsrc)],
                        cout=info.cout, ci_lo=info.ci_lo, dg=info.dg, thin=info.thin, tc=info.tc))
    return out
