"""Parameter inventory of the CRFP_DSV hot path.

The 118-tensor state_dict of the reference model `CRFP_DSV(mid_channels=C)`
(/root/reference/model/CRFP.py:1388-1481) written down as a
table so that the module shell, the oracle and the weight generators agree on
names and shapes without importing the reference.
"""
from __future__ import annotations

from collections import OrderedDict


def _conv(d, name, cout, cin, k=3):
    d[name + ".weight"] = (cout, cin, k, k)
    d[name + ".bias"] = (cout,)


def fnet_param_shapes(prefix="spynet."):
    """FNet (/root/reference/model/CRFP.py:743-795); attribute name `spynet`."""
    d = OrderedDict()
    widths = [("encoder1", 6, 32), ("encoder2", 32, 64), ("encoder3", 64, 128),
              ("decoder1", 128, 256), ("decoder2", 256, 128), ("decoder3", 128, 64)]
    for name, cin, cout in widths:
        _conv(d, f"{prefix}{name}.0", cout, cin)
        _conv(d, f"{prefix}{name}.2", cout, cout)
    _conv(d, f"{prefix}flow.0", 32, 64)
    _conv(d, f"{prefix}flow.2", 2, 32)
    return d


def dcn_module_param_shapes(prefix, C, dg, repeat, pre_offset, pixelshuffle):
    """DCN_module (/root/reference/model/CRFP.py:282-322)."""
    d = OrderedDict()
    if pre_offset:
        if pixelshuffle:
            _conv(d, f"{prefix}upsample.upsample_conv", C * 16, C * 8)
        _conv(d, f"{prefix}conv_fuse", C, 2 * C)
    _conv(d, f"{prefix}dcn_block.0", C, 2 * C + 2)
    _conv(d, f"{prefix}dcn_block.2", C, C)
    if repeat:
        _conv(d, f"{prefix}dcn_offset", dg * 2, C)
        _conv(d, f"{prefix}dcn_mask", dg, C)
    else:
        _conv(d, f"{prefix}dcn_offset", dg * 18, C)
        _conv(d, f"{prefix}dcn_mask", dg * 9, C)
    _conv(d, f"{prefix}dcn", C, C)
    return d


VARIANTS = ("dsv", "v15", "v13")   # CRFP_DSV (v18), CRFP (v15), CRFP_simple (v13)


def crfp_param_shapes(variant="dsv", mid_channels=32, y_only=False):
    """Parameters of CRFP_DSV (CRFP.py:1388-1481), CRFP (CRFP.py:1102-1221) or CRFP_simple (CRFP.py:817-936) with
    hr_dcn=True, offset_prop=True, in module order.  The three share every module name; they differ in
    forward_resblocks_k.main.0 input width, upsample / upsample_post widths (DSV's 24/8 channel split)."""
    if variant == "dsv":
        return crfp_dsv_param_shapes(mid_channels, y_only)
    if variant not in ("v15", "v13"):
        raise ValueError(variant)
    C = mid_channels
    c = C // 8
    k_in = 3 if variant == "v15" else 2
    d = OrderedDict()
    d.update(fnet_param_shapes())
    d.update(dcn_module_param_shapes("dcn_0.", C, 8, False, False, False))
    d.update(dcn_module_param_shapes("dcn_1.", C, 8, False, True, False))
    d.update(dcn_module_param_shapes("dcn_2.", C, 8, False, True, False))
    d.update(dcn_module_param_shapes("dcn_3.", c, 1, True, True, True))
    _conv(d, "encoder_lr.slice1.0", C, 3)
    _conv(d, "encoder_lr.slice1.2", C, C)
    _conv(d, "encoder_hr.slice1.0", c, 6)
    _conv(d, "encoder_hr.slice1.2", c, c)
    _conv(d, "conv_tttf", c, 2 * c)
    for k in range(3):
        _conv(d, f"forward_resblocks_{k}.main.0", C, k_in * C)
        _conv(d, f"forward_resblocks_{k}.main.2.0.conv1", C, C)
        _conv(d, f"forward_resblocks_{k}.main.2.0.conv2", C, C)
    _conv(d, "forward_resblocks_3.main.0", c, k_in * c)
    _conv(d, "forward_resblocks_3.main.2.0.conv1", c, c)
    _conv(d, "forward_resblocks_3.main.2.0.conv2", c, c)
    _conv(d, "downsample.downsample_conv", C, c * 16)
    _conv(d, "upsample.upsample_conv", C * 4, C)
    _conv(d, "upsample_post.upsample_conv", c * 16, C)
    _conv(d, "conv_last", 1 if y_only else 3, c)
    return d


def crfp_dsv_param_shapes(mid_channels=32, y_only=False):
    """All parameters of CRFP_DSV(hr_dcn=True, offset_prop=True) in module order."""
    C = mid_channels
    c = C // 8
    d = OrderedDict()
    d.update(fnet_param_shapes())
    d.update(dcn_module_param_shapes("dcn_0.", C, 8, False, False, False))
    d.update(dcn_module_param_shapes("dcn_1.", C, 8, False, True, False))
    d.update(dcn_module_param_shapes("dcn_2.", C, 8, False, True, False))
    d.update(dcn_module_param_shapes("dcn_3.", c, 1, True, True, True))
    _conv(d, "encoder_lr.slice1.0", C, 3)
    _conv(d, "encoder_lr.slice1.2", C, C)
    _conv(d, "encoder_hr.slice1.0", c, 6)
    _conv(d, "encoder_hr.slice1.2", c, c)
    _conv(d, "conv_tttf", c, 2 * c)
    for k in range(3):
        _conv(d, f"forward_resblocks_{k}.main.0", C, 2 * C)
        _conv(d, f"forward_resblocks_{k}.main.2.0.conv1", C, C)
        _conv(d, f"forward_resblocks_{k}.main.2.0.conv2", C, C)
    _conv(d, "forward_resblocks_3.main.0", c, 2 * c)
    _conv(d, "forward_resblocks_3.main.2.0.conv1", c, c)
    _conv(d, "forward_resblocks_3.main.2.0.conv2", c, c)
    _conv(d, "downsample.downsample_conv", C, c * 16)
    _conv(d, "upsample.upsample_conv", (C * 3 // 4) * 4, C)
    _conv(d, "upsample_post.upsample_conv", c * 16, C * 3 // 4)
    _conv(d, "conv_last", 1 if y_only else 3, c)
    return d
