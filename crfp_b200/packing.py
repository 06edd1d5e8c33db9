"""Weight repacking: reference OIHW conv / DCNv2 weights -> the layouts libcrfp_b200 consumes.

conv  (crfp_conv3x3_fwd):  weight[(tap*cin_packed + ci)*cout_packed + co], tap = ky*3+kx; each input source
      of the channel concat is padded to a multiple of 4 channels, the total to a multiple of 8; a source
      read through pixel_unshuffle(4) uses the packed channel order (dy*4+dx)*HRc + ch for the reference's
      channel ch*16 + dy*4 + dx (/root/reference/model/CRFP.py:28-42).
DCNv2 (crfp_dcn_v2_fwd):   weight[k*cout_pad + co] with k = (g*9 + tap)*(C/dg) + c_in_group.
"""
from __future__ import annotations

import torch

SRC_PLAIN, SRC_UNSHUFFLE4 = 0, 1


def cout_packed(cout: int) -> int:
    return 4 if cout <= 4 else (cout + 31) // 32 * 32


def cin_packed(c_list) -> int:
    q = sum((c + 3) // 4 for c in c_list)
    return (q * 4 + 7) // 8 * 8


def input_index_map(c_list, modes=None, ci_lo: int = 0):
    """packed input channel -> original input channel of the OIHW weight (-1 = zero padding)."""
    modes = modes or [SRC_PLAIN] * len(c_list)
    idx = []
    base = ci_lo
    for c, mode in zip(c_list, modes):
        if mode == SRC_UNSHUFFLE4:
            hc = c // 16
            for sub in range(16):
                for ch in range(hc):
                    idx.append(base + ch * 16 + sub)
        else:
            idx.extend(range(base, base + c))
        idx.extend([-1] * ((-c) % 4))
        base += c
    idx.extend([-1] * (cin_packed(c_list) - len(idx)))
    return idx


def pack_conv(weight: torch.Tensor, bias: torch.Tensor, c_list, modes=None, ci_lo: int = 0):
    """OIHW (cout, cin, 3, 3) + bias -> ([9, cin_packed, cout_packed], [cout_packed]) fp32 contiguous."""
    cout = weight.shape[0]
    cp, op = cin_packed(c_list), cout_packed(cout)
    idx = input_index_map(c_list, modes, ci_lo)
    w = weight.detach().to(torch.float32)
    sel = torch.tensor([i if i >= 0 else 0 for i in idx], device=w.device, dtype=torch.long)
    valid = torch.tensor([1.0 if i >= 0 else 0.0 for i in idx], device=w.device)
    wp = w[:, sel] * valid.view(1, -1, 1, 1)                    # (cout, cin_packed, 3, 3)
    wp = wp.permute(2, 3, 1, 0).reshape(9, cp, cout)            # (tap, ci, co)
    out = torch.zeros(9, cp, op, device=w.device, dtype=torch.float32)
    out[:, :, :cout] = wp
    b = torch.zeros(op, device=w.device, dtype=torch.float32)
    b[:cout] = bias.detach().to(torch.float32)
    return out.contiguous(), b.contiguous()


def pack_dcn(weight: torch.Tensor, bias: torch.Tensor, dg: int):
    """DCNv2 weight (cout, C, 3, 3) -> [9*C, cout_pad4] with k = (g*9+t)*(C/dg)+c; bias -> [cout_pad4]."""
    cout, c = weight.shape[0], weight.shape[1]
    cpg = c // dg
    op = (cout + 3) // 4 * 4
    w = weight.detach().to(torch.float32).reshape(cout, dg, cpg, 9)   # (o, g, c, t)
    wk = w.permute(1, 3, 2, 0).reshape(dg * 9 * cpg, cout)            # (g, t, c) -> k
    out = torch.zeros(dg * 9 * cpg, op, device=w.device, dtype=torch.float32)
    out[:, :cout] = wk
    b = torch.zeros(op, device=w.device, dtype=torch.float32)
    b[:cout] = bias.detach().to(torch.float32)
    return out.contiguous(), b.contiguous()


def tc_cout_tile(cout: int):
    """(nt, ntiles) exactly as crfp_tc_cout_tile computes them."""
    tiles = (cout + 127) // 128
    per = (cout + tiles - 1) // tiles
    return max(16, (per + 15) // 16 * 16), tiles


def pack_conv_tc(weight: torch.Tensor, bias: torch.Tensor, c_list, ci_lo: int = 0):
    """OIHW -> bf16 [ntiles][9][kc][nt][8] for crfp_conv3x3_tc_fwd (+ fp32 bias [ntiles*nt]).

    `c_list` are the REAL channels each source contributes; every source occupies ceil(c/8) chunks of 8 (the
    caller's buffers are zero in the padding channels), the chunk total is rounded up to even."""
    cout = weight.shape[0]
    nt, ntiles = tc_cout_tile(cout)
    idx, base = [], ci_lo
    for c in c_list:
        idx.extend(range(base, base + c))
        idx.extend([-1] * ((-c) % 8))
        base += c
    if (len(idx) // 8) % 2:
        idx.extend([-1] * 8)
    kc = len(idx) // 8
    w = weight.detach().to(torch.float32)
    sel = torch.tensor([i if i >= 0 else 0 for i in idx], device=w.device, dtype=torch.long)
    valid = torch.tensor([1.0 if i >= 0 else 0.0 for i in idx], device=w.device)
    wp = (w[:, sel] * valid.view(1, -1, 1, 1)).reshape(cout, kc, 8, 9)          # (o, kc, j, tap)
    full = torch.zeros(ntiles * nt, kc, 8, 9, device=w.device)
    full[:cout] = wp
    out = full.view(ntiles, nt, kc, 8, 9).permute(0, 4, 2, 1, 3).contiguous()    # (tile, tap, kc, n, j)
    b = torch.zeros(ntiles * nt, device=w.device, dtype=torch.float32)
    b[:cout] = bias.detach().to(torch.float32)
    return out.to(torch.bfloat16).contiguous(), b.contiguous()


def pack_layer(info: dict, sd):
    """Pack one entry of the library's layer table (crfp_dsv_layer_info) from a state_dict."""
    w, b = sd[info["key"] + ".weight"], sd[info["key"] + ".bias"]
    if info["kind"] == 1:
        return pack_dcn(w, b, info["dg"])
    if info["kind"] == 2:  # fused heads: dcn_offset ++ dcn_mask along cout
        w = torch.cat([w, sd[info["key2"] + ".weight"]], dim=0)
        b = torch.cat([b, sd[info["key2"] + ".bias"]], dim=0)
    return pack_conv(w, b, info["c"], info["mode"], info["ci_lo"])
