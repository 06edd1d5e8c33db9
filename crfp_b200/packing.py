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


_INDEX_CACHE = {}


def _conv_index_tensors(c_list, modes, ci_lo, device):
    """(gather index, validity mask) of input_index_map on `device`, built once per layer signature: the training
    step re-packs every layer's weights every iteration and must not pay a host->device copy each time."""
    key = (c_list, modes, ci_lo, str(device))
    hit = _INDEX_CACHE.get(key)
    if hit is None:
        idx = input_index_map(list(c_list), list(modes) if modes else None, ci_lo)
        sel = torch.tensor([i if i >= 0 else 0 for i in idx], device=device, dtype=torch.long)
        valid = torch.tensor([1.0 if i >= 0 else 0.0 for i in idx], device=device)
        hit = _INDEX_CACHE[key] = (sel, valid)
    return hit


def pack_conv(weight: torch.Tensor, bias: torch.Tensor, c_list, modes=None, ci_lo: int = 0):
    """OIHW (cout, cin, 3, 3) + bias -> ([9, cin_packed, cout_packed], [cout_packed]) fp32 contiguous."""
    cout = weight.shape[0]
    cp, op = cin_packed(c_list), cout_packed(cout)
    w = weight.detach().to(torch.float32)
    sel, valid = _conv_index_tensors(tuple(c_list), tuple(modes) if modes else None, ci_lo, w.device)
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


def _tc3_smem_bytes(kc_real, kc_total, nt):
    return (2 * 9 * kc_total * nt + 6 * kc_total * 130) * 16 + (2 * kc_real + 1) * 130 * 16 + (nt + 31) // 32 * 32 * 4 + (18 * nt + 32) * 4


def tc3_cout_tile(cout: int, cin: int):
    """(nt, ntiles) exactly as crfp_tc3_cout_tile computes them (None if cin is too large for the kernel)."""
    kc_real = cin // 8
    kc_total = kc_real + kc_real % 2
    tiles = (cout + 111) // 112
    while tiles <= cout:
        per = (cout + tiles - 1) // tiles
        nt = max(16, (per + 15) // 16 * 16)
        if _tc3_smem_bytes(kc_real, kc_total, nt) <= 225 * 1024:
            return nt, tiles
        if nt == 16:
            break
        tiles += 1
    return None


def split_bf16(x: torch.Tensor, dtype=torch.bfloat16):
    """x (fp32) -> (hi, lo) of `dtype` (bf16: CRFP_PREC_TC3; fp16: CRFP_PREC_HALF) with hi = dtype(x), lo = dtype(x - hi)."""
    hi = x.to(dtype)
    lo = (x - hi.to(torch.float32)).to(dtype)
    return hi, lo


def pack_conv_tc3(weight: torch.Tensor, bias: torch.Tensor, c_list, ci_lo: int = 0, extra: int = 0, modes=None,
                  dtype=torch.bfloat16):
    """OIHW fp32 -> (w_hi, w_lo bf16 [ntiles][9][kc][nt][8], bias fp32 [ntiles*nt], w_extra fp32 [9][extra][ntiles*nt])
    for crfp_conv3x3_tc3_fwd.  `c_list`: channels of the tensor-core sources (multiples of 8); `extra`: trailing
    input channels convolved on the CUDA cores (the 2 flow channels of dcn_block.0)."""
    cout = weight.shape[0]
    assert all(c % 8 == 0 for c in c_list)
    k = sum(c_list)
    tile = tc3_cout_tile(cout, k)
    if tile is None:
        raise ValueError(f"conv_tc3: {k} input channels do not fit the shared-memory rings (max 64)")
    nt, ntiles = tile
    kc = k // 8
    kc += kc % 2
    w = weight.detach().to(torch.float32)
    wm = torch.zeros(ntiles * nt, kc * 8, 9, device=w.device)
    idx = input_index_map(c_list, modes, ci_lo)[:k]           # packed K index -> original input channel
    assert all(i >= 0 for i in idx)
    wm[:cout, :k] = w[:, torch.tensor(idx, device=w.device)].reshape(cout, k, 9)
    packed = wm.view(ntiles, nt, kc, 8, 9).permute(0, 4, 2, 1, 3).contiguous()      # (tile, tap, kc, n, j)
    hi, lo = split_bf16(packed, dtype)
    b = torch.zeros(ntiles * nt, device=w.device, dtype=torch.float32)
    b[:cout] = bias.detach().to(torch.float32)
    wx = None
    if extra:
        wx = torch.zeros(9, extra, ntiles * nt, device=w.device, dtype=torch.float32)
        wx[:, :, :cout] = w[:, ci_lo + k:ci_lo + k + extra].reshape(cout, extra, 9).permute(2, 1, 0)
        wx = wx.contiguous()
    return hi.contiguous(), lo.contiguous(), b.contiguous(), wx


def pack_conv_tc3_device(weight: torch.Tensor, bias, k: int, extra: int = 0, lo: int = 0, transposed: bool = False,
                         nout: int | None = None, k_lo: int = 0):
    """pack_conv_tc3 as ONE kernel launch on the device (crfp_pack_conv_tc3; the training step repacks every layer every
    iteration).  transposed=True packs the backward-data operator of input channels [lo, lo + nout) (k = cout of `weight`)."""
    import ctypes as C
    from . import _lib as L
    cout_w, cin_w = weight.shape[0], weight.shape[1]
    nout = cout_w if nout is None else nout
    nt, ntiles = tc3_cout_tile(nout, k)
    kc = k // 8 + (k // 8) % 2
    dev = weight.device
    hi = torch.empty(ntiles, 9, kc, nt, 8, device=dev, dtype=torch.bfloat16)
    lo_t = torch.empty_like(hi)
    bp = torch.empty(ntiles * nt, device=dev, dtype=torch.float32)
    wx = torch.empty(9, extra, ntiles * nt, device=dev, dtype=torch.float32) if extra else None
    w = weight.detach()
    assert w.dtype == torch.float32 and w.is_contiguous()
    L.check(L.lib().crfp_pack_conv_tc3(w.data_ptr(), None if (bias is None or transposed) else bias.detach().data_ptr(), cout_w,
                                       cin_w, int(transposed), lo, nout, k, extra, k_lo, hi.data_ptr(), lo_t.data_ptr(), bp.data_ptr(),
                                       wx.data_ptr() if wx is not None else None,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pack_conv_tc3")
    return hi, lo_t, bp, wx


def pack_dcn_tc(weight: torch.Tensor, bias: torch.Tensor, dg: int):
    """DCNv2 weight -> bf16 UMMA B operand [K/8][cout][8] (k = (g*9+t)*(C/dg)+c) + fp32 bias, for crfp_dcn_v2_tc_fwd."""
    wk, b = pack_dcn(weight, bias, dg)                     # [K, cout]
    k, cout = wk.shape
    out = wk.view(k // 8, 8, cout).permute(0, 2, 1).contiguous()
    return out.to(torch.bfloat16).contiguous(), b


def pack_dcn_tc3(weight: torch.Tensor, bias: torch.Tensor, dg: int, dtype=torch.bfloat16):
    """DCNv2 weight -> (hi, lo) bf16 (fp16: CRFP_PREC_HALF) UMMA B operands [K/8][cout][8] + fp32 bias, for crfp_dcn_v2_tc3_fwd."""
    wk, b = pack_dcn(weight, bias, dg)
    k, cout = wk.shape
    hi, lo = split_bf16(wk.view(k // 8, 8, cout).permute(0, 2, 1).contiguous(), dtype)
    return hi.contiguous(), lo.contiguous(), b


def pack_align_heads(w_off: torch.Tensor, b_off: torch.Tensor, w_msk: torch.Tensor, b_msk: torch.Tensor, dtype=torch.bfloat16):
    """dcn_offset (144,32,3,3) + dcn_mask (72,32,3,3) -> the B operand of crfp_dcn_align_fused's heads GEMM:
    bf16 [6 sixths][hi | lo][6 chunks][224][8] and fp32 bias [224].
    Column n = 3*s + {0: dy, 1: dx, 2: mask} for sample s = g*9 + t (offset channels 2s, 2s+1; mask channel s), 216..223
    zero.  K = tap*32 + c (tap-major), chunk kc = K // 8 = tap*4 + c // 8; a sixth is 6 consecutive chunks."""
    assert tuple(w_off.shape) == (144, 32, 3, 3) and tuple(w_msk.shape) == (72, 32, 3, 3)
    dev = w_off.device
    w = torch.zeros(224, 32, 9, device=dev, dtype=torch.float32)
    b = torch.zeros(224, device=dev, dtype=torch.float32)
    wo, wm = w_off.detach().to(torch.float32).reshape(144, 32, 9), w_msk.detach().to(torch.float32).reshape(72, 32, 9)
    s = torch.arange(72, device=dev)
    w[3 * s], w[3 * s + 1], w[3 * s + 2] = wo[2 * s], wo[2 * s + 1], wm[s]
    b[3 * s], b[3 * s + 1], b[3 * s + 2] = b_off.detach().float()[2 * s], b_off.detach().float()[2 * s + 1], b_msk.detach().float()[s]
    wk = w.permute(0, 2, 1).reshape(224, 288)                       # [n][K = tap*32 + c]
    chunks = wk.view(224, 36, 8).permute(1, 0, 2).contiguous()      # [kc][n][8]
    hi, lo = split_bf16(chunks, dtype)
    packed = torch.stack([hi.view(6, 6, 224, 8), lo.view(6, 6, 224, 8)], dim=1).contiguous()   # [sixth][hi|lo][6][224][8]
    return packed, b.contiguous()


def pack_layer_tc(info: dict, sd):
    """bf16-storage tensor-core packing (experimental CRFP_PREC_BF16) of a layer with crfp_layer_info.tc in (1, 2)."""
    w, b = sd[info["key"] + ".weight"], sd[info["key"] + ".bias"]
    if info["tc"] == 2:
        return pack_dcn_tc(w, b, info["dg"])
    if info["kind"] == 2:
        w = torch.cat([w, sd[info["key2"] + ".weight"]], dim=0)
        b = torch.cat([b, sd[info["key2"] + ".bias"]], dim=0)
    return pack_conv_tc(w, b, info["c"], info["ci_lo"])


def pack_layer_tc3(info: dict, sd, dtype=torch.bfloat16):
    """CRFP_PREC_TC3 (bf16) / CRFP_PREC_HALF (fp16) packing of a layer with crfp_layer_info.tc != 0:
    (w_hi, w_lo, bias, w_extra or None)."""
    w, b = sd[info["key"] + ".weight"], sd[info["key"] + ".bias"]
    if info["tc"] == 2:
        hi, lo, bb = pack_dcn_tc3(w, b, info["dg"], dtype)
        return hi, lo, bb, None
    if info["tc"] == 4:
        # K-split (> 64 input channels): one packing per 64-channel slice, [passes][ntiles][9][8][nt][8]; bias [2][ntiles*nt] =
        # (bias, zeros) — pass 0 adds the bias, the later passes add the earlier partial sums instead
        cin = sum(info["c"])
        assert len(info["c"]) == 1 and cin % 64 == 0 and cin > 64
        parts = [pack_conv_tc3(w[:, k:k + 64].contiguous(), b, [64], 0, 0, None, dtype) for k in range(0, cin, 64)]
        hi = torch.stack([q[0] for q in parts]).contiguous()
        lo = torch.stack([q[1] for q in parts]).contiguous()
        bb = torch.stack([parts[0][2], torch.zeros_like(parts[0][2])]).contiguous()
        return hi, lo, bb, None
    if info["kind"] == 2:
        w = torch.cat([w, sd[info["key2"] + ".weight"]], dim=0)
        b = torch.cat([b, sd[info["key2"] + ".bias"]], dim=0)
    c_tc = [c for c in info["c"] if c % 8 == 0]
    extra = info["c"][-1] if info["c"][-1] % 8 else 0
    assert sum(c_tc) + extra == sum(info["c"]) and extra in (0, 2)
    m_tc = [m for c, m in zip(info["c"], info["mode"]) if c % 8 == 0]
    return pack_conv_tc3(w, b, c_tc, info["ci_lo"], extra, m_tc, dtype)


def pack_layer(info: dict, sd):
    """Pack one entry of the library's layer table (crfp_dsv_layer_info) from a state_dict."""
    w, b = sd[info["key"] + ".weight"], sd[info["key"] + ".bias"]
    if info["kind"] == 1:
        return pack_dcn(w, b, info["dg"])
    if info["kind"] == 2:  # fused heads: dcn_offset ++ dcn_mask along cout
        w = torch.cat([w, sd[info["key2"] + ".weight"]], dim=0)
        b = torch.cat([b, sd[info["key2"] + ".bias"]], dim=0)
    return pack_conv(w, b, info["c"], info["mode"], info["ci_lo"])
