// fp32-accurate tensor-core 3x3 convolution: fp32 NHWC activations in HBM, 3 x bf16 split products on tcgen05.
//
// Every fp32 operand is split as x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits kept) and the
// contraction is evaluated as A_hi*W_hi + A_lo*W_hi + A_hi*W_lo with fp32 accumulation in TMEM (the dropped
// A_lo*W_lo term is ~2^-18 relative).  This keeps the <= 1e-3 fp32 parity bar of BASELINE.json (measured ~1e-4 end
// to end) while the dense L1 contractions of CRFP (/root/reference/model/CRFP.py:303-317, 433-552, 154-193) run on
// the 5th-gen tensor cores; MMA work triples but these layers are HBM/latency bound, not MMA bound.
//
// Structure is that of conv_tc.cu (128-pixel row tiles, 3x3 taps as shifted descriptor start addresses into a
// ring of staged rows in the K-major no-swizzle UMMA layout) plus a conversion stage: each input row is fetched
// once with cp.async into an fp32 staging buffer, then split into the hi / lo bf16 rings by all threads while
// the previous row's MMAs and epilogue are in flight.  A 2-channel fp32 "extra" source (the optical flow input of
// dcn_block.0) is convolved on the CUDA cores in the epilogue instead of being padded into the K dimension.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace crfp {

constexpr int T3M = 128;
constexpr int T3WP = 130;

// Stage one input row (tile + halo) of every source into the fp32 staging buffer, laid out [px][kc_real] 32-byte
// records = the global NHWC order, so a dense source row is copied as a flat run of 16-byte cp.async pieces.
__device__ __forceinline__ void tc3_stage_row(const Tc3Params& P, float4* stage, int n, int y, int x0, int tid) {
  const bool yin = (y >= 0 && y < P.h);
  const int kcr = P.kc_real;
#pragma unroll 1
  for (int s = 0; s < P.nsrc; ++s) {
    const int q = P.src_c[s] >> 2;                 // 16-byte pieces per pixel of this source
    const int items = T3WP * q;
    const float* rowp = P.src[s] + (((size_t)n * P.h + (yin ? y : 0)) * (size_t)P.w) * P.src_cstride[s] + P.src_coffset[s];
    float4* st = stage + P.kstart[s] * 2;          // first piece of this source inside a pixel's staging record
    int px = tid / q, j = tid - px * q;            // one division per source per row
    const int dpx = 128 / q, dj = 128 - dpx * q;
    for (int it = tid; it < items; it += 128) {
      const int x = x0 + px - 1;
      const bool in = yin && x >= 0 && x < P.w;
      const float* g = in ? rowp + (size_t)x * P.src_cstride[s] + j * 4 : P.src[s];
      umma::cp_async16(st + px * (kcr * 2 + 1) + j, g, in ? 16u : 0u);
      px += dpx; j += dj;
      if (j >= q) { j -= q; ++px; }
    }
  }
}

__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    const float2 hf = __bfloat1622float2(hh);
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = umma::pack_bf16(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// staged fp32 row [px][kc] -> hi / lo bf16 records [kc][px] of one ring slot (optionally x regional mask fg)
__device__ __forceinline__ void tc3_convert_row(const Tc3Params& P, const float4* stage, uint4* slot_hi, uint4* slot_lo,
                                                int n, int y, int x0, int tid) {
  const int kcr = P.kc_real;
#pragma unroll 1
  for (int px = tid; px < T3WP; px += 128) {
    float f = 1.f;
    if (P.fg != nullptr) {
      const int x = x0 + px - 1;
      f = (y >= 0 && y < P.h && x >= 0 && x < P.w) ? __ldg(P.fg + (size_t)n * P.fg_clip_stride + (size_t)y * P.w + x) : 0.f;
    }
    const float4* sp = stage + px * (kcr * 2 + 1);   // +1: odd pitch keeps the LDS.128 conflict-free
    for (int kc = 0; kc < kcr; ++kc) {
      float4 a = sp[2 * kc], b = sp[2 * kc + 1];
      if (P.fg != nullptr) { a.x *= f; a.y *= f; a.z *= f; a.w *= f; b.x *= f; b.y *= f; b.z *= f; b.w *= f; }
      uint4 hi, lo;
      split8(a, b, hi, lo);
      slot_hi[kc * T3WP + px] = hi;
      slot_lo[kc * T3WP + px] = lo;
    }
  }
}

// One output row: 9 taps x KC/2 K-steps x 3 split products.  All operand offsets are compile-time multiples of the
// runtime slot bases / NT, so each tcgen05.mma costs a couple of integer adds on the single issuing thread.
template <int KC>
__device__ __forceinline__ void tc3_issue_row(uint64_t dAh, uint64_t dAl, uint64_t dBh, uint64_t dBl, uint32_t s0, uint32_t s1,
                                              uint32_t s2, uint32_t NT, uint32_t idesc, uint32_t taddr, uint32_t leader) {
  const uint32_t ahl = (uint32_t)dAh, ahh = (uint32_t)(dAh >> 32), all_ = (uint32_t)dAl, alh = (uint32_t)(dAl >> 32);
  const uint32_t bhl = (uint32_t)dBh, bhh = (uint32_t)(dBh >> 32), bll = (uint32_t)dBl, blh = (uint32_t)(dBl >> 32);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    const uint32_t arow = (ky == 0 ? s0 : ky == 1 ? s1 : s2) + kx;
#pragma unroll
    for (int ks = 0; ks < KC / 2; ++ks) {
      const uint32_t a = arow + 2 * ks * T3WP, b = (uint32_t)(tap * KC + 2 * ks) * NT;
      const uint64_t dah = umma::desc_advance(ahl, ahh, a), dal = umma::desc_advance(all_, alh, a);
      const uint64_t dbh = umma::desc_advance(bhl, bhh, b), dbl = umma::desc_advance(bll, blh, b);
      umma::mma_bf16(taddr, dah, dbh, idesc, (tap | ks) != 0 ? 1u : 0u);
      umma::mma_bf16(taddr, dal, dbh, idesc, 1u);
      umma::mma_bf16(taddr, dah, dbl, idesc, 1u);
    }
  }
}

// N-concatenated form (cout tile of 32): the hi and lo halves of the weights sit side by side as ONE 64-wide B operand
// ([tap][kc][hi 32 | lo 32]), so A_hi x [W_hi | W_lo] is a single N = 64 MMA (columns 0..31 and 32..63 of the
// accumulator) and A_lo x W_hi an N = 32 MMA into columns 0..31: 2 instead of 3 MMAs per K step (the MMA cost here is
// dominated by streaming the 4 KB A operand, 45 / 48 cycles for N = 32 / 64).  The epilogue adds the two halves.
template <int KC>
__device__ __forceinline__ void tc3_issue_row_cat(uint64_t dAh, uint64_t dAl, uint64_t dB, uint32_t s0, uint32_t s1, uint32_t s2,
                                                  uint32_t idesc64, uint32_t idesc32, uint32_t taddr) {
  const uint32_t ahl = (uint32_t)dAh, ahh = (uint32_t)(dAh >> 32), all_ = (uint32_t)dAl, alh = (uint32_t)(dAl >> 32);
  const uint32_t bl = (uint32_t)dB, bh = (uint32_t)(dB >> 32);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    const uint32_t arow = (ky == 0 ? s0 : ky == 1 ? s1 : s2) + kx;
#pragma unroll
    for (int ks = 0; ks < KC / 2; ++ks) {
      const uint32_t a = arow + 2 * ks * T3WP, b = (uint32_t)(tap * KC + 2 * ks) * 64;
      const uint64_t dah = umma::desc_advance(ahl, ahh, a), dal = umma::desc_advance(all_, alh, a);
      const uint64_t db = umma::desc_advance(bl, bh, b);
      umma::mma_bf16(taddr, dah, db, idesc64, (tap | ks) != 0 ? 1u : 0u);
      umma::mma_bf16(taddr, dal, db, idesc32, 1u);
    }
  }
}

// CRFP_PREC_HALF: the activation operand is ONE fp16 tile.  N-concatenated weights: a single N = 64 MMA per K step;
// otherwise A x W_hi + A x W_lo.
template <int KC>
__device__ __forceinline__ void tc3_issue_row_half_cat(uint64_t dA, uint64_t dB, uint32_t s0, uint32_t s1, uint32_t s2,
                                                       uint32_t idesc64, uint32_t taddr) {
  const uint32_t al = (uint32_t)dA, ah = (uint32_t)(dA >> 32), bl = (uint32_t)dB, bh = (uint32_t)(dB >> 32);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap % 3;
    const uint32_t arow = (ky == 0 ? s0 : ky == 1 ? s1 : s2) + kx;
#pragma unroll
    for (int ks = 0; ks < KC / 2; ++ks)
      umma::mma_bf16(taddr, umma::desc_advance(al, ah, arow + 2 * ks * T3WP), umma::desc_advance(bl, bh, (uint32_t)(tap * KC + 2 * ks) * 64),
                     idesc64, (tap | ks) != 0 ? 1u : 0u);
  }
}
__device__ __noinline__ void tc3_issue_row_half_generic(uint64_t dA, uint64_t dBh, uint64_t dBl, uint32_t s0, uint32_t s1, uint32_t s2,
                                                        uint32_t NT, int KC, uint32_t idesc, uint32_t taddr) {
  const uint32_t al = (uint32_t)dA, ah = (uint32_t)(dA >> 32);
  const uint32_t bhl = (uint32_t)dBh, bhh = (uint32_t)(dBh >> 32), bll = (uint32_t)dBl, blh = (uint32_t)(dBl >> 32);
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap - ky * 3;
    uint32_t a = (ky == 0 ? s0 : ky == 1 ? s1 : s2) + kx, b = (uint32_t)(tap * KC) * NT;
    for (int ks = 0; ks < KC / 2; ++ks, a += 2 * T3WP, b += 2 * NT) {
      const uint64_t da = umma::desc_advance(al, ah, a);
      umma::mma_bf16(taddr, da, umma::desc_advance(bhl, bhh, b), idesc, (tap | ks) != 0 ? 1u : 0u);
      umma::mma_bf16(taddr, da, umma::desc_advance(bll, blh, b), idesc, 1u);
    }
  }
}

__device__ __noinline__ void tc3_issue_row_generic(uint64_t dAh, uint64_t dAl, uint64_t dBh, uint64_t dBl, uint32_t s0,
                                                   uint32_t s1, uint32_t s2, uint32_t NT, int KC, uint32_t idesc, uint32_t taddr, uint32_t leader) {
  const uint32_t ahl = (uint32_t)dAh, ahh = (uint32_t)(dAh >> 32), all_ = (uint32_t)dAl, alh = (uint32_t)(dAl >> 32);
  const uint32_t bhl = (uint32_t)dBh, bhh = (uint32_t)(dBh >> 32), bll = (uint32_t)dBl, blh = (uint32_t)(dBl >> 32);
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int ky = tap / 3, kx = tap - ky * 3;
    uint32_t a = (ky == 0 ? s0 : ky == 1 ? s1 : s2) + kx, b = (uint32_t)(tap * KC) * NT;
    for (int ks = 0; ks < KC / 2; ++ks, a += 2 * T3WP, b += 2 * NT) {
      const uint64_t dah = umma::desc_advance(ahl, ahh, a), dal = umma::desc_advance(all_, alh, a);
      const uint64_t dbh = umma::desc_advance(bhl, bhh, b), dbl = umma::desc_advance(bll, blh, b);
      umma::mma_bf16(taddr, dah, dbh, idesc, (tap | ks) != 0 ? 1u : 0u);
      umma::mma_bf16(taddr, dal, dbh, idesc, 1u);
      umma::mma_bf16(taddr, dah, dbl, idesc, 1u);
    }
  }
}

// Epilogue of one 32-channel chunk of one pixel: bias, 2-channel extra source, activation / DCN head, residual,
// post-scale, store (fp32 NHWC segments or pixel shuffle).  Shared by both kernel variants.
__device__ __forceinline__ void tc3_epilogue_chunk(const Tc3Params& P, float* v, const float* sBiasC, const float* sWxC,
                                                   const float2* ex, float2 fl, int NT, int cbase, int nvalid, int n, int y,
                                                   int x, size_t pix, const float4* rpre = nullptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += sBiasC[i];
      if (sWxC != nullptr) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float4* wx0 = reinterpret_cast<const float4*>(sWxC + (tap * 2) * NT);
          const float4* wx1 = reinterpret_cast<const float4*>(sWxC + (tap * 2 + 1) * NT);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 a = wx0[i4], b = wx1[i4];
            v[4 * i4 + 0] = fmaf(ex[tap].x, a.x, fmaf(ex[tap].y, b.x, v[4 * i4 + 0]));
            v[4 * i4 + 1] = fmaf(ex[tap].x, a.y, fmaf(ex[tap].y, b.y, v[4 * i4 + 1]));
            v[4 * i4 + 2] = fmaf(ex[tap].x, a.z, fmaf(ex[tap].y, b.z, v[4 * i4 + 2]));
            v[4 * i4 + 3] = fmaf(ex[tap].x, a.w, fmaf(ex[tap].y, b.w, v[4 * i4 + 3]));
          }
        }
      }
      if (P.res_pre && P.residual != nullptr) {   // K-split pass 2..: the earlier passes' partial sums, BEFORE the activation
        const float* rp = P.residual + pix * P.res_cstride + P.res_coffset + cbase;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (4 * j >= nvalid) break;
          const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + 4 * j));
          v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w;
        }
      }
      if (P.act == CRFP_ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lrelu01(v[i]);
      } else if (P.act == CRFP_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (P.act == CRFP_ACT_DCN_HEAD) {
        // offsets: mag * tanh(v) + flow = (mag + flow) - 2 mag / (exp(2v) + 1);  masks: sigmoid(v) = 1 - 1 / (exp(v) + 1).
        // Raw ex2 / rcp MUFU ops (5 instructions per channel, no branches, 32 independent chains per chunk); chunks
        // never straddle the offset / mask boundary in the fused head layout (144 | 72, chunk bases multiples of 16).
        if (cbase + 32 <= P.head_split && !(cbase & 1)) {
          const float m2 = -2.f * P.head_mag, by = P.head_mag + fl.y, bx = P.head_mag + fl.x;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = fmaf(m2, rcp_approx(ex2_approx(v[i] * 2.885390081777927f) + 1.f), (i & 1) ? bx : by);
        } else if (cbase >= P.head_split) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 1.f - rcp_approx(ex2_approx(v[i] * 1.4426950408889634f) + 1.f);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int cc = cbase + i;
            const bool off = cc < P.head_split;
            const float k = off ? 2.f : 1.f;
            const float t = 1.f - k * rcp_approx(ex2_approx(k * 1.4426950408889634f * v[i]) + 1.f);
            v[i] = off ? fmaf(P.head_mag, t, (cc & 1) ? fl.x : fl.y) : t;
          }
        }
      }
      if (P.res_pre) {
      } else if (rpre != nullptr) {   // residual of this chunk already in registers (loaded while the MMAs were running)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j] += rpre[j].x; v[4 * j + 1] += rpre[j].y; v[4 * j + 2] += rpre[j].z; v[4 * j + 3] += rpre[j].w;
        }
      } else if (P.residual != nullptr) {
        const float* rp = P.residual + pix * P.res_cstride + P.res_coffset + cbase;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (4 * j >= nvalid) break;
          const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + 4 * j));
          v[4 * j] += rv.x; v[4 * j + 1] += rv.y; v[4 * j + 2] += rv.z; v[4 * j + 3] += rv.w;
        }
      }
      if (P.post_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= P.post_scale;
      }
      if (P.out_kind == TC_OUT_F32 && (P.ndst == 1 || cbase >= P.dst_c[0] || cbase + nvalid <= P.dst_c[0])) {
        // the whole chunk lands in one destination segment (always, for this network's layouts): one pointer
        // computation, then stores at constant offsets
        const bool s1 = P.ndst > 1 && cbase >= P.dst_c[0];
        float* op = reinterpret_cast<float*>(s1 ? P.dst[1] : P.dst[0]) + pix * (size_t)(s1 ? P.dst_cstride[1] : P.dst_cstride[0]) +
                    (s1 ? P.dst_coffset[1] + cbase - P.dst_c[0] : P.dst_coffset[0] + cbase);
        if (nvalid == 32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < nvalid) *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else if (P.out_kind == TC_OUT_F32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int cc = cbase + 4 * j;
          if (4 * j >= nvalid) break;
          int seg = 0, cl = cc;
          if (P.ndst > 1 && cc >= P.dst_c[0]) { seg = 1; cl = cc - P.dst_c[0]; }
          float* op = reinterpret_cast<float*>(P.dst[seg]) + pix * P.dst_cstride[seg] + P.dst_coffset[seg] + cl;
          *reinterpret_cast<float4*>(op) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else if (P.shuffle_r == 2 && nvalid == 32 && !((P.dst_cstride[0] | P.dst_coffset[0]) & 3)) {
        // PixelShufflePack x2 (upsample: 32 -> 96 @LR -> 24 channels @L1): conv channel o*4 + dy*2 + dx -> output channel o
        // of sub-pixel (dy, dx); a 32-channel chunk is 8 consecutive output channels of each of the 4 sub-pixels = two
        // 16-byte stores per sub-pixel instead of 32 scattered 4-byte stores
        const int Wo = P.w * 2;
        float* ob = reinterpret_cast<float*>(P.dst[0]) + P.dst_coffset[0] + (cbase >> 2);
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const size_t opix = ((size_t)n * (P.h * 2) + (y * 2 + (sub >> 1))) * (size_t)Wo + (x * 2 + (sub & 1));
          float4* op = reinterpret_cast<float4*>(ob + opix * P.dst_cstride[0]);
          op[0] = make_float4(v[sub], v[4 + sub], v[8 + sub], v[12 + sub]);
          op[1] = make_float4(v[16 + sub], v[20 + sub], v[24 + sub], v[28 + sub]);
        }
      } else {  // TC_OUT_SHUFFLE_F32
        const int r_ = P.shuffle_r, rr = r_ * r_;
        const int Wo = P.w * r_;
        float* ob = reinterpret_cast<float*>(P.dst[0]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int cc = cbase + i;
          if (i >= nvalid) break;
          const int o = cc / rr, sub = cc - o * rr;
          const int dy = sub / r_, dx = sub - dy * r_;
          const size_t opix = ((size_t)n * (P.h * r_) + (y * r_ + dy)) * (size_t)Wo + (x * r_ + dx);
          ob[opix * P.dst_cstride[0] + P.dst_coffset[0] + o] = v[i];
        }
      }
}

__global__ void __launch_bounds__(128) conv_tc3_kernel(const Tc3Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int KC = P.kc_total, NT = P.nt;
  const int wrecs = 9 * KC * NT;          // records per weight half
  const int slot_recs = KC * T3WP;        // records per ring slot half
  const long long t_start = clock64();
  pdl_trigger();
  uint4* sWh = reinterpret_cast<uint4*>(smem);
  uint4* sWl = sWh + wrecs;
  uint4* sAh = sWl + wrecs;               // [3 slots][KC][130]
  uint4* sAl = sAh + 3 * slot_recs;
  float4* sStage = reinterpret_cast<float4*>(sAl + 3 * slot_recs);   // [kc_real][130][2]
  float* sBias = reinterpret_cast<float*>(sStage + (2 * P.kc_real + 1) * T3WP);
  float* sWx = sBias + ((NT + 31) & ~31);   // extra 2-channel source weights [9][2][NT] (optional); bias padded to 32
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (keeps the MMA issue path uniform)
  const int cotile = blockIdx.z % P.ntiles, n = blockIdx.z / P.ntiles;
  const int x0 = blockIdx.x * T3M;
  const int y_begin = blockIdx.y * P.rows_per_cta;
  const int y_end = min(P.h, y_begin + P.rows_per_cta);

  {
    const uint4* gwh = reinterpret_cast<const uint4*>(P.weight_hi) + (size_t)cotile * wrecs;
    const uint4* gwl = reinterpret_cast<const uint4*>(P.weight_lo) + (size_t)cotile * wrecs;
    for (int i = tid; i < wrecs; i += 128) {
      umma::cp_async16(sWh + i, gwh + i, 16u);
      umma::cp_async16(sWl + i, gwl + i, 16u);
    }
    umma::cp_async_commit();
    for (int i = tid; i < ((NT + 31) & ~31); i += 128) sBias[i] = (i < NT) ? P.bias[cotile * NT + i] : 0.f;
    if (P.extra != nullptr)
      for (int i = tid; i < 18 * NT; i += 128) sWx[i] = P.w_extra[(size_t)(i / NT) * (P.ntiles * NT) + cotile * NT + (i % NT)];
    for (int kc = P.kc_real; kc < KC; ++kc)  // K padding chunk: zero in every slot, both halves
      for (int i = tid; i < 3 * T3WP; i += 128) {
        sAh[(i / T3WP) * slot_recs + kc * T3WP + (i % T3WP)] = make_uint4(0u, 0u, 0u, 0u);
        sAl[(i / T3WP) * slot_recs + kc * T3WP + (i % T3WP)] = make_uint4(0u, 0u, 0u, 0u);
      }
  }
  uint32_t ncols = 32;
  while ((int)ncols < NT) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  pdl_wait();
  // prologue: rows y_begin-1 and y_begin go through the staging buffer one after the other
  for (int r = -1; r <= 0; ++r) {
    const int yy = y_begin + r, sl = (yy + 3) % 3;
    tc3_stage_row(P, sStage, n, yy, x0, tid);
    umma::cp_async_commit();
    umma::cp_async_wait<0>();
    __syncthreads();
    tc3_convert_row(P, sStage, sAh + sl * slot_recs, sAl + sl * slot_recs, n, yy, x0, tid);
    __syncthreads();
  }
  tc3_stage_row(P, sStage, n, y_begin + 1, x0, tid);
  umma::cp_async_commit();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_bf16(T3M, NT);
  // base descriptors (record 0 of each operand region); the issue loop only advances their start addresses
  const uint64_t dAh = umma::make_desc(umma::smem_u32(sAh), T3WP * 16, 128), dAl = umma::make_desc(umma::smem_u32(sAl), T3WP * 16, 128);
  const uint64_t dBh = umma::make_desc(umma::smem_u32(sWh), (uint32_t)NT * 16, 128), dBl = umma::make_desc(umma::smem_u32(sWl), (uint32_t)NT * 16, 128);
  uint32_t phase = 0;
  const int x = x0 + tid;
  const bool xvalid = x < P.w;

  const bool trace = (P.dbg != nullptr) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
  if (trace) { P.dbg[0] = t_start; P.dbg[1] = clock64(); }
  for (int y = y_begin; y < y_end; ++y) {
    long long* tr = trace ? P.dbg + 8 * (y - y_begin + 1) : nullptr;
    if (tr) tr[0] = clock64();
    umma::cp_async_wait<0>();
    __syncthreads();  // staged row y+1 is visible to everyone; MMA(y-1) has been waited for by all threads
    if (tr) tr[1] = clock64();
    {
      const int sl = (y + 1 + 3) % 3;
      tc3_convert_row(P, sStage, sAh + sl * slot_recs, sAl + sl * slot_recs, n, y + 1, x0, tid);
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tr) tr[2] = clock64();
    if (y + 1 < y_end) tc3_stage_row(P, sStage, n, y + 2, x0, tid);
    umma::cp_async_commit();
    if (tr) tr[3] = clock64();
    if (warp == 0 && umma::elect_one()) {   // warp-uniform branch + elect.sync: one thread, uniform datapath
      umma::fence_after_sync();
      const uint32_t leader = 1u;
      const uint32_t s0 = (uint32_t)(((y - 1 + 3) % 3) * slot_recs), s1 = (uint32_t)(((y + 3) % 3) * slot_recs),
                     s2 = (uint32_t)(((y + 1 + 3) % 3) * slot_recs);
      if (KC == 4)
        tc3_issue_row<4>(dAh, dAl, dBh, dBl, s0, s1, s2, (uint32_t)NT, idesc, taddr, leader);
      else if (KC == 8)
        tc3_issue_row<8>(dAh, dAl, dBh, dBl, s0, s1, s2, (uint32_t)NT, idesc, taddr, leader);
      else
        tc3_issue_row_generic(dAh, dAl, dBh, dBl, s0, s1, s2, (uint32_t)NT, KC, idesc, taddr, leader);
      umma::mma_commit(&bar);
    }
    // work that does not need the accumulator: flow at this pixel, the 2-channel extra source taps
    const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
    float2 fl = make_float2(0.f, 0.f);
    if (P.act == CRFP_ACT_DCN_HEAD && xvalid) fl = __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2));
    float2 ex[9];
    if (P.extra != nullptr) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        ex[tap] = (xvalid && yy >= 0 && yy < P.h && xx >= 0 && xx < P.w)
                      ? __ldg(reinterpret_cast<const float2*>(P.extra + (((size_t)n * P.h + yy) * (size_t)P.w + xx) * 2))
                      : make_float2(0.f, 0.f);
      }
    }
    if (tr) tr[4] = clock64();
    umma::mbar_wait(&bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    if (tr) tr[5] = clock64();

    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
      umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
      const int cbase = cotile * NT + c0;
      if (!xvalid || cbase >= P.cout) continue;
      const int nvalid = min(min(32, NT - c0), P.cout - cbase);
      tc3_epilogue_chunk(P, v, sBias + c0, P.extra != nullptr ? sWx + c0 : nullptr, ex, fl, NT, cbase, nvalid, n, y, x, pix);
    }
    umma::fence_before_sync();
    if (tr) tr[6] = clock64();
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, ncols);
}


// =====================================================================================================================
// Warp-specialised, software-pipelined variant (default): 384 threads = 4 epilogue warps (TMEM lanes 0..127) + 2 MMA
// warps (even / odd output rows) + 6 producer warps.  Producers load fp32 rows straight from global (coalesced
// LDG.128, one or more rows ahead in registers), split them hi/lo and store them into a 4-slot ring of UMMA operand
// rows; an MMA warp issues its output row's 9 x KC/2 x 3 tcgen05.mma into its TMEM accumulator; the epilogue warps
// drain the other accumulator.  Stages meet only through
// mbarriers: full[slot] (producers -> MMA), acc_full[buf] (tcgen05.commit -> epilogue and, as "rows <= v are
// consumed", -> producers), acc_empty[buf] (epilogue -> MMA).  All three stages of consecutive rows overlap.
constexpr int WS_EPI = 128, WS_NPROD = 192, WS_THREADS = 384, WS_SLOTS = 4;

// Producer loop with register prefetch.  Every producer thread owns up to RPT fixed records of a row (record = 8
// channels of one pixel of one source; same (source, pixel, channel group) for every row, so all addressing is hoisted
// out of the row loop).  The loads of row u+D are issued before row u is split and stored: D+1 rows of HBM latency
// overlap with the conversion and with the wait for a free ring slot.
// CRFP_PREC_HALF: 8 channels -> one record of fp16 (no lo half)
__device__ __forceinline__ uint4 cvt8_f16(const float4 a, const float4 b) {
  return make_uint4(umma::pack_f16(a.x, a.y), umma::pack_f16(a.z, a.w), umma::pack_f16(b.x, b.y), umma::pack_f16(b.z, b.w));
}

template <int RPT, int D, bool HALF>
__device__ __forceinline__ void ws_producer_loop(const Tc3Params& P, uint4* sAh, uint4* sAl, int slot_recs, int n, int y_begin,
                                                 int rows_out, int x0, int ptid, uint64_t* full_bar, uint64_t* accf_bar,
                                                 long long* tr) {
  const float* rp[RPT];
  int rstride[RPT], dst[RPT];
  const int nrec_total = T3WP * P.kc_real;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int id = ptid + k * WS_NPROD;
    rp[k] = nullptr; rstride[k] = 0; dst[k] = -1;
    if (id < nrec_total) {
      int s = 0, local = id;
      while (s + 1 < P.nsrc && local >= T3WP * (P.src_c[s] >> 3)) { local -= T3WP * (P.src_c[s] >> 3); ++s; }
      const int cps = P.src_c[s] >> 3;
      const int px = local / cps, j = local - px * cps;
      const int x = x0 + px - 1;
      dst[k] = (P.kstart[s] + j) * T3WP + px;
      if (x >= 0 && x < P.w) {
        if (P.src_mode[s] == CRFP_SRC_UNSHUFFLE4) {
          // pixel_unshuffle(4) of a dense 4-channel HR plane: record j = HR row 4y + j/2, HR pixels 4x + 2*(j%2), +1
          rp[k] = P.src[s] + ((size_t)n * (P.h * 4) * (size_t)(P.w * 4) + (size_t)(j >> 1) * (P.w * 4) + (size_t)x * 4 + (j & 1) * 2) * 4;
          rstride[k] = 4 * (P.w * 4) * 4;
        } else {
          rp[k] = P.src[s] + ((size_t)n * P.h * P.w + x) * (size_t)P.src_cstride[s] + P.src_coffset[s] + j * 8;
          rstride[k] = P.w * P.src_cstride[s];
        }
      }
    }
  }
  float4 ra[D + 1][RPT], rb[D + 1][RPT];
  float rf[D + 1][RPT];
  auto load_row = [&](int y, float4* a, float4* b, float* f) {
    const bool yin = (y >= 0 && y < P.h);
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      f[k] = 1.f;
      if (yin && rp[k] != nullptr) {
        const float4* g = reinterpret_cast<const float4*>(rp[k] + (long long)y * rstride[k]);
        a[k] = __ldg(g);
        b[k] = __ldg(g + 1);
        if (P.fg != nullptr) f[k] = __ldg(P.fg + (size_t)n * P.fg_clip_stride + (size_t)y * P.w + (x0 + dst[k] % T3WP - 1));
      }
    }
  };
  const int nrows = rows_out + 2;
#pragma unroll
  for (int d = 0; d < D; ++d)
    if (d < nrows) load_row(y_begin - 1 + d, ra[d], rb[d], rf[d]);
  for (int u = 0; u < nrows; ++u) {
    const int slot = u & (WS_SLOTS - 1);
    if (u + D < nrows) load_row(y_begin - 1 + u + D, ra[D], rb[D], rf[D]);
    if (tr && u < 60) tr[u * 4 + 0] = clock64();
    if (u >= WS_SLOTS) umma::mbar_wait_safe(&accf_bar[(u - 4) & 1], (uint32_t)(((u - 4) >> 1) & 1));
    if (tr && u < 60) tr[u * 4 + 1] = clock64();
    uint4* hi = sAh + slot * slot_recs;
    uint4* lo = sAl + slot * slot_recs;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (dst[k] < 0) continue;
      float4 a = ra[0][k], b = rb[0][k];
      if (P.fg != nullptr) {
        const float f = rf[0][k];
        a.x *= f; a.y *= f; a.z *= f; a.w *= f;
        b.x *= f; b.y *= f; b.z *= f; b.w *= f;
      }
      if (HALF) {
        hi[dst[k]] = cvt8_f16(a, b);
      } else {
        uint4 h, l;
        split8(a, b, h, l);
        hi[dst[k]] = h;
        lo[dst[k]] = l;
      }
    }
    umma::fence_proxy_async();
    umma::mbar_arrive(&full_bar[slot]);
    if (tr && u < 60) tr[u * 4 + 2] = clock64();
#pragma unroll
    for (int d = 0; d < D; ++d) {
#pragma unroll
      for (int k = 0; k < RPT; ++k) { ra[d][k] = ra[d + 1][k]; rb[d][k] = rb[d + 1][k]; rf[d][k] = rf[d + 1][k]; }
    }
  }
}

template <bool HALF>
__global__ void __launch_bounds__(WS_THREADS, 1) conv_tc3_ws_kernel(const Tc3Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[WS_SLOTS], accf_bar[2], acce_bar[2];
  __shared__ uint32_t tmem_base_s;
  const int KC = P.kc_total, NT = P.nt;
  const int wrecs = 9 * KC * NT;
  const int slot_recs = KC * T3WP;
  uint4* sWh = reinterpret_cast<uint4*>(smem);
  uint4* sWl = sWh + wrecs;
  uint4* sAh = sWl + wrecs;                 // [4 slots][KC][130]
  uint4* sAl = sAh + WS_SLOTS * slot_recs;
  float* sBias = reinterpret_cast<float*>(sAl + WS_SLOTS * slot_recs);
  float* sWx = sBias + ((NT + 31) & ~31);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int cotile = blockIdx.z % P.ntiles, n = blockIdx.z / P.ntiles;
  const int x0 = blockIdx.x * T3M;
  const int y_begin = blockIdx.y * P.rows_per_cta;
  const int y_end = min(P.h, y_begin + P.rows_per_cta);
  const int rows_out = y_end - y_begin;
  // profiling aid (crfp_conv3x3_tc3_trace): clock64 stamps of one interior CTA, per role and row
  const bool trace_cta = P.dbg != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2 && blockIdx.z == 0;
  if (trace_cta && tid == 0) P.dbg[0] = clock64();

  pdl_trigger();   // constant-only prologue below overlaps the previous kernel's tail (PDL)
  {
    const uint4* gwh = reinterpret_cast<const uint4*>(P.weight_hi) + (size_t)cotile * wrecs;
    const uint4* gwl = reinterpret_cast<const uint4*>(P.weight_lo) + (size_t)cotile * wrecs;
    if (P.ncat) {   // [tap][kc][hi NT | lo NT] (NT = 32)
      for (int i = tid; i < wrecs; i += WS_THREADS) {
        const int blk = i / NT, r = i - blk * NT;
        umma::cp_async16(sWh + blk * 2 * NT + r, gwh + i, 16u);
        umma::cp_async16(sWh + blk * 2 * NT + NT + r, gwl + i, 16u);
      }
    } else {
      for (int i = tid; i < wrecs; i += WS_THREADS) {
        umma::cp_async16(sWh + i, gwh + i, 16u);
        umma::cp_async16(sWl + i, gwl + i, 16u);
      }
    }
    umma::cp_async_commit();
    for (int i = tid; i < ((NT + 31) & ~31); i += WS_THREADS) sBias[i] = (i < NT) ? P.bias[cotile * NT + i] : 0.f;
    if (P.extra != nullptr)
      for (int i = tid; i < 18 * NT; i += WS_THREADS) sWx[i] = P.w_extra[(size_t)(i / NT) * (P.ntiles * NT) + cotile * NT + (i % NT)];
    for (int kc = P.kc_real; kc < KC; ++kc)
      for (int i = tid; i < WS_SLOTS * T3WP; i += WS_THREADS) {
        sAh[(i / T3WP) * slot_recs + kc * T3WP + (i % T3WP)] = make_uint4(0u, 0u, 0u, 0u);
        sAl[(i / T3WP) * slot_recs + kc * T3WP + (i % T3WP)] = make_uint4(0u, 0u, 0u, 0u);
      }
    umma::cp_async_wait<0>();
  }
  uint32_t ncols = 32;
  while ((int)ncols < NT * (P.ncat ? 2 : 1)) ncols <<= 1;
  if (warp == 4) umma::tmem_alloc(&tmem_base_s, 2 * ncols);
  if (tid == 0) {
    for (int i = 0; i < WS_SLOTS; ++i) umma::mbar_init(&full_bar[i], WS_NPROD);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&accf_bar[i], 1); umma::mbar_init(&acce_bar[i], WS_EPI); }
    umma::fence_mbar_init();
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  pdl_wait();      // activations (sources, residual, flow, destinations) are only touched from here on
  if (trace_cta && tid == 0) P.dbg[1] = clock64();

  if (warp >= 6) {
    // ------------------------------------------------------------------ producers
    const int ptid = tid - 192;
    long long* tr = (trace_cta && ptid == 0) ? P.dbg + 1024 : nullptr;        // ws trace, role 1: producers
    if (P.kc_real <= 4)
      ws_producer_loop<3, 2, HALF>(P, sAh, sAl, slot_recs, n, y_begin, rows_out, x0, ptid, full_bar, accf_bar, tr);
    else
      ws_producer_loop<6, 0, HALF>(P, sAh, sAl, slot_recs, n, y_begin, rows_out, x0, ptid, full_bar, accf_bar, tr);
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ MMA issuers: warp 4 = even rows (accumulator
    // 0), warp 5 = odd rows (accumulator 1).  While one warp is blocked feeding the tensor pipe, the other one has
    // already passed the barriers of the next row, so the pipe never drains between rows.
    const uint32_t idesc = HALF ? umma::make_idesc_f16(T3M, NT) : umma::make_idesc_bf16(T3M, NT);
    const uint64_t dAh = umma::make_desc(umma::smem_u32(sAh), T3WP * 16, 128), dAl = umma::make_desc(umma::smem_u32(sAl), T3WP * 16, 128);
    const uint64_t dBh = umma::make_desc(umma::smem_u32(sWh), (uint32_t)NT * 16, 128), dBl = umma::make_desc(umma::smem_u32(sWl), (uint32_t)NT * 16, 128);
    long long* tr = (trace_cta && (tid & 31) == 0) ? P.dbg + 2048 : nullptr;  // role 2: MMA issuers
    for (int v = warp - 4; v < rows_out; v += 2) {
      const int u2 = v + 2, b = v & 1;
      if (tr && v < 60) tr[v * 4 + 0] = clock64();
      umma::mbar_wait_safe(&full_bar[v & 3], (uint32_t)((v >> 2) & 1));
      umma::mbar_wait_safe(&full_bar[(v + 1) & 3], (uint32_t)(((v + 1) >> 2) & 1));
      umma::mbar_wait_safe(&full_bar[u2 & 3], (uint32_t)((u2 >> 2) & 1));
      if (tr && v < 60) tr[v * 4 + 1] = clock64();
      umma::mbar_wait_safe(&acce_bar[b], (uint32_t)(((v >> 1) & 1) ^ 1));
      if (tr && v < 60) tr[v * 4 + 2] = clock64();
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t s0 = (uint32_t)((v & 3) * slot_recs), s1 = (uint32_t)(((v + 1) & 3) * slot_recs),
                       s2 = (uint32_t)(((v + 2) & 3) * slot_recs);
        const uint32_t acc = taddr + (uint32_t)b * ncols;
        if (HALF) {
          if (P.ncat && KC == 8)
            tc3_issue_row_half_cat<8>(dAh, umma::make_desc(umma::smem_u32(sWh), 64 * 16, 128), s0, s1, s2, umma::make_idesc_f16(T3M, 64), acc);
          else if (P.ncat)
            tc3_issue_row_half_cat<4>(dAh, umma::make_desc(umma::smem_u32(sWh), 64 * 16, 128), s0, s1, s2, umma::make_idesc_f16(T3M, 64), acc);
          else
            tc3_issue_row_half_generic(dAh, dBh, dBl, s0, s1, s2, (uint32_t)NT, KC, idesc, acc);
        } else if (P.ncat && KC == 8)
          tc3_issue_row_cat<8>(dAh, dAl, umma::make_desc(umma::smem_u32(sWh), 64 * 16, 128), s0, s1, s2,
                               umma::make_idesc_bf16(T3M, 64), idesc, acc);
        else if (P.ncat)
          tc3_issue_row_cat<4>(dAh, dAl, umma::make_desc(umma::smem_u32(sWh), 64 * 16, 128), s0, s1, s2,
                               umma::make_idesc_bf16(T3M, 64), idesc, acc);
        else if (KC == 4)
          tc3_issue_row<4>(dAh, dAl, dBh, dBl, s0, s1, s2, (uint32_t)NT, idesc, acc, 1u);
        else if (KC == 8)
          tc3_issue_row<8>(dAh, dAl, dBh, dBl, s0, s1, s2, (uint32_t)NT, idesc, acc, 1u);
        else
          tc3_issue_row_generic(dAh, dAl, dBh, dBl, s0, s1, s2, (uint32_t)NT, KC, idesc, acc, 1u);
        umma::mma_commit(&accf_bar[b]);
      }
      __syncwarp();
      if (tr && v < 60) tr[v * 4 + 3] = clock64();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (thread = pixel = TMEM lane)
    long long* tr = (trace_cta && tid == 0) ? P.dbg + 3072 : nullptr;         // role 3: epilogue
    const int x = x0 + tid;
    const bool xvalid = x < P.w;
    const bool shuffle4_fast = P.out_kind == TC_OUT_SHUFFLE_F32 && P.shuffle_r == 4 && NT == 64 && P.cout == 64 &&
                               P.dst_cstride[0] == 4 && P.dst_coffset[0] == 0 && P.residual == nullptr &&
                               P.extra == nullptr && (P.act == CRFP_ACT_NONE || P.act == CRFP_ACT_LRELU);
    // Coalesced epilogue for the plain 32-channel tiles (most L1 / LR layers): the accumulator is read in the 16x256b fragment
    // layout, where a quad of threads owns 8 consecutive channels of one pixel, so every store instruction of a warp writes
    // eight full 32-byte sectors (the one-pixel-per-thread layout writes 32 half sectors), and a thread needs only 8 bias
    // values — kept in registers for the whole CTA instead of 8 LDS.128 per row queued behind the UMMA operand reads.
    // (not in the fp16 instantiation: the extra live ranges spill at its 168-register cap — measured 409 vs 435 fps)
    const bool fast16 = !HALF && P.fast16 && NT == 32 && P.out_kind == TC_OUT_F32 && P.extra == nullptr && P.cout % 8 == 0 &&
                        (P.act == CRFP_ACT_NONE || P.act == CRFP_ACT_LRELU || P.act == CRFP_ACT_RELU) &&
                        (P.ndst == 1 || P.dst_c[0] % 8 == 0);
    if (fast16) {
      const int lane = tid & 31, q2 = (lane & 3) * 2, rr = lane >> 2;
      const int cbase = cotile * NT;
      float bias8[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) { bias8[2 * g] = sBias[8 * g + q2]; bias8[2 * g + 1] = sBias[8 * g + q2 + 1]; }
      // per column group (loop invariant): destination pointer of pixel 0 and pixel stride, or NULL past cout
      float* gptr[4];
      int gstr[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int cc = cbase + 8 * g + q2;
        const bool s1 = P.ndst > 1 && cc >= P.dst_c[0];
        gstr[g] = s1 ? P.dst_cstride[1] : P.dst_cstride[0];
        gptr[g] = (cbase + 8 * g < P.cout)
                      ? reinterpret_cast<float*>(s1 ? P.dst[1] : P.dst[0]) + (s1 ? P.dst_coffset[1] + cc - P.dst_c[0] : P.dst_coffset[0] + cc)
                      : nullptr;
      }
      const float* resb = P.residual != nullptr ? P.residual + P.res_coffset + cbase + q2 : nullptr;
      const int px0 = x0 + 32 * warp + rr;
      const bool lrelu = P.act == CRFP_ACT_LRELU, relu = P.act == CRFP_ACT_RELU, scale = P.post_scale != 1.f;
      for (int v = 0; v < rows_out; ++v) {
        const int y = y_begin + v, b = v & 1;
        const size_t rowpix = ((size_t)n * P.h + y) * (size_t)P.w;
        float2 rp[4][4];
        if (resb != nullptr) {   // residual (or K-split partial sums) fetched while the MMAs of this row run
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int px = px0 + 8 * h;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              rp[h][g] = (px < P.w && gptr[g] != nullptr)
                             ? __ldg(reinterpret_cast<const float2*>(resb + (rowpix + px) * P.res_cstride + 8 * g))
                             : make_float2(0.f, 0.f);
          }
        }
        if (tr && v < 60) tr[v * 4 + 0] = clock64();
        umma::mbar_wait_safe(&accf_bar[b], (uint32_t)((v >> 1) & 1));
        if (tr && v < 60) tr[v * 4 + 1] = clock64();
        umma::fence_after_sync();
        float a0[16], a1[16];
        const uint32_t tb = taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)b * ncols;
        umma::tmem_ld16x256b_x4(tb, a0);
        umma::tmem_ld16x256b_x4(tb + (16u << 16), a1);
        if (P.ncat) {   // columns 32..63 hold A_hi x W_lo
          float c0[16], c1[16];
          umma::tmem_ld16x256b_x4(tb + 32u, c0);
          umma::tmem_ld16x256b_x4(tb + (16u << 16) + 32u, c1);
          umma::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) { a0[i] += c0[i]; a1[i] += c1[i]; }
        } else {
          umma::tmem_ld_wait();
        }
        if (tr && v < 60) tr[v * 4 + 3] = clock64();
        umma::fence_before_sync();
        umma::mbar_arrive(&acce_bar[b]);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const int px = px0 + 8 * h;
          if (px >= P.w) continue;
          const size_t pix = rowpix + px;
          const float* av = (h < 2) ? a0 : a1;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (gptr[g] == nullptr) continue;
            float v0 = av[4 * g + 2 * (h & 1)] + bias8[2 * g], v1 = av[4 * g + 2 * (h & 1) + 1] + bias8[2 * g + 1];
            if (P.res_pre && resb != nullptr) { v0 += rp[h][g].x; v1 += rp[h][g].y; }
            if (lrelu) { v0 = lrelu01(v0); v1 = lrelu01(v1); }
            else if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
            if (!P.res_pre && resb != nullptr) { v0 += rp[h][g].x; v1 += rp[h][g].y; }
            if (scale) { v0 *= P.post_scale; v1 *= P.post_scale; }
            *reinterpret_cast<float2*>(gptr[g] + pix * (size_t)gstr[g]) = make_float2(v0, v1);
          }
        }
        if (tr && v < 60) tr[v * 4 + 2] = clock64();
      }
    } else
    for (int v = 0; v < rows_out; ++v) {
      const int y = y_begin + v, b = v & 1;
      const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
      float2 fl = make_float2(0.f, 0.f);
      if (P.act == CRFP_ACT_DCN_HEAD && xvalid) fl = __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2));
      float2 ex[9];
      if (P.extra != nullptr) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
          ex[tap] = (xvalid && yy >= 0 && yy < P.h && xx >= 0 && xx < P.w)
                        ? __ldg(reinterpret_cast<const float2*>(P.extra + (((size_t)n * P.h + yy) * (size_t)P.w + xx) * 2))
                        : make_float2(0.f, 0.f);
        }
      }
      float4 rpre[8];
      const bool use_rpre = P.residual != nullptr && !P.res_pre && xvalid && NT >= 32 && cotile * NT + 32 <= P.cout;
      if (use_rpre) {   // the residual of the first 32-channel chunk is fetched while the MMAs of this row run
        const float* rpp = P.residual + pix * P.res_cstride + P.res_coffset + cotile * NT;
#pragma unroll
        for (int j = 0; j < 8; ++j) rpre[j] = __ldg(reinterpret_cast<const float4*>(rpp + 4 * j));
      }
      if (tr && v < 60) tr[v * 4 + 0] = clock64();
      umma::mbar_wait_safe(&accf_bar[b], (uint32_t)((v >> 1) & 1));
      if (tr && v < 60) tr[v * 4 + 1] = clock64();
      umma::fence_after_sync();
      if (shuffle4_fast) {
        // PixelShufflePack x4 with 64 conv channels -> 4-channel HR plane: hold all 64 values, then every (dy) row of
        // the 4x4 sub-pixel block is 4 consecutive float4 = 64 contiguous bytes per thread, 2 KB per warp.
        float v0[32], v1[32];
        umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)b * ncols, v0);
        umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)b * ncols + 32u, v1);
        umma::fence_before_sync();
        umma::mbar_arrive(&acce_bar[b]);
        if (!xvalid) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v0[i] += sBias[i];
          v1[i] += sBias[32 + i];
          if (P.act == CRFP_ACT_LRELU) { v0[i] = lrelu01(v0[i]); v1[i] = lrelu01(v1[i]); }
          v0[i] *= P.post_scale;
          v1[i] *= P.post_scale;
        }
        float* ob = reinterpret_cast<float*>(P.dst[0]);
        const int Wo = P.w * 4;
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
          float4* op = reinterpret_cast<float4*>(ob + (((size_t)n * (P.h * 4) + (y * 4 + dy)) * (size_t)Wo + (size_t)x * 4) * 4);
#pragma unroll
          for (int dx = 0; dx < 4; ++dx) op[dx] = make_float4(v0[dy * 4 + dx], v0[16 + dy * 4 + dx], v1[dy * 4 + dx], v1[16 + dy * 4 + dx]);
        }
        continue;
      }
      for (int c0 = 0; c0 < NT; c0 += 32) {
        float vv[32];
        umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)b * ncols + (uint32_t)c0, vv);
        if (P.ncat) {   // columns 32..63 hold A_hi x W_lo
          float v2[32];
          umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)b * ncols + 32u, v2);
#pragma unroll
          for (int i = 0; i < 32; ++i) vv[i] += v2[i];
        }
        if (tr && v < 60 && c0 == 0) tr[v * 4 + 3] = clock64();
        if (c0 + 32 >= NT) {  // last chunk is in registers: the accumulator can be overwritten
          umma::fence_before_sync();
          umma::mbar_arrive(&acce_bar[b]);
        }
        const int cbase = cotile * NT + c0;
        if (!xvalid || cbase >= P.cout) continue;
        const int nvalid = min(min(32, NT - c0), P.cout - cbase);
        tc3_epilogue_chunk(P, vv, sBias + c0, P.extra != nullptr ? sWx + c0 : nullptr, ex, fl, NT, cbase, nvalid, n, y, x, pix,
                           (c0 == 0 && use_rpre) ? rpre : nullptr);
      }
      if (tr && v < 60) tr[v * 4 + 2] = clock64();
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 4) umma::tmem_dealloc(taddr, 2 * ncols);
}

static size_t tc3_smem_bytes(int kc_real, int kc_total, int nt, bool extra) {
  return (size_t)(2 * 9 * kc_total * nt + 6 * kc_total * T3WP) * 16 + (size_t)(2 * kc_real + 1) * T3WP * 16 +
         (size_t)((nt + 31) & ~31) * 4 + (extra ? (size_t)(18 * nt + 32) * 4 : 0);
}

// cout tile for the split kernel: weights are resident twice (hi, lo), so tiles are smaller than conv_tc's
// (also bounded by shared memory: the tile count grows until weights + rings + staging fit in 225 KB)
int tc3_cout_tile(int cout, int kc_real, int* nt, int* ntiles) {
  const int kc_total = (kc_real + 1) & ~1;
  for (int tiles = (cout + 111) / 112; tiles <= cout; ++tiles) {
    const int per = (cout + tiles - 1) / tiles;
    int n = (per + 15) & ~15;
    if (n < 16) n = 16;
    if (tc3_smem_bytes(kc_real, kc_total, n, true) <= (size_t)225 * 1024) {
      *nt = n; *ntiles = tiles;
      return CRFP_OK;
    }
    if (n == 16) break;
  }
  return CRFP_ERR_UNSUPPORTED;
}

int launch_conv_tc3(Tc3Params p, cudaStream_t st) {
  if (p.nsrc < 1 || p.nsrc > 3) return CRFP_ERR_BAD_SHAPE;
  int kc = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    if (!p.src[s]) return CRFP_ERR_NULL;
    if (p.src_c[s] % 8 || p.src_cstride[s] % 4 || p.src_coffset[s] % 4 || ((uintptr_t)p.src[s] & 15)) return CRFP_ERR_BAD_SHAPE;
    if (p.src_mode[s] == CRFP_SRC_UNSHUFFLE4 && (p.src_c[s] != 64 || p.src_cstride[s] != 4 || p.src_coffset[s] != 0)) return CRFP_ERR_UNSUPPORTED;
    p.kstart[s] = kc;
    kc += p.src_c[s] / 8;
  }
  p.kc_real = kc;
  p.kc_total = (kc + 1) & ~1;
  CRFP_TRY(tc3_cout_tile(p.cout, p.kc_real, &p.nt, &p.ntiles));
  if (!p.weight_hi || !p.weight_lo || !p.bias || !p.dst[0]) return CRFP_ERR_NULL;
  if (p.extra && !p.w_extra) return CRFP_ERR_NULL;
  if (p.out_kind != TC_OUT_F32 && p.out_kind != TC_OUT_SHUFFLE_F32) return CRFP_ERR_UNSUPPORTED;
  if (p.out_kind == TC_OUT_F32)
    for (int s = 0; s < p.ndst; ++s)
      if (p.dst_cstride[s] % 4 || p.dst_coffset[s] % 4 || (s == 0 && p.ndst > 1 && p.dst_c[0] % 4)) return CRFP_ERR_BAD_SHAPE;
  if (p.residual && (p.res_cstride % 4 || p.res_coffset % 4)) return CRFP_ERR_BAD_SHAPE;
  if (p.post_scale == 0.f) p.post_scale = 1.f;
  const int strips = ceil_div(p.w, T3M);
  const int per_seg = strips * p.n * p.ntiles;
  static const bool use_v1_env = (getenv("CRFP_TC3_V1") != nullptr);   // A/B switch: the non-specialised kernel
  const bool use_v1 = use_v1_env && !p.half;                            // (it has no CRFP_PREC_HALF variant)
  for (int s = 0; s < p.nsrc; ++s)
    if (use_v1 && p.src_mode[s] != CRFP_SRC_PLAIN) return CRFP_ERR_UNSUPPORTED;
  if (!use_v1) {
    // warp-specialised pipeline: one CTA per SM, one wave
    const size_t smem = (size_t)(2 * 9 * p.kc_total * p.nt + 2 * WS_SLOTS * p.kc_total * T3WP) * 16 +
                        (size_t)((p.nt + 31) & ~31) * 4 + (p.extra ? (size_t)(18 * p.nt + 32) * 4 : 0);
    if (smem > 227 * 1024) return CRFP_ERR_UNSUPPORTED;
    static const bool no_cat = (getenv("CRFP_TC3_NOCAT") != nullptr);
    static const bool cat4 = (getenv("CRFP_TC3_NOCAT4") == nullptr);   // also for the 32-channel layers (+0.5 %)
    static const bool fast16_env = (getenv("CRFP_TC3_NOFAST16") == nullptr);   // A/B: one-pixel-per-thread epilogue
    p.fast16 = fast16_env ? 1 : 0;
    p.ncat = (!no_cat && (p.kc_total == 8 || (cat4 && p.kc_total == 4)) && p.nt == 32 && p.out_kind == TC_OUT_F32) ? 1 : 0;
    int segs = 148 / per_seg;
    if (segs < 1) segs = 1;
    if (segs > ceil_div(p.h, 4)) segs = ceil_div(p.h, 4);
    p.rows_per_cta = ceil_div(p.h, segs);
    segs = ceil_div(p.h, p.rows_per_cta);
    void (*kern)(const Tc3Params) = p.half ? conv_tc3_ws_kernel<true> : conv_tc3_ws_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
    dim3 grid(strips, segs, p.n * p.ntiles);
    launch_k_ws(kern, dim3(grid), dim3(WS_THREADS), (size_t)(smem), st, p);
    return check_launch();
  }
  const size_t smem = tc3_smem_bytes(p.kc_real, p.kc_total, p.nt, p.extra != nullptr);
  if (smem > 227 * 1024) return CRFP_ERR_UNSUPPORTED;
  const int ctas_per_sm = (int)((227 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((227 * 1024) / (smem + 1024));
  int segs = (148 * (ctas_per_sm > 4 ? 4 : ctas_per_sm)) / per_seg;
  if (segs < 1) segs = 1;
  if (segs > ceil_div(p.h, 4)) segs = ceil_div(p.h, 4);
  p.rows_per_cta = ceil_div(p.h, segs);
  segs = ceil_div(p.h, p.rows_per_cta);
  cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  dim3 grid(strips, segs, p.n * p.ntiles);
  launch_k(conv_tc3_kernel, dim3(grid), dim3(128), (size_t)(smem), st, p);
  return check_launch();
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_tc3_cout_tile(int cout, int cin, int32_t* nt, int32_t* ntiles) {
  if (!nt || !ntiles || cout <= 0 || cin <= 0 || cin % 8) return CRFP_ERR_BAD_SHAPE;
  int a = 0, b = 0;
  CRFP_TRY(tc3_cout_tile(cout, cin / 8, &a, &b));
  *nt = a; *ntiles = b;
  return CRFP_OK;
}

static int tc3_fwd_impl(const crfp_conv_tc3_desc* d, long long* trace, crfp_stream stream) {
  if (!d) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0 || d->cout <= 0) return CRFP_ERR_BAD_SHAPE;
  if ((long long)d->n * d->h * d->w == 0) return CRFP_OK;
  if (d->ndst < 1 || d->ndst > 2 || d->nsrc < 1 || d->nsrc > 3) return CRFP_ERR_BAD_SHAPE;
  Tc3Params p;
  memset(&p, 0, sizeof(p));
  p.n = d->n; p.h = d->h; p.w = d->w; p.nsrc = d->nsrc;
  for (int s = 0; s < d->nsrc; ++s) {
    p.src[s] = d->src[s].ptr; p.src_c[s] = d->src[s].c; p.src_cstride[s] = d->src[s].cstride;
    p.src_coffset[s] = d->src[s].coffset; p.src_mode[s] = d->src[s]._pad;   /* crfp_tc3_src.mode */
  }
  p.cout = d->cout; p.act = d->act;
  p.weight_hi = reinterpret_cast<const __nv_bfloat16*>(d->weight_hi);
  p.weight_lo = reinterpret_cast<const __nv_bfloat16*>(d->weight_lo);
  p.bias = d->bias;
  p.extra = d->extra; p.w_extra = d->w_extra;
  p.out_kind = d->out_kind; p.shuffle_r = d->shuffle_r; p.ndst = d->ndst;
  for (int s = 0; s < d->ndst; ++s) {
    p.dst[s] = d->dst[s].ptr; p.dst_c[s] = d->dst[s].c; p.dst_cstride[s] = d->dst[s].cstride;
    p.dst_coffset[s] = d->dst[s].coffset;
  }
  p.residual = d->residual; p.res_cstride = d->res_cstride; p.res_coffset = d->res_coffset;
  p.flow = d->flow; p.head_split = d->head_split; p.head_mag = d->head_mag; p.post_scale = d->post_scale;
  if (d->act == CRFP_ACT_DCN_HEAD && !d->flow) return CRFP_ERR_NULL;
  if (d->out_kind == CRFP_TC_OUT_SHUFFLE_F32 && (d->shuffle_r < 1 || d->cout % (d->shuffle_r * d->shuffle_r))) return CRFP_ERR_BAD_SHAPE;
  p.dbg = trace;
  p.half = d->half ? 1 : 0;
  p.res_pre = d->res_pre ? 1 : 0;
  return launch_conv_tc3(p, (cudaStream_t)stream);
}

extern "C" int crfp_conv3x3_tc3_fwd(const crfp_conv_tc3_desc* d, crfp_stream stream) { return tc3_fwd_impl(d, nullptr, stream); }
extern "C" int crfp_conv3x3_tc3_trace(const crfp_conv_tc3_desc* d, long long* trace, crfp_stream stream) {
  if (!trace) return CRFP_ERR_NULL;
  return tc3_fwd_impl(d, trace, stream);
}

// profiling aid: same as crfp_conv3x3_tc3_fwd plus a per-phase clock64 trace of CTA (0,0,0) into `trace` (device int64
// [rows_per_cta+1][8]: row r+1 = {loop top, row staged, converted+synced, next row issued, MMAs issued, MMAs done, epilogue done})

extern "C" size_t crfp_sizeof_conv_tc3_desc(void) { return sizeof(crfp_conv_tc3_desc); }
