// fp32 SIMT 3x3 convolution for thin layers (cout <= 4): the 4-channel HR planes of CRFP
// (forward_resblocks_3, dcn_3 block / fuse / heads, encoder_hr, conv_tttf, conv_last) and FNet's 32->2 head.
// These are HBM/L1-bandwidth bound (36..90 MAC per 16..40 B), so: one thread per output pixel, float4
// channel-quads loaded straight from global (NHWC: one LDG.128 per tap per quad, neighbours hit L1),
// weights broadcast from shared memory, every elementwise consumer fused into the epilogue:
//   EPI_STD      bias + LeakyReLU/ReLU/tanh*256/DCN-head + residual + post_scale
//   EPI_BLEND    conv_tttf + fovea blend + LeakyReLU:  S = lrelu(m*F + (1-m)*S)      (model/CRFP.py:1672-1675)
//   EPI_OUT_NCHW conv_last + bilinear x8 base of the LR frame, planar NCHW store     (model/CRFP.py:1678-1683)
#include <stdlib.h>

#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace crfp {

// Shared epilogue of the thin kernels for one output pixel (v = conv + bias, 4 channels).  EPI >= 0: the epilogue kind is a
// compile-time constant (the persistent TMA kernel is instantiated per kind: a third of the code, no instruction-fetch
// stalls on the untaken variants); EPI < 0: runtime dispatch on P.epi.
template <int EPI = -1>
__device__ __forceinline__ void thin_epilogue(const ConvParams& P, int n, int y, int x, float e0, float e1, float e2, float e3) {
  float v[4] = {e0, e1, e2, e3};
  const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
  const int epi = EPI >= 0 ? EPI : P.epi;

  if (epi == EPI_BLEND) {
    // m*F + (1-m)*S with m in {0,1}: a select (identical for finite operands, immune to values the mask discards)
    const bool m = P.mask[(size_t)n * P.mask_clip_stride + (size_t)y * P.w + x] != 0;
    const float4 so = __ldg(reinterpret_cast<const float4*>(P.blend_old + pix * 4));
    float4 o;
    o.x = lrelu01(m ? v[0] : so.x);
    o.y = lrelu01(m ? v[1] : so.y);
    o.z = lrelu01(m ? v[2] : so.z);
    o.w = lrelu01(m ? v[3] : so.w);
    *reinterpret_cast<float4*>(P.dst[0] + pix * 4) = o;
    return;
  }
  if (epi == EPI_OUT_NCHW) {
    // base = nn.Upsample(x8, bilinear, align_corners=False)(lr): rscale = 1/8
    const int hl = P.h >> 3, wl = P.w >> 3;
    int y0, y1, x0, x1;
    float ly, lx;
    bilin_src(y, 0.125f, hl, y0, y1, ly);
    bilin_src(x, 0.125f, wl, x0, x1, lx);
    const float* lb = P.base_lr4 + (size_t)n * P.base_clip_stride;
    const float4 p00 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y0 * wl + x0) * 4));
    const float4 p01 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y0 * wl + x1) * 4));
    const float4 p10 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y1 * wl + x0) * 4));
    const float4 p11 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y1 * wl + x1) * 4));
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float bs[3] = {hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x),
                         hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y),
                         hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z)};
    float* ob = P.dst[0] + (size_t)n * P.out_clip_stride + (size_t)y * P.w + x;
    const size_t plane = (size_t)P.h * P.w;
#pragma unroll
    for (int c = 0; c < 3; ++c)      // static indices: v[] / bs[] stay in registers
      if (c < P.out_planes) ob[c * plane] = v[c] + bs[c];
    return;
  }

  if (P.act == CRFP_ACT_DCN_HEAD) {
    const float2 fl = __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // same branch-free formulation as the tensor-core heads: 1 - k / (exp(k v) + 1), k = 2 (tanh) / 1 (sigmoid)
      const bool off = c < P.head_split;
      const float k = off ? 2.f : 1.f;
      const float t = 1.f - __fdividef(k, __expf(k * v[c]) + 1.f);
      v[c] = off ? fmaf(P.head_mag, t, (c & 1) ? fl.x : fl.y) : t;
    }
  } else if (P.act == CRFP_ACT_TANH256) {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = tanhf(v[c]) * 256.f;
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = apply_act(v[c], P.act);
  }
  if (P.residual != nullptr) {
    const float* rp = P.residual + pix * P.res_cstride + P.res_coffset;
    if (P.cout == 4 && ((P.res_cstride | P.res_coffset) & 3) == 0) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(rp));
      v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < P.cout) v[c] += __ldg(rp + c);
    }
  }
  const float ps = (P.post_scale == 0.f) ? 1.f : P.post_scale;
  float* op = P.dst[0] + pix * P.dst_cstride[0] + P.dst_coffset[0];
  if (P.cout == 4 && ((P.dst_cstride[0] | P.dst_coffset[0]) & 3) == 0) {
    *reinterpret_cast<float4*>(op) = make_float4(v[0] * ps, v[1] * ps, v[2] * ps, v[3] * ps);
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < P.cout) op[c] = v[c] * ps;
    // zero the padding channels of a 4-wide destination pixel so float4 consumers read zeros
    if (P.dst_c[0] > P.cout) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c >= P.cout && c < P.dst_c[0]) op[c] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) conv_thin_kernel(const ConvParams P) {
  extern __shared__ __align__(16) float s_w[];  // [9][cin_packed][4]
  pdl_trigger();
  const int nw4 = 9 * P.cin_packed;
  for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < nw4; i += blockDim.x * blockDim.y)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(P.weight) + i);
  pdl_wait();
  __syncthreads();

  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= P.w || y >= P.h) return;

  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int nq = P.cin_packed >> 2;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const float4* wt = reinterpret_cast<const float4*>(s_w) + (size_t)(ky * 3 + kx) * P.cin_packed;
      for (int q = 0; q < nq; ++q) {
        const float4 v = load_quad_fg(P, q, n, yy, xx);
        const float4 w0 = wt[q * 4 + 0], w1 = wt[q * 4 + 1], w2 = wt[q * 4 + 2], w3 = wt[q * 4 + 3];
        a0 = fmaf(v.x, w0.x, a0); a1 = fmaf(v.x, w0.y, a1); a2 = fmaf(v.x, w0.z, a2); a3 = fmaf(v.x, w0.w, a3);
        a0 = fmaf(v.y, w1.x, a0); a1 = fmaf(v.y, w1.y, a1); a2 = fmaf(v.y, w1.z, a2); a3 = fmaf(v.y, w1.w, a3);
        a0 = fmaf(v.z, w2.x, a0); a1 = fmaf(v.z, w2.y, a1); a2 = fmaf(v.z, w2.z, a2); a3 = fmaf(v.z, w2.w, a3);
        a0 = fmaf(v.w, w3.x, a0); a1 = fmaf(v.w, w3.y, a1); a2 = fmaf(v.w, w3.z, a2); a3 = fmaf(v.w, w3.w, a3);
      }
    }
  }
  const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias));
  thin_epilogue(P, n, y, x, a0 + b.x, a1 + b.y, a2 + b.z, a3 + b.w);
}

// ---------------------------------------------------------------------------------------------------------------
// v2 for the HR planes (<= 3 input quads): CTA = 32x32 output pixels, 256 threads, each thread owns a 4-pixel COLUMN
// strip (x = lane, 4 consecutive rows) x 4 output channels.  The (34x34) halo tile of every input quad is staged once
// in shared memory (coalesced float4 loads, zero padding / regional mask applied by the loader); per (quad, kx) a
// thread reads its 6-row column (conflict-free LDS.128, reused by the 3 ky taps) and 12 broadcast weight float4s.
// 16 accumulators, ~54 LDS.128 per 576 FFMA: FFMA bound (the 4-channel convs are ~2x over their HBM time in fp32).
// ROWS = rows per thread (4: 256 threads / CTA, 8: 128 threads / CTA with twice the FFMA per shared-memory load).
template <int NQ, int ROWS>
__global__ void __launch_bounds__(1024 / ROWS, ROWS == 4 ? 3 : 4) conv_thin4_kernel(const ConvParams P) {
  constexpr int TS = 32, HS = TS + 2, PITCH = HS + 1, NT = 1024 / ROWS;
  extern __shared__ __align__(16) float smem_t4[];
  float4* s_in = reinterpret_cast<float4*>(smem_t4);            // [NQ][34][35]
  float4* s_w = s_in + NQ * HS * PITCH;                         // [9][cin_packed]
  const int tid = threadIdx.x + threadIdx.y * 32;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
  pdl_trigger();
  pdl_wait();
  if (P.tile_flags != nullptr && P.tile_flags[((size_t)n * P.tiles_y + blockIdx.y) * P.tiles_x + blockIdx.x] == 0) {
    // tile outside the (dilated) fovea: the conv result would be multiplied by a zero mask
    if (P.tile_mode == 2) {
      const int x = x0 + threadIdx.x;
      for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ROWS * threadIdx.y + r;
        if (x < P.w && y < P.h) {
          const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
          const float4 so = __ldg(reinterpret_cast<const float4*>(P.blend_old + pix * 4));
          *reinterpret_cast<float4*>(P.dst[0] + pix * 4) = make_float4(lrelu01(so.x), lrelu01(so.y), lrelu01(so.z), lrelu01(so.w));
        }
      }
    }
    return;
  }
  for (int i = tid; i < 9 * P.cin_packed; i += NT) s_w[i] = __ldg(reinterpret_cast<const float4*>(P.weight) + i);
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int kind = (P.fg != nullptr) ? 2 : P.qkind[q];
    const float* base = P.qptr[q] + (size_t)n * P.h * P.w * P.qcs[q];
    for (int r = tid; r < HS * HS; r += NT) {
      const int py = r / HS, px = r - py * HS;
      const int y = y0 + py - 1, x = x0 + px - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kind == 2) {
        v = load_quad_fg(P, q, n, y, x);
      } else if (y >= 0 && y < P.h && x >= 0 && x < P.w) {
        const float* g = base + ((size_t)y * P.w + x) * P.qcs[q];
        if (kind == 0) {
          v = __ldg(reinterpret_cast<const float4*>(g));
        } else {
          const float2 t = __ldg(reinterpret_cast<const float2*>(g));
          v.x = t.x; v.y = t.y;
        }
      }
      s_in[(q * HS + py) * PITCH + px] = v;
    }
  }
  __syncthreads();
  const int tx = threadIdx.x, ty = threadIdx.y;
  float acc[ROWS][4];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      float4 col[ROWS + 2];
#pragma unroll
      for (int r = 0; r < ROWS + 2; ++r) col[r] = s_in[(q * HS + ROWS * ty + r) * PITCH + tx + kx];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float4* wt = s_w + (ky * 3 + kx) * P.cin_packed + q * 4;
        const float4 w0 = wt[0], w1 = wt[1], w2 = wt[2], w3 = wt[3];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float4 v = col[r + ky];
          acc[r][0] = fmaf(v.x, w0.x, fmaf(v.y, w1.x, fmaf(v.z, w2.x, fmaf(v.w, w3.x, acc[r][0]))));
          acc[r][1] = fmaf(v.x, w0.y, fmaf(v.y, w1.y, fmaf(v.z, w2.y, fmaf(v.w, w3.y, acc[r][1]))));
          acc[r][2] = fmaf(v.x, w0.z, fmaf(v.y, w1.z, fmaf(v.z, w2.z, fmaf(v.w, w3.z, acc[r][2]))));
          acc[r][3] = fmaf(v.x, w0.w, fmaf(v.y, w1.w, fmaf(v.z, w2.w, fmaf(v.w, w3.w, acc[r][3]))));
        }
      }
    }
  }
  const int x = x0 + tx;
  if (x >= P.w) return;
  const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias));
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int y = y0 + ROWS * ty + r;
    if (y >= P.h) break;
    thin_epilogue(P, n, y, x, acc[r][0] + b.x, acc[r][1] + b.y, acc[r][2] + b.z, acc[r][3] + b.w);
  }
}

__device__ __forceinline__ void umma_cp16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void umma_cp8(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
// tile outside the (dilated) fovea: nothing to compute; tile_mode 2 (blend layer) still has to carry the old state
// through the LeakyReLU of the blend (S <- lrelu(0*F + 1*S))
__device__ __forceinline__ void thin4_skip_tile(const ConvParams& P, int tile, int per_img, int tiles_x) {
  if (P.tile_mode != 2) return;
  const int n = tile / per_img, tr = tile - n * per_img;
  const int y0 = (tr / tiles_x) * 32, x = (tr % tiles_x) * 32 + threadIdx.x;
  for (int r = 0; r < 4; ++r) {
    const int y = y0 + 4 * threadIdx.y + r;
    if (x < P.w && y < P.h) {
      const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
      const float4 so = __ldg(reinterpret_cast<const float4*>(P.blend_old + pix * 4));
      *reinterpret_cast<float4*>(P.dst[0] + pix * 4) = make_float4(lrelu01(so.x), lrelu01(so.y), lrelu01(so.z), lrelu01(so.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// v3: persistent + double buffered.  Same per-thread arithmetic as conv_thin4_kernel<NQ, 4>, but a CTA walks tiles
// blockIdx.x, +gridDim.x, ... and the (34x34) halo tile of the NEXT tile is fetched with cp.async (zero fill outside
// the image) into the other half of a 2-stage shared-memory ring while the current tile is being convolved, so the
// load latency of a tile is hidden behind the FFMAs of the previous one instead of behind other CTAs.
// Sources must be plain 4-channel (kind 0) or 2-channel (kind 1: 8-byte cp.async, upper half of the quad stays zero).
template <int NQ>
__global__ void __launch_bounds__(256, NQ == 1 ? 3 : 2) conv_thin4p_kernel(const ConvParams P) {
  constexpr int TS = 32, HS = TS + 2, PITCH = HS + 1, STAGE = NQ * HS * PITCH;
  extern __shared__ __align__(16) float smem_t4[];
  float4* s_in = reinterpret_cast<float4*>(smem_t4);            // [2][NQ][34][35]
  float4* s_w = s_in + 2 * STAGE;                               // [9][cin_packed]
  const int tid = threadIdx.x + threadIdx.y * 32;
  const int tiles_x = (P.w + TS - 1) / TS, tiles_y = (P.h + TS - 1) / TS;
  const int per_img = tiles_x * tiles_y, total = per_img * P.n;
  pdl_trigger();
  for (int i = tid; i < 2 * STAGE; i += 256) s_in[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // pad column + kind-1 upper halves
  for (int i = tid; i < 9 * P.cin_packed; i += 256) s_w[i] = __ldg(reinterpret_cast<const float4*>(P.weight) + i);
  pdl_wait();
  __syncthreads();

  auto tile_live = [&](int tile) {   // tiles outside the (dilated) fovea are skipped
    return P.tile_flags == nullptr || P.tile_flags[tile] != 0;   // flags are [n][tiles_y][tiles_x] = tile index order
  };
  // halo slots of this thread (the same five (py, px) positions for every tile): hoisted out of the tile loop
  constexpr int NSLOT = (HS * HS + 255) / 256;
  int slot_yx[NSLOT];
#pragma unroll
  for (int k = 0; k < NSLOT; ++k) {
    const int r = tid + k * 256;
    slot_yx[k] = (r < HS * HS) ? (((r / HS) << 8) | (r % HS)) : -1;
  }
  auto prefetch = [&](int tile, int stage) {
    const int n = tile / per_img, tr = tile - n * per_img;
    const int y0 = (tr / tiles_x) * TS, x0 = (tr % tiles_x) * TS;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float* base = P.qptr[q] + (size_t)n * P.h * P.w * P.qcs[q];
      const int kind = P.qkind[q];
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) {
        if (slot_yx[k] < 0) continue;
        const int py = slot_yx[k] >> 8, px = slot_yx[k] & 255;
        const int y = y0 + py - 1, x = x0 + px - 1;
        const bool in = y >= 0 && y < P.h && x >= 0 && x < P.w;
        const float* g = in ? base + ((size_t)y * P.w + x) * P.qcs[q] : base;
        float4* dst = s_in + stage * STAGE + (q * HS + py) * PITCH + px;
        if (kind == 0) umma_cp16(dst, g, in ? 16u : 0u);
        else umma_cp8(dst, g, in ? 8u : 0u);
      }
    }
  };

  int tile = blockIdx.x;
  while (tile < total && !tile_live(tile)) {   // leading skipped tiles
    thin4_skip_tile(P, tile, per_img, tiles_x);
    tile += gridDim.x;
  }
  if (tile < total) prefetch(tile, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int stage = 0;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float4 bias = __ldg(reinterpret_cast<const float4*>(P.bias));
  while (tile < total) {
    int next = tile + gridDim.x;
    while (next < total && !tile_live(next)) {
      thin4_skip_tile(P, next, per_img, tiles_x);
      next += gridDim.x;
    }
    if (next < total) prefetch(next, stage ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float4* sb = s_in + stage * STAGE;
    // packed fp32 FMAs (FFMA2, sm_100): a scalar 3-register FFMA issues every other cycle per scheduler, the 2-wide form
    // retires two per issue; accumulators are (co 0,1) / (co 2,3) pairs, the input value is duplicated into a pair
    float2 acc[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float4 col[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] = sb[(q * HS + 4 * ty + r) * PITCH + tx + kx];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float4* wt = s_w + (ky * 3 + kx) * P.cin_packed + q * 4;
          const float4 w0 = wt[0], w1 = wt[1], w2 = wt[2], w3 = wt[3];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 v = col[r + ky];
            const float2 vx = make_float2(v.x, v.x), vy = make_float2(v.y, v.y), vz = make_float2(v.z, v.z), vw = make_float2(v.w, v.w);
            acc[r][0] = __ffma2_rn(vx, make_float2(w0.x, w0.y), __ffma2_rn(vy, make_float2(w1.x, w1.y),
                        __ffma2_rn(vz, make_float2(w2.x, w2.y), __ffma2_rn(vw, make_float2(w3.x, w3.y), acc[r][0]))));
            acc[r][1] = __ffma2_rn(vx, make_float2(w0.z, w0.w), __ffma2_rn(vy, make_float2(w1.z, w1.w),
                        __ffma2_rn(vz, make_float2(w2.z, w2.w), __ffma2_rn(vw, make_float2(w3.z, w3.w), acc[r][1]))));
          }
        }
      }
    }
    const int n = tile / per_img, tr = tile - n * per_img;
    const int y0 = (tr / tiles_x) * TS, x = (tr % tiles_x) * TS + tx;
    if (x < P.w) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int y = y0 + 4 * ty + r;
        if (y < P.h) thin_epilogue(P, n, y, x, acc[r][0].x + bias.x, acc[r][0].y + bias.y, acc[r][1].x + bias.z, acc[r][1].y + bias.w);
      }
    }
    __syncthreads();   // everybody is done with this stage before the next prefetch overwrites it
    tile = next;
    stage ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// TMA-fed variant of the persistent kernel: the per-thread halo prefetch (five cp.async with their address arithmetic per
// quad, measured at 18-29 % of the kernel) becomes ONE cp.async.bulk.tensor per quad and tile issued by thread 0; out-of-image
// halo pixels are the TMA's zero fill.  The tile of a quad is a dense [34][34] float4 block (a warp's 32 consecutive
// float4 of one row are conflict-free at any pitch).  2-channel sources (kind 1: the flow input of dcn_3.dcn_block.0) keep the
// 8-byte cp.async path into the same layout.  Completion: one mbarrier per stage (tx bytes), so the per-tile __syncthreads
// in front of the FFMA loop is only needed when a cp.async quad is present.
constexpr int T4T_TILE_BYTES = 34 * 34 * 16;
constexpr int T4T_TILE_PITCH = (T4T_TILE_BYTES + 127) / 128 * 128;   // 18560: TMA destinations 128-byte aligned

template <int NQ, int EPI>
__global__ void __launch_bounds__(256, NQ == 1 ? 3 : 2) conv_thin4t_kernel(const ConvParams P, const __grid_constant__ CUtensorMap tm0,
                                                                           const __grid_constant__ CUtensorMap tm1,
                                                                           const __grid_constant__ CUtensorMap tm2) {
  constexpr int TS = 32, HS = TS + 2, STAGE = NQ * T4T_TILE_PITCH;   // bytes per stage
  extern __shared__ __align__(128) unsigned char smem_t4t[];
  __shared__ uint64_t full[2];
  float4* s_w = reinterpret_cast<float4*>(smem_t4t + 2 * STAGE);     // [9][cin_packed]
  const int tid = threadIdx.x + threadIdx.y * 32;
  const int tiles_x = (P.w + TS - 1) / TS, tiles_y = (P.h + TS - 1) / TS;
  const int per_img = tiles_x * tiles_y, total = per_img * P.n;
  bool any_k1 = false;
  int ntma = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    any_k1 = any_k1 || P.qkind[q] == 1;
    ntma += P.qkind[q] == 0 ? 1 : 0;
  }
  pdl_trigger();
  if (any_k1)
    for (int i = tid; i < 2 * STAGE / 16; i += 256) reinterpret_cast<float4*>(smem_t4t)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 9 * P.cin_packed; i += 256) s_w[i] = __ldg(reinterpret_cast<const float4*>(P.weight) + i);
  if (tid == 0) {
    umma::mbar_init(&full[0], 1);
    umma::mbar_init(&full[1], 1);
    umma::fence_mbar_init();
    umma::tma_prefetch_desc(&tm0);
    if (NQ > 1) umma::tma_prefetch_desc(&tm1);
    if (NQ > 2) umma::tma_prefetch_desc(&tm2);
  }
  umma::fence_proxy_async();   // the zero fill above (generic proxy) precedes any TMA write into the same bytes
  pdl_wait();
  __syncthreads();

  auto tile_live = [&](int tile) { return P.tile_flags == nullptr || P.tile_flags[tile] != 0; };
  auto issue = [&](int tile, int stage) {
    const int n = tile / per_img, tr = tile - n * per_img;
    const int y0 = (tr / tiles_x) * TS, x0 = (tr % tiles_x) * TS;
    if (tid == 0) {
      umma::mbar_arrive_expect_tx(&full[stage], (uint32_t)(ntma * T4T_TILE_BYTES));
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (P.qkind[q] != 0) continue;
        const CUtensorMap* tm = q == 0 ? &tm0 : (q == 1 ? &tm1 : &tm2);
        umma::tma_load_4d(smem_t4t + stage * STAGE + q * T4T_TILE_PITCH, tm, &full[stage], 0, x0 - 1, y0 - 1, n);
      }
    }
    if (any_k1) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (P.qkind[q] != 1) continue;
        const float* base = P.qptr[q] + (size_t)n * P.h * P.w * P.qcs[q];
        float4* dstq = reinterpret_cast<float4*>(smem_t4t + stage * STAGE + q * T4T_TILE_PITCH);
        for (int r = tid; r < HS * HS; r += 256) {
          const int py = r / HS, px = r - py * HS;
          const int y = y0 + py - 1, x = x0 + px - 1;
          const bool in = y >= 0 && y < P.h && x >= 0 && x < P.w;
          const float* g = in ? base + ((size_t)y * P.w + x) * P.qcs[q] : base;
          umma_cp8(dstq + r, g, in ? 8u : 0u);
        }
      }
    }
  };

  int tile = blockIdx.x;
  while (tile < total && !tile_live(tile)) {
    thin4_skip_tile(P, tile, per_img, tiles_x);
    tile += gridDim.x;
  }
  if (tile < total) issue(tile, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int stage = 0;
  uint32_t phase = 0;   // bit s = parity the next wait on full[s] expects
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float4 bias = __ldg(reinterpret_cast<const float4*>(P.bias));
  while (tile < total) {
    int next = tile + gridDim.x;
    while (next < total && !tile_live(next)) {
      thin4_skip_tile(P, next, per_img, tiles_x);
      next += gridDim.x;
    }
    if (next < total) issue(next, stage ^ 1);
    if (any_k1) {
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    umma::mbar_wait(&full[stage], (phase >> stage) & 1u);
    phase ^= 1u << stage;
    if (any_k1) __syncthreads();   // the cp.async quad was written by all threads
    const float4* sb = reinterpret_cast<const float4*>(smem_t4t + stage * STAGE);
    float2 acc[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float4* sq = sb + q * (T4T_TILE_PITCH / 16);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float4 col[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) col[r] = sq[(4 * ty + r) * HS + tx + kx];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float4* wt = s_w + (ky * 3 + kx) * P.cin_packed + q * 4;
          const float4 w0 = wt[0], w1 = wt[1], w2 = wt[2], w3 = wt[3];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float4 v = col[r + ky];
            const float2 vx = make_float2(v.x, v.x), vy = make_float2(v.y, v.y), vz = make_float2(v.z, v.z), vw = make_float2(v.w, v.w);
            acc[r][0] = __ffma2_rn(vx, make_float2(w0.x, w0.y), __ffma2_rn(vy, make_float2(w1.x, w1.y),
                        __ffma2_rn(vz, make_float2(w2.x, w2.y), __ffma2_rn(vw, make_float2(w3.x, w3.y), acc[r][0]))));
            acc[r][1] = __ffma2_rn(vx, make_float2(w0.z, w0.w), __ffma2_rn(vy, make_float2(w1.z, w1.w),
                        __ffma2_rn(vz, make_float2(w2.z, w2.w), __ffma2_rn(vw, make_float2(w3.z, w3.w), acc[r][1]))));
          }
        }
      }
    }
    const int n = tile / per_img, tr = tile - n * per_img;
    const int y0 = (tr / tiles_x) * TS, x = (tr % tiles_x) * TS + tx;
    if (x < P.w) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int y = y0 + 4 * ty + r;
        if (y < P.h) thin_epilogue<EPI>(P, n, y, x, acc[r][0].x + bias.x, acc[r][0].y + bias.y, acc[r][1].x + bias.z, acc[r][1].y + bias.w);
      }
    }
    __syncthreads();   // everybody is done reading this stage before the next TMA / cp.async refills it
    tile = next;
    stage ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

typedef CUresult (*thin_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static thin_tmap_encode_fn thin_tmap_encoder() {
  static thin_tmap_encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (thin_tmap_encode_fn)p;
  }();
  return fn;
}

int launch_conv_thin(const ConvParams& p_in, cudaStream_t st) {
  ConvParams p = p_in;
  if (p.cout > 4 || p.cout_packed != 4) return CRFP_ERR_BAD_SHAPE;
  {  // per-quad fast addressing
    const int nqr = p.qstart[p.nsrc];
    for (int q = 0; q < 3; ++q) { p.qptr[q] = p.src[0]; p.qcs[q] = p.src_cstride[0]; p.qkind[q] = 2; }
    for (int q = 0; q < nqr && q < 3; ++q) {
      int s = 0;
      if (p.nsrc > 1 && q >= p.qstart[1]) s = 1;
      if (p.nsrc > 2 && q >= p.qstart[2]) s = 2;
      const int lc = (q - p.qstart[s]) * 4, rem = p.src_c[s] - lc;
      const float* ptr = p.src[s] + p.src_coffset[s] + lc;
      p.qptr[q] = ptr; p.qcs[q] = p.src_cstride[s];
      if (p.src_mode[s] != CRFP_SRC_PLAIN) continue;
      const bool phys4 = (p.src_coffset[s] + lc + 4 <= p.src_cstride[s]);   // 4 floats physically present
      if ((rem >= 4 || (phys4 && rem == 4)) && ((uintptr_t)ptr & 15) == 0 && (p.src_cstride[s] & 3) == 0) p.qkind[q] = 0;
      else if (rem == 2 && ((uintptr_t)ptr & 7) == 0 && (p.src_cstride[s] & 1) == 0) p.qkind[q] = 1;
    }
  }
  if (p.out_mode != CRFP_OUT_NHWC && p.epi == EPI_STD) return CRFP_ERR_UNSUPPORTED;
  dim3 block(32, 8);
  const int nq = p.qstart[p.nsrc];   // real input quads
  if (p.tile_flags != nullptr && !(nq >= 1 && nq <= 3)) return CRFP_ERR_UNSUPPORTED;
  if (nq >= 1 && nq <= 3 && ((long long)p.h * p.w >= 32 * 32 || p.tile_flags != nullptr)) {
    dim3 grid(ceil_div(p.w, 32), ceil_div(p.h, 32), p.n);
    const size_t smem4 = ((size_t)nq * 34 * 35 + 9 * p.cin_packed) * 16;
    static const int rows = getenv("CRFP_THIN_ROWS") ? atoi(getenv("CRFP_THIN_ROWS")) : 4;   // 8 measured 7 % slower end to end
    static const bool persistent = getenv("CRFP_THIN_V2") == nullptr;   // A/B switch back to one CTA per tile
    bool plain = p.fg == nullptr;
    for (int q = 0; q < nq; ++q) plain = plain && (p.qkind[q] == 0 || p.qkind[q] == 1);
    if (persistent && plain) {
      const size_t smemp = ((size_t)2 * nq * 34 * 35 + 9 * p.cin_packed) * 16;
      const int total = grid.x * grid.y * grid.z;
      static const int sms = [] {   // one process per GPU: the SM count of the current device is queried once
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        return v;
      }();
      const int ctas = sms * (nq == 1 ? 3 : 2);
      dim3 pgrid(total < ctas ? total : ctas);
      static const bool use_tma = getenv("CRFP_THIN_NOTMA") == nullptr;   // A/B: per-thread cp.async halo prefetch
      thin_tmap_encode_fn enc = use_tma ? thin_tmap_encoder() : nullptr;
      bool tma_ok = enc != nullptr && p.qkind[0] == 0;
      CUtensorMap tm[3];
      memset(tm, 0, sizeof(tm));
      for (int q = 0; q < nq && tma_ok; ++q) {
        if (p.qkind[q] != 0) continue;
        const cuuint64_t gdim[4] = {4, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.n};
        const cuuint64_t gstr[3] = {(cuuint64_t)p.qcs[q] * 4, (cuuint64_t)p.w * p.qcs[q] * 4, (cuuint64_t)p.h * p.w * p.qcs[q] * 4};
        const cuuint32_t box[4] = {4, 34, 34, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        tma_ok = enc(&tm[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.qptr[q]), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }
      if (tma_ok) {
        const size_t smemt = (size_t)2 * nq * T4T_TILE_PITCH + (size_t)9 * p.cin_packed * 16;
#define CRFP_THIN4T_E(NQ_, EPI_)                                                                                      \
  do {                                                                                                                \
    cudaFuncSetAttribute(conv_thin4t_kernel<NQ_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemt);     \
    launch_k(conv_thin4t_kernel<NQ_, EPI_>, dim3(pgrid), dim3(32, 8), (size_t)(smemt), st, p, tm[0], tm[1], tm[2]);    \
  } while (0)
#define CRFP_THIN4T(NQ_)                                                  \
  do {                                                                    \
    if (p.epi == EPI_STD) CRFP_THIN4T_E(NQ_, EPI_STD);                    \
    else if (p.epi == EPI_BLEND) CRFP_THIN4T_E(NQ_, EPI_BLEND);           \
    else CRFP_THIN4T_E(NQ_, EPI_OUT_NCHW);                                \
  } while (0)
        if (nq == 1) CRFP_THIN4T(1); else if (nq == 2) CRFP_THIN4T(2); else CRFP_THIN4T(3);
#undef CRFP_THIN4T
#undef CRFP_THIN4T_E
        return check_launch();
      }
#define CRFP_THIN4P(NQ_)                                                                                       \
  do {                                                                                                         \
    cudaFuncSetAttribute(conv_thin4p_kernel<NQ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemp);    \
    launch_k(conv_thin4p_kernel<NQ_>, dim3(pgrid), dim3(32, 8), (size_t)(smemp), st, p);                       \
  } while (0)
      if (nq == 1) CRFP_THIN4P(1); else if (nq == 2) CRFP_THIN4P(2); else CRFP_THIN4P(3);
#undef CRFP_THIN4P
      return check_launch();
    }
#define CRFP_THIN4(NQ_, R_)                                                                                        \
  do {                                                                                                             \
    cudaFuncSetAttribute(conv_thin4_kernel<NQ_, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);     \
    launch_k(conv_thin4_kernel<NQ_, R_>, dim3(grid), dim3(32, 32 / R_), (size_t)(smem4), st, p);                   \
  } while (0)
    if (rows == 4) {
      if (nq == 1) CRFP_THIN4(1, 4); else if (nq == 2) CRFP_THIN4(2, 4); else CRFP_THIN4(3, 4);
    } else {
      if (nq == 1) CRFP_THIN4(1, 8); else if (nq == 2) CRFP_THIN4(2, 8); else CRFP_THIN4(3, 8);
    }
#undef CRFP_THIN4
    return check_launch();
  }
  dim3 grid(ceil_div(p.w, 32), ceil_div(p.h, 8), p.n);
  const size_t smem = (size_t)9 * p.cin_packed * 4 * sizeof(float);
  if (smem > 48 * 1024) return CRFP_ERR_UNSUPPORTED;
  launch_k(conv_thin_kernel, dim3(grid), dim3(block), (size_t)(smem), st, p);
  return check_launch();
}

int launch_conv(const ConvParams& p, cudaStream_t st) {
  if (p.cout_packed == 4) return launch_conv_thin(p, st);
  return launch_conv_wide(p, st);
}

}  // namespace crfp
