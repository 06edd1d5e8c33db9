// DCNv2 align kernel on the tensor cores (bf16 storage): the modulated deformable im2col is gathered straight into
// the UMMA A-operand tile in shared memory and contracted with tcgen05.mma — `columns` never exists in HBM.
//
// Reference: dcn_v2.DCNv2.forward (/root/reference/model/CRFP.py:318-320,350); SURVEY.md 8(a) a7.
//   CTA = 16x8 output pixels (M = 128), 256 threads.
//   gather   every (pixel, 8-wide K chunk) record = 2 (group,tap) samples x 4 channels: offsets/mask read as one
//            float4 + one float2 (fp32, exact sampling positions), 4 bilinear corners x 8 B (4 bf16) per sample,
//            fp32 lerp x mask -> 8 bf16 -> one 16-byte st.shared into A[kc][pixel] (K-major, no swizzle,
//            LBO = 129*16 B so that consecutive-kc lanes hit distinct banks, SBO = 128 B)
//   contract 18 x tcgen05.mma M128 N32 K16, B = W[32][288] bf16 resident in smem, fp32 accumulators in TMEM
//   epilogue TMEM -> registers (+bias) -> bf16 NHWC, 64 B per pixel.
// One CTA per SM on purpose: the gather's working set (tile + reach of the offsets) must stay in L1.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace crfp {

constexpr int DTW = 16, DTH = 8;    // tile
constexpr int DKC = 36;             // 288 / 8 K chunks
constexpr int DAP = 129;            // records per kc row of A (128 + 1 pad)

struct DcnTcParams {
  int n, h, w;
  const __nv_bfloat16* x; int x_cstride, x_coffset;       // bf16 NHWC, 32 channels used
  const float* offset; int off_cstride, off_coffset;      // fp32: (dy,dx) per (g,t)
  const float* mask; int mask_cstride, mask_coffset;      // fp32 per (g,t)
  const __nv_bfloat16* weight;                            // [36][32][8] bf16, k = (g*9+t)*4+c
  const float* bias;                                      // [32]
  __nv_bfloat16* out; int out_cstride, out_coffset;
};

__device__ __forceinline__ void acc4_bf16(float* a, const __nv_bfloat16* p, float wgt) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  a[0] += wgt * lo.x; a[1] += wgt * lo.y; a[2] += wgt * hi.x; a[3] += wgt * hi.y;
}

__device__ __forceinline__ void dcn_sample_bf16(const DcnTcParams& P, const __nv_bfloat16* img, int gt, int y, int x,
                                                float dy, float dx, float m, float* v) {
  const int g = gt / 9, t = gt - g * 9;
  const int i = t / 3, j = t - i * 3;
  int y0, x0;
  float w00, w01, w10, w11;
  dcn_corner_w(dcn_pos(y, i, dy), dcn_pos(x, j, dx), P.h, P.w, y0, x0, w00, w01, w10, w11);
  v[0] = v[1] = v[2] = v[3] = 0.f;
  const __nv_bfloat16* p = img + ((long long)y0 * P.w + x0) * P.x_cstride + g * 4;
  if (w00 != 0.f) acc4_bf16(v, p, w00);
  if (w01 != 0.f) acc4_bf16(v, p + P.x_cstride, w01);
  if (w10 != 0.f) acc4_bf16(v, p + (long long)P.w * P.x_cstride, w10);
  if (w11 != 0.f) acc4_bf16(v, p + (long long)P.w * P.x_cstride + P.x_cstride, w11);
  v[0] *= m; v[1] *= m; v[2] *= m; v[3] *= m;
}

__global__ void __launch_bounds__(256, 1) dcn_tc_kernel(const DcnTcParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[32];
  uint4* sB = reinterpret_cast<uint4*>(smem);   // [36][32] records
  uint4* sA = sB + DKC * 32;                     // [36][129] records
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tiles_x = (P.w + DTW - 1) / DTW;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int n = blockIdx.y;
  const int x0t = tx * DTW, y0t = ty * DTH;

  pdl_trigger();
  for (int i = tid; i < DKC * 32; i += 256) umma::cp_async16(sB + i, reinterpret_cast<const uint4*>(P.weight) + i, 16u);
  umma::cp_async_commit();
  if (tid < 32) s_bias[tid] = P.bias[tid];
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 32);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }

  pdl_wait();
  // ---- gather: 128 px x 36 chunks = 4608 records, 18 per thread; consecutive lanes = consecutive kc of one pixel
  const __nv_bfloat16* img = P.x + (size_t)n * P.h * P.w * P.x_cstride + P.x_coffset;
#pragma unroll 2
  for (int idx = tid; idx < 128 * DKC; idx += 256) {
    const int m = idx / DKC, kc = idx - m * DKC;
    const int y = y0t + (m >> 4), x = x0t + (m & 15);
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (y < P.h && x < P.w) {
      const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
      const float4 off = __ldg(reinterpret_cast<const float4*>(P.offset + pix * P.off_cstride + P.off_coffset + kc * 4));
      const float2 mk = __ldg(reinterpret_cast<const float2*>(P.mask + pix * P.mask_cstride + P.mask_coffset + kc * 2));
      float v0[4], v1[4];
      dcn_sample_bf16(P, img, 2 * kc, y, x, off.x, off.y, mk.x, v0);
      dcn_sample_bf16(P, img, 2 * kc + 1, y, x, off.z, off.w, mk.y, v1);
      rec.x = umma::pack_bf16(v0[0], v0[1]);
      rec.y = umma::pack_bf16(v0[2], v0[3]);
      rec.z = umma::pack_bf16(v1[0], v1[1]);
      rec.w = umma::pack_bf16(v1[2], v1[3]);
    }
    sA[kc * DAP + m] = rec;
  }
  umma::cp_async_wait<0>();
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(128, 32);
    const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
#pragma unroll 1
    for (int ks = 0; ks < DKC / 2; ++ks) {
      const uint64_t da = umma::make_desc(a0 + (uint32_t)(2 * ks) * (DAP * 16), DAP * 16, 128);
      const uint64_t db = umma::make_desc(b0 + (uint32_t)(2 * ks) * (32 * 16), 32 * 16, 128);
      umma::mma_bf16(taddr, da, db, idesc, ks != 0 ? 1u : 0u);
    }
    umma::mma_commit(&bar);
  }
  if (warp < 4) {
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    float v[32];
    umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16), v);
    const int m = tid;  // TMEM lane = pixel index in the tile
    const int y = y0t + (m >> 4), x = x0t + (m & 15);
    if (y < P.h && x < P.w) {
      const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
      __nv_bfloat16* op = P.out + pix * P.out_cstride + P.out_coffset;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = umma::pack_bf16(v[8 * j + 0] + s_bias[8 * j + 0], v[8 * j + 1] + s_bias[8 * j + 1]);
        o.y = umma::pack_bf16(v[8 * j + 2] + s_bias[8 * j + 2], v[8 * j + 3] + s_bias[8 * j + 3]);
        o.z = umma::pack_bf16(v[8 * j + 4] + s_bias[8 * j + 4], v[8 * j + 5] + s_bias[8 * j + 5]);
        o.w = umma::pack_bf16(v[8 * j + 6] + s_bias[8 * j + 6], v[8 * j + 7] + s_bias[8 * j + 7]);
        *reinterpret_cast<uint4*>(op + 8 * j) = o;
      }
    }
    umma::fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, 32);
}

// ------------------------------------------------------------------------------------------------ fp32-accurate variant
// fp32 NHWC input / output; the modulated columns are split hi/lo into bf16 and contracted as
// A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (fp32 TMEM accumulation).  K = 288 is processed in two halves of 144 so that
// the hi+lo A tiles (2 x 37 KB) + hi+lo weights (37 KB) leave >= 110 KB of L1 for the gather's working set.
constexpr int D3KH = 18;  // K chunks per half

struct DcnTc3Params {
  int n, h, w;
  const float* x; int x_cstride, x_coffset;
  const float* offset; int off_cstride, off_coffset;
  const float* mask; int mask_cstride, mask_coffset;
  const __nv_bfloat16* w_hi;  // [36][32][8]
  const __nv_bfloat16* w_lo;
  const float* bias;
  float* out; int out_cstride, out_coffset;
  const float* flow_hint;     // optional NHWC 2-channel flow: centres the shared-memory sampling window
  int head_raw;               // offset / mask are raw head-conv outputs: the sampler applies tanh / sigmoid / + flow
  const float* head_flow;     // NHWC 2-channel flow added to the offsets (head_raw)
  float head_mag;
  int32_t* dbg_y0;            // optional parity dump: floor(py), floor(px) of every sample, int32 [n,h,w,72]
  int32_t* dbg_x0;
};

__device__ __forceinline__ void dcn_sample_f32(const DcnTc3Params& P, const float* img, int gt, int y, int x, float dy,
                                               float dx, float m, float* v) {
  const int g = gt / 9, t = gt - g * 9;
  const int i = t / 3, j = t - i * 3;
  int y0, x0;
  float w00, w01, w10, w11;
  dcn_corner_w(dcn_pos(y, i, dy), dcn_pos(x, j, dx), P.h, P.w, y0, x0, w00, w01, w10, w11);
  v[0] = v[1] = v[2] = v[3] = 0.f;
  const float* p = img + ((long long)y0 * P.w + x0) * P.x_cstride + g * 4;
#define CRFP_C4(ptr, wgt)                                                  \
  if ((wgt) != 0.f) {                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(ptr));         \
    v[0] += (wgt) * t4.x; v[1] += (wgt) * t4.y; v[2] += (wgt) * t4.z; v[3] += (wgt) * t4.w; \
  }
  CRFP_C4(p, w00)
  CRFP_C4(p + P.x_cstride, w01)
  CRFP_C4(p + (long long)P.w * P.x_cstride, w10)
  CRFP_C4(p + (long long)P.w * P.x_cstride + P.x_cstride, w11)
#undef CRFP_C4
  v[0] *= m; v[1] *= m; v[2] *= m; v[3] *= m;
}

__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = umma::pack_bf16(a - __bfloat162float(h0), b - __bfloat162float(h1));
}

// Shared-memory sampling window: for each K half (4 deformable groups = 64 B of every input pixel) the
// (16+2R) x (8+2R) neighbourhood of the tile, shifted by the rounded optical flow at the tile centre, is staged once
// with cp.async (zero filled outside the image).  Samples whose 2x2 footprint lies inside the window are gathered
// with LDS.128; the rest (large residual offsets) fall back to global loads — always correct, never assumed.
constexpr int DWR = 10;
constexpr int DWW = DTW + 2 * DWR, DWH = DTH + 2 * DWR;   // 36 x 28 pixels

__device__ __forceinline__ void dcn_load_window(const DcnTc3Params& P, float4* sWin, int n, int wy0, int wx0, int half, int tid) {
  const float* img = P.x + (size_t)n * P.h * P.w * P.x_cstride + P.x_coffset + half * 16;
  for (int i = tid; i < DWH * DWW * 4; i += blockDim.x) {
    const int gl = i & 3, pix = i >> 2;
    const int py = pix / DWW, px = pix - py * DWW;
    const int gy = wy0 + py, gx = wx0 + px;
    const bool in = gy >= 0 && gy < P.h && gx >= 0 && gx < P.w;
    const float* g = in ? img + ((size_t)gy * P.w + gx) * P.x_cstride + gl * 4 : P.x;
    umma::cp_async16(sWin + i, g, in ? 16u : 0u);
  }
}

__device__ __forceinline__ void dcn_sample_win(const DcnTc3Params& P, const float* img, const float4* sWin, int wy0, int wx0,
                                               int half, int gt, int y, int x, float dy, float dx, float m, float* v) {
  const int g = gt / 9, t = gt - g * 9;
  const int i = t / 3, j = t - i * 3;
  int y0, x0;
  float w00, w01, w10, w11;
  dcn_corner_w(dcn_pos(y, i, dy), dcn_pos(x, j, dx), P.h, P.w, y0, x0, w00, w01, w10, w11);
  const int wy = y0 - wy0, wx = x0 - wx0;
  if (wy >= 0 && wy + 1 < DWH && wx >= 0 && wx + 1 < DWW) {
    const float4* p = sWin + (wy * DWW + wx) * 4 + (g - 4 * half);
    const float4 c00 = p[0], c01 = p[4], c10 = p[DWW * 4], c11 = p[DWW * 4 + 4];
    v[0] = (w00 * c00.x + w01 * c01.x + w10 * c10.x + w11 * c11.x) * m;
    v[1] = (w00 * c00.y + w01 * c01.y + w10 * c10.y + w11 * c11.y) * m;
    v[2] = (w00 * c00.z + w01 * c01.z + w10 * c10.z + w11 * c11.z) * m;
    v[3] = (w00 * c00.w + w01 * c01.w + w10 * c10.w + w11 * c11.w) * m;
    return;
  }
  // outside the staged window: global gather
  v[0] = v[1] = v[2] = v[3] = 0.f;
  const float* p = img + ((long long)y0 * P.w + x0) * P.x_cstride + g * 4;
#define CRFP_C4(ptr, wgt)                                                  \
  if ((wgt) != 0.f) {                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(ptr));         \
    v[0] += (wgt) * t4.x; v[1] += (wgt) * t4.y; v[2] += (wgt) * t4.z; v[3] += (wgt) * t4.w; \
  }
  CRFP_C4(p, w00)
  CRFP_C4(p + P.x_cstride, w01)
  CRFP_C4(p + (long long)P.w * P.x_cstride, w10)
  CRFP_C4(p + (long long)P.w * P.x_cstride + P.x_cstride, w11)
#undef CRFP_C4
  v[0] *= m; v[1] *= m; v[2] *= m; v[3] *= m;
}

constexpr int WQC = 10;                       // K chunks per quarter stage (9 real + 1 zero)
constexpr int WNW = 2;                        // window ring depth
constexpr int WSAMP = 384;                    // sampler threads
constexpr int WWR = 12;                       // reach: 10 (max |residual offset|) + 1 (tap) + 1 (bilinear corner)
constexpr int WWW = DTW + 2 * WWR, WWH = DTH + 2 * WWR;   // 40 x 32 pixels
constexpr int WWIN_FLOATS = WWH * WWW * 8;    // 10240 floats = 40 KB per window
constexpr int WA_RECS = WQC * DAP;            // uint4 records of one A stage (hi or lo)

// window variant for the persistent kernel: the window holds 8 channels (the 2 deformable groups of K quarter `q`) per
// pixel, [32][40][8] floats as written by the TMA tile load; `gtr` = (group, tap) index relative to the quarter (0..17)
__device__ __forceinline__ void dcn_sample_win8(const DcnTc3Params& P, const float* img, const float4* sWin, int wy0, int wx0,
                                                int q, int gtr, int y, int x, float dy, float dx, float m, float* v,
                                                long long dbg_idx) {
  const int gl = gtr / 9, t = gtr - gl * 9;
  const int i = t / 3, j = t - i * 3;
  int y0, x0;
  float w00, w01, w10, w11;
  dcn_corner_w(dcn_pos(y, i, dy), dcn_pos(x, j, dx), P.h, P.w, y0, x0, w00, w01, w10, w11);
  if (P.dbg_y0 != nullptr) { P.dbg_y0[dbg_idx] = y0; P.dbg_x0[dbg_idx] = x0; }   // parity dump (uniform branch)
  const int wy = y0 - wy0, wx = x0 - wx0;
  if (wy >= 0 && wy + 1 < WWH && wx >= 0 && wx + 1 < WWW) {
    const float4* p = sWin + (wy * WWW + wx) * 2 + gl;
    const float4 c00 = p[0], c01 = p[2], c10 = p[WWW * 2], c11 = p[WWW * 2 + 2];
    // packed fp32 FMAs (FFMA2): channel pairs (0,1) / (2,3), corner weight as the broadcast operand
    const float2 k00 = make_float2(w00, w00), k01 = make_float2(w01, w01), k10 = make_float2(w10, w10), k11 = make_float2(w11, w11);
    const float2 mm = make_float2(m, m), z = make_float2(0.f, 0.f);
    float2 a = __ffma2_rn(k00, make_float2(c00.x, c00.y), z), b = __ffma2_rn(k00, make_float2(c00.z, c00.w), z);
    a = __ffma2_rn(k01, make_float2(c01.x, c01.y), a); b = __ffma2_rn(k01, make_float2(c01.z, c01.w), b);
    a = __ffma2_rn(k10, make_float2(c10.x, c10.y), a); b = __ffma2_rn(k10, make_float2(c10.z, c10.w), b);
    a = __ffma2_rn(k11, make_float2(c11.x, c11.y), a); b = __ffma2_rn(k11, make_float2(c11.z, c11.w), b);
    a = __ffma2_rn(a, mm, z); b = __ffma2_rn(b, mm, z);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    return;
  }
  // outside the staged window: global gather
  v[0] = v[1] = v[2] = v[3] = 0.f;
  const float* p = img + ((long long)y0 * P.w + x0) * P.x_cstride + (2 * q + gl) * 4;
#define CRFP_C4(ptr, wgt)                                                  \
  if ((wgt) != 0.f) {                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(ptr));         \
    v[0] += (wgt) * t4.x; v[1] += (wgt) * t4.y; v[2] += (wgt) * t4.z; v[3] += (wgt) * t4.w; \
  }
  CRFP_C4(p, w00)
  CRFP_C4(p + P.x_cstride, w01)
  CRFP_C4(p + (long long)P.w * P.x_cstride, w10)
  CRFP_C4(p + (long long)P.w * P.x_cstride + P.x_cstride, w11)
#undef CRFP_C4
  v[0] *= m; v[1] *= m; v[2] *= m; v[3] *= m;
}

__global__ void __launch_bounds__(512, 1) dcn_tc3_kernel(const DcnTc3Params P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[32];
  uint4* sBh = reinterpret_cast<uint4*>(smem);   // [36][32]
  uint4* sBl = sBh + DKC * 32;
  uint4* sAh = sBl + DKC * 32;                   // [18][129]
  uint4* sAl = sAh + D3KH * DAP;
  float4* sWin = reinterpret_cast<float4*>(sAl + D3KH * DAP);   // [28][36][4]
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tiles_x = (P.w + DTW - 1) / DTW;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int n = blockIdx.y;
  const int x0t = tx * DTW, y0t = ty * DTH;

  // constant-only prologue (overlaps the previous kernel under PDL): weights -> smem, bias, TMEM, barrier
  pdl_trigger();
  for (int i = tid; i < DKC * 32; i += 512) {
    umma::cp_async16(sBh + i, reinterpret_cast<const uint4*>(P.w_hi) + i, 16u);
    umma::cp_async16(sBl + i, reinterpret_cast<const uint4*>(P.w_lo) + i, 16u);
  }
  if (tid < 32) s_bias[tid] = P.bias[tid];
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 32);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  pdl_wait();
  // window origin: tile - reach, shifted by the rounded flow at the tile centre (the offsets are flow + residual)
  int wy0 = y0t - DWR, wx0 = x0t - DWR;
  if (P.flow_hint != nullptr) {
    const int cy = min(y0t + DTH / 2, P.h - 1), cx = min(x0t + DTW / 2, P.w - 1);
    const float2 fl = __ldg(reinterpret_cast<const float2*>(P.flow_hint + (((size_t)n * P.h + cy) * (size_t)P.w + cx) * 2));
    wy0 += (int)rintf(fminf(fmaxf(fl.y, -4096.f), 4096.f));
    wx0 += (int)rintf(fminf(fmaxf(fl.x, -4096.f), 4096.f));
  }
  dcn_load_window(P, sWin, n, wy0, wx0, 0, tid);
  umma::cp_async_commit();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_bf16(128, 32);
  const float* img = P.x + (size_t)n * P.h * P.w * P.x_cstride + P.x_coffset;

  for (int half = 0; half < 2; ++half) {
    umma::cp_async_wait<0>();
    __syncthreads();   // this half's window (and, first time, the weights) have landed for everyone
    // ---- gather this half's 18 K chunks for the 128 pixels: 2304 records, <= 5 per thread (16 warps).  All offset / mask loads
    //      (the long-latency HBM stream) are issued up front, then the samples are taken from the window.
    float4 offs[5];
    float2 mks[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int idx = tid + r * 512;
      const int m = idx / D3KH, kl = idx - m * D3KH, kc = half * D3KH + kl;
      const int y = y0t + (m >> 4), x = x0t + (m & 15);
      offs[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      mks[r] = make_float2(0.f, 0.f);
      if (idx < 128 * D3KH && y < P.h && x < P.w) {
        const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
        offs[r] = __ldg(reinterpret_cast<const float4*>(P.offset + pix * P.off_cstride + P.off_coffset + kc * 4));
        mks[r] = __ldg(reinterpret_cast<const float2*>(P.mask + pix * P.mask_cstride + P.mask_coffset + kc * 2));
      }
    }
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int idx = tid + r * 512;
      if (idx >= 128 * D3KH) break;
      const int m = idx / D3KH, kl = idx - m * D3KH, kc = half * D3KH + kl;
      const int y = y0t + (m >> 4), x = x0t + (m & 15);
      uint4 rh = make_uint4(0u, 0u, 0u, 0u), rl = rh;
      if (y < P.h && x < P.w) {
        float v0[4], v1[4];
        dcn_sample_win(P, img, sWin, wy0, wx0, half, 2 * kc, y, x, offs[r].x, offs[r].y, mks[r].x, v0);
        dcn_sample_win(P, img, sWin, wy0, wx0, half, 2 * kc + 1, y, x, offs[r].z, offs[r].w, mks[r].y, v1);
        split_pair(v0[0], v0[1], rh.x, rl.x);
        split_pair(v0[2], v0[3], rh.y, rl.y);
        split_pair(v1[0], v1[1], rh.z, rl.z);
        split_pair(v1[2], v1[3], rh.w, rl.w);
      }
      sAh[kl * DAP + m] = rh;
      sAl[kl * DAP + m] = rl;
    }
    umma::fence_proxy_async();
    __syncthreads();   // A tiles complete; the window buffer is free again
    if (half == 0) {
      dcn_load_window(P, sWin, n, wy0, wx0, 1, tid);   // overlaps with the MMAs below
      umma::cp_async_commit();
    }
    if (warp == 0 && umma::elect_one()) {
      umma::fence_after_sync();
      const uint64_t dAh = umma::make_desc(umma::smem_u32(sAh), DAP * 16, 128), dAl = umma::make_desc(umma::smem_u32(sAl), DAP * 16, 128);
      const uint64_t dBh = umma::make_desc(umma::smem_u32(sBh) + (uint32_t)(half * D3KH) * (32 * 16), 32 * 16, 128);
      const uint64_t dBl = umma::make_desc(umma::smem_u32(sBl) + (uint32_t)(half * D3KH) * (32 * 16), 32 * 16, 128);
      const uint32_t ahl = (uint32_t)dAh, ahh = (uint32_t)(dAh >> 32), all_ = (uint32_t)dAl, alh = (uint32_t)(dAl >> 32);
      const uint32_t bhl = (uint32_t)dBh, bhh = (uint32_t)(dBh >> 32), bll = (uint32_t)dBl, blh = (uint32_t)(dBl >> 32);
#pragma unroll
      for (int ks = 0; ks < D3KH / 2; ++ks) {
        const uint64_t dah = umma::desc_advance(ahl, ahh, 2 * ks * DAP), dal = umma::desc_advance(all_, alh, 2 * ks * DAP);
        const uint64_t dbh = umma::desc_advance(bhl, bhh, 2 * ks * 32), dbl = umma::desc_advance(bll, blh, 2 * ks * 32);
        umma::mma_bf16(taddr, dah, dbh, idesc, (half | ks) != 0 ? 1u : 0u);
        umma::mma_bf16(taddr, dal, dbh, idesc, 1u);
        umma::mma_bf16(taddr, dah, dbl, idesc, 1u);
      }
      umma::mma_commit(&bar);
    }
    // the A tiles are overwritten by the next half: everybody waits for this half's MMAs
    umma::mbar_wait(&bar, (uint32_t)half);
    umma::fence_after_sync();
  }
  if (warp < 4) {
    float v[32];
    umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16), v);
    const int m = tid;
    const int y = y0t + (m >> 4), x = x0t + (m & 15);
    if (y < P.h && x < P.w) {
      const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
      float* op = P.out + pix * P.out_cstride + P.out_coffset;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j] + s_bias[4 * j], v[4 * j + 1] + s_bias[4 * j + 1],
                                                             v[4 * j + 2] + s_bias[4 * j + 2], v[4 * j + 3] + s_bias[4 * j + 3]);
    }
    umma::fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, 32);
}

// ------------------------------------------------------------------------------------------------ persistent pipeline
// Same arithmetic as dcn_tc3_kernel, restructured so that the four phases of a tile overlap instead of alternating:
//   * one CTA per SM, each walking tiles blockIdx.x, +gridDim.x, ...; the weights are loaded once per CTA;
//   * K = 288 is cut into 4 quarters = 2 deformable groups = 8 input channels each (9 real K chunks + 1 zero chunk so
//     that a quarter is 5 K16 steps);
//   * the sampling window of a quarter (32 x 40 pixels x 8 channels, 40 KB) is fetched by ONE TMA tensor-tile load
//     (cp.async.bulk.tensor.4d, zero fill outside the image) into a ring of 2 buffers, 2 quarters ahead;
//   * 12 sampler warps gather + modulate + split into a ring of 2 A stages (the next quarter's offsets / masks are
//     already in registers when a quarter starts);
//   * warp 0 issues the 15 tcgen05.mma of a full stage, commits the stage back to the samplers, then re-arms the
//     window ring; after the 4th quarter warps 0-3 read the accumulator from TMEM and store the tile.
// mbarriers: win_full[2] (TMA bytes), a_full[2] (384 sampler arrivals), a_empty[2] / acc_full (tcgen05.commit).
// "Window slot free" needs no barrier of its own: a_full[k] completes only after every sampler has finished
// reading the window of quarter k.
__global__ void __launch_bounds__(512, 1) dcn_tc3_ws_kernel(const DcnTc3Params P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t win_full[WNW], a_full[2], a_empty[2], acc_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[32];
  __shared__ int2 s_org[WNW];
  float* sWin = reinterpret_cast<float*>(smem);                              // [2][32][40][8]
  uint4* sBh = reinterpret_cast<uint4*>(smem + WNW * WWIN_FLOATS * 4);       // [4][10][32]
  uint4* sBl = sBh + 4 * WQC * 32;
  uint4* sAh = sBl + 4 * WQC * 32;                                           // [2][10][129]
  uint4* sAl = sAh + 2 * WA_RECS;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int tiles_x = (P.w + DTW - 1) / DTW, tiles_y = (P.h + DTH - 1) / DTH;
  const int tiles_img = tiles_x * tiles_y, total = tiles_img * P.n;
  const int my_tiles = ((int)blockIdx.x < total) ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int nq = 4 * my_tiles;

  // ---- constant-only prologue (overlaps the previous kernel under PDL)
  pdl_trigger();
  for (int i = tid; i < DKC * 32; i += 512) {
    const int c = i >> 5, co = i & 31;
    const int dst = ((c / 9) * WQC + (c % 9)) * 32 + co;
    umma::cp_async16(sBh + dst, reinterpret_cast<const uint4*>(P.w_hi) + i, 16u);
    umma::cp_async16(sBl + dst, reinterpret_cast<const uint4*>(P.w_lo) + i, 16u);
  }
  umma::cp_async_commit();
  for (int i = tid; i < 4 * 32; i += 512) {   // the zero K chunk of every quarter
    sBh[((i >> 5) * WQC + 9) * 32 + (i & 31)] = make_uint4(0u, 0u, 0u, 0u);
    sBl[((i >> 5) * WQC + 9) * 32 + (i & 31)] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < 2 * DAP; i += 512) {
    sAh[(i / DAP) * WA_RECS + 9 * DAP + (i % DAP)] = make_uint4(0u, 0u, 0u, 0u);
    sAl[(i / DAP) * WA_RECS + 9 * DAP + (i % DAP)] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 32) s_bias[tid] = P.bias[tid];
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 32);
  if (tid == 0) {
    for (int i = 0; i < WNW; ++i) umma::mbar_init(&win_full[i], 1);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&a_full[i], WSAMP); umma::mbar_init(&a_empty[i], 1); }
    umma::mbar_init(&acc_full, 1);
    umma::fence_mbar_init();
    umma::tma_prefetch_desc(&tmap);
  }
  umma::cp_async_wait<0>();
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  pdl_wait();   // activations (x, offsets, masks, flow hint, out) are only touched from here on

  if (warp < 4) {
    // ================================================================ window loads + MMA issue (warp 0) + epilogue
    // window of quarter kk: origin = tile - reach, shifted by the rounded flow at the tile centre
    auto issue_window = [&](int kk) {
      const int tile = (int)blockIdx.x + (kk >> 2) * (int)gridDim.x;
      const int n = tile / tiles_img, tr = tile - n * tiles_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y0t = ty * DTH, x0t = tx * DTW;
      int wy0 = y0t - WWR, wx0 = x0t - WWR;
      if (P.flow_hint != nullptr) {
        const int cy = min(y0t + DTH / 2, P.h - 1), cx = min(x0t + DTW / 2, P.w - 1);
        const float2 fl = __ldg(reinterpret_cast<const float2*>(P.flow_hint + (((size_t)n * P.h + cy) * (size_t)P.w + cx) * 2));
        wy0 += (int)rintf(fminf(fmaxf(fl.y, -4096.f), 4096.f));
        wx0 += (int)rintf(fminf(fmaxf(fl.x, -4096.f), 4096.f));
      }
      const int slot = kk % WNW;
      s_org[slot] = make_int2(wy0, wx0);
      umma::mbar_arrive_expect_tx(&win_full[slot], (uint32_t)(WWIN_FLOATS * 4));
      umma::tma_load_4d(sWin + slot * WWIN_FLOATS, &tmap, &win_full[slot], 8 * (kk & 3), wx0, wy0, n);
    };
    if (warp == 0 && umma::elect_one())
      for (int kk = 0; kk < min(WNW, nq); ++kk) issue_window(kk);
    __syncwarp();
    const uint32_t idesc = umma::make_idesc_bf16(128, 32);
    for (int it = 0; it < my_tiles; ++it) {
      if (warp == 0) {
        for (int q = 0; q < 4; ++q) {
          const int k = it * 4 + q, stg = k & 1;
          umma::mbar_wait_safe(&a_full[stg], (uint32_t)((k >> 1) & 1));
          umma::fence_after_sync();
          if (umma::elect_one()) {
            const uint64_t dAh = umma::make_desc(umma::smem_u32(sAh + stg * WA_RECS), DAP * 16, 128);
            const uint64_t dAl = umma::make_desc(umma::smem_u32(sAl + stg * WA_RECS), DAP * 16, 128);
            const uint64_t dBh = umma::make_desc(umma::smem_u32(sBh + q * WQC * 32), 32 * 16, 128);
            const uint64_t dBl = umma::make_desc(umma::smem_u32(sBl + q * WQC * 32), 32 * 16, 128);
            const uint32_t ahl = (uint32_t)dAh, ahh = (uint32_t)(dAh >> 32), all_ = (uint32_t)dAl, alh = (uint32_t)(dAl >> 32);
            const uint32_t bhl = (uint32_t)dBh, bhh = (uint32_t)(dBh >> 32), bll = (uint32_t)dBl, blh = (uint32_t)(dBl >> 32);
#pragma unroll
            for (int ks = 0; ks < WQC / 2; ++ks) {
              const uint64_t dah = umma::desc_advance(ahl, ahh, 2 * ks * DAP), dal = umma::desc_advance(all_, alh, 2 * ks * DAP);
              const uint64_t dbh = umma::desc_advance(bhl, bhh, 2 * ks * 32), dbl = umma::desc_advance(bll, blh, 2 * ks * 32);
              umma::mma_bf16(taddr, dah, dbh, idesc, (q | ks) != 0 ? 1u : 0u);
              umma::mma_bf16(taddr, dal, dbh, idesc, 1u);
              umma::mma_bf16(taddr, dah, dbl, idesc, 1u);
            }
            umma::mma_commit(&a_empty[stg]);
            if (q == 3) umma::mma_commit(&acc_full);
            if (k + WNW < nq) issue_window(k + WNW);   // a_full[k] => every sampler is done with window k: its slot is free
          }
          __syncwarp();
        }
      }
      // ---- epilogue of this tile: thread = pixel = TMEM lane
      umma::mbar_wait_safe(&acc_full, (uint32_t)(it & 1));
      umma::fence_after_sync();
      float v[32];
      umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16), v);
      umma::fence_before_sync();
      umma::named_bar_sync(1, 128);     // all four quadrants are in registers: warp 0 may overwrite the accumulator
      umma::fence_after_sync();
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int n = tile / tiles_img, tr = tile - n * tiles_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const int y = ty * DTH + (tid >> 4), x = tx * DTW + (tid & 15);
      if (y < P.h && x < P.w) {
        const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
        float* op = P.out + pix * P.out_cstride + P.out_coffset;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j] + s_bias[4 * j], v[4 * j + 1] + s_bias[4 * j + 1],
                                                               v[4 * j + 2] + s_bias[4 * j + 2], v[4 * j + 3] + s_bias[4 * j + 3]);
      }
    }
  } else {
    // ================================================================ samplers: 1152 records per quarter, 3 per thread
    const int st = tid - 128;
    int rm[3], rkl[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int idx = st + r * WSAMP;
      rm[r] = idx / 9;
      rkl[r] = idx - rm[r] * 9;
    }
    float4 offs[3], noffs[3];
    float2 mks[3], nmks[3];
    float2 fls[3], nfls[3];     // head_raw: flow at the record's pixel
    auto load_offsets = [&](int kk, float4* o, float2* mk, float2* fl) {
      const int tile = (int)blockIdx.x + (kk >> 2) * (int)gridDim.x, q = kk & 3;
      const int n = tile / tiles_img, tr = tile - n * tiles_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int y = ty * DTH + (rm[r] >> 4), x = tx * DTW + (rm[r] & 15);
        o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        mk[r] = make_float2(0.f, 0.f);
        fl[r] = make_float2(0.f, 0.f);
        if (y < P.h && x < P.w) {
          const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
          const int kc = q * 9 + rkl[r];
          o[r] = __ldg(reinterpret_cast<const float4*>(P.offset + pix * P.off_cstride + P.off_coffset + kc * 4));
          mk[r] = __ldg(reinterpret_cast<const float2*>(P.mask + pix * P.mask_cstride + P.mask_coffset + kc * 2));
          if (P.head_raw) fl[r] = __ldg(reinterpret_cast<const float2*>(P.head_flow + pix * 2));
        }
      }
    };
    if (nq > 0) load_offsets(0, offs, mks, fls);
    for (int k = 0; k < nq; ++k) {
      if (k + 1 < nq) load_offsets(k + 1, noffs, nmks, nfls);   // in flight while this quarter is sampled
      const int tile = (int)blockIdx.x + (k >> 2) * (int)gridDim.x, q = k & 3, stg = k & 1, slot = k % WNW;
      const int n = tile / tiles_img, tr = tile - n * tiles_img;
      const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
      const float* img = P.x + (size_t)n * P.h * P.w * P.x_cstride + P.x_coffset;
      umma::mbar_wait_safe(&win_full[slot], (uint32_t)((k / WNW) & 1));
      if (k >= 2) umma::mbar_wait_safe(&a_empty[stg], (uint32_t)(((k >> 1) - 1) & 1));
      const int2 org = s_org[slot];
      const float4* win = reinterpret_cast<const float4*>(sWin + slot * WWIN_FLOATS);
      uint4* ah = sAh + stg * WA_RECS;
      uint4* al = sAl + stg * WA_RECS;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int m = rm[r], kl = rkl[r];
        const int y = ty * DTH + (m >> 4), x = tx * DTW + (m & 15);
        uint4 rh = make_uint4(0u, 0u, 0u, 0u), rl = rh;
        if (y < P.h && x < P.w) {
          float v0[4], v1[4];
          float4 of = offs[r];
          float2 mk = mks[r];
          if (P.head_raw) {   // DCN_module.forward's head activations (CRFP.py:337-349), fused into the sampler
            of.x = head_offset_act(of.x, P.head_mag, fls[r].y); of.y = head_offset_act(of.y, P.head_mag, fls[r].x);
            of.z = head_offset_act(of.z, P.head_mag, fls[r].y); of.w = head_offset_act(of.w, P.head_mag, fls[r].x);
            mk.x = head_mask_act(mk.x); mk.y = head_mask_act(mk.y);
          }
          const long long dbg = ((((long long)n * P.h + y) * P.w + x) * 72) + q * 18 + 2 * kl;
          dcn_sample_win8(P, img, win, org.x, org.y, q, 2 * kl, y, x, of.x, of.y, mk.x, v0, dbg);
          dcn_sample_win8(P, img, win, org.x, org.y, q, 2 * kl + 1, y, x, of.z, of.w, mk.y, v1, dbg + 1);
          split_pair(v0[0], v0[1], rh.x, rl.x);
          split_pair(v0[2], v0[3], rh.y, rl.y);
          split_pair(v1[0], v1[1], rh.z, rl.z);
          split_pair(v1[2], v1[3], rh.w, rl.w);
        }
        ah[kl * DAP + m] = rh;
        al[kl * DAP + m] = rl;
      }
      umma::fence_proxy_async();
      umma::mbar_arrive(&a_full[stg]);
#pragma unroll
      for (int r = 0; r < 3; ++r) { offs[r] = noffs[r]; mks[r] = nmks[r]; fls[r] = nfls[r]; }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, 32);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (tmap_encode_fn)p;
  }();
  return fn;
}

static int sm_count() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
}

// x viewed as a 4-D tensor {channel, x, y, image}; one box = 8 channels x 40 x 32 pixels of one image
static int launch_dcn_tc3_ws(const DcnTc3Params& p, cudaStream_t st) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc) return CRFP_ERR_UNSUPPORTED;
  const float* base = p.x + p.x_coffset;
  if ((uintptr_t)base & 15) return CRFP_ERR_BAD_SHAPE;
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {32, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.n};
  const cuuint64_t gstr[3] = {(cuuint64_t)p.x_cstride * 4, (cuuint64_t)p.w * p.x_cstride * 4, (cuuint64_t)p.h * p.w * p.x_cstride * 4};
  const cuuint32_t box[4] = {8, WWW, WWH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return CRFP_ERR_UNSUPPORTED;
  const size_t smem = (size_t)WNW * WWIN_FLOATS * 4 + (size_t)(2 * 4 * WQC * 32 + 2 * 2 * WA_RECS) * 16;   // 81920 + 40960 + 82560
  cudaError_t e = cudaFuncSetAttribute(dcn_tc3_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  const int total = ceil_div(p.w, DTW) * ceil_div(p.h, DTH) * p.n;
  const int grid = total < sm_count() ? total : sm_count();
  launch_k_ws(dcn_tc3_ws_kernel, dim3(grid), dim3(512), smem, st, p, tmap);
  return check_launch();
}

int launch_dcn_tc3(const crfp_dcn_desc& d, const void* w_lo, const float* flow_hint, cudaStream_t st) {
  if ((long long)d.n * d.h * d.w == 0) return CRFP_OK;
  if (!(d.c == 32 && d.dg == 8 && d.cout == 32 && !d.shared_taps)) return CRFP_ERR_UNSUPPORTED;
  if (!w_lo) return CRFP_ERR_NULL;
  if ((d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 3) return CRFP_ERR_BAD_SHAPE;
  if (((d.off_cstride | d.off_coffset) & 3) || ((d.mask_cstride | d.mask_coffset) & 1)) return CRFP_ERR_BAD_SHAPE;
  DcnTc3Params p;
  p.n = d.n; p.h = d.h; p.w = d.w;
  p.x = d.x; p.x_cstride = d.x_cstride; p.x_coffset = d.x_coffset;
  p.offset = d.offset; p.off_cstride = d.off_cstride; p.off_coffset = d.off_coffset;
  p.mask = d.mask; p.mask_cstride = d.mask_cstride; p.mask_coffset = d.mask_coffset;
  p.w_hi = reinterpret_cast<const __nv_bfloat16*>(d.weight); p.w_lo = reinterpret_cast<const __nv_bfloat16*>(w_lo);
  p.bias = d.bias;
  p.out = d.out; p.out_cstride = d.out_cstride; p.out_coffset = d.out_coffset;
  p.flow_hint = flow_hint;
  p.head_raw = d.head_raw; p.head_flow = d.head_flow; p.head_mag = d.head_mag;
  p.dbg_y0 = d.dbg_y0; p.dbg_x0 = d.dbg_x0;
  if (d.head_raw && !d.head_flow) return CRFP_ERR_NULL;
  if ((d.dbg_y0 != nullptr) != (d.dbg_x0 != nullptr)) return CRFP_ERR_NULL;
  static const bool use_v1 = (getenv("CRFP_DCN_V1") != nullptr);   // A/B switch: the non-persistent kernel
  if (!use_v1) return launch_dcn_tc3_ws(p, st);
  if (d.head_raw || d.dbg_y0) return CRFP_ERR_UNSUPPORTED;        // raw heads / index dump: persistent kernel only
  const size_t smem = (size_t)(2 * DKC * 32 + 2 * D3KH * DAP + DWH * DWW * 4) * 16;  // 36864 + 74304 + 64512 B
  cudaError_t e = cudaFuncSetAttribute(dcn_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  dim3 grid(ceil_div(d.w, DTW) * ceil_div(d.h, DTH), d.n);
  launch_k(dcn_tc3_kernel, dim3(grid), dim3(512), (size_t)(smem), st, p);
  return check_launch();
}

int launch_dcn_tc(const crfp_dcn_desc& d, cudaStream_t st) {
  if ((long long)d.n * d.h * d.w == 0) return CRFP_OK;
  if (!(d.c == 32 && d.dg == 8 && d.cout == 32 && !d.shared_taps)) return CRFP_ERR_UNSUPPORTED;
  if (d.head_raw || d.dbg_y0 || d.dbg_x0) return CRFP_ERR_UNSUPPORTED;
  if ((d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 7) return CRFP_ERR_BAD_SHAPE;
  if (((d.off_cstride | d.off_coffset) & 3) || ((d.mask_cstride | d.mask_coffset) & 1)) return CRFP_ERR_BAD_SHAPE;
  DcnTcParams p;
  p.n = d.n; p.h = d.h; p.w = d.w;
  p.x = reinterpret_cast<const __nv_bfloat16*>(d.x); p.x_cstride = d.x_cstride; p.x_coffset = d.x_coffset;
  p.offset = d.offset; p.off_cstride = d.off_cstride; p.off_coffset = d.off_coffset;
  p.mask = d.mask; p.mask_cstride = d.mask_cstride; p.mask_coffset = d.mask_coffset;
  p.weight = reinterpret_cast<const __nv_bfloat16*>(d.weight); p.bias = d.bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(d.out); p.out_cstride = d.out_cstride; p.out_coffset = d.out_coffset;
  const size_t smem = (size_t)(DKC * 32 + DKC * DAP) * 16;  // 18432 + 74304 = 92736 B
  cudaError_t e = cudaFuncSetAttribute(dcn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  dim3 grid(ceil_div(d.w, DTW) * ceil_div(d.h, DTH), d.n);
  launch_k(dcn_tc_kernel, dim3(grid), dim3(256), (size_t)(smem), st, p);
  return check_launch();
}

}  // namespace crfp

using namespace crfp;

// bf16 variant of crfp_dcn_v2_fwd: x / out are bf16 NHWC, weight bf16 [36][32][8] (k = (g*9+t)*4+c), offset / mask /
// bias fp32.  Same descriptor struct; pointers are reinterpreted.
// fp32-accurate tensor-core variant: x / out fp32 NHWC; d->weight = hi part, weight_lo = lo part of the bf16 split of
// the [36][32][8] packed weight (k = (g*9+t)*4+c).
extern "C" int crfp_dcn_v2_tc3_fwd(const crfp_dcn_desc* d, const void* weight_lo, const float* flow_hint, crfp_stream stream) {
  if (!d || !d->x || !d->offset || !d->mask || !d->weight || !d->bias || !d->out || !weight_lo) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0) return CRFP_ERR_BAD_SHAPE;
  return launch_dcn_tc3(*d, weight_lo, flow_hint, (cudaStream_t)stream);
}

extern "C" int crfp_dcn_v2_tc_fwd(const crfp_dcn_desc* d, crfp_stream stream) {
  if (!d || !d->x || !d->offset || !d->mask || !d->weight || !d->bias || !d->out) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0) return CRFP_ERR_BAD_SHAPE;
  return launch_dcn_tc(*d, (cudaStream_t)stream);
}
