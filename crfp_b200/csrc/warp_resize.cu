// flow_warp (bilinear backward warp), bilinear resize, 2x2 average pool and NCHW<->NHWC layout kernels.
// All are pure HBM-bandwidth kernels: NHWC float4 channel-quads, one thread per (pixel, quad), coalesced.
//
// flow_warp replaces /root/reference/model/CRFP.py:90-130 (meshgrid + normalise + F.grid_sample): the
// sampling position is computed with the SAME fp32 operation sequence (no FMA contraction) so that the
// integer corner indices agree bit for bit with the reference's CPU path:
//   g  = 2.0f*(x + flow_x)/(W-1) - 1.0f                (CRFP.py:118-121)
//   ix = ((g + 1.0f)/2)*(W-1)                          (ATen grid_sampler_unnormalize, align_corners=True)
#include "common.cuh"

namespace crfp {

__device__ __forceinline__ float warp_coord(int i, float f, int size) {
  const float denom = (float)((size - 1) > 1 ? (size - 1) : 1);
  const float gf = __fadd_rn((float)i, f);
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, gf), denom), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

__global__ void __launch_bounds__(256) flow_warp_kernel(const crfp_warp_desc D) {
  pdl_trigger();
  pdl_wait();
  const int cq = D.c >> 2;
  const long long total = (long long)D.n * D.h * D.w * cq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = (int)(idx % cq);
  const long long pix = idx / cq;
  const int x = (int)(pix % D.w);
  const int y = (int)((pix / D.w) % D.h);
  const int n = (int)(pix / ((long long)D.w * D.h));
  const float2 fl = __ldg(reinterpret_cast<const float2*>(D.flow + pix * 2));
  float ix = warp_coord(x, fl.x, D.w);
  float iy = warp_coord(y, fl.y, D.h);
  if (D.border) {  // padding_mode='border': clip_coordinates
    ix = fminf(fmaxf(ix, 0.f), (float)(D.w - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(D.h - 1));
  }
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  // grid_sample weights: nw = (x1-ix)(y1-iy), ne = (ix-x0)(y1-iy), sw = (x1-ix)(iy-y0), se = (ix-x0)(iy-y0)
  const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
  const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
  const float* base = D.x + (size_t)n * D.h * D.w * D.x_cstride + D.x_coffset + q * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vx0 = (x0 >= 0 && x0 < D.w), vx1 = (x1 >= 0 && x1 < D.w);
  const bool vy0 = (y0 >= 0 && y0 < D.h), vy1 = (y1 >= 0 && y1 < D.h);
#define CRFP_ACC(vy, vx, yy, xx, wgt)                                                                  \
  if ((vy) && (vx)) {                                                                                  \
    const float4 t = __ldg(reinterpret_cast<const float4*>(base + ((size_t)(yy) * D.w + (xx)) * D.x_cstride)); \
    const float w_ = (wgt);                                                                            \
    acc.x += t.x * w_; acc.y += t.y * w_; acc.z += t.z * w_; acc.w += t.w * w_;                        \
  }
  CRFP_ACC(vy0, vx0, y0, x0, wx0 * wy0)
  CRFP_ACC(vy0, vx1, y0, x1, wx1 * wy0)
  CRFP_ACC(vy1, vx0, y1, x0, wx0 * wy1)
  CRFP_ACC(vy1, vx1, y1, x1, wx1 * wy1)
#undef CRFP_ACC
  *reinterpret_cast<float4*>(D.out + (size_t)pix * D.out_cstride + D.out_coffset + q * 4) = acc;
}

// The four L1 warps of a CRFP_DSV frame in ONE launch (model/CRFP.py:1570-1577): P (32 ch) -> P_w and the three 8-channel
// slices of the recurrent L1 state -> channels 24..31 of level k's input, all along the same flow.  One thread per
// (pixel, 4-channel quad); 14 quads per pixel; the same coordinate arithmetic as flow_warp_kernel.
__global__ void __launch_bounds__(256) flow_warp_l1_kernel(int n, int h, int w, const float* __restrict__ flow,
                                                           const float* __restrict__ P, float* __restrict__ P_w,
                                                           const float* __restrict__ state, float* __restrict__ cur0,
                                                           float* __restrict__ cur1, float* __restrict__ cur2) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * h * w * 14;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = (int)(idx % 14);
  const long long pix = idx / 14;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const int b = (int)(pix / ((long long)w * h));
  const float2 fl = __ldg(reinterpret_cast<const float2*>(flow + pix * 2));
  const float ix = warp_coord(x, fl.x, w), iy = warp_coord(y, fl.y, h);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
  const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
  const float* base;
  float* out;
  int cs;
  if (q < 8) { base = P + (size_t)b * h * w * 32 + q * 4; cs = 32; out = P_w + pix * 32 + q * 4; }
  else {
    const int j = q - 8, k = j >> 1;
    base = state + (size_t)b * h * w * 24 + j * 4; cs = 24;
    out = (k == 0 ? cur0 : k == 1 ? cur1 : cur2) + pix * 32 + 24 + (j & 1) * 4;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vx0 = (x0 >= 0 && x0 < w), vx1 = (x1 >= 0 && x1 < w);
  const bool vy0 = (y0 >= 0 && y0 < h), vy1 = (y1 >= 0 && y1 < h);
#define CRFP_ACC(vy, vx, yy, xx, wgt)                                                                  \
  if ((vy) && (vx)) {                                                                                  \
    const float4 t = __ldg(reinterpret_cast<const float4*>(base + ((size_t)(yy) * w + (xx)) * cs));   \
    const float w_ = (wgt);                                                                            \
    acc.x += t.x * w_; acc.y += t.y * w_; acc.z += t.z * w_; acc.w += t.w * w_;                        \
  }
  CRFP_ACC(vy0, vx0, y0, x0, wx0 * wy0)
  CRFP_ACC(vy0, vx1, y0, x1, wx1 * wy0)
  CRFP_ACC(vy1, vx0, y1, x0, wx0 * wy1)
  CRFP_ACC(vy1, vx1, y1, x1, wx1 * wy1)
#undef CRFP_ACC
  *reinterpret_cast<float4*>(out) = acc;
}

__global__ void __launch_bounds__(256) flow_warp_indices_kernel(int n, int h, int w, const float* __restrict__ flow,
                                                                int32_t* __restrict__ x0, int32_t* __restrict__ y0) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * h * w;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % w), y = (int)((pix / w) % h);
  const float2 fl = __ldg(reinterpret_cast<const float2*>(flow + pix * 2));
  x0[pix] = (int)floorf(warp_coord(x, fl.x, w));
  y0[pix] = (int)floorf(warp_coord(y, fl.y, h));
}

// ---- bilinear resize (align_corners=False), NHWC, all channels of the pixel
__global__ void __launch_bounds__(256) resize_bilinear_kernel(int n, int hin, int win, int c, const float* __restrict__ in,
                                                              int hout, int wout, float rh, float rw, float mul,
                                                              float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * hout * wout * c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  const long long pix = idx / c;
  const int x = (int)(pix % wout);
  const int y = (int)((pix / wout) % hout);
  const int b = (int)(pix / ((long long)wout * hout));
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_src(y, rh, hin, y0, y1, ly);
  bilin_src(x, rw, win, x0, x1, lx);
  const float* ib = in + (size_t)b * hin * win * c + ch;
  const float v00 = __ldg(ib + ((size_t)y0 * win + x0) * c), v01 = __ldg(ib + ((size_t)y0 * win + x1) * c);
  const float v10 = __ldg(ib + ((size_t)y1 * win + x0) * c), v11 = __ldg(ib + ((size_t)y1 * win + x1) * c);
  const float hy = 1.f - ly, hx = 1.f - lx;
  out[idx] = (hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11)) * mul;
}

// vectorised variants: one thread = VEC consecutive channels of one output pixel (float2 for the 2-channel flows,
// float4 for feature maps); same arithmetic per channel as the scalar kernel
template <typename VT>
__global__ void __launch_bounds__(256) resize_bilinear_vec_kernel(int n, int hin, int win, int cv, const VT* __restrict__ in,
                                                                  int hout, int wout, float rh, float rw, float mul,
                                                                  VT* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * hout * wout * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % cv);
  const long long pix = idx / cv;
  const int x = (int)(pix % wout);
  const int y = (int)((pix / wout) % hout);
  const int b = (int)(pix / ((long long)wout * hout));
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_src(y, rh, hin, y0, y1, ly);
  bilin_src(x, rw, win, x0, x1, lx);
  const VT* ib = in + (size_t)b * hin * win * cv + ch;
  const VT v00 = __ldg(ib + ((size_t)y0 * win + x0) * cv), v01 = __ldg(ib + ((size_t)y0 * win + x1) * cv);
  const VT v10 = __ldg(ib + ((size_t)y1 * win + x0) * cv), v11 = __ldg(ib + ((size_t)y1 * win + x1) * cv);
  const float hy = 1.f - ly, hx = 1.f - lx;
  VT o;
  o.x = (hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x)) * mul;
  o.y = (hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y)) * mul;
  if constexpr (sizeof(VT) == 16) {
    o.z = (hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z)) * mul;
    o.w = (hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w)) * mul;
  }
  out[idx] = o;
}

__global__ void __launch_bounds__(256) avgpool2_kernel(int n, int hin, int win, int c, const float* __restrict__ in,
                                                       float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int ho = hin >> 1, wo = win >> 1;
  const long long total = (long long)n * ho * wo * c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  const long long pix = idx / c;
  const int x = (int)(pix % wo);
  const int y = (int)((pix / wo) % ho);
  const int b = (int)(pix / ((long long)wo * ho));
  const float* ib = in + (((size_t)b * hin + 2 * y) * win + 2 * x) * c + ch;
  const float s = (__ldg(ib) + __ldg(ib + c)) + (__ldg(ib + (size_t)win * c) + __ldg(ib + (size_t)win * c + c));
  out[idx] = s * 0.25f;
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(int n, int c, int h, int w, const float* __restrict__ in,
                                                           long long in_image_stride, int cpad, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * h * w;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const long long hw = (long long)h * w;
  const int b = (int)(pix / hw);
  const long long p = pix - (long long)b * hw;
  const float* ib = in + (size_t)b * in_image_stride + p;
  float* ob = out + (size_t)pix * cpad;
  for (int ch = 0; ch < cpad; ++ch) ob[ch] = (ch < c) ? __ldg(ib + (size_t)ch * hw) : 0.f;
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(int n, int c, int h, int w, const float* __restrict__ in,
                                                           int cs, int co, float* __restrict__ out,
                                                           long long out_image_stride) {
  pdl_trigger();
  pdl_wait();
  const long long hw = (long long)h * w;
  const long long total = (long long)n * hw * c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long p = idx % hw;
  const int ch = (int)((idx / hw) % c);
  const int b = (int)(idx / (hw * c));
  out[(size_t)b * out_image_stride + (size_t)ch * hw + p] = __ldg(in + ((size_t)b * hw + p) * cs + co + ch);
}

// ---- bf16 flow_warp: thread = (pixel, 8-channel chunk); fp32 positions / weights, bf16 storage
__global__ void __launch_bounds__(256) flow_warp_bf16_kernel(const crfp_warp_desc D) {
  pdl_trigger();
  pdl_wait();
  const int cq = D.c >> 3;
  const long long total = (long long)D.n * D.h * D.w * cq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = (int)(idx % cq);
  const long long pix = idx / cq;
  const int x = (int)(pix % D.w);
  const int y = (int)((pix / D.w) % D.h);
  const int n = (int)(pix / ((long long)D.w * D.h));
  const float2 fl = __ldg(reinterpret_cast<const float2*>(D.flow + pix * 2));
  const float ix = warp_coord(x, fl.x, D.w), iy = warp_coord(y, fl.y, D.h);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
  const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(D.x) + (size_t)n * D.h * D.w * D.x_cstride + D.x_coffset + q * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool vx0 = (x0 >= 0 && x0 < D.w), vx1 = (x1 >= 0 && x1 < D.w);
  const bool vy0 = (y0 >= 0 && y0 < D.h), vy1 = (y1 >= 0 && y1 < D.h);
#define CRFP_ACC8(vy, vx, yy, xx, wgt)                                                                        \
  if ((vy) && (vx)) {                                                                                         \
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)(yy) * D.w + (xx)) * D.x_cstride));   \
    const __nv_bfloat162* t2 = reinterpret_cast<const __nv_bfloat162*>(&t);                                   \
    const float w_ = (wgt);                                                                                   \
    for (int k = 0; k < 4; ++k) {                                                                             \
      const float2 f = __bfloat1622float2(t2[k]);                                                             \
      acc[2 * k] += f.x * w_; acc[2 * k + 1] += f.y * w_;                                                     \
    }                                                                                                         \
  }
  CRFP_ACC8(vy0, vx0, y0, x0, wx0 * wy0)
  CRFP_ACC8(vy0, vx1, y0, x1, wx1 * wy0)
  CRFP_ACC8(vy1, vx0, y1, x0, wx0 * wy1)
  CRFP_ACC8(vy1, vx1, y1, x1, wx1 * wy1)
#undef CRFP_ACC8
  __nv_bfloat162 o[4];
  for (int k = 0; k < 4; ++k) o[k] = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(D.out) + (size_t)pix * D.out_cstride + D.out_coffset + q * 8) =
      *reinterpret_cast<uint4*>(o);
}

// flow_lv3 = up2(flow)*2 (model/CRFP.py:1565) written twice: fp32 NHWC2 (warps, DCN heads) and bf16 NHWC8
// [fx, fy, 0 x6] (third source of the tensor-core dcn_block conv)
__global__ void __launch_bounds__(256) flow_up2_dual_kernel(int n, int h, int w, const float* __restrict__ flow,
                                                            float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf8) {
  pdl_trigger();
  pdl_wait();
  const int ho = 2 * h, wo = 2 * w;
  const long long total = (long long)n * ho * wo;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % wo), y = (int)((pix / wo) % ho);
  const int b = (int)(pix / ((long long)wo * ho));
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_src(y, 0.5f, h, y0, y1, ly);
  bilin_src(x, 0.5f, w, x0, x1, lx);
  const float2* ib = reinterpret_cast<const float2*>(flow) + (size_t)b * h * w;
  const float2 v00 = __ldg(ib + (size_t)y0 * w + x0), v01 = __ldg(ib + (size_t)y0 * w + x1);
  const float2 v10 = __ldg(ib + (size_t)y1 * w + x0), v11 = __ldg(ib + (size_t)y1 * w + x1);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float fx = (hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x)) * 2.f;
  const float fy = (hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y)) * 2.f;
  reinterpret_cast<float2*>(out_f32)[pix] = make_float2(fx, fy);
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  __nv_bfloat162 p = __floats2bfloat162_rn(fx, fy);
  o.x = *reinterpret_cast<uint32_t*>(&p);
  reinterpret_cast<uint4*>(out_bf8)[pix] = o;
}

static inline unsigned grid1d(long long total) { return (unsigned)((total + 255) / 256); }

int launch_flow_warp_bf16(const crfp_warp_desc& d, cudaStream_t st) {
  if (d.c % 8 || ((d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 7)) return CRFP_ERR_BAD_SHAPE;
  const long long total = (long long)d.n * d.h * d.w * (d.c / 8);
  if (total == 0) return CRFP_OK;
  launch_k(flow_warp_bf16_kernel, dim3(grid1d(total)), dim3(256), (size_t)(0), st, d);
  return check_launch();
}

int launch_flow_up2_dual(int n, int h, int w, const float* flow, float* out_f32, void* out_bf8, cudaStream_t st) {
  const long long total = (long long)n * 4 * h * w;
  launch_k(flow_up2_dual_kernel, dim3(grid1d(total)), dim3(256), (size_t)(0), st, n, h, w, flow, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf8));
  return check_launch();
}

int launch_flow_warp_l1(int n, int h, int w, const float* flow, const float* P, float* P_w, const float* state, float* cur0,
                        float* cur1, float* cur2, cudaStream_t st) {
  const long long total = (long long)n * h * w * 14;
  if (total == 0) return CRFP_OK;
  launch_k(flow_warp_l1_kernel, dim3(grid1d(total)), dim3(256), (size_t)(0), st, n, h, w, flow, P, P_w, state, cur0, cur1, cur2);
  return check_launch();
}

int launch_flow_warp(const crfp_warp_desc& d, cudaStream_t st) {
  if (d.c % 4 || ((d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 3)) return CRFP_ERR_BAD_SHAPE;
  const long long total = (long long)d.n * d.h * d.w * (d.c / 4);
  if (total == 0) return CRFP_OK;
  launch_k(flow_warp_kernel, dim3(grid1d(total)), dim3(256), (size_t)(0), st, d);
  return check_launch();
}

// ---- fovea paste: what the data loader does on the host (fvs[:, y:y+fv, x:x+fv] = patch, mks[...] = 1) for all
// frames of a clip in one launch; `clear` = zero the rectangles instead (the previous call's positions)
__global__ void fovea_paste_kernel(const float* __restrict__ patch, const int32_t* __restrict__ coords, int fv, int H, int W,
                                   float* __restrict__ fvs, uint8_t* __restrict__ mks, int clear) {
  const int f = blockIdx.y;   // frame index over n * t
  const int y0 = min(max(coords[2 * f], 0), H - fv), x0 = min(max(coords[2 * f + 1], 0), W - fv);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fv * fv; i += gridDim.x * blockDim.x) {
    const int dy = i / fv, dx = i - dy * fv;
    const size_t o = (size_t)(y0 + dy) * W + (x0 + dx);
    mks[(size_t)f * H * W + o] = clear ? 0 : 1;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      fvs[((size_t)f * 3 + c) * H * W + o] = clear ? 0.f : patch[((size_t)f * 3 + c) * fv * fv + i];
  }
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_flow_warp_fwd(const crfp_warp_desc* d, crfp_stream stream) {
  if (!d || !d->x || !d->flow || !d->out) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0 || d->c <= 0) return CRFP_ERR_BAD_SHAPE;
  return launch_flow_warp(*d, (cudaStream_t)stream);
}

extern "C" int crfp_flow_warp_bf16_fwd(const crfp_warp_desc* d, crfp_stream stream) {
  if (!d || !d->x || !d->flow || !d->out) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0 || d->c <= 0 || d->border) return CRFP_ERR_BAD_SHAPE;
  return launch_flow_warp_bf16(*d, (cudaStream_t)stream);
}

extern "C" int crfp_flow_warp_indices(int n, int h, int w, const float* flow, int32_t* x0, int32_t* y0,
                                      crfp_stream stream) {
  if (!flow || !x0 || !y0) return CRFP_ERR_NULL;
  const long long total = (long long)n * h * w;
  if (total <= 0) return CRFP_ERR_BAD_SHAPE;
  flow_warp_indices_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(n, h, w, flow, x0, y0);
  return check_launch();
}

extern "C" int crfp_resize_bilinear(int n, int hin, int win, int c, const float* in, int hout, int wout, float rscale_h,
                                    float rscale_w, float mul, float* out, crfp_stream stream) {
  if (!in || !out) return CRFP_ERR_NULL;
  if (n <= 0 || hin <= 0 || win <= 0 || c <= 0 || hout <= 0 || wout <= 0) return CRFP_ERR_BAD_SHAPE;
  const long long total = (long long)n * hout * wout * c;
  if (c % 4 == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
    launch_k(resize_bilinear_vec_kernel<float4>, dim3(grid1d(total / 4)), dim3(256), (size_t)(0), (cudaStream_t)stream, n, hin, win, c / 4,
             reinterpret_cast<const float4*>(in), hout, wout, rscale_h, rscale_w, mul, reinterpret_cast<float4*>(out));
    return check_launch();
  }
  if (c % 2 == 0 && (((uintptr_t)in | (uintptr_t)out) & 7) == 0) {
    launch_k(resize_bilinear_vec_kernel<float2>, dim3(grid1d(total / 2)), dim3(256), (size_t)(0), (cudaStream_t)stream, n, hin, win, c / 2,
             reinterpret_cast<const float2*>(in), hout, wout, rscale_h, rscale_w, mul, reinterpret_cast<float2*>(out));
    return check_launch();
  }
  launch_k(resize_bilinear_kernel, dim3(grid1d(total)), dim3(256), (size_t)(0), (cudaStream_t)stream, n, hin, win, c, in, hout, wout, rscale_h,
                                                                         rscale_w, mul, out);
  return check_launch();
}

extern "C" int crfp_fovea_paste(const float* patch, const int32_t* coords, int frames, int fv, int H, int W, float* fvs,
                                uint8_t* mks, int clear, crfp_stream stream) {
  if (!coords || !fvs || !mks || (!clear && !patch)) return CRFP_ERR_NULL;
  if (frames < 0 || fv <= 0 || fv > H || fv > W) return CRFP_ERR_BAD_SHAPE;
  if (frames == 0) return CRFP_OK;
  int bx = (fv * fv + 255) / 256;
  if (bx > 64) bx = 64;
  launch_k(fovea_paste_kernel, dim3(bx, frames), dim3(256), (size_t)(0), (cudaStream_t)stream, patch, coords, fv, H, W, fvs, mks, clear);
  return check_launch();
}

extern "C" int crfp_avgpool2(int n, int hin, int win, int c, const float* in, float* out, crfp_stream stream) {
  if (!in || !out) return CRFP_ERR_NULL;
  if (n <= 0 || hin < 2 || win < 2 || c <= 0) return CRFP_ERR_BAD_SHAPE;
  const long long total = (long long)n * (hin / 2) * (win / 2) * c;
  launch_k(avgpool2_kernel, dim3(grid1d(total)), dim3(256), (size_t)(0), (cudaStream_t)stream, n, hin, win, c, in, out);
  return check_launch();
}

extern "C" int crfp_nchw_to_nhwc(int n, int c, int h, int w, const float* in, long long in_image_stride, int cpad,
                                 float* out, crfp_stream stream) {
  if (!in || !out) return CRFP_ERR_NULL;
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || cpad < c) return CRFP_ERR_BAD_SHAPE;
  launch_k(nchw_to_nhwc_kernel, dim3(grid1d((long long)n * h * w)), dim3(256), (size_t)(0), (cudaStream_t)stream, n, c, h, w, in, in_image_stride,
                                                                                    cpad, out);
  return check_launch();
}

extern "C" int crfp_nhwc_to_nchw(int n, int c, int h, int w, const float* in, int in_cstride, int in_coffset,
                                 float* out, long long out_image_stride, crfp_stream stream) {
  if (!in || !out) return CRFP_ERR_NULL;
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return CRFP_ERR_BAD_SHAPE;
  launch_k(nhwc_to_nchw_kernel, dim3(grid1d((long long)n * h * w * c)), dim3(256), (size_t)(0), (cudaStream_t)stream, n, c, h, w, in, in_cstride,
                                                                                        in_coffset, out,
                                                                                        out_image_stride);
  return check_launch();
}
