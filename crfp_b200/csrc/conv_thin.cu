// fp32 SIMT 3x3 convolution for thin layers (cout <= 4): the 4-channel HR planes of CRFP
// (forward_resblocks_3, dcn_3 block / fuse / heads, encoder_hr, conv_tttf, conv_last) and FNet's 32->2 head.
// These are HBM/L1-bandwidth bound (36..90 MAC per 16..40 B), so: one thread per output pixel, float4
// channel-quads loaded straight from global (NHWC: one LDG.128 per tap per quad, neighbours hit L1),
// weights broadcast from shared memory, every elementwise consumer fused into the epilogue:
//   EPI_STD      bias + LeakyReLU/ReLU/tanh*256/DCN-head + residual + post_scale
//   EPI_BLEND    conv_tttf + fovea blend + LeakyReLU:  S = lrelu(m*F + (1-m)*S)      (model/CRFP.py:1672-1675)
//   EPI_OUT_NCHW conv_last + bilinear x8 base of the LR frame, planar NCHW store     (model/CRFP.py:1678-1683)
#include "common.cuh"

namespace crfp {

__global__ void __launch_bounds__(256) conv_thin_kernel(const ConvParams P) {
  extern __shared__ __align__(16) float s_w[];  // [9][cin_packed][4]
  const int nw4 = 9 * P.cin_packed;
  for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < nw4; i += blockDim.x * blockDim.y)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(P.weight) + i);
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
  float v[4] = {a0 + b.x, a1 + b.y, a2 + b.z, a3 + b.w};
  const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;

  if (P.epi == EPI_BLEND) {
    const float m = P.mask[(size_t)n * P.mask_clip_stride + (size_t)y * P.w + x] ? 1.f : 0.f;
    const float4 so = __ldg(reinterpret_cast<const float4*>(P.blend_old + pix * 4));
    float4 o;
    o.x = lrelu01(m * v[0] + (1.f - m) * so.x);
    o.y = lrelu01(m * v[1] + (1.f - m) * so.y);
    o.z = lrelu01(m * v[2] + (1.f - m) * so.z);
    o.w = lrelu01(m * v[3] + (1.f - m) * so.w);
    *reinterpret_cast<float4*>(P.dst[0] + pix * 4) = o;
    return;
  }
  if (P.epi == EPI_OUT_NCHW) {
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
    for (int c = 0; c < P.out_planes; ++c) ob[c * plane] = v[c] + bs[c];
    return;
  }

  if (P.act == CRFP_ACT_DCN_HEAD) {
    const float2 fl = __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < P.head_split)
        v[c] = P.head_mag * tanhf(v[c]) + ((c & 1) ? fl.x : fl.y);
      else
        v[c] = sigmoidf_(v[c]);
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
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < P.cout) v[c] += __ldg(rp + c);
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
    if (P.dst_c[0] > P.cout)
      for (int c = P.cout; c < P.dst_c[0] && c < 4; ++c) op[c] = 0.f;
  }
}

int launch_conv_thin(const ConvParams& p, cudaStream_t st) {
  if (p.cout > 4 || p.cout_packed != 4) return CRFP_ERR_BAD_SHAPE;
  if (p.out_mode != CRFP_OUT_NHWC && p.epi == EPI_STD) return CRFP_ERR_UNSUPPORTED;
  dim3 block(32, 8);
  dim3 grid(ceil_div(p.w, 32), ceil_div(p.h, 8), p.n);
  const size_t smem = (size_t)9 * p.cin_packed * 4 * sizeof(float);
  if (smem > 48 * 1024) return CRFP_ERR_UNSUPPORTED;
  conv_thin_kernel<<<grid, block, smem, st>>>(p);
  return check_launch();
}

int launch_conv(const ConvParams& p, cudaStream_t st) {
  if (p.cout_packed == 4) return launch_conv_thin(p, st);
  return launch_conv_wide(p, st);
}

}  // namespace crfp
