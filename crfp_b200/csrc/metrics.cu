// GPU-side input synthesis and evaluation metrics.
//
//  crfp_fovea_from_gt   `fovea_generator`'s per-pixel work (/root/reference/dataset/reds.py:190-226): for every frame of a
//                       clip, mask = 1 inside the frame's rectangle, fvs = GT * mask — one pass over the clip instead of the
//                       reference's per-frame zeros_like + slice assignment + multiply on the CPU.
//  crfp_psnr_ssim       `calc_psnr_and_ssim_cuda` (/root/reference/utils.py:165-254) fused: the five 11x11 Gaussian moments
//                       (mu1, mu2, E[x^2], E[y^2], E[xy]; sigma 1.5, zero padding) of a 32x32 tile are built separably in
//                       shared memory, the SSIM map and the squared error are multiplied by the mask and reduced to ONE
//                       partial triple per CTA (sum ssim*m, sum err^2*m, sum m) — the reference runs 5 depthwise convs + ~15
//                       pointwise kernels over full planes.  Partials are written per CTA (no atomics: deterministic) and the
//                       host adds them in float64.
#include "common.cuh"

namespace crfp {

__global__ void __launch_bounds__(256) fovea_from_gt_kernel(int frames, int c, int H, int W, const float* __restrict__ gt,
                                                            const int32_t* __restrict__ rects, float* __restrict__ fvs,
                                                            uint8_t* __restrict__ mks) {
  const long long hw = (long long)H * W;
  const long long total = (long long)frames * hw;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = (int)(idx / hw);
  const long long p = idx - (long long)f * hw;
  const int y = (int)(p / W), x = (int)(p - (long long)y * W);
  const int4 r = __ldg(reinterpret_cast<const int4*>(rects) + f);   // y0, x0, y1, x1
  const bool in = y >= r.x && y < r.z && x >= r.y && x < r.w;
  mks[idx] = in ? 1 : 0;
  for (int ch = 0; ch < c; ++ch) {
    const long long o = ((long long)f * c + ch) * hw + p;
    fvs[o] = in ? __ldg(gt + o) : 0.f;
  }
}

// frames as the reference SAVES them (trainer.py:446-474, 535-537): (sr * 255).clip(0, 255).round() -> uint8; 16 pixels per
// thread (4 x LDG.128 -> one STG.128), round-half-to-even like torch.round
__global__ void __launch_bounds__(256) quantize_u8_kernel(long long count16, long long count, const float* __restrict__ in,
                                                          uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto q = [](float v) -> uint32_t { return (uint32_t)__float2int_rn(fminf(fmaxf(v * 255.f, 0.f), 255.f)); };
  if (i < count16) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i * 4 + k);
      w[k] = q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  } else if (i == count16) {
    for (long long j = count16 * 16; j < count; ++j) out[j] = (uint8_t)q(in[j]);   // tail (< 16 elements)
  }
}

constexpr int ST = 32;          // output tile
constexpr int SR = 5;           // window radius (11 taps)
constexpr int SW = ST + 2 * SR; // 42

struct SsimWin { float g[11]; };

// grid: (tiles_x, tiles_y, B*C); partial[(b*C + ch) * tiles + tile] = {sum ssim*m, sum err^2*m, sum m}
__global__ void __launch_bounds__(256) psnr_ssim_kernel(int C, int H, int W, const float* __restrict__ a, const float* __restrict__ b,
                                                        const float* __restrict__ maskf, const uint8_t* __restrict__ masku,
                                                        const SsimWin win, float* __restrict__ partial) {
  __shared__ float sa[SW][SW + 1], sb[SW][SW + 1];
  __shared__ float hm[5][SW][ST + 1];   // horizontally filtered: a, b, a*a, b*b, a*b
  __shared__ float red[3][8];
  const int tid = threadIdx.x;
  const int plane = blockIdx.z, bimg = plane / C;
  const int x0 = blockIdx.x * ST, y0 = blockIdx.y * ST;
  const float* pa = a + (size_t)plane * H * W;
  const float* pb = b + (size_t)plane * H * W;
  for (int i = tid; i < SW * SW; i += 256) {
    const int r = i / SW, cidx = i - r * SW;
    const int y = y0 + r - SR, x = x0 + cidx - SR;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    sa[r][cidx] = in ? __ldg(pa + (size_t)y * W + x) : 0.f;   // F.conv2d zero padding
    sb[r][cidx] = in ? __ldg(pb + (size_t)y * W + x) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < SW * ST; i += 256) {
    const int r = i / ST, cidx = i - r * ST;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float va = sa[r][cidx + k], vb = sb[r][cidx + k], g = win.g[k];
      m1 = fmaf(g, va, m1); m2 = fmaf(g, vb, m2);
      s11 = fmaf(g, va * va, s11); s22 = fmaf(g, vb * vb, s22); s12 = fmaf(g, va * vb, s12);
    }
    hm[0][r][cidx] = m1; hm[1][r][cidx] = m2; hm[2][r][cidx] = s11; hm[3][r][cidx] = s22; hm[4][r][cidx] = s12;
  }
  __syncthreads();
  float acc_s = 0.f, acc_e = 0.f, acc_m = 0.f;
  for (int i = tid; i < ST * ST; i += 256) {
    const int r = i / ST, cidx = i - r * ST;
    const int y = y0 + r, x = x0 + cidx;
    if (y >= H || x >= W) continue;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = win.g[k];
      m1 = fmaf(g, hm[0][r + k][cidx], m1); m2 = fmaf(g, hm[1][r + k][cidx], m2);
      s11 = fmaf(g, hm[2][r + k][cidx], s11); s22 = fmaf(g, hm[3][r + k][cidx], s22); s12 = fmaf(g, hm[4][r + k][cidx], s12);
    }
    const float mu11 = m1 * m1, mu22 = m2 * m2, mu12 = m1 * m2;
    const float v1 = s11 - mu11, v2 = s22 - mu22, v12 = s12 - mu12;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float ssim = ((2.f * mu12 + C1) * (2.f * v12 + C2)) / ((mu11 + mu22 + C1) * (v1 + v2 + C2));
    const size_t mi = (size_t)bimg * H * W + (size_t)y * W + x;
    const float m = maskf != nullptr ? __ldg(maskf + mi) : (masku != nullptr ? (masku[mi] ? 1.f : 0.f) : 1.f);
    const float d = sa[r + SR][cidx + SR] - sb[r + SR][cidx + SR];
    acc_s += ssim * m; acc_e += d * d * m; acc_m += m;
  }
  // CTA reduction in a fixed order (deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
    acc_e += __shfl_xor_sync(0xffffffffu, acc_e, o);
    acc_m += __shfl_xor_sync(0xffffffffu, acc_m, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = acc_s; red[1][tid >> 5] = acc_e; red[2][tid >> 5] = acc_m; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
    for (int wq = 0; wq < 8; ++wq) s += red[tid][wq];
    const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    partial[((size_t)plane * gridDim.x * gridDim.y + tile) * 3 + tid] = s;
  }
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_fovea_from_gt(const float* gt, const int32_t* rects, int frames, int c, int H, int W, float* fvs,
                                  uint8_t* mks, crfp_stream stream) {
  if (!gt || !rects || !fvs || !mks) return CRFP_ERR_NULL;
  if (frames < 0 || c <= 0 || H <= 0 || W <= 0) return CRFP_ERR_BAD_SHAPE;
  if (frames == 0) return CRFP_OK;
  if ((uintptr_t)rects & 15) return CRFP_ERR_BAD_SHAPE;
  const long long total = (long long)frames * H * W;
  fovea_from_gt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(frames, c, H, W, gt, rects, fvs, mks);
  return check_launch();
}

extern "C" int crfp_quantize_u8(const float* in, uint8_t* out, long long count, crfp_stream stream) {
  if (!in || !out) return CRFP_ERR_NULL;
  if (count < 0 || ((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return CRFP_ERR_BAD_SHAPE;
  if (count == 0) return CRFP_OK;
  const long long c16 = count / 16;
  quantize_u8_kernel<<<(unsigned)((c16 + 1 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c16, count, in, out);
  return check_launch();
}

extern "C" int crfp_psnr_ssim_tiles(int H, int W, int32_t* tiles_x, int32_t* tiles_y) {
  if (!tiles_x || !tiles_y || H <= 0 || W <= 0) return CRFP_ERR_BAD_SHAPE;
  *tiles_x = ceil_div(W, ST); *tiles_y = ceil_div(H, ST);
  return CRFP_OK;
}

extern "C" int crfp_psnr_ssim(int B, int C, int H, int W, const float* img1, const float* img2, const float* mask_f32,
                              const uint8_t* mask_u8, const float* window11, float* partial, crfp_stream stream) {
  if (!img1 || !img2 || !window11 || !partial) return CRFP_ERR_NULL;
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (long long)B * C > 65535) return CRFP_ERR_BAD_SHAPE;
  SsimWin win;
  for (int i = 0; i < 11; ++i) win.g[i] = window11[i];   // HOST pointer: the 11 taps ride in the kernel parameters
  dim3 grid(ceil_div(W, ST), ceil_div(H, ST), B * C);
  psnr_ssim_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, H, W, img1, img2, mask_f32, mask_u8, win, partial);
  return check_launch();
}
