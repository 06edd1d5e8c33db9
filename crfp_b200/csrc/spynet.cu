// Operators of the legacy SPyNet flow pyramid (/root/reference/model/CRFP.py:554-741, 145-152) that
// the CRFP hot path does not already provide: the 7x7 `conv(act(x))` layers, the align_corners=True bilinear x2 of the
// flow, and a per-channel affine (ImageNet normalisation, per-axis flow rescale).  The pyramid's other steps reuse the
// library's avg-pool, align_corners=False resize and border-mode flow_warp kernels.
//
// Secondary path (no shipped model calls SPyNet, only the legacy runtime classes): plain, sync-free SIMT kernels over
// dense fp32 NHWC tensors — one thread per output element — which also compile for the host emulation of the CPU test
// suite (CRFP_HOST_EMU, tests/tools/hostemu).
#ifdef CRFP_HOST_EMU
#include "cuda_shim.h"
#else
#include "common.cuh"
#define CRFP_LAUNCH(kernel, grid, block, st, ...) kernel<<<(grid), (block), 0, (st)>>>(__VA_ARGS__)
#endif

namespace crfp {

// out[b,y,x,co] = bias[co] + sum_{ky,kx,ci} W[ky*k+kx][ci][co] * act(x[b, y+ky-k/2, x+kx-k/2, ci]) (+ residual[b,y,x,co]);
// act = ReLU when relu_in (applied to the INPUT, CRFP.py:151-152), zero padding.
// Threads of a warp: consecutive co of one pixel first (weights coalesced, activations warp-broadcast).
__global__ void __launch_bounds__(256) conv_kxk_kernel(int n, int h, int w, int cin, int cout, int k, int relu_in,
                                                       const float* __restrict__ x, const float* __restrict__ wt,
                                                       const float* __restrict__ bias, const float* __restrict__ residual,
                                                       float* __restrict__ out) {
  const long long total = (long long)n * h * w * cout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % cout);
  const long long pix = idx / cout;
  const int px = (int)(pix % w);
  const int py = (int)((pix / w) % h);
  const long long b = pix / ((long long)w * h);
  const int r = k / 2;
  float acc = bias[co];
  for (int ky = 0; ky < k; ++ky) {
    const int yi = py + ky - r;
    if (yi < 0 || yi >= h) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int xi = px + kx - r;
      if (xi < 0 || xi >= w) continue;
      const float* xp = x + ((b * h + yi) * w + xi) * cin;
      const float* wp = wt + (long long)(ky * k + kx) * cin * cout + co;
#pragma unroll 4
      for (int ci = 0; ci < cin; ++ci) {
        float v = xp[ci];
        if (relu_in) v = v > 0.f ? v : 0.f;
        acc += v * wp[(long long)ci * cout];
      }
    }
  }
  if (residual != nullptr) acc += residual[idx];
  out[idx] = acc;
}

// F.interpolate(mode='bilinear', align_corners=True): src = dst * (in-1)/(out-1) (0 when out == 1); out = value * mul
__global__ void __launch_bounds__(256) resize_bilinear_ac_kernel(int n, int hin, int win, int c, const float* __restrict__ in,
                                                                 int hout, int wout, float mul, float* __restrict__ out) {
  const long long total = (long long)n * hout * wout * c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  const long long pix = idx / c;
  const int x = (int)(pix % wout);
  const int y = (int)((pix / wout) % hout);
  const long long b = pix / ((long long)wout * hout);
  const float sh = hout > 1 ? (float)(hin - 1) / (float)(hout - 1) : 0.f;
  const float sw = wout > 1 ? (float)(win - 1) / (float)(wout - 1) : 0.f;
  const float fy = sh * (float)y, fx = sw * (float)x;
  int y0 = (int)fy, x0 = (int)fx;
  if (y0 > hin - 1) y0 = hin - 1;
  if (x0 > win - 1) x0 = win - 1;
  const int y1 = y0 + (y0 < hin - 1 ? 1 : 0), x1 = x0 + (x0 < win - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* ib = in + b * hin * win * c + ch;
  const float v00 = ib[((long long)y0 * win + x0) * c], v01 = ib[((long long)y0 * win + x1) * c];
  const float v10 = ib[((long long)y1 * win + x0) * c], v11 = ib[((long long)y1 * win + x1) * c];
  out[idx] = (hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11)) * mul;
}

// out[p][ch] = ((in[p][ch] - sub[ch]) / div[ch]) * mul[ch] for ch < c_in; channels [c_in, c_out) of out are zero
__global__ void __launch_bounds__(256) channel_affine_kernel(long long npix, int c_in, int c_out, const float* __restrict__ in,
                                                             const float* __restrict__ sub, const float* __restrict__ div,
                                                             const float* __restrict__ mul, float* __restrict__ out) {
  const long long total = npix * c_out;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c_out);
  const long long p = idx / c_out;
  out[idx] = ch < c_in ? ((in[p * c_in + ch] - sub[ch]) / div[ch]) * mul[ch] : 0.f;
}

static inline unsigned spy_blocks(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_conv_kxk_fwd(int n, int h, int w, int cin, int cout, int k, int relu_in, const float* x,
                                 const float* weight, const float* bias, const float* residual, float* out,
                                 crfp_stream stream) {
  if (n < 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || k < 1 || k > 15 || (k & 1) == 0) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!x || !weight || !bias || !out) return CRFP_ERR_NULL;
  const long long total = (long long)n * h * w * cout;
  CRFP_LAUNCH(conv_kxk_kernel, dim3(spy_blocks(total)), dim3(256), (cudaStream_t)stream, n, h, w, cin, cout, k, relu_in, x,
              weight, bias, residual, out);
  return check_launch();
}

extern "C" int crfp_resize_bilinear_ac(int n, int hin, int win, int c, const float* in, int hout, int wout, float mul,
                                       float* out, crfp_stream stream) {
  if (n < 0 || hin <= 0 || win <= 0 || c <= 0 || hout <= 0 || wout <= 0) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!in || !out) return CRFP_ERR_NULL;
  const long long total = (long long)n * hout * wout * c;
  CRFP_LAUNCH(resize_bilinear_ac_kernel, dim3(spy_blocks(total)), dim3(256), (cudaStream_t)stream, n, hin, win, c, in, hout,
              wout, mul, out);
  return check_launch();
}

extern "C" int crfp_channel_affine(long long npix, int c_in, int c_out, const float* in, const float* sub, const float* div,
                                   const float* mul, float* out, crfp_stream stream) {
  if (npix < 0 || c_in <= 0 || c_out < c_in) return CRFP_ERR_BAD_SHAPE;
  if (npix == 0) return CRFP_OK;
  if (!in || !sub || !div || !mul || !out) return CRFP_ERR_NULL;
  CRFP_LAUNCH(channel_affine_kernel, dim3(spy_blocks(npix * c_out)), dim3(256), (cudaStream_t)stream, npix, c_in, c_out, in,
              sub, div, mul, out);
  return check_launch();
}
