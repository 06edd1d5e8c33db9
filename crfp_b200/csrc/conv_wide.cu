// fp32 SIMT direct 3x3 convolution for "wide" layers (cout > 4): NHWC in, channel-concatenated sources,
// fused bias / activation / residual / DCN-head epilogue / channel-split or pixel-shuffle stores.
//
// Replaces nn.Conv2d + LeakyReLU/ReLU + torch.cat + torch.chunk + F.pixel_shuffle + pixel_unshuffle of
// /root/reference/model/CRFP.py:154-193, 239-279, 303-317, 433-552.
//
// Tiling: one CTA = 32 (x) by 2*RPT (y) output pixels by 32 output channels.  256 threads = 8 warps:
// warp -> (row group, 8-channel group), lane -> x.  Each thread owns RPT rows x 8 channels = 8*RPT fp32
// accumulators and walks the packed input channels in chunks of 8 staged (transposed to [ci][y][x]) in
// shared memory together with the chunk's [9][8][32] weights: per (ci, kx) a thread issues RPT+2
// conflict-free LDS.32 for its input column and 6 warp-broadcast LDS.128 for the weights, then 24*RPT FFMA.
#include "common.cuh"

namespace crfp {

constexpr int TW = 32;          // tile width (pixels) = lanes
constexpr int CO_T = 32;        // output channels per CTA
constexpr int CI_CH = 8;        // packed input channels per smem stage
constexpr int SROW = TW + 2 + 2;  // smem row pitch (34 used, padded to 36)

template <int RPT>
struct WideSmem {
  static constexpr int TH = 2 * RPT;
  float in[CI_CH][TH + 2][SROW];
  float w[9][CI_CH][CO_T];
};

template <int RPT>
__global__ void __launch_bounds__(256, 2) conv_wide_kernel(const ConvParams P) {
  pdl_trigger();
  pdl_wait();
  constexpr int TH = 2 * RPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WideSmem<RPT>& S = *reinterpret_cast<WideSmem<RPT>*>(smem_raw);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int cg = warp & 3;   // 8-channel group inside the 32-channel CTA tile
  const int rg = warp >> 2;  // row group
  const int co_tiles = P.cout_packed / CO_T;
  const int n = blockIdx.z / co_tiles;
  const int co_base = (blockIdx.z - n * co_tiles) * CO_T;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

  // accumulators as channel pairs: packed fp32 FMAs (FFMA2, sm_100) with the input value as the broadcast operand
  float2 acc2[RPT][4];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc2[r][c] = make_float2(0.f, 0.f);

  const int nchunks = P.cin_packed / CI_CH;
  for (int ch = 0; ch < nchunks; ++ch) {
    // ---- stage inputs: (TH+2) x (TW+2) halo pixels x 2 quads, transposed to [ci][y][x]
    constexpr int NPIX = (TH + 2) * (TW + 2);
    for (int it = tid; it < NPIX * 2; it += 256) {
      const int q = it & 1, pix = it >> 1;
      const int py = pix / (TW + 2), px = pix - py * (TW + 2);
      const float4 v = load_quad_fg(P, ch * 2 + q, n, y0 + py - 1, x0 + px - 1);
      S.in[q * 4 + 0][py][px] = v.x;
      S.in[q * 4 + 1][py][px] = v.y;
      S.in[q * 4 + 2][py][px] = v.z;
      S.in[q * 4 + 3][py][px] = v.w;
    }
    // ---- stage weights: [9][8][32] slice of [9][cin_packed][cout_packed]
    for (int it = tid; it < 9 * CI_CH * (CO_T / 4); it += 256) {
      const int c4 = it & 7, row = it >> 3;  // row = tap*8 + ci
      const int tap = row >> 3, ci = row & 7;
      const float4 v = __ldg(reinterpret_cast<const float4*>(
          P.weight + ((size_t)tap * P.cin_packed + ch * CI_CH + ci) * P.cout_packed + co_base + c4 * 4));
      *reinterpret_cast<float4*>(&S.w[tap][ci][c4 * 4]) = v;
    }
    __syncthreads();

#pragma unroll 2
    for (int ci = 0; ci < CI_CH; ++ci) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float col[RPT + 2];
#pragma unroll
        for (int r = 0; r < RPT + 2; ++r) col[r] = S.in[ci][rg * RPT + r][lane + kx];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float4 w0 = *reinterpret_cast<const float4*>(&S.w[ky * 3 + kx][ci][cg * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&S.w[ky * 3 + kx][ci][cg * 8 + 4]);
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            const float2 a = make_float2(col[r + ky], col[r + ky]);
            acc2[r][0] = __ffma2_rn(a, make_float2(w0.x, w0.y), acc2[r][0]);
            acc2[r][1] = __ffma2_rn(a, make_float2(w0.z, w0.w), acc2[r][1]);
            acc2[r][2] = __ffma2_rn(a, make_float2(w1.x, w1.y), acc2[r][2]);
            acc2[r][3] = __ffma2_rn(a, make_float2(w1.z, w1.w), acc2[r][3]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  const int x = x0 + lane;
  const int cch = co_base + cg * 8;  // first conv output channel of this thread
  if (x >= P.w || cch >= P.cout) return;
  float bias[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(P.bias + cch));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(P.bias + cch + 4));
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
    bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  const float ps = (P.post_scale == 0.f) ? 1.f : P.post_scale;
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int y = y0 + rg * RPT + r;
    if (y >= P.h) break;
    const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = ((c & 1) ? acc2[r][c >> 1].y : acc2[r][c >> 1].x) + bias[c];
    if (P.act == CRFP_ACT_DCN_HEAD) {
      const float2 fl = __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2));
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cc = cch + c;
        if (cc < P.head_split)
          v[c] = P.head_mag * tanhf(v[c]) + ((cc & 1) ? fl.x : fl.y);
        else
          v[c] = sigmoidf_(v[c]);
      }
    } else if (P.act == CRFP_ACT_TANH256) {
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = tanhf(v[c]) * 256.f;
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = apply_act(v[c], P.act);
    }
    if (P.residual != nullptr) {
      const float* rp = P.residual + pix * P.res_cstride + P.res_coffset + cch;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (cch + c < P.cout) v[c] += __ldg(rp + c);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] *= ps;

    if (P.out_mode == CRFP_OUT_NHWC) {
      // the 8-channel group lies in one destination segment when segment sizes are multiples of 8;
      // otherwise fall back to per-channel routing
      int seg = 0, cl = cch;
      if (P.ndst > 1 && cch >= P.dst_c[0]) { seg = 1; cl = cch - P.dst_c[0]; }
      const bool whole = (cl + 8 <= P.dst_c[seg]) && (((P.dst_cstride[seg] | P.dst_coffset[seg]) & 3) == 0);
      if (P.out_bf16) {
        const bool whole8 = (cl + 8 <= P.dst_c[seg]) && (((P.dst_cstride[seg] | P.dst_coffset[seg]) & 7) == 0);
        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(P.dst[seg]) + pix * P.dst_cstride[seg] + P.dst_coffset[seg] + cl;
        if (whole8) {
          __nv_bfloat162 o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) o[c] = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
          *reinterpret_cast<uint4*>(ob) = *reinterpret_cast<uint4*>(o);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (cch + c < P.cout && cl + c < P.dst_c[seg]) ob[c] = __float2bfloat16(v[c]);
        }
      } else if (whole) {
        float* op = P.dst[seg] + pix * P.dst_cstride[seg] + P.dst_coffset[seg] + cl;
        *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = cch + c;
          if (cc >= P.cout) break;
          int sg = 0, l = cc;
          if (P.ndst > 1 && cc >= P.dst_c[0]) { sg = 1; l = cc - P.dst_c[0]; }
          P.dst[sg][pix * P.dst_cstride[sg] + P.dst_coffset[sg] + l] = v[c];
        }
      }
    } else {  // pixel shuffle: conv channel o*r*r + dy*r + dx -> (y*r+dy, x*r+dx, o)
      const int r_ = P.shuffle_r, rr = r_ * r_;
      const int Ho = P.h * r_, Wo = P.w * r_;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cc = cch + c;
        if (cc >= P.cout) break;
        const int o = cc / rr, sub = cc - o * rr;
        const int dy = sub / r_, dx = sub - dy * r_;
        const size_t opix = ((size_t)n * Ho + (y * r_ + dy)) * (size_t)Wo + (x * r_ + dx);
        if (P.out_bf16)
          reinterpret_cast<__nv_bfloat16*>(P.dst[0])[opix * P.dst_cstride[0] + P.dst_coffset[0] + o] = __float2bfloat16(v[c]);
        else
          P.dst[0][opix * P.dst_cstride[0] + P.dst_coffset[0] + o] = v[c];
      }
    }
  }
}

int launch_conv_wide(const ConvParams& p, cudaStream_t st) {
  if (p.cout_packed % CO_T != 0 || p.cin_packed % CI_CH != 0) return CRFP_ERR_BAD_SHAPE;
  const int co_tiles = p.cout_packed / CO_T;
  const long long blocks8 = (long long)ceil_div(p.w, TW) * ceil_div(p.h, 16) * p.n * co_tiles;
  const bool small = blocks8 < 2 * 148;
  if (!small) {
    dim3 grid(ceil_div(p.w, TW), ceil_div(p.h, 16), p.n * co_tiles);
    size_t smem = sizeof(WideSmem<8>);
    static thread_local bool attr8 = false;
    if (!attr8) {
      cudaFuncSetAttribute(conv_wide_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr8 = true;
    }
    launch_k(conv_wide_kernel<8>, dim3(grid), dim3(256), (size_t)(smem), st, p);
  } else {
    dim3 grid(ceil_div(p.w, TW), ceil_div(p.h, 8), p.n * co_tiles);
    size_t smem = sizeof(WideSmem<4>);
    static thread_local bool attr4 = false;
    if (!attr4) {
      cudaFuncSetAttribute(conv_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr4 = true;
    }
    launch_k(conv_wide_kernel<4>, dim3(grid), dim3(256), (size_t)(smem), st, p);
  }
  return check_launch();
}

}  // namespace crfp
