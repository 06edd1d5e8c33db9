// DCNv2 modulated deformable 3x3 convolution — the "align kernel" of CRFP.
//
// Replaces dcn_v2.DCNv2.forward(input, offset, mask) (/root/reference/model/CRFP.py:318-320,350; external CUDA
// extension jinfagang/DCNv2_latest: modulated_deformable_im2col + cuBLAS SGEMM with a `columns` buffer of
// 9*C floats per pixel round-tripping HBM).  Here the im2col never leaves the SM:
//
//  L1 kernel (C=32, dg=8, cout=32): CTA = 8x8 output pixels.
//    phase 1  gather: each (pixel, group, tap) sample = 4 bilinear corners x one float4 (the group's 4 channels
//             are contiguous in NHWC) -> modulated column tile col[64][288] in shared memory
//    phase 2  contraction col[64x288] x W[288x32] in fp32 FFMA from shared memory, K split over the 8 warps,
//             partial sums reduced through shared memory, + bias, NHWC float4 store.
//  HR kernel (C=4, dg=1, shared offsets): one thread per pixel, 9 taps x 4 corners float4 gathers (L1-resident
//             neighbourhood), 144 FFMA, float4 store.  HBM-bound by construction (44 B/px algorithmic).
//
// Sampling semantics follow DCNv2 / torchvision exactly: p = (y-1+i) + dy evaluated as one
// fp32 add; sample = 0 if p <= -1 or p >= size; corners outside the image contribute 0.
#include "common.cuh"

namespace crfp {

struct Corner {
  int y0, x0;
  float w00, w01, w10, w11;  // 0 where the corner is out of range or the sample is out of range
};

// position -> corner set: dcn_pos.cuh holds the one shared definition
__device__ __forceinline__ Corner dcn_corner(float py, float px, int H, int W) {
  Corner c;
  dcn_corner_w(py, px, H, W, c.y0, c.x0, c.w00, c.w01, c.w10, c.w11);
  return c;
}

__device__ __forceinline__ float4 dcn_sample4(const float* __restrict__ img, int cs, int W, const Corner& c) {
  // img points at channel 0 of the 4-channel group at pixel (0,0) of this image
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = img + ((long long)c.y0 * W + c.x0) * cs;
  if (c.w00 != 0.f) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.x += c.w00 * t.x; r.y += c.w00 * t.y; r.z += c.w00 * t.z; r.w += c.w00 * t.w; }
  if (c.w01 != 0.f) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + cs)); r.x += c.w01 * t.x; r.y += c.w01 * t.y; r.z += c.w01 * t.z; r.w += c.w01 * t.w; }
  if (c.w10 != 0.f) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + (long long)W * cs)); r.x += c.w10 * t.x; r.y += c.w10 * t.y; r.z += c.w10 * t.z; r.w += c.w10 * t.w; }
  if (c.w11 != 0.f) { const float4 t = __ldg(reinterpret_cast<const float4*>(p + (long long)W * cs + cs)); r.x += c.w11 * t.x; r.y += c.w11 * t.y; r.z += c.w11 * t.z; r.w += c.w11 * t.w; }
  return r;
}

// ------------------------------------------------------------------------------------------------ L1 kernel
constexpr int DT = 8;             // 8x8 pixel tile
constexpr int DPIX = DT * DT;     // 64
constexpr int DK = 288;           // 8 groups * 9 taps * 4 channels
constexpr int DKP = DK + 4;       // padded row pitch (292 % 32 == 4 -> conflict-free float4 column reads)
constexpr int DCO = 32;

__global__ void __launch_bounds__(256, 2) dcn_l1_kernel(const crfp_dcn_desc D) {
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                 // [288][32]
  float* s_col = smem + DK * DCO;    // [64][292]; reused as the K-split reduction scratch [8][64][32]

  const int tid = threadIdx.x;
  const int tiles_x = (D.w + DT - 1) / DT;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int n = blockIdx.y;
  const int x0 = tx * DT, y0 = ty * DT;

  pdl_trigger();
  for (int i = tid; i < DK * DCO / 4; i += 256)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(D.weight) + i);
  pdl_wait();

  // ---- phase 1: gather.  64 px * 72 (group,tap) samples, 18 rounds of 256 threads
  const float* img = D.x + (size_t)n * D.h * D.w * D.x_cstride + D.x_coffset;
#pragma unroll 2
  for (int s = tid; s < DPIX * 72; s += 256) {
    const int p = s / 72, gt = s - p * 72;
    const int y = y0 + (p >> 3), x = x0 + (p & 7);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < D.h && x < D.w) {
      const size_t pix = ((size_t)n * D.h + y) * (size_t)D.w + x;
      const float2 off = __ldg(reinterpret_cast<const float2*>(D.offset + pix * D.off_cstride + D.off_coffset + gt * 2));
      const float m = __ldg(D.mask + pix * D.mask_cstride + D.mask_coffset + gt);
      const int g = gt / 9, t = gt - g * 9;
      const int i = t / 3, j = t - i * 3;
      const Corner c = dcn_corner(dcn_pos(y, i, off.x), dcn_pos(x, j, off.y), D.h, D.w);
      if (D.dbg_y0 != nullptr) { D.dbg_y0[pix * 72 + gt] = c.y0; D.dbg_x0[pix * 72 + gt] = c.x0; }
      v = dcn_sample4(img + g * 4, D.x_cstride, D.w, c);
      v.x *= m; v.y *= m; v.z *= m; v.w *= m;
    }
    *reinterpret_cast<float4*>(s_col + p * DKP + gt * 4) = v;
  }
  __syncthreads();

  // ---- phase 2: col[64][288] x W[288][32], K split over 8 warps (36 = 9 float4 steps each)
  const int lane = tid & 31, ks = tid >> 5;
  const int pxg = lane & 7, cog = lane >> 3;  // thread owns pixels pxg + 8*i (i<8), channels cog*8..+7
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
#pragma unroll 1
  for (int k4 = 0; k4 < 9; ++k4) {
    const int k = ks * 36 + k4 * 4;
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(s_col + (pxg + 8 * i) * DKP + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (k + kk) * DCO + cog * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (k + kk) * DCO + cog * 8 + 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float av = (kk == 0) ? a[i].x : (kk == 1) ? a[i].y : (kk == 2) ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
        acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
        acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
        acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
      }
    }
  }
  __syncthreads();  // everyone is done reading s_col; reuse it as scratch[ks][p][co]
  float* scr = s_col;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* d = scr + ((ks * DPIX) + (pxg + 8 * i)) * DCO + cog * 8;
    *reinterpret_cast<float4*>(d) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
  __syncthreads();
  // 64 px * 32 co = 2048 outputs = 512 float4; 2 per thread
  for (int o = tid; o < DPIX * DCO / 4; o += 256) {
    const int p = o >> 3, c4 = o & 7;
    float4 sum = __ldg(reinterpret_cast<const float4*>(D.bias) + c4);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const float4 t = *reinterpret_cast<const float4*>(scr + (s * DPIX + p) * DCO + c4 * 4);
      sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
    }
    const int y = y0 + (p >> 3), x = x0 + (p & 7);
    if (y < D.h && x < D.w) {
      const size_t pix = ((size_t)n * D.h + y) * (size_t)D.w + x;
      *reinterpret_cast<float4*>(D.out + pix * D.out_cstride + D.out_coffset + c4 * 4) = sum;
    }
  }
}

// ------------------------------------------------------------------------------------------------ HR kernel
// C = 4, dg = 1, cout = 4.  shared_taps: offset = [dy, dx], mask = [m] per pixel; otherwise 18 / 9 channels.
__global__ void __launch_bounds__(256) dcn_hr_kernel(const crfp_dcn_desc D) {
  __shared__ float4 s_w[36];  // [k = t*4 + c][co 0..3]
  __shared__ float4 s_b;
  const int tid = threadIdx.x + threadIdx.y * blockDim.x;
  pdl_trigger();
  if (tid < 36) s_w[tid] = __ldg(reinterpret_cast<const float4*>(D.weight) + tid);
  if (tid == 36) s_b = __ldg(reinterpret_cast<const float4*>(D.bias));
  pdl_wait();
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int n = blockIdx.z;
  if (x >= D.w || y >= D.h) return;
  const size_t pix = ((size_t)n * D.h + y) * (size_t)D.w + x;
  const float* offp = D.offset + pix * D.off_cstride + D.off_coffset;
  const float* mp = D.mask + pix * D.mask_cstride + D.mask_coffset;
  const float* img = D.x + (size_t)n * D.h * D.w * D.x_cstride + D.x_coffset;
  float dys = 0.f, dxs = 0.f, ms = 0.f;
  if (D.shared_taps) { dys = __ldg(offp); dxs = __ldg(offp + 1); ms = __ldg(mp); }
  float2 a01 = make_float2(s_b.x, s_b.y), a23 = make_float2(s_b.z, s_b.w);   // packed fp32 FMAs (FFMA2)
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int i = t / 3, j = t - i * 3;
    float dy = dys, dx = dxs, m = ms;
    if (!D.shared_taps) { dy = __ldg(offp + 2 * t); dx = __ldg(offp + 2 * t + 1); m = __ldg(mp + t); }
    const Corner c = dcn_corner(dcn_pos(y, i, dy), dcn_pos(x, j, dx), D.h, D.w);
    if (D.dbg_y0 != nullptr) { D.dbg_y0[pix * 9 + t] = c.y0; D.dbg_x0[pix * 9 + t] = c.x0; }
    float4 v = dcn_sample4(img, D.x_cstride, D.w, c);
    v.x *= m; v.y *= m; v.z *= m; v.w *= m;
    const float4 w0 = s_w[t * 4 + 0], w1 = s_w[t * 4 + 1], w2 = s_w[t * 4 + 2], w3 = s_w[t * 4 + 3];
    const float2 vx = make_float2(v.x, v.x), vy = make_float2(v.y, v.y), vz = make_float2(v.z, v.z), vw = make_float2(v.w, v.w);
    a01 = __ffma2_rn(vx, make_float2(w0.x, w0.y), __ffma2_rn(vy, make_float2(w1.x, w1.y),
          __ffma2_rn(vz, make_float2(w2.x, w2.y), __ffma2_rn(vw, make_float2(w3.x, w3.y), a01))));
    a23 = __ffma2_rn(vx, make_float2(w0.z, w0.w), __ffma2_rn(vy, make_float2(w1.z, w1.w),
          __ffma2_rn(vz, make_float2(w2.z, w2.w), __ffma2_rn(vw, make_float2(w3.z, w3.w), a23))));
  }
  *reinterpret_cast<float4*>(D.out + pix * D.out_cstride + D.out_coffset) = make_float4(a01.x, a01.y, a23.x, a23.y);
}

__global__ void __launch_bounds__(256) dcn_indices_kernel(const crfp_dcn_desc D, int32_t* __restrict__ y0o,
                                                          int32_t* __restrict__ x0o) {
  const int gts = D.dg * 9;
  const long long total = (long long)D.n * D.h * D.w * gts;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gt = (int)(idx % gts);
  const long long pix = idx / gts;
  const int x = (int)(pix % D.w), y = (int)((pix / D.w) % D.h);
  const int t = gt % 9, i = t / 3, j = t - i * 3;
  const float* offp = D.offset + pix * D.off_cstride + D.off_coffset;
  float dy, dx;
  if (D.shared_taps) { const int g = gt / 9; dy = offp[g]; dx = offp[D.dg + g]; }
  else { dy = offp[gt * 2]; dx = offp[gt * 2 + 1]; }
  y0o[idx] = (int)floorf(dcn_pos(y, i, dy));
  x0o[idx] = (int)floorf(dcn_pos(x, j, dx));
}

int launch_dcn(const crfp_dcn_desc& d, cudaStream_t st) {
  if ((long long)d.n * d.h * d.w == 0) return CRFP_OK;
  if (d.head_raw) return CRFP_ERR_UNSUPPORTED;                    // raw heads: tensor-core align kernel only
  if ((d.dbg_y0 != nullptr) != (d.dbg_x0 != nullptr)) return CRFP_ERR_NULL;
  if (d.c == 32 && d.dg == 8 && d.cout == 32 && !d.shared_taps) {
    if (((d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 3) || ((d.off_cstride | d.off_coffset) & 1))
      return CRFP_ERR_BAD_SHAPE;
    const size_t smem = (size_t)(DK * DCO + DPIX * DKP) * sizeof(float);  // 36864 + 74752 = 111616 B
    static_assert(8 * DPIX * DCO <= DPIX * DKP, "reduction scratch must fit in the column tile");
    cudaError_t e = cudaFuncSetAttribute(dcn_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
    dim3 grid(ceil_div(d.w, DT) * ceil_div(d.h, DT), d.n);
    launch_k(dcn_l1_kernel, dim3(grid), dim3(256), (size_t)(smem), st, d);
    return check_launch();
  }
  if (d.c == 4 && d.dg == 1 && d.cout == 4) {
    if ((d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 3) return CRFP_ERR_BAD_SHAPE;
    dim3 block(32, 8), grid(ceil_div(d.w, 32), ceil_div(d.h, 8), d.n);
    launch_k(dcn_hr_kernel, dim3(grid), dim3(block), (size_t)(0), st, d);
    return check_launch();
  }
  return CRFP_ERR_UNSUPPORTED;
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_dcn_v2_fwd(const crfp_dcn_desc* d, crfp_stream stream) {
  if (!d || !d->x || !d->offset || !d->mask || !d->weight || !d->bias || !d->out) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0) return CRFP_ERR_BAD_SHAPE;
  return launch_dcn(*d, (cudaStream_t)stream);
}

extern "C" int crfp_dcn_v2_indices(const crfp_dcn_desc* d, int32_t* y0, int32_t* x0, crfp_stream stream) {
  if (!d || !d->offset || !y0 || !x0) return CRFP_ERR_NULL;
  const long long total = (long long)d->n * d->h * d->w * d->dg * 9;
  if (total <= 0) return CRFP_ERR_BAD_SHAPE;
  dcn_indices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*d, y0, x0);
  return check_launch();
}

extern "C" size_t crfp_sizeof_dcn_desc(void) { return sizeof(crfp_dcn_desc); }
