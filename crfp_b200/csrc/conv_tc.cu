// Tensor-core implicit-GEMM 3x3 convolution for the L1 (2h x 2w) layers of CRFP: bf16 NHWC activations,
// bf16 weights, fp32 accumulation in TMEM via tcgen05.mma (sm_100a), fused bias / LeakyReLU / ReLU / residual /
// DCN-head / channel-split / pixel-shuffle epilogues.
//
// Replaces the same reference code as conv_wide.cu (/root/reference/model/CRFP.py:154-193, 303-317, 433-552)
// for the dense contractions (cout 32..224, cin 24..72): resblocks, dcn_block, conv_fuse, offset/mask heads,
// upsample_post, dcn_3.upsample.
//
// Mapping.  GEMM M = 128 consecutive output pixels of one image row, N = cout tile (<= 128), K = 9 taps x cin.
// A CTA owns a 128-pixel-wide column strip and walks down `rows_per_cta` rows.  Input rows are staged ONCE
// (cp.async, zero-filled outside the image = the conv's zero padding) into a 4-slot ring, each slot laid out
// as [cin/8][130 pixels] 16-byte records — the canonical K-major no-swizzle UMMA layout with SBO = 128 B and
// LBO = 130*16 B — so the nine shifted windows of the 3x3 stencil are nine descriptor start addresses into the
// same ring (slot of row y+ky-1, +kx*16 B): no im2col copy, no re-read of the input.  One elected thread issues
// 9*cin/16 tcgen05.mma (M128 x N x K16) per output row and commits to an mbarrier; the 4 warps then pull their
// 32 TMEM lanes (tcgen05.ld 32x32b) and run the epilogue straight to global memory.  Several CTAs per SM overlap
// each other's load / MMA / epilogue phases.
#include "common.cuh"
#include "umma.cuh"

namespace crfp {

constexpr int TCM = 128;   // pixels per tile row (UMMA M)
constexpr int TCWP = 130;  // staged pixels per row (tile + 1-px halo each side)

__device__ __forceinline__ void tc_load_row(const TcParams& P, uint4* slot, int n, int y, int x0, int tid) {
  const bool yin = (y >= 0 && y < P.h);
  const int kcr = P.kc_real;
  const int items = kcr * TCWP;
  for (int it = tid; it < items; it += 128) {
    const int px = it / kcr, kc = it - px * kcr;
    const int x = x0 + px - 1;
    int s = 0;
    if (P.nsrc > 1 && kc >= P.kstart[1]) s = 1;
    if (P.nsrc > 2 && kc >= P.kstart[2]) s = 2;
    const bool in = yin && x >= 0 && x < P.w;
    const __nv_bfloat16* g = P.src[s];
    if (in) g += (((size_t)n * P.h + y) * (size_t)P.w + x) * P.src_cstride[s] + P.src_coffset[s] + (kc - P.kstart[s]) * 8;
    umma::cp_async16(slot + kc * TCWP + px, g, in ? 16u : 0u);
  }
}

__global__ void __launch_bounds__(128) conv_tc_kernel(const TcParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int KC = P.kc_total, NT = P.nt;
  uint4* sW = reinterpret_cast<uint4*>(smem);             // [9][KC][NT] records
  uint4* sA = sW + 9 * KC * NT;                           // [4 slots][KC][130] records
  float* sBias = reinterpret_cast<float*>(sA + 4 * KC * TCWP);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int cotile = blockIdx.z % P.ntiles, n = blockIdx.z / P.ntiles;
  const int x0 = blockIdx.x * TCM;
  const int y_begin = blockIdx.y * P.rows_per_cta;
  const int y_end = min(P.h, y_begin + P.rows_per_cta);
  const int slot_recs = KC * TCWP;

  pdl_trigger();
  {  // weights slab of this cout tile + bias
    const uint4* gw = reinterpret_cast<const uint4*>(P.weight) + (size_t)cotile * 9 * KC * NT;
    for (int i = tid; i < 9 * KC * NT; i += 128) umma::cp_async16(sW + i, gw + i, 16u);
    for (int i = tid; i < ((NT + 31) & ~31); i += 128) sBias[i] = (i < NT) ? P.bias[cotile * NT + i] : 0.f;
    // K padding chunks (never loaded) must be finite: zero them once in every slot
    for (int kc = P.kc_real; kc < KC; ++kc)
      for (int i = tid; i < 4 * TCWP; i += 128)
        sA[(i / TCWP) * slot_recs + kc * TCWP + (i % TCWP)] = make_uint4(0u, 0u, 0u, 0u);
  }
  uint32_t ncols = 32;
  while ((int)ncols < NT) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  pdl_wait();
  for (int r = -1; r <= 1; ++r) {
    tc_load_row(P, sA + ((y_begin + r) & 3) * slot_recs, n, y_begin + r, x0, tid);
    umma::cp_async_commit();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_bf16(TCM, NT);
  const uint32_t sA_addr = umma::smem_u32(sA), sW_addr = umma::smem_u32(sW);
  const uint32_t lboA = TCWP * 16, lboB = (uint32_t)NT * 16;
  uint32_t phase = 0;

  const int x = x0 + tid;
  const bool xvalid = x < P.w;

  for (int y = y_begin; y < y_end; ++y) {
    if (y + 1 < y_end) tc_load_row(P, sA + ((y + 2) & 3) * slot_recs, n, y + 2, x0, tid);
    umma::cp_async_commit();
    umma::cp_async_wait<1>();
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t a_base = sA_addr + (uint32_t)(((y + ky - 1) & 3) * slot_recs + kx) * 16;
        const uint32_t b_base = sW_addr + (uint32_t)(tap * KC * NT) * 16;
        for (int ks = 0; ks < KC / 2; ++ks) {
          const uint64_t da = umma::make_desc(a_base + (uint32_t)(2 * ks) * lboA, lboA, 128);
          const uint64_t db = umma::make_desc(b_base + (uint32_t)(2 * ks) * lboB, lboB, 128);
          umma::mma_bf16(taddr, da, db, idesc, (tap | ks) != 0 ? 1u : 0u);
        }
      }
      umma::mma_commit(&bar);
    }
    umma::mbar_wait(&bar, phase);
    phase ^= 1;
    umma::fence_after_sync();

    // ---------------- epilogue: thread = pixel (TMEM lane), 32 channels per chunk
    const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
    float2 fl = make_float2(0.f, 0.f);
    if (P.act == CRFP_ACT_DCN_HEAD && xvalid) fl = __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2));
    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
      umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
      const int cbase = cotile * NT + c0;  // first conv output channel of this chunk
      if (!xvalid || cbase >= P.cout) continue;
      // channels of this chunk that are real: inside this cout tile and inside cout
      const int nvalid = min(min(32, NT - c0), P.cout - cbase);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += sBias[c0 + i];
      if (P.act == CRFP_ACT_LRELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lrelu01(v[i]);
      } else if (P.act == CRFP_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      } else if (P.act == CRFP_ACT_DCN_HEAD) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int cc = cbase + i;
          v[i] = (cc < P.head_split) ? P.head_mag * tanhf(v[i]) + ((cc & 1) ? fl.x : fl.y) : sigmoidf_(v[i]);
        }
      }
      if (P.residual != nullptr) {
        const __nv_bfloat16* rp = P.residual + pix * P.res_cstride + P.res_coffset + cbase;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (8 * j >= nvalid) break;
          const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + 8 * j));
          const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(r2[k]);
            v[8 * j + 2 * k] += f.x;
            v[8 * j + 2 * k + 1] += f.y;
          }
        }
      }
      if (P.post_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= P.post_scale;
      }
      if (P.out_kind == TC_OUT_BF16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = cbase + 8 * j;
          if (8 * j >= nvalid) break;
          int seg = 0, cl = cc;
          if (P.ndst > 1 && cc >= P.dst_c[0]) { seg = 1; cl = cc - P.dst_c[0]; }
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(P.dst[seg]) + pix * P.dst_cstride[seg] + P.dst_coffset[seg] + cl;
          uint4 o;
          o.x = umma::pack_bf16(v[8 * j + 0], v[8 * j + 1]);
          o.y = umma::pack_bf16(v[8 * j + 2], v[8 * j + 3]);
          o.z = umma::pack_bf16(v[8 * j + 4], v[8 * j + 5]);
          o.w = umma::pack_bf16(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(op) = o;
        }
      } else if (P.out_kind == TC_OUT_F32) {
        float* op = reinterpret_cast<float*>(P.dst[0]) + pix * P.dst_cstride[0] + P.dst_coffset[0] + cbase;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (4 * j >= nvalid) break;
          *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else {  // TC_OUT_SHUFFLE_F32: F.pixel_shuffle(r): conv channel o*r*r + dy*r + dx -> (y*r+dy, x*r+dx, o), fp32
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
    umma::fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, ncols);
}

void tc_cout_tile(int cout, int* nt, int* ntiles) {
  const int tiles = (cout + 127) / 128;
  const int per = (cout + tiles - 1) / tiles;
  *ntiles = tiles;
  *nt = (per + 15) & ~15;
  if (*nt < 16) *nt = 16;
}

int launch_conv_tc(TcParams p, cudaStream_t st) {
  if (p.nsrc < 1 || p.nsrc > 3) return CRFP_ERR_BAD_SHAPE;
  int kc = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    if (!p.src[s]) return CRFP_ERR_NULL;
    if (p.src_c[s] % 8 || p.src_cstride[s] % 8 || p.src_coffset[s] % 8 || ((uintptr_t)p.src[s] & 15)) return CRFP_ERR_BAD_SHAPE;
    p.kstart[s] = kc;
    kc += p.src_c[s] / 8;
  }
  p.kc_real = kc;
  p.kc_total = (kc + 1) & ~1;
  tc_cout_tile(p.cout, &p.nt, &p.ntiles);
  if (!p.weight || !p.bias || !p.dst[0]) return CRFP_ERR_NULL;
  if (p.out_kind == TC_OUT_BF16) {
    for (int s = 0; s < p.ndst; ++s)
      if (p.dst_cstride[s] % 8 || p.dst_coffset[s] % 8 || (s == 0 && p.ndst > 1 && p.dst_c[0] % 8)) return CRFP_ERR_BAD_SHAPE;
  }
  if (p.residual && (p.res_cstride % 8 || p.res_coffset % 8)) return CRFP_ERR_BAD_SHAPE;
  if (p.post_scale == 0.f) p.post_scale = 1.f;
  const int strips = ceil_div(p.w, TCM);
  const int per_seg = strips * p.n * p.ntiles;
  int segs = 296 / per_seg;
  if (segs < 1) segs = 1;
  if (segs > ceil_div(p.h, 4)) segs = ceil_div(p.h, 4);
  p.rows_per_cta = ceil_div(p.h, segs);
  segs = ceil_div(p.h, p.rows_per_cta);
  const size_t smem = (size_t)(9 * p.kc_total * p.nt + 4 * p.kc_total * TCWP) * 16 + (size_t)((p.nt + 31) & ~31) * 4;
  if (smem > 227 * 1024) return CRFP_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  dim3 grid(strips, segs, p.n * p.ntiles);
  launch_k(conv_tc_kernel, dim3(grid), dim3(128), (size_t)(smem), st, p);
  return check_launch();
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_tc_cout_tile(int cout, int32_t* nt, int32_t* ntiles) {
  if (!nt || !ntiles || cout <= 0) return CRFP_ERR_BAD_SHAPE;
  int a, b;
  tc_cout_tile(cout, &a, &b);
  *nt = a; *ntiles = b;
  return CRFP_OK;
}

extern "C" int crfp_conv3x3_tc_fwd(const crfp_conv_tc_desc* d, crfp_stream stream) {
  if (!d) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0 || d->cout <= 0) return CRFP_ERR_BAD_SHAPE;
  if ((long long)d->n * d->h * d->w == 0) return CRFP_OK;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n = d->n; p.h = d->h; p.w = d->w; p.nsrc = d->nsrc;
  for (int s = 0; s < d->nsrc && s < 3; ++s) {
    p.src[s] = reinterpret_cast<const __nv_bfloat16*>(d->src[s].ptr);
    p.src_c[s] = d->src[s].c; p.src_cstride[s] = d->src[s].cstride; p.src_coffset[s] = d->src[s].coffset;
  }
  p.cout = d->cout; p.act = d->act;
  p.weight = reinterpret_cast<const __nv_bfloat16*>(d->weight); p.bias = d->bias;
  p.out_kind = d->out_kind; p.shuffle_r = d->shuffle_r; p.ndst = d->ndst;
  if (d->ndst < 1 || d->ndst > 2) return CRFP_ERR_BAD_SHAPE;
  for (int s = 0; s < d->ndst; ++s) {
    p.dst[s] = d->dst[s].ptr; p.dst_c[s] = d->dst[s].c; p.dst_cstride[s] = d->dst[s].cstride;
    p.dst_coffset[s] = d->dst[s].coffset;
  }
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual); p.res_cstride = d->res_cstride;
  p.res_coffset = d->res_coffset;
  p.flow = d->flow; p.head_split = d->head_split; p.head_mag = d->head_mag; p.post_scale = d->post_scale;
  if (d->act == CRFP_ACT_DCN_HEAD && !d->flow) return CRFP_ERR_NULL;
  if (d->out_kind == CRFP_TC_OUT_SHUFFLE_F32 && (d->shuffle_r < 1 || d->cout % (d->shuffle_r * d->shuffle_r))) return CRFP_ERR_BAD_SHAPE;
  return launch_conv_tc(p, (cudaStream_t)stream);
}

extern "C" size_t crfp_sizeof_conv_tc_desc(void) { return sizeof(crfp_conv_tc_desc); }
