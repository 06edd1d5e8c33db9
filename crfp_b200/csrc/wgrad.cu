// Weight gradients of the 3x3 convs and of the DCNv2 contraction, round-2 kernels (the gradients
// cuDNN's wgrad and `_ext.dcn_v2_backward` produce inside `loss.backward()`, /root/reference/trainer.py:246-250).
//
// Round 1's kernels (bwd.cu: one thread per (tap, 4 ci, 4 co) block streaming all pixels from L2, thousands of pixel chunks
// meeting in atomicAdd) ran at ~3 % of the FFMA peak: 57 % of a training step.  Two shared-memory / register-tiled
// replacements, both exact fp32:
//
//   conv_wgrad_tile_kernel   layers with cin % 4 == 0 and cout % 4 == 0 (L1 / LR / FNet).  CTA = 192 threads = (ky, 4-ci
//                            block, 4-co block) over a 32 ci x 32 co block of dW; pixel chunks (one image row segment of 64
//                            pixels + halo) are staged with cp.async, double buffered; a thread slides a 3-pixel window of
//                            its x row through registers: per pixel 2 LDS.128 feed 24 FFMA2 (all three kx taps).  Each CTA
//                            ends in ONE coalesced write of its partial sums; wgrad_reduce2_kernel adds the partial rows
//                            (<= ~300 per layer) — no same-address atomics.  FLAT mode = the DCNv2 weight gradient
//                            (plain col^T dout product, K = 288 handled as nine 32-channel blocks).
//   conv_wgrad_thin_kernel   the 4-channel HR layers (cin slices of <= 4 channels, cout <= 4): one thread per pixel column,
//                            48 accumulators (one ky, three kx, 4x4 channels), coalesced loads along x, warp-shuffle +
//                            shared-memory reduction, 52 atomics per CTA.  FLAT mode = the HR DCNv2 weight gradient (K = 36).
//
// The host emulation used by the CPU suite (tests/tools/hostemu) cannot run shared-memory kernels: it keeps building
// bwd.cu's sync-free kernels; these kernels are checked against them (and against torch autograd) on the GPU.
#include <stdlib.h>

#include "common.cuh"

namespace crfp {

__device__ __forceinline__ void wg_cp16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}

// acc[kx][ci][co pair] += x_kx[ci] * g[co]
__device__ __forceinline__ void wg_fma48(float2 (&acc)[3][4][2], const float4& xm, const float4& xc, const float4& xp,
                                         const float4& gv) {
  const float2 g01 = make_float2(gv.x, gv.y), g23 = make_float2(gv.z, gv.w);
  const float xs[3][4] = {{xm.x, xm.y, xm.z, xm.w}, {xc.x, xc.y, xc.z, xc.w}, {xp.x, xp.y, xp.z, xp.w}};
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float2 xv = make_float2(xs[kx][a], xs[kx][a]);
      acc[kx][a][0] = __ffma2_rn(xv, g01, acc[kx][a][0]);
      acc[kx][a][1] = __ffma2_rn(xv, g23, acc[kx][a][1]);
    }
}

// ------------------------------------------------------------------------------------------------ tiled (wide layers)
constexpr int WG_MAX_ENT = 16;

struct WgTile {
  int rows, h, w;          // rows = n*h image rows
  int cin, cout;           // channels (= pixel strides) of x and g
  int ncob;                // 32-wide co blocks; blockIdx.x = cib * ncob + cob
  int nchunks, segs;       // pixel chunks = nent * rows * segs, segs = ceil(w / SEG)
  int nent;                // (x, g) pairs of identical shape handled by this launch (frames of the recurrence), <= WG_MAX_ENT
  const float* x[WG_MAX_ENT];
  const float* g[WG_MAX_ENT];
  float* partial;          // [gridDim.y][pitch]
  long long pitch;         // taps*cin*cout + cout
};

template <bool FLAT>
__global__ void __launch_bounds__(192, FLAT ? 2 : 3) conv_wgrad_tile_kernel(const WgTile A) {
  constexpr int SEG = FLAT ? 32 : 64;
  constexpr int XPX = FLAT ? SEG : SEG + 2;   // pixels per x tile row
  constexpr int XCH = FLAT ? 96 : 32;         // channels per pixel in the x tile
  constexpr int XS = 3 * XPX * XCH, GS = SEG * 32, STAGE = XS + GS;
  extern __shared__ __align__(16) float wg_smem[];
  const int tid = threadIdx.x;
  const int j = tid >> 6, ciq = (tid >> 3) & 7, coq = tid & 7;
  const int cib = blockIdx.x / A.ncob, cob = blockIdx.x - cib * A.ncob;
  const int ci0 = FLAT ? 0 : cib * 32, co0 = cob * 32;

  auto load = [&](int chunk, int stage) {
    float* sx = wg_smem + stage * STAGE;
    float* sg = sx + XS;
    const int per_ent = A.rows * A.segs;
    const int ent = chunk / per_ent, lc = chunk - ent * per_ent;
    const int r = lc / A.segs, x0 = (lc - r * A.segs) * SEG;
    const float* __restrict__ xe = A.x[ent];
    const float* __restrict__ ge = A.g[ent];
    if (FLAT) {
      for (int i = tid; i < 3 * SEG * 24; i += 192) {
        const int q = i % 24, px = (i / 24) % SEG, jj = i / (24 * SEG);
        const int col = x0 + px;
        const bool in = col < A.w;
        const float* src = in ? xe + ((long long)r * A.w + col) * A.cin + jj * 96 + q * 4 : xe;
        wg_cp16(sx + (jj * XPX + px) * XCH + q * 4, src, in ? 16u : 0u);
      }
    } else {
      const int y = r % A.h;
      for (int i = tid; i < 3 * XPX * 8; i += 192) {
        const int q = i & 7, px = (i >> 3) % XPX, jj = (i >> 3) / XPX;
        const int col = x0 - 1 + px, yi = y + jj - 1;
        const bool in = col >= 0 && col < A.w && yi >= 0 && yi < A.h && ci0 + q * 4 < A.cin;
        const float* src = in ? xe + ((long long)(r + jj - 1) * A.w + col) * A.cin + ci0 + q * 4 : xe;
        wg_cp16(sx + (jj * XPX + px) * XCH + q * 4, src, in ? 16u : 0u);
      }
    }
    for (int i = tid; i < SEG * 8; i += 192) {
      const int q = i & 7, px = i >> 3;
      const int col = x0 + px;
      const bool in = col < A.w && co0 + q * 4 < A.cout;
      const float* src = in ? ge + ((long long)r * A.w + col) * A.cout + co0 + q * 4 : ge;
      wg_cp16(sg + px * 32 + q * 4, src, in ? 16u : 0u);
    }
  };

  float2 acc[3][4][2];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[kx][a][0] = acc[kx][a][1] = make_float2(0.f, 0.f);
  float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool do_bias = j == 1 && ciq == 0 && cib == 0;

  int chunk = blockIdx.y, stage = 0;
  if (chunk < A.nchunks) load(chunk, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  while (chunk < A.nchunks) {
    const int next = chunk + gridDim.y;
    if (next < A.nchunks) load(next, stage ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* xt = wg_smem + stage * STAGE + j * XPX * XCH + ciq * 4;
    const float* gt = wg_smem + stage * STAGE + XS + coq * 4;
    if (FLAT) {
#pragma unroll 4
      for (int px = 0; px < SEG; ++px) {
        const float4 xm = *reinterpret_cast<const float4*>(xt + px * XCH);
        const float4 xc = *reinterpret_cast<const float4*>(xt + px * XCH + 32);
        const float4 xp = *reinterpret_cast<const float4*>(xt + px * XCH + 64);
        const float4 gv = *reinterpret_cast<const float4*>(gt + px * 32);
        wg_fma48(acc, xm, xc, xp, gv);
        if (do_bias) { gsum.x += gv.x; gsum.y += gv.y; gsum.z += gv.z; gsum.w += gv.w; }
      }
    } else {
      float4 xm = *reinterpret_cast<const float4*>(xt);
      float4 xc = *reinterpret_cast<const float4*>(xt + XCH);
#pragma unroll 4
      for (int px = 0; px < SEG; ++px) {
        const float4 xp = *reinterpret_cast<const float4*>(xt + (px + 2) * XCH);
        const float4 gv = *reinterpret_cast<const float4*>(gt + px * 32);
        wg_fma48(acc, xm, xc, xp, gv);
        if (do_bias) { gsum.x += gv.x; gsum.y += gv.y; gsum.z += gv.z; gsum.w += gv.w; }
        xm = xc;
        xc = xp;
      }
    }
    __syncthreads();   // everybody is done with this stage before the next prefetch overwrites it
    chunk = next;
    stage ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // one partial row per CTA row (blockIdx.y); the (ci, co) blocks of blockIdx.x fill disjoint parts of it
  float* prow = A.partial + (long long)blockIdx.y * A.pitch;
  const int co = co0 + coq * 4;
  if (co < A.cout) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int ci = FLAT ? (j * 3 + kx) * 32 + ciq * 4 + a : ci0 + ciq * 4 + a;
        if (ci >= A.cin) continue;
        const long long idx = FLAT ? (long long)ci * A.cout + co : ((long long)(j * 3 + kx) * A.cin + ci) * A.cout + co;
        *reinterpret_cast<float4*>(prow + idx) = make_float4(acc[kx][a][0].x, acc[kx][a][0].y, acc[kx][a][1].x, acc[kx][a][1].y);
      }
    if (do_bias) *reinterpret_cast<float4*>(prow + (A.pitch - A.cout) + co) = gsum;
  }
}

// dw[(tap*cin_total + cin_off + ci)*cout + co] += sum over the partial rows; db[co] += the bias partials (db may be NULL)
__global__ void __launch_bounds__(256) wgrad_reduce2_kernel(int prows, int taps, int cin, int cout, int cin_total, int cin_off,
                                                            const float* __restrict__ partial, float* __restrict__ dw,
                                                            float* __restrict__ db) {
  __shared__ float red[8][33];
  const int elems = taps * cin * cout;
  const int idx = blockIdx.x * 32 + threadIdx.x;
  const long long pitch = (long long)elems + cout;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (idx < elems + cout) {
    int p = threadIdx.y;
    for (; p + 24 < prows; p += 32) {
      s0 += partial[p * pitch + idx];
      s1 += partial[(p + 8) * pitch + idx];
      s2 += partial[(p + 16) * pitch + idx];
      s3 += partial[(p + 24) * pitch + idx];
    }
    for (; p < prows; p += 8) s0 += partial[p * pitch + idx];
  }
  red[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.y != 0 || idx >= elems + cout) return;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sum += red[k][threadIdx.x];
  if (idx < elems) {
    const int co = idx % cout, ci = (idx / cout) % cin, tap = idx / (cout * cin);
    atomicAdd(dw + ((long long)tap * cin_total + cin_off + ci) * cout + co, sum);
  } else if (db != nullptr) {
    atomicAdd(db + (idx - elems), sum);
  }
}

// ------------------------------------------------------------------------------------------------ thin (HR layers)
struct WgThin {
  int rows, h, w;
  int cin;                 // pixel stride of x (channels of the source tensor)
  int cq, nci;             // first channel of this launch's quad, valid channels in it (1..4)
  int cout;                // <= 4, pixel stride of g
  int cin_total, cin_off;
  int rb;                  // image rows per CTA
  int flat;                // 1: DCNv2 HR weight gradient — the (ky, kx) "taps" are the nine 4-channel blocks of col[.][36]
  int vecx, vecg;          // 16-byte loads allowed
  int nent;                // (x, g) pairs handled by this launch; blockIdx.y walks nent * rows image rows
  const float* x[WG_MAX_ENT];
  const float* g[WG_MAX_ENT];
  float* dw;
  float* db;
};

__device__ __forceinline__ float4 wg_ld4(const float* p, int nvalid, bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  v.x = __ldg(p);
  if (nvalid > 1) v.y = __ldg(p + 1);
  if (nvalid > 2) v.z = __ldg(p + 2);
  if (nvalid > 3) v.w = __ldg(p + 3);
  return v;
}

// grid (ceil(w/128), ceil(rows/rb), 3 = ky), block 128 (threads along x)
__global__ void __launch_bounds__(128) conv_wgrad_thin_kernel(const WgThin A) {
  __shared__ float red[4][52];
  const int px = blockIdx.x * 128 + threadIdx.x;
  const int ky = blockIdx.z;
  const long long all_rows = (long long)A.nent * A.rows;
  const long long r0 = (long long)blockIdx.y * A.rb;
  long long r1 = r0 + A.rb;
  if (r1 > all_rows) r1 = all_rows;
  float2 acc[3][4][2];
#pragma unroll
  for (int kx = 0; kx < 3; ++kx)
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[kx][a][0] = acc[kx][a][1] = make_float2(0.f, 0.f);
  float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool do_bias = ky == 1 && A.db != nullptr;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  if (px < A.w && r0 < r1) {
    // one row of operands is fetched while the previous one is consumed (the loop is latency bound otherwise: four dependent
    // loads, then 48 FMAs); loads outside the image are predicated to zero instead of skipping the row
    const bool vecx = A.vecx != 0, vecg = A.vecg != 0;
    struct RowOps { float4 g, xm, xc, xp; };
    auto load_row = [&](int ent, int r, int y) {
      RowOps o;
      const float* __restrict__ xe = A.x[ent];
      o.g = wg_ld4(A.g[ent] + ((long long)r * A.w + px) * A.cout, A.cout, vecg);
      if (A.flat) {
        const float* xb = xe + ((long long)r * A.w + px) * A.cin + ky * 12;
        o.xm = wg_ld4(xb, 4, vecx);
        o.xc = wg_ld4(xb + 4, 4, vecx);
        o.xp = wg_ld4(xb + 8, 4, vecx);
      } else {
        const int yi = y + ky - 1;
        const bool ok = yi >= 0 && yi < A.h;
        const float* xr = xe + ((long long)(r + ky - 1) * A.w + px) * A.cin + A.cq;
        o.xm = (ok && px > 0) ? wg_ld4(xr - A.cin, A.nci, vecx) : zero;
        o.xc = ok ? wg_ld4(xr, A.nci, vecx) : zero;
        o.xp = (ok && px + 1 < A.w) ? wg_ld4(xr + A.cin, A.nci, vecx) : zero;
      }
      return o;
    };
    int ent = (int)(r0 / A.rows);
    int r = (int)(r0 - (long long)ent * A.rows);
    int y = r % A.h;
    RowOps cur = load_row(ent, r, y);
    for (long long rr = r0; rr < r1; ++rr) {
      int en = ent, rn = r + 1, yn = y + 1;
      if (yn == A.h) yn = 0;
      if (rn == A.rows) { rn = 0; ++en; }
      RowOps nxt = cur;
      if (rr + 1 < r1) nxt = load_row(en, rn, yn);
      if (do_bias) { gsum.x += cur.g.x; gsum.y += cur.g.y; gsum.z += cur.g.z; gsum.w += cur.g.w; }
      wg_fma48(acc, cur.xm, cur.xc, cur.xp, cur.g);
      cur = nxt; ent = en; r = rn; y = yn;
    }
  }
  // CTA reduction: butterfly inside the warp, 4 warps through shared memory, 52 atomics per CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 52; ++i) {
    float v;
    if (i < 48) {
      const float2 pr = acc[i >> 4][(i >> 2) & 3][(i >> 1) & 1];
      v = (i & 1) ? pr.y : pr.x;
    } else {
      v = i == 48 ? gsum.x : (i == 49 ? gsum.y : (i == 50 ? gsum.z : gsum.w));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= 52) return;
  const float sum = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
  if (i < 48) {
    const int kx = i >> 4, a = (i >> 2) & 3, c = i & 3;
    if (a >= A.nci || c >= A.cout) return;
    const int tap = ky * 3 + kx;
    const long long row = A.flat ? (long long)tap * 4 + a : (long long)tap * A.cin_total + A.cin_off + A.cq + a;
    atomicAdd(A.dw + row * A.cout + c, sum);
  } else if (do_bias && i - 48 < A.cout) {
    atomicAdd(A.db + (i - 48), sum);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int sm_count() {
  static const int n = []() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
  }();
  return n;
}
static bool wg_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool wg_enabled() {
  static const bool on = getenv("CRFP_WGRAD_V1") == nullptr;   // A/B: round 1's kernels
  return on;
}

static bool wg_thin_ok(int cin, int cout, int taps) {
  if (taps == 9) return cout <= 4 && cin <= 8;
  return cout <= 4 && cin == 36;
}
static bool wg_tile_ok(int cin, int cout, int taps) {
  if (cout % 4 != 0 || cout <= 4) return false;
  if (taps == 9) return cin % 4 == 0;
  return cin == 288;
}

// CTA rows (= partial rows) of the tiled kernel
static int wg_tile_workers(long long rows, int w, int cin, int cout, int taps) {
  const int seg = taps == 9 ? 64 : 32;
  const long long nchunks = rows * ((w + seg - 1) / seg);
  const int gx = (taps == 9 ? (cin + 31) / 32 : 1) * ((cout + 31) / 32);
  static const int per_sm = getenv("CRFP_WGRAD_CTAS") ? atoi(getenv("CRFP_WGRAD_CTAS")) : 3;   // CTAs per SM of the tile kernel (measured: 3 -> 31.7 ms, 2 -> 32.3 ms per V7 step)
  long long workers = ((long long)(taps == 9 ? per_sm : 2) * sm_count() + gx - 1) / gx;
  if (workers > nchunks) workers = nchunks;
  if (workers < 1) workers = 1;
  return (int)workers;
}

size_t wgrad_workspace_floats(long long rows, int w, int cin, int cout, int taps) {
  if (!wg_enabled() || !wg_tile_ok(cin, cout, taps)) return 0;
  return (size_t)wg_tile_workers(rows, w, cin, cout, taps) * ((size_t)taps * cin * cout + cout);
}

// returns CRFP_OK when one of the round-2 kernels took the job, 1 when the caller should fall back to round 1's.
// nent (x, g) pairs of identical shape (rows = n*h image rows each) are summed into the same dw / db by ONE launch.
int launch_bwd_weight_v2(long long rows, int h, int w, int cin, int cout, int taps, int cin_total, int cin_off,
                         const float* const* xs, const float* const* gs, int nent, float* dw, float* db, float* workspace,
                         size_t ws_floats, cudaStream_t st) {
  if (!wg_enabled() || nent < 1 || nent > WG_MAX_ENT || rows * nent * w >= (1LL << 31)) return 1;
  bool al = true;
  for (int e = 0; e < nent; ++e) al = al && wg_aligned16(xs[e]) && wg_aligned16(gs[e]);
  if (wg_thin_ok(cin, cout, taps)) {
    WgThin A;
    A.rows = (int)rows; A.h = h; A.w = w; A.cin = cin; A.cout = cout; A.cin_total = cin_total; A.cin_off = cin_off;
    A.flat = taps == 1; A.dw = dw; A.nent = nent;
    for (int e = 0; e < nent; ++e) { A.x[e] = xs[e]; A.g[e] = gs[e]; }
    A.vecg = cout == 4 && al;
    const int gx = (w + 127) / 128;
    const long long all_rows = rows * nent;
    long long rb = (all_rows * gx * 3 + 4LL * sm_count() - 1) / (4LL * sm_count());
    if (rb < 8) rb = 8;
    if (rb > 64) rb = 64;
    A.rb = (int)rb;
    const dim3 grid(gx, (unsigned)((all_rows + rb - 1) / rb), 3);
    if (grid.y > 65535) return 1;
    if (taps == 1) {
      A.cq = 0; A.nci = 4; A.db = db; A.vecx = al;
      conv_wgrad_thin_kernel<<<grid, 128, 0, st>>>(A);
      return check_launch();
    }
    for (int cq = 0; cq < cin; cq += 4) {
      A.cq = cq; A.nci = cin - cq < 4 ? cin - cq : 4;
      A.db = cq == 0 ? db : nullptr;
      A.vecx = (cin % 4 == 0) && al;
      conv_wgrad_thin_kernel<<<grid, 128, 0, st>>>(A);
      CRFP_TRY(check_launch());
    }
    return CRFP_OK;
  }
  if (!wg_tile_ok(cin, cout, taps) || workspace == nullptr || !al || !wg_aligned16(workspace)) return 1;
  const int workers = wg_tile_workers(rows * nent, w, cin, cout, taps);
  const long long pitch = (long long)taps * cin * cout + cout;
  if ((size_t)workers * (size_t)pitch > ws_floats) return 1;
  WgTile A;
  A.rows = (int)rows; A.h = h; A.w = w; A.cin = cin; A.cout = cout;
  A.ncob = (cout + 31) / 32;
  const int seg = taps == 9 ? 64 : 32;
  A.segs = (w + seg - 1) / seg;
  if (rows * nent * A.segs >= (1LL << 31)) return 1;
  A.nchunks = (int)(rows * nent * A.segs);
  A.nent = nent;
  for (int e = 0; e < nent; ++e) { A.x[e] = xs[e]; A.g[e] = gs[e]; }
  A.partial = workspace; A.pitch = pitch;
  const int ncib = taps == 9 ? (cin + 31) / 32 : 1;
  const dim3 grid(ncib * A.ncob, workers);
  if (taps == 9) {
    constexpr size_t smem = 2 * (3 * 66 * 32 + 64 * 32) * sizeof(float);
    cudaFuncSetAttribute(conv_wgrad_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_wgrad_tile_kernel<false><<<grid, 192, smem, st>>>(A);
  } else {
    constexpr size_t smem = 2 * (3 * 32 * 96 + 32 * 32) * sizeof(float);
    cudaFuncSetAttribute(conv_wgrad_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_wgrad_tile_kernel<true><<<grid, 192, smem, st>>>(A);
  }
  CRFP_TRY(check_launch());
  const int total = (int)pitch;
  wgrad_reduce2_kernel<<<dim3((total + 31) / 32), dim3(32, 8), 0, st>>>(workers, taps, cin, cout, cin_total, cin_off, workspace, dw, db);
  return check_launch();
}

}  // namespace crfp
