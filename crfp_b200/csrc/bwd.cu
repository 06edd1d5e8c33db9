// Backward kernels of the training step (BASELINE.json configs[4]): the gradients the
// reference obtains from ATen / cuDNN autograd and from `_ext.dcn_v2_backward` of jinfagang/DCNv2_latest
// (/root/reference/trainer.py:246-250 `loss.backward()` through model/CRFP.py:90-130 flow_warp, :350 DCNv2, the 3x3
// convs, nn.Upsample and AvgPool2d), plus the Charbonnier loss (loss/loss.py:116-124) and the Adam update
// (trainer.py:149,250).
//
// Round-1 design: first correct versions, then the three fixes the first ncu launch list asked for (vector DCN backward
// with a transposed weight, per-source weight gradients, backward-data moved onto the tiled forward conv kernel — see
// DESIGN.md 11).  Every kernel here is a sync-free, shared-memory-free SIMT kernel — one thread per output element /
// register tile (gather form) or per contribution (scatter form with atomicAdd) over dense fp32 NHWC tensors, coalesced
// along the channel axis.  That keeps them testable without a GPU: the same kernel bodies and entry points compile
// with g++ against tests/tools/hostemu/cuda_shim.h (CRFP_HOST_EMU; test infrastructure only) and are checked against
// torch autograd in the CPU suite.  Tensor-core / shared-memory-tiled weight gradients are round-2 work.
//
// Conventions: all tensors dense NHWC (pixel stride == channel count); "accumulated" outputs are += (the caller
// zero-fills them once per step), everything else is overwritten.
#include <stdlib.h>
#ifdef CRFP_HOST_EMU
#include "cuda_shim.h"
#include "dcn_pos.cuh"
#else
#include "common.cuh"
#define CRFP_LAUNCH(kernel, grid, block, st, ...) kernel<<<(grid), (block), 0, (st)>>>(__VA_ARGS__)
#endif

namespace crfp {

static inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------ activations
// g = dy * act'(v) with act'(v) recovered from the saved forward OUTPUT (out > 0 <=> v > 0 for LeakyReLU / ReLU).
__global__ void __launch_bounds__(256) act_bwd_kernel(long long count, int act, const float* __restrict__ dy,
                                                      const float* __restrict__ out, float* __restrict__ g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float d = dy[i];
  const bool pos = out[i] > 0.f;
  g[i] = pos ? d : (act == CRFP_ACT_LRELU ? 0.1f * d : 0.f);
}

// ------------------------------------------------------------------------------------------------ 3x3 conv, data
// dx[b,y,x,ci] = sum_{ky,kx,co} g[b, y+1-ky, x+1-kx, co] * W[co,cin_off+ci,ky,kx];  wt_t = [tap][co][cin_total] (ci
// fastest, so that the threads of one pixel read consecutive weights; g[.., co] is a warp-wide broadcast).  One call
// produces the gradient of ONE source of the forward's channel concat (channels [cin_off, cin_off+cin) of the weight).
__global__ void __launch_bounds__(256) conv3x3_bwd_data_kernel(int n, int h, int w, int cin, int cout, int cin_total,
                                                               int cin_off, const float* __restrict__ g,
                                                               const float* __restrict__ wt_t, float* __restrict__ dx) {
  const long long total = (long long)n * h * w * cin;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ci = (int)(idx % cin);
  const long long pix = idx / cin;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const long long b = pix / ((long long)w * h);
  float acc = 0.f;
  for (int ky = 0; ky < 3; ++ky) {
    const int yo = y + 1 - ky;
    if (yo < 0 || yo >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xo = x + 1 - kx;
      if (xo < 0 || xo >= w) continue;
      const float* gp = g + ((b * h + yo) * w + xo) * cout;
      const float* wp = wt_t + (long long)(ky * 3 + kx) * cout * cin_total + cin_off + ci;
#pragma unroll 4
      for (int co = 0; co < cout; ++co) acc += gp[co] * wp[(long long)co * cin_total];
    }
  }
  dx[idx] = acc;
}

// Register-tiled variant for cin % 4 == 0 and cout % 4 == 0: one thread per (pixel, 4 input channels); per 4 output
// channels it issues 1 LDG.128 of g (warp broadcast) + 4 LDG.128 of weights (coalesced over ci) for 16 FMAs.
__global__ void __launch_bounds__(256) conv3x3_bwd_data_v4_kernel(int n, int h, int w, int cin, int cout, int cin_total,
                                                                  int cin_off, const float* __restrict__ g,
                                                                  const float* __restrict__ wt_t, float* __restrict__ dx) {
  const int cq = cin >> 2;
  const long long total = (long long)n * h * w * cq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ci = (int)(idx % cq) * 4;
  const long long pix = idx / cq;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const long long b = pix / ((long long)w * h);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int ky = 0; ky < 3; ++ky) {
    const int yo = y + 1 - ky;
    if (yo < 0 || yo >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int xo = x + 1 - kx;
      if (xo < 0 || xo >= w) continue;
      const float* gp = g + ((b * h + yo) * w + xo) * cout;
      const float* wp = wt_t + (long long)(ky * 3 + kx) * cout * cin_total + cin_off + ci;
#pragma unroll 2
      for (int co = 0; co < cout; co += 4) {
        const float4 gv = *reinterpret_cast<const float4*>(gp + co);
        const float4 w0 = *reinterpret_cast<const float4*>(wp + (long long)co * cin_total);
        const float4 w1 = *reinterpret_cast<const float4*>(wp + (long long)(co + 1) * cin_total);
        const float4 w2 = *reinterpret_cast<const float4*>(wp + (long long)(co + 2) * cin_total);
        const float4 w3 = *reinterpret_cast<const float4*>(wp + (long long)(co + 3) * cin_total);
        acc.x += gv.x * w0.x + gv.y * w1.x + gv.z * w2.x + gv.w * w3.x;
        acc.y += gv.x * w0.y + gv.y * w1.y + gv.z * w2.y + gv.w * w3.y;
        acc.z += gv.x * w0.z + gv.y * w1.z + gv.z * w2.z + gv.w * w3.z;
        acc.w += gv.x * w0.w + gv.y * w1.w + gv.z * w2.w + gv.w * w3.w;
      }
    }
  }
  *reinterpret_cast<float4*>(dx + pix * cin + ci) = acc;
}

// ------------------------------------------------------------------------------------------------ 3x3 conv, weights
// dw[tap][cin_off + ci][co] += sum_p x[p + shift(tap)][ci] * g[p][co];  db[co] += sum_p g[p][co]  (dw = [tap][cin_total]
// [cout]; one call handles ONE source of the forward's channel concat, so the concat is never materialised).
// One thread per weight element (co fastest: g loads coalesced, x loads warp-broadcast) and per block of image rows;
// partial sums meet in dw through atomicAdd.  taps == 1 is the plain (pixels x cin)^T (pixels x cout) product used for
// the DCN weight gradient (x = the modulated column buffer).
//
// Thread mapping (host-chosen): plain — one thread per weight element, 128 per CTA, unit-stride pixel loop;
// pixel lanes (few elements, the 4-channel HR layers: `lanes` > 1) — the CTA holds `lanes` copies of the `epad`-padded
// element set and copy `pl` takes every lanes-th pixel, so that the CTA's threads are not mostly idle.
// PARTIAL (pixel-lane mapping only): instead of the atomics, every (chunk, lane) writes its partial sums to row
// blockIdx.y * lanes + pl of `partial` ([rows][taps*cin*cout + cout], the bias partials behind the weights) and
// wgrad_reduce_kernel adds the rows up: no same-cache-line atomic contention, so the chunk count is free to grow.
template <bool UNIT, bool PARTIAL>   // UNIT: plain mapping (lanes == 1), unit-stride loop with compile-time pointer increments
__global__ void __launch_bounds__(128) conv_bwd_weight_kernel(int rows, int h, int w, int cin, int cout, int taps,
                                                              int cin_total, int cin_off, int rows_per_block, int xsegs,
                                                              int lanes, int epad, const float* __restrict__ x,
                                                              const float* __restrict__ g, float* __restrict__ dw,
                                                              float* __restrict__ db, float* __restrict__ partial) {
  const int e = UNIT ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : (int)(threadIdx.x % epad);
  const int pl = UNIT ? 0 : (int)(threadIdx.x / epad);
  const int step = UNIT ? 1 : lanes;
  if (e >= taps * cin * cout || pl >= step) return;
  const int co = e % cout;
  const int ci = (e / cout) % cin;
  const int tap = e / (cout * cin);
  const int ky = (taps == 9) ? tap / 3 : 1, kx = (taps == 9) ? tap % 3 : 1;
  const bool do_bias = (PARTIAL || db != nullptr) && ci == 0 && tap == ((taps == 9) ? 4 : 0);
  // blockIdx.y = (row block, column segment): rows [r0, r1), columns [c0, c1)
  const long long r0 = (long long)(blockIdx.y / xsegs) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  const int wseg = (w + xsegs - 1) / xsegs;
  const int c0 = (int)(blockIdx.y % xsegs) * wseg;
  const int c1 = (c0 + wseg < w) ? c0 + wseg : w;
  const int xlo = (kx == 0 && c0 == 0) ? 1 : c0, xhi = (kx == 2 && c1 == w) ? w - 1 : c1;
  float acc = 0.f, gsum = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const int y = (int)(r % h);
    const int yi = y + ky - 1;
    const float* gp = g + r * w * cout + co;
    if (do_bias)
      for (int xx = c0 + pl; xx < c1; xx += step) gsum += gp[(long long)xx * cout];
    if (yi < 0 || yi >= h) continue;
    const float* xp = x + (r + (ky - 1)) * w * cin + ci;
#pragma unroll 4
    for (int xx = xlo + pl; xx < xhi; xx += step) acc += gp[(long long)xx * cout] * xp[(long long)(xx + kx - 1) * cin];
  }
  if (PARTIAL) {
    float* prow = partial + ((long long)blockIdx.y * lanes + pl) * ((long long)taps * cin * cout + cout);
    prow[e] = acc;
    if (ci == 0 && tap == ((taps == 9) ? 4 : 0)) prow[(long long)taps * cin * cout + co] = gsum;
    return;
  }
  atomicAdd(dw + ((long long)tap * cin_total + cin_off + ci) * cout + co, acc);
  if (do_bias) atomicAdd(db + co, gsum);
}

// Register-tiled variant for cin % 4 == 0 and cout % 4 == 0: one thread per (tap, 4 ci, 4 co) block of dw; per pixel
// 2 LDG.128 (x: warp broadcast over the co blocks, g: coalesced) feed 16 FMAs.
template <bool UNIT, bool PARTIAL>
__global__ void __launch_bounds__(128) conv_bwd_weight_v4_kernel(int rows, int h, int w, int cin, int cout, int taps,
                                                                 int cin_total, int cin_off, int rows_per_block, int xsegs,
                                                                 int lanes, int epad, const float* __restrict__ x,
                                                                 const float* __restrict__ g, float* __restrict__ dw,
                                                                 float* __restrict__ db, float* __restrict__ partial) {
  const int cq = cin >> 2, oq = cout >> 2;
  const int e = UNIT ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : (int)(threadIdx.x % epad);
  const int pl = UNIT ? 0 : (int)(threadIdx.x / epad);
  const int step = UNIT ? 1 : lanes;
  if (e >= taps * cq * oq || pl >= step) return;
  const int co = (e % oq) * 4;
  const int ci = ((e / oq) % cq) * 4;
  const int tap = e / (oq * cq);
  const int ky = (taps == 9) ? tap / 3 : 1, kx = (taps == 9) ? tap % 3 : 1;
  const bool do_bias = (PARTIAL || db != nullptr) && ci == 0 && tap == ((taps == 9) ? 4 : 0);
  const long long r0 = (long long)(blockIdx.y / xsegs) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  const int wseg = (w + xsegs - 1) / xsegs;
  const int c0 = (int)(blockIdx.y % xsegs) * wseg;
  const int c1 = (c0 + wseg < w) ? c0 + wseg : w;
  float acc[4][4];
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int xlo = (kx == 0 && c0 == 0) ? 1 : c0, xhi = (kx == 2 && c1 == w) ? w - 1 : c1;
  for (long long r = r0; r < r1; ++r) {
    const int y = (int)(r % h);
    const int yi = y + ky - 1;
    const float* gp = g + r * w * cout + co;
    if (do_bias)
      for (int xx = c0 + pl; xx < c1; xx += step) {
        const float4 gv = *reinterpret_cast<const float4*>(gp + (long long)xx * cout);
        gsum.x += gv.x; gsum.y += gv.y; gsum.z += gv.z; gsum.w += gv.w;
      }
    if (yi < 0 || yi >= h) continue;
    const float* xp = x + (r + (ky - 1)) * w * cin + ci;
#pragma unroll 4
    for (int xx = xlo + pl; xx < xhi; xx += step) {
      const float4 gv = *reinterpret_cast<const float4*>(gp + (long long)xx * cout);
      const float4 xv = *reinterpret_cast<const float4*>(xp + (long long)(xx + kx - 1) * cin);
      acc[0][0] += xv.x * gv.x; acc[0][1] += xv.x * gv.y; acc[0][2] += xv.x * gv.z; acc[0][3] += xv.x * gv.w;
      acc[1][0] += xv.y * gv.x; acc[1][1] += xv.y * gv.y; acc[1][2] += xv.y * gv.z; acc[1][3] += xv.y * gv.w;
      acc[2][0] += xv.z * gv.x; acc[2][1] += xv.z * gv.y; acc[2][2] += xv.z * gv.z; acc[2][3] += xv.z * gv.w;
      acc[3][0] += xv.w * gv.x; acc[3][1] += xv.w * gv.y; acc[3][2] += xv.w * gv.z; acc[3][3] += xv.w * gv.w;
    }
  }
  if (PARTIAL) {
    float* prow = partial + ((long long)blockIdx.y * lanes + pl) * ((long long)taps * cin * cout + cout);
    float* pd = prow + ((long long)tap * cin + ci) * cout + co;
    for (int a = 0; a < 4; ++a)
      for (int c = 0; c < 4; ++c) pd[(long long)a * cout + c] = acc[a][c];
    if (ci == 0 && tap == ((taps == 9) ? 4 : 0)) {
      float* pb = prow + (long long)taps * cin * cout + co;
      pb[0] = gsum.x; pb[1] = gsum.y; pb[2] = gsum.z; pb[3] = gsum.w;
    }
    return;
  }
  float* d = dw + ((long long)tap * cin_total + cin_off + ci) * cout + co;
  for (int a = 0; a < 4; ++a)
    for (int c = 0; c < 4; ++c) atomicAdd(d + (long long)a * cout + c, acc[a][c]);
  if (do_bias) {
    atomicAdd(db + co, gsum.x); atomicAdd(db + co + 1, gsum.y); atomicAdd(db + co + 2, gsum.z); atomicAdd(db + co + 3, gsum.w);
  }
}

// Opt-in variant (CRFP_WGRAD_KX3=1) of the 4x4-tile kernel: one thread owns the THREE kx taps of one (ky, 4 ci, 4 co)
// block and slides a 3-pixel window of x through registers, so a pixel step is 1 new LDG.128 of x + 1 of g for 48 FMAs
// (the plain kernel: 2 loads for 16).  ncu on the plain kernel showed a 29 % L1 hit rate and ~1.3 TB/s of L2 traffic per
// launch; this cuts the x traffic by 3.  CPU-emulation-verified only (written after round 1's GPU budget was spent).
__global__ void __launch_bounds__(128) conv_bwd_weight_v4x3_kernel(int rows, int h, int w, int cin, int cout, int cin_total,
                                                                   int cin_off, int rows_per_block, int xsegs,
                                                                   const float* __restrict__ x, const float* __restrict__ g,
                                                                   float* __restrict__ dw, float* __restrict__ db) {
  const int cq = cin >> 2, oq = cout >> 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * cq * oq) return;
  const int co = (e % oq) * 4;
  const int ci = ((e / oq) % cq) * 4;
  const int ky = e / (oq * cq);
  const bool do_bias = (db != nullptr) && ci == 0 && ky == 1;
  const long long r0 = (long long)(blockIdx.y / xsegs) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  const int wseg = (w + xsegs - 1) / xsegs;
  const int c0 = (int)(blockIdx.y % xsegs) * wseg;
  const int c1 = (c0 + wseg < w) ? c0 + wseg : w;
  float acc[3][4][4];
  for (int k = 0; k < 3; ++k)
    for (int a = 0; a < 4; ++a)
      for (int c = 0; c < 4; ++c) acc[k][a][c] = 0.f;
  float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = r0; r < r1; ++r) {
    const int y = (int)(r % h);
    const int yi = y + ky - 1;
    const float* gp = g + r * w * cout + co;
    if (yi < 0 || yi >= h) continue;        // (never the bias thread: ky == 1 is always inside)
    const float* xp = x + (r + (ky - 1)) * w * cin + ci;
    float4 xm1 = (c0 > 0) ? *reinterpret_cast<const float4*>(xp + (long long)(c0 - 1) * cin) : z4;
    float4 x0 = (c0 < c1) ? *reinterpret_cast<const float4*>(xp + (long long)c0 * cin) : z4;
#pragma unroll 2
    for (int xx = c0; xx < c1; ++xx) {
      const float4 gv = *reinterpret_cast<const float4*>(gp + (long long)xx * cout);
      const float4 xp1 = (xx + 1 < w) ? *reinterpret_cast<const float4*>(xp + (long long)(xx + 1) * cin) : z4;
      if (do_bias) { gsum.x += gv.x; gsum.y += gv.y; gsum.z += gv.z; gsum.w += gv.w; }
#define CRFP_OUTER(K_, XV)                                                                                          \
      acc[K_][0][0] += XV.x * gv.x; acc[K_][0][1] += XV.x * gv.y; acc[K_][0][2] += XV.x * gv.z; acc[K_][0][3] += XV.x * gv.w; \
      acc[K_][1][0] += XV.y * gv.x; acc[K_][1][1] += XV.y * gv.y; acc[K_][1][2] += XV.y * gv.z; acc[K_][1][3] += XV.y * gv.w; \
      acc[K_][2][0] += XV.z * gv.x; acc[K_][2][1] += XV.z * gv.y; acc[K_][2][2] += XV.z * gv.z; acc[K_][2][3] += XV.z * gv.w; \
      acc[K_][3][0] += XV.w * gv.x; acc[K_][3][1] += XV.w * gv.y; acc[K_][3][2] += XV.w * gv.z; acc[K_][3][3] += XV.w * gv.w;
      CRFP_OUTER(0, xm1)
      CRFP_OUTER(1, x0)
      CRFP_OUTER(2, xp1)
#undef CRFP_OUTER
      xm1 = x0;
      x0 = xp1;
    }
  }
  for (int kx = 0; kx < 3; ++kx) {
    float* d = dw + ((long long)(ky * 3 + kx) * cin_total + cin_off + ci) * cout + co;
    for (int a = 0; a < 4; ++a)
      for (int c = 0; c < 4; ++c) atomicAdd(d + (long long)a * cout + c, acc[kx][a][c]);
  }
  if (do_bias) {
    atomicAdd(db + co, gsum.x); atomicAdd(db + co + 1, gsum.y); atomicAdd(db + co + 2, gsum.z); atomicAdd(db + co + 3, gsum.w);
  }
}

// second stage of the PARTIAL mode: one thread per weight (and bias) element sums the `prows` partial rows
__global__ void __launch_bounds__(128) wgrad_reduce_kernel(int prows, int taps, int cin, int cout, int cin_total, int cin_off,
                                                           const float* __restrict__ partial, float* __restrict__ dw,
                                                           float* __restrict__ db) {
  const int elems = taps * cin * cout;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= elems + cout) return;
  float sum = 0.f;
  const long long pitch = (long long)elems + cout;
  for (int p = 0; p < prows; ++p) sum += partial[p * pitch + idx];
  if (idx < elems) {
    const int co = idx % cout, ci = (idx / cout) % cin, tap = idx / (cout * cin);
    atomicAdd(dw + ((long long)tap * cin_total + cin_off + ci) * cout + co, sum);
  } else if (db != nullptr) {
    atomicAdd(db + (idx - elems), sum);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int env_int(const char* name, int dflt, int lo, int hi) {
  const char* v = getenv(name);
  if (v == nullptr) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}

// two-stage (PARTIAL) geometry of a thin layer: pixel lanes per CTA and number of partial rows; 0 rows = not a thin layer
static void thin_plan(long long rows, int w, int cin, int cout, int taps, bool v4, int* lanes, int* epad, long long* chunks,
                      int* xsegs, long long* rpb) {
  const int elems = v4 ? taps * (cin / 4) * (cout / 4) : taps * cin * cout;
  *lanes = 1; *epad = 0; *chunks = 0; *xsegs = 1; *rpb = 1;
  if (elems > 64) return;
  *epad = 16;
  while (*epad < elems) *epad *= 2;
  *lanes = 128 / *epad;
  long long c = 8192 / *lanes;                      // <= 8192 partial rows (a few MB), ~1e5 threads in flight
  if (c <= rows) {
    *rpb = (rows + c - 1) / c;
  } else {
    *xsegs = (int)(c / rows);
    if (*xsegs > w / (32 * *lanes)) *xsegs = w / (32 * *lanes);
    if (*xsegs < 1) *xsegs = 1;
  }
  *chunks = ((rows + *rpb - 1) / *rpb) * *xsegs;
  if (*chunks > 65535) { *xsegs = 1; *chunks = (rows + *rpb - 1) / *rpb; }
}

static int launch_bwd_weight(long long rows, int h, int w, int cin, int cout, int taps, int cin_total, int cin_off,
                             const float* x, const float* g, float* dw, float* db, cudaStream_t st,
                             float* workspace = nullptr, size_t ws_floats = 0) {
#ifndef CRFP_HOST_EMU
  {  // round-2 kernels (wgrad.cu): shared-memory tiled / register-tiled, no same-address atomics per pixel chunk
    const int s2 = launch_bwd_weight_v2(rows, h, w, cin, cout, taps, cin_total, cin_off, &x, &g, 1, dw, db, workspace, ws_floats, st);
    if (s2 != 1) return s2;
  }
#endif
  if (workspace != nullptr) {                       // two-stage reduction for the thin layers (opt-in by the caller)
    const bool v4t = (cin % 4 == 0) && (cout % 4 == 0) && aligned16(x) && aligned16(g);
    int lanes_t, epad_t, xsegs_t;
    long long chunks_t, rpb_t;
    thin_plan(rows, w, cin, cout, taps, v4t, &lanes_t, &epad_t, &chunks_t, &xsegs_t, &rpb_t);
    const long long pitch = (long long)taps * cin * cout + cout;
    if (chunks_t > 0 && (size_t)(chunks_t * lanes_t * pitch) <= ws_floats) {
      const dim3 grid(1, (unsigned)chunks_t), block(128);
      float* nobias = nullptr;
      if (v4t)
        CRFP_LAUNCH((conv_bwd_weight_v4_kernel<false, true>), grid, block, st, (int)rows, h, w, cin, cout, taps, cin_total,
                    cin_off, (int)rpb_t, xsegs_t, lanes_t, epad_t, x, g, dw, nobias, workspace);
      else
        CRFP_LAUNCH((conv_bwd_weight_kernel<false, true>), grid, block, st, (int)rows, h, w, cin, cout, taps, cin_total,
                    cin_off, (int)rpb_t, xsegs_t, lanes_t, epad_t, x, g, dw, nobias, workspace);
      CRFP_TRY(check_launch());
      const int total = taps * cin * cout + cout;
      CRFP_LAUNCH(wgrad_reduce_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), st, (int)(chunks_t * lanes_t), taps,
                  cin, cout, cin_total, cin_off, (const float*)workspace, dw, db);
      return check_launch();
    }
  }
  const bool v4 = (cin % 4 == 0) && (cout % 4 == 0) && aligned16(x) && aligned16(g);   // (dw is only touched by atomics)
  const int elems = v4 ? taps * (cin / 4) * (cout / 4) : taps * cin * cout;
  // A/B knobs (read once): CTA size cap for the plain mapping and the thread target that sizes the pixel chunks
  // (measured on the B200, profiles/r01/v6_train_wgrad_ab.txt: neither knob moves the step time by more than 2 %)
  static const int bd_cap = env_int("CRFP_WGRAD_BD", 128, 32, 128);
  static const int thread_target = env_int("CRFP_WGRAD_THREADS", 524288, 1024, 1 << 24);
  // thread mapping, see the kernel comment
  int lanes = 1, epad = 0, bd;
  unsigned gx;
  if (elems <= 64) {
    epad = 16;
    while (epad < elems) epad *= 2;
    lanes = 128 / epad;
    bd = 128;
    gx = 1;
  } else {
    gx = (unsigned)((elems + bd_cap - 1) / bd_cap);
    bd = (int)((elems + gx - 1) / gx);
    bd = (bd + 31) / 32 * 32;
  }
  // How many pixel chunks (= CTAs per element block)?  Every chunk ends in one atomicAdd per weight element, and ncu
  // showed the launch time tracking the atomic count (~1e11 atomics/s device-wide, ~1 per 3 clocks on one cache line):
  // use just enough chunks to put ~160 k threads in flight, and for the 4-channel layers (a few cache lines of dw) at
  // most 512 partial sums per element.
  long long chunks = (long long)thread_target / ((long long)gx * bd);
  if ((long long)taps * cin * cout <= 1024 && chunks * lanes > 512) chunks = 512 / lanes;
  if (chunks < 1) chunks = 1;
  long long rpb = 1;
  int xsegs = 1;
  if (chunks <= rows) {
    rpb = (rows + chunks - 1) / chunks;
  } else {                       // more chunks than rows: split every row into column segments of >= 32 pixels per lane
    xsegs = (int)(chunks / rows);
    if (xsegs > w / (32 * lanes)) xsegs = w / (32 * lanes);
    if (xsegs < 1) xsegs = 1;
  }
  unsigned gy = (unsigned)((rows + rpb - 1) / rpb);
  if ((long long)gy * xsegs > 65535) xsegs = (int)(65535 / gy);
  gy *= (unsigned)xsegs;
  static const int kx3 = env_int("CRFP_WGRAD_KX3", 0, 0, 1);
  if (kx3 && v4 && taps == 9 && lanes == 1) {        // opt-in sliding-window variant: a third of the thread tiles
    const int elems3 = 3 * (cin / 4) * (cout / 4);
    const unsigned gx3 = (unsigned)((elems3 + 127) / 128);
    int bd3 = (int)((elems3 + gx3 - 1) / gx3);
    bd3 = (bd3 + 31) / 32 * 32;
    long long chunks3 = (long long)thread_target / ((long long)gx3 * bd3);
    if (chunks3 < 1) chunks3 = 1;
    long long rpb3 = 1;
    int xsegs3 = 1;
    if (chunks3 <= rows) {
      rpb3 = (rows + chunks3 - 1) / chunks3;
    } else {
      xsegs3 = (int)(chunks3 / rows);
      if (xsegs3 > w / 32) xsegs3 = w / 32;
      if (xsegs3 < 1) xsegs3 = 1;
    }
    unsigned gy3 = (unsigned)((rows + rpb3 - 1) / rpb3);
    if ((long long)gy3 * xsegs3 > 65535) xsegs3 = (int)(65535 / gy3);
    gy3 *= (unsigned)xsegs3;
    CRFP_LAUNCH(conv_bwd_weight_v4x3_kernel, dim3(gx3, gy3), dim3(bd3), st, (int)rows, h, w, cin, cout, cin_total, cin_off,
                (int)rpb3, xsegs3, x, g, dw, db);
    return check_launch();
  }
  float* nopartial = nullptr;
#define CRFP_WGRAD_ARGS (int)rows, h, w, cin, cout, taps, cin_total, cin_off, (int)rpb, xsegs, lanes, epad, x, g, dw, db, nopartial
  if (v4 && lanes == 1) CRFP_LAUNCH((conv_bwd_weight_v4_kernel<true, false>), dim3(gx, gy), dim3(bd), st, CRFP_WGRAD_ARGS);
  else if (v4) CRFP_LAUNCH((conv_bwd_weight_v4_kernel<false, false>), dim3(gx, gy), dim3(bd), st, CRFP_WGRAD_ARGS);
  else if (lanes == 1) CRFP_LAUNCH((conv_bwd_weight_kernel<true, false>), dim3(gx, gy), dim3(bd), st, CRFP_WGRAD_ARGS);
  else CRFP_LAUNCH((conv_bwd_weight_kernel<false, false>), dim3(gx, gy), dim3(bd), st, CRFP_WGRAD_ARGS);
#undef CRFP_WGRAD_ARGS
  return check_launch();
}

// ------------------------------------------------------------------------------------------------ fovea blend
// S' = lrelu(m * F + (1 - m) * S, 0.1) (model/CRFP.py:1672-1675; m = the (n,H,W,1) fovea mask) and its backward: one kernel each
// instead of five / six pointwise ATen kernels over HR planes.  One thread per (pixel, channel quad); C % 4 == 0.
__global__ void __launch_bounds__(256) fovea_blend_fwd_kernel(long long total4, int cq, const float* __restrict__ f,
                                                              const float* __restrict__ s, const float* __restrict__ m,
                                                              float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const float mk = m[idx / cq], mi = 1.f - mk;
  const float4 fv = *reinterpret_cast<const float4*>(f + idx * 4), sv = *reinterpret_cast<const float4*>(s + idx * 4);
  float4 o;
  o.x = mk * fv.x + mi * sv.x; o.y = mk * fv.y + mi * sv.y; o.z = mk * fv.z + mi * sv.z; o.w = mk * fv.w + mi * sv.w;
  o.x = o.x > 0.f ? o.x : 0.1f * o.x; o.y = o.y > 0.f ? o.y : 0.1f * o.y;
  o.z = o.z > 0.f ? o.z : 0.1f * o.z; o.w = o.w > 0.f ? o.w : 0.1f * o.w;
  *reinterpret_cast<float4*>(out + idx * 4) = o;
}

__global__ void __launch_bounds__(256) fovea_blend_bwd_kernel(long long total4, int cq, const float* __restrict__ dout,
                                                              const float* __restrict__ out, const float* __restrict__ m,
                                                              float* __restrict__ df, float* __restrict__ ds) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const float mk = m[idx / cq], mi = 1.f - mk;
  const float4 dv = *reinterpret_cast<const float4*>(dout + idx * 4), ov = *reinterpret_cast<const float4*>(out + idx * 4);
  float4 g;
  g.x = ov.x > 0.f ? dv.x : 0.1f * dv.x; g.y = ov.y > 0.f ? dv.y : 0.1f * dv.y;
  g.z = ov.z > 0.f ? dv.z : 0.1f * dv.z; g.w = ov.w > 0.f ? dv.w : 0.1f * dv.w;
  *reinterpret_cast<float4*>(df + idx * 4) = make_float4(g.x * mk, g.y * mk, g.z * mk, g.w * mk);
  *reinterpret_cast<float4*>(ds + idx * 4) = make_float4(g.x * mi, g.y * mi, g.z * mi, g.w * mi);
}

// ------------------------------------------------------------------------------------------------ DCN heads activation
// The pointwise tail of DCN_module.forward between the fused offset / mask conv and DCNv2 (model/CRFP.py:337-347):
//   offset[.., 2k+e] = mag * tanh(heads[.., 2k'+e]) + flow[.., 1-e]     (flow is (dx, dy); offsets are (dy, dx) pairs)
//   mask[.., k]      = sigmoid(heads[.., noff + k'])
// nk (offset pair, mask) outputs per pixel; repeat: the heads hold ONE pair and ONE mask per pixel (k' = 0, noff = 2) that
// the nine taps share (the HR module, dg = 1).  Replaces six ATen kernels (slice copies, tanh, mul, flip, repeat, add,
// sigmoid) per call and their ten backward kernels.  One thread per (pixel, k): sync-free like the rest of this file.
__global__ void __launch_bounds__(256) dcn_heads_act_fwd_kernel(long long total, int nk, int repeat, float mag,
                                                                const float* __restrict__ heads, const float* __restrict__ flow,
                                                                float* __restrict__ offset, float* __restrict__ mask) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long pix = idx / nk;
  const int k = (int)(idx - pix * nk);
  const int kin = repeat ? 0 : k, noff = repeat ? 2 : 2 * nk, ch = repeat ? 3 : 3 * nk;
  const float* hp = heads + pix * ch;
  const float fx = flow[pix * 2], fy = flow[pix * 2 + 1];
  offset[pix * 2 * nk + 2 * k] = mag * tanhf(hp[2 * kin]) + fy;
  offset[pix * 2 * nk + 2 * k + 1] = mag * tanhf(hp[2 * kin + 1]) + fx;
  mask[pix * nk + k] = 1.f / (1.f + expf(-hp[noff + kin]));
}

// dheads (overwritten) and dflow (overwritten): one thread per (pixel, INPUT head index k' ); the thread with k' == 0 also
// reduces the pixel's flow gradient
__global__ void __launch_bounds__(256) dcn_heads_act_bwd_kernel(long long total, int nk, int repeat, float mag,
                                                                const float* __restrict__ heads, const float* __restrict__ doffset,
                                                                const float* __restrict__ dmask, float* __restrict__ dheads,
                                                                float* __restrict__ dflow) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int nin = repeat ? 1 : nk;
  const long long pix = idx / nin;
  const int kin = (int)(idx - pix * nin);
  const int noff = 2 * nin, ch = 3 * nin;
  const float* hp = heads + pix * ch;
  const float* dop = doffset + pix * 2 * nk;
  const float* dmp = dmask + pix * nk;
  float gy = 0.f, gx = 0.f, gm = 0.f;
  if (repeat) {
    for (int k = 0; k < nk; ++k) { gy += dop[2 * k]; gx += dop[2 * k + 1]; gm += dmp[k]; }
  } else {
    gy = dop[2 * kin]; gx = dop[2 * kin + 1]; gm = dmp[kin];
  }
  const float ty = tanhf(hp[2 * kin]), tx = tanhf(hp[2 * kin + 1]);
  const float sg = 1.f / (1.f + expf(-hp[noff + kin]));
  dheads[pix * ch + 2 * kin] = gy * mag * (1.f - ty * ty);
  dheads[pix * ch + 2 * kin + 1] = gx * mag * (1.f - tx * tx);
  dheads[pix * ch + noff + kin] = gm * sg * (1.f - sg);
  if (kin == 0) {
    float fy = 0.f, fx = 0.f;
    for (int k = 0; k < nk; ++k) { fy += dop[2 * k]; fx += dop[2 * k + 1]; }
    dflow[pix * 2] = fx;
    dflow[pix * 2 + 1] = fy;
  }
}

// ------------------------------------------------------------------------------------------------ DCNv2
// One thread per (pixel, deformable group, tap).  For its C/dg channels it rebuilds the 4 bilinear corners, forms
// gcol[c] = sum_co W[k][co] * dout[pix][co] and emits
//   col[pix][k]            = m * val[c]                                  (consumed by the weight-gradient kernel)
//   dmask[pix][g*9+t]      = sum_c gcol[c] * val[c]
//   doffset[pix][(g*9+t)*2 + {0,1}] = m * sum_c gcol[c] * d val[c] / d{py,px}
//   dx[corner][g*cpg + c] += gcol[c] * m * (corner weight)               (atomicAdd scatter)
// The coordinate derivative follows dmcn_get_coordinate_weight of jinfagang/DCNv2_latest (0 when the sample lies
// outside (-1,H) x (-1,W), per-corner zero padding otherwise), which torchvision's deform_conv2d backward reproduces
// everywhere except exactly at py == -1 / px == -1.
__global__ void __launch_bounds__(128) dcn_bwd_sample_kernel(const crfp_dcn_bwd_desc D) {
  const int gts = D.dg * 9;
  const long long total = (long long)D.n * D.h * D.w * gts;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gt = (int)(idx % gts);
  const long long pix = idx / gts;
  const int x = (int)(pix % D.w);
  const int y = (int)((pix / D.w) % D.h);
  const long long b = pix / ((long long)D.w * D.h);
  const int t = gt % 9, g = gt / 9, i = t / 3, j = t - i * 3;
  const int cpg = D.c / D.dg;
  const int K = gts * cpg;
  const float oy = D.offset[pix * gts * 2 + gt * 2], ox = D.offset[pix * gts * 2 + gt * 2 + 1];
  const float m = D.mask[pix * gts + gt];
  const float py = dcn_pos(y, i, oy), px = dcn_pos(x, j, ox);   // dcn_pos.cuh: the forward's expression
  const float fy = floorf(py), fx = floorf(px);
  const int y0 = (int)fy, x0 = (int)fx;
  const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
  const bool inside = (py > -1.f) && (py < (float)D.h) && (px > -1.f) && (px < (float)D.w);
  const bool vy0 = inside && y0 >= 0, vy1 = inside && (y0 + 1 <= D.h - 1);
  const bool vx0 = x0 >= 0, vx1 = (x0 + 1 <= D.w - 1);
  const bool ok00 = vy0 && vx0, ok01 = vy0 && vx1, ok10 = vy1 && vx0, ok11 = vy1 && vx1;
  const long long img = b * D.h * D.w;
  const long long p00 = (img + (long long)y0 * D.w + x0) * D.c, p01 = p00 + D.c;
  const long long p10 = p00 + (long long)D.w * D.c, p11 = p10 + D.c;
  const float* dout = D.dout + pix * D.cout;
  float sdm = 0.f, sdy = 0.f, sdx = 0.f;
  for (int cc = 0; cc < cpg; ++cc) {
    const int ch = g * cpg + cc;
    const int k = gt * cpg + cc;
    const float* wk = D.weight + (long long)k * D.cout;
    float gc = 0.f;
    for (int co = 0; co < D.cout; ++co) gc += wk[co] * dout[co];
    const float v00 = ok00 ? D.x[p00 + ch] : 0.f, v01 = ok01 ? D.x[p01 + ch] : 0.f;
    const float v10 = ok10 ? D.x[p10 + ch] : 0.f, v11 = ok11 ? D.x[p11 + ch] : 0.f;
    const float val = hy * hx * v00 + hy * lx * v01 + ly * hx * v10 + ly * lx * v11;
    D.col[pix * K + k] = m * val;
    sdm += gc * val;
    sdy += gc * (hx * (v10 - v00) + lx * (v11 - v01));
    sdx += gc * (hy * (v01 - v00) + ly * (v11 - v10));
    const float gm = gc * m;
    if (ok00) atomicAdd(D.dx + p00 + ch, gm * hy * hx);
    if (ok01) atomicAdd(D.dx + p01 + ch, gm * hy * lx);
    if (ok10) atomicAdd(D.dx + p10 + ch, gm * ly * hx);
    if (ok11) atomicAdd(D.dx + p11 + ch, gm * ly * lx);
  }
  D.doffset[pix * gts * 2 + gt * 2] = m * sdy;
  D.doffset[pix * gts * 2 + gt * 2 + 1] = m * sdx;
  D.dmask[pix * gts + gt] = sdm;
}

// Vector variant for C/dg == 4 and cout % 4 == 0 (both CRFP configurations: C=32/dg=8 and C=4/dg=1): the group's 4
// channels travel as one float4 — corner loads, the column store and the dx scatter (one 16-byte vector atomic per
// corner instead of 4 scalar ones) — and gcol comes from the TRANSPOSED weight weight_t[co][K]: the lanes of a warp
// hold consecutive (group, tap) slots, so `weight_t + co*K + gt*4` is one coalesced LDG.128 per output channel (the
// [K][cout] layout made every lane walk its own 128-byte rows: 32 wavefronts per load, measured 0.95 ms per launch).
__global__ void __launch_bounds__(128) dcn_bwd_sample_v4_kernel(const crfp_dcn_bwd_desc D) {
  const int gts = D.dg * 9;
  const long long total = (long long)D.n * D.h * D.w * gts;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gt = (int)(idx % gts);
  const long long pix = idx / gts;
  const int x = (int)(pix % D.w);
  const int y = (int)((pix / D.w) % D.h);
  const long long b = pix / ((long long)D.w * D.h);
  const int t = gt % 9, g = gt / 9, i = t / 3, j = t - i * 3;
  const int K = gts * 4;
  const float oy = D.offset[pix * gts * 2 + gt * 2], ox = D.offset[pix * gts * 2 + gt * 2 + 1];
  const float m = D.mask[pix * gts + gt];
  const float py = dcn_pos(y, i, oy), px = dcn_pos(x, j, ox);   // dcn_pos.cuh: the forward's expression
  const float fy = floorf(py), fx = floorf(px);
  const int y0 = (int)fy, x0 = (int)fx;
  const float ly = py - fy, lx = px - fx, hy = 1.f - ly, hx = 1.f - lx;
  const bool inside = (py > -1.f) && (py < (float)D.h) && (px > -1.f) && (px < (float)D.w);
  const bool vy0 = inside && y0 >= 0, vy1 = inside && (y0 + 1 <= D.h - 1);
  const bool vx0 = x0 >= 0, vx1 = (x0 + 1 <= D.w - 1);
  const bool ok00 = vy0 && vx0, ok01 = vy0 && vx1, ok10 = vy1 && vx0, ok11 = vy1 && vx1;
  const long long img = b * D.h * D.w;
  const int ch = g * 4;
  const long long p00 = (img + (long long)y0 * D.w + x0) * D.c + ch, p01 = p00 + D.c;
  const long long p10 = p00 + (long long)D.w * D.c, p11 = p10 + D.c;
  const float* dout = D.dout + pix * D.cout;
  const float* wt = D.weight_t + gt * 4;
  float4 gc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (int co = 0; co < D.cout; co += 4) {
    const float4 dv = *reinterpret_cast<const float4*>(dout + co);
    const float4 w0 = *reinterpret_cast<const float4*>(wt + (long long)co * K);
    const float4 w1 = *reinterpret_cast<const float4*>(wt + (long long)(co + 1) * K);
    const float4 w2 = *reinterpret_cast<const float4*>(wt + (long long)(co + 2) * K);
    const float4 w3 = *reinterpret_cast<const float4*>(wt + (long long)(co + 3) * K);
    gc.x += w0.x * dv.x + w1.x * dv.y + w2.x * dv.z + w3.x * dv.w;
    gc.y += w0.y * dv.x + w1.y * dv.y + w2.y * dv.z + w3.y * dv.w;
    gc.z += w0.z * dv.x + w1.z * dv.y + w2.z * dv.z + w3.z * dv.w;
    gc.w += w0.w * dv.x + w1.w * dv.y + w2.w * dv.z + w3.w * dv.w;
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 v00 = ok00 ? *reinterpret_cast<const float4*>(D.x + p00) : z4;
  const float4 v01 = ok01 ? *reinterpret_cast<const float4*>(D.x + p01) : z4;
  const float4 v10 = ok10 ? *reinterpret_cast<const float4*>(D.x + p10) : z4;
  const float4 v11 = ok11 ? *reinterpret_cast<const float4*>(D.x + p11) : z4;
  const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
  float4 val;
  val.x = w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
  val.y = w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
  val.z = w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
  val.w = w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
  *reinterpret_cast<float4*>(D.col + pix * K + gt * 4) = make_float4(m * val.x, m * val.y, m * val.z, m * val.w);
  const float sdm = gc.x * val.x + gc.y * val.y + gc.z * val.z + gc.w * val.w;
  const float sdy = gc.x * (hx * (v10.x - v00.x) + lx * (v11.x - v01.x)) + gc.y * (hx * (v10.y - v00.y) + lx * (v11.y - v01.y)) +
                    gc.z * (hx * (v10.z - v00.z) + lx * (v11.z - v01.z)) + gc.w * (hx * (v10.w - v00.w) + lx * (v11.w - v01.w));
  const float sdx = gc.x * (hy * (v01.x - v00.x) + ly * (v11.x - v10.x)) + gc.y * (hy * (v01.y - v00.y) + ly * (v11.y - v10.y)) +
                    gc.z * (hy * (v01.z - v00.z) + ly * (v11.z - v10.z)) + gc.w * (hy * (v01.w - v00.w) + ly * (v11.w - v10.w));
  const float4 gm = make_float4(gc.x * m, gc.y * m, gc.z * m, gc.w * m);
  if (ok00) atomicAdd(reinterpret_cast<float4*>(D.dx + p00), make_float4(gm.x * w00, gm.y * w00, gm.z * w00, gm.w * w00));
  if (ok01) atomicAdd(reinterpret_cast<float4*>(D.dx + p01), make_float4(gm.x * w01, gm.y * w01, gm.z * w01, gm.w * w01));
  if (ok10) atomicAdd(reinterpret_cast<float4*>(D.dx + p10), make_float4(gm.x * w10, gm.y * w10, gm.z * w10, gm.w * w10));
  if (ok11) atomicAdd(reinterpret_cast<float4*>(D.dx + p11), make_float4(gm.x * w11, gm.y * w11, gm.z * w11, gm.w * w11));
  D.doffset[pix * gts * 2 + gt * 2] = m * sdy;
  D.doffset[pix * gts * 2 + gt * 2 + 1] = m * sdx;
  D.dmask[pix * gts + gt] = sdm;
}

// ------------------------------------------------------------------------------------------------ flow_warp
// Same fp32 coordinate sequence as the forward (warp_resize.cu::warp_coord).  d ix / d flow_x = 1 (the normalise /
// un-normalise pair of CRFP.py:118-121 + grid_sample align_corners=True cancels), zeros padding: only in-range
// corners carry value or gradient (ATen grid_sampler_2d_backward).
__device__ __forceinline__ float warp_coord_bwd(int i, float f, int size) {
  const float denom = (float)((size - 1) > 1 ? (size - 1) : 1);
  const float gf = __fadd_rn((float)i, f);
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, gf), denom), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

// one thread per pixel: loops the channels, dx scattered with atomicAdd, dflow[pix] = (d/dflow_x, d/dflow_y)
__global__ void __launch_bounds__(128) flow_warp_bwd_kernel(int n, int h, int w, int c, const float* __restrict__ xin,
                                                            const float* __restrict__ flow, const float* __restrict__ dy,
                                                            float* __restrict__ dx, float* __restrict__ dflow) {
  const long long total = (long long)n * h * w;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % w);
  const int y = (int)((pix / w) % h);
  const long long b = pix / ((long long)w * h);
  const float ix = warp_coord_bwd(x, flow[pix * 2], w);
  const float iy = warp_coord_bwd(y, flow[pix * 2 + 1], h);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
  const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
  const bool vx0 = (x0 >= 0 && x0 < w), vx1 = (x1 >= 0 && x1 < w);
  const bool vy0 = (y0 >= 0 && y0 < h), vy1 = (y1 >= 0 && y1 < h);
  const bool ok00 = vy0 && vx0, ok01 = vy0 && vx1, ok10 = vy1 && vx0, ok11 = vy1 && vx1;
  const long long img = b * h * w;
  const long long p00 = (img + (long long)y0 * w + x0) * c, p01 = p00 + c, p10 = p00 + (long long)w * c, p11 = p10 + c;
  const float* d = dy + pix * c;
  float gix = 0.f, giy = 0.f;
  for (int ch = 0; ch < c; ++ch) {
    const float go = d[ch];
    const float v00 = ok00 ? xin[p00 + ch] : 0.f, v01 = ok01 ? xin[p01 + ch] : 0.f;
    const float v10 = ok10 ? xin[p10 + ch] : 0.f, v11 = ok11 ? xin[p11 + ch] : 0.f;
    gix += go * (wy0 * (v01 - v00) + wy1 * (v11 - v10));
    giy += go * (wx0 * (v10 - v00) + wx1 * (v11 - v01));
    if (dx != nullptr) {
      if (ok00) atomicAdd(dx + p00 + ch, go * wx0 * wy0);
      if (ok01) atomicAdd(dx + p01 + ch, go * wx1 * wy0);
      if (ok10) atomicAdd(dx + p10 + ch, go * wx0 * wy1);
      if (ok11) atomicAdd(dx + p11 + ch, go * wx1 * wy1);
    }
  }
  if (dflow != nullptr) {
    // chain through ((g+1)/2)*(size-1) and 2*(i+f)/max(size-1,1) - 1, in the order autograd applies them
    const float mx = (float)(w - 1) * 0.5f, my = (float)(h - 1) * 0.5f;
    const float dnx = (float)((w - 1) > 1 ? (w - 1) : 1), dny = (float)((h - 1) > 1 ? (h - 1) : 1);
    dflow[pix * 2] = (gix * mx) * 2.0f / dnx;
    dflow[pix * 2 + 1] = (giy * my) * 2.0f / dny;
  }
}

// ------------------------------------------------------------------------------------------------ resize / pool
__device__ __forceinline__ void bilin_src_bwd(int dst, float rscale, int size, int& i0, int& i1, float& l1) {
  float s = rscale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > size - 1) i0 = size - 1;
  i1 = i0 + ((i0 < size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

// scatter form: one thread per OUTPUT element of the forward, 4 atomicAdds into dx (zero-filled by the caller)
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(int n, int hin, int win, int c, int hout, int wout,
                                                                  float rh, float rw, float mul,
                                                                  const float* __restrict__ dy, float* __restrict__ dx) {
  const long long total = (long long)n * hout * wout * c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  const long long pix = idx / c;
  const int x = (int)(pix % wout);
  const int y = (int)((pix / wout) % hout);
  const long long b = pix / ((long long)wout * hout);
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_src_bwd(y, rh, hin, y0, y1, ly);
  bilin_src_bwd(x, rw, win, x0, x1, lx);
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float go = dy[idx] * mul;
  float* ob = dx + b * hin * win * c + ch;
  atomicAdd(ob + ((long long)y0 * win + x0) * c, go * hy * hx);
  atomicAdd(ob + ((long long)y0 * win + x1) * c, go * hy * lx);
  atomicAdd(ob + ((long long)y1 * win + x0) * c, go * ly * hx);
  atomicAdd(ob + ((long long)y1 * win + x1) * c, go * ly * lx);
}

// AvgPool2d(2,2) backward, gather form: dx[b,y,x,c] = dy[b,y/2,x/2,c] / 4 (0 in a trailing odd row / column)
__global__ void __launch_bounds__(256) avgpool2_bwd_kernel(int n, int hin, int win, int c, const float* __restrict__ dy,
                                                           float* __restrict__ dx) {
  const long long total = (long long)n * hin * win * c;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  const long long pix = idx / c;
  const int x = (int)(pix % win);
  const int y = (int)((pix / win) % hin);
  const long long b = pix / ((long long)win * hin);
  const int ho = hin / 2, wo = win / 2;
  const int yo = y >> 1, xo = x >> 1;
  dx[idx] = (yo < ho && xo < wo) ? 0.25f * dy[((b * ho + yo) * wo + xo) * c + ch] : 0.f;
}

// ------------------------------------------------------------------------------------------------ loss / optimiser
// Charbonnier (loss/loss.py:116-124, reduction 'mean'): loss_sum += sum sqrt(d^2 + eps);
// dpred = grad_scale * d / sqrt(d^2 + eps) with grad_scale = loss_weight / count.
__global__ void __launch_bounds__(256) charbonnier_kernel(long long count, const float* __restrict__ pred,
                                                          const float* __restrict__ target, float eps, float grad_scale,
                                                          float* __restrict__ loss_sum, float* __restrict__ dpred) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float local = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float d = pred[i] - target[i];
    const float s = sqrtf(d * d + eps);
    local += s;
    if (dpred != nullptr) dpred[i] = grad_scale * d / s;
  }
  atomicAdd(loss_sum, local);
}

// torch.optim.Adam (no weight decay, no amsgrad), written out as torch's _single_tensor_adam applies it
__global__ void __launch_bounds__(256) adam_kernel(long long count, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float beta1,
                                                   float beta2, float eps, float step_size, float bc2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.f - beta1);
  const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - step_size * (mi / denom);
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_act_bwd(long long count, int act, const float* dy, const float* out, float* g, crfp_stream stream) {
  if (count < 0) return CRFP_ERR_BAD_SHAPE;
  if (act != CRFP_ACT_LRELU && act != CRFP_ACT_RELU) return CRFP_ERR_UNSUPPORTED;
  if (count == 0) return CRFP_OK;
  if (!dy || !out || !g) return CRFP_ERR_NULL;
  CRFP_LAUNCH(act_bwd_kernel, dim3(blocks_for(count, 256)), dim3(256), (cudaStream_t)stream, count, act, dy, out, g);
  return check_launch();
}

extern "C" int crfp_conv3x3_bwd_data(int n, int h, int w, int cin, int cout, int cin_total, int cin_off, const float* g,
                                     const float* weight_t, float* dx, crfp_stream stream) {
  if (n < 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || cin_off < 0 || cin_off + cin > cin_total) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!g || !weight_t || !dx) return CRFP_ERR_NULL;
  if (cin % 4 == 0 && cout % 4 == 0 && cin_total % 4 == 0 && cin_off % 4 == 0 && aligned16(g) && aligned16(weight_t) &&
      aligned16(dx)) {
    const long long total4 = (long long)n * h * w * (cin / 4);
    CRFP_LAUNCH(conv3x3_bwd_data_v4_kernel, dim3(blocks_for(total4, 256)), dim3(256), (cudaStream_t)stream, n, h, w, cin, cout,
                cin_total, cin_off, g, weight_t, dx);
    return check_launch();
  }
  const long long total = (long long)n * h * w * cin;
  CRFP_LAUNCH(conv3x3_bwd_data_kernel, dim3(blocks_for(total, 256)), dim3(256), (cudaStream_t)stream, n, h, w, cin, cout,
              cin_total, cin_off, g, weight_t, dx);
  return check_launch();
}

extern "C" int crfp_conv3x3_bwd_weight(int n, int h, int w, int cin, int cout, int cin_total, int cin_off, const float* x,
                                       const float* g, float* dw, float* db, float* workspace, size_t ws_floats,
                                       crfp_stream stream) {
  if (n < 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || cin_off < 0 || cin_off + cin > cin_total) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!x || !g || !dw) return CRFP_ERR_NULL;
  return launch_bwd_weight((long long)n * h, h, w, cin, cout, 9, cin_total, cin_off, x, g, dw, db, (cudaStream_t)stream,
                           workspace, ws_floats);
}

// `count` (x, g) pairs of identical shape (the frames of the recurrence) accumulated into the same dw / db: one launch of the
// round-2 kernels per 16 pairs (xs / gs are HOST arrays of device pointers); shapes those kernels do not take go pair by pair
// through crfp_conv3x3_bwd_weight's path.  workspace: crfp_conv3x3_bwd_weight_workspace(n * min(count, 16), ...) floats.
extern "C" int crfp_conv3x3_bwd_weight_batched(int count, const float* const* xs, const float* const* gs, int n, int h, int w,
                                               int cin, int cout, int cin_total, int cin_off, float* dw, float* db,
                                               float* workspace, size_t ws_floats, crfp_stream stream) {
  if (count < 0 || n < 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || cin_off < 0 || cin_off + cin > cin_total) return CRFP_ERR_BAD_SHAPE;
  if (count == 0 || n == 0) return CRFP_OK;
  if (!xs || !gs || !dw) return CRFP_ERR_NULL;
  for (int e = 0; e < count; ++e)
    if (!xs[e] || !gs[e]) return CRFP_ERR_NULL;
  for (int e0 = 0; e0 < count; e0 += 16) {
    const int ne = count - e0 < 16 ? count - e0 : 16;
    int s2 = 1;
#ifndef CRFP_HOST_EMU
    s2 = launch_bwd_weight_v2((long long)n * h, h, w, cin, cout, 9, cin_total, cin_off, xs + e0, gs + e0, ne, dw, db, workspace,
                              ws_floats, (cudaStream_t)stream);
#endif
    if (s2 == 1) {
      for (int e = e0; e < e0 + ne; ++e)
        CRFP_TRY(launch_bwd_weight((long long)n * h, h, w, cin, cout, 9, cin_total, cin_off, xs[e], gs[e], dw, db,
                                   (cudaStream_t)stream, workspace, ws_floats));
    } else if (s2 != CRFP_OK) {
      return s2;
    }
  }
  return CRFP_OK;
}

extern "C" size_t crfp_conv3x3_bwd_weight_workspace(int n, int h, int w, int cin, int cout) {
  if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return 0;
#ifndef CRFP_HOST_EMU
  {
    const size_t v2 = wgrad_workspace_floats((long long)n * h, w, cin, cout, 9);
    if (v2) return v2;
  }
#endif
  int lanes, epad, xsegs;
  long long chunks, rpb;
  thin_plan((long long)n * h, w, cin, cout, 9, (cin % 4 == 0) && (cout % 4 == 0), &lanes, &epad, &chunks, &xsegs, &rpb);
  if (chunks <= 0) return 0;
  // sized for the scalar mapping as well (an unaligned pointer falls back to it): it never needs more rows
  int lanes_s, epad_s, xsegs_s;
  long long chunks_s, rpb_s;
  thin_plan((long long)n * h, w, cin, cout, 9, false, &lanes_s, &epad_s, &chunks_s, &xsegs_s, &rpb_s);
  long long prows = chunks * lanes;
  if (chunks_s * lanes_s > prows) prows = chunks_s * lanes_s;
  return (size_t)(prows * (9LL * cin * cout + cout));
}

extern "C" int crfp_fovea_blend_fwd(long long npix, int c, const float* f, const float* s, const float* mask, float* out,
                                    crfp_stream stream) {
  if (npix < 0 || c <= 0 || c % 4) return CRFP_ERR_BAD_SHAPE;
  if (npix == 0) return CRFP_OK;
  if (!f || !s || !mask || !out) return CRFP_ERR_NULL;
  if (!aligned16(f) || !aligned16(s) || !aligned16(out)) return CRFP_ERR_BAD_SHAPE;
  const long long total4 = npix * (c / 4);
  CRFP_LAUNCH(fovea_blend_fwd_kernel, dim3(blocks_for(total4, 256)), dim3(256), (cudaStream_t)stream, total4, c / 4, f, s, mask, out);
  return check_launch();
}

extern "C" int crfp_fovea_blend_bwd(long long npix, int c, const float* dout, const float* out, const float* mask, float* df,
                                    float* ds, crfp_stream stream) {
  if (npix < 0 || c <= 0 || c % 4) return CRFP_ERR_BAD_SHAPE;
  if (npix == 0) return CRFP_OK;
  if (!dout || !out || !mask || !df || !ds) return CRFP_ERR_NULL;
  if (!aligned16(dout) || !aligned16(out) || !aligned16(df) || !aligned16(ds)) return CRFP_ERR_BAD_SHAPE;
  const long long total4 = npix * (c / 4);
  CRFP_LAUNCH(fovea_blend_bwd_kernel, dim3(blocks_for(total4, 256)), dim3(256), (cudaStream_t)stream, total4, c / 4, dout, out, mask,
              df, ds);
  return check_launch();
}

extern "C" int crfp_dcn_heads_act_fwd(long long npix, int nk, int repeat, float mag, const float* heads, const float* flow,
                                      float* offset, float* mask, crfp_stream stream) {
  if (npix < 0 || nk <= 0) return CRFP_ERR_BAD_SHAPE;
  if (npix == 0) return CRFP_OK;
  if (!heads || !flow || !offset || !mask) return CRFP_ERR_NULL;
  const long long total = npix * nk;
  CRFP_LAUNCH(dcn_heads_act_fwd_kernel, dim3(blocks_for(total, 256)), dim3(256), (cudaStream_t)stream, total, nk, repeat ? 1 : 0, mag,
              heads, flow, offset, mask);
  return check_launch();
}

extern "C" int crfp_dcn_heads_act_bwd(long long npix, int nk, int repeat, float mag, const float* heads, const float* doffset,
                                      const float* dmask, float* dheads, float* dflow, crfp_stream stream) {
  if (npix < 0 || nk <= 0) return CRFP_ERR_BAD_SHAPE;
  if (npix == 0) return CRFP_OK;
  if (!heads || !doffset || !dmask || !dheads || !dflow) return CRFP_ERR_NULL;
  const long long total = npix * (repeat ? 1 : nk);
  CRFP_LAUNCH(dcn_heads_act_bwd_kernel, dim3(blocks_for(total, 256)), dim3(256), (cudaStream_t)stream, total, nk, repeat ? 1 : 0, mag,
              heads, doffset, dmask, dheads, dflow);
  return check_launch();
}

extern "C" int crfp_dcn_v2_bwd(const crfp_dcn_bwd_desc* d, crfp_stream stream) {
  if (!d) return CRFP_ERR_NULL;
  if (d->n < 0 || d->h <= 0 || d->w <= 0 || d->c <= 0 || d->cout <= 0 || d->dg <= 0 || d->c % d->dg != 0)
    return CRFP_ERR_BAD_SHAPE;
  if (d->n == 0) return CRFP_OK;
  if (!d->x || !d->offset || !d->mask || !d->weight || !d->dout || !d->dx || !d->doffset || !d->dmask || !d->col || !d->dweight)
    return CRFP_ERR_NULL;
  const long long total = (long long)d->n * d->h * d->w * d->dg * 9;
  const bool v4 = d->weight_t != nullptr && d->c == 4 * d->dg && d->cout % 4 == 0 && aligned16(d->x) && aligned16(d->dx) &&
                  aligned16(d->dout) && aligned16(d->col) && aligned16(d->weight_t);
  if (v4)
    CRFP_LAUNCH(dcn_bwd_sample_v4_kernel, dim3(blocks_for(total, 128)), dim3(128), (cudaStream_t)stream, *d);
  else
    CRFP_LAUNCH(dcn_bwd_sample_kernel, dim3(blocks_for(total, 128)), dim3(128), (cudaStream_t)stream, *d);
  CRFP_TRY(check_launch());
  // dweight[k][co] += col^T dout, dbias[co] += sum dout
  return launch_bwd_weight((long long)d->n * d->h, d->h, d->w, 9 * d->c, d->cout, 1, 9 * d->c, 0, d->col, d->dout, d->dweight,
                           d->dbias, (cudaStream_t)stream, d->wg_workspace, d->wg_ws_floats);
}

extern "C" size_t crfp_dcn_v2_bwd_workspace(int n, int h, int w, int c, int cout) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || cout <= 0) return 0;
#ifndef CRFP_HOST_EMU
  return wgrad_workspace_floats((long long)n * h, w, 9 * c, cout, 1);
#else
  return 0;
#endif
}

extern "C" int crfp_flow_warp_bwd(int n, int h, int w, int c, const float* x, const float* flow, const float* dy,
                                  float* dx, float* dflow, crfp_stream stream) {
  if (n < 0 || h <= 0 || w <= 0 || c <= 0) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!x || !flow || !dy || (!dx && !dflow)) return CRFP_ERR_NULL;
  const long long total = (long long)n * h * w;
  CRFP_LAUNCH(flow_warp_bwd_kernel, dim3(blocks_for(total, 128)), dim3(128), (cudaStream_t)stream, n, h, w, c, x, flow, dy, dx,
              dflow);
  return check_launch();
}

extern "C" int crfp_resize_bilinear_bwd(int n, int hin, int win, int c, int hout, int wout, float rscale_h, float rscale_w,
                                        float mul, const float* dy, float* dx, crfp_stream stream) {
  if (n < 0 || hin <= 0 || win <= 0 || c <= 0 || hout <= 0 || wout <= 0) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!dy || !dx) return CRFP_ERR_NULL;
  const long long total = (long long)n * hout * wout * c;
  CRFP_LAUNCH(resize_bilinear_bwd_kernel, dim3(blocks_for(total, 256)), dim3(256), (cudaStream_t)stream, n, hin, win, c, hout,
              wout, rscale_h, rscale_w, mul, dy, dx);
  return check_launch();
}

extern "C" int crfp_avgpool2_bwd(int n, int hin, int win, int c, const float* dy, float* dx, crfp_stream stream) {
  if (n < 0 || hin < 2 || win < 2 || c <= 0) return CRFP_ERR_BAD_SHAPE;
  if (n == 0) return CRFP_OK;
  if (!dy || !dx) return CRFP_ERR_NULL;
  const long long total = (long long)n * hin * win * c;
  CRFP_LAUNCH(avgpool2_bwd_kernel, dim3(blocks_for(total, 256)), dim3(256), (cudaStream_t)stream, n, hin, win, c, dy, dx);
  return check_launch();
}

extern "C" int crfp_charbonnier_fwd_bwd(long long count, const float* pred, const float* target, float eps,
                                        float grad_scale, float* loss_sum, float* dpred, crfp_stream stream) {
  if (count <= 0) return CRFP_ERR_BAD_SHAPE;
  if (!pred || !target || !loss_sum) return CRFP_ERR_NULL;
  unsigned blocks = blocks_for(count, 256);
  if (blocks > 592) blocks = 592;  // 4 CTAs per SM x 148 SMs, grid-stride
  CRFP_LAUNCH(charbonnier_kernel, dim3(blocks), dim3(256), (cudaStream_t)stream, count, pred, target, eps, grad_scale,
              loss_sum, dpred);
  return check_launch();
}

extern "C" int crfp_adam_step(long long count, float* p, const float* g, float* m, float* v, float beta1, float beta2,
                              float eps, float step_size, float bc2_sqrt, crfp_stream stream) {
  if (count < 0) return CRFP_ERR_BAD_SHAPE;
  if (count == 0) return CRFP_OK;
  if (!p || !g || !m || !v) return CRFP_ERR_NULL;
  CRFP_LAUNCH(adam_kernel, dim3(blocks_for(count, 256)), dim3(256), (cudaStream_t)stream, count, p, g, m, v, beta1, beta2, eps,
              step_size, bc2_sqrt);
  return check_launch();
}

extern "C" size_t crfp_sizeof_dcn_bwd_desc(void) { return sizeof(crfp_dcn_bwd_desc); }
