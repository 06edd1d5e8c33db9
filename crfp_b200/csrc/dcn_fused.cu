// dcn_align_fused: the tail of DCN_module.forward (/root/reference/model/CRFP.py:337-350) as ONE kernel —
//     offset = 10 * tanh(dcn_offset(z)) + flow.flip(1).repeat(72);  mask = sigmoid(dcn_mask(z));  out = DCNv2(x, offset, mask)
// The 216-channel offset / mask tensor (864 B per L1 pixel, written by one kernel and re-read by the next: 400 MB of HBM
// traffic per level at 360x640) never exists: the head convolution's accumulator stays in TMEM and the sampler threads read
// their own raw offsets straight from it.  SURVEY.md 8(b) `dcn_align_fused`, 8(d) "fused align block".
//
// Persistent, one CTA (512 threads) per SM, tiles of 8 x 16 output pixels (M = 128).  Per tile, two phases that reuse
// the same shared memory (everything here is bound by the shared-memory pipe, so overlapping them would buy little):
//
//  H  heads GEMM  [128 px] x [K = 288 = 9 taps x 32 ch] x [N = 224 = 72 samples x (dy, dx, mask) + 8 pad] on tcgen05 as
//     3 x bf16 split products (A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 in TMEM columns 0..223).
//     A: the 10 x 18 halo tile of z, split hi / lo once, K-major no-swizzle records; tap (ky, kx) is a descriptor START
//        ADDRESS into it and the 8-pixel tile rows are addressed with SBO = 10 records (160 B).
//     B: 258 KB of split weights do not fit: they stream from L2 in six 42 KB "sixths" (6 K chunks each) through a ring of
//        two slots with cp.async.bulk + mbarrier; the next tile's first sixth is prefetched during phase S.
//  S  for each K quarter (2 deformable groups): two TMA tensor-tile loads (one 33 x 40 x 4-channel plane per group) of the sampling window of x
//     (origin shifted by the rounded flow at the tile centre; samples outside it fall back to global loads), 12 sampler
//     warps = (TMEM lane quadrant, third of the quarter's 18 samples): tcgen05.ld of the thread's 18 raw head values, bias,
//     tanh / sigmoid / + flow (same ex2 / rcp arithmetic as the unfused path), bilinear gather with LDS.128, modulate,
//     split hi / lo into the UMMA A stage (thread = pixel: conflict-free stores), 15 tcgen05.mma into TMEM columns
//     224..255; after the 4th quarter warps 0-3 drain the 32-channel result.
//
// Shared memory (204 KB): [W slot 0 42 KB][W slot 1 42 KB | window B 40 KB][DCN weights 40 KB][A stage 40 KB | z hi/lo 23 KB][window A 40 KB].
// Window A is free during phase H, so the first quarter's window is fetched while the heads GEMM runs.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace crfp {

constexpr int FTW = 8, FTH = 16;               // tile
constexpr int FHW = FTW + 2;                   // halo tile width (10); height FTH + 2 = 18
constexpr int FHPX = FHW * (FTH + 2);          // 180 halo pixels
constexpr int FZP = 184;                       // records per 8-channel plane of the z operand
constexpr int FNH = 224;                       // heads GEMM N
constexpr int FWREC = 6 * FNH;                 // records per sixth and half (hi / lo)
constexpr int FWSLOT = 2 * FWREC * 16;         // 43008 B
constexpr int FWR = 12;                        // window reach: 10 (max |residual offset|) + 1 (tap) + 1 (bilinear corner)
constexpr int FWW = FTW + 2 * FWR + 1, FWH = FTH + 2 * FWR;   // 33 x 40 pixels: the odd width is the row pitch of the staged
                                                              // planes (528 B), so the 4 pixel rows of a warp fall into different banks
constexpr int FWPLANE = FWH * FWW;             // float4 records per deformable-group plane
constexpr int FWIN_BYTES = 2 * FWPLANE * 16;   // 42240: two group planes [gl][40][33][4 floats]
constexpr int FQC = 10;                        // K chunks per DCN quarter (9 real + 1 zero)
constexpr int FAP = 128;                       // records per chunk row of the DCN A stage
constexpr int FA_RECS = FQC * FAP;
constexpr int FSAMP = 384;

constexpr int OFF_W0 = 0;
constexpr int OFF_W1 = FWSLOT;                          // also window B
constexpr int OFF_DW = 2 * FWSLOT;                      // DCN weights hi | lo: 2 x 4 x 10 x 32 records
constexpr int OFF_A = OFF_DW + 2 * 4 * FQC * 32 * 16;   // A stage hi | lo
constexpr int OFF_WA = OFF_A + 2 * FA_RECS * 16;        // window A
constexpr int OFF_Z = OFF_A;                            // z hi | lo share the A stage (phase H only)
constexpr int FUSED_SMEM = OFF_WA + FWIN_BYTES;         // 208896
static_assert(FWIN_BYTES <= FWSLOT, "window B must fit in W slot 1");
static_assert(2 * 4 * FZP * 16 <= 2 * FA_RECS * 16, "z operand must fit in the A stage");
static_assert(OFF_W1 % 128 == 0 && OFF_WA % 128 == 0, "TMA destinations must be 128-byte aligned");

struct FusedParams {
  int n, h, w;
  const float* z; int z_cstride, z_coffset;
  const float* flow;
  const float* x; int x_cstride, x_coffset;
  const uint4* heads_w;        // [6 sixths][hi 6*224 | lo 6*224] records
  const float* heads_b;        // [224], sample-major (dy, dx, m) triples
  const __nv_bfloat16* dw_hi;  // [36][32][8]
  const __nv_bfloat16* dw_lo;
  const float* dbias;
  float* out; int out_cstride, out_coffset;
  float head_mag;
  int32_t* dbg_y0;
  int32_t* dbg_x0;
  long long* trace;            // optional clock64 timeline of CTA 0 (profiling aid): [tile][16]
};

__device__ __forceinline__ void fused_split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    const float2 hf = __bfloat1622float2(hh);
    h[k] = *reinterpret_cast<const uint32_t*>(&hh);
    l[k] = umma::pack_bf16(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void fused_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = umma::pack_bf16(a - __bfloat162float(h0), b - __bfloat162float(h1));
}

// HALF = CRFP_PREC_HALF: every activation operand (z, the modulated columns) is ONE fp16 tile, the weights are fp16 hi / lo
template <bool HALF>
__global__ void __launch_bounds__(512, 1) dcn_align_fused_kernel(const FusedParams P, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t w_full[2], w_empty[2], win_full[2], a_full, a_empty, hacc_full, dacc_full, dacc_empty;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_hbias[FNH];
  __shared__ float s_dbias[32];
  __shared__ int2 s_org[2];
  auto sW = [&](int slot) { return reinterpret_cast<uint4*>(smem + (slot ? OFF_W1 : OFF_W0)); };      // W ring slot
  auto sWinP = [&](int slot) { return reinterpret_cast<float*>(smem + (slot ? OFF_W1 : OFF_WA)); };   // window A / B
  uint4* sBh = reinterpret_cast<uint4*>(smem + OFF_DW);          // [4][10][32]
  uint4* sBl = sBh + 4 * FQC * 32;
  uint4* sAh = reinterpret_cast<uint4*>(smem + OFF_A);           // [10][128]
  uint4* sAl = sAh + FA_RECS;
  uint4* sZh = reinterpret_cast<uint4*>(smem + OFF_Z);           // [4][184]
  uint4* sZ1h = reinterpret_cast<uint4*>(smem + OFF_WA);         // second tile of a round: [hi 4][184] | [lo 4][184]
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int tiles_x = (P.w + FTW - 1) / FTW, tiles_y = (P.h + FTH - 1) / FTH;
  const int tiles_img = tiles_x * tiles_y, total = tiles_img * P.n;

  // ---- constant-only prologue (overlaps the previous kernel under PDL)
  pdl_trigger();
  for (int i = tid; i < 36 * 32; i += 512) {   // DCN weights [36][32] -> [4 quarters][10][32], chunk 9 of a quarter is zero
    const int c = i >> 5, co = i & 31;
    const int dst = ((c / 9) * FQC + (c % 9)) * 32 + co;
    umma::cp_async16(sBh + dst, reinterpret_cast<const uint4*>(P.dw_hi) + i, 16u);
    umma::cp_async16(sBl + dst, reinterpret_cast<const uint4*>(P.dw_lo) + i, 16u);
  }
  umma::cp_async_commit();
  for (int i = tid; i < 4 * 32; i += 512) {
    sBh[((i >> 5) * FQC + 9) * 32 + (i & 31)] = make_uint4(0u, 0u, 0u, 0u);
    sBl[((i >> 5) * FQC + 9) * 32 + (i & 31)] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < FNH; i += 512) s_hbias[i] = P.heads_b[i];
  if (tid < 32) s_dbias[tid] = P.dbias[tid];
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&w_full[i], 1); umma::mbar_init(&w_empty[i], 1); umma::mbar_init(&win_full[i], 1); }
    umma::mbar_init(&a_full, FSAMP); umma::mbar_init(&a_empty, 1);
    umma::mbar_init(&hacc_full, 1); umma::mbar_init(&dacc_full, 1); umma::mbar_init(&dacc_empty, 128);
    umma::fence_mbar_init();
    umma::tma_prefetch_desc(&tmap);
  }
  umma::cp_async_wait<0>();
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base_s;
  const uint32_t dacc = taddr + (uint32_t)(2 * FNH);   // DCN accumulator: TMEM columns 448..479
  pdl_wait();   // activations (z, flow, x, out) are only touched from here on

  // Work list: ROUNDS of up to two tiles that share one pass over the streamed head weights (the W stream from L2 is what
  // bounds phase H: 258 KB per pass, so two tiles per pass halve it).  Full rounds take tiles 2*(r*G + b), +1; the
  // remainder (< 2G tiles) is dealt one tile per CTA, then a second one, so no CTA does more than ceil(total / G) tiles.
  const int G = (int)gridDim.x, bx = (int)blockIdx.x;
  const int full_rounds = total / (2 * G);
  const int rem_base = 2 * G * full_rounds;
  const int rem0 = rem_base + bx, rem1 = rem_base + G + bx;
  const int rounds = full_rounds + (rem0 < total ? 1 : 0);

  const uint32_t idesc_h = HALF ? umma::make_idesc_f16(128, FNH) : umma::make_idesc_bf16(128, FNH);
  const uint32_t idesc_d = HALF ? umma::make_idesc_f16(128, 32) : umma::make_idesc_bf16(128, 32);
  // the W stream is owned by ONE thread (warp 0's elected lane); sixth s of a round lands in slot s & 1
  auto load_sixth = [&](int s) {
    umma::mbar_arrive_expect_tx(&w_full[s & 1], (uint32_t)FWSLOT);
    umma::bulk_load(sW(s & 1), reinterpret_cast<const unsigned char*>(P.heads_w) + (size_t)s * FWSLOT, (uint32_t)FWSLOT, &w_full[s & 1]);
  };
  if (warp == 0 && rounds > 0 && umma::elect_one()) load_sixth(0);
  __syncwarp();

  int kq = 0;        // quarters processed so far by this CTA (barrier phase bookkeeping)
  int ntile = 0;     // tiles processed so far
  for (int rd = 0; rd < rounds; ++rd) {
    int tl[2];
    if (rd < full_rounds) { tl[0] = 2 * (rd * G + bx); tl[1] = tl[0] + 1; }
    else { tl[0] = rem0; tl[1] = rem1 < total ? rem1 : -1; }
    const int nt = tl[1] >= 0 ? 2 : 1;
    int tn[2], ty0[2], tx0[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t_ = tl[u] >= 0 ? tl[u] : 0;
      tn[u] = t_ / tiles_img;
      const int trm = t_ - tn[u] * tiles_img;
      const int tyi = trm / tiles_x;
      ty0[u] = tyi * FTH; tx0[u] = (trm - tyi * tiles_x) * FTW;
    }
    long long* tr = (P.trace != nullptr && blockIdx.x == 0 && rd < 8) ? P.trace + rd * 32 : nullptr;
    if (tr && tid == 0) tr[0] = clock64();
    // flow at the tile centres: centres the sampling windows (loaded here: its latency hides behind the z conversion)
    float2 flc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    if (warp == 0) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int cy = min(ty0[u] + FTH / 2, P.h - 1), cx = min(tx0[u] + FTW / 2, P.w - 1);
        flc[u] = __ldg(reinterpret_cast<const float2*>(P.flow + (((size_t)tn[u] * P.h + cy) * (size_t)P.w + cx) * 2));
      }
    }
    // ================================================================ phase H: z halos -> hi / lo operands
    // tile 0's operand lives in the A stage, tile 1's in window A (both idle during phase H)
    {
      float4 za[3], zb[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {     // 2 x 720 records over 512 threads: all loads of a thread in flight before the split
        const int i = tid + u * 512;
        za[u] = make_float4(0.f, 0.f, 0.f, 0.f); zb[u] = za[u];
        const int sel = i >= FHPX * 4 ? 1 : 0, il = i - sel * FHPX * 4;
        if (i < nt * FHPX * 4) {
          const int p = il >> 2, c8 = il & 3;
          const int hy = p / FHW, hx = p - hy * FHW;
          const int gy = ty0[sel] - 1 + hy, gx = tx0[sel] - 1 + hx;
          if (gy >= 0 && gy < P.h && gx >= 0 && gx < P.w) {
            const float4* g = reinterpret_cast<const float4*>(P.z + (((size_t)tn[sel] * P.h + gy) * (size_t)P.w + gx) * P.z_cstride + P.z_coffset + c8 * 8);
            za[u] = __ldg(g); zb[u] = __ldg(g + 1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = tid + u * 512;
        const int sel = i >= FHPX * 4 ? 1 : 0, il = i - sel * FHPX * 4;
        if (i < nt * FHPX * 4) {
          uint4* zh = sel ? sZ1h : sZh;
          if (HALF) {
            zh[(il & 3) * FZP + (il >> 2)] = make_uint4(umma::pack_f16(za[u].x, za[u].y), umma::pack_f16(za[u].z, za[u].w),
                                                        umma::pack_f16(zb[u].x, zb[u].y), umma::pack_f16(zb[u].z, zb[u].w));
          } else {
            uint4 hi, lo;
            fused_split8(za[u], zb[u], hi, lo);
            zh[(il & 3) * FZP + (il >> 2)] = hi;
            zh[4 * FZP + (il & 3) * FZP + (il >> 2)] = lo;
          }
        }
      }
    }
    umma::fence_proxy_async();
    __syncthreads();
    if (tr && tid == 0) tr[1] = clock64();   // z operands ready

    // window of global quarter k (tile k >> 2 of the round, K quarter k & 3) into slot k & 1
    auto issue_window = [&](int k) {
      const int u = k >> 2, q = k & 3;
      const int wy0 = ty0[u] - FWR + (int)rintf(fminf(fmaxf(flc[u].y, -4096.f), 4096.f));
      const int wx0 = tx0[u] - FWR + (int)rintf(fminf(fmaxf(flc[u].x, -4096.f), 4096.f));
      const int slot = k & 1;
      s_org[slot] = make_int2(wy0, wx0);
      umma::mbar_arrive_expect_tx(&win_full[slot], (uint32_t)FWIN_BYTES);
      umma::tma_load_4d(sWinP(slot), &tmap, &win_full[slot], 8 * q, wx0, wy0, tn[u]);                    // group 2q
      umma::tma_load_4d(sWinP(slot) + FWPLANE * 4, &tmap, &win_full[slot], 8 * q + 4, wx0, wy0, tn[u]);   // group 2q + 1
    };
    if (warp == 0) {
      if (umma::elect_one()) {
        umma::fence_after_sync();
        load_sixth(1);     // slot 1 (= window B of the previous round) is free since the round-end barrier
        const uint64_t dZh = umma::make_desc(umma::smem_u32(sZh), FZP * 16, FHW * 16), dZl = umma::make_desc(umma::smem_u32(sZh + 4 * FZP), FZP * 16, FHW * 16);
        const uint64_t dYh = umma::make_desc(umma::smem_u32(sZ1h), FZP * 16, FHW * 16), dYl = umma::make_desc(umma::smem_u32(sZ1h + 4 * FZP), FZP * 16, FHW * 16);
        const uint32_t zhl = (uint32_t)dZh, zhh = (uint32_t)(dZh >> 32), zll = (uint32_t)dZl, zlh = (uint32_t)(dZl >> 32);
        const uint32_t yhl = (uint32_t)dYh, yhh = (uint32_t)(dYh >> 32), yll = (uint32_t)dYl, ylh = (uint32_t)(dYl >> 32);
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
          const int slot = s & 1;
          umma::mbar_wait_safe(&w_full[slot], (uint32_t)((rd * 3 + (s >> 1)) & 1));
          if (tr && s == 0) tr[2] = clock64();   // first W sixth landed
          const uint64_t dWh = umma::make_desc(umma::smem_u32(sW(slot)), FNH * 16, 128);
          const uint64_t dWl = umma::make_desc(umma::smem_u32(sW(slot) + FWREC), FNH * 16, 128);
          const uint32_t whl = (uint32_t)dWh, whh = (uint32_t)(dWh >> 32), wll = (uint32_t)dWl, wlh = (uint32_t)(dWl >> 32);
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            const int ks = s * 3 + jj, tap = ks >> 1, half = ks & 1;
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t arec = (uint32_t)((2 * half) * FZP + ky * FHW + kx), brec = (uint32_t)(2 * jj * FNH);
            const uint64_t dbh = umma::desc_advance(whl, whh, brec), dbl = umma::desc_advance(wll, wlh, brec);
            const uint64_t dah = umma::desc_advance(zhl, zhh, arec), dal = umma::desc_advance(zll, zlh, arec);
            umma::mma_bf16(taddr, dah, dbh, idesc_h, ks != 0 ? 1u : 0u);
            if (!HALF) umma::mma_bf16(taddr, dal, dbh, idesc_h, 1u);
            umma::mma_bf16(taddr, dah, dbl, idesc_h, 1u);
            if (nt == 2) {   // the second tile of the round reuses the sixth while it is resident
              const uint64_t dch = umma::desc_advance(yhl, yhh, arec), dcl = umma::desc_advance(yll, ylh, arec);
              umma::mma_bf16(taddr + (uint32_t)FNH, dch, dbh, idesc_h, ks != 0 ? 1u : 0u);
              if (!HALF) umma::mma_bf16(taddr + (uint32_t)FNH, dcl, dbh, idesc_h, 1u);
              umma::mma_bf16(taddr + (uint32_t)FNH, dch, dbl, idesc_h, 1u);
            }
          }
          umma::mma_commit(&w_empty[slot]);
          if (s == 5) umma::mma_commit(&hacc_full);
          if (s >= 1) {   // sixth s-1 is consumed: refill its slot with sixth s+1, or (s == 5) the NEXT round's first sixth
            umma::mbar_wait_safe(&w_empty[(s - 1) & 1], (uint32_t)((rd * 3 + ((s - 1) >> 1)) & 1));
            if (s < 5) load_sixth(s + 1);
            else if (rd + 1 < rounds) load_sixth(0);
          }
        }
      }
      __syncwarp();
    }

    // ================================================================ phase S: 4 quarters per tile, tiles back to back
    umma::mbar_wait_safe(&hacc_full, (uint32_t)(rd & 1));   // heads accumulators complete; both z operands and W slot 1 are free
    umma::fence_after_sync();
    if (tr && tid == 0) tr[3] = clock64();     // heads GEMMs done
    const int nk = 4 * nt;
    if (warp < 4) {
      // epilogue of tile u of the round: thread = pixel = TMEM lane
      auto epilogue = [&](int u) {
        umma::mbar_wait_safe(&dacc_full, (uint32_t)((ntile + u) & 1));
        umma::fence_after_sync();
        float v[32];
        umma::tmem_ld32(dacc + ((uint32_t)(32 * warp) << 16), v);
        umma::fence_before_sync();
        umma::mbar_arrive(&dacc_empty);    // the accumulator may be overwritten by the next tile's first quarter
        const int y = ty0[u] + (tid >> 3), x = tx0[u] + (tid & 7);
        if (y < P.h && x < P.w) {
          const size_t pix = ((size_t)tn[u] * P.h + y) * (size_t)P.w + x;
          float* op = P.out + pix * P.out_cstride + P.out_coffset;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[4 * j] + s_dbias[4 * j], v[4 * j + 1] + s_dbias[4 * j + 1],
                                                                 v[4 * j + 2] + s_dbias[4 * j + 2], v[4 * j + 3] + s_dbias[4 * j + 3]);
        }
      };
      if (warp == 0) {
        if (umma::elect_one()) { issue_window(0); issue_window(1); }
        __syncwarp();
        for (int k = 0; k < nk; ++k) {
          const int q = k & 3;
          umma::mbar_wait_safe(&a_full, (uint32_t)((kq + k) & 1));
          if (k == 4) umma::mbar_wait_safe(&dacc_empty, (uint32_t)(ntile & 1));   // tile 0's result has left TMEM (all 4 warps)
          umma::fence_after_sync();
          if (umma::elect_one()) {
            const uint64_t dAh = umma::make_desc(umma::smem_u32(sAh), FAP * 16, 128), dAl = umma::make_desc(umma::smem_u32(sAl), FAP * 16, 128);
            const uint64_t dBh = umma::make_desc(umma::smem_u32(sBh + q * FQC * 32), 32 * 16, 128);
            const uint64_t dBl = umma::make_desc(umma::smem_u32(sBl + q * FQC * 32), 32 * 16, 128);
            const uint32_t ahl = (uint32_t)dAh, ahh = (uint32_t)(dAh >> 32), all_ = (uint32_t)dAl, alh = (uint32_t)(dAl >> 32);
            const uint32_t bhl = (uint32_t)dBh, bhh = (uint32_t)(dBh >> 32), bll = (uint32_t)dBl, blh = (uint32_t)(dBl >> 32);
#pragma unroll
            for (int ks = 0; ks < FQC / 2; ++ks) {
              const uint64_t dah = umma::desc_advance(ahl, ahh, 2 * ks * FAP), dal = umma::desc_advance(all_, alh, 2 * ks * FAP);
              const uint64_t dbh = umma::desc_advance(bhl, bhh, 2 * ks * 32), dbl = umma::desc_advance(bll, blh, 2 * ks * 32);
              umma::mma_bf16(dacc, dah, dbh, idesc_d, (q | ks) != 0 ? 1u : 0u);
              if (!HALF) umma::mma_bf16(dacc, dal, dbh, idesc_d, 1u);
              umma::mma_bf16(dacc, dah, dbl, idesc_d, 1u);
            }
            umma::mma_commit(&a_empty);
            if (q == 3) umma::mma_commit(&dacc_full);
            if (k + 2 < nk) issue_window(k + 2);   // a_full(k): every sampler is done with window k, its slot is free
          }
          __syncwarp();
          if (q == 3) epilogue(k >> 2);   // warp 0's quarter of the tile's result (drained while the samplers go on)
        }
      } else {
        for (int u = 0; u < nt; ++u) epilogue(u);   // warps 1-3 wait here from the start of phase S
      }
    } else {
      // ---- samplers: warp = (TMEM lane quadrant warp % 4, third j of the quarter's 18 samples); thread = pixel
      const int jthird = (warp - 4) >> 2;
      const int m = 32 * (warp & 3) + lane;
      for (int k = 0; k < nk; ++k) {
        const int u = k >> 2, q = k & 3;
        const int y = ty0[u] + (m >> 3), x = tx0[u] + (m & 7);
        const bool valid = y < P.h && x < P.w;
        const size_t pix = ((size_t)tn[u] * P.h + (valid ? y : 0)) * (size_t)P.w + (valid ? x : 0);
        const float* img = P.x + (size_t)tn[u] * P.h * P.w * P.x_cstride + P.x_coffset;
        const float2 fl = valid ? __ldg(reinterpret_cast<const float2*>(P.flow + pix * 2)) : make_float2(0.f, 0.f);
        // raw head outputs of this thread's 6 samples: 18 consecutive TMEM columns (dy, dx, m per sample)
        const int col = 54 * q + 18 * jthird;
        float raw[18];
        umma::tmem_ld16(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(u * FNH + col), raw);
        umma::tmem_ld2(taddr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(u * FNH + col + 16), raw + 16);
#pragma unroll
        for (int i = 0; i < 18; ++i) raw[i] += s_hbias[col + i];
        if (tr && tid == 128) tr[4 + 3 * k] = clock64();        // sampler: raw offsets in registers
        umma::mbar_wait_safe(&win_full[k & 1], (uint32_t)(((kq + k) >> 1) & 1));
        if (tr && tid == 128) tr[5 + 3 * k] = clock64();        // window landed
        const int2 org = s_org[k & 1];
        const float4* win = reinterpret_cast<const float4*>(sWinP(k & 1));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int kl = 3 * jthird + r;
          // both samples of the record: positions, corner weights and window addresses first, then all 8 LDS.128 back to
          // back (no branch in between), then the arithmetic; samples that leave the window are redone from global
          // memory afterwards (rare)
          float sv[2][4];
          int sy0[2], sx0[2];
          float sw[2][4], smk[2], spy[2], spx[2];
          bool sin[2];
          const float4* sp[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int gtr = 2 * kl + e, gl = gtr / 9, t = gtr - gl * 9;
            const int i = t / 3, j = t - i * 3;
            const float dy = head_offset_act(raw[6 * r + 3 * e + 0], P.head_mag, fl.y);
            const float dx = head_offset_act(raw[6 * r + 3 * e + 1], P.head_mag, fl.x);
            smk[e] = head_mask_act(raw[6 * r + 3 * e + 2]);
            // the staged planes are zero outside the image (TMA out-of-bounds fill), so in-window samples need no
            // per-corner validity: out-of-image corners contribute w * 0 exactly like the reference's zero padding
            spy[e] = dcn_pos(y, i, dy); spx[e] = dcn_pos(x, j, dx);
            const float fy = floorf(spy[e]), fx = floorf(spx[e]);
            sy0[e] = (int)fy; sx0[e] = (int)fx;
            const float ly = spy[e] - fy, lx = spx[e] - fx, hy = 1.f - ly, hx = 1.f - lx;
            sw[e][0] = hy * hx; sw[e][1] = hy * lx; sw[e][2] = ly * hx; sw[e][3] = ly * lx;
            const int wy = sy0[e] - org.x, wx = sx0[e] - org.y;
            sin[e] = wy >= 0 && wy + 1 < FWH && wx >= 0 && wx + 1 < FWW;
            sp[e] = win + gl * FWPLANE + (sin[e] ? wy * FWW + wx : 0);
          }
          float4 c[2][4];
#pragma unroll
          for (int e = 0; e < 2; ++e) { c[e][0] = sp[e][0]; c[e][1] = sp[e][1]; c[e][2] = sp[e][FWW]; c[e][3] = sp[e][FWW + 1]; }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float2 zz = make_float2(0.f, 0.f), mm = make_float2(smk[e], smk[e]);
            float2 a = zz, b = zz;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const float2 kw = make_float2(sw[e][kk], sw[e][kk]);
              a = __ffma2_rn(kw, make_float2(c[e][kk].x, c[e][kk].y), a);
              b = __ffma2_rn(kw, make_float2(c[e][kk].z, c[e][kk].w), b);
            }
            a = __ffma2_rn(a, mm, zz); b = __ffma2_rn(b, mm, zz);
            sv[e][0] = a.x; sv[e][1] = a.y; sv[e][2] = b.x; sv[e][3] = b.y;
          }
          if (valid) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gtr = 2 * kl + e;
              if (P.dbg_y0 != nullptr) {
                const long long dbg = (long long)pix * 72 + q * 18 + gtr;
                P.dbg_y0[dbg] = sy0[e]; P.dbg_x0[dbg] = sx0[e];
              }
              if (!sin[e]) {   // outside the staged window: global gather with the full per-corner validity
                const int gl = gtr / 9;
                int gy0, gx0;
                dcn_corner_w(spy[e], spx[e], P.h, P.w, gy0, gx0, sw[e][0], sw[e][1], sw[e][2], sw[e][3]);
                const float* p = img + ((long long)sy0[e] * P.w + sx0[e]) * P.x_cstride + (2 * q + gl) * 4;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
#define CRFP_C4(ptr, wgt)                                                  \
  if ((wgt) != 0.f) {                                                      \
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(ptr));         \
    v[0] += (wgt) * t4.x; v[1] += (wgt) * t4.y; v[2] += (wgt) * t4.z; v[3] += (wgt) * t4.w; \
  }
                CRFP_C4(p, sw[e][0])
                CRFP_C4(p + P.x_cstride, sw[e][1])
                CRFP_C4(p + (long long)P.w * P.x_cstride, sw[e][2])
                CRFP_C4(p + (long long)P.w * P.x_cstride + P.x_cstride, sw[e][3])
#undef CRFP_C4
                sv[e][0] = v[0] * smk[e]; sv[e][1] = v[1] * smk[e]; sv[e][2] = v[2] * smk[e]; sv[e][3] = v[3] * smk[e];
              }
            }
          }
          uint4 rh = make_uint4(0u, 0u, 0u, 0u), rl = rh;
          if (valid) {
            if (HALF) {
              rh = make_uint4(umma::pack_f16(sv[0][0], sv[0][1]), umma::pack_f16(sv[0][2], sv[0][3]),
                              umma::pack_f16(sv[1][0], sv[1][1]), umma::pack_f16(sv[1][2], sv[1][3]));
            } else {
              fused_split_pair(sv[0][0], sv[0][1], rh.x, rl.x);
              fused_split_pair(sv[0][2], sv[0][3], rh.y, rl.y);
              fused_split_pair(sv[1][0], sv[1][1], rh.z, rl.z);
              fused_split_pair(sv[1][2], sv[1][3], rh.w, rl.w);
            }
          }
          if (r == 0) {
            if (k > 0) umma::mbar_wait_safe(&a_empty, (uint32_t)((kq + k - 1) & 1));   // the MMAs of the previous quarter have read the A stage
            else if (tid - 128 < FAP) {   // a z operand lived here during phase H: restore the zero K chunk of the stage
              sAh[9 * FAP + (tid - 128)] = make_uint4(0u, 0u, 0u, 0u);
              sAl[9 * FAP + (tid - 128)] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
          sAh[kl * FAP + m] = rh;
          if (!HALF) sAl[kl * FAP + m] = rl;
        }
        umma::fence_proxy_async();
        umma::mbar_arrive(&a_full);
        if (tr && tid == 128) tr[6 + 3 * k] = clock64();        // quarter sampled
      }
      umma::fence_before_sync();
    }
    kq += nk;
    ntile += nt;
    __syncthreads();   // round end: TMEM, the A stage, both windows / z operands / W slot 1 may be reused
    umma::fence_after_sync();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, 512);
}

typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn fused_tmap_encoder() {
  static tmap_encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (tmap_encode_fn)p;
  }();
  return fn;
}

int launch_align_fused(const crfp_align_fused_desc& d, cudaStream_t st, long long* trace) {
  if ((long long)d.n * d.h * d.w == 0) return CRFP_OK;
  if (!d.z || !d.flow || !d.x || !d.heads_w || !d.heads_b || !d.dcn_w_hi || !d.dcn_w_lo || !d.dcn_b || !d.out) return CRFP_ERR_NULL;
  if ((d.dbg_y0 != nullptr) != (d.dbg_x0 != nullptr)) return CRFP_ERR_NULL;
  if (d.n < 0 || d.h <= 0 || d.w <= 0) return CRFP_ERR_BAD_SHAPE;
  if ((d.z_cstride | d.z_coffset | d.x_cstride | d.x_coffset | d.out_cstride | d.out_coffset) & 3) return CRFP_ERR_BAD_SHAPE;
  if (((uintptr_t)d.heads_w & 15) || ((uintptr_t)(d.z + d.z_coffset) & 15) || ((uintptr_t)(d.out + d.out_coffset) & 15)) return CRFP_ERR_BAD_SHAPE;
  tmap_encode_fn enc = fused_tmap_encoder();
  if (!enc) return CRFP_ERR_UNSUPPORTED;
  const float* base = d.x + d.x_coffset;
  if ((uintptr_t)base & 15) return CRFP_ERR_BAD_SHAPE;
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {32, (cuuint64_t)d.w, (cuuint64_t)d.h, (cuuint64_t)d.n};
  const cuuint64_t gstr[3] = {(cuuint64_t)d.x_cstride * 4, (cuuint64_t)d.w * d.x_cstride * 4, (cuuint64_t)d.h * d.w * d.x_cstride * 4};
  const cuuint32_t box[4] = {4, FWW, FWH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return CRFP_ERR_UNSUPPORTED;
  FusedParams p;
  p.n = d.n; p.h = d.h; p.w = d.w;
  p.z = d.z; p.z_cstride = d.z_cstride; p.z_coffset = d.z_coffset;
  p.flow = d.flow;
  p.x = d.x; p.x_cstride = d.x_cstride; p.x_coffset = d.x_coffset;
  p.heads_w = reinterpret_cast<const uint4*>(d.heads_w); p.heads_b = d.heads_b;
  p.dw_hi = reinterpret_cast<const __nv_bfloat16*>(d.dcn_w_hi); p.dw_lo = reinterpret_cast<const __nv_bfloat16*>(d.dcn_w_lo);
  p.dbias = d.dcn_b;
  p.out = d.out; p.out_cstride = d.out_cstride; p.out_coffset = d.out_coffset;
  p.head_mag = d.head_mag;
  p.dbg_y0 = d.dbg_y0; p.dbg_x0 = d.dbg_x0;
  p.trace = trace;
  void (*kern)(const FusedParams, const CUtensorMap) = d.half ? dcn_align_fused_kernel<true> : dcn_align_fused_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const int total = ceil_div(d.w, FTW) * ceil_div(d.h, FTH) * d.n;
  const int grid = total < sms ? total : sms;
  launch_k_ws(kern, dim3(grid), dim3(512), (size_t)FUSED_SMEM, st, p, tmap);
  return check_launch();
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_dcn_align_fused(const crfp_align_fused_desc* d, crfp_stream stream) {
  if (!d) return CRFP_ERR_NULL;
  return launch_align_fused(*d, (cudaStream_t)stream, nullptr);
}

// profiling aid: same launch + a clock64 timeline of CTA 0's first 16 tiles into trace[16][16] (device int64):
// [0] tile start, [1] z operand ready, [2] first W sixth landed, [3] heads GEMM done, then per quarter q:
// [4+3q] sampler has its raw offsets, [5+3q] window landed and A stage free, [6+3q] quarter sampled
extern "C" int crfp_dcn_align_fused_trace(const crfp_align_fused_desc* d, long long* trace, crfp_stream stream) {
  if (!d || !trace) return CRFP_ERR_NULL;
  return launch_align_fused(*d, (cudaStream_t)stream, trace);
}

extern "C" size_t crfp_sizeof_align_fused_desc(void) { return sizeof(crfp_align_fused_desc); }
