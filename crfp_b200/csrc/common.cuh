// Shared helpers for libcrfp_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string.h>
#include <stdint.h>

#include "../../include/crfp_b200.h"
#include "dcn_pos.cuh"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcrfp_b200 is written for sm_100a (B200) only"
#endif

namespace crfp {

// ---- status / launch accounting (thread-local, no global mutable state shared between callers)
void note_cuda_error(cudaError_t e);
void count_launch();

inline int check_launch() {
  cudaError_t e = cudaPeekAtLastError();
  count_launch();
  if (e != cudaSuccess) {
    note_cuda_error(e);
    (void)cudaGetLastError();
    return CRFP_ERR_CUDA;
  }
  return CRFP_OK;
}

#define CRFP_TRY(expr)          \
  do {                          \
    int _s = (expr);            \
    if (_s != CRFP_OK) return _s; \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL): a kernel launched with programmaticStreamSerialization may start while
// its predecessor is still running; its launch latency, CTA scheduling and constant-only prologue (weights -> smem,
// TMEM allocation, mbarrier init) then overlap the predecessor's tail.  pdl_wait() (griddepcontrol.wait) precedes
// the first access to any activation.
// CRFP_PDL = all | ws | none (default ws): measured in the whole-frame chain, early launch pays between the persistent
// one-CTA-per-SM tensor-core kernels (their weight / TMEM / barrier prologue hides behind the previous kernel's tail)
// and costs ~3 % when the multi-wave SIMT kernels are launched early as well.
bool pdl_enabled(bool persistent);
template <bool PERSISTENT, typename... KArgs, typename... Args>
inline void launch_k_impl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled(PERSISTENT) ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_k_impl<false>(kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
template <typename... KArgs, typename... Args>
inline void launch_k_ws(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_k_impl<true>(kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- internal conv description (superset of the public crfp_conv_desc)
enum {
  EPI_STD = 0,       // bias, act, residual, post_scale, NHWC segments / pixel shuffle
  EPI_BLEND = 1,     // thin only: S = lrelu(m*v + (1-m)*S_old)       (model/CRFP.py:1674-1675)
  EPI_OUT_NCHW = 2   // thin only: planar NCHW output + bilinear x8 base of the LR frame (model/CRFP.py:1678-1683)
};

struct ConvParams {
  int n, h, w;
  int nsrc;
  const float* src[3];
  int src_c[3], src_cstride[3], src_coffset[3], src_mode[3];
  int qstart[4];     // first packed quad of each source; qstart[nsrc] = total real quads
  int cin_packed;    // multiple of 8
  int cout;          // real
  int cout_packed;   // wide: multiple of 32; thin: 4
  int act;
  const float* weight;
  const float* bias;
  int out_mode, shuffle_r;
  int ndst;
  float* dst[2];
  int dst_c[2], dst_cstride[2], dst_coffset[2];
  const float* residual;
  int res_cstride, res_coffset;
  const float* flow;
  int head_split;
  float post_scale, head_mag;
  // special epilogues (thin kernels)
  int epi;
  const uint8_t* mask;            // EPI_BLEND: bool (1,H,W) per clip
  long long mask_clip_stride;
  const float* blend_old;         // EPI_BLEND: NHWC 4ch tensor holding S_old
  const float* fg;                // optional regional mask fp32 (1,H,W) per clip multiplied into every source
  long long fg_clip_stride;
  const float* base_lr4;          // EPI_OUT_NCHW: NHWC4 LR frame
  long long base_clip_stride;
  long long out_clip_stride;      // EPI_OUT_NCHW: floats between clips of the planar output
  int out_planes;                 // EPI_OUT_NCHW: number of planes written (3)
  int out_bf16;                   // wide kernel: destinations are bf16 (strides / offsets in bf16 elements)
  // fast per-quad addressing for the thin kernels (filled by launch_conv_thin): kind 0 = aligned float4,
  // 1 = aligned float2 (+2 zero channels), 2 = generic (load_quad)
  const float* qptr[3];
  int qcs[3], qkind[3];
  // fovea tile skipping (thin4 kernels, 32x32 tiles): flags[n][tiles_y][tiles_x] != 0 where the conv is needed;
  // tile_mode 1: untouched tiles are skipped; 2 (EPI_BLEND): untouched tiles get S = lrelu(S_old) without the conv
  const uint8_t* tile_flags;
  int tiles_x, tiles_y, tile_mode;
};

// ---- tensor-core (tcgen05) conv description: bf16 NHWC sources, channel counts multiples of 8
enum { TC_OUT_BF16 = 0, TC_OUT_F32 = 1, TC_OUT_SHUFFLE_F32 = 2 };

struct TcParams {
  int n, h, w;
  int nsrc;
  const __nv_bfloat16* src[3];
  int src_c[3], src_cstride[3], src_coffset[3];
  int kstart[3];          // first 8-channel chunk of each source
  int kc_real, kc_total;  // chunks of 8 input channels; kc_total = kc_real rounded up to even (K step = 16)
  int cout, nt, ntiles;   // nt = UMMA N (cout tile, multiple of 16, <= 128)
  const __nv_bfloat16* weight;  // [ntiles][9][kc_total][nt][8]
  const float* bias;            // [ntiles*nt]
  int act;
  int out_kind, shuffle_r;
  int ndst;
  void* dst[2];
  int dst_c[2], dst_cstride[2], dst_coffset[2];
  const __nv_bfloat16* residual;
  int res_cstride, res_coffset;
  const float* flow;
  int head_split;
  float head_mag, post_scale;
  int rows_per_cta;
};
void tc_cout_tile(int cout, int* nt, int* ntiles);
int launch_conv_tc(TcParams p, cudaStream_t st);

// ---- fp32-accurate tensor-core conv (3 x bf16 split): fp32 NHWC sources, channel counts multiples of 8
struct Tc3Params {
  int n, h, w;
  int nsrc;
  const float* src[3];
  int src_c[3], src_cstride[3], src_coffset[3];
  int src_mode[3];        // CRFP_SRC_PLAIN or CRFP_SRC_UNSHUFFLE4 (dense 4-channel (4h x 4w) plane, c = 64)
  int kstart[3];
  int kc_real, kc_total;
  int cout, nt, ntiles;
  const __nv_bfloat16* weight_hi;  // [ntiles][9][kc_total][nt][8]
  const __nv_bfloat16* weight_lo;
  const float* bias;               // [ntiles*nt]
  const float* extra;              // optional NHWC 2-channel fp32 source convolved in the epilogue (flow)
  const float* w_extra;            // fp32 [9][2][ntiles*nt]
  const float* fg;                 // optional regional mask (1,h,w) per clip multiplied into the sources
  long long fg_clip_stride;
  int act;
  int out_kind, shuffle_r;
  int ndst;
  void* dst[2];
  int dst_c[2], dst_cstride[2], dst_coffset[2];
  const float* residual;
  int res_cstride, res_coffset;
  const float* flow;
  int head_split;
  float head_mag, post_scale;
  int rows_per_cta;
  int ncat;         // ws kernel: hi | lo weights as one 64-wide B operand (2 MMAs per K step)
  int half;         // CRFP_PREC_HALF: fp16 operands, activations as ONE product (no lo half), weights fp16 hi / lo
  int fast16;       // ws kernel: coalesced 16x256b epilogue for the plain 32-channel tiles (CRFP_TC3_NOFAST16 turns it off)
  int res_pre;      // residual is added BEFORE the activation (K-split layers: the partial sums of the earlier passes)
  long long* dbg;   // optional per-phase clock64 trace of CTA (0,0,0): [row][8] (debug / profiling only)
};
int tc3_cout_tile(int cout, int kc_real, int* nt, int* ntiles);
int launch_conv_tc3(Tc3Params p, cudaStream_t st);

int conv_params_from_desc(const crfp_conv_desc* d, ConvParams* p);
int launch_conv_wide(const ConvParams& p, cudaStream_t st);
int launch_conv_thin(const ConvParams& p, cudaStream_t st);
int launch_conv(const ConvParams& p, cudaStream_t st);  // dispatch on cout
int launch_flow_warp(const crfp_warp_desc& d, cudaStream_t st);
int launch_flow_warp_l1(int n, int h, int w, const float* flow, const float* P, float* P_w, const float* state, float* cur0,
                        float* cur1, float* cur2, cudaStream_t st);
int launch_dcn(const crfp_dcn_desc& d, cudaStream_t st);
int launch_dcn_tc(const crfp_dcn_desc& d, cudaStream_t st);
int launch_dcn_tc3(const crfp_dcn_desc& d, const void* w_lo, const float* flow_hint, cudaStream_t st);
// round-2 weight-gradient kernels (wgrad.cu); launch_bwd_weight_v2 returns 1 when the shape is not theirs
size_t wgrad_workspace_floats(long long rows, int w, int cin, int cout, int taps);
int launch_bwd_weight_v2(long long rows, int h, int w, int cin, int cout, int taps, int cin_total, int cin_off,
                         const float* const* xs, const float* const* gs, int nent, float* dw, float* db, float* workspace,
                         size_t ws_floats, cudaStream_t st);
int launch_align_fused(const crfp_align_fused_desc& d, cudaStream_t st, long long* trace = nullptr);
int launch_flow_warp_bf16(const crfp_warp_desc& d, cudaStream_t st);
int launch_flow_up2_dual(int n, int h, int w, const float* flow, float* out_f32, void* out_bf8, cudaStream_t st);

// ---- device helpers
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float lrelu01(float v) { return v > 0.f ? v : 0.1f * v; }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == CRFP_ACT_LRELU) return lrelu01(v);
  if (act == CRFP_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }
// MUFU-based versions for the tensor-core epilogues (abs error ~1e-7, far inside the parity budget)
__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
__device__ __forceinline__ float fast_tanh(float v) { return 1.f - __fdividef(2.f, __expf(2.f * v) + 1.f); }

// raw MUFU ops + the DCN head activations built on them (model/CRFP.py:337-349): shared by the conv epilogue
// (CRFP_ACT_DCN_HEAD) and by the align kernel's sampler (crfp_dcn_desc.head_raw) so both give identical bits
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// mag * tanh(v) + fl = (mag + fl) - 2 mag / (exp(2v) + 1)
__device__ __forceinline__ float head_offset_act(float v, float mag, float fl) {
  return fmaf(-2.f * mag, rcp_approx(ex2_approx(v * 2.885390081777927f) + 1.f), mag + fl);
}
// sigmoid(v) = 1 - 1 / (exp(v) + 1)
__device__ __forceinline__ float head_mask_act(float v) { return 1.f - rcp_approx(ex2_approx(v * 1.4426950408889634f) + 1.f); }

// Packed-quad loader shared by the conv kernels: 4 consecutive packed input channels (quad `vq`) of the
// channel-concatenated input at pixel (n, y, x); zero outside the image and in padding channels.
__device__ __forceinline__ float4 load_quad(const ConvParams& P, int vq, int n, int y, int x) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y < 0 || y >= P.h || x < 0 || x >= P.w) return r;
  int s = 0;
  if (P.nsrc > 1 && vq >= P.qstart[1]) s = 1;
  if (P.nsrc > 2 && vq >= P.qstart[2]) s = 2;
  if (vq >= P.qstart[P.nsrc]) return r;
  const int lq = vq - P.qstart[s];
  const float* base = P.src[s];
  const int cs = P.src_cstride[s], co = P.src_coffset[s];
  if (P.src_mode[s] == CRFP_SRC_UNSHUFFLE4) {
    const int hc4 = P.src_c[s] >> 6;  // quads per HR pixel (HR channels / 4)
    const int sub = lq / hc4, qq = lq - sub * hc4;
    const int dy = sub >> 2, dx = sub & 3;
    const size_t pix = ((size_t)n * (P.h * 4) + (y * 4 + dy)) * (size_t)(P.w * 4) + (x * 4 + dx);
    return __ldg(reinterpret_cast<const float4*>(base + pix * cs + co + qq * 4));
  }
  const size_t pix = ((size_t)n * P.h + y) * (size_t)P.w + x;
  const float* p = base + pix * cs + co + lq * 4;
  const int rem = P.src_c[s] - lq * 4;
  const bool vec_ok = ((cs | co) & 3) == 0;
  if (vec_ok && (rem >= 4 || co + lq * 4 + 4 <= cs)) {
    r = __ldg(reinterpret_cast<const float4*>(p));
    if (rem < 4) {
      if (rem < 3) r.z = 0.f;
      if (rem < 2) r.y = 0.f;
      r.w = 0.f;
    }
    return r;
  }
  r.x = __ldg(p);
  if (rem > 1) r.y = __ldg(p + 1);
  if (rem > 2) r.z = __ldg(p + 2);
  if (rem > 3) r.w = __ldg(p + 3);
  return r;
}

__device__ __forceinline__ float4 load_quad_fg(const ConvParams& P, int vq, int n, int y, int x) {
  float4 v = load_quad(P, vq, n, y, x);
  if (P.fg != nullptr && y >= 0 && y < P.h && x >= 0 && x < P.w) {
    const float f = __ldg(P.fg + (size_t)n * P.fg_clip_stride + (size_t)y * P.w + x);
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
  }
  return v;
}

// nn.Upsample / F.interpolate bilinear, align_corners=False: source index and lerp weight
__device__ __forceinline__ void bilin_src(int dst, float rscale, int size, int& i0, int& i1, float& l1) {
  float s = rscale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > size - 1) i0 = size - 1;
  i1 = i0 + ((i0 < size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

}  // namespace crfp
