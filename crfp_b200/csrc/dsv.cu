// CRFP_DSV composite entry points: the clip-level stage (flows + LR features) and one recurrent frame step.
// Reference: /root/reference/model/CRFP.py:1483-1508 (compute_flow), 797-814 (FNet), 1510-1686 (forward).
// Everything here is host-side sequencing of the library's own kernels on the caller's stream, inside the
// caller's workspace; no allocation, no synchronisation.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace crfp {

// ------------------------------------------------------------------------------------------------ layer table
enum Layer {
  L_FNET_E1_0, L_FNET_E1_2, L_FNET_E2_0, L_FNET_E2_2, L_FNET_E3_0, L_FNET_E3_2,
  L_FNET_D1_0, L_FNET_D1_2, L_FNET_D2_0, L_FNET_D2_2, L_FNET_D3_0, L_FNET_D3_2,
  L_FNET_F0, L_FNET_F2,
  L_ENC_LR_0, L_ENC_LR_2,
  L_UPSAMPLE, L_DOWNSAMPLE,
  L_DCN0_B0, L_DCN0_B2, L_DCN0_HEADS, L_DCN0_DCN,
  L_DCN1_B0, L_DCN1_B2, L_DCN1_FUSE, L_DCN1_HEADS, L_DCN1_DCN,
  L_DCN2_B0, L_DCN2_B2, L_DCN2_FUSE, L_DCN2_HEADS, L_DCN2_DCN,
  L_RES0_IN, L_RES0_IN_FIRST, L_RES0_C1, L_RES0_C2,
  L_RES1_IN, L_RES1_IN_FIRST, L_RES1_C1, L_RES1_C2,
  L_RES2_IN, L_RES2_IN_FIRST, L_RES2_C1, L_RES2_C2,
  L_UPSAMPLE_POST,
  L_DCN3_UP, L_DCN3_B0, L_DCN3_B2, L_DCN3_FUSE, L_DCN3_HEADS, L_DCN3_DCN,
  L_RES3_IN, L_RES3_IN_FIRST, L_RES3_C1, L_RES3_C2,
  L_ENC_HR_0, L_ENC_HR_2, L_TTTF, L_LAST,
  L_COUNT
};
static_assert(L_COUNT <= CRFP_DSV_MAX_LAYERS, "layer table too large");

#define CONV1(key, cin, cout) {key, nullptr, 0, 1, {cin, 0, 0}, {0, 0, 0}, cout, 0, 0, (cout) <= 4}
static const crfp_layer_info kLayers[L_COUNT] = {
    {"spynet.encoder1.0", nullptr, 0, 2, {3, 3, 0}, {0, 0, 0}, 32, 0, 0, 0},
    CONV1("spynet.encoder1.2", 32, 32),
    CONV1("spynet.encoder2.0", 32, 64), CONV1("spynet.encoder2.2", 64, 64),
    CONV1("spynet.encoder3.0", 64, 128), CONV1("spynet.encoder3.2", 128, 128),
    CONV1("spynet.decoder1.0", 128, 256), CONV1("spynet.decoder1.2", 256, 256),
    CONV1("spynet.decoder2.0", 256, 128), CONV1("spynet.decoder2.2", 128, 128),
    CONV1("spynet.decoder3.0", 128, 64), CONV1("spynet.decoder3.2", 64, 64),
    CONV1("spynet.flow.0", 64, 32), CONV1("spynet.flow.2", 32, 2),
    CONV1("encoder_lr.slice1.0", 3, 32), CONV1("encoder_lr.slice1.2", 32, 32),
    CONV1("upsample.upsample_conv", 32, 96),
    {"downsample.downsample_conv", nullptr, 0, 1, {64, 0, 0}, {CRFP_SRC_UNSHUFFLE4, 0, 0}, 32, 0, 0, 0},
#define DCN_L1(k, fuse)                                                                        \
    {"dcn_" #k ".dcn_block.0", nullptr, 0, 3, {32, 32, 2}, {0, 0, 0}, 32, 0, 0, 0},           \
    CONV1("dcn_" #k ".dcn_block.2", 32, 32),                                                   \
    fuse                                                                                       \
    {"dcn_" #k ".dcn_offset", "dcn_" #k ".dcn_mask", 2, 1, {32, 0, 0}, {0, 0, 0}, 216, 0, 8, 0}, \
    {"dcn_" #k ".dcn", nullptr, 1, 1, {32, 0, 0}, {0, 0, 0}, 32, 0, 8, 0},
#define FUSE_L1(k) {"dcn_" #k ".conv_fuse", nullptr, 0, 2, {32, 32, 0}, {0, 0, 0}, 32, 0, 0, 0},
    DCN_L1(0, )
    DCN_L1(1, FUSE_L1(1))
    DCN_L1(2, FUSE_L1(2))
#define RES_L1(k)                                                                                       \
    {"forward_resblocks_" #k ".main.0", nullptr, 0, 2, {32, 32, 0}, {0, 0, 0}, 32, 0, 0, 0},            \
    {"forward_resblocks_" #k ".main.0", nullptr, 0, 1, {24, 0, 0}, {0, 0, 0}, 32, 0, 0, 0},             \
    CONV1("forward_resblocks_" #k ".main.2.0.conv1", 32, 32),                                           \
    CONV1("forward_resblocks_" #k ".main.2.0.conv2", 32, 32),
    RES_L1(0) RES_L1(1) RES_L1(2)
    CONV1("upsample_post.upsample_conv", 24, 64),
    CONV1("dcn_3.upsample.upsample_conv", 32, 64),
    {"dcn_3.dcn_block.0", nullptr, 0, 3, {4, 4, 2}, {0, 0, 0}, 4, 0, 0, 1},
    CONV1("dcn_3.dcn_block.2", 4, 4),
    {"dcn_3.conv_fuse", nullptr, 0, 2, {4, 4, 0}, {0, 0, 0}, 4, 0, 0, 1},
    {"dcn_3.dcn_offset", "dcn_3.dcn_mask", 2, 1, {4, 0, 0}, {0, 0, 0}, 3, 0, 1, 1},
    {"dcn_3.dcn", nullptr, 1, 1, {4, 0, 0}, {0, 0, 0}, 4, 0, 1, 1},
    {"forward_resblocks_3.main.0", nullptr, 0, 2, {4, 4, 0}, {0, 0, 0}, 4, 0, 0, 1},
    {"forward_resblocks_3.main.0", nullptr, 0, 1, {4, 0, 0}, {0, 0, 0}, 4, 0, 0, 1},
    CONV1("forward_resblocks_3.main.2.0.conv1", 4, 4),
    CONV1("forward_resblocks_3.main.2.0.conv2", 4, 4),
    CONV1("encoder_hr.slice1.0", 6, 4), CONV1("encoder_hr.slice1.2", 4, 4),
    {"conv_tttf", nullptr, 0, 2, {4, 4, 0}, {0, 0, 0}, 4, 0, 0, 1},
    CONV1("conv_last", 4, 3),
};

// ---- per-variant layer tables: CRFP (v15) / CRFP_simple (v13) reuse DSV's modules with different widths
static crfp_layer_info g_tables[3][L_COUNT];
static bool g_tables_ready = false;
static int layer_tc_kind(int li);
static const crfp_layer_info* layer_table(int variant) {
  if (!g_tables_ready) {
    for (int v = 0; v < 3; ++v) {
      for (int i = 0; i < L_COUNT; ++i) { g_tables[v][i] = kLayers[i]; g_tables[v][i].tc = layer_tc_kind(i); }
      if (v == CRFP_VARIANT_DSV) continue;
      crfp_layer_info* T = g_tables[v];
      T[L_UPSAMPLE].cout = 128;                                       // PixelShufflePack(32 -> 32, x2)
      T[L_UPSAMPLE_POST].c[0] = 32;                                   // PixelShufflePack(32 -> 4, x4)
      static const int rin[3] = {L_RES0_IN, L_RES1_IN, L_RES2_IN}, rfi[3] = {L_RES0_IN_FIRST, L_RES1_IN_FIRST, L_RES2_IN_FIRST};
      for (int k = 0; k < 3; ++k) {
        T[rfi[k]].c[0] = 32;
        if (v == CRFP_VARIANT_V15) { T[rin[k]].nsrc = 3; T[rin[k]].c[2] = 32; T[rin[k]].tc = 0; }   // 96 input channels: SIMT
      }
      if (v == CRFP_VARIANT_V15) { T[L_RES3_IN].nsrc = 3; T[L_RES3_IN].c[2] = 4; }
    }
    g_tables_ready = true;
  }
  return g_tables[(variant >= 0 && variant < 3) ? variant : 0];
}

// tensor-core packing kind per layer: 0 none, 1 conv_tc (pack_conv_tc over the same source split), 2 dcn_tc
static int layer_tc_kind(int li) {
  switch (li) {
    case L_DCN0_B0: case L_DCN0_B2: case L_DCN0_HEADS:
    case L_DCN1_B0: case L_DCN1_B2: case L_DCN1_FUSE: case L_DCN1_HEADS:
    case L_DCN2_B0: case L_DCN2_B2: case L_DCN2_FUSE: case L_DCN2_HEADS:
    case L_RES0_IN: case L_RES0_IN_FIRST: case L_RES0_C1: case L_RES0_C2:
    case L_RES1_IN: case L_RES1_IN_FIRST: case L_RES1_C1: case L_RES1_C2:
    case L_RES2_IN: case L_RES2_IN_FIRST: case L_RES2_C1: case L_RES2_C2:
    case L_UPSAMPLE_POST: case L_DCN3_UP:
      return 1;
    case L_DCN0_DCN: case L_DCN1_DCN: case L_DCN2_DCN:
      return 2;
    case L_FNET_E1_2: case L_FNET_E2_0: case L_FNET_E2_2: case L_FNET_E3_0: case L_FNET_D3_2: case L_FNET_F0:
    case L_ENC_LR_2: case L_UPSAMPLE: case L_DOWNSAMPLE:
      return 3;
    // K-split: 128 / 256 input channels as 2 / 4 passes of 64 through the tensor-core conv; the later passes add the partial
    // sums of the earlier ones before the activation (Tc3Params.res_pre).  packing: [passes][ntiles][9][8][nt][8] hi / lo,
    // bias [2][ntiles*nt] = (bias, zeros)
    case L_FNET_E3_2: case L_FNET_D1_0: case L_FNET_D1_2: case L_FNET_D2_0: case L_FNET_D2_2: case L_FNET_D3_0:
      return 4;
    default:
      return 0;
  }
}

// ------------------------------------------------------------------------------------------------ conv builder
struct CB {
  ConvParams p;
  const crfp_dsv_weights* W_ = nullptr;
  int li_ = -1;
  CB(int n, int h, int w) {
    memset(&p, 0, sizeof(p));
    p.n = n; p.h = h; p.w = w;
    p.epi = EPI_STD; p.out_mode = CRFP_OUT_NHWC; p.post_scale = 1.f;
  }
  CB& src(const float* ptr, int c, int cs, int co = 0, int mode = CRFP_SRC_PLAIN) {
    const int s = p.nsrc++;
    p.src[s] = ptr; p.src_c[s] = c; p.src_cstride[s] = cs; p.src_coffset[s] = co; p.src_mode[s] = mode;
    return *this;
  }
  CB& layer(const crfp_dsv_weights* w, int li) {
    p.weight = w->layer[li].w; p.bias = w->layer[li].b; p.cout = layer_table(w->variant)[li].cout;
    W_ = w; li_ = li;
    return *this;
  }
  // TC3 precision: the same layer on the tensor cores (3 x bf16 split) when it is eligible
  bool tc3_eligible() const {
    if (!W_ || li_ < 0 || (W_->precision != CRFP_PREC_TC3 && W_->precision != CRFP_PREC_HALF) || !W_->layer_tc[li_].w_hi ||
        !W_->layer_tc[li_].w_lo)
      return false;
    if (p.epi != EPI_STD || p.out_bf16 || p.cout <= 4) return false;
    if (layer_tc_kind(li_) == 4)
      return ksplit_enabled() && p.nsrc == 1 && p.src_mode[0] == CRFP_SRC_PLAIN && p.src_c[0] > 64 && p.src_c[0] % 64 == 0 &&
             p.ndst == 1 && p.out_mode == CRFP_OUT_NHWC && p.residual == nullptr && p.act != CRFP_ACT_DCN_HEAD;
    int kc = 0;
    for (int s = 0; s < p.nsrc; ++s) {
      if (p.src_mode[s] == CRFP_SRC_UNSHUFFLE4) {
        if (p.src_c[s] != 64 || p.src_cstride[s] != 4 || p.src_coffset[s] != 0) return false;
      } else if (p.src_mode[s] != CRFP_SRC_PLAIN) {
        return false;
      }
      if (p.src_c[s] % 8 == 0) { kc += p.src_c[s] / 8; continue; }
      if (p.src_c[s] == 2 && s == p.nsrc - 1 && p.src_cstride[s] == 2 && p.src_coffset[s] == 0 && W_->layer_tc[li_].w_extra) continue;
      return false;
    }
    return kc >= 1 && kc <= 8;
  }
  static bool ksplit_enabled() {
    static const bool on = getenv("CRFP_NO_KSPLIT") == nullptr;   // A/B: the fp32 SIMT conv for the > 64-channel layers
    return on;
  }
  int run_tc3_ksplit(cudaStream_t st) const {
    int32_t nt = 0, ntiles = 0;
    CRFP_TRY(crfp_tc3_cout_tile(p.cout, 64, &nt, &ntiles));
    const size_t pass_elems = (size_t)ntiles * 9 * 8 * nt * 8;
    const int passes = p.src_c[0] / 64;
    for (int k = 0; k < passes; ++k) {
      Tc3Params t;
      memset(&t, 0, sizeof(t));
      t.n = p.n; t.h = p.h; t.w = p.w;
      t.nsrc = 1;
      t.src[0] = p.src[0]; t.src_c[0] = 64; t.src_cstride[0] = p.src_cstride[0]; t.src_coffset[0] = p.src_coffset[0] + 64 * k;
      t.src_mode[0] = CRFP_SRC_PLAIN;
      t.cout = p.cout;
      t.act = (k == passes - 1) ? p.act : CRFP_ACT_NONE;
      t.weight_hi = reinterpret_cast<const __nv_bfloat16*>(W_->layer_tc[li_].w_hi) + k * pass_elems;
      t.weight_lo = reinterpret_cast<const __nv_bfloat16*>(W_->layer_tc[li_].w_lo) + k * pass_elems;
      t.bias = W_->layer_tc[li_].b + (k == 0 ? 0 : (size_t)ntiles * nt);
      t.out_kind = TC_OUT_F32; t.ndst = 1;
      t.dst[0] = p.dst[0]; t.dst_c[0] = p.dst_c[0]; t.dst_cstride[0] = p.dst_cstride[0]; t.dst_coffset[0] = p.dst_coffset[0];
      if (k > 0) {   // in place: every thread reads the partial sum of exactly the elements it then overwrites
        t.residual = p.dst[0]; t.res_cstride = p.dst_cstride[0]; t.res_coffset = p.dst_coffset[0]; t.res_pre = 1;
      }
      t.post_scale = (k == passes - 1) ? p.post_scale : 1.f;
      CRFP_TRY(launch_conv_tc3(t, st));
    }
    return CRFP_OK;
  }
  int run_tc3(cudaStream_t st) const {
    if (layer_tc_kind(li_) == 4) return run_tc3_ksplit(st);
    Tc3Params t;
    memset(&t, 0, sizeof(t));
    t.n = p.n; t.h = p.h; t.w = p.w;
    for (int s = 0; s < p.nsrc; ++s) {
      if (p.src_c[s] % 8 == 0) {
        const int k = t.nsrc++;
        t.src[k] = p.src[s]; t.src_c[k] = p.src_c[s]; t.src_cstride[k] = p.src_cstride[s]; t.src_coffset[k] = p.src_coffset[s];
        t.src_mode[k] = p.src_mode[s];
      } else {
        t.extra = p.src[s];
        t.w_extra = W_->layer_tc[li_].w_extra;
      }
    }
    t.cout = p.cout; t.act = p.act;
    t.weight_hi = reinterpret_cast<const __nv_bfloat16*>(W_->layer_tc[li_].w_hi);
    t.weight_lo = reinterpret_cast<const __nv_bfloat16*>(W_->layer_tc[li_].w_lo);
    t.bias = W_->layer_tc[li_].b;
    t.fg = p.fg; t.fg_clip_stride = p.fg_clip_stride;
    t.out_kind = (p.out_mode == CRFP_OUT_SHUFFLE) ? TC_OUT_SHUFFLE_F32 : TC_OUT_F32;
    t.shuffle_r = p.shuffle_r; t.ndst = p.ndst;
    for (int s = 0; s < p.ndst; ++s) {
      t.dst[s] = p.dst[s]; t.dst_c[s] = p.dst_c[s]; t.dst_cstride[s] = p.dst_cstride[s]; t.dst_coffset[s] = p.dst_coffset[s];
    }
    t.residual = p.residual; t.res_cstride = p.res_cstride; t.res_coffset = p.res_coffset;
    t.flow = p.flow; t.head_split = p.head_split; t.head_mag = p.head_mag; t.post_scale = p.post_scale;
    // CRFP_PREC_HALF keeps the flow network at the full 3-product split (its weights stay bf16-packed): the flow is
    // 256 * tanh(.) and feeds every warp and every DCN offset
    t.half = (W_->precision == CRFP_PREC_HALF && li_ > L_FNET_F2) ? 1 : 0;
    return launch_conv_tc3(t, st);
  }
  CB& act(int a) { p.act = a; return *this; }
  CB& dst(float* ptr, int c, int cs, int co = 0) {
    const int s = p.ndst++;
    p.dst[s] = ptr; p.dst_c[s] = c; p.dst_cstride[s] = cs; p.dst_coffset[s] = co;
    return *this;
  }
  CB& shuffle(int r) { p.out_mode = CRFP_OUT_SHUFFLE; p.shuffle_r = r; return *this; }
  CB& res(const float* ptr, int cs, int co = 0) { p.residual = ptr; p.res_cstride = cs; p.res_coffset = co; return *this; }
  CB& scale(float s) { p.post_scale = s; return *this; }
  CB& head(const float* flow, int split, float mag) {
    p.act = CRFP_ACT_DCN_HEAD; p.flow = flow; p.head_split = split; p.head_mag = mag; return *this;
  }
  CB& fg(const float* ptr, long long clip_stride) { p.fg = ptr; p.fg_clip_stride = clip_stride; return *this; }
  int run(cudaStream_t st) {
    int q = 0;
    for (int s = 0; s < p.nsrc; ++s) { p.qstart[s] = q; q += (p.src_c[s] + 3) / 4; }
    for (int s = p.nsrc; s < 4; ++s) p.qstart[s] = q;
    p.cin_packed = (q * 4 + 7) & ~7;
    p.cout_packed = crfp_conv_cout_packed(p.cout);
    if (tc3_eligible()) return run_tc3(st);
    if (!p.weight || !p.bias) return CRFP_ERR_NULL;
    return launch_conv(p, st);
  }
};

// tensor-core conv builder (bf16 sources)
struct TB {
  TcParams p;
  TB(int n, int h, int w) {
    memset(&p, 0, sizeof(p));
    p.n = n; p.h = h; p.w = w;
    p.out_kind = TC_OUT_BF16; p.post_scale = 1.f;
  }
  TB& src(const void* ptr, int c, int cs, int co = 0) {
    const int s = p.nsrc++;
    p.src[s] = reinterpret_cast<const __nv_bfloat16*>(ptr); p.src_c[s] = c; p.src_cstride[s] = cs; p.src_coffset[s] = co;
    return *this;
  }
  TB& layer(const crfp_dsv_weights* w, int li) {
    p.weight = reinterpret_cast<const __nv_bfloat16*>(w->layer_tc[li].w_hi); p.bias = w->layer_tc[li].b;
    p.cout = layer_table(w->variant)[li].cout;
    return *this;
  }
  TB& act(int a) { p.act = a; return *this; }
  TB& dst(void* ptr, int c, int cs, int co = 0) {
    const int s = p.ndst++;
    p.dst[s] = ptr; p.dst_c[s] = c; p.dst_cstride[s] = cs; p.dst_coffset[s] = co;
    return *this;
  }
  TB& f32() { p.out_kind = TC_OUT_F32; return *this; }
  TB& shuffle_f32(int r) { p.out_kind = TC_OUT_SHUFFLE_F32; p.shuffle_r = r; return *this; }
  TB& res(const void* ptr, int cs, int co = 0) {
    p.residual = reinterpret_cast<const __nv_bfloat16*>(ptr); p.res_cstride = cs; p.res_coffset = co; return *this;
  }
  TB& scale(float s) { p.post_scale = s; return *this; }
  TB& head(const float* flow, int split, float mag) {
    p.act = CRFP_ACT_DCN_HEAD; p.flow = flow; p.head_split = split; p.head_mag = mag; return *this;
  }
  int run(cudaStream_t st) { return launch_conv_tc(p, st); }
};

// ------------------------------------------------------------------------------------------------ small kernels
// Fovea compositing (model/CRFP.py:1542-1547): hr_in[...,0:3] = fvs*m + up8(lr)*(1-m), hr_in[...,3:6] = up8(lr),
// hr_in[...,6:8] = 0.  fvs planar NCHW, mask uint8, lr NHWC4.
__global__ void __launch_bounds__(256) fovea_compose_kernel(int n, int H, int W, const float* __restrict__ fvs,
                                                            long long fvs_cs, const uint8_t* __restrict__ mks,
                                                            long long mks_cs, const float* __restrict__ lr4,
                                                            long long lr4_cs, float* __restrict__ out,
                                                            const uint8_t* __restrict__ flags, int tiles_x, int tiles_y) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const long long hw = (long long)H * W;
  const int b = (int)(pix / hw);
  const long long p = pix - (long long)b * hw;
  const int y = (int)(p / W), x = (int)(p - (long long)y * W);
  if (flags != nullptr && flags[((size_t)b * tiles_y + (y >> 5)) * tiles_x + (x >> 5)] == 0) return;
  const int hl = H >> 3, wl = W >> 3;
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_src(y, 0.125f, hl, y0, y1, ly);
  bilin_src(x, 0.125f, wl, x0, x1, lx);
  const float* lb = lr4 + (size_t)b * lr4_cs;
  const float4 p00 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y0 * wl + x0) * 4));
  const float4 p01 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y0 * wl + x1) * 4));
  const float4 p10 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y1 * wl + x0) * 4));
  const float4 p11 = __ldg(reinterpret_cast<const float4*>(lb + ((size_t)y1 * wl + x1) * 4));
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float b0 = hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
  const float b1 = hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
  const float b2 = hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
  const float m = mks[(size_t)b * mks_cs + p] ? 1.f : 0.f;
  const float* fb = fvs + (size_t)b * fvs_cs + p;
  const float f0 = __ldg(fb), f1 = __ldg(fb + hw), f2 = __ldg(fb + 2 * hw);
  float4* o = reinterpret_cast<float4*>(out + (size_t)pix * 8);
  o[0] = make_float4(f0 * m + b0 * (1.f - m), f1 * m + b1 * (1.f - m), f2 * m + b2 * (1.f - m), b0);
  o[1] = make_float4(b1, b2, 0.f, 0.f);
}

// flags[n][ty][tx] = any(mask) over the 32x32 tile dilated by one tile on every side (covers the 3-pixel receptive
// field of encoder_hr.0 -> encoder_hr.2 -> conv_tttf around every mask pixel).  Two tiny passes: per-tile any() with
// 8-byte loads (one block per tile row), then a 3x3 dilation over the tile grid.
__global__ void __launch_bounds__(256) fovea_tile_any_kernel(int H, int W, const uint8_t* __restrict__ mks, long long mks_cs,
                                                             int tiles_x, int tiles_y, uint8_t* __restrict__ any) {
  extern __shared__ int s_any[];
  pdl_trigger();
  pdl_wait();
  const int ty = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < tiles_x; i += 256) s_any[i] = 0;
  __syncthreads();
  const int y_lo = ty * 32, rows = min(32, H - y_lo);
  const uint8_t* mb = mks + (size_t)b * mks_cs + (size_t)y_lo * W;
  if ((W & 7) == 0 && (((uintptr_t)mb) & 7) == 0) {
    const int ppr = W >> 3;                    // 8-byte pieces per row
    for (int i = threadIdx.x; i < rows * ppr; i += 256) {
      const int r = i / ppr, p = i - r * ppr;
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(mb + (size_t)r * W) + p);
      if (v.x | v.y) s_any[p >> 2] = 1;
    }
  } else {
    for (int i = threadIdx.x; i < rows * W; i += 256) {
      const int r = i / W, x = i - r * W;
      if (mb[(size_t)r * W + x]) s_any[x >> 5] = 1;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tiles_x; i += 256) any[((size_t)b * tiles_y + ty) * tiles_x + i] = s_any[i] ? 1 : 0;
}

__global__ void __launch_bounds__(256) fovea_tile_dilate_kernel(int tiles_x, int tiles_y, int n, const uint8_t* __restrict__ any,
                                                                uint8_t* __restrict__ flags) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= tiles_x * tiles_y * n) return;
  const int tx = i % tiles_x, ty = (i / tiles_x) % tiles_y, b = i / (tiles_x * tiles_y);
  int f = 0;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = ty + dy, xx = tx + dx;
      if (yy >= 0 && yy < tiles_y && xx >= 0 && xx < tiles_x) f |= any[((size_t)b * tiles_y + yy) * tiles_x + xx];
    }
  flags[i] = f ? 1 : 0;
}

static int launch_compose(int n, int H, int W, const crfp_dsv_frame_desc* d, float* out, const uint8_t* flags, int tiles_x,
                          int tiles_y, cudaStream_t st) {
  const long long total = (long long)n * H * W;
  launch_k(fovea_compose_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st, n, H, W, d->fvs, d->fvs_clip_stride, d->mks,
                                                                       d->mks_clip_stride, d->lr4, d->lr4_clip_stride,
                                                                       out, flags, tiles_x, tiles_y);
  return check_launch();
}

// gather `n` strided images into a contiguous run (used to make per-frame slices of clip tensors dense)
__global__ void __launch_bounds__(256) gather_images_kernel(int n, long long per_image, const float* __restrict__ in,
                                                            long long in_stride, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)n * per_image;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int b = (int)(idx / per_image);
  out[idx] = __ldg(in + (size_t)b * in_stride + (idx - (long long)b * per_image));
}

static int gather_images(int n, long long per_image, const float* in, long long in_stride, float* out, cudaStream_t st) {
  const long long total = (long long)n * per_image;
  launch_k(gather_images_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st, n, per_image, in, in_stride, out);
  return check_launch();
}

// ------------------------------------------------------------------------------------------------ workspace carving
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* b) : base((char*)b), off(0) {}
  float* take(size_t nfloats) {
    float* p = base ? (float*)(base + off) : nullptr;
    off += align_up(nfloats * sizeof(float), 256);
    return p;
  }
};

struct FNetWs {
  float *a, *b, *flowtmp, *prev4;
};
static const int kPrepChunk = 8;  // images per FNet / encoder chunk

static size_t carve_prepare(const crfp_dsv_shape* s, void* ws, FNetWs* f, float** enc_tmp) {
  Carver c(ws);
  const size_t hw = (size_t)s->h * s->w;
  f->a = c.take(kPrepChunk * hw * 64);
  f->b = c.take(kPrepChunk * hw * 64);
  f->flowtmp = c.take(kPrepChunk * hw * 2);
  f->prev4 = c.take((size_t)s->n * hw * 4);
  *enc_tmp = c.take(kPrepChunk * hw * 32);
  return c.off;
}

// FNet over `cnt` consecutive (cur, prev) NHWC4 image pairs -> flows (cnt, h, w, 2)   (model/CRFP.py:797-814)
static int run_fnet(const crfp_dsv_weights* W, int cnt, int h, int w, const float* cur4, const float* prev4,
                    const FNetWs& f, float* flows, cudaStream_t st) {
  const int h2 = h / 2, w2 = w / 2, h4 = h2 / 2, w4 = w2 / 2, h8 = h4 / 2, w8 = w4 / 2;
  if (h8 < 1 || w8 < 1) return CRFP_ERR_BAD_SHAPE;
  float *A = f.a, *B = f.b;
  CRFP_TRY(CB(cnt, h, w).src(cur4, 3, 4).src(prev4, 3, 4).layer(W, L_FNET_E1_0).act(CRFP_ACT_RELU).dst(A, 32, 32).run(st));
  CRFP_TRY(CB(cnt, h, w).src(A, 32, 32).layer(W, L_FNET_E1_2).act(CRFP_ACT_RELU).dst(B, 32, 32).run(st));
  CRFP_TRY(crfp_avgpool2(cnt, h, w, 32, B, A, st));
  CRFP_TRY(CB(cnt, h2, w2).src(A, 32, 32).layer(W, L_FNET_E2_0).act(CRFP_ACT_RELU).dst(B, 64, 64).run(st));
  CRFP_TRY(CB(cnt, h2, w2).src(B, 64, 64).layer(W, L_FNET_E2_2).act(CRFP_ACT_RELU).dst(A, 64, 64).run(st));
  CRFP_TRY(crfp_avgpool2(cnt, h2, w2, 64, A, B, st));
  CRFP_TRY(CB(cnt, h4, w4).src(B, 64, 64).layer(W, L_FNET_E3_0).act(CRFP_ACT_RELU).dst(A, 128, 128).run(st));
  CRFP_TRY(CB(cnt, h4, w4).src(A, 128, 128).layer(W, L_FNET_E3_2).act(CRFP_ACT_RELU).dst(B, 128, 128).run(st));
  CRFP_TRY(crfp_avgpool2(cnt, h4, w4, 128, B, A, st));
  CRFP_TRY(CB(cnt, h8, w8).src(A, 128, 128).layer(W, L_FNET_D1_0).act(CRFP_ACT_RELU).dst(B, 256, 256).run(st));
  CRFP_TRY(CB(cnt, h8, w8).src(B, 256, 256).layer(W, L_FNET_D1_2).act(CRFP_ACT_RELU).dst(A, 256, 256).run(st));
  CRFP_TRY(crfp_resize_bilinear(cnt, h8, w8, 256, A, 2 * h8, 2 * w8, 0.5f, 0.5f, 1.f, B, st));
  CRFP_TRY(CB(cnt, 2 * h8, 2 * w8).src(B, 256, 256).layer(W, L_FNET_D2_0).act(CRFP_ACT_RELU).dst(A, 128, 128).run(st));
  CRFP_TRY(CB(cnt, 2 * h8, 2 * w8).src(A, 128, 128).layer(W, L_FNET_D2_2).act(CRFP_ACT_RELU).dst(B, 128, 128).run(st));
  CRFP_TRY(crfp_resize_bilinear(cnt, 2 * h8, 2 * w8, 128, B, 4 * h8, 4 * w8, 0.5f, 0.5f, 1.f, A, st));
  CRFP_TRY(CB(cnt, 4 * h8, 4 * w8).src(A, 128, 128).layer(W, L_FNET_D3_0).act(CRFP_ACT_RELU).dst(B, 64, 64).run(st));
  CRFP_TRY(CB(cnt, 4 * h8, 4 * w8).src(B, 64, 64).layer(W, L_FNET_D3_2).act(CRFP_ACT_RELU).dst(A, 64, 64).run(st));
  CRFP_TRY(crfp_resize_bilinear(cnt, 4 * h8, 4 * w8, 64, A, 8 * h8, 8 * w8, 0.5f, 0.5f, 1.f, B, st));
  const int hf = 8 * h8, wf = 8 * w8;
  CRFP_TRY(CB(cnt, hf, wf).src(B, 64, 64).layer(W, L_FNET_F0).act(CRFP_ACT_RELU).dst(A, 32, 32).run(st));
  const bool same = (hf == h && wf == w);
  float* fl = same ? flows : f.flowtmp;
  CRFP_TRY(CB(cnt, hf, wf).src(A, 32, 32).layer(W, L_FNET_F2).act(CRFP_ACT_TANH256).dst(fl, 2, 2).run(st));
  if (!same)
    CRFP_TRY(crfp_resize_bilinear(cnt, hf, wf, 2, fl, h, w, (float)hf / (float)h, (float)wf / (float)w, 1.f, flows, st));
  return CRFP_OK;
}

struct FrameWs {
  // L1 (2h x 2w)
  float *P, *P_w, *cur[3], *t1, *t2, *offf[2], *A, *r0, *r1, *prop, *flow_l1, *flow8, *om, *fg_l1;
  // HR (8h x 8w)
  float *S0_w, *q, *po, *h1, *h2, *h3, *om3, *A3, *g0, *g1, *S_pre, *hr_in, *e1, *x_hr, *flow_hr, *tile_flags;
  // dense per-frame copies of strided clip slices
  float *x_lr_d, *flow_d;
};

static size_t carve_frame(const crfp_dsv_shape* s, void* ws, FrameWs* f) {
  Carver c(ws);
  const size_t n = s->n, hw = (size_t)s->h * s->w;
  const size_t l1 = n * hw * 4, hr = n * hw * 64;
  f->P = c.take(l1 * 32); f->P_w = c.take(l1 * 32);
  for (int k = 0; k < 3; ++k) f->cur[k] = c.take(l1 * 32);
  f->t1 = c.take(l1 * 32); f->t2 = c.take(l1 * 32);
  f->offf[0] = c.take(l1 * 32); f->offf[1] = c.take(l1 * 32);
  f->A = c.take(l1 * 32); f->r0 = c.take(l1 * 32); f->r1 = c.take(l1 * 32);
  f->prop = c.take(l1 * 24);
  f->flow_l1 = c.take(l1 * 2);
  f->flow8 = c.take(l1 * 4);
  f->om = c.take(l1 * 216);
  f->fg_l1 = c.take(l1);
  f->S0_w = c.take(hr * 4); f->q = c.take(hr * 4); f->po = c.take(hr * 4);
  f->h1 = c.take(hr * 4); f->h2 = c.take(hr * 4); f->h3 = c.take(hr * 4);
  f->om3 = c.take(hr * 4); f->A3 = c.take(hr * 4);
  f->g0 = c.take(hr * 4); f->g1 = c.take(hr * 4); f->S_pre = c.take(hr * 4);
  f->hr_in = c.take(hr * 8); f->e1 = c.take(hr * 4); f->x_hr = c.take(hr * 4);
  f->flow_hr = c.take(hr * 2);
  f->tile_flags = c.take(2 * (n * (size_t)((s->h * 8 + 31) / 32) * ((s->w * 8 + 31) / 32) / 4 + 64));   // flags + raw any()
  f->x_lr_d = c.take(n * hw * 32);
  f->flow_d = c.take(n * hw * 2);
  return c.off;
}


// Optional fork / join of the off-chain work onto a second stream (crfp_dsv_frame_desc.aux_stream)
struct Aux {
  cudaStream_t main, aux;
  cudaEvent_t ev[3];
  bool on;
  // everything enqueued on `to` after this call runs after everything enqueued on `from` before it
  int order(cudaStream_t from, cudaStream_t to, int e) const {
    if (!on) return CRFP_OK;
    cudaError_t r = cudaEventRecord(ev[e], from);
    if (r == cudaSuccess) r = cudaStreamWaitEvent(to, ev[e], 0);
    if (r != cudaSuccess) { note_cuda_error(r); return CRFP_ERR_CUDA; }
    return CRFP_OK;
  }
};

// Offset / mask heads + DCNv2 of one L1 level (DCN_module.forward, model/CRFP.py:337-350): ONE 216-channel conv
// into `om`, then the align kernel.  Tensor-core precision: the conv stores the RAW head outputs and the align kernel's
// sampler applies 10*tanh + flow / sigmoid itself (crfp_dcn_desc.head_raw) — the conv epilogue drops to bias + store
// (it was MUFU / issue bound with the activations in it); CRFP_HEAD_EPI=1 restores the epilogue form for A/B.
static int run_heads_dcn(const crfp_dsv_weights* W, int n, int h1, int w1, const float* z, const float* flow_l1, float* om,
                         const float* P, float* A, int l_heads, int l_dcn, cudaStream_t st) {
  static const bool head_epi = (getenv("CRFP_HEAD_EPI") != nullptr);
  static const bool no_fused = (getenv("CRFP_ALIGN_UNFUSED") != nullptr);   // A/B switch: heads conv + align kernel
  const bool half = W->precision == CRFP_PREC_HALF;
  const bool tc = (W->precision == CRFP_PREC_TC3 || half) && W->layer_tc[l_dcn].w_hi && W->layer_tc[l_dcn].w_lo;
  if (half && !(W->layer_tc[l_heads].w_fused && W->layer_tc[l_heads].b_fused)) return CRFP_ERR_NULL;   // half: fused kernel only
  if (tc && ((!no_fused && !head_epi) || half) && W->layer_tc[l_heads].w_fused && W->layer_tc[l_heads].b_fused) {
    // ONE kernel: the offset / mask tensor stays in TMEM (dcn_fused.cu)
    crfp_align_fused_desc fd;
    memset(&fd, 0, sizeof(fd));
    fd.n = n; fd.h = h1; fd.w = w1;
    fd.z = z; fd.z_cstride = 32;
    fd.flow = flow_l1;
    fd.x = P; fd.x_cstride = 32;
    fd.heads_w = W->layer_tc[l_heads].w_fused; fd.heads_b = W->layer_tc[l_heads].b_fused;
    fd.dcn_w_hi = W->layer_tc[l_dcn].w_hi; fd.dcn_w_lo = W->layer_tc[l_dcn].w_lo; fd.dcn_b = W->layer_tc[l_dcn].b;
    fd.out = A; fd.out_cstride = 32;
    fd.head_mag = 10.f;
    fd.half = half ? 1 : 0;
    return launch_align_fused(fd, st);
  }
  const bool raw = tc && !head_epi;
  if (raw)
    CRFP_TRY(CB(n, h1, w1).src(z, 32, 32).layer(W, l_heads).dst(om, 216, 216).run(st));
  else
    CRFP_TRY(CB(n, h1, w1).src(z, 32, 32).layer(W, l_heads).head(flow_l1, 144, 10.f).dst(om, 216, 216).run(st));
  crfp_dcn_desc dd;
  memset(&dd, 0, sizeof(dd));
  dd.n = n; dd.h = h1; dd.w = w1; dd.c = 32; dd.cout = 32; dd.dg = 8;
  dd.x = P; dd.x_cstride = 32;
  dd.offset = om; dd.off_cstride = 216; dd.off_coffset = 0;
  dd.mask = om; dd.mask_cstride = 216; dd.mask_coffset = 144;
  dd.weight = W->layer[l_dcn].w; dd.bias = W->layer[l_dcn].b;
  dd.out = A; dd.out_cstride = 32;
  if (tc) {
    dd.weight = reinterpret_cast<const float*>(W->layer_tc[l_dcn].w_hi); dd.bias = W->layer_tc[l_dcn].b;
    if (raw) { dd.head_raw = 1; dd.head_flow = flow_l1; dd.head_mag = 10.f; }
    return launch_dcn_tc3(dd, W->layer_tc[l_dcn].w_lo, flow_l1, st);
  }
  return launch_dcn(dd, st);
}

// L1 stage of CRFP (v15, model/CRFP.py:1260-1340) and CRFP_simple (v13, model/CRFP.py:968-1050): no DSV split; the HR
// state is warped first and both the state and its warp go through `downsample`; v15 feeds the warped planes into
// every residual block as a third concat source.  Same kernels, same hand-off (q, po, S0_w, flow_hr) to the HR stage.
static int frame_l1_v1x(const crfp_dsv_frame_desc* d, const crfp_dsv_weights* W, const FrameWs& f, const float* x_lr,
                        cudaStream_t st) {
  const crfp_dsv_shape* s = &d->shape;
  const int n = s->n, h = s->h, w = s->w;
  const int h1 = 2 * h, w1 = 2 * w, H = 8 * h, Wd = 8 * w;
  const size_t hw = (size_t)h * w;
  const bool three = (W->variant == CRFP_VARIANT_V15);
  static const int lr1[3] = {L_RES0_C1, L_RES1_C1, L_RES2_C1};
  static const int lr2[3] = {L_RES0_C2, L_RES1_C2, L_RES2_C2};
  float* cur = f.cur[0];
  CRFP_TRY(CB(n, h, w).src(x_lr, 32, 32).layer(W, L_UPSAMPLE).shuffle(2).dst(cur, 32, 32).run(st));
  if (!d->first) {
    const float* flow = d->flow;
    if (n > 1 && d->flow_clip_stride != (long long)(hw * 2)) {
      CRFP_TRY(gather_images(n, (long long)hw * 2, d->flow, d->flow_clip_stride, f.flow_d, st));
      flow = f.flow_d;
    }
    CRFP_TRY(crfp_resize_bilinear(n, h, w, 2, flow, h1, w1, 0.5f, 0.5f, 2.f, f.flow_l1, st));
    CRFP_TRY(crfp_resize_bilinear(n, h, w, 2, flow, H, Wd, 0.125f, 0.125f, 8.f, f.flow_hr, st));
    crfp_warp_desc wd;
    memset(&wd, 0, sizeof(wd));
    wd.n = n; wd.h = H; wd.w = Wd; wd.c = 4; wd.x = d->state_hr; wd.x_cstride = 4; wd.flow = f.flow_hr;
    wd.out = f.S0_w; wd.out_cstride = 4;
    CRFP_TRY(launch_flow_warp(wd, st));
    CRFP_TRY(CB(n, h1, w1).src(f.S0_w, 64, 4, 0, CRFP_SRC_UNSHUFFLE4).layer(W, L_DOWNSAMPLE).dst(f.P_w, 32, 32).run(st));
    CRFP_TRY(CB(n, h1, w1).src(d->state_hr, 64, 4, 0, CRFP_SRC_UNSHUFFLE4).layer(W, L_DOWNSAMPLE).dst(f.P, 32, 32).run(st));
    static const int lb0[3] = {L_DCN0_B0, L_DCN1_B0, L_DCN2_B0};
    static const int lb2[3] = {L_DCN0_B2, L_DCN1_B2, L_DCN2_B2};
    static const int lfu[3] = {-1, L_DCN1_FUSE, L_DCN2_FUSE};
    static const int lhd[3] = {L_DCN0_HEADS, L_DCN1_HEADS, L_DCN2_HEADS};
    static const int ldc[3] = {L_DCN0_DCN, L_DCN1_DCN, L_DCN2_DCN};
    static const int lri[3] = {L_RES0_IN, L_RES1_IN, L_RES2_IN};
    const float* offfeat = nullptr;
    for (int k = 0; k < 3; ++k) {
      CRFP_TRY(CB(n, h1, w1).src(cur, 32, 32).src(f.P_w, 32, 32).src(f.flow_l1, 2, 2).layer(W, lb0[k])
                   .act(CRFP_ACT_LRELU).dst(f.t1, 32, 32).run(st));
      float* z = f.offf[k & 1];
      if (k == 0) {
        CRFP_TRY(CB(n, h1, w1).src(f.t1, 32, 32).layer(W, lb2[k]).act(CRFP_ACT_LRELU).dst(z, 32, 32).run(st));
      } else {
        CRFP_TRY(CB(n, h1, w1).src(f.t1, 32, 32).layer(W, lb2[k]).act(CRFP_ACT_LRELU).dst(f.t2, 32, 32).run(st));
        CRFP_TRY(CB(n, h1, w1).src(f.t2, 32, 32).src(offfeat, 32, 32).layer(W, lfu[k]).act(CRFP_ACT_LRELU)
                     .dst(z, 32, 32).run(st));
      }
      offfeat = z;
      CRFP_TRY(run_heads_dcn(W, n, h1, w1, z, f.flow_l1, f.om, f.P, f.A, lhd[k], ldc[k], st));
      CB in(n, h1, w1);
      in.src(cur, 32, 32).src(f.A, 32, 32);
      if (three) in.src(f.P_w, 32, 32);
      in.layer(W, lri[k]).act(CRFP_ACT_LRELU).dst(f.r0, 32, 32);
      CRFP_TRY(in.run(st));
      CRFP_TRY(CB(n, h1, w1).src(f.r0, 32, 32).layer(W, lr1[k]).act(CRFP_ACT_RELU).dst(f.r1, 32, 32).run(st));
      float* nxt = f.cur[(k + 1) % 3];
      CRFP_TRY(CB(n, h1, w1).src(f.r1, 32, 32).layer(W, lr2[k]).res(f.r0, 32).dst(nxt, 32, 32).run(st));
      cur = nxt;
    }
    CRFP_TRY(CB(n, h1, w1).src(cur, 32, 32).layer(W, L_UPSAMPLE_POST).act(CRFP_ACT_LRELU).shuffle(4).dst(f.q, 4, 4).run(st));
    CRFP_TRY(CB(n, h1, w1).src(offfeat, 32, 32).layer(W, L_DCN3_UP).shuffle(4).scale(2.f).dst(f.po, 4, 4).run(st));
  } else {
    static const int lrf[3] = {L_RES0_IN_FIRST, L_RES1_IN_FIRST, L_RES2_IN_FIRST};
    for (int k = 0; k < 3; ++k) {   // cat([cur, zeros, ...]): only weight[:, :32] contributes   (CRFP.py:1342-1356)
      CRFP_TRY(CB(n, h1, w1).src(cur, 32, 32).layer(W, lrf[k]).act(CRFP_ACT_LRELU).dst(f.r0, 32, 32).run(st));
      CRFP_TRY(CB(n, h1, w1).src(f.r0, 32, 32).layer(W, lr1[k]).act(CRFP_ACT_RELU).dst(f.r1, 32, 32).run(st));
      float* nxt = f.cur[(k + 1) % 3];
      CRFP_TRY(CB(n, h1, w1).src(f.r1, 32, 32).layer(W, lr2[k]).res(f.r0, 32).dst(nxt, 32, 32).run(st));
      cur = nxt;
    }
    CRFP_TRY(CB(n, h1, w1).src(cur, 32, 32).layer(W, L_UPSAMPLE_POST).act(CRFP_ACT_LRELU).shuffle(4).dst(f.q, 4, 4).run(st));
  }
  return CRFP_OK;
}

// L1 stage in bf16 mode: every 2h x 2w layer on tcgen05 (conv_tc / dcn_tc), bf16 storage, fp32 accumulation.
// Produces the same hand-off to the HR stage as the fp32 path: q, po (fp32 HR), S0_w, flow_hr, new state_l1 (bf16).
static int frame_l1_bf16(const crfp_dsv_frame_desc* d, const crfp_dsv_weights* W, const FrameWs& f, const float* x_lr,
                         cudaStream_t st) {
  const crfp_dsv_shape* s = &d->shape;
  const int n = s->n, h = s->h, w = s->w;
  const int h1 = 2 * h, w1 = 2 * w, H = 8 * h, Wd = 8 * w;
  const size_t hw = (size_t)h * w;
  for (int li = 0; li < L_COUNT; ++li)
    if ((layer_tc_kind(li) == 1 || layer_tc_kind(li) == 2) && (!W->layer_tc[li].w_hi || !W->layer_tc[li].b)) return CRFP_ERR_NULL;
  // feat_prop_lv0 = PixelShufflePack(x_lr) (fp32 SIMT conv at LR resolution, bf16 store)   (CRFP.py:1560)
  {
    CB c(n, h, w);
    c.src(x_lr, 32, 32).layer(W, L_UPSAMPLE).shuffle(2).dst(d->first ? f.prop : f.cur[0], 24, d->first ? 24 : 32);
    c.p.out_bf16 = 1;
    CRFP_TRY(c.run(st));
  }
  if (!d->first) {
    const float* flow = d->flow;
    if (n > 1 && d->flow_clip_stride != (long long)(hw * 2)) {
      CRFP_TRY(gather_images(n, (long long)hw * 2, d->flow, d->flow_clip_stride, f.flow_d, st));
      flow = f.flow_d;
    }
    CRFP_TRY(launch_flow_up2_dual(n, h, w, flow, f.flow_l1, f.flow8, st));
    CRFP_TRY(crfp_resize_bilinear(n, h, w, 2, flow, H, Wd, 0.125f, 0.125f, 8.f, f.flow_hr, st));
    {  // P = downsample(S0): fp32 HR state read through pixel_unshuffle(4), bf16 store         (CRFP.py:1569)
      CB c(n, h1, w1);
      c.src(d->state_hr, 64, 4, 0, CRFP_SRC_UNSHUFFLE4).layer(W, L_DOWNSAMPLE).dst(f.P, 32, 32);
      c.p.out_bf16 = 1;
      CRFP_TRY(c.run(st));
    }
    crfp_warp_desc wd;
    memset(&wd, 0, sizeof(wd));
    wd.n = n; wd.h = h1; wd.w = w1; wd.c = 32; wd.x = f.P; wd.x_cstride = 32; wd.flow = f.flow_l1;
    wd.out = f.P_w; wd.out_cstride = 32;
    CRFP_TRY(launch_flow_warp_bf16(wd, st));
    wd.h = H; wd.w = Wd; wd.c = 4; wd.x = d->state_hr; wd.x_cstride = 4; wd.flow = f.flow_hr;
    wd.out = f.S0_w; wd.out_cstride = 4;
    CRFP_TRY(launch_flow_warp(wd, st));
    for (int k = 0; k < 3; ++k) {
      wd.h = h1; wd.w = w1; wd.c = 8; wd.x = d->state_l1; wd.x_cstride = 24; wd.x_coffset = 8 * k; wd.flow = f.flow_l1;
      wd.out = f.cur[k]; wd.out_cstride = 32; wd.out_coffset = 24;
      CRFP_TRY(launch_flow_warp_bf16(wd, st));
    }
    static const int lb0[3] = {L_DCN0_B0, L_DCN1_B0, L_DCN2_B0};
    static const int lb2[3] = {L_DCN0_B2, L_DCN1_B2, L_DCN2_B2};
    static const int lfu[3] = {-1, L_DCN1_FUSE, L_DCN2_FUSE};
    static const int lhd[3] = {L_DCN0_HEADS, L_DCN1_HEADS, L_DCN2_HEADS};
    static const int ldc[3] = {L_DCN0_DCN, L_DCN1_DCN, L_DCN2_DCN};
    static const int lri[3] = {L_RES0_IN, L_RES1_IN, L_RES2_IN};
    static const int lr1[3] = {L_RES0_C1, L_RES1_C1, L_RES2_C1};
    static const int lr2[3] = {L_RES0_C2, L_RES1_C2, L_RES2_C2};
    const float* offfeat = nullptr;
    for (int k = 0; k < 3; ++k) {
      float* cur = f.cur[k];
      CRFP_TRY(TB(n, h1, w1).src(cur, 32, 32).src(f.P_w, 32, 32).src(f.flow8, 8, 8).layer(W, lb0[k]).act(CRFP_ACT_LRELU)
                   .dst(f.t1, 32, 32).run(st));
      float* z = f.offf[k & 1];
      if (k == 0) {
        CRFP_TRY(TB(n, h1, w1).src(f.t1, 32, 32).layer(W, lb2[k]).act(CRFP_ACT_LRELU).dst(z, 32, 32).run(st));
      } else {
        CRFP_TRY(TB(n, h1, w1).src(f.t1, 32, 32).layer(W, lb2[k]).act(CRFP_ACT_LRELU).dst(f.t2, 32, 32).run(st));
        CRFP_TRY(TB(n, h1, w1).src(f.t2, 32, 32).src(offfeat, 32, 32).layer(W, lfu[k]).act(CRFP_ACT_LRELU)
                     .dst(z, 32, 32).run(st));
      }
      offfeat = z;
      CRFP_TRY(TB(n, h1, w1).src(z, 32, 32).layer(W, lhd[k]).head(f.flow_l1, 144, 10.f).f32().dst(f.om, 216, 216).run(st));
      crfp_dcn_desc dd;
      memset(&dd, 0, sizeof(dd));
      dd.n = n; dd.h = h1; dd.w = w1; dd.c = 32; dd.cout = 32; dd.dg = 8;
      dd.x = f.P; dd.x_cstride = 32;
      dd.offset = f.om; dd.off_cstride = 216; dd.off_coffset = 0;
      dd.mask = f.om; dd.mask_cstride = 216; dd.mask_coffset = 144;
      dd.weight = reinterpret_cast<const float*>(W->layer_tc[ldc[k]].w_hi); dd.bias = W->layer_tc[ldc[k]].b;
      dd.out = f.A; dd.out_cstride = 32;
      CRFP_TRY(launch_dcn_tc(dd, st));
      CRFP_TRY(TB(n, h1, w1).src(cur, 32, 32).src(f.A, 32, 32).layer(W, lri[k]).act(CRFP_ACT_LRELU).dst(f.r0, 32, 32).run(st));
      CRFP_TRY(TB(n, h1, w1).src(f.r0, 32, 32).layer(W, lr1[k]).act(CRFP_ACT_RELU).dst(f.r1, 32, 32).run(st));
      float* nxt = (k < 2) ? f.cur[k + 1] : f.prop;
      const int nxt_cs = (k < 2) ? 32 : 24;
      CRFP_TRY(TB(n, h1, w1).src(f.r1, 32, 32).layer(W, lr2[k]).res(f.r0, 32).dst(nxt, 24, nxt_cs)
                   .dst(d->state_l1, 8, 24, 8 * k).run(st));
    }
    CRFP_TRY(TB(n, h1, w1).src(f.prop, 24, 24).layer(W, L_UPSAMPLE_POST).act(CRFP_ACT_LRELU).shuffle_f32(4)
                 .dst(f.q, 4, 4).run(st));
    CRFP_TRY(TB(n, h1, w1).src(offfeat, 32, 32).layer(W, L_DCN3_UP).shuffle_f32(4).scale(2.f).dst(f.po, 4, 4).run(st));
  } else {
    static const int lrf[3] = {L_RES0_IN_FIRST, L_RES1_IN_FIRST, L_RES2_IN_FIRST};
    static const int lr1[3] = {L_RES0_C1, L_RES1_C1, L_RES2_C1};
    static const int lr2[3] = {L_RES0_C2, L_RES1_C2, L_RES2_C2};
    float* pa = f.prop;
    float* pb = f.cur[0];
    for (int k = 0; k < 3; ++k) {
      CRFP_TRY(TB(n, h1, w1).src(pa, 24, 24).layer(W, lrf[k]).act(CRFP_ACT_LRELU).dst(f.r0, 32, 32).run(st));
      CRFP_TRY(TB(n, h1, w1).src(f.r0, 32, 32).layer(W, lr1[k]).act(CRFP_ACT_RELU).dst(f.r1, 32, 32).run(st));
      CRFP_TRY(TB(n, h1, w1).src(f.r1, 32, 32).layer(W, lr2[k]).res(f.r0, 32).dst(pb, 24, 24)
                   .dst(d->state_l1, 8, 24, 8 * k).run(st));
      float* tmp = pa; pa = pb; pb = tmp;
    }
    CRFP_TRY(TB(n, h1, w1).src(pa, 24, 24).layer(W, L_UPSAMPLE_POST).act(CRFP_ACT_LRELU).shuffle_f32(4)
                 .dst(f.q, 4, 4).run(st));
  }
  return CRFP_OK;
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_dsv_num_layers(void) { return L_COUNT; }

extern "C" int crfp_dsv_layer_info(int i, crfp_layer_info* info) {
  if (!info) return CRFP_ERR_NULL;
  if (i < 0 || i >= L_COUNT) return CRFP_ERR_BAD_SHAPE;
  *info = layer_table(CRFP_VARIANT_DSV)[i];
  return CRFP_OK;
}

extern "C" int crfp_layer_info_variant(int variant, int i, crfp_layer_info* info) {
  if (!info) return CRFP_ERR_NULL;
  if (i < 0 || i >= L_COUNT || variant < 0 || variant > 2) return CRFP_ERR_BAD_SHAPE;
  *info = layer_table(variant)[i];
  return CRFP_OK;
}

extern "C" size_t crfp_sizeof_dsv_weights(void) { return sizeof(crfp_dsv_weights); }
extern "C" size_t crfp_sizeof_dsv_frame_desc(void) { return sizeof(crfp_dsv_frame_desc); }

static int check_shape(const crfp_dsv_shape* s) {
  if (!s) return CRFP_ERR_NULL;
  if (s->n <= 0 || s->t <= 0 || s->h < 8 || s->w < 8) return CRFP_ERR_BAD_SHAPE;
  if (s->mid_channels != 32) return CRFP_ERR_UNSUPPORTED;
  return CRFP_OK;
}

extern "C" size_t crfp_dsv_prepare_workspace(const crfp_dsv_shape* s) {
  if (check_shape(s) != CRFP_OK) return 0;
  FNetWs f; float* e;
  return carve_prepare(s, nullptr, &f, &e);
}

extern "C" int crfp_dsv_prepare(const crfp_dsv_shape* s, const crfp_dsv_weights* W, const float* lrs,
                                const float* prev_lr, float* lr4, float* x_lr, float* flows, void* workspace,
                                size_t ws_bytes, crfp_stream stream) {
  CRFP_TRY(check_shape(s));
  if (!W || !lrs || !lr4 || !x_lr || !flows || !workspace) return CRFP_ERR_NULL;
  if (W->nlayers != L_COUNT || W->mid_channels != 32) return CRFP_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  FNetWs f; float* enc_tmp;
  if (carve_prepare(s, workspace, &f, &enc_tmp) > ws_bytes) return CRFP_ERR_WORKSPACE;
  const int n = s->n, t = s->t, h = s->h, w = s->w;
  const size_t hw = (size_t)h * w;
  CRFP_TRY(crfp_nchw_to_nhwc(n * t, 3, h, w, lrs, (long long)(3 * hw), 4, lr4, st));
  // encoder_lr over all n*t frames, in chunks                                    (model/LTE.py:34-51)
  for (int i0 = 0; i0 < n * t; i0 += kPrepChunk) {
    const int cnt = (n * t - i0 < kPrepChunk) ? (n * t - i0) : kPrepChunk;
    CRFP_TRY(CB(cnt, h, w).src(lr4 + (size_t)i0 * hw * 4, 3, 4).layer(W, L_ENC_LR_0).act(CRFP_ACT_LRELU)
                 .dst(enc_tmp, 32, 32).run(st));
    CRFP_TRY(CB(cnt, h, w).src(enc_tmp, 32, 32).layer(W, L_ENC_LR_2).act(CRFP_ACT_LRELU)
                 .dst(x_lr + (size_t)i0 * hw * 32, 32, 32).run(st));
  }
  // flows: frame i (cur) -> frame i-1 (prev), per clip so that pairs are consecutive images
  if (prev_lr) CRFP_TRY(crfp_nchw_to_nhwc(n, 3, h, w, prev_lr, (long long)(3 * hw), 4, f.prev4, st));
  for (int b = 0; b < n; ++b) {
    if (prev_lr)
      CRFP_TRY(run_fnet(W, 1, h, w, lr4 + (size_t)(b * t) * hw * 4, f.prev4 + (size_t)b * hw * 4, f,
                        flows + (size_t)(b * t) * hw * 2, st));
    for (int i0 = 1; i0 < t; i0 += kPrepChunk) {
      const int cnt = (t - i0 < kPrepChunk) ? (t - i0) : kPrepChunk;
      CRFP_TRY(run_fnet(W, cnt, h, w, lr4 + (size_t)(b * t + i0) * hw * 4, lr4 + (size_t)(b * t + i0 - 1) * hw * 4, f,
                        flows + (size_t)(b * t + i0) * hw * 2, st));
    }
  }
  return CRFP_OK;
}

extern "C" size_t crfp_dsv_frame_workspace(const crfp_dsv_shape* s) {
  if (check_shape(s) != CRFP_OK) return 0;
  FrameWs f;
  return carve_frame(s, nullptr, &f);
}

extern "C" int crfp_dsv_frame(const crfp_dsv_frame_desc* d, const crfp_dsv_weights* W, void* workspace,
                              size_t ws_bytes, crfp_stream stream) {
  if (!d || !W || !workspace) return CRFP_ERR_NULL;
  const crfp_dsv_shape* s = &d->shape;
  CRFP_TRY(check_shape(s));
  if (W->nlayers != L_COUNT || W->mid_channels != 32) return CRFP_ERR_BAD_SHAPE;
  if (!d->lr4 || !d->x_lr || !d->fvs || !d->mks || !d->state_hr || !d->state_l1 || !d->out) return CRFP_ERR_NULL;
  if (!d->first && !d->flow) return CRFP_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  FrameWs f;
  if (carve_frame(s, workspace, &f) > ws_bytes) return CRFP_ERR_WORKSPACE;
  const int n = s->n, h = s->h, w = s->w;
  const int h1 = 2 * h, w1 = 2 * w, H = 8 * h, Wd = 8 * w;
  const size_t hw = (size_t)h * w;

  // dense copies of this frame's LR features / flow (the clip tensors are (b, t, ...) so a frame is strided)
  const float* x_lr = d->x_lr;
  if (n > 1 && d->x_lr_clip_stride != (long long)(hw * 32)) {
    CRFP_TRY(gather_images(n, (long long)hw * 32, d->x_lr, d->x_lr_clip_stride, f.x_lr_d, st));
    x_lr = f.x_lr_d;
  }
  const float* flow_dense = d->flow;
  if (!d->first && n > 1 && d->flow_clip_stride != (long long)(hw * 2)) {
    CRFP_TRY(gather_images(n, (long long)hw * 2, d->flow, d->flow_clip_stride, f.flow_d, st));
    flow_dense = f.flow_d;
  }
  const int tiles_x = (Wd + 31) / 32, tiles_y = (H + 31) / 32;
  const uint8_t* flags = nullptr;

  // ---- off-chain work: depends only on the inputs and on the previous frame's HR state.  With an aux stream it runs
  // beside the L1 chain (forked here, joined before the HR stage); without one it is enqueued in place.
  Aux ax;
  ax.main = st; ax.aux = (cudaStream_t)d->aux_stream;
  ax.on = d->aux_stream != nullptr && W->variant == CRFP_VARIANT_DSV && W->precision != CRFP_PREC_BF16;
  for (int i = 0; i < 3; ++i) {
    ax.ev[i] = (cudaEvent_t)d->aux_events[i];
    if (ax.on && !ax.ev[i]) return CRFP_ERR_NULL;
  }
  auto hr_side_work = [&](cudaStream_t s2, bool with_warp) -> int {
    if (with_warp && !d->first) {   // flow_lv0 = up8(flow)*8, warped HR state                              (CRFP.py:1566,1571)
      CRFP_TRY(crfp_resize_bilinear(n, h, w, 2, flow_dense, H, Wd, 0.125f, 0.125f, 8.f, f.flow_hr, s2));
      crfp_warp_desc wd;
      memset(&wd, 0, sizeof(wd));
      wd.n = n; wd.h = H; wd.w = Wd; wd.c = 4; wd.x = d->state_hr; wd.x_cstride = 4; wd.flow = f.flow_hr;
      wd.out = f.S0_w; wd.out_cstride = 4;
      CRFP_TRY(launch_flow_warp(wd, s2));
    }
    // fovea compositing + encoder_hr for this frame                                          (CRFP.py:1542-1547)
    // Outside the mask the blend keeps S (0*F + 1*S): encoder_hr and conv_tttf are only CONSUMED within 3 pixels of a
    // mask pixel, so tiles whose one-tile-dilated neighbourhood holds no mask pixel skip them (bit-identical output).
    if (d->skip_outside_fovea) {
      uint8_t* fl = reinterpret_cast<uint8_t*>(f.tile_flags);
      uint8_t* any = fl + (((size_t)n * tiles_x * tiles_y + 255) & ~(size_t)255);
      launch_k(fovea_tile_any_kernel, dim3(tiles_y, n), dim3(256), (size_t)tiles_x * sizeof(int), s2, H, Wd, d->mks, d->mks_clip_stride, tiles_x, tiles_y, any);
      CRFP_TRY(check_launch());
      launch_k(fovea_tile_dilate_kernel, dim3((n * tiles_x * tiles_y + 255) / 256), dim3(256), (size_t)0, s2, tiles_x, tiles_y, n, (const uint8_t*)any, fl);
      CRFP_TRY(check_launch());
      flags = fl;
    }
    CRFP_TRY(launch_compose(n, H, Wd, d, f.hr_in, flags, tiles_x, tiles_y, s2));
    CB e0(n, H, Wd);
    e0.src(f.hr_in, 6, 8).layer(W, L_ENC_HR_0).act(CRFP_ACT_LRELU).dst(f.e1, 4, 4);
    e0.p.tile_flags = flags; e0.p.tiles_x = tiles_x; e0.p.tiles_y = tiles_y; e0.p.tile_mode = 1;
    CRFP_TRY(e0.run(s2));
    CB e2(n, H, Wd);
    e2.src(f.e1, 4, 4).layer(W, L_ENC_HR_2).act(CRFP_ACT_LRELU).dst(f.x_hr, 4, 4);
    e2.p.tile_flags = flags; e2.p.tiles_x = tiles_x; e2.p.tiles_y = tiles_y; e2.p.tile_mode = 1;
    CRFP_TRY(e2.run(s2));
    return CRFP_OK;
  };
  bool hr_side_done = false;
  if (ax.on) {
    CRFP_TRY(ax.order(ax.main, ax.aux, 0));
    CRFP_TRY(hr_side_work(ax.aux, true));
    hr_side_done = true;
  }

  if (W->variant != CRFP_VARIANT_DSV) {
    if (W->variant != CRFP_VARIANT_V15 && W->variant != CRFP_VARIANT_V13) return CRFP_ERR_UNSUPPORTED;
    if (d->fg || W->precision == CRFP_PREC_BF16) return CRFP_ERR_UNSUPPORTED;
    CRFP_TRY(frame_l1_v1x(d, W, f, x_lr, st));
  } else if (W->precision == CRFP_PREC_BF16) {
    if (d->fg) return CRFP_ERR_UNSUPPORTED;  // regional masking is only wired in the fp32 path
    CRFP_TRY(frame_l1_bf16(d, W, f, x_lr, st));
  } else {
    // feat_prop_lv0 = PixelShufflePack(x_lr): 32 -> 96 @LR, shuffle x2 -> 24ch @L1           (CRFP.py:1560)
    float* prop_dst = d->first ? f.prop : f.cur[0];
    const int prop_cs = d->first ? 24 : 32;
    CRFP_TRY(CB(n, h, w).src(x_lr, 32, 32).layer(W, L_UPSAMPLE).shuffle(2).dst(prop_dst, 24, prop_cs).run(st));

    const float* prop24 = nullptr;  // input of upsample_post
    if (!d->first) {
      const float* flow = flow_dense;
      // flow_lv3 = up2(flow)*2                                                                (CRFP.py:1565)
      CRFP_TRY(crfp_resize_bilinear(n, h, w, 2, flow, h1, w1, 0.5f, 0.5f, 2.f, f.flow_l1, st));
      // P = downsample(S0): pixel_unshuffle(4) + conv 64 -> 32                                (CRFP.py:1569)
      CRFP_TRY(CB(n, h1, w1).src(d->state_hr, 64, 4, 0, CRFP_SRC_UNSHUFFLE4).layer(W, L_DOWNSAMPLE).dst(f.P, 32, 32).run(st));
      // warps                                                                                 (CRFP.py:1570-1577)
      // P -> P_w and the warped feat_lv{k} (channels 24..31 of level k's `cur`): one launch for all four warps
      CRFP_TRY(launch_flow_warp_l1(n, h1, w1, f.flow_l1, f.P, f.P_w, d->state_l1, f.cur[0], f.cur[1], f.cur[2], st));
      const float* fg_l1 = nullptr;
      if (d->fg) {  // streaming regional mask at L1: bilinear x0.25                            (CRFP_test.py:2299-2300)
        if (n > 1 && d->fg_clip_stride != (long long)H * Wd) return CRFP_ERR_UNSUPPORTED;
        CRFP_TRY(crfp_resize_bilinear(n, H, Wd, 1, d->fg, h1, w1, 4.f, 4.f, 1.f, f.fg_l1, st));
        fg_l1 = f.fg_l1;
      }
      static const int lb0[3] = {L_DCN0_B0, L_DCN1_B0, L_DCN2_B0};
      static const int lb2[3] = {L_DCN0_B2, L_DCN1_B2, L_DCN2_B2};
      static const int lfu[3] = {-1, L_DCN1_FUSE, L_DCN2_FUSE};
      static const int lhd[3] = {L_DCN0_HEADS, L_DCN1_HEADS, L_DCN2_HEADS};
      static const int ldc[3] = {L_DCN0_DCN, L_DCN1_DCN, L_DCN2_DCN};
      static const int lri[3] = {L_RES0_IN, L_RES1_IN, L_RES2_IN};
      static const int lr1[3] = {L_RES0_C1, L_RES1_C1, L_RES2_C1};
      static const int lr2[3] = {L_RES0_C2, L_RES1_C2, L_RES2_C2};
      const float* offfeat = nullptr;
      for (int k = 0; k < 3; ++k) {
        float* cur = f.cur[k];
        // DCN_module (CRFP.py:324-352)
        CRFP_TRY(CB(n, h1, w1).src(cur, 32, 32).src(f.P_w, 32, 32).src(f.flow_l1, 2, 2).layer(W, lb0[k])
                     .act(CRFP_ACT_LRELU).dst(f.t1, 32, 32).run(st));
        float* z = f.offf[k & 1];
        if (k == 0) {
          CRFP_TRY(CB(n, h1, w1).src(f.t1, 32, 32).layer(W, lb2[k]).act(CRFP_ACT_LRELU).dst(z, 32, 32).run(st));
        } else {
          CRFP_TRY(CB(n, h1, w1).src(f.t1, 32, 32).layer(W, lb2[k]).act(CRFP_ACT_LRELU).dst(f.t2, 32, 32).run(st));
          CRFP_TRY(CB(n, h1, w1).src(f.t2, 32, 32).src(offfeat, 32, 32).layer(W, lfu[k]).act(CRFP_ACT_LRELU)
                       .dst(z, 32, 32).run(st));
        }
        offfeat = z;
        if (k == 2 && ax.on) {   // dcn_3.upsample only needs level 2's offset feature: beside the rest of the chain
          CRFP_TRY(ax.order(ax.main, ax.aux, 1));
          CRFP_TRY(CB(n, h1, w1).src(offfeat, 32, 32).layer(W, L_DCN3_UP).shuffle(4).scale(2.f).dst(f.po, 4, 4).run(ax.aux));
        }
        CRFP_TRY(run_heads_dcn(W, n, h1, w1, z, f.flow_l1, f.om, f.P, f.A, lhd[k], ldc[k], st));
        // ResidualBlocksWithInputConv on cat(cur, A) (CRFP.py:1589-1596)
        CB in(n, h1, w1);
        in.src(cur, 32, 32).src(f.A, 32, 32).layer(W, lri[k]).act(CRFP_ACT_LRELU).dst(f.r0, 32, 32);
        if (fg_l1 && k > 0) in.fg(fg_l1, (long long)h1 * w1);
        CRFP_TRY(in.run(st));
        CRFP_TRY(CB(n, h1, w1).src(f.r0, 32, 32).layer(W, lr1[k]).act(CRFP_ACT_RELU).dst(f.r1, 32, 32).run(st));
        // split: first 24 channels propagate, last 8 become feat_lv{k} of the next frame
        float* nxt = (k < 2) ? f.cur[k + 1] : f.prop;
        const int nxt_cs = (k < 2) ? 32 : 24;
        CRFP_TRY(CB(n, h1, w1).src(f.r1, 32, 32).layer(W, lr2[k]).res(f.r0, 32).dst(nxt, 24, nxt_cs)
                     .dst(d->state_l1, 8, 24, 8 * k).run(st));
      }
      prop24 = f.prop;
      // L3                                                                                    (CRFP.py:1625-1630)
      CRFP_TRY(CB(n, h1, w1).src(prop24, 24, 24).layer(W, L_UPSAMPLE_POST).act(CRFP_ACT_LRELU).shuffle(4)
                   .dst(f.q, 4, 4).run(st));
      if (!ax.on)
        CRFP_TRY(CB(n, h1, w1).src(offfeat, 32, 32).layer(W, L_DCN3_UP).shuffle(4).scale(2.f).dst(f.po, 4, 4).run(st));
    } else {
      // first frame: no alignment; cat([prop, zeros32, zeros8]) == only weight[:, :24] contributes (CRFP.py:1634-1670)
      static const int lrf[3] = {L_RES0_IN_FIRST, L_RES1_IN_FIRST, L_RES2_IN_FIRST};
      static const int lr1[3] = {L_RES0_C1, L_RES1_C1, L_RES2_C1};
      static const int lr2[3] = {L_RES0_C2, L_RES1_C2, L_RES2_C2};
      float* pa = f.prop;       // 24ch ping
      float* pb = f.cur[0];     // reuse as 24ch pong (pixel stride 24)
      for (int k = 0; k < 3; ++k) {
        CRFP_TRY(CB(n, h1, w1).src(pa, 24, 24).layer(W, lrf[k]).act(CRFP_ACT_LRELU).dst(f.r0, 32, 32).run(st));
        CRFP_TRY(CB(n, h1, w1).src(f.r0, 32, 32).layer(W, lr1[k]).act(CRFP_ACT_RELU).dst(f.r1, 32, 32).run(st));
        CRFP_TRY(CB(n, h1, w1).src(f.r1, 32, 32).layer(W, lr2[k]).res(f.r0, 32).dst(pb, 24, 24)
                     .dst(d->state_l1, 8, 24, 8 * k).run(st));
        float* tmp = pa; pa = pb; pb = tmp;
      }
      prop24 = pa;
      CRFP_TRY(CB(n, h1, w1).src(prop24, 24, 24).layer(W, L_UPSAMPLE_POST).act(CRFP_ACT_LRELU).shuffle(4)
                   .dst(f.q, 4, 4).run(st));
    }
  }

  // ---------------- HR stage (8h x 8w, 4 channels, fp32 in both precisions)
  if (ax.on) CRFP_TRY(ax.order(ax.aux, ax.main, 2));   // join: S0_w, flow_hr, x_hr, flags, po are complete
  if (!hr_side_done && W->variant == CRFP_VARIANT_DSV && W->precision != CRFP_PREC_BF16) {
    CRFP_TRY(hr_side_work(st, true));
    hr_side_done = true;
  }
  if (!d->first) {
    CRFP_TRY(CB(n, H, Wd).src(f.q, 4, 4).src(f.S0_w, 4, 4).src(f.flow_hr, 2, 2).layer(W, L_DCN3_B0).act(CRFP_ACT_LRELU)
                 .dst(f.h1, 4, 4).run(st));
    CRFP_TRY(CB(n, H, Wd).src(f.h1, 4, 4).layer(W, L_DCN3_B2).act(CRFP_ACT_LRELU).dst(f.h2, 4, 4).run(st));
    CRFP_TRY(CB(n, H, Wd).src(f.h2, 4, 4).src(f.po, 4, 4).layer(W, L_DCN3_FUSE).act(CRFP_ACT_LRELU).dst(f.h3, 4, 4).run(st));
    CRFP_TRY(CB(n, H, Wd).src(f.h3, 4, 4).layer(W, L_DCN3_HEADS).head(f.flow_hr, 2, 10.f).dst(f.om3, 4, 4).run(st));
    crfp_dcn_desc dd;
    memset(&dd, 0, sizeof(dd));
    dd.n = n; dd.h = H; dd.w = Wd; dd.c = 4; dd.cout = 4; dd.dg = 1; dd.shared_taps = 1;
    dd.x = d->state_hr; dd.x_cstride = 4;
    dd.offset = f.om3; dd.off_cstride = 4; dd.off_coffset = 0;
    dd.mask = f.om3; dd.mask_cstride = 4; dd.mask_coffset = 2;
    dd.weight = W->layer[L_DCN3_DCN].w; dd.bias = W->layer[L_DCN3_DCN].b;
    dd.out = f.A3; dd.out_cstride = 4;
    CRFP_TRY(launch_dcn(dd, st));
    CB in3(n, H, Wd);
    in3.src(f.q, 4, 4).src(f.A3, 4, 4);
    if (W->variant == CRFP_VARIANT_V15) in3.src(f.S0_w, 4, 4);   // cat([q, A3, warped state]) (CRFP.py:1332)
    in3.layer(W, L_RES3_IN).act(CRFP_ACT_LRELU).dst(f.g0, 4, 4);
    if (d->fg) in3.fg(d->fg, d->fg_clip_stride);
    CRFP_TRY(in3.run(st));
  } else {
    CRFP_TRY(CB(n, H, Wd).src(f.q, 4, 4).layer(W, L_RES3_IN_FIRST).act(CRFP_ACT_LRELU).dst(f.g0, 4, 4).run(st));
  }
  CRFP_TRY(CB(n, H, Wd).src(f.g0, 4, 4).layer(W, L_RES3_C1).act(CRFP_ACT_RELU).dst(f.g1, 4, 4).run(st));
  CRFP_TRY(CB(n, H, Wd).src(f.g1, 4, 4).layer(W, L_RES3_C2).res(f.g0, 4).dst(f.S_pre, 4, 4).run(st));

  if (!hr_side_done)   // v15 / v13 / bf16 L1 paths resize the flow and warp the HR state themselves: only the fovea part is left
    CRFP_TRY(hr_side_work(st, false));
  // conv_tttf + blend + LeakyReLU -> new state                                             (CRFP.py:1672-1675)
  {
    CB b(n, H, Wd);
    b.src(f.S_pre, 4, 4).src(f.x_hr, 4, 4).layer(W, L_TTTF).dst(d->state_hr, 4, 4);
    b.p.epi = EPI_BLEND; b.p.mask = d->mks; b.p.mask_clip_stride = d->mks_clip_stride; b.p.blend_old = f.S_pre;
    b.p.tile_flags = flags; b.p.tiles_x = tiles_x; b.p.tiles_y = tiles_y; b.p.tile_mode = 2;
    CRFP_TRY(b.run(st));
  }
  // out = conv_last(S) + up8(lr)                                                           (CRFP.py:1678-1683)
  {
    CB o(n, H, Wd);
    o.src(d->state_hr, 4, 4).layer(W, L_LAST).dst(d->out, 3, 3);
    o.p.epi = EPI_OUT_NCHW; o.p.base_lr4 = d->lr4; o.p.base_clip_stride = d->lr4_clip_stride;
    o.p.out_clip_stride = d->out_clip_stride; o.p.out_planes = 3;
    CRFP_TRY(o.run(st));
  }
  return CRFP_OK;
}
