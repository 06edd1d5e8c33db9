/*
 * crfp_b200.h — C ABI of libcrfp_b200.so: the B200 (sm_100a) implementation of CRFP's recurrent
 * cross-resolution propagation hot path (reference: eugenelet/CRFP, model/CRFP.py).
 *
 * Conventions
 *   - plain C: device pointers + sizes, no torch types; every pointer is CALLER-OWNED device memory.
 *   - every entry point is asynchronous on the given `cudaStream_t` (passed as void*), never
 *     allocates, never synchronises, keeps no global mutable state, and returns 0 (CRFP_OK) or a
 *     negative crfp_status; it never throws.
 *   - activations are fp32 NHWC ("channels-last": a pixel's channels are contiguous); a tensor is
 *     addressed as ptr[((n*H + y)*W + x)*cstride + coffset + c] so that channel slices of a wider
 *     tensor can be read/written in place (this is how torch.cat / torch.chunk of the reference
 *     disappear).  The user-facing NCHW tensors (lrs, fvs, mks, output) are read/written directly.
 *   - conv weights are passed PACKED, see crfp_conv3x3_fwd.
 *
 * Reference interfaces replaced (file:line under /root/reference):
 *   crfp_flow_warp_fwd        flow_warp(x, flow)                         model/CRFP.py:90-130
 *   crfp_dcn_v2_fwd           dcn_v2.DCNv2.forward(input, offset, mask)  model/CRFP.py:318-320,350
 *                             (= _ext.dcn_v2_forward of jinfagang/DCNv2_latest, README.md:26)
 *   crfp_conv3x3_fwd          nn.Conv2d(3x3,s1,p1) + bias + LeakyReLU/ReLU + residual + torch.cat of
 *                             the inputs + torch.chunk / F.pixel_shuffle / pixel_unshuffle of the
 *                             output:  model/CRFP.py:154-193, 239-279, 28-42, 433-552, 303-317
 *   crfp_resize_bilinear      nn.Upsample(bilinear, align_corners=False) model/CRFP.py:1471-1478,
 *                             F.interpolate(size=) model/CRFP.py:808-812
 *   crfp_avgpool2             nn.AvgPool2d(2,2)                          model/CRFP.py:755,762,769
 *   crfp_dsv_prepare          CRFP_DSV.compute_flow + encoder_lr         model/CRFP.py:1483-1508,1540
 *   crfp_dsv_frame            one iteration of the t-loop of CRFP_DSV.forward   model/CRFP.py:1555-1684
 *                             incl. fovea compositing + encoder_hr (1542-1547) for that frame
 *   crfp_*_bwd, crfp_charbonnier_fwd_bwd, crfp_adam_step   loss.backward() + optimizer.step()   trainer.py:246-250
 *   crfp_conv_kxk_fwd, crfp_resize_bilinear_ac, crfp_channel_affine   SPyNet (legacy)        model/CRFP.py:554-741
 */
#ifndef CRFP_B200_H
#define CRFP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRFP_ABI_VERSION 1

typedef enum {
  CRFP_OK = 0,
  CRFP_ERR_BAD_SHAPE = -1,    /* inconsistent sizes / channel counts */
  CRFP_ERR_UNSUPPORTED = -2,  /* valid request the library has no kernel for */
  CRFP_ERR_WORKSPACE = -3,    /* workspace too small */
  CRFP_ERR_CUDA = -4,         /* a CUDA runtime call or launch failed (see crfp_last_cuda_error) */
  CRFP_ERR_NULL = -5          /* required pointer is NULL */
} crfp_status;

typedef void* crfp_stream; /* cudaStream_t */

/* ------------------------------------------------------------------ misc */
int crfp_abi_version(void);
const char* crfp_status_string(int status);
/* last cudaError_t seen by this thread inside the library, as text (thread-local) */
const char* crfp_last_cuda_error(void);
/* number of kernel launches issued by this library from the calling thread since the last reset */
long long crfp_launch_count(void);
void crfp_launch_count_reset(void);
/* accounting hook for callers that capture this library's launches into a CUDA graph: a capture enqueues nothing
 * (add -n), every replay of the graph launches its n kernels (add +n) */
void crfp_launch_count_add(long long n);
/* device properties check: 0 if the current device is sm_100 (B200), CRFP_ERR_UNSUPPORTED otherwise */
int crfp_check_device(void);

/* tcgen05 plumbing self-test: D[m][n] = sum_k A[m+shift][k]*B[n][k] (m<128) with bf16 A[rowsA][K], B[N][K]
 * row-major in HBM, fp32 D[128][N]; exercises the shifted-window operand addressing of the conv kernels. */
int crfp_selftest_umma(int rowsA, int K, int N, int shift, const void* A, const void* B, float* D, crfp_stream stream);

/* micro-benchmark: `ctas` CTAs each issue `reps` back-to-back tcgen05.mma (M128 x N x K16 bf16) from one thread;
 * cycles_out[cta] = SM cycles from first issue to completion (device buffer of `ctas` int64). */
/* as crfp_selftest_umma with an explicit A-operand SBO (16-byte records between 8-row groups):
 * D[m][n] = sum_k A[shift + (m/8)*sbo_recs + m%8][k] * B[n][k] */
int crfp_selftest_umma_sbo(int rowsA, int K, int N, int shift, int sbo_recs, const void* A, const void* B, float* D,
                           crfp_stream stream);
int crfp_selftest_umma_rate(int N, int reps, int ctas, long long* cycles_out, crfp_stream stream);

/* ------------------------------------------------------------------ activations / epilogues */
enum {
  CRFP_ACT_NONE = 0,
  CRFP_ACT_LRELU = 1,     /* LeakyReLU(0.1) */
  CRFP_ACT_RELU = 2,
  CRFP_ACT_DCN_HEAD = 3,  /* channels [0, head_split): max_mag*tanh(v) + flow (even ch: flow y, odd ch: flow x);
                             channels [head_split, cout): sigmoid(v)        (model/CRFP.py:337-340,349) */
  CRFP_ACT_TANH256 = 4    /* tanh(v) * 256                                   (model/CRFP.py:807) */
};

enum { CRFP_SRC_PLAIN = 0, CRFP_SRC_UNSHUFFLE4 = 1 };
enum { CRFP_OUT_NHWC = 0, CRFP_OUT_SHUFFLE = 1 };

typedef struct {
  const float* ptr; /* NHWC base of image 0 */
  int32_t c;        /* channels taken from this source */
  int32_t cstride;  /* floats per pixel of the underlying tensor */
  int32_t coffset;  /* first channel taken */
  int32_t mode;     /* CRFP_SRC_PLAIN, or CRFP_SRC_UNSHUFFLE4: the source is a (4h x 4w) plane with `c/16`
                       channels read through pixel_unshuffle(4) (model/CRFP.py:28-42); packed channel order
                       of such a source is (dy*4+dx)*(c/16) + ch, see crfp_conv_pack_index */
} crfp_src;

typedef struct {
  float* ptr;      /* NHWC base of image 0 */
  int32_t c;       /* number of output channels routed here (segments are consecutive) */
  int32_t cstride;
  int32_t coffset;
  int32_t _pad;
} crfp_dst;

/*
 * 3x3, stride 1, pad 1 convolution over the channel-concatenation of up to 3 NHWC sources.
 *   packed input channel space: each source is padded up to a multiple of 4 channels, the total up to
 *   a multiple of 8 (= cin_packed); packed output channels: cout padded up to a multiple of
 *   crfp_conv_cout_pad(cout).  weight[(tap*cin_packed + ci)*cout_packed + co], tap = ky*3+kx, zero in
 *   all padding; bias[cout_packed].
 *   out_mode CRFP_OUT_NHWC: conv channel co goes to the dst segment that contains it.
 *   out_mode CRFP_OUT_SHUFFLE: F.pixel_shuffle(r): conv channel o*r*r + dy*r + dx goes to pixel
 *   (y*r+dy, x*r+dx), channel o of dst[0] (an (h*r x w*r) plane).
 *   epilogue order: v = acc + bias; v = act(v); v += residual; v *= post_scale.
 */
typedef struct {
  int32_t n, h, w;
  int32_t nsrc;
  crfp_src src[3];
  int32_t cout;
  int32_t act;
  const float* weight;
  const float* bias;
  int32_t out_mode;
  int32_t shuffle_r;
  int32_t ndst;
  int32_t head_split;   /* CRFP_ACT_DCN_HEAD only */
  crfp_dst dst[2];
  const float* residual; /* optional NHWC, added after the activation */
  int32_t res_cstride;
  int32_t res_coffset;
  const float* flow;     /* CRFP_ACT_DCN_HEAD only: NHWC 2-channel flow (x, y) at the output resolution */
  float post_scale;      /* 0 is treated as 1 */
  float head_mag;        /* CRFP_ACT_DCN_HEAD: max_residue_magnitude (10) */
} crfp_conv_desc;

int crfp_conv3x3_fwd(const crfp_conv_desc* d, crfp_stream stream);
/* packed sizes for a conv over sources with `c[i]` channels */
int crfp_conv_cin_packed(int nsrc, const int32_t* c);
int crfp_conv_cout_packed(int cout);
size_t crfp_sizeof_conv_desc(void);

/* ------------------------------------------------------------------ tensor-core conv (bf16 storage) */
/*
 * Same convolution on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulation in TMEM): sources are bf16 NHWC
 * with channel counts / strides / offsets that are multiples of 8; the packed K space is the plain
 * concatenation of the sources (padded to a multiple of 16 channels).
 *   weight: bf16 [ntiles][9][kc][nt][8] with (nt, ntiles) = crfp_tc_cout_tile(cout), kc = K/8:
 *           element = W[tile*nt + n][kc*8 + j][tap]; bias fp32 [ntiles*nt].
 *   out_kind CRFP_TC_OUT_BF16: bf16 NHWC, up to 2 channel segments (multiples of 8);
 *            CRFP_TC_OUT_F32: fp32 NHWC into dst[0]; CRFP_TC_OUT_SHUFFLE_F32: pixel_shuffle(r) into fp32 dst[0].
 *   residual: optional bf16 NHWC.  act / head_* / post_scale as in crfp_conv_desc.
 */
enum { CRFP_TC_OUT_BF16 = 0, CRFP_TC_OUT_F32 = 1, CRFP_TC_OUT_SHUFFLE_F32 = 2 };
typedef struct {
  const void* ptr; /* bf16 NHWC */
  int32_t c, cstride, coffset, _pad;
} crfp_tc_src;
typedef struct {
  void* ptr;
  int32_t c, cstride, coffset, _pad;
} crfp_tc_dst;
typedef struct {
  int32_t n, h, w;
  int32_t nsrc;
  crfp_tc_src src[3];
  int32_t cout;
  int32_t act;
  const void* weight;
  const float* bias;
  int32_t out_kind;
  int32_t shuffle_r;
  int32_t ndst;
  int32_t head_split;
  crfp_tc_dst dst[2];
  const void* residual; /* bf16 NHWC */
  int32_t res_cstride;
  int32_t res_coffset;
  const float* flow;
  float post_scale;
  float head_mag;
} crfp_conv_tc_desc;
int crfp_conv3x3_tc_fwd(const crfp_conv_tc_desc* d, crfp_stream stream);
int crfp_tc_cout_tile(int cout, int32_t* nt, int32_t* ntiles);
size_t crfp_sizeof_conv_tc_desc(void);

/*
 * fp32-ACCURATE tensor-core conv: fp32 NHWC sources/outputs, operands split hi/lo into bf16 inside the kernel,
 * A_hi*W_hi + A_lo*W_hi + A_hi*W_lo accumulated in fp32 TMEM (error ~2^-17 relative: keeps the fp32 parity bar).
 *   sources: fp32 NHWC, channel counts multiples of 8 (strides/offsets multiples of 4)
 *   weight_hi / weight_lo: bf16 [ntiles][9][kc][nt][8], (nt, ntiles) = crfp_tc3_cout_tile(cout, cin); bias fp32
 *   extra / w_extra: optional 2-channel fp32 NHWC source (the flow input of dcn_block.0) convolved on the CUDA
 *           cores in the epilogue with fp32 weights [9][2][ntiles*nt]
 *   out_kind: CRFP_TC_OUT_F32 (up to 2 channel segments, multiples of 4) or CRFP_TC_OUT_SHUFFLE_F32.
 */
typedef struct {
  const float* ptr;
  int32_t c, cstride, coffset;
  int32_t _pad; /* = mode: CRFP_SRC_PLAIN, or CRFP_SRC_UNSHUFFLE4 (dense 4-channel (4h x 4w) plane read through
                   pixel_unshuffle(4): c = 64, cstride = 4, coffset = 0; packed channel order (dy*4+dx)*4 + ch) */
} crfp_tc3_src;
typedef struct {
  int32_t n, h, w;
  int32_t nsrc;
  crfp_tc3_src src[3];
  int32_t cout;
  int32_t act;
  const void* weight_hi;
  const void* weight_lo;
  const float* bias;
  const float* extra;
  const float* w_extra;
  int32_t out_kind;
  int32_t shuffle_r;
  int32_t ndst;
  int32_t head_split;
  crfp_tc_dst dst[2];
  const float* residual;
  int32_t res_cstride;
  int32_t res_coffset;
  const float* flow;
  float post_scale;
  float head_mag;
  int32_t half;    /* != 0: CRFP_PREC_HALF arithmetic (weight_hi / weight_lo are fp16, activations one fp16 product) */
  int32_t res_pre; /* != 0: `residual` is added BEFORE the activation: a conv over more than 64 input channels runs as passes of
                      64 whose partial sums meet here (pass 1: bias, no activation; last pass: zero bias, residual = the
                      destination itself, the layer's activation) */
} crfp_conv_tc3_desc;
int crfp_conv3x3_tc3_fwd(const crfp_conv_tc3_desc* d, crfp_stream stream);
/* profiling aid: same launch + a clock64 trace of CTA (0,0,0): trace[0..1] = kernel start / loop start, then per row
 * trace[8*(r+1) + k], k = loop top, row staged, converted, next row issued, MMAs issued, MMAs done, epilogue done */
int crfp_conv3x3_tc3_trace(const crfp_conv_tc3_desc* d, long long* trace, crfp_stream stream);
/* cout tiling for `cin` tensor-core input channels (multiple of 8); CRFP_ERR_UNSUPPORTED when cin is too large for
 * the shared-memory resident rings (cin > 64) */
int crfp_tc3_cout_tile(int cout, int cin, int32_t* nt, int32_t* ntiles);
size_t crfp_sizeof_conv_tc3_desc(void);
/*
 * Packs an OIHW fp32 nn.Conv2d weight (+ bias) into crfp_conv3x3_tc3_fwd's operands in one launch (the training step
 * repacks every layer after every optimiser update).  Logical operator V[o][i][tap], o < nout, i < k + extra:
 *   transposed == 0: V[o][i][tap] = W[o][lo + i][tap]            (forward; nout == cout_w)
 *   transposed == 1: V[o][i][tap] = W[k_lo + i][lo + o][8 - tap] (backward data of input channels [lo, lo + nout): the
 *                    transposed, 180-degree rotated kernel; K slice [k_lo, k_lo + k) of W's output channels; extra == 0,
 *                    bias ignored -> zeros).  k_lo must be 0 when transposed == 0 (use `lo`).
 * w_hi / w_lo: bf16 [ntiles][9][kc][nt][8] with (nt, ntiles) = crfp_tc3_cout_tile(nout, k), kc = k/8 rounded up to even;
 * bias_packed fp32 [ntiles*nt]; w_extra fp32 [9][extra][ntiles*nt] (extra > 0 only).
 */
int crfp_pack_conv_tc3(const float* weight, const float* bias, int cout_w, int cin_w, int transposed, int lo, int nout, int k,
                       int extra, int k_lo, void* w_hi, void* w_lo, float* bias_packed, float* w_extra, crfp_stream stream);

/* ------------------------------------------------------------------ flow_warp */
/*
 * out[n,y,x,c] = bilinear(x_in[n,:,:,c], y + flow[n,y,x,1], x + flow[n,y,x,0]); corners outside the
 * image contribute 0 (padding_mode='zeros', border=0) or are clamped (padding_mode='border', border=1);
 * sampling positions are computed with the reference's fp32 op sequence (normalise to [-1,1], then
 * grid_sample align_corners=True) so that floor() agrees bit for bit.
 */
typedef struct {
  int32_t n, h, w, c;      /* c % 4 == 0 */
  const float* x; int32_t x_cstride, x_coffset;
  const float* flow;       /* NHWC, 2 channels (x, y) */
  float* out; int32_t out_cstride, out_coffset;
  int32_t border;
  int32_t _pad;
} crfp_warp_desc;
int crfp_flow_warp_fwd(const crfp_warp_desc* d, crfp_stream stream);
/* same, x / out bf16 NHWC with c % 8 == 0 (zeros padding only) */
int crfp_flow_warp_bf16_fwd(const crfp_warp_desc* d, crfp_stream stream);
/* debug/parity: the integer corner indices (x0, y0) flow_warp uses at every pixel: int32 [n,h,w] each */
int crfp_flow_warp_indices(int n, int h, int w, const float* flow, int32_t* x0, int32_t* y0, crfp_stream stream);
size_t crfp_sizeof_warp_desc(void);

/* ------------------------------------------------------------------ DCNv2 */
/*
 * Modulated deformable 3x3 convolution (stride 1, pad 1, dilation 1), DCNv2 semantics:
 *   out[n,y,x,o] = b[o] + sum_{c,i,j} W[o,c,i,j] * m[n,y,x,g*9+t] * bilinear(in[n,:,:,c], py, px),
 *   t = i*3+j, g = c / (C/dg), py = (y-1+i) + off[n,y,x,(g*9+t)*2], px = (x-1+j) + off[...+1].
 * offset and mask are NHWC with their own pixel strides / channel offsets (so they may live in one fused
 * "heads" tensor).  shared_taps=1: one (dy,dx) and one mask per deformable group shared by the 9 taps
 * (the `repeat=True` HR module, model/CRFP.py:341-347): offset has 2*dg channels ordered [dy(g)..., dx(g)...]
 * and mask dg channels.
 * weight packed as [k][co] with k = (g*9+t)*(C/dg) + c_in_group, co padded to a multiple of 4; bias[co].
 * Supported: (C=32, dg=8, cout=32) and (C=4, dg=1, cout=4).
 *
 * head_raw != 0 (the fused form of DCN_module.forward, model/CRFP.py:337-349; tensor-core align kernel only): `offset`
 * and `mask` hold the RAW outputs of the dcn_offset / dcn_mask convolutions and the sampler applies
 *   dy = head_mag * tanh(raw) + head_flow[n,y,x,1],  dx = head_mag * tanh(raw) + head_flow[n,y,x,0],  m = sigmoid(raw)
 * itself (same ex2 / rcp arithmetic as the conv epilogue's CRFP_ACT_DCN_HEAD), so the activation never costs a pass.
 * dbg_y0 / dbg_x0 (optional, parity): the kernel that PRODUCES the output also writes floor(py), floor(px) of every
 * sample it takes: int32 [n,h,w,dg*9] each (shared_taps: the 9 taps of the one group).
 */
typedef struct {
  int32_t n, h, w;
  int32_t c, cout, dg;
  int32_t shared_taps;
  int32_t head_raw;
  const float* x; int32_t x_cstride, x_coffset;
  const float* offset; int32_t off_cstride, off_coffset;
  const float* mask; int32_t mask_cstride, mask_coffset;
  const float* weight;
  const float* bias;
  float* out; int32_t out_cstride, out_coffset;
  const float* head_flow;       /* NHWC 2-channel flow (x, y) at this resolution; required when head_raw != 0 */
  float head_mag; int32_t _pad;
  int32_t* dbg_y0; int32_t* dbg_x0;
} crfp_dcn_desc;
int crfp_dcn_v2_fwd(const crfp_dcn_desc* d, crfp_stream stream);
/* tensor-core variant (C=32, dg=8, cout=32): x / out bf16 NHWC, weight bf16 [36][32][8] with k = (g*9+t)*4+c,
 * offset / mask / bias fp32; the gather writes the UMMA A tile in shared memory, tcgen05.mma contracts it. */
int crfp_dcn_v2_tc_fwd(const crfp_dcn_desc* d, crfp_stream stream);
/* fp32-accurate tensor-core variant: x / out fp32 NHWC; d->weight and weight_lo are the hi / lo bf16 halves of the
 * [36][32][8] packed weight; columns are split hi/lo in the kernel, 3 products accumulate in fp32 TMEM.
 * flow_hint (optional, NHWC 2-channel flow at this resolution) centres the shared-memory sampling window of each tile;
 * it only affects speed (samples outside the window are gathered from global memory). */
int crfp_dcn_v2_tc3_fwd(const crfp_dcn_desc* d, const void* weight_lo, const float* flow_hint, crfp_stream stream);
/*
 * dcn_align_fused — the tail of DCN_module.forward (model/CRFP.py:337-350) as ONE kernel:
 *   offset = head_mag * tanh(conv3x3(z; dcn_offset)) + flow.flip(1).repeat(72)   (144 channels)
 *   mask   = sigmoid(conv3x3(z; dcn_mask))                                        (72 channels)
 *   out    = DCNv2(x, offset, mask; dcn.weight, dcn.bias)      C = 32, dg = 8, cout = 32
 * The offset / mask tensor never reaches HBM: the head convolution's accumulator stays in TMEM and the sampler threads
 * read their raw offsets from it.  fp32 NHWC tensors; 3 x bf16 split products on tcgen05 (fp32-grade).
 *   z        32-channel offset feature (cstride / coffset in floats, multiples of 4); flow: dense NHWC 2-channel (x, y)
 *   x        32-channel tensor to align
 *   heads_w  bf16 [6][2][6][224][8]: six K "sixths" (6 chunks of 8 input channels of the tap-major K = tap*32 + c), each
 *            the hi half then the lo half, [chunk][n][8]; column n = 3*(g*9+t) + {0: dy, 1: dx, 2: mask}, columns 216..223
 *            zero (crfp_b200.packing.pack_align_heads); heads_b fp32 [224] in the same column order
 *   dcn_w_hi / dcn_w_lo / dcn_b: as crfp_dcn_v2_tc3_fwd ([36][32][8] packing, k = (g*9+t)*4 + c)
 *   dbg_y0 / dbg_x0: optional int32 [n,h,w,72] dump of floor(py), floor(px) of every sample taken (parity)
 */
typedef struct {
  int32_t n, h, w, _pad;
  const float* z; int32_t z_cstride, z_coffset;
  const float* flow;
  const float* x; int32_t x_cstride, x_coffset;
  const void* heads_w;
  const float* heads_b;
  const void* dcn_w_hi;
  const void* dcn_w_lo;
  const float* dcn_b;
  float* out; int32_t out_cstride, out_coffset;
  float head_mag; int32_t half;   /* half != 0: CRFP_PREC_HALF arithmetic (all packed weights fp16) */
  int32_t* dbg_y0; int32_t* dbg_x0;
} crfp_align_fused_desc;
int crfp_dcn_align_fused(const crfp_align_fused_desc* d, crfp_stream stream);
/* profiling aid: same launch + a clock64 timeline of CTA 0's first 16 tiles into trace[16][16] (device int64) */
int crfp_dcn_align_fused_trace(const crfp_align_fused_desc* d, long long* trace, crfp_stream stream);
size_t crfp_sizeof_align_fused_desc(void);
/* debug/parity: floor(py), floor(px) for every (pixel, group, tap): int32 [n,h,w,dg*9] each (non-shared form) */
int crfp_dcn_v2_indices(const crfp_dcn_desc* d, int32_t* y0, int32_t* x0, crfp_stream stream);
size_t crfp_sizeof_dcn_desc(void);

/* ------------------------------------------------------------------ resampling / layout */
/*
 * Bilinear resize, align_corners=False, NHWC, c channels (all of the pixel: cstride == c):
 * src = max((dst+0.5)*rscale - 0.5, 0).  rscale_{h,w} = 1/scale_factor for nn.Upsample(scale_factor=..)
 * or in/out for F.interpolate(size=..); out = value * mul.
 */
int crfp_resize_bilinear(int n, int hin, int win, int c, const float* in, int hout, int wout, float rscale_h,
                         float rscale_w, float mul, float* out, crfp_stream stream);
int crfp_avgpool2(int n, int hin, int win, int c, const float* in, float* out, crfp_stream stream);
/*
 * Fovea paste = the data loader's `ref[:, y:y+fv, x:x+fv] = patch; ref_sp[...] = 1` (dataset/reds.py:196-201) on the
 * device for `frames` = n*t frames at once: patch [frames][3][fv][fv], coords int32 [frames][2] = (y, x) top-left in
 * HR pixels (clamped to the frame), fvs [frames][3][H][W], mks [frames][H][W] (bytes).  clear != 0 zeroes the
 * rectangles instead (used with the previous call's coords so that persistent fvs / mks buffers stay exact).
 */
int crfp_fovea_paste(const float* patch, const int32_t* coords, int frames, int fv, int H, int W, float* fvs,
                     uint8_t* mks, int clear, crfp_stream stream);
/*
 * The per-pixel half of the data loader's `fovea_generator` (dataset/reds.py:190-226) for a whole clip: gt / fvs planar
 * [frames][c][H][W] fp32, rects int32 [frames][4] = (y0, x0, y1, x1) half-open and already clipped to the frame (16-byte
 * aligned), mks [frames][H][W] bytes: mks = 1 inside the rectangle, fvs = gt * mask.
 */
int crfp_fovea_from_gt(const float* gt, const int32_t* rects, int frames, int c, int H, int W, float* fvs, uint8_t* mks,
                       crfp_stream stream);
/*
 * `calc_psnr_and_ssim_cuda` (utils.py:165-254) fused into one pass: planar img1 / img2 [B][C][H][W] in [0,1], optional
 * mask [B][H][W] (fp32 or bytes; both NULL = all ones), window11 = the 11 normalised Gaussian taps (HOST pointer).
 * Writes per-CTA partial sums partial[B*C][tiles_y*tiles_x][3] = (sum ssim*m, sum (img1-img2)^2*m, sum m) with
 * (tiles_x, tiles_y) = crfp_psnr_ssim_tiles(H, W); the caller adds them up (float64) — deterministic, no atomics.
 */
int crfp_psnr_ssim_tiles(int H, int W, int32_t* tiles_x, int32_t* tiles_y);
/* frames as the reference SAVES them (trainer.py:446-474, 535-537): out[i] = uint8(round(clip(in[i] * 255, 0, 255))),
 * round-half-to-even; in / out 16-byte aligned.  Used to stream uint8 frames to the host (4x fewer PCIe bytes). */
int crfp_quantize_u8(const float* in, uint8_t* out, long long count, crfp_stream stream);
int crfp_psnr_ssim(int B, int C, int H, int W, const float* img1, const float* img2, const float* mask_f32,
                   const uint8_t* mask_u8, const float* window11, float* partial, crfp_stream stream);
/* NCHW (with an explicit image stride in floats) -> NHWC with cpad >= c channels (extra channels zero) */
int crfp_nchw_to_nhwc(int n, int c, int h, int w, const float* in, long long in_image_stride, int cpad, float* out,
                      crfp_stream stream);
int crfp_nhwc_to_nchw(int n, int c, int h, int w, const float* in, int in_cstride, int in_coffset, float* out,
                      long long out_image_stride, crfp_stream stream);

/* ------------------------------------------------------------------ CRFP_DSV composite entry points */
/* conv layers of CRFP_DSV in the order the library expects them in crfp_dsv_weights.layer[] */
#define CRFP_DSV_MAX_LAYERS 72
typedef struct {
  const float* w; /* packed as the consuming kernel expects (see crfp_dsv_layer_info) */
  const float* b;
} crfp_layer;

enum { CRFP_PREC_FP32 = 0, CRFP_PREC_BF16 = 1, CRFP_PREC_TC3 = 2, CRFP_PREC_HALF = 3 };
/* model variants sharing every kernel and parameter name (model/CRFP.py): CRFP_DSV :1387 (24/8 split state),
 * CRFP :1101 (3-way concat, HR state warped before down-sampling), CRFP_simple :816 (2-way concat) */
enum { CRFP_VARIANT_DSV = 0, CRFP_VARIANT_V15 = 1, CRFP_VARIANT_V13 = 2 };
typedef struct {
  const void* w_hi;     /* bf16 tensor-core packing (TC3: hi half of the split) */
  const void* w_lo;     /* TC3: lo half (NULL in bf16 mode) */
  const float* b;
  const float* w_extra; /* TC3: fp32 [9][2][cout_packed] weights of a trailing 2-channel source, else NULL */
  const void* w_fused;  /* TC3, dcn_k heads layers: crfp_align_fused_desc.heads_w packing (NULL: unfused heads + align) */
  const float* b_fused; /* ... and heads_b */
} crfp_layer_tc;
typedef struct {
  int32_t mid_channels; /* 32 */
  int32_t nlayers;      /* must equal crfp_dsv_num_layers() */
  int32_t precision;    /* CRFP_PREC_FP32: every layer fp32 SIMT FFMA (parity <= 1e-3);
                           CRFP_PREC_TC3 : fp32 storage, the dense layers (cin <= 64) and the L1 DCN contraction on
                                           tcgen05 as 3 x bf16 split products with fp32 TMEM accumulation — fp32-grade
                                           (parity <= 1e-3), the default of the Python shell;
                           CRFP_PREC_HALF: the reduced-precision tier (north_star's "bf16" tier, <= 5e-3): same kernels and fp32
                                           storage as TC3, but every tensor-core ACTIVATION operand is rounded once to fp16
                                           (11-bit mantissa; bf16's 8 bits measured 1.1e-2) and used as ONE product against
                                           the fp16 hi / lo split weights: a third of TC3's MMAs and half its operand bytes
                                           through shared memory; weights packed with dtype fp16;
                           CRFP_PREC_BF16: experimental, L1 layers with bf16 STORAGE (not parity-certified) */
  int32_t variant;      /* CRFP_VARIANT_DSV (CRFP_DSV, v18), CRFP_VARIANT_V15 (CRFP), CRFP_VARIANT_V13 (CRFP_simple) */
  crfp_layer layer[CRFP_DSV_MAX_LAYERS];
  crfp_layer_tc layer_tc[CRFP_DSV_MAX_LAYERS]; /* entries of layers with crfp_layer_info.tc != 0 (else all NULL) */
} crfp_dsv_weights;

/*
 * Static description of layer `i` so that the host can pack a state_dict without duplicating tables:
 *   key      state_dict prefix ("dcn_0.dcn_block.0", ...), weight = key+".weight", bias = key+".bias"
 *   kind     0 conv (crfp_conv3x3_fwd packing), 1 DCN (crfp_dcn_v2_fwd packing),
 *            2 fused heads: conv packing of cat(key.weight, key2.weight) along cout (dcn_offset ++ dcn_mask)
 *   nsrc/c   the channel split of the input concat (conv packing pads each source)
 *   mode     per-source CRFP_SRC_*
 *   ci_lo    first input channel of the ORIGINAL weight this packed layer consumes (first-frame variants
 *            use only a slice of the input channels, model/CRFP.py:1637 and SURVEY.md App. A)
 */
typedef struct {
  const char* key;
  const char* key2;
  int32_t kind;
  int32_t nsrc;
  int32_t c[3];
  int32_t mode[3];
  int32_t cout;
  int32_t ci_lo;
  int32_t dg;
  int32_t thin; /* 1: consumed by the thin-channel kernels (cout <= 4): weight [9][cin_packed][4] */
  int32_t tc;   /* tensor-core packing into crfp_dsv_weights.layer_tc[i]: 0 none; 1 conv (TC3: crfp_conv3x3_tc3_fwd
                   packing of the sources with c % 8 == 0, a trailing 2-channel source becomes w_extra; BF16:
                   crfp_conv3x3_tc_fwd packing); 2 DCN (crfp_dcn_v2_tc3_fwd / crfp_dcn_v2_tc_fwd packing);
                   3 conv, TC3 only */
  int32_t _pad;
} crfp_layer_info;
int crfp_dsv_num_layers(void);
int crfp_dsv_layer_info(int i, crfp_layer_info* info);               /* CRFP_VARIANT_DSV */
int crfp_layer_info_variant(int variant, int i, crfp_layer_info* info);

typedef struct {
  int32_t n, t, h, w;     /* clip batch, frames in this call, LR size */
  int32_t mid_channels;
  int32_t _pad;
} crfp_dsv_shape;           /* (the model variant travels in crfp_dsv_weights.variant) */

/*
 * Clip-level stage for frames [0, t) of `n` clips:
 *   lrs      NCHW fp32 (n, t, 3, h, w) contiguous
 *   prev_lr  optional NCHW (n, 3, h, w): frame preceding lrs[:,0] (streaming); NULL in clip mode
 *   lr4      out: NHWC4 copy of lrs, image index b*t+i                       [n*t*h*w*4]
 *   x_lr     out: encoder_lr features NHWC 32ch, image index b*t+i           [n*t*h*w*C]
 *   flows    out: NHWC 2ch flow from frame i to frame i-1, image index b*t+i [n*t*h*w*2]; entry i=0 is
 *            written only when prev_lr != NULL (otherwise left untouched)
 */
size_t crfp_dsv_prepare_workspace(const crfp_dsv_shape* s);
int crfp_dsv_prepare(const crfp_dsv_shape* s, const crfp_dsv_weights* wts, const float* lrs, const float* prev_lr,
                     float* lr4, float* x_lr, float* flows, void* workspace, size_t ws_bytes, crfp_stream stream);

/*
 * One recurrent step for `n` clips (model/CRFP.py:1555-1684).  All per-frame inputs are passed as the pointer
 * to clip 0's frame plus the stride (in elements) between clips.
 *   first        1: no previous state (i == 0 branch), state buffers are only written
 *   lr4/x_lr/flow   this frame's slices of the crfp_dsv_prepare outputs (clip stride in floats)
 *   fvs          NCHW fp32 (3, 8h, 8w) fovea frame; mks: bool/uint8 (1, 8h, 8w)
 *   fg           optional uint8/float? NULL in clip mode (streaming regional mask, fp32 (1,8h,8w))
 *   state_hr     in/out NHWC (n, 8h, 8w, 4): feat_prop_lv3 (S)
 *   state_l1     in/out NHWC (n, 2h, 2w, 24): [feat_lv0 | feat_lv1 | feat_lv2]
 *   out          NCHW fp32 (3, 8h, 8w) per clip
 *   aux_stream   optional second stream + 3 caller-created events (cudaEvent_t, timing disabled): the work of a frame
 *                that does not sit on the L1 chain (x8 flow resize + warp of the HR state, fovea tile flags, compositing,
 *                encoder_hr, dcn_3.upsample) is forked onto it and joined back before the HR stage, so these HBM-bound
 *                kernels fill the launch gaps of the latency-bound tensor-core chain.  The fork starts after everything
 *                enqueued on `stream` before the call and everything is joined back into `stream` before the call
 *                returns (also valid under stream capture).  NULL: single-stream execution.  Results are bit-identical.
 */
typedef struct {
  crfp_dsv_shape shape;   /* t ignored */
  int32_t first;
  int32_t skip_outside_fovea; /* 1: evaluate encoder_hr/conv_tttf only on tiles whose (dilated) mask is non-empty
                                 (bit-identical) */
  const float* lr4;   long long lr4_clip_stride;
  const float* x_lr;  long long x_lr_clip_stride;
  const float* flow;  long long flow_clip_stride;
  const float* fvs;   long long fvs_clip_stride;
  const uint8_t* mks; long long mks_clip_stride;
  const float* fg;    long long fg_clip_stride;
  float* state_hr;
  float* state_l1;
  float* out;         long long out_clip_stride;
  crfp_stream aux_stream;
  void* aux_events[3];
} crfp_dsv_frame_desc;
size_t crfp_dsv_frame_workspace(const crfp_dsv_shape* s);
int crfp_dsv_frame(const crfp_dsv_frame_desc* d, const crfp_dsv_weights* wts, void* workspace, size_t ws_bytes,
                   crfp_stream stream);
size_t crfp_sizeof_dsv_weights(void);
size_t crfp_sizeof_dsv_frame_desc(void);

/* ------------------------------------------------------------------ training: backward kernels, loss, optimiser */
/*
 * Gradients the reference obtains from `loss.backward()` (trainer.py:246-250) through ATen / cuDNN autograd and
 * `_ext.dcn_v2_backward(input, weight, bias, offset, mask, grad_output, ...)` of jinfagang/DCNv2_latest (README.md:26).
 * All tensors here are DENSE fp32 NHWC (pixel stride == channel count).  Outputs documented as "accumulated" are +=
 * (zero-fill them once per step); the others are overwritten.
 */
/* g = dy * act'(v), act in {CRFP_ACT_LRELU, CRFP_ACT_RELU}; the sign of v is taken from the saved forward output */
int crfp_act_bwd(long long count, int act, const float* dy, const float* out, float* g, crfp_stream stream);
/* nn.Conv2d(3x3,s1,p1) backward for ONE source of the forward's channel concat (the concat is never materialised):
 * the source owns input channels [cin_off, cin_off + cin) of the layer's cin_total.
 * backward-data: dx[n,h,w,cin] (dense, overwritten) from g[n,h,w,cout]; weight_t = [tap][cout][cin_total], tap = ky*3+kx */
int crfp_conv3x3_bwd_data(int n, int h, int w, int cin, int cout, int cin_total, int cin_off, const float* g,
                          const float* weight_t, float* dx, crfp_stream stream);
/* backward-weight: rows [cin_off, cin_off+cin) of dw[tap][cin_total][cout] (accumulated) from x[n,h,w,cin] (dense) and
 * g[n,h,w,cout]; db[cout] (accumulated; pass it with one source only, NULL otherwise) */
int crfp_conv3x3_bwd_weight(int n, int h, int w, int cin, int cout, int cin_total, int cin_off, const float* x,
                            const float* g, float* dw, float* db, float* workspace, size_t ws_floats, crfp_stream stream);
/* optional two-stage reduction for the thin (few-channel) layers: with a workspace of at least this many floats (0 = the
 * layer is not thin, pass NULL) the pixel chunks write partial sums there and a second kernel adds them up instead of
 * contending for the same few cache lines of dw with atomics; workspace == NULL selects the atomic path */
size_t crfp_conv3x3_bwd_weight_workspace(int n, int h, int w, int cin, int cout);
/* `count` (x, g) pairs of identical shape — the t frames of the recurrence, whose weight gradients the trainer defers to the
 * end of the backward pass — accumulated into the same dw / db by one launch per 16 pairs.  xs / gs: HOST arrays of device
 * pointers.  workspace: crfp_conv3x3_bwd_weight_workspace(n * min(count, 16), h, w, cin, cout) floats (may be NULL: atomics). */
int crfp_conv3x3_bwd_weight_batched(int count, const float* const* xs, const float* const* gs, int n, int h, int w, int cin,
                                    int cout, int cin_total, int cin_off, float* dw, float* db, float* workspace,
                                    size_t ws_floats, crfp_stream stream);
/*
 * DCNv2 backward (= dcn_v2_backward): non-shared layout only (offset dg*18, mask dg*9 channels per pixel).
 *   weight   [K][cout], K = 9*c, k = (g*9+t)*(c/dg) + c_in_group (crfp_dcn_v2_fwd packing with cout % 4 == 0)
 *   dx       [n,h,w,c]      accumulated (atomic scatter)
 *   doffset  [n,h,w,dg*18], dmask [n,h,w,dg*9]   overwritten
 *   dweight  [K][cout]      accumulated;  dbias [cout] accumulated (may be NULL)
 *   col      workspace [n*h*w][K] floats: the modulated columns, rebuilt here and contracted with dout
 *   weight_t optional [cout][K] transpose of `weight`: enables the vector kernel when c == 4*dg and cout % 4 == 0
 */
/*
 * The pointwise tail of DCN_module.forward between the fused offset / mask conv and DCNv2 (model/CRFP.py:337-347), forward
 * and backward, one kernel each (NHWC fp32, npix = n*h*w):
 *   offset[.., 2k+e] = mag * tanh(heads[.., 2k'+e]) + flow[.., 1-e],  mask[.., k] = sigmoid(heads[.., noff + k'])
 * nk pairs / masks per pixel out; repeat == 0: heads has 3*nk channels (2*nk offsets then nk masks), k' = k; repeat != 0: heads
 * has 3 channels (one pair, one mask) shared by the nk taps (the HR module).  bwd overwrites dheads and dflow.
 */
/* Fovea blend of the training forward: out = lrelu(mask * f + (1 - mask) * s, 0.1), mask (npix) one value per pixel, f / s / out
 * (npix, c) NHWC fp32 with c % 4 == 0 (model/CRFP.py:1672-1675); bwd: df = g * mask, ds = g * (1 - mask), g = dout * lrelu'(out). */
int crfp_fovea_blend_fwd(long long npix, int c, const float* f, const float* s, const float* mask, float* out, crfp_stream stream);
int crfp_fovea_blend_bwd(long long npix, int c, const float* dout, const float* out, const float* mask, float* df, float* ds,
                         crfp_stream stream);
int crfp_dcn_heads_act_fwd(long long npix, int nk, int repeat, float mag, const float* heads, const float* flow, float* offset,
                           float* mask, crfp_stream stream);
int crfp_dcn_heads_act_bwd(long long npix, int nk, int repeat, float mag, const float* heads, const float* doffset,
                           const float* dmask, float* dheads, float* dflow, crfp_stream stream);
typedef struct {
  int32_t n, h, w;
  int32_t c, cout, dg;
  const float* x;
  const float* offset;
  const float* mask;
  const float* weight;
  const float* dout;
  float* dx;
  float* doffset;
  float* dmask;
  float* dweight;
  float* dbias;
  float* col;
  const float* weight_t;
  float* wg_workspace;   /* optional: crfp_dcn_v2_bwd_workspace() floats for the tiled weight-gradient kernel (NULL: atomics) */
  size_t wg_ws_floats;
} crfp_dcn_bwd_desc;
int crfp_dcn_v2_bwd(const crfp_dcn_bwd_desc* d, crfp_stream stream);
size_t crfp_dcn_v2_bwd_workspace(int n, int h, int w, int c, int cout);
size_t crfp_sizeof_dcn_bwd_desc(void);
/* flow_warp backward (zeros padding): dx [n,h,w,c] accumulated (may be NULL), dflow [n,h,w,2] overwritten (may be NULL) */
int crfp_flow_warp_bwd(int n, int h, int w, int c, const float* x, const float* flow, const float* dy, float* dx,
                       float* dflow, crfp_stream stream);
/* crfp_resize_bilinear backward: dx [n,hin,win,c] accumulated from dy [n,hout,wout,c] */
int crfp_resize_bilinear_bwd(int n, int hin, int win, int c, int hout, int wout, float rscale_h, float rscale_w, float mul,
                             const float* dy, float* dx, crfp_stream stream);
/* crfp_avgpool2 backward: dx [n,hin,win,c] overwritten from dy [n,hin/2,win/2,c] */
int crfp_avgpool2_bwd(int n, int hin, int win, int c, const float* dy, float* dx, crfp_stream stream);
/* Charbonnier loss (loss/loss.py:116-124): *loss_sum += sum sqrt((pred-target)^2 + eps) (divide by count for the mean);
 * dpred (may be NULL) = grad_scale * (pred-target)/sqrt(.) — grad_scale = loss_weight / count for reduction='mean' */
int crfp_charbonnier_fwd_bwd(long long count, const float* pred, const float* target, float eps, float grad_scale,
                             float* loss_sum, float* dpred, crfp_stream stream);
/* one torch.optim.Adam update (trainer.py:149: no weight decay / amsgrad) over a flat parameter range:
 * step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t) */
int crfp_adam_step(long long count, float* p, const float* g, float* m, float* v, float beta1, float beta2, float eps,
                   float step_size, float bc2_sqrt, crfp_stream stream);

/* ------------------------------------------------------------------ SPyNet operators (legacy flow pyramid) */
/* the pieces of SPyNet.compute_flow (model/CRFP.py:593-650) the hot path does not already provide; dense fp32 NHWC */
/* `conv(act(x))` of SPyNetBasicModule (model/CRFP.py:145-152, 687-741): k x k (odd), stride 1, pad k/2, optional ReLU on
 * the INPUT, optional residual added to the output (flow_up + module(...), :647).  weight = [k*k][cin][cout] */
int crfp_conv_kxk_fwd(int n, int h, int w, int cin, int cout, int k, int relu_in, const float* x, const float* weight,
                      const float* bias, const float* residual, float* out, crfp_stream stream);
/* F.interpolate(mode='bilinear', align_corners=True) (model/CRFP.py:635-639); out = value * mul */
int crfp_resize_bilinear_ac(int n, int hin, int win, int c, const float* in, int hout, int wout, float mul, float* out,
                            crfp_stream stream);
/* out[p][ch] = ((in[p][ch] - sub[ch]) / div[ch]) * mul[ch], ch < c_in; extra output channels [c_in, c_out) are zero:
 * the ImageNet normalisation (:609-610) into a 4-channel plane, and the per-axis flow rescale (:683-684) */
int crfp_channel_affine(long long npix, int c_in, int c_out, const float* in, const float* sub, const float* div,
                        const float* mul, float* out, crfp_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* CRFP_B200_H */
