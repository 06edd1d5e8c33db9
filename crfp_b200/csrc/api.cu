// C-ABI plumbing: status strings, launch accounting, device check, public conv entry point.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace crfp {

static thread_local cudaError_t g_last_err = cudaSuccess;
static thread_local long long g_launches = 0;

void note_cuda_error(cudaError_t e) { g_last_err = e; }
bool pdl_enabled(bool persistent) {
  static const int mode = [] {   // 0 none, 1 persistent tensor-core kernels only, 2 all
    if (getenv("CRFP_NO_PDL") != nullptr) return 0;
    const char* m = getenv("CRFP_PDL");
    if (m == nullptr) return 1;
    return !strcmp(m, "all") ? 2 : !strcmp(m, "none") ? 0 : 1;
  }();
  return mode == 2 || (mode == 1 && persistent);
}
void count_launch() { ++g_launches; }

static int pad4(int c) { return (c + 3) & ~3; }

int conv_params_from_desc(const crfp_conv_desc* d, ConvParams* p) {
  memset(p, 0, sizeof(*p));
  if (d->n < 0 || d->h <= 0 || d->w <= 0 || d->nsrc < 1 || d->nsrc > 3 || d->cout <= 0) return CRFP_ERR_BAD_SHAPE;
  if (!d->weight || !d->bias) return CRFP_ERR_NULL;
  p->n = d->n; p->h = d->h; p->w = d->w; p->nsrc = d->nsrc;
  int q = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const crfp_src& sr = d->src[s];
    if (!sr.ptr) return CRFP_ERR_NULL;
    if (sr.c <= 0 || sr.cstride <= 0 || sr.coffset < 0) return CRFP_ERR_BAD_SHAPE;
    if (sr.mode == CRFP_SRC_UNSHUFFLE4) {
      if (sr.c % 64 != 0 || ((sr.cstride | sr.coffset) & 3)) return CRFP_ERR_UNSUPPORTED;
      if (sr.coffset + sr.c / 16 > sr.cstride) return CRFP_ERR_BAD_SHAPE;
    } else {
      if (sr.mode != CRFP_SRC_PLAIN) return CRFP_ERR_UNSUPPORTED;
      if (sr.coffset + sr.c > sr.cstride) return CRFP_ERR_BAD_SHAPE;
    }
    if (((uintptr_t)sr.ptr & 15) && (((sr.cstride | sr.coffset) & 3) == 0)) return CRFP_ERR_BAD_SHAPE;
    p->src[s] = sr.ptr; p->src_c[s] = sr.c; p->src_cstride[s] = sr.cstride; p->src_coffset[s] = sr.coffset;
    p->src_mode[s] = sr.mode;
    p->qstart[s] = q;
    q += pad4(sr.c) / 4;
  }
  for (int s = d->nsrc; s < 4; ++s) p->qstart[s] = q;
  p->cin_packed = ((q * 4) + 7) & ~7;
  p->cout = d->cout;
  p->cout_packed = crfp_conv_cout_packed(d->cout);
  p->act = d->act;
  p->weight = d->weight; p->bias = d->bias;
  p->out_mode = d->out_mode; p->shuffle_r = d->shuffle_r;
  if (d->out_mode == CRFP_OUT_SHUFFLE) {
    if (d->shuffle_r < 1 || d->cout % (d->shuffle_r * d->shuffle_r) != 0 || d->ndst != 1) return CRFP_ERR_BAD_SHAPE;
  } else if (d->out_mode != CRFP_OUT_NHWC) {
    return CRFP_ERR_UNSUPPORTED;
  }
  if (d->ndst < 1 || d->ndst > 2) return CRFP_ERR_BAD_SHAPE;
  p->ndst = d->ndst;
  int csum = 0;
  for (int s = 0; s < d->ndst; ++s) {
    if (!d->dst[s].ptr) return CRFP_ERR_NULL;
    p->dst[s] = d->dst[s].ptr; p->dst_c[s] = d->dst[s].c; p->dst_cstride[s] = d->dst[s].cstride;
    p->dst_coffset[s] = d->dst[s].coffset;
    csum += d->dst[s].c;
  }
  if (d->out_mode == CRFP_OUT_NHWC && csum < d->cout) return CRFP_ERR_BAD_SHAPE;
  p->residual = d->residual; p->res_cstride = d->res_cstride; p->res_coffset = d->res_coffset;
  p->flow = d->flow; p->head_split = d->head_split;
  p->post_scale = d->post_scale; p->head_mag = d->head_mag;
  if (d->act == CRFP_ACT_DCN_HEAD && !d->flow) return CRFP_ERR_NULL;
  if (d->act < CRFP_ACT_NONE || d->act > CRFP_ACT_TANH256) return CRFP_ERR_UNSUPPORTED;
  p->epi = EPI_STD;
  return CRFP_OK;
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_abi_version(void) { return CRFP_ABI_VERSION; }

extern "C" const char* crfp_status_string(int s) {
  switch (s) {
    case CRFP_OK: return "ok";
    case CRFP_ERR_BAD_SHAPE: return "bad shape";
    case CRFP_ERR_UNSUPPORTED: return "unsupported configuration";
    case CRFP_ERR_WORKSPACE: return "workspace too small";
    case CRFP_ERR_CUDA: return "CUDA error";
    case CRFP_ERR_NULL: return "null pointer";
    default: return "unknown status";
  }
}

extern "C" const char* crfp_last_cuda_error(void) { return cudaGetErrorString(g_last_err); }
extern "C" long long crfp_launch_count(void) { return g_launches; }
extern "C" void crfp_launch_count_reset(void) { g_launches = 0; }
extern "C" void crfp_launch_count_add(long long n) { g_launches += n; }

extern "C" int crfp_check_device(void) {
  // cudaDeviceGetAttribute, not cudaGetDeviceProperties: the latter takes 2.5-3.7 ms per call on the B200 box (measured,
  // profiles/r02/r2i_stream_host.txt) and used to run once per forward — half of a streaming call's latency
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  return (major == 10) ? CRFP_OK : CRFP_ERR_UNSUPPORTED;
}

extern "C" int crfp_conv_cin_packed(int nsrc, const int32_t* c) {
  if (nsrc < 1 || nsrc > 3 || !c) return CRFP_ERR_BAD_SHAPE;
  int q = 0;
  for (int s = 0; s < nsrc; ++s) q += (c[s] + 3) / 4;
  return ((q * 4) + 7) & ~7;
}

extern "C" int crfp_conv_cout_packed(int cout) { return cout <= 4 ? 4 : ((cout + 31) & ~31); }

extern "C" int crfp_conv3x3_fwd(const crfp_conv_desc* d, crfp_stream stream) {
  if (!d) return CRFP_ERR_NULL;
  ConvParams p;
  CRFP_TRY(conv_params_from_desc(d, &p));
  if ((long long)p.n * p.h * p.w == 0) return CRFP_OK;
  return launch_conv(p, (cudaStream_t)stream);
}

extern "C" size_t crfp_sizeof_conv_desc(void) { return sizeof(crfp_conv_desc); }
extern "C" size_t crfp_sizeof_warp_desc(void) { return sizeof(crfp_warp_desc); }
