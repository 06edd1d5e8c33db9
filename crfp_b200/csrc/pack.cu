// Device-side weight packing for the training step: the optimiser changes every weight every iteration, so every layer's
// tensor-core operands (crfp_conv3x3_tc3_fwd: hi / lo bf16 [ntiles][9][kc][nt][8] + padded bias + fp32 weights of the
// 2-channel extra source) are rebuilt once per step.  The torch formulation (crfp_b200/packing.py pack_conv_tc3: index,
// permute, two casts, a subtraction ...) is a dozen tiny kernels per layer; this is ONE launch per layer and also packs the
// backward-data operator (the transposed, 180-degree-rotated kernel of nn.Conv2d's autograd) straight from the OIHW
// parameter, without materialising the transposed weight.
#include "common.cuh"

namespace crfp {

struct PackTc3 {
  const float* weight;     // OIHW fp32 [cout_w][cin_w][3][3]
  const float* bias;       // [nout] or NULL (zeros)
  int cout_w, cin_w;
  int transposed;          // 0: V[o][i][tap] = W[o][lo + i][tap];  1 (backward data): V[o][i][tap] = W[i][lo + o][8 - tap]
  int lo;
  int k_lo;                // transposed only: first row (output channel of W) of this K slice
  int nout, k, extra;      // logical outputs, tensor-core inputs, trailing fp32 inputs
  int nt, ntiles, kc;      // kc = K chunks of 8 incl. the pad chunk
  __nv_bfloat16* hi;
  __nv_bfloat16* lo_out;
  float* bias_p;           // [ntiles*nt]
  float* wx;               // [9][extra][ntiles*nt] or NULL
};

__device__ __forceinline__ float pack_v(const PackTc3& P, int o, int i, int tap) {
  if (P.transposed) return P.weight[((size_t)(P.k_lo + i) * P.cin_w + P.lo + o) * 9 + (8 - tap)];
  return P.weight[((size_t)o * P.cin_w + P.lo + i) * 9 + tap];
}

__global__ void __launch_bounds__(256) pack_conv_tc3_kernel(const PackTc3 P) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nmain = P.ntiles * 9 * P.kc * P.nt * 8;
  const int npad = P.ntiles * P.nt;
  if (idx < nmain) {
    const int j = idx & 7;
    int r = idx >> 3;
    const int n = r % P.nt; r /= P.nt;
    const int kc = r % P.kc; r /= P.kc;
    const int tap = r % 9;
    const int tile = r / 9;
    const int o = tile * P.nt + n, i = kc * 8 + j;
    const float v = (o < P.nout && i < P.k) ? pack_v(P, o, i, tap) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    P.hi[idx] = h;
    P.lo_out[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    return;
  }
  int e = idx - nmain;
  if (e < npad) {
    P.bias_p[e] = (P.bias != nullptr && e < P.nout) ? P.bias[e] : 0.f;
    return;
  }
  e -= npad;
  if (P.wx != nullptr && e < 9 * P.extra * npad) {
    const int o = e % npad, x = (e / npad) % P.extra, tap = e / (npad * P.extra);
    P.wx[e] = o < P.nout ? pack_v(P, o, P.k + x, tap) : 0.f;
  }
}

}  // namespace crfp

using namespace crfp;

extern "C" int crfp_pack_conv_tc3(const float* weight, const float* bias, int cout_w, int cin_w, int transposed, int lo, int nout,
                                  int k, int extra, int k_lo, void* w_hi, void* w_lo, float* bias_packed, float* w_extra,
                                  crfp_stream stream) {
  if (!weight || !w_hi || !w_lo || !bias_packed) return CRFP_ERR_NULL;
  if (cout_w <= 0 || cin_w <= 0 || nout <= 0 || k <= 0 || k % 8 != 0 || extra < 0 || lo < 0) return CRFP_ERR_BAD_SHAPE;
  if (extra > 0 && !w_extra) return CRFP_ERR_NULL;
  if (transposed) {
    if (extra != 0 || k_lo < 0 || k_lo + k > cout_w || lo + nout > cin_w) return CRFP_ERR_BAD_SHAPE;
  } else {
    if (nout != cout_w || k_lo != 0 || lo + k + extra > cin_w) return CRFP_ERR_BAD_SHAPE;
  }
  int32_t nt = 0, ntiles = 0;
  CRFP_TRY(crfp_tc3_cout_tile(nout, k, &nt, &ntiles));
  PackTc3 P;
  P.weight = weight; P.bias = bias; P.cout_w = cout_w; P.cin_w = cin_w; P.transposed = transposed; P.lo = lo; P.k_lo = k_lo;
  P.nout = nout; P.k = k; P.extra = extra; P.nt = nt; P.ntiles = ntiles;
  P.kc = k / 8 + (k / 8) % 2;
  P.hi = (__nv_bfloat16*)w_hi; P.lo_out = (__nv_bfloat16*)w_lo; P.bias_p = bias_packed; P.wx = extra ? w_extra : nullptr;
  const long long total = (long long)ntiles * 9 * P.kc * nt * 8 + (long long)ntiles * nt * (1 + 9 * extra);
  pack_conv_tc3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P);
  return check_launch();
}
