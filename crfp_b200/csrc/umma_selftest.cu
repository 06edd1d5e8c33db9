// Self-test of the hand-written tcgen05 plumbing (descriptors, TMEM alloc/ld, mbarrier commit) and of the
// "shifted window" addressing the implicit-GEMM convolution relies on:
//   D[m][n] = sum_k A[m + shift][k] * B[n][k],  m < 128, A staged once as [K/8][rowsA] 16-byte records.
#include "common.cuh"
#include "umma.cuh"

namespace crfp {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const __nv_bfloat16* __restrict__ A,
                                                            const __nv_bfloat16* __restrict__ B, float* __restrict__ D,
                                                            int rowsA, int K, int N, int shift, int sbo_recs) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint4* sA = reinterpret_cast<uint4*>(smem);               // [K/8][rowsA]
  uint4* sB = sA + (size_t)(K / 8) * rowsA;                 // [K/8][N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (K / 8) * rowsA; i += 128) {
    const int kc = i / rowsA, r = i - kc * rowsA;
    sA[i] = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kc * 8);
  }
  for (int i = tid; i < (K / 8) * N; i += 128) {
    const int kc = i / N, r = i - kc * N;
    sB[i] = *reinterpret_cast<const uint4*>(B + (size_t)r * K + kc * 8);
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(128, N);
    const uint32_t lboA = (uint32_t)rowsA * 16, lboB = (uint32_t)N * 16;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t da = umma::make_desc(umma::smem_u32(sA) + (uint32_t)(2 * ks) * lboA + (uint32_t)shift * 16, lboA,
                                          (uint32_t)sbo_recs * 16);
      const uint64_t db = umma::make_desc(umma::smem_u32(sB) + (uint32_t)(2 * ks) * lboB, lboB, 128);
      umma::mma_bf16(taddr, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    umma::tmem_ld32(taddr + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, v);
    for (int i = 0; i < 32; ++i)
      if (c0 + i < N) D[(size_t)(32 * warp + lane) * N + c0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, ncols);
}

}  // namespace crfp

extern "C" int crfp_selftest_umma(int rowsA, int K, int N, int shift, const void* A, const void* B, float* D,
                                  crfp_stream stream) {
  using namespace crfp;
  if (!A || !B || !D) return CRFP_ERR_NULL;
  if (K % 16 || N % 16 || N < 16 || N > 256 || rowsA < 128 + shift || shift < 0) return CRFP_ERR_BAD_SHAPE;
  const size_t smem = (size_t)(K / 8) * (rowsA + N) * 16;
  if (smem > 200 * 1024) return CRFP_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, rowsA,
                                                              K, N, shift, 8);
  return check_launch();
}

// Same with an explicit stride between the 8-row groups of the A operand (SBO, in 16-byte records): row m of the MMA reads
// record shift + (m / 8) * sbo_recs + m % 8.  sbo_recs = 10 is how the fused align kernel addresses an 8-pixel-wide tile
// inside a 10-pixel-wide halo tile (one descriptor start address per 3x3 tap).
extern "C" int crfp_selftest_umma_sbo(int rowsA, int K, int N, int shift, int sbo_recs, const void* A, const void* B, float* D,
                                      crfp_stream stream) {
  using namespace crfp;
  if (!A || !B || !D) return CRFP_ERR_NULL;
  if (K % 16 || N % 16 || N < 16 || N > 256 || shift < 0 || sbo_recs < 8 || rowsA < shift + 15 * sbo_recs + 8) return CRFP_ERR_BAD_SHAPE;
  const size_t smem = (size_t)(K / 8) * (rowsA + N) * 16;
  if (smem > 200 * 1024) return CRFP_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { note_cuda_error(e); return CRFP_ERR_CUDA; }
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, rowsA,
                                                              K, N, shift, sbo_recs);
  return check_launch();
}

// ---- micro-benchmark: cycles per tcgen05.mma (M128 x N x K16, bf16, no-swizzle K-major operands) issued back to
// back by one thread with precomputed descriptors; used to size the conv kernels' tiles.
namespace crfp {
__global__ void __launch_bounds__(128) umma_rate_kernel(int N, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint4* s = reinterpret_cast<uint4*>(smem);
  for (int i = tid; i < (2 * 136 + 2 * 256) ; i += 128) s[i] = make_uint4(0u, 0u, 0u, 0u);
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base, ncols);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t taddr = tmem_base;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_bf16(128, N);
    const uint64_t da0 = umma::make_desc(umma::smem_u32(s), 136 * 16, 128);
    const uint64_t da1 = umma::make_desc(umma::smem_u32(s) + 16, 136 * 16, 128);
    const uint64_t db = umma::make_desc(umma::smem_u32(s + 2 * 136), (uint32_t)N * 16, 128);
    t0 = clock64();
    for (int r = 0; r < reps; r += 2) {
      umma::mma_bf16(taddr, da0, db, idesc, 1u);
      umma::mma_bf16(taddr, da1, db, idesc, 1u);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(taddr, ncols);
}
}  // namespace crfp

extern "C" int crfp_selftest_umma_rate(int N, int reps, int ctas, long long* cycles_out, crfp_stream stream) {
  using namespace crfp;
  if (!cycles_out || N % 16 || N < 16 || N > 256 || reps < 2 || ctas < 1) return CRFP_ERR_BAD_SHAPE;
  const size_t smem = (size_t)(2 * 136 + 2 * 256) * 16;
  umma_rate_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(N, reps, cycles_out);
  return check_launch();
}
