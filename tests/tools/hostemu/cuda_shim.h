// TEST INFRASTRUCTURE ONLY — a serial host emulation of the CUDA execution model, just large enough to compile
// crfp_b200/csrc/bwd.cu (sync-free, shared-memory-free SIMT kernels: one thread per element + atomicAdd) with g++
// and run the SAME kernel bodies and the SAME C-ABI entry points on host pointers.  This container has no GPU; the
// emulation lets the `-m "not gpu"` tests check the arithmetic, the indexing and the argument validation of the
// backward kernels against torch autograd before they ever reach a B200.  Nothing under crfp_b200/ uses it: the
// product path loads only the nvcc-built libcrfp_b200.so and raises when that is missing.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../../include/crfp_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef void* cudaStream_t;
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

static thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }
static inline float4 atomicAdd(float4* p, float4 v) {  // sm_90+ vector atomic (red.global.add.v4.f32)
  const float4 o = *p; p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w; return o;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }

namespace crfp {
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int check_launch() { return CRFP_OK; }
inline int sm_count() { return 148; }

template <class F>
inline void emu_run(dim3 grid, dim3 block, F f) {
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              blockIdx = dim3(bx, by, bz);
              threadIdx = dim3(tx, ty, tz);
              f();
            }
}
}  // namespace crfp

#define CRFP_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != CRFP_OK) return _s; \
  } while (0)

#define CRFP_LAUNCH(kernel, grid, block, st, ...) \
  do { (void)(st); crfp::emu_run((grid), (block), [&]() { kernel(__VA_ARGS__); }); } while (0)
