// Shared helpers for the physicsml_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PML_OK 0
#define PML_ERR_BAD_ARG (-1)
#define PML_ERR_UNSUPPORTED (-2)
#define PML_ERR_LAUNCH (-3)

// Every entry point returns 0 or a negative error code and never throws; launch errors are reported
// (without synchronising) via cudaPeekAtLastError so that CUDA-graph capture keeps working.
#define PML_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                 \
    if (e__ != cudaSuccess) return PML_ERR_LAUNCH;           \
  } while (0)

namespace pml {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (read-once) global loads: keep them out of L1 so gathered node rows stay cached
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace pml
