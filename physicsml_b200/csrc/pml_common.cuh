// Shared helpers for the physicsml_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PML_OK 0
#define PML_ERR_BAD_ARG (-1)
#define PML_ERR_UNSUPPORTED (-2)
#define PML_ERR_LAUNCH (-3)

// Every entry point returns 0 or a negative error code and never throws; launch errors are reported
// (without synchronising) via cudaGetLastError, which also CLEARS a non-sticky launch error so that one failed launch
// does not make every later entry point report PML_ERR_LAUNCH; CUDA-graph capture keeps working.
#define PML_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
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

// packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2): one instruction, two channels
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2dup(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }

// *(float2*)p += (a, b) without a return value (p 8-byte aligned), with an L2 eviction policy (createpolicy)
__device__ __forceinline__ void red_add_f32x2(float* p, float a, float b, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
}

__device__ __forceinline__ void st_stream2(float* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y));
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncSetAttribute / SM counts are per device: launchers cache their one-time configuration per device ordinal
constexpr int kMaxDevices = 64;
static inline int device_slot() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

}  // namespace pml
