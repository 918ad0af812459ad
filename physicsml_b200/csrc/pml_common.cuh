// Shared helpers for the physicsml_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

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

// Programmatic dependent launch.  A training step on a molecule batch is ~460 kernels of a few microseconds: the gap between
// two dependent kernels (grid drain, completion, next grid's launch and CTA scheduling) is a sizeable part of each.
// A kernel that starts with pdl_enter() and is launched by launch_pdl() may be made resident while its predecessor still
// runs; it blocks in griddepcontrol.wait until every predecessor grid has completed and flushed its memory, so the
// ordering a stream gives is unchanged - only the launch latency overlaps.  launch_dependents comes FIRST so that the
// successor of this kernel can be set up as early as possible; nothing that allocates a shared resource another grid
// may be waiting for (TMEM) happens before the wait.  Both instructions are no-ops in a kernel launched normally.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// OFF unless PML_PDL=1 is in the environment: the effect is inside the run-to-run spread (see launch_pdl below)
static inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PML_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

template <bool PDL, class... KArgs, class... Args>
static inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (PDL && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// The small, frequent kernels (general GEMM, activations, row / column permutes: ~160 launches of a config-2 step) carry
// the attribute when PML_PDL=1.  Measured: back-to-back node-sized GEMMs 11.7 -> 11.2 us each; the config-2 step 4.40 ->
// 4.36 ms in one A/B pair and 4.43 -> 4.48 ms in another (two boxes) - no effect outside the +-1 % run-to-run spread,
// hence off by default.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  return launch_ex<true>(kernel, grid, block, smem, stream, args...);
}
// The large kernels (fused edge kernels, streaming GEMMs, contraction, element kernels) are NOT: an early-resident
// successor holds shared memory / TMEM / CTA slots the running kernels of both streams want (measured with the attribute on
// every kernel: config 2 4.40 -> 4.44 ms, config 4 32.3 -> 33.0 ms, config 5 16.5 -> 16.7 ms).  They keep pdl_enter(),
// a no-op for them, so that flipping one is a one-word change.
template <class... KArgs, class... Args>
static inline cudaError_t launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  return launch_ex<false>(kernel, grid, block, smem, stream, args...);
}

// cudaFuncSetAttribute / SM counts are per device: launchers cache their one-time configuration per device ordinal
constexpr int kMaxDevices = 64;
static inline int device_slot() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

}  // namespace pml
