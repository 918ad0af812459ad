// Fused multi-tensor Adam (amsgrad, L2 weight decay) — SURVEY §8f row N3.  The reference trains with
// torch.optim.Adam(lr, amsgrad=True[, weight_decay]) (models/mace/supervised/default_configs.py:25-33,
// mace_module.py:131-157); for the ~60 small parameter tensors of a MACE / NequIP model torch's capturable foreach
// implementation launches ~10 kernels per step (0.32 ms of a 9.7 ms config-2 step).  Here one kernel updates every
// tensor: a chunk table maps CTAs to (tensor, offset).  Same update rule as torch.optim.adam._single_tensor_adam:
//   g += wd p ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; vmax = max(vmax, v)
//   p -= lr / (1 - b1^t) * m / (sqrt(vmax) / sqrt(1 - b2^t) + eps)
// The step counter lives in device memory so that a captured CUDA graph advances it on every replay.
#include "pml_common.cuh"

extern "C" {
typedef struct {
  float* p;        // parameter
  const float* g;  // gradient
  float* m;        // exp_avg
  float* v;        // exp_avg_sq
  float* vmax;     // max_exp_avg_sq
  long long n;     // elements
} pml_adam_tensor;
typedef struct {
  int tensor;
  int pad;
  long long offset;
} pml_adam_chunk;
}

namespace pml {

constexpr int kAdamChunk = 4096;

__global__ void adam_step_kernel(float* step) { *step += 1.f; }

__global__ void __launch_bounds__(256) adam_amsgrad_kernel(const pml_adam_tensor* __restrict__ tensors,
                                                           const pml_adam_chunk* __restrict__ chunks,
                                                           const float* __restrict__ step, float lr, float b1, float b2,
                                                           float omb1, float omb2, float eps, float wd) {
  const pml_adam_chunk ch = chunks[blockIdx.x];
  const pml_adam_tensor t = tensors[ch.tensor];
  const float s = *step;
  const float bc1 = 1.f - powf(b1, s), bc2 = 1.f - powf(b2, s);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const long long end = min(t.n, ch.offset + (long long)kAdamChunk);
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    float p = t.p[i];
    const float g = t.g[i] + wd * p;
    const float m = b1 * t.m[i] + omb1 * g;
    const float v = b2 * t.v[i] + omb2 * g * g;
    const float vm = fmaxf(t.vmax[i], v);
    t.m[i] = m;
    t.v[i] = v;
    t.vmax[i] = vm;
    p -= step_size * m / (sqrtf(vm) * inv_sqrt_bc2 + eps);
    t.p[i] = p;
  }
}

}  // namespace pml

// tensors / chunks: DEVICE tables (nchunks entries of kAdamChunk = 4096 elements each); step: device scalar, incremented
// first (one extra 1-thread launch).  one_minus_b1 / one_minus_b2 are passed separately because 1 - (float)0.999 is not
// (float)(1 - 0.999).  Everything on the caller's stream.
extern "C" int pml_adam_amsgrad(const pml_adam_tensor* tensors, const pml_adam_chunk* chunks, int nchunks, float* step,
                                float lr, float b1, float b2, float one_minus_b1, float one_minus_b2, float eps, float wd,
                                cudaStream_t stream) {
  if (nchunks <= 0) return PML_OK;
  pml::adam_step_kernel<<<1, 1, 0, stream>>>(step);
  pml::adam_amsgrad_kernel<<<nchunks, 256, 0, stream>>>(tensors, chunks, step, lr, b1, b2, one_minus_b1, one_minus_b2, eps, wd);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
