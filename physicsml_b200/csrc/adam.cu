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
// per-tensor hyper-parameters: one launch serves every param group (the reference builds ONE GROUP PER PARAMETER with
// its own weight_decay, mace_module.py:131-157) and every tensor keeps its own step counter like torch.optim.Adam
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* vmax;
  long long n;
  const float* step;  // device scalar, already incremented for this step
  float lr, b1, b2, omb1, omb2, eps, wd, gscale;  // gscale multiplies the gradient first (1/world_size of a summed all-reduce)
} pml_adam_tensor2;
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

__global__ void adam_steps_kernel(float* const* __restrict__ steps, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) *steps[i] += 1.f;
}

__global__ void __launch_bounds__(256) adam_amsgrad_groups_kernel(const pml_adam_tensor2* __restrict__ tensors,
                                                                  const pml_adam_chunk* __restrict__ chunks) {
  const pml_adam_chunk ch = chunks[blockIdx.x];
  const pml_adam_tensor2 t = tensors[ch.tensor];
  const float s = *t.step;
  const float bc1 = 1.f - powf(t.b1, s), bc2 = 1.f - powf(t.b2, s);
  const float step_size = t.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const long long end = min(t.n, ch.offset + (long long)kAdamChunk);
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    float p = t.p[i];
    const float g = t.g[i] * t.gscale + t.wd * p;
    const float m = t.b1 * t.m[i] + t.omb1 * g;
    const float v = t.b2 * t.v[i] + t.omb2 * g * g;
    const float vm = fmaxf(t.vmax[i], v);
    t.m[i] = m;
    t.v[i] = v;
    t.vmax[i] = vm;
    p -= step_size * m / (sqrtf(vm) * inv_sqrt_bc2 + t.eps);
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

// Every param group in one launch: `steps` is a DEVICE table of the nsteps distinct step-counter pointers the tensors
// refer to (incremented first by one small launch), hyper-parameters travel in the tensor table.
extern "C" int pml_adam_amsgrad_groups(const pml_adam_tensor2* tensors, const pml_adam_chunk* chunks, int nchunks,
                                       float* const* steps, int nsteps, cudaStream_t stream) {
  if (nchunks <= 0) return PML_OK;
  if (nsteps > 0) pml::adam_steps_kernel<<<(nsteps + 127) / 128, 128, 0, stream>>>(steps, nsteps);
  pml::adam_amsgrad_groups_kernel<<<nchunks, 256, 0, stream>>>(tensors, chunks);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// ---- gradient packing for the data-parallel all-reduce ---------------------------------------------------------------------
// flat[t.offset + i] = t.src ? t.src[i] : 0 for every tensor of a bucket in ONE launch (chunk table as above), so that one
// NCCL all-reduce per bucket replaces torch.cat + per-tensor copy_ (the reference leaves this to Lightning DDP,
// lightning/model.py:159-163).  A NULL src (a parameter without gradient on this rank) packs zeros: every rank reduces
// the same layout.
extern "C" {
typedef struct {
  const float* src;
  long long offset;  // element offset into flat
  long long n;
} pml_pack_tensor;
}
namespace pml {
__global__ void __launch_bounds__(256) pack_f32_kernel(const pml_pack_tensor* __restrict__ tensors,
                                                       const pml_adam_chunk* __restrict__ chunks, float* __restrict__ flat) {
  const pml_adam_chunk ch = chunks[blockIdx.x];
  const pml_pack_tensor t = tensors[ch.tensor];
  const long long end = min(t.n, ch.offset + (long long)kAdamChunk);
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) flat[t.offset + i] = t.src ? t.src[i] : 0.f;
}
}  // namespace pml
extern "C" int pml_pack_f32(const pml_pack_tensor* tensors, const pml_adam_chunk* chunks, int nchunks, float* flat,
                            cudaStream_t stream) {
  if (nchunks <= 0) return PML_OK;
  pml::pack_f32_kernel<<<nchunks, 256, 0, stream>>>(tensors, chunks, flat);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
