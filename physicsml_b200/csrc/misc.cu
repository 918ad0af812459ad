// Small elementwise / segmented kernels on the path: normalised SiLU (and tanh) with first and second derivative
// forms for the radial MLP and readout (e3nn nn.FullyConnectedNet / normalize2mom; reference blocks.py:157-160,11-28),
// the per-graph pooled readout (mace.py:252-258) and the receiver-CSR build used by the fused edge kernels.
#include <stdio.h>

#include "gemm_problem.h"
#include "pml_common.cuh"

namespace pml {

// kind 0: SiLU, 1: tanh.   order 0: c*f(x) ; 1: g*c*f'(x) ; 2: g*v*c*f''(x)
template <int KIND, int ORDER>
__device__ __forceinline__ float act_value(float t, float c, float g, float v) {
  float r;
  if (KIND == 0) {
    const float s = 1.f / (1.f + expf(-t));
    if (ORDER == 0) r = t * s;
    else if (ORDER == 1) r = s * (1.f + t * (1.f - s));
    else r = s * (1.f - s) * (2.f + t * (1.f - 2.f * s));
  } else {
    const float th = tanhf(t);
    if (ORDER == 0) r = th;
    else if (ORDER == 1) r = 1.f - th * th;
    else r = -2.f * th * (1.f - th * th);
  }
  r *= c;
  if (ORDER >= 1) r *= g;
  if (ORDER >= 2) r *= v;
  return r;
}

template <int KIND, int ORDER>
__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                  const float* __restrict__ v, float c, float* __restrict__ y,
                                                  long long n) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = act_value<KIND, ORDER>(x[i], c, ORDER >= 1 ? g[i] : 0.f, ORDER >= 2 ? v[i] : 0.f);
}

// 16 bytes per thread and operand: the radial-MLP activations are edge-sized streams (config 4: 14 launches over
// 2 M x 64 values; the scalar kernel above ran them at 2.4 TB/s)
template <int KIND, int ORDER>
__global__ void __launch_bounds__(256) act_kernel_v4(const float4* __restrict__ x, const float4* __restrict__ g,
                                                     const float4* __restrict__ v, float c, float4* __restrict__ y,
                                                     long long n4) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 t = __ldcs(x + i);
  float4 gg = make_float4(0.f, 0.f, 0.f, 0.f), vv = gg, r;
  if (ORDER >= 1) gg = __ldcs(g + i);
  if (ORDER >= 2) vv = __ldcs(v + i);
  r.x = act_value<KIND, ORDER>(t.x, c, gg.x, vv.x);
  r.y = act_value<KIND, ORDER>(t.y, c, gg.y, vv.y);
  r.z = act_value<KIND, ORDER>(t.z, c, gg.z, vv.z);
  r.w = act_value<KIND, ORDER>(t.w, c, gg.w, vv.w);
  y[i] = r;
}

// out[batch[n], f] += x[n, f]   (batch sorted ascending; G tiny) — one thread per (n, f)
__global__ void __launch_bounds__(256) pool_sum_kernel(const float* __restrict__ x, const long long* __restrict__ batch,
                                                       float* __restrict__ out, int N, int F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * F) return;
  const int n = (int)(i / F), f = (int)(i - (long long)n * F);
  atomicAdd(out + (size_t)batch[n] * F + f, x[i]);
}

__global__ void __launch_bounds__(256) pool_gather_kernel(const float* __restrict__ g, const long long* __restrict__ batch,
                                                          float* __restrict__ out, int N, int F) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * F) return;
  const int n = (int)(i / F), f = (int)(i - (long long)n * F);
  out[i] = g[(size_t)batch[n] * F + f];
}

// ---- receiver-CSR build: histogram -> (scan on the caller's side or single-block scan) -> stable fill ----------------
__global__ void __launch_bounds__(256) csr_count_kernel(const long long* __restrict__ idx, int* __restrict__ counts, int E) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) atomicAdd(counts + (int)idx[e], 1);
}

// single-CTA exclusive scan (N up to a few hundred thousand is fine: one pass of 1024-wide chunks)
__global__ void __launch_bounds__(1024) csr_scan_kernel(const int* __restrict__ counts, int* __restrict__ rowptr, int N) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < N ? counts[i] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      warp_tot[threadIdx.x] = w;
    }
    __syncthreads();
    const int woff = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    const int incl = carry + woff + s;
    if (i < N) rowptr[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) rowptr[N] = carry;
}

// int64 [2,E] edge_index -> int32 sender / receiver copies
__global__ void __launch_bounds__(256) split_edge_index_kernel(const long long* __restrict__ ei, int* __restrict__ a,
                                                               int* __restrict__ b, int E) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) {
    a[e] = (int)ei[e];
    b[e] = (int)ei[(size_t)E + e];
  }
}

// [N, C*S] irreps layout (blocks [C, 2l+1]) <-> [N, C, S] channel-major layout (reference irreps_tools.py:47-67)
__global__ void __launch_bounds__(256) irreps_to_channel_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                long long total, int C, int S, int lmax, int inverse) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes the channel-major tensor: (n, c, s)
  const int s = (int)(i % S);
  const long long nc = i / S;
  const int c = (int)(nc % C);
  const long long n = nc / C;
  int l = 0;
  while ((l + 1) * (l + 1) <= s) ++l;
  (void)lmax;
  const int m = s - l * l, d = 2 * l + 1;
  const long long j = n * (long long)C * S + (long long)C * l * l + (long long)c * d + m;
  if (inverse) y[j] = x[i];
  else y[i] = x[j];
}

// test utility: fill the shared memory of every SM with NaN, so that a kernel reading shared memory it never wrote (the
// residue of whatever ran before) turns a rare, tiny discrepancy into a NaN
__global__ void poison_smem_kernel(int words) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = __int_as_float(0x7fc00000);
  __syncthreads();
  if (sm[(threadIdx.x * 31) % words] == 0.f) printf("unreachable\n");
}

}  // namespace pml

extern "C" int pml_debug_poison_smem(cudaStream_t stream) {
  int dev = 0, sms = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  static bool configured[pml::kMaxDevices] = {};
  const int slot = pml::device_slot();
  if (!configured[slot]) {
    if (cudaFuncSetAttribute(pml::poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin) != cudaSuccess)
      return PML_ERR_LAUNCH;
    configured[slot] = true;
  }
  pml::poison_smem_kernel<<<sms, 256, max_optin, stream>>>(max_optin / 4);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_act(const float* x, const float* g, const float* v, float c, int kind, int order, float* y,
                       long long n, cudaStream_t stream) {
  if (n <= 0) return PML_OK;
  const bool vec = (n % 4 == 0) && !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                      reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(v)) & 15);
  const int grid = pml::ceil_div(vec ? n / 4 : n, 256);
#define PML_ACT_CASE(K, O)                                                                                          \
  if (kind == K && order == O) {                                                                                    \
    if (vec)                                                                                                        \
      pml::launch_pdl(pml::act_kernel_v4<K, O>, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const float4*>(x), \
                      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(v), c,                    \
                      reinterpret_cast<float4*>(y), n / 4);                                                         \
    else                                                                                                            \
      pml::launch_pdl(pml::act_kernel<K, O>, dim3(grid), dim3(256), 0, stream, x, g, v, c, y, n);                   \
  }
  PML_ACT_CASE(0, 0) PML_ACT_CASE(0, 1) PML_ACT_CASE(0, 2) PML_ACT_CASE(1, 0) PML_ACT_CASE(1, 1) PML_ACT_CASE(1, 2)
#undef PML_ACT_CASE
  if (kind < 0 || kind > 1 || order < 0 || order > 2) return PML_ERR_BAD_ARG;
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// out [G, F] must be zero-initialised
extern "C" int pml_pool_sum(const float* x, const long long* batch, float* out, int N, int F, cudaStream_t stream) {
  if (N <= 0 || F <= 0) return PML_OK;
  pml::pool_sum_kernel<<<pml::ceil_div((long long)N * F, 256), 256, 0, stream>>>(x, batch, out, N, F);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_pool_gather(const float* g, const long long* batch, float* out, int N, int F, cudaStream_t stream) {
  if (N <= 0 || F <= 0) return PML_OK;
  pml::pool_gather_kernel<<<pml::ceil_div((long long)N * F, 256), 256, 0, stream>>>(g, batch, out, N, F);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// dst[n, j] = src[n, colmap[j]]: re-orders the columns of a node-feature row.  Used to hand the grouped GEMM a COMPONENT-MAJOR
// copy (m*mul + u inside every irrep block) of an operand stored in the e3nn order (u*(2l+1) + m) or the channel layout
// ([C, S]): the reduction index of the per-irrep mixing then has unit stride, which is the difference between 16-byte
// coalesced operand loads and one 32-byte sector per element (config 5: 254 -> ~40 us for the cotangent of a 256-channel
// o3.Linear).  Rows are a few KB and L2-resident; the pass costs a few microseconds.
namespace pml {
__global__ void __launch_bounds__(256) permute_columns_kernel(const float* __restrict__ src, long long src_row,
                                                              const int* __restrict__ colmap, float* __restrict__ dst,
                                                              int dst_row, long long total) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / dst_row;
  const int j = (int)(i - n * dst_row);
  dst[i] = __ldg(src + n * src_row + __ldg(colmap + j));
}
}  // namespace pml

extern "C" int pml_permute_columns(const float* src, long long src_row, const int* colmap, float* dst, int dst_row, long long N,
                                   cudaStream_t stream) {
  const long long total = N * dst_row;
  if (total <= 0) return PML_OK;
  pml::launch_pdl(pml::permute_columns_kernel, dim3(pml::ceil_div(total, 256)), dim3(256), 0, stream, src, src_row, colmap, dst, dst_row, total);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// ---- species-sorted node order for the element-indexed linears (element_linear.cu's header explains the op) ----------
// FullyConnectedTensorProduct(x, one-hot attrs) is "one [mul_in, mul_out] matrix per chemical element": with the nodes
// GROUPED by element every (element, irrep) pair is a dense GEMM over a contiguous row range, which is what the tensor
// cores want (the FFMA kernels of element_linear.cu reach ~6 % of the fp32 pipe on a batch of molecules).
//   order[pos]  = node ids grouped by element, ascending inside a group (deterministic: a stable counting sort)
//   seg[0..Z]   = start of every element's group in `order`; seg[Z + 1] = 1 when some attrs row is not one-hot
// One CTA: N is the number of nodes of a batch (1e3 .. 1e5); Z <= 128.
namespace pml {
constexpr int kSpeciesMaxZ = 128;

__device__ __forceinline__ int one_hot_species(const float* __restrict__ row, int Z) {
  int nz = 0, last = -1;
  bool unit = true;
  for (int e = 0; e < Z; ++e) {
    const float y = __ldg(row + e);
    if (y != 0.f) {
      ++nz;
      last = e;
      unit = unit && (y == 1.f);
    }
  }
  return (nz == 1 && unit) ? last : -1;
}

__global__ void __launch_bounds__(1024) species_order_kernel(const float* __restrict__ attrs, int N, int Z,
                                                             int* __restrict__ order, int* __restrict__ seg) {
  // group Z collects the nodes whose row is not one-hot: they go to the tail of `order`, which stays a permutation
  __shared__ int cnt[kSpeciesMaxZ + 1], base[kSpeciesMaxZ + 1], run[kSpeciesMaxZ + 1];
  __shared__ int wcnt[32][kSpeciesMaxZ + 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = Z + 1;
  if (tid < G) cnt[tid] = 0, run[tid] = 0;
  __syncthreads();
  for (int n = tid; n < N; n += blockDim.x) {
    const int s = one_hot_species(attrs + (size_t)n * Z, Z);
    atomicAdd(&cnt[s < 0 ? Z : s], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int v = 0; v < G; ++v) {
      base[v] = acc;
      if (v <= Z) seg[v] = acc;  // seg[Z] = number of one-hot rows
      acc += cnt[v];
    }
    seg[Z + 1] = cnt[Z] > 0 ? 1 : 0;
  }
  const int chunks = (N + (int)blockDim.x - 1) / (int)blockDim.x;
  for (int c = 0; c < chunks; ++c) {
    __syncthreads();  // base / run of the previous chunk are final; wcnt is free
    for (int i = tid; i < 32 * G; i += blockDim.x) wcnt[i / G][i % G] = 0;
    __syncthreads();
    const int n = c * (int)blockDim.x + tid;
    int s = -1;
    if (n < N) {
      s = one_hot_species(attrs + (size_t)n * Z, Z);
      if (s < 0) s = Z;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, s);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (s >= 0 && rank == 0) wcnt[warp][s] = __popc(peers);
    __syncthreads();
    if (s >= 0) {
      int off = 0;
      for (int w = 0; w < warp; ++w) off += wcnt[w][s];
      order[base[s] + run[s] + off + rank] = n;
    }
    __syncthreads();
    if (tid < G) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += wcnt[w][tid];
      run[tid] += tot;
    }
  }
}

// The per-batch GEMM problem tables of the element-indexed linears: `tmpl` holds one problem per (path, element) with the
// element-independent fields filled in by the host plan; here the row range of the element's group is patched in
//   mode 0: M = group size, rows start at seg[v]  ;  mode 1: the reduction (chunk 0) runs over the group
// entirely on the device, so a captured CUDA graph stays valid when a new batch is loaded into its static buffers.
// A batch whose attrs are not one-hot (seg[Z + 1]) poisons every problem with alpha = NaN instead of computing a wrong result.
__global__ void segment_problems_kernel(const pml_gemm_problem* __restrict__ tmpl, const pml_seg_meta* __restrict__ meta,
                                        int nprob, const int* __restrict__ seg, int Z, pml_gemm_problem* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nprob) return;
  pml_gemm_problem p = tmpl[i];
  const pml_seg_meta m = meta[i];
  const long long start = seg[m.v];
  const int count = seg[m.v + 1] - seg[m.v];
  for (int c = 0; c < p.nchunks; ++c) {
    p.a_off[c] += start * m.a_mult;
    p.b_off[c] += start * m.b_mult;
  }
  p.c_off += start * m.c_mult;
  if (m.mode == 0) p.M = count;
  else p.kc[0] = count;
  if (seg[Z + 1] != 0) p.alpha = __int_as_float(0x7fc00000);
  out[i] = p;
}

// dst[pos, j] = src[rows[pos], colmap[j]]   (scatter == 0: gather rows, permute columns; dst rows are dense)
// dst[rows[pos], colmap[j]] = src[pos, j]   (scatter == 1: the inverse; src rows are dense)
__global__ void __launch_bounds__(256) permute_rows_cols_kernel(const float* __restrict__ src, long long src_row,
                                                                const int* __restrict__ rows, const int* __restrict__ colmap,
                                                                float* __restrict__ dst, long long dst_row, int ncols,
                                                                long long total, int scatter) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long pos = i / ncols;
  const int j = (int)(i - pos * ncols);
  const long long r = __ldg(rows + pos);
  const int cj = __ldg(colmap + j);
  if (scatter) dst[r * dst_row + cj] = __ldg(src + pos * src_row + j);
  else dst[pos * dst_row + j] = __ldg(src + r * src_row + cj);
}
}  // namespace pml

extern "C" int pml_species_order(const float* attrs, int N, int Z, int* order, int* seg, cudaStream_t stream) {
  if (Z <= 0 || Z > pml::kSpeciesMaxZ || N < 0) return PML_ERR_UNSUPPORTED;
  pml::species_order_kernel<<<1, 1024, 0, stream>>>(attrs, N, Z, order, seg);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_segment_problems(const pml_gemm_problem* tmpl, const pml_seg_meta* meta, int nprob, const int* seg, int Z,
                                    pml_gemm_problem* out, cudaStream_t stream) {
  if (nprob <= 0) return PML_OK;
  pml::segment_problems_kernel<<<pml::ceil_div(nprob, 128), 128, 0, stream>>>(tmpl, meta, nprob, seg, Z, out);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_permute_rows_cols(const float* src, long long src_row, const int* rows, const int* colmap, float* dst,
                                     long long dst_row, int ncols, long long N, int scatter, cudaStream_t stream) {
  const long long total = N * ncols;
  if (total <= 0) return PML_OK;
  pml::launch_pdl(pml::permute_rows_cols_kernel, dim3(pml::ceil_div(total, 256)), dim3(256), 0, stream, src, src_row, rows, colmap,
                  dst, dst_row, ncols, total, scatter);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// Fused force-matching loss (reference lightning/module.py:65-95 + mace_module.py:159-185 with the stock
// MSELoss on y_graph_scalars and y_node_vector):  loss = w_e mean((e - e_t)^2) + w_f mean((-gpos - f_t)^2), forces = -gpos.
// One CTA, fixed summation order: deterministic (the eager chain is ~8 kernels forward and ~8 backward).
namespace pml {
__global__ void __launch_bounds__(1024) ef_mse_fwd_kernel(const float* __restrict__ e, const float* __restrict__ e_t, int ne,
                                                          const float* __restrict__ gpos, const float* __restrict__ f_t, int nf,
                                                          float w_e, float w_f, float* __restrict__ out) {
  float se = 0.f, sf = 0.f;
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    const float d = e[i] - e_t[i];
    se = fmaf(d, d, se);
  }
  for (int i = threadIdx.x; i < nf; i += blockDim.x) {
    const float d = -gpos[i] - f_t[i];
    sf = fmaf(d, d, sf);
  }
  __shared__ float red[2][32];
  se = warp_sum(se), sf = warp_sum(sf);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = se, red[1][threadIdx.x >> 5] = sf;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    se = threadIdx.x < nw ? red[0][threadIdx.x] : 0.f;
    sf = threadIdx.x < nw ? red[1][threadIdx.x] : 0.f;
    se = warp_sum(se), sf = warp_sum(sf);
    if (threadIdx.x == 0) out[0] = w_e * se / (float)max(ne, 1) + w_f * sf / (float)max(nf, 1);
  }
}
__global__ void __launch_bounds__(256) ef_mse_bwd_kernel(const float* __restrict__ e, const float* __restrict__ e_t, int ne,
                                                         const float* __restrict__ gpos, const float* __restrict__ f_t, int nf,
                                                         float w_e, float w_f, const float* __restrict__ gl,
                                                         float* __restrict__ de, float* __restrict__ dgpos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float g = gl[0];
  if (i < ne) de[i] = g * w_e * 2.f * (e[i] - e_t[i]) / (float)ne;
  const int j = i - ne;
  if (j >= 0 && j < nf) dgpos[j] = g * w_f * 2.f * (gpos[j] + f_t[j]) / (float)nf;
}
}  // namespace pml

extern "C" int pml_ef_mse_fwd(const float* e, const float* e_t, int ne, const float* gpos, const float* f_t, int nf, float w_e,
                              float w_f, float* out, cudaStream_t stream) {
  pml::ef_mse_fwd_kernel<<<1, 1024, 0, stream>>>(e, e_t, ne, gpos, f_t, nf, w_e, w_f, out);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_ef_mse_bwd(const float* e, const float* e_t, int ne, const float* gpos, const float* f_t, int nf, float w_e,
                              float w_f, const float* gl, float* de, float* dgpos, cudaStream_t stream) {
  if (ne + nf <= 0) return PML_OK;
  pml::ef_mse_bwd_kernel<<<pml::ceil_div(ne + nf, 256), 256, 0, stream>>>(e, e_t, ne, gpos, f_t, nf, w_e, w_f, gl, de, dgpos);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// Verlet-skin test of the MD loop (md.py): *out = max_i |pos_i - ref_i|^2 (out zero-initialised by the caller; the
// squared displacement is non-negative, so its IEEE bits order like an unsigned integer)
namespace pml {
__global__ void __launch_bounds__(256) max_disp2_kernel(const float* __restrict__ pos, const float* __restrict__ ref, int N,
                                                        unsigned int* __restrict__ out) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const float dx = pos[3 * i] - ref[3 * i], dy = pos[3 * i + 1] - ref[3 * i + 1], dz = pos[3 * i + 2] - ref[3 * i + 2];
    m = fmaxf(m, dx * dx + dy * dy + dz * dz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}
}  // namespace pml

extern "C" int pml_max_disp2(const float* pos, const float* ref, int N, float* out, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  int grid = pml::ceil_div(N, 256);
  if (grid > 592) grid = 592;
  pml::max_disp2_kernel<<<grid, 256, 0, stream>>>(pos, ref, N, reinterpret_cast<unsigned int*>(out));
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// perm[rowptr[n] + k] = id of the k-th edge (ascending id = stable order) whose index is n.  Step 1 fills every segment
// in arrival order with an atomic cursor; step 2 sorts each segment by edge id in shared memory (one CTA per node), which
// makes the grouping — and with it the summation order of the scatter kernels — deterministic.  Segments longer than
// 2048 take an in-place odd-even transposition sort in global memory (slow, rare, still exact).
namespace pml {
__global__ void __launch_bounds__(256) csr_group_fill_kernel(const long long* __restrict__ idx, const int* __restrict__ rowptr,
                                                             int* __restrict__ cursor, int* __restrict__ perm, int E) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int n = (int)idx[e];
  perm[rowptr[n] + atomicAdd(cursor + n, 1)] = e;
}
constexpr int kGroupSortMax = 2048;
__global__ void __launch_bounds__(128) csr_group_sort_kernel(int* __restrict__ perm, const int* __restrict__ rowptr, int N) {
  __shared__ int s[kGroupSortMax];
  const int beg = rowptr[blockIdx.x], len = rowptr[blockIdx.x + 1] - beg;
  if (len <= 1) return;
  if (len > kGroupSortMax) {
    int* a = perm + beg;
    for (int phase = 0; phase < len; ++phase) {
      for (int t = 2 * threadIdx.x + (phase & 1); t + 1 < len; t += 2 * blockDim.x) {
        const int x = a[t], y = a[t + 1];
        if (x > y) a[t] = y, a[t + 1] = x;
      }
      __syncthreads();
    }
    return;
  }
  int P = 2;
  while (P < len) P <<= 1;
  for (int t = threadIdx.x; t < P; t += blockDim.x) s[t] = t < len ? perm[beg + t] : 0x7fffffff;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += blockDim.x) {
        const int u = t ^ j;
        if (u > t) {
          const bool up = (t & k) == 0;
          const int a = s[t], b = s[u];
          if ((a > b) == up) s[t] = b, s[u] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < len; t += blockDim.x) perm[beg + t] = s[t];
}
}  // namespace pml

// cursor: scratch [N], zero-initialised by the caller; perm [E]
extern "C" int pml_csr_group(const long long* idx, const int* rowptr, int* cursor, int* perm, int E, int N,
                             cudaStream_t stream) {
  if (E <= 0 || N <= 0) return PML_OK;
  pml::csr_group_fill_kernel<<<pml::ceil_div(E, 256), 256, 0, stream>>>(idx, rowptr, cursor, perm, E);
  pml::csr_group_sort_kernel<<<N, 128, 0, stream>>>(perm, rowptr, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// rowptr [N+1]; counts is scratch [N], zero-initialised by the caller
extern "C" int pml_csr_rowptr(const long long* idx, int* counts, int* rowptr, int E, int N, cudaStream_t stream) {
  if (E > 0) pml::csr_count_kernel<<<pml::ceil_div(E, 256), 256, 0, stream>>>(idx, counts, E);
  pml::csr_scan_kernel<<<1, 1024, 0, stream>>>(counts, rowptr, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// exclusive scan of int32 counts[n] into out[n+1] (single CTA)
extern "C" int pml_scan_i32(const int* counts, int* out, int n, cudaStream_t stream) {
  pml::csr_scan_kernel<<<1, 1024, 0, stream>>>(counts, out, n);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_split_edge_index(const long long* edge_index, int* row0, int* row1, int E, cudaStream_t stream) {
  if (E <= 0) return PML_OK;
  pml::split_edge_index_kernel<<<pml::ceil_div(E, 256), 256, 0, stream>>>(edge_index, row0, row1, E);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_irreps_to_channel(const float* x, float* y, long long N, int C, int lmax, int inverse,
                                     cudaStream_t stream) {
  const int S = (lmax + 1) * (lmax + 1);
  const long long total = N * C * S;
  if (total <= 0) return PML_OK;
  pml::irreps_to_channel_kernel<<<pml::ceil_div(total, 256), 256, 0, stream>>>(x, y, total, C, S, lmax, inverse);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
