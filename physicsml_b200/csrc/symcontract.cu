// Symmetric contraction (MACE product basis) — forward, cotangent and directional second-order kernels.
//
// Replaces  Contraction.forward  (reference src/physicsml/models/mace/modules/symmetric_contraction.py:218-239),
// which evaluates dense einsums over >97%-zero generalised-CG tensors.  Here the symmetrised sparse polynomial
//     out[b, c, M] = sum_nu sum_k W_nu[z_b, k, c] * sum_mono coef * x[b,c,i_1] ... x[b,c,i_nu]
// is unrolled at build time (gen/sym_gen.cuh); one thread owns one (node, channel), keeps x[b,c,:] in registers and
// streams the element-dependent weights W[z, k, c] coalesced along c.  Compute-bound (fp32 FMA), not an HBM kernel.
//
// Compiled once per contraction signature: -DPML_SYM_ID=<id>.
#include "pml_common.cuh"

// The generated bodies are straight-line code of 4 k (fwd) to 30 k (dbl) instructions, far beyond the instruction
// caches (ncu: stall_no_instruction 8.5 cycles per issue, issue slots 15-33 % busy).  One CTA per SM whose warps
// re-converge at a barrier after every weight keeps all warps of the SM inside the same few cache lines, so the
// instruction stream is fetched once per SM instead of once per warp.
#ifndef PML_SYM_BLOCK
#define PML_SYM_BLOCK 256
#endif
#ifndef PML_SYM_SYNC
#define PML_SYM_SYNC __syncthreads();
#endif
#include "gen/sym_gen.cuh"

#ifndef PML_SYM_ID
#error "compile with -DPML_SYM_ID=<spec id>"
#endif

#define PML_SYM_MAX_OUT 3
#define PML_SYM_MAX_NU 3

extern "C" {
// Weight tables of one SymmetricContraction: w[a][nu-1] -> [Z, K[a][nu-1], C] row-major (reference parameter layout:
// weights_max = nu = corr, weights[i] = nu = corr-1-i; symmetric_contraction.py:153-156,209-213).
typedef struct {
  const float* w[PML_SYM_MAX_OUT][PML_SYM_MAX_NU];
  float* gw[PML_SYM_MAX_OUT][PML_SYM_MAX_NU];  // gradient accumulators (same shapes), may be null
  int K[PML_SYM_MAX_OUT][PML_SYM_MAX_NU];
} pml_sym_weights;
}

namespace pml {

// Element lookup: node_attrs rows are one-hot in every reference configuration; a general (dense) row falls back to
// the literal definition  W_eff[k,c] = sum_e attrs[b,e] W[e,k,c]   (einsum "...,ekc,be->...", symmetric_contraction.py:125-129).
struct NodeElem {
  int z;  // >= 0: one-hot index; -1: dense row
  const float* row;
  int Z;
  __device__ __forceinline__ NodeElem(const float* attrs, int b, int Z_) : row(attrs + (size_t)b * Z_), Z(Z_) {
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
    z = (nz == 1 && unit) ? last : -1;
  }
};

struct WLoad {
  const pml_sym_weights& t;
  const NodeElem& el;
  int C, c;
  __device__ __forceinline__ float load(int a, int nu, int k) const {
    const float* base = t.w[a][nu];
    const int K = t.K[a][nu];
    if (el.z >= 0) return __ldg(base + ((size_t)el.z * K + k) * C + c);
    float s = 0.f;
    for (int e = 0; e < el.Z; ++e) {
      const float y = __ldg(el.row + e);
      if (y != 0.f) s = fmaf(y, __ldg(base + ((size_t)e * K + k) * C + c), s);
    }
    return s;
  }
};

struct WGrad {
  const pml_sym_weights& t;
  const NodeElem& el;
  int C, c;
  bool active;
  __device__ __forceinline__ void add(int a, int nu, int k, float v) const {
    float* base = t.gw[a][nu];
    if (!active || base == nullptr) return;
    const int K = t.K[a][nu];
    if (el.z >= 0) {
      atomicAdd(base + ((size_t)el.z * K + k) * C + c, v);
      return;
    }
    for (int e = 0; e < el.Z; ++e) {
      const float y = __ldg(el.row + e);
      if (y != 0.f) atomicAdd(base + ((size_t)e * K + k) * C + c, y * v);
    }
  }
};

template <int ID>
__global__ void __launch_bounds__(PML_SYM_BLOCK) sym_fwd_kernel(const float* __restrict__ x, const float* __restrict__ attrs,
                                                      pml_sym_weights W, float* __restrict__ out, int N, int C, int Z) {
  pdl_enter();
  using T = SymCode<ID>;
  // no early exit: every thread of the CTA takes part in the barriers of the generated body
  const long long tot = (long long)N * C, idx_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx_raw < tot;
  const long long idx = active ? idx_raw : tot - 1;
  const int b = (int)(idx / C), c = (int)(idx - (long long)b * C);
  float xv[T::S], o[T::DOUT];
#pragma unroll
  for (int i = 0; i < T::S; ++i) xv[i] = __ldg(x + (size_t)idx * T::S + i);
#pragma unroll
  for (int i = 0; i < T::DOUT; ++i) o[i] = 0.f;
  const NodeElem el(attrs, b, Z);
  const WLoad wl{W, el, C, c};
  T::fwd(xv, wl, o);
  if (active) T::store_out(out + (size_t)b * T::DOUT * C, C, c, o);
}

template <int ID, bool WANT_W>
__global__ void __launch_bounds__(PML_SYM_BLOCK, 3) sym_bwd_kernel(const float* __restrict__ x, const float* __restrict__ attrs,
                                                      pml_sym_weights W, const float* __restrict__ g,
                                                      float* __restrict__ dx, int N, int C, int Z) {
  pdl_enter();
  using T = SymCode<ID>;
  // no early exit: every thread of the CTA takes part in the barriers of the generated body
  const long long tot = (long long)N * C, idx_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx_raw < tot;
  const long long idx = active ? idx_raw : tot - 1;
  const int b = (int)(idx / C), c = (int)(idx - (long long)b * C);
  float xv[T::S], gv[T::DOUT], d[T::S];
#pragma unroll
  for (int i = 0; i < T::S; ++i) {
    xv[i] = __ldg(x + (size_t)idx * T::S + i);
    d[i] = 0.f;
  }
  T::load_out(g + (size_t)b * T::DOUT * C, C, c, gv);
  const NodeElem el(attrs, b, Z);
  const WLoad wl{W, el, C, c};
  const WGrad wg{W, el, C, c, active};
  T::template bwd<WANT_W>(xv, wl, gv, d, wg);
  if (dx != nullptr && active) {
#pragma unroll
    for (int i = 0; i < T::S; ++i) dx[(size_t)idx * T::S + i] = d[i];
  }
}

// Gradient of  phi'(x, W, g; v) = d/de <g, out(x + e v, W)>  w.r.t. (g, x, W): JVP, Hessian-vector product and the
// weight cotangent that the double-backward pass of force-matching training needs (reference computes it through
// generic autograd over the einsum graph; lightning/module.py:39-40 sets create_graph=True).
template <int ID, bool WANT_W>
__global__ void __launch_bounds__(PML_SYM_BLOCK, 3) sym_dbl_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                      const float* __restrict__ attrs, pml_sym_weights W,
                                                      const float* __restrict__ g, float* __restrict__ dg,
                                                      float* __restrict__ ddx, int N, int C, int Z) {
  pdl_enter();
  using T = SymCode<ID>;
  // no early exit: every thread of the CTA takes part in the barriers of the generated body
  const long long tot = (long long)N * C, idx_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx_raw < tot;
  const long long idx = active ? idx_raw : tot - 1;
  const int b = (int)(idx / C), c = (int)(idx - (long long)b * C);
  float xv[T::S], vv[T::S], gv[T::DOUT], o[T::DOUT], d[T::S];
#pragma unroll
  for (int i = 0; i < T::S; ++i) {
    xv[i] = __ldg(x + (size_t)idx * T::S + i);
    vv[i] = __ldg(v + (size_t)idx * T::S + i);
    d[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < T::DOUT; ++i) o[i] = 0.f;
  T::load_out(g + (size_t)b * T::DOUT * C, C, c, gv);
  const NodeElem el(attrs, b, Z);
  const WLoad wl{W, el, C, c};
  const WGrad wg{W, el, C, c, active};
  T::template dbl<WANT_W>(xv, vv, wl, gv, o, d, wg);
  if (dg != nullptr && active) T::store_out(dg + (size_t)b * T::DOUT * C, C, c, o);
  if (ddx != nullptr && active) {
#pragma unroll
    for (int i = 0; i < T::S; ++i) ddx[(size_t)idx * T::S + i] = d[i];
  }
}

}  // namespace pml

#define PML_CAT_(a, b) a##b
#define PML_CAT(a, b) PML_CAT_(a, b)

static inline bool pml_sym_has_gw(const pml_sym_weights& W) {
  for (int a = 0; a < PML_SYM_MAX_OUT; ++a)
    for (int n = 0; n < PML_SYM_MAX_NU; ++n)
      if (W.gw[a][n] != nullptr) return true;
  return false;
}

extern "C" int PML_CAT(pml_sym_fwd_launch_, PML_SYM_ID)(const float* x, const float* attrs, const pml_sym_weights* W,
                                                         float* out, int N, int C, int Z, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  const long long tot = (long long)N * C;
  pml::launch_plain(pml::sym_fwd_kernel<PML_SYM_ID>, dim3(pml::ceil_div(tot, PML_SYM_BLOCK)), dim3(PML_SYM_BLOCK), 0, stream, x, attrs, *W, out, N, C, Z);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// gradient accumulators W->gw[][] (if any) must be zero-initialised by the caller.
extern "C" int PML_CAT(pml_sym_bwd_launch_, PML_SYM_ID)(const float* x, const float* attrs, const pml_sym_weights* W,
                                                         const float* g, float* dx, int N, int C, int Z,
                                                         cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  const long long tot = (long long)N * C;
  if (pml_sym_has_gw(*W))
    pml::launch_plain(pml::sym_bwd_kernel<PML_SYM_ID, true>, dim3(pml::ceil_div(tot, PML_SYM_BLOCK)), dim3(PML_SYM_BLOCK), 0, stream, x, attrs, *W, g, dx, N, C, Z);
  else
    pml::launch_plain(pml::sym_bwd_kernel<PML_SYM_ID, false>, dim3(pml::ceil_div(tot, PML_SYM_BLOCK)), dim3(PML_SYM_BLOCK), 0, stream, x, attrs, *W, g, dx, N, C, Z);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int PML_CAT(pml_sym_dbl_launch_, PML_SYM_ID)(const float* x, const float* v, const float* attrs,
                                                         const pml_sym_weights* W, const float* g, float* dg,
                                                         float* ddx, int N, int C, int Z, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  const long long tot = (long long)N * C;
  if (pml_sym_has_gw(*W))
    pml::launch_plain(pml::sym_dbl_kernel<PML_SYM_ID, true>, dim3(pml::ceil_div(tot, PML_SYM_BLOCK)), dim3(PML_SYM_BLOCK), 0, stream, x, v, attrs, *W, g, dg, ddx, N, C, Z);
  else
    pml::launch_plain(pml::sym_dbl_kernel<PML_SYM_ID, false>, dim3(pml::ceil_div(tot, PML_SYM_BLOCK)), dim3(PML_SYM_BLOCK), 0, stream, x, v, attrs, *W, g, dg, ddx, N, C, Z);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
