// Element-indexed per-irrep linear maps: o3.FullyConnectedTensorProduct(x, node_attrs) with scalar (one-hot) attrs.
//
// Replaces  mix_layer (reference mace/modules/blocks.py:183-188,231-232), residual_connection_layer (blocks.py:72-78,88-92)
// and NequIP's self_connection_layer (nequip/modules/interaction_block.py:73-82,91-92).  The reference contracts a dense
// [mul_in, Z, mul_out] weight with the one-hot row, wasting a factor Z of work; here the element row selects the
// [mul_in, mul_out] slab directly (dense attrs rows are still handled, term by term, for generality):
//     out[n, w, m] = alpha * sum_v attrs[n, v] * sum_u W[u, v, w] * x[n, u, m]
#include "pml_common.cuh"

#define PML_EL_MAX_PATHS 8

extern "C" {
typedef struct {
  int npaths;
  int d[PML_EL_MAX_PATHS];            // 2l+1
  int mul_in[PML_EL_MAX_PATHS];
  int mul_out[PML_EL_MAX_PATHS];
  long long w_off[PML_EL_MAX_PATHS];  // offset of this path's [mul_in, Z, mul_out] slab in the flat weight
  long long x_off[PML_EL_MAX_PATHS];  // x[n, x_off + u*d + m]
  long long o_off[PML_EL_MAX_PATHS];  // out[n, o_off + w*o_ws + m]
  long long o_ws[PML_EL_MAX_PATHS];
  float alpha[PML_EL_MAX_PATHS];
  long long x_rs, o_rs;               // row strides
} pml_el_plan;
}

namespace pml {

__device__ __forceinline__ int one_hot_index(const float* __restrict__ row, int Z) {
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

template <int D>
__global__ void __launch_bounds__(128) el_fwd_kernel(const float* __restrict__ x, const float* __restrict__ attrs,
                                                     const float* __restrict__ W, pml_el_plan P, float* __restrict__ out,
                                                     int N, int Z, int accumulate) {
  pdl_enter();
  const int n = blockIdx.x, p = blockIdx.y;
  if (P.d[p] != D) return;
  extern __shared__ float xs[];  // [mul_in * D]
  const int MI = P.mul_in[p], MO = P.mul_out[p];
  const float* xr = x + (size_t)n * P.x_rs + P.x_off[p];
  for (int i = threadIdx.x; i < MI * D; i += blockDim.x) xs[i] = __ldg(xr + i);
  __syncthreads();
  const float* arow = attrs + (size_t)n * Z;
  const int z = one_hot_index(arow, Z);
  const float* Wp = W + P.w_off[p];
  for (int w = threadIdx.x; w < MO; w += blockDim.x) {
    float acc[D];
#pragma unroll
    for (int m = 0; m < D; ++m) acc[m] = 0.f;
    for (int v = (z >= 0 ? z : 0); v < (z >= 0 ? z + 1 : Z); ++v) {
      const float a = z >= 0 ? 1.f : __ldg(arow + v);
      if (a == 0.f) continue;
      for (int u = 0; u < MI; ++u) {
        const float wv = a * __ldg(Wp + ((size_t)u * Z + v) * MO + w);
#pragma unroll
        for (int m = 0; m < D; ++m) acc[m] = fmaf(wv, xs[u * D + m], acc[m]);
      }
    }
    float* o = out + (size_t)n * P.o_rs + P.o_off[p] + (size_t)w * P.o_ws[p];
#pragma unroll
    for (int m = 0; m < D; ++m) {
      const float r = P.alpha[p] * acc[m];
      o[m] = accumulate ? o[m] + r : r;
    }
  }
}

// dx[n, u, m] = alpha * sum_v a_v sum_w W[u, v, w] g[n, w, m]
// One CTA per (path, element v, chunk of 256 nodes): the chunk's nodes that carry element v are compacted
// (deterministic ballot scan) and the [mul_in, mul_out] slab of that element is staged ONCE in shared memory (coalesced
// along w, odd pitch) for all of them — one CTA per node, as in the forward, would re-read the 64 KB slab per node with
// every lane on its own cache line (ncu: 7.4 ms of a 86 ms config-4 evaluation).  Thread = input channel u, EL_XNB nodes
// at a time with their g rows broadcast from shared memory.  Nodes whose attrs row is not one-hot appear in several
// element lists, so results are accumulated with atomics into a zero-initialised dx (exactly one addend per entry, hence
// still deterministic, for one-hot rows).
constexpr int EL_XCHUNK = 256;

template <int D>
__global__ void __launch_bounds__(256) el_bwd_x_kernel(const float* __restrict__ g, const float* __restrict__ attrs,
                                                       const float* __restrict__ W, pml_el_plan P, float* __restrict__ dx,
                                                       int N, int Z, int WC, int chunk) {
  pdl_enter();
  constexpr int NB = D >= 5 ? 4 : 8;  // nodes per thread and pass
  constexpr int DP = D == 1 ? 1 : (D == 3 ? 4 : 8);  // staged row width: 128-bit shared-memory loads of the g rows
  const int p = blockIdx.z / Z, v = blockIdx.z - p * Z;
  if (P.d[p] != D) return;
  const int MI = P.mul_in[p], MO = P.mul_out[p];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_list[EL_XCHUNK];
  __shared__ float s_coef[EL_XCHUNK];
  __shared__ int s_wcnt[8];
  extern __shared__ float dyn[];
  const int wc = min(WC, MO);                // columns of the slab staged at a time (all of them when they fit)
  const int pitch = wc | 1;                  // odd: thread u reads Ws[u][w] without bank conflicts
  float* Ws = dyn;                           // [MI][pitch]
  float* gs = dyn + (((size_t)MI * pitch + 3) & ~(size_t)3);  // [NB * nsub][wc * DP], 16-byte aligned

  // `chunk` <= EL_XCHUNK nodes per CTA: a batch of molecules (N ~ 1e3) is cut finer than a large structure so that
  // its (path, element, chunk) groups cover all SMs (config 2: 96 busy CTAs of 54 nodes at 256 -> 336 of 14 at 64)
  const int n = blockIdx.y * chunk + tid;
  const float a = (tid < chunk && n < N) ? __ldg(attrs + (size_t)n * Z + v) : 0.f;
  const unsigned bal = __ballot_sync(0xffffffffu, a != 0.f);
  if (lane == 0) s_wcnt[warp] = __popc(bal);
  __syncthreads();
  int off = 0, total = 0;
#pragma unroll
  for (int w8 = 0; w8 < 8; ++w8) {
    if (w8 < warp) off += s_wcnt[w8];
    total += s_wcnt[w8];
  }
  if (total == 0) return;  // uniform
  if (a != 0.f) {
    const int pos = off + __popc(bal & ((1u << lane) - 1u));
    s_list[pos] = n;
    s_coef[pos] = a;
  }
  // thread = (input channel u, node sub-group): with mul_in < 256 the spare threads take further nodes of the pass
  const int lanes_u = MI >= 256 ? 256 : MI;
  const int nsub = 256 / lanes_u;
  const int u0 = tid % lanes_u, sub = tid / lanes_u;
  const int pass = NB * nsub;
  const float* Wp = W + P.w_off[p];
  // the reduction over w is cut into slab pieces of wc columns; every piece adds its share into dx (same thread, same
  // order every time, so the result stays deterministic for one-hot attrs)
  for (int w0 = 0; w0 < MO; w0 += wc) {
    const int wn = min(wc, MO - w0);
    __syncthreads();  // the previous piece (and the node list) is no longer being read / is complete
    for (int i = tid; i < MI * wn; i += 256) {
      const int u = i / wn, w = i - u * wn;
      Ws[u * pitch + w] = __ldg(Wp + ((size_t)u * Z + v) * MO + w0 + w);
    }
    for (int j0 = 0; j0 < total; j0 += pass) {
      const int nb = min(pass, total - j0);
      __syncthreads();
      for (int i = tid; i < nb * wn * DP; i += 256) {
        const int j = i / (wn * DP), rem = i - j * (wn * DP);
        const int w = rem / DP, m = rem - w * DP;
        gs[(j * wc + w) * DP + m] =
            m < D ? __ldg(g + (size_t)s_list[j0 + j] * P.o_rs + P.o_off[p] + (size_t)(w0 + w) * P.o_ws[p] + m) : 0.f;
      }
      __syncthreads();
      if (sub < nsub && sub * NB < nb) {
        const float* gsub = gs + (size_t)sub * NB * wc * DP;
        for (int u = u0; u < MI; u += lanes_u) {
          float acc[NB][D];
#pragma unroll
          for (int j = 0; j < NB; ++j)
#pragma unroll
            for (int m = 0; m < D; ++m) acc[j][m] = 0.f;
          const float* wr = Ws + u * pitch;
          for (int w = 0; w < wn; ++w) {
            const float wv = wr[w];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              float gv[DP];
              if (DP == 1) {
                gv[0] = gsub[j * wc + w];
              } else {
#pragma unroll
                for (int c = 0; c < DP / 4; ++c) {
                  const float4 t = reinterpret_cast<const float4*>(gsub + (size_t)(j * wc + w) * DP)[c];
                  gv[4 * c] = t.x, gv[4 * c + 1] = t.y, gv[4 * c + 2] = t.z, gv[4 * c + 3] = t.w;
                }
              }
#pragma unroll
              for (int m = 0; m < D; ++m) acc[j][m] = fmaf(wv, gv[m], acc[j][m]);
            }
          }
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const int jj = sub * NB + j;
            if (jj < nb) {
              float* o = dx + (size_t)s_list[j0 + jj] * P.x_rs + P.x_off[p] + (size_t)u * D;
              const float c = P.alpha[p] * s_coef[j0 + jj];
#pragma unroll
              for (int m = 0; m < D; ++m) atomicAdd(o + m, c * acc[j][m]);
            }
          }
        }
      }
    }
  }
}

// dW[u, v, w] = alpha * sum_n a[n, v] sum_m x[n, u, m] g[n, w, m]
// One CTA per (path, element v, 16x16 tile of (u, w), chunk of 256 nodes): the chunk's nodes that carry element v are
// compacted (deterministic ballot scan), then processed 8 at a time from shared memory.  Chunk partials are combined
// with fp32 atomics (dW zero-initialised by the caller).
constexpr int EL_CHUNK = 256, EL_NB = 8;

template <int D>
__global__ void __launch_bounds__(256) el_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ attrs, pml_el_plan P,
                                                       float* __restrict__ dW, int N, int Z) {
  pdl_enter();
  const int p = blockIdx.z / Z, v = blockIdx.z - p * Z;
  if (P.d[p] != D) return;
  const int MI = P.mul_in[p], MO = P.mul_out[p];
  constexpr int T = 32;  // output tile T x T per CTA, 2 x 2 entries per thread: every staged x / g value feeds two FMAs
  const int tiles_w = (MO + T - 1) / T;
  const int tu = blockIdx.x / tiles_w, tw = blockIdx.x - tu * tiles_w;
  if (tu * T >= MI) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_list[EL_CHUNK];
  __shared__ float s_coef[EL_CHUNK];
  __shared__ int s_wcnt[8];
  __shared__ float xs[EL_NB][T * D], gs[EL_NB][T * D];

  const int n = blockIdx.y * EL_CHUNK + tid;
  const float a = n < N ? __ldg(attrs + (size_t)n * Z + v) : 0.f;
  const unsigned bal = __ballot_sync(0xffffffffu, a != 0.f);
  if (lane == 0) s_wcnt[warp] = __popc(bal);
  __syncthreads();
  int off = 0, total = 0;
#pragma unroll
  for (int w8 = 0; w8 < 8; ++w8) {
    if (w8 < warp) off += s_wcnt[w8];
    total += s_wcnt[w8];
  }
  if (total == 0) return;  // uniform
  if (a != 0.f) {
    const int pos = off + __popc(bal & ((1u << lane) - 1u));
    s_list[pos] = n;
    s_coef[pos] = a;
  }
  __syncthreads();

  const int ul = tid >> 4, wl = tid & 15;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int j0 = 0; j0 < total; j0 += EL_NB) {
    const int nb = min(EL_NB, total - j0);
    for (int i = tid; i < nb * T * D; i += 256) {
      const int j = i / (T * D), rem = i - j * (T * D);
      const int uu = rem / D, m = rem - uu * D;
      const int node = s_list[j0 + j];
      const int gu = tu * T + uu, gw = tw * T + uu;
      xs[j][rem] = gu < MI ? __ldg(x + (size_t)node * P.x_rs + P.x_off[p] + (size_t)gu * D + m) : 0.f;
      gs[j][rem] = gw < MO ? __ldg(g + (size_t)node * P.o_rs + P.o_off[p] + (size_t)gw * P.o_ws[p] + m) : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
      float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int m = 0; m < D; ++m) {
        const float x0 = xs[j][ul * D + m], x1 = xs[j][(ul + 16) * D + m];
        const float g0 = gs[j][wl * D + m], g1 = gs[j][(wl + 16) * D + m];
        s[0][0] = fmaf(x0, g0, s[0][0]);
        s[0][1] = fmaf(x0, g1, s[0][1]);
        s[1][0] = fmaf(x1, g0, s[1][0]);
        s[1][1] = fmaf(x1, g1, s[1][1]);
      }
      const float c = s_coef[j0 + j];
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) acc[a2][b2] = fmaf(c, s[a2][b2], acc[a2][b2]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
    for (int b2 = 0; b2 < 2; ++b2) {
      const int u = tu * T + ul + 16 * a2, w = tw * T + wl + 16 * b2;
      if (u < MI && w < MO && acc[a2][b2] != 0.f)
        atomicAdd(dW + P.w_off[p] + ((size_t)u * Z + v) * MO + w, P.alpha[p] * acc[a2][b2]);
    }
}

}  // namespace pml

static inline bool pml_el_has_d(const pml_el_plan* plan, int d) {
  for (int p = 0; p < plan->npaths; ++p)
    if (plan->d[p] == d) return true;
  return false;
}
#define PML_EL_FOR_D(CALL) \
  if (pml_el_has_d(plan, 1)) { CALL(1) } if (pml_el_has_d(plan, 3)) { CALL(3) } \
  if (pml_el_has_d(plan, 5)) { CALL(5) } if (pml_el_has_d(plan, 7)) { CALL(7) }

extern "C" int pml_element_linear_fwd(const float* x, const float* attrs, const float* W, const pml_el_plan* plan,
                                      float* out, int N, int Z, int accumulate, cudaStream_t stream) {
  if (N <= 0 || plan->npaths <= 0) return PML_OK;
  dim3 grid(N, plan->npaths);
  int max_mi = 0;
  for (int p = 0; p < plan->npaths; ++p) max_mi = plan->mul_in[p] > max_mi ? plan->mul_in[p] : max_mi;
#define CALL(D) pml::launch_plain(pml::el_fwd_kernel<D>, dim3(grid), dim3(128), (size_t)max_mi * D * sizeof(float), stream, x, attrs, W, *plan, out, N, Z, accumulate);
  PML_EL_FOR_D(CALL)
#undef CALL
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// dx must be zero-initialised by the caller (contributions are accumulated atomically; one addend per entry when the
// attrs rows are one-hot).
extern "C" int pml_element_linear_bwd_x(const float* g, const float* attrs, const float* W, const pml_el_plan* plan,
                                        float* dx, int N, int Z, int accumulate, cudaStream_t stream) {
  (void)accumulate;
  if (N <= 0 || plan->npaths <= 0) return PML_OK;
  const int chunk = N <= 8192 ? 64 : pml::EL_XCHUNK;
  dim3 grid(1, (N + chunk - 1) / chunk, plan->npaths * Z);
  // per D: the largest slab piece (multiple of 32 columns, or all columns) whose staging fits 160 KB of shared memory
  size_t smem[8] = {0};
  int wcs[8] = {0};
  const size_t budget = 160 * 1024;
  for (int d = 1; d <= 7; d += 2) {
    int wc = 1 << 30;
    bool any = false;
    for (int p = 0; p < plan->npaths; ++p) {
      if (plan->d[p] != d) continue;
      any = true;
      const int mi = plan->mul_in[p], mo = plan->mul_out[p];
      const int nb = d >= 5 ? 4 : 8, lanes_u = mi >= 256 ? 256 : mi;
      const int dp = d == 1 ? 1 : (d == 3 ? 4 : 8);
      const size_t per_col = ((size_t)mi + (size_t)nb * (256 / lanes_u) * dp) * sizeof(float);
      int fit = (int)((budget - (size_t)(mi + 4) * sizeof(float)) / per_col);
      if (fit >= mo) fit = mo;
      else fit = fit / 32 * 32;
      if (fit < 1) return PML_ERR_UNSUPPORTED;
      if (fit < wc) wc = fit;
    }
    if (!any) continue;
    wcs[d] = wc;
    for (int p = 0; p < plan->npaths; ++p) {
      if (plan->d[p] != d) continue;
      const int mi = plan->mul_in[p], mo = plan->mul_out[p];
      const int nb = d >= 5 ? 4 : 8, lanes_u = mi >= 256 ? 256 : mi;
      const int w = wc < mo ? wc : mo;
      const int dp = d == 1 ? 1 : (d == 3 ? 4 : 8);
      const size_t need = ((((size_t)mi * (w | 1) + 3) & ~(size_t)3) + (size_t)nb * (256 / lanes_u) * w * dp) * sizeof(float);
      if (need > smem[d]) smem[d] = need;
    }
  }
#define CALL(D)                                                                                                          \
  {                                                                                                                      \
    static size_t configured = 0;                                                                                        \
    if (smem[D] > configured) {                                                                                          \
      if (cudaFuncSetAttribute(pml::el_bwd_x_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem[D]) !=    \
          cudaSuccess)                                                                                                   \
        return PML_ERR_LAUNCH;                                                                                           \
      configured = smem[D];                                                                                              \
    }                                                                                                                    \
    pml::launch_plain(pml::el_bwd_x_kernel<D>, dim3(grid), dim3(256), smem[D], stream, g, attrs, W, *plan, dx, N, Z, wcs[D], chunk);                        \
  }
  PML_EL_FOR_D(CALL)
#undef CALL
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// dW must be zero-initialised by the caller (node-chunk partials are accumulated atomically).
extern "C" int pml_element_linear_bwd_w(const float* x, const float* g, const float* attrs, const pml_el_plan* plan,
                                        float* dW, int N, int Z, int accumulate, cudaStream_t stream) {
  (void)accumulate;
  if (plan->npaths <= 0 || N <= 0) return PML_OK;
  int max_t = 0;
  for (int p = 0; p < plan->npaths; ++p) {
    const int t = ((plan->mul_in[p] + 31) / 32) * ((plan->mul_out[p] + 31) / 32);
    max_t = t > max_t ? t : max_t;
  }
  dim3 grid(max_t, (N + pml::EL_CHUNK - 1) / pml::EL_CHUNK, plan->npaths * Z);
#define CALL(D) pml::launch_plain(pml::el_bwd_w_kernel<D>, dim3(grid), dim3(256), 0, stream, x, g, attrs, *plan, dW, N, Z);
  PML_EL_FOR_D(CALL)
#undef CALL
  PML_CHECK_LAUNCH();
  return PML_OK;
}
