// Fused channelwise ('uvu') edge tensor product + gather + segmented scatter-sum, and its cotangent kernel.
//
// Replaces the reference op chain   w_h[sender] -> o3.TensorProduct(uvu, external weights) -> scatter(sum)
//   MACE   : src/physicsml/models/mace/modules/blocks.py:214-226   (gather edge_index[0], scatter edge_index[1])
//   NequIP : src/physicsml/models/nequip/modules/interaction_block.py:95-96 (gather [1], scatter [0])
// without ever materialising the [E, d_tp] message tensor.
//
//   out[n, blk(p), u, k] = sum_{e : scatter(e) = n}  R[e, p*C+u] * sum_{ij} c_p[i,j,k] * h[gather(e), blk1(p), u, i] * Y[e, l2(p), j]
//
// Work decomposition: one warp owns (scatter node n, slab of 32 channels) and walks that node's edge segment
// (CSR over the scatter index), so the reduction over edges is a register accumulation — deterministic, no atomics.
// R (the dominant HBM stream, 4*NP*C bytes / edge) is read exactly once, coalesced along channels, with streaming
// loads that do not allocate in L1; gathered node rows and Y stay cacheable.
//
// Compiled once per tensor-product signature: -DPML_TP_ID=<id> selects the generated TPCode<id> (gen/tp_gen.cuh).
#include "pml_common.cuh"
#include "gen/tp_gen.cuh"

#ifndef PML_TP_ID
#error "compile with -DPML_TP_ID=<spec id>"
#endif

namespace pml {

template <int NV>
struct Log2 {
  static constexpr int value = 1 + Log2<NV / 2>::value;
};
template <>
struct Log2<1> {
  static constexpr int value = 0;
};

// Sum a per-lane vector v[NV] (NV a power of two <= 32) over the 32 lanes with NV-1+log-steps shuffles instead of
// 5*NV: recursive halving — after the call v[0] of lane l holds the total of component (l >> (5 - log2 NV)).
template <int NV>
__device__ __forceinline__ void warp_reduce_vec(float (&v)[NV], int lane) {
  int cnt = NV;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (cnt > 1) {
      const int half = cnt / 2;
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) {
        if (i < half) {
          const float mine = up ? v[half + i] : v[i];
          const float other = up ? v[i] : v[half + i];
          v[i] = mine + __shfl_xor_sync(0xffffffffu, other, off);
        }
      }
      cnt = half;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}

__host__ __device__ constexpr int next_pow2(int x) { return x <= 1 ? 1 : 2 * next_pow2((x + 1) / 2); }

template <int ID>
__global__ void __launch_bounds__(128) tp_fwd_kernel(const float* __restrict__ h, const float* __restrict__ Y,
                                                     const float* __restrict__ R, const int* __restrict__ gather,
                                                     const int* __restrict__ rowptr, const int* __restrict__ perm,
                                                     float* __restrict__ out, int N, int C, int nslab) {
  using T = TPCode<ID>;
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (long long)N * nslab) return;
  const int n = (int)(item / nslab);
  const int u = (int)(item - (long long)n * nslab) * 32 + lane;
  const bool active = u < C;
  const int uc = active ? u : C - 1;  // clamp: inactive lanes compute on valid memory and discard

  float acc[T::DOUT];
#pragma unroll
  for (int k = 0; k < T::DOUT; ++k) acc[k] = 0.f;

  const int beg = __ldg(rowptr + n), end = __ldg(rowptr + n + 1);
  const size_t rstride = (size_t)T::NP * C;
  const size_t hstride = (size_t)T::DIN * C;

  int q = beg;
  // two edges per iteration: all loads of both edges are issued before the first FMA
  for (; q + 1 < end; q += 2) {
    const int e0 = perm ? __ldg(perm + q) : q;
    const int e1 = perm ? __ldg(perm + q + 1) : q + 1;
    const int s0 = __ldg(gather + e0), s1 = __ldg(gather + e1);
    float R0[T::NP], R1[T::NP], Y0[T::NSH], Y1[T::NSH], h0[T::DIN], h1[T::DIN];
#pragma unroll
    for (int p = 0; p < T::NP; ++p) {
      R0[p] = ld_stream(R + (size_t)e0 * rstride + (size_t)p * C + uc);
      R1[p] = ld_stream(R + (size_t)e1 * rstride + (size_t)p * C + uc);
    }
#pragma unroll
    for (int j = 0; j < T::NSH; ++j) {
      Y0[j] = __ldg(Y + (size_t)e0 * T::NSH + j);
      Y1[j] = __ldg(Y + (size_t)e1 * T::NSH + j);
    }
    T::load_h(h + (size_t)s0 * hstride, C, uc, h0);
    T::load_h(h + (size_t)s1 * hstride, C, uc, h1);
    T::fwd(h0, Y0, R0, acc);
    T::fwd(h1, Y1, R1, acc);
  }
  if (q < end) {
    const int e0 = perm ? __ldg(perm + q) : q;
    const int s0 = __ldg(gather + e0);
    float R0[T::NP], Y0[T::NSH], h0[T::DIN];
#pragma unroll
    for (int p = 0; p < T::NP; ++p) R0[p] = ld_stream(R + (size_t)e0 * rstride + (size_t)p * C + uc);
#pragma unroll
    for (int j = 0; j < T::NSH; ++j) Y0[j] = __ldg(Y + (size_t)e0 * T::NSH + j);
    T::load_h(h + (size_t)s0 * hstride, C, uc, h0);
    T::fwd(h0, Y0, R0, acc);
  }
  if (active) T::store_out(out + (size_t)n * T::DOUT * C, C, u, acc);
}

// Cotangents of the trilinear form, grouped by the SAME scatter segments as the forward so that the large row
// g[n] (4*DOUT*C bytes) is loaded once per node and kept in registers:
//   dR[e,p,u]  (no reduction)                       -> streamed out, coalesced
//   dY[e,j]    = sum over channels                  -> vector warp reduction (+ atomics across channel slabs)
//   dh[gather(e)] += ...                            -> fp32 red.global.add (gather side is the small side: DIN << DOUT)
template <int ID, bool WANT_H, bool WANT_Y, bool WANT_R>
__global__ void __launch_bounds__(128) tp_bwd_kernel(const float* __restrict__ h, const float* __restrict__ Y,
                                                     const float* __restrict__ R, const float* __restrict__ g,
                                                     const int* __restrict__ gather, const int* __restrict__ rowptr,
                                                     const int* __restrict__ perm, float* __restrict__ dh,
                                                     float* __restrict__ dY, float* __restrict__ dR, int N, int C,
                                                     int nslab) {
  using T = TPCode<ID>;
  constexpr int NV = next_pow2(T::NSH);
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (long long)N * nslab) return;
  const int n = (int)(item / nslab);
  const int u = (int)(item - (long long)n * nslab) * 32 + lane;
  const bool active = u < C;
  const int uc = active ? u : C - 1;

  float gv[T::DOUT];
  T::load_out(g + (size_t)n * T::DOUT * C, C, uc, gv);
  if (!active) {
#pragma unroll
    for (int k = 0; k < T::DOUT; ++k) gv[k] = 0.f;
  }
  const int beg = __ldg(rowptr + n), end = __ldg(rowptr + n + 1);
  const size_t rstride = (size_t)T::NP * C;
  const size_t hstride = (size_t)T::DIN * C;

  for (int q = beg; q < end; ++q) {
    const int e = perm ? __ldg(perm + q) : q;
    const int s = __ldg(gather + e);
    float Rv[T::NP], Yv[T::NSH], hv[T::DIN];
#pragma unroll
    for (int p = 0; p < T::NP; ++p) Rv[p] = (WANT_H || WANT_Y) ? ld_stream(R + (size_t)e * rstride + (size_t)p * C + uc) : 0.f;
#pragma unroll
    for (int j = 0; j < T::NSH; ++j) Yv[j] = __ldg(Y + (size_t)e * T::NSH + j);
    T::load_h(h + (size_t)s * hstride, C, uc, hv);

    float dhv[T::DIN], dYv[NV], dRv[T::NP];
#pragma unroll
    for (int i = 0; i < T::DIN; ++i) dhv[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) dYv[j] = 0.f;
    T::template bwd<WANT_H, WANT_Y, WANT_R, NV>(hv, Yv, Rv, gv, dhv, dYv, dRv);

    if (WANT_R && active) {
#pragma unroll
      for (int p = 0; p < T::NP; ++p) st_stream(dR + (size_t)e * rstride + (size_t)p * C + u, dRv[p]);
    }
    if (WANT_H && active) T::atomic_add_h(dh + (size_t)s * hstride, C, u, dhv);
    if (WANT_Y) {
      warp_reduce_vec<NV>(dYv, lane);
      constexpr int sh = 5 - Log2<NV>::value;
      const int j = lane >> sh;
      if ((lane & ((1 << sh) - 1)) == 0 && j < T::NSH) {
        if (nslab == 1)
          dY[(size_t)e * T::NSH + j] = dYv[0];
        else
          atomicAdd(dY + (size_t)e * T::NSH + j, dYv[0]);
      }
    }
  }
}

}  // namespace pml

#define PML_CAT_(a, b) a##b
#define PML_CAT(a, b) PML_CAT_(a, b)

// Launchers (one pair per compiled signature; dispatched by id in capi.cu).  Streams are the caller's.
extern "C" int PML_CAT(pml_tp_fwd_launch_, PML_TP_ID)(const float* h, const float* Y, const float* R, const int* gather,
                                                       const int* rowptr, const int* perm, float* out, int N, int C,
                                                       cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  const int nslab = (C + 31) / 32;
  const int wpb = 4;
  const long long items = (long long)N * nslab;
  pml::tp_fwd_kernel<PML_TP_ID><<<pml::ceil_div(items, wpb), wpb * 32, 0, stream>>>(h, Y, R, gather, rowptr, perm, out, N, C, nslab);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// dh and (when C > 32) dY must be zero-initialised by the caller; dR is fully overwritten.
extern "C" int PML_CAT(pml_tp_bwd_launch_, PML_TP_ID)(const float* h, const float* Y, const float* R, const float* g,
                                                       const int* gather, const int* rowptr, const int* perm, float* dh,
                                                       float* dY, float* dR, int N, int C, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  const int nslab = (C + 31) / 32;
  const int wpb = 4;
  const long long items = (long long)N * nslab;
  const int grid = pml::ceil_div(items, wpb);
  const int mask = (dh ? 1 : 0) | (dY ? 2 : 0) | (dR ? 4 : 0);
#define PML_TP_BWD_CASE(M, H_, Y_, R_)                                                                              \
  case M:                                                                                                           \
    pml::tp_bwd_kernel<PML_TP_ID, H_, Y_, R_><<<grid, wpb * 32, 0, stream>>>(h, Y, R, g, gather, rowptr, perm, dh, dY, \
                                                                            dR, N, C, nslab);                       \
    break;
  switch (mask) {
    PML_TP_BWD_CASE(1, true, false, false)
    PML_TP_BWD_CASE(2, false, true, false)
    PML_TP_BWD_CASE(3, true, true, false)
    PML_TP_BWD_CASE(4, false, false, true)
    PML_TP_BWD_CASE(5, true, false, true)
    PML_TP_BWD_CASE(6, false, true, true)
    PML_TP_BWD_CASE(7, true, true, true)
    default:
      return PML_OK;
  }
#undef PML_TP_BWD_CASE
  PML_CHECK_LAUNCH();
  return PML_OK;
}
