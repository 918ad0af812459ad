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
#ifndef PML_GEN_TP  // a build variant (physicsml_b200.build --variant) points this at its own generated header
#define PML_GEN_TP "gen/tp_gen.cuh"
#endif
#include PML_GEN_TP

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
                                                     float* __restrict__ out, int N, int C, int nslab, int mm) {
  pdl_enter();
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
  if (active) {  // mm: component-major blocks (m*C + u) instead of the e3nn order (u*(2l+1) + m)
    if (mm & 1) T::template store_out<true>(out + (size_t)n * T::DOUT * C, C, u, acc);
    else T::template store_out<false>(out + (size_t)n * T::DOUT * C, C, u, acc);
  }
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
                                                     int nslab, int mm) {
  pdl_enter();
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
  if (mm & 1) T::template load_out<true>(g + (size_t)n * T::DOUT * C, C, uc, gv);
  else T::template load_out<false>(g + (size_t)n * T::DOUT * C, C, uc, gv);
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
      for (int p = 0; p < T::NP; ++p) {
        float* q_ = dR + (size_t)e * rstride + (size_t)p * C + u;
        if (mm & 2) atomicAdd(q_, dRv[p]);  // accumulate onto an existing dR (second-order pass)
        else st_stream(q_, dRv[p]);
      }
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


// =====================================================================================================================
// Pipelined variants (the ones that run whenever the row sizes are 16-byte multiples): persistent CTAs, each owning a
// contiguous range of scatter nodes holding ~E/grid edges.  Every edge's operands — the R row (NP*C floats) and the
// gathered h row (DIN*C floats) by `cp.async.bulk` (TMA 1-D bulk copies), the Y row by 4-byte cp.async — are streamed
// into an NSTAGE-deep shared-memory ring guarded by full/empty mbarriers, several edges ahead of the arithmetic, so no
// thread ever waits on a global load.  The ring runs across node boundaries (no per-node ramp-up): bytes in flight
// per SM = resident CTAs x NSTAGE x row bytes (~200 KB at C=128, lmax 3) instead of what the register file of the
// load-compute-load loop above can hold (~20 KB) — the bound of a latency-limited stream.  Each thread owns two
// channels and computes on the packed fp32x2 pipe (FFMA2), with the 3j contraction factored to about one packed
// instruction per non-zero coefficient; without that the kernels are issue-bound, not HBM-bound.
// =====================================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 eviction policies: R / dR are touched exactly once (evict-first) so that the gathered node rows, the dh rows
// under reduction and Y stay resident; without them ncu shows 13.8 GB (fwd) / 30.8 GB (cotangent) of DRAM traffic for
// 11.1 / 21.7 GB of algorithmic bytes at config 4 — the stream keeps flushing the 49 MB of node rows out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void st_stream2_hint(float* p, float2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int kPipeMaxStages = 8;

// T is a path group (TPGroup<ID, g>; the whole signature when it has a single group): it sees its own contiguous slice
// of the R row (paths P0 .. P0+NP) and of the gathered h row (slots H_OFF .. H_OFF+DIN) and owns its output blocks.
template <class T>
struct PipeLayout {
  static constexpr int YPAD = (2 * T::NSH + 3) & ~3;  // Y is staged as duplicated pairs (Y_j, Y_j): packed operands
  __host__ __device__ static int r_floats(int C) { return T::NP * C; }
  __host__ __device__ static int h_floats(int C) { return T::DIN * C; }
  __host__ __device__ static int stage_floats(int C) { return r_floats(C) + h_floats(C) + YPAD + 4; }  // + {edge id, gather id}
  __host__ __device__ static bool rows_ok(int C) {
    return (r_floats(C) % 4 == 0) && (h_floats(C) % 4 == 0) && (C % 2 == 0) && C <= 1984 && 2 * T::NSH <= 32;
  }
  __host__ __device__ static bool supported(int C) {
    // DOUT <= 48: two channels of accumulators (or of the cotangent row) must fit the register file without spilling
    // (wider signatures are cut into path groups by the code generator); slices of a row must start 16-byte aligned
    return rows_ok(C) && T::DOUT <= 48 && (T::NGROUPS == 1 || C % 4 == 0);
  }
  __host__ __device__ static int consumer_threads(int C) { return ((C / 2 + 31) / 32) * 32; }
};

// Work distribution: the node range is cut into tiles of TN consecutive scatter nodes (~kTileEdges edges each) and
// CTA b takes tiles b, b+G, b+2G, ...  All CTAs advance at about the same pace, so at any moment the grid works on one
// window of ~G consecutive tiles: the node rows it gathers (h) and reduces into (dh) are shared through L2 while they
// are hot instead of being spread over the whole kernel duration (with one contiguous block of nodes per CTA the
// cotangent kernel moved 29 GB through DRAM for 21.7 GB of algorithmic traffic: dh rows were evicted between reds).
constexpr int kTileEdges = 512;

struct TileSched {
  int TN, ntiles, G;  // nodes per tile, number of tiles, CTAs that take part (the others exit)
  // Tiles are walked from the END of the node range: the R rows (and, in the cotangent pass, the g rows) of the last
  // nodes are the ones their producer kernel wrote last, i.e. the part of a tensor larger than L2 that is still
  // resident when this kernel starts.
  __device__ __forceinline__ int first_node(int tile) const { return (ntiles - 1 - tile) * TN; }
};

__device__ __forceinline__ TileSched tile_schedule(const int* __restrict__ rowptr, int N) {
  const int E = __ldg(rowptr + N);
  long long want = (long long)E / kTileEdges;  // tiles of ~kTileEdges edges ...
  if (want < 3ll * gridDim.x) want = 3ll * gridDim.x;  // ... but at least three per CTA (evens out the tile sizes)
  TileSched t;
  t.TN = (int)((long long)N / want);
  if (t.TN < 1) t.TN = 1;
  t.ntiles = (N + t.TN - 1) / t.TN;
  const int waves = (t.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
  t.G = (t.ntiles + waves - 1) / waves;  // same number of waves with an even share per CTA
  return t;
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Producer: walks the CTA's tiles edge by edge.  The (edge id, gather id) pairs are fetched 32 edges at a time, one
// chunk ahead (the chunk after the current one may already belong to the next tile), the bounds of the next tile one
// tile ahead, so no load is ever waited for right after it was issued; edges are handed out with shuffles.
constexpr int kPipeHeaderBytes = 256;  // 16 mbarriers (128 B) + spare, then the stage ring

template <class T, bool LOAD_R>
struct PipeProducer {
  using L = PipeLayout<T>;
  const float *h, *Y, *R;
  const int *gather, *perm, *rowptr;
  float* stages;
  uint64_t *full, *empty;
  int N, C, nstage;
  TileSched ts;
  int ftile, fq, fq_hi;  // fetch cursor: tile being fetched and its remaining edge range
  int nq, nq_hi;         // edge range of the CTA's next tile (prefetched)
  int issued;            // edges issued so far
  int pos;               // position inside the current chunk
  int stage;             // ring slot the next edge goes to
  uint32_t phase;
  int e_cur, s_cur, cnt_cur, e_nxt, s_nxt, cnt_nxt;

  __device__ __forceinline__ void tile_bounds(int tile, int& q, int& q_hi) const {
    q = 0, q_hi = 0;
    if (tile < ts.ntiles) {
      const int n0 = ts.first_node(tile), n1 = min(N, n0 + ts.TN);
      q = __ldg(rowptr + n0), q_hi = __ldg(rowptr + n1);
    }
  }
  __device__ __forceinline__ void fetch(int lane, int& e, int& s, int& cnt) {
    while (fq >= fq_hi && ftile < ts.ntiles) {  // next non-empty tile of this CTA
      ftile += ts.G;
      fq = nq, fq_hi = nq_hi;
      tile_bounds(ftile + ts.G, nq, nq_hi);
    }
    e = 0, s = 0, cnt = 0;
    if (fq >= fq_hi) return;
    cnt = min(32, fq_hi - fq);
    if (lane < cnt) {
      const int q = fq + lane;
      e = perm ? __ldg(perm + q) : q;
      s = __ldg(gather + e);
    }
    fq += cnt;
  }
  __device__ __forceinline__ void start(int lane) {
    issued = 0, pos = 0, stage = 0, phase = 0;
    ftile = blockIdx.x;
    tile_bounds(ftile, fq, fq_hi);
    tile_bounds(ftile + ts.G, nq, nq_hi);
    fetch(lane, e_cur, s_cur, cnt_cur);
    fetch(lane, e_nxt, s_nxt, cnt_nxt);
  }
  // one contiguous edge range instead of tiles (cotangent pass)
  __device__ __forceinline__ void start_range(int lane, int q_lo, int q_hi) {
    issued = 0, pos = 0, stage = 0, phase = 0;
    ftile = ts.ntiles;
    fq = q_lo, fq_hi = q_hi, nq = 0, nq_hi = 0;
    fetch(lane, e_cur, s_cur, cnt_cur);
    fetch(lane, e_nxt, s_nxt, cnt_nxt);
  }
  __device__ __forceinline__ bool done() const { return cnt_cur == 0; }
  // issue up to `limit` edges while fewer than `max_ahead` edges are in flight beyond `consumed`; a slot still being
  // read by another consumer warp ends the attempt unless `blocking`
  __device__ __forceinline__ void issue(int limit, int consumed, int max_ahead, bool blocking, int lane) {
    const int rfl = L::r_floats(C), hfl = L::h_floats(C), sfl = L::stage_floats(C);
    for (int n = 0; n < limit && cnt_cur > 0 && issued - consumed < max_ahead; ++n) {
      const uint32_t eb = smem_u32(empty + stage);
      if (blocking)
        mbar_wait(eb, phase ^ 1u);
      else if (!mbar_test(eb, phase ^ 1u))
        break;
      const int e = __shfl_sync(0xffffffffu, e_cur, pos);
      const int s = __shfl_sync(0xffffffffu, s_cur, pos);
      float* st = stages + (size_t)stage * sfl;
      const uint32_t fb = smem_u32(full + stage);
      if (lane == 0) {
        reinterpret_cast<int2*>(st + rfl + hfl + L::YPAD)[0] = make_int2(e, s);  // released by the arrive below
        mbar_expect_tx(fb, (LOAD_R ? (uint32_t)rfl * 4u : 0u) + (uint32_t)hfl * 4u);
        if (LOAD_R)
          bulk_g2s_hint(smem_u32(st), R + ((size_t)e * T::NP_FULL + T::P0) * C, (uint32_t)rfl * 4u, fb, l2_policy_evict_first());
        bulk_g2s_hint(smem_u32(st + rfl), h + ((size_t)s * T::DIN_FULL + T::H_OFF) * C, (uint32_t)hfl * 4u, fb, l2_policy_evict_last());
      }
      if (lane < 2 * T::NSH) cp_async4(smem_u32(st + rfl + hfl + lane), Y + (size_t)e * T::NSH + (lane >> 1));
      cp_async_mbar_arrive_noinc(fb);
      ++issued;
      if (++stage == nstage) {
        stage = 0;
        phase ^= 1u;
      }
      if (++pos == cnt_cur) {
        pos = 0;
        e_cur = e_nxt, s_cur = s_nxt, cnt_cur = cnt_nxt;
        if (cnt_cur > 0) fetch(lane, e_nxt, s_nxt, cnt_nxt);
      }
    }
  }
};

__device__ __forceinline__ void pipe_init(uint64_t* full, uint64_t* empty, int nstage, int consumer_warps) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(smem_u32(full + s), 33);               // 32 cp.async arrivals + the expect_tx arrival
      mbar_init(smem_u32(empty + s), consumer_warps);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
}

// Y pairs from shared memory: (Y_j, Y_j) sits at float offset 2j (16-byte vector loads)
template <int NSH, int YPAD>
__device__ __forceinline__ void load_y2_smem(const float* ys, float2 (&Yv)[NSH]) {
  float tmp[YPAD];
#pragma unroll
  for (int j = 0; j < YPAD / 4; ++j) {
    const float4 v = reinterpret_cast<const float4*>(ys)[j];
    tmp[4 * j] = v.x, tmp[4 * j + 1] = v.y, tmp[4 * j + 2] = v.z, tmp[4 * j + 3] = v.w;
  }
#pragma unroll
  for (int j = 0; j < NSH; ++j) Yv[j] = make_float2(tmp[2 * j], tmp[2 * j + 1]);
}

// blockDim.x = 32*ceil(C/64) consumer threads: two channels each, packed fp32x2 arithmetic.  The copies are issued by
// a dedicated extra warp (DEDICATED: the ring is refilled the moment a slot is released, independent of anybody's
// arithmetic) or by warp 0 between its own edges (no extra warp: a third more CTAs fit when registers bound occupancy).
#define PML_PIPE_PROLOGUE(LOAD_R_, RANGED_)                                                                                     \
  extern __shared__ __align__(128) unsigned char smem_raw[];                                                           \
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                                                              \
  uint64_t* empty = full + kPipeMaxStages;                                                                             \
  float* stages = reinterpret_cast<float*>(smem_raw + kPipeHeaderBytes);                                               \
  const int nthr_c = blockDim.x - (DEDICATED ? 32 : 0);                                                                \
  const int consumer_warps = nthr_c >> 5;                                                                              \
  TileSched ts;                                                                                                        \
  int q_lo = 0, q_hi = 0;                                                                                              \
  if (RANGED_) { /* even split of the edge list, walked from the end like the tiles */                                 \
    const long long E_ = __ldg(rowptr + N), b_ = (long long)gridDim.x - 1 - blockIdx.x;                                \
    q_lo = (int)(E_ * b_ / gridDim.x), q_hi = (int)(E_ * (b_ + 1) / gridDim.x);                                        \
    ts.TN = 1, ts.ntiles = 0, ts.G = (int)gridDim.x;                                                                   \
    if (q_lo >= q_hi) return;                                                                                          \
  } else {                                                                                                             \
    ts = tile_schedule(rowptr, N);                                                                                     \
    if ((int)blockIdx.x >= ts.G) return;                                                                               \
  }                                                                                                                    \
  pipe_init(full, empty, nstage, consumer_warps);                                                                      \
  __syncthreads();                                                                                                     \
  const int lane = threadIdx.x & 31;                                                                                   \
  const bool is_producer = !DEDICATED && threadIdx.x < 32;                                                             \
  PipeProducer<T, LOAD_R_> prod;                                                                                      \
  prod.h = h, prod.Y = Y, prod.R = R, prod.gather = gather, prod.perm = perm, prod.rowptr = rowptr;                    \
  prod.stages = stages, prod.full = full, prod.empty = empty, prod.N = N, prod.C = C, prod.nstage = nstage, prod.ts = ts; \
  if (DEDICATED && (int)threadIdx.x >= nthr_c) {                                                                       \
    if (RANGED_) prod.start_range(lane, q_lo, q_hi); else prod.start(lane);                                                                                               \
    while (!prod.done()) prod.issue(1 << 30, 0, 1 << 30, true, lane);                                                  \
    return;                                                                                                            \
  }                                                                                                                    \
  if (is_producer) {                                                                                                   \
    if (RANGED_) prod.start_range(lane, q_lo, q_hi); else prod.start(lane);                                                                                               \
    prod.issue(nstage, 0, nstage, true, lane);                                                                         \
  }                                                                                                                    \
  const int u = 2 * threadIdx.x; /* channels (u, u+1) */                                                               \
  const bool active = u < C;                                                                                           \
  const int uc = active ? u : C - 2;                                                                                   \
  const int rfl = L::r_floats(C), hfl = L::h_floats(C), sfl = L::stage_floats(C);                                      \
  int stage = 0, consumed = 0;                                                                                         \
  uint32_t phase = 0;

template <class T, bool DEDICATED>
__global__ void tp_fwd_pipe_kernel(const float* __restrict__ h, const float* __restrict__ Y, const float* __restrict__ R,
                                   const int* __restrict__ gather, const int* __restrict__ rowptr,
                                   const int* __restrict__ perm, float* __restrict__ out, int N, int C, int nstage, int mm) {
  pdl_enter();
  using L = PipeLayout<T>;
  PML_PIPE_PROLOGUE(true, false)
  for (int tile = blockIdx.x; tile < ts.ntiles; tile += ts.G) {
    const int n0 = ts.first_node(tile), n1 = min(N, n0 + ts.TN);
    int q = __ldg(rowptr + n0);
    int end = __ldg(rowptr + n0 + 1);
    for (int n = n0; n < n1; ++n) {
      const int end_next = (n + 2 <= N) ? __ldg(rowptr + n + 2) : end;  // prefetched one node ahead
      float2 acc[T::DOUT];
#pragma unroll
      for (int k = 0; k < T::DOUT; ++k) acc[k] = make_float2(0.f, 0.f);
      for (; q < end; ++q) {
        const float* st = stages + (size_t)stage * sfl;
        if (is_producer && prod.issued == consumed) prod.issue(1, consumed, 1, true, lane);  // never wait for an edge nobody requested
        mbar_wait(smem_u32(full + stage), phase);
        float2 Rv[T::NP], Yv[T::NSH], hv[T::DIN];
#pragma unroll
        for (int p = 0; p < T::NP; ++p) Rv[p] = *reinterpret_cast<const float2*>(st + p * C + uc);
        T::load_h2_smem(st + rfl, C, uc, hv);
        load_y2_smem<T::NSH, L::YPAD>(st + rfl + hfl, Yv);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(empty + stage));
        if (++stage == nstage) {
          stage = 0;
          phase ^= 1u;
        }
        ++consumed;
        if (is_producer) prod.issue(2, consumed, nstage, false, lane);
        T::fwd2(hv, Yv, Rv, acc);
      }
      if (active) {
        if (mm & 1) T::template store_out2<true>(out + (size_t)n * T::DOUT_FULL * C, C, u, acc);
        else T::template store_out2<false>(out + (size_t)n * T::DOUT_FULL * C, C, u, acc);
      }
      end = end_next;
    }
  }
}

#ifdef PML_TP_BWD_MAXREG  // tuning builds: cap the cotangent kernel's registers (one more CTA per SM, at the price of spills)
#define PML_BWD_REGS __maxnreg__(PML_TP_BWD_MAXREG)
#else
#define PML_BWD_REGS
#endif
template <class T, bool DEDICATED, bool WANT_H, bool WANT_Y, bool WANT_R>
__global__ void PML_BWD_REGS tp_bwd_pipe_kernel(const float* __restrict__ h, const float* __restrict__ Y, const float* __restrict__ R,
                                   const float* __restrict__ g, const int* __restrict__ gather,
                                   const int* __restrict__ rowptr, const int* __restrict__ perm, float* __restrict__ dh,
                                   float* __restrict__ dY, float* __restrict__ dR, int N, int C, int nstage, int mm) {
  pdl_enter();
  using L = PipeLayout<T>;
  constexpr int NV = next_pow2(T::NSH);
  PML_PIPE_PROLOGUE((WANT_H || WANT_Y), true)
  const size_t rstride = (size_t)T::NP_FULL * C, hstride = (size_t)T::DIN_FULL * C;  // rows of the FULL tensors
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  // The cotangent pass has no per-node output (dR, dY are per edge, dh is accumulated atomically), so a node's edges
  // may be split between CTAs: every CTA gets the same number of edges whatever the node degrees are.
  int n;  // node holding edge q_lo: rowptr[n] <= q_lo < rowptr[n + 1]
  {
    int lo = 0, hi = N;  // smallest m in (lo, hi] with rowptr[m] > q_lo
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(rowptr + mid) > q_lo) hi = mid; else lo = mid;
    }
    n = hi - 1;
  }
  {
    int q = q_lo;
    int end = min(__ldg(rowptr + n + 1), q_hi);
    for (;; ++n) {
      const int end_next = (n + 2 <= N) ? min(__ldg(rowptr + n + 2), q_hi) : q_hi;
      float2 gv[T::DOUT];
      if (q < end) {
        if (mm & 1) T::template load_out2<true>(g + (size_t)n * T::DOUT_FULL * C, C, uc, gv);
        else T::template load_out2<false>(g + (size_t)n * T::DOUT_FULL * C, C, uc, gv);
        if (!active) {
#pragma unroll
          for (int k = 0; k < T::DOUT; ++k) gv[k] = make_float2(0.f, 0.f);
        }
      }
      for (; q < end; ++q) {
        const float* st = stages + (size_t)stage * sfl;
        if (is_producer && prod.issued == consumed) prod.issue(1, consumed, 1, true, lane);
        mbar_wait(smem_u32(full + stage), phase);
        const int2 es = reinterpret_cast<const int2*>(st + rfl + hfl + L::YPAD)[0];  // {edge id, gather id}
        const int e = es.x;
        float2 Rv[T::NP], Yv[T::NSH], hv[T::DIN];
#pragma unroll
        for (int p = 0; p < T::NP; ++p)
          Rv[p] = (WANT_H || WANT_Y) ? *reinterpret_cast<const float2*>(st + p * C + uc) : make_float2(0.f, 0.f);
        T::load_h2_smem(st + rfl, C, uc, hv);
        load_y2_smem<T::NSH, L::YPAD>(st + rfl + hfl, Yv);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(empty + stage));
        if (++stage == nstage) {
          stage = 0;
          phase ^= 1u;
        }
        ++consumed;
        if (is_producer) prod.issue(2, consumed, nstage, false, lane);

        float2 dhv[T::DIN], dRv[T::NP];
        float dYv[NV];
#pragma unroll
        for (int i = 0; i < T::DIN; ++i) dhv[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NV; ++j) dYv[j] = 0.f;
        T::template bwd2<WANT_H, WANT_Y, WANT_R, NV>(hv, Yv, Rv, gv, dhv, dYv, dRv);

        if (WANT_R && active) {
          float* drow = dR + (size_t)e * rstride + (size_t)T::P0 * C + u;
#pragma unroll
          for (int p = 0; p < T::NP; ++p) {
            // mm bit 1: ADD into an existing dR row (the two dR contributions of the second-order pass are accumulated
            // here by L2 reductions — no load in the SM — instead of a separate R-sized add kernel)
            if (mm & 2) red_add_f32x2(drow + p * C, dRv[p].x, dRv[p].y, pol_stream);
            else st_stream2_hint(drow + p * C, dRv[p], pol_stream);
          }
        }
        if (WANT_H) {
          if (active) T::atomic_add_h2(dh + (size_t)es.y * hstride, C, u, dhv, pol_keep);
        }
        if (WANT_Y) {
          warp_reduce_vec<NV>(dYv, lane);
          constexpr int sh = 5 - Log2<NV>::value;
          const int j = lane >> sh;
          if ((lane & ((1 << sh) - 1)) == 0 && j < T::NSH) {
            if (consumer_warps == 1 && T::NGROUPS == 1)
              dY[(size_t)e * T::NSH + j] = dYv[0];
            else  // several consumer warps, or several path groups (one launch each), add into the same row
              atomicAdd(dY + (size_t)e * T::NSH + j, dYv[0]);
          }
        }
      }
      if (q >= q_hi) break;
      end = end_next;
    }
  }
}
#undef PML_PIPE_PROLOGUE

// =====================================================================================================================
// Intra-CTA path groups: ONE ring of full rows per CTA, consumed by TPGroups<ID>::N warp sets, each running the code of
// one path group (TPGroup<ID, g>) on its own slice of the staged rows.  Compared with one launch per group the R row,
// the gathered h row and Y are staged once per edge, and compared with a single group a thread holds only its group's
// part of the accumulator / cotangent row: half (or a third of) the registers, twice (three times) the warps per edge.
// =====================================================================================================================
struct PipeCtx {
  float* stages;
  uint64_t *full, *empty;
  const int* rowptr;
  TileSched ts;
  int nstage, sfl, rfl, hfl, C, N, mm, tpg, lane, q_lo, q_hi;
  bool is_producer;
};

template <int ID, class TG, class PROD>
__device__ __forceinline__ void fwd_group_body(const PipeCtx& c, PROD& prod, int grp, float* __restrict__ out) {
  using TF = TPCode<ID>;
  using L = PipeLayout<TF>;
  const int C = c.C, lane = c.lane;
  const int u = 2 * ((int)threadIdx.x - grp * c.tpg);
  const bool active = u < C;
  const int uc = active ? u : C - 2;
  int stage = 0, consumed = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < c.ts.ntiles; tile += c.ts.G) {
    const int n0 = c.ts.first_node(tile), n1 = min(c.N, n0 + c.ts.TN);
    int q = __ldg(c.rowptr + n0);
    int end = __ldg(c.rowptr + n0 + 1);
    for (int n = n0; n < n1; ++n) {
      const int end_next = (n + 2 <= c.N) ? __ldg(c.rowptr + n + 2) : end;
      float2 acc[TG::DOUT];
#pragma unroll
      for (int k = 0; k < TG::DOUT; ++k) acc[k] = make_float2(0.f, 0.f);
      for (; q < end; ++q) {
        const float* st = c.stages + (size_t)stage * c.sfl;
        if (c.is_producer && prod.issued == consumed) prod.issue(1, consumed, 1, true, lane);
        mbar_wait(smem_u32(c.full + stage), phase);
        float2 Rv[TG::NP], Yv[TG::NSH], hv[TG::DIN];
#pragma unroll
        for (int p = 0; p < TG::NP; ++p) Rv[p] = *reinterpret_cast<const float2*>(st + (TG::P0 + p) * C + uc);
        TG::load_h2_smem(st + c.rfl + TG::H_OFF * C, C, uc, hv);
        load_y2_smem<TG::NSH, L::YPAD>(st + c.rfl + c.hfl, Yv);
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(c.empty + stage));
        if (++stage == c.nstage) {
          stage = 0;
          phase ^= 1u;
        }
        ++consumed;
        if (c.is_producer) prod.issue(2, consumed, c.nstage, false, lane);
        TG::fwd2(hv, Yv, Rv, acc);
      }
      if (active) {
        if (c.mm & 1) TG::template store_out2<true>(out + (size_t)n * TG::DOUT_FULL * C, C, u, acc);
        else TG::template store_out2<false>(out + (size_t)n * TG::DOUT_FULL * C, C, u, acc);
      }
      end = end_next;
    }
  }
}

template <int ID, int G, class PROD>
__device__ __forceinline__ void fwd_group_dispatch(const PipeCtx& c, PROD& prod, int grp, float* __restrict__ out) {
  if constexpr (G < TPGroups<ID>::N) {
    if (grp == G) fwd_group_body<ID, TPGroup<ID, G>>(c, prod, grp, out);
    else fwd_group_dispatch<ID, G + 1>(c, prod, grp, out);
  }
}

template <int ID, bool DEDICATED>
__global__ void tp_fwd_pipe_g_kernel(const float* __restrict__ h, const float* __restrict__ Y, const float* __restrict__ R,
                                     const int* __restrict__ gather, const int* __restrict__ rowptr,
                                     const int* __restrict__ perm, float* __restrict__ out, int N, int C, int nstage, int mm) {
  pdl_enter();
  using TF = TPCode<ID>;
  using L = PipeLayout<TF>;
  constexpr int NG = TPGroups<ID>::N;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + kPipeMaxStages;
  PipeCtx c;
  c.stages = reinterpret_cast<float*>(smem_raw + kPipeHeaderBytes);
  c.full = full, c.empty = empty, c.rowptr = rowptr, c.nstage = nstage, c.C = C, c.N = N, c.mm = mm;
  c.tpg = L::consumer_threads(C);
  const int nthr_c = NG * c.tpg;
  c.ts = tile_schedule(rowptr, N);
  if ((int)blockIdx.x >= c.ts.G) return;
  pipe_init(full, empty, nstage, nthr_c >> 5);
  __syncthreads();
  c.lane = threadIdx.x & 31;
  c.rfl = L::r_floats(C), c.hfl = L::h_floats(C), c.sfl = L::stage_floats(C);
  c.is_producer = !DEDICATED && threadIdx.x < 32;
  PipeProducer<TF, true> prod;
  prod.h = h, prod.Y = Y, prod.R = R, prod.gather = gather, prod.perm = perm, prod.rowptr = rowptr;
  prod.stages = c.stages, prod.full = full, prod.empty = empty, prod.N = N, prod.C = C, prod.nstage = nstage, prod.ts = c.ts;
  if (DEDICATED && (int)threadIdx.x >= nthr_c) {
    prod.start(c.lane);
    while (!prod.done()) prod.issue(1 << 30, 0, 1 << 30, true, c.lane);
    return;
  }
  if (c.is_producer) {
    prod.start(c.lane);
    prod.issue(nstage, 0, nstage, true, c.lane);
  }
  fwd_group_dispatch<ID, 0>(c, prod, (int)threadIdx.x / c.tpg, out);
}

template <int ID, class TG, bool WANT_H, bool WANT_Y, bool WANT_R, class PROD>
__device__ __forceinline__ void bwd_group_body(const PipeCtx& c, PROD& prod, int grp, const float* __restrict__ g,
                                               float* __restrict__ dh, float* __restrict__ dY, float* __restrict__ dR) {
  using TF = TPCode<ID>;
  using L = PipeLayout<TF>;
  constexpr int NV = next_pow2(TG::NSH);
  constexpr int NG = TPGroups<ID>::N;
  const int C = c.C, lane = c.lane, N = c.N;
  const int u = 2 * ((int)threadIdx.x - grp * c.tpg);
  const bool active = u < C;
  const int uc = active ? u : C - 2;
  const size_t rstride = (size_t)TG::NP_FULL * C, hstride = (size_t)TG::DIN_FULL * C;
  const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
  int stage = 0, consumed = 0;
  uint32_t phase = 0;
  int n;  // node holding edge q_lo
  {
    int lo = 0, hi = N;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(c.rowptr + mid) > c.q_lo) hi = mid; else lo = mid;
    }
    n = hi - 1;
  }
  int q = c.q_lo;
  int end = min(__ldg(c.rowptr + n + 1), c.q_hi);
  for (;; ++n) {
    const int end_next = (n + 2 <= N) ? min(__ldg(c.rowptr + n + 2), c.q_hi) : c.q_hi;
    float2 gv[TG::DOUT];
    if (q < end) {
      if (c.mm & 1) TG::template load_out2<true>(g + (size_t)n * TG::DOUT_FULL * C, C, uc, gv);
      else TG::template load_out2<false>(g + (size_t)n * TG::DOUT_FULL * C, C, uc, gv);
      if (!active) {
#pragma unroll
        for (int k = 0; k < TG::DOUT; ++k) gv[k] = make_float2(0.f, 0.f);
      }
    }
    for (; q < end; ++q) {
      const float* st = c.stages + (size_t)stage * c.sfl;
      if (c.is_producer && prod.issued == consumed) prod.issue(1, consumed, 1, true, lane);
      mbar_wait(smem_u32(c.full + stage), phase);
      const int2 es = reinterpret_cast<const int2*>(st + c.rfl + c.hfl + L::YPAD)[0];  // {edge id, gather id}
      const int e = es.x;
      float2 Rv[TG::NP], Yv[TG::NSH], hv[TG::DIN];
#pragma unroll
      for (int p = 0; p < TG::NP; ++p)
        Rv[p] = (WANT_H || WANT_Y) ? *reinterpret_cast<const float2*>(st + (TG::P0 + p) * C + uc) : make_float2(0.f, 0.f);
      TG::load_h2_smem(st + c.rfl + TG::H_OFF * C, C, uc, hv);
      load_y2_smem<TG::NSH, L::YPAD>(st + c.rfl + c.hfl, Yv);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(c.empty + stage));
      if (++stage == c.nstage) {
        stage = 0;
        phase ^= 1u;
      }
      ++consumed;
      if (c.is_producer) prod.issue(2, consumed, c.nstage, false, lane);

      float2 dhv[TG::DIN], dRv[TG::NP];
      float dYv[NV];
#pragma unroll
      for (int i = 0; i < TG::DIN; ++i) dhv[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NV; ++j) dYv[j] = 0.f;
      TG::template bwd2<WANT_H, WANT_Y, WANT_R, NV>(hv, Yv, Rv, gv, dhv, dYv, dRv);

      if (WANT_R && active) {
        float* drow = dR + (size_t)e * rstride + (size_t)TG::P0 * C + u;
#pragma unroll
        for (int p = 0; p < TG::NP; ++p) {
          if (c.mm & 2) red_add_f32x2(drow + p * C, dRv[p].x, dRv[p].y, pol_stream);
          else st_stream2_hint(drow + p * C, dRv[p], pol_stream);
        }
      }
      if (WANT_H) {
        if (active) TG::atomic_add_h2(dh + (size_t)es.y * hstride, C, u, dhv, pol_keep);
      }
      if (WANT_Y) {
        warp_reduce_vec<NV>(dYv, lane);
        constexpr int sh = 5 - Log2<NV>::value;
        const int j = lane >> sh;
        if ((lane & ((1 << sh) - 1)) == 0 && j < TG::NSH) {
          if (NG == 1 && c.tpg == 32)
            dY[(size_t)e * TG::NSH + j] = dYv[0];
          else
            atomicAdd(dY + (size_t)e * TG::NSH + j, dYv[0]);
        }
      }
    }
    if (q >= c.q_hi) break;
    end = end_next;
  }
}

template <int ID, int G, bool WANT_H, bool WANT_Y, bool WANT_R, class PROD>
__device__ __forceinline__ void bwd_group_dispatch(const PipeCtx& c, PROD& prod, int grp, const float* __restrict__ g,
                                                   float* __restrict__ dh, float* __restrict__ dY, float* __restrict__ dR) {
  if constexpr (G < TPGroups<ID>::N) {
    if (grp == G) bwd_group_body<ID, TPGroup<ID, G>, WANT_H, WANT_Y, WANT_R>(c, prod, grp, g, dh, dY, dR);
    else bwd_group_dispatch<ID, G + 1, WANT_H, WANT_Y, WANT_R>(c, prod, grp, g, dh, dY, dR);
  }
}

template <int ID, bool DEDICATED, bool WANT_H, bool WANT_Y, bool WANT_R>
__global__ void tp_bwd_pipe_g_kernel(const float* __restrict__ h, const float* __restrict__ Y, const float* __restrict__ R,
                                     const float* __restrict__ g, const int* __restrict__ gather,
                                     const int* __restrict__ rowptr, const int* __restrict__ perm, float* __restrict__ dh,
                                     float* __restrict__ dY, float* __restrict__ dR, int N, int C, int nstage, int mm) {
  pdl_enter();
  using TF = TPCode<ID>;
  using L = PipeLayout<TF>;
  constexpr int NG = TPGroups<ID>::N;
  constexpr bool LOAD_R = WANT_H || WANT_Y;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty = full + kPipeMaxStages;
  PipeCtx c;
  c.stages = reinterpret_cast<float*>(smem_raw + kPipeHeaderBytes);
  c.full = full, c.empty = empty, c.rowptr = rowptr, c.nstage = nstage, c.C = C, c.N = N, c.mm = mm;
  c.tpg = L::consumer_threads(C);
  const int nthr_c = NG * c.tpg;
  {  // even split of the edge list between the CTAs, walked from the end
    const long long E_ = __ldg(rowptr + N), b_ = (long long)gridDim.x - 1 - blockIdx.x;
    c.q_lo = (int)(E_ * b_ / gridDim.x), c.q_hi = (int)(E_ * (b_ + 1) / gridDim.x);
    c.ts.TN = 1, c.ts.ntiles = 0, c.ts.G = (int)gridDim.x;
    if (c.q_lo >= c.q_hi) return;
  }
  pipe_init(full, empty, nstage, nthr_c >> 5);
  __syncthreads();
  c.lane = threadIdx.x & 31;
  c.rfl = L::r_floats(C), c.hfl = L::h_floats(C), c.sfl = L::stage_floats(C);
  c.is_producer = !DEDICATED && threadIdx.x < 32;
  PipeProducer<TF, LOAD_R> prod;
  prod.h = h, prod.Y = Y, prod.R = R, prod.gather = gather, prod.perm = perm, prod.rowptr = rowptr;
  prod.stages = c.stages, prod.full = full, prod.empty = empty, prod.N = N, prod.C = C, prod.nstage = nstage, prod.ts = c.ts;
  if (DEDICATED && (int)threadIdx.x >= nthr_c) {
    prod.start_range(c.lane, c.q_lo, c.q_hi);
    while (!prod.done()) prod.issue(1 << 30, 0, 1 << 30, true, c.lane);
    return;
  }
  if (c.is_producer) {
    prod.start_range(c.lane, c.q_lo, c.q_hi);
    prod.issue(nstage, 0, nstage, true, c.lane);
  }
  bwd_group_dispatch<ID, 0, WANT_H, WANT_Y, WANT_R>(c, prod, (int)threadIdx.x / c.tpg, g, dh, dY, dR);
}

struct PipeCfg {
  int nstage, smem, grid;
};

// stage count and persistent grid for a kernel instantiation (cached per (C, small); queries only, no allocation /
// sync).  Large graphs: 4 CTAs per SM with 7-stage rings (the arithmetic of 8 consumer warps already fills the fp32
// pipe; more CTAs only fight over it: cfg4 72% vs 86% of HBM peak).  Small graphs (a training batch of molecules: ~30
// edges per node, ~100 edges per CTA): 5 CTAs per SM with 6-stage rings — the ramp dominates and more concurrent
// streams shorten it (cfg2 59% vs 54%).
template <auto kernel>  // one cache per kernel instantiation
static bool pipe_config(int threads, int stage_bytes, int C, bool small, PipeCfg* cfg) {
  static int cached_C[2 * kMaxDevices];  // zero-initialised: C == 0 never occurs, so 0 means "empty"
  static PipeCfg cached[2 * kMaxDevices];
  static bool cached_ok[2 * kMaxDevices];
  int dev = 0, sms = 0, max_optin = 0;
  cudaGetDevice(&dev);
  const int slot = 2 * device_slot() + (small ? 1 : 0);  // per device: function attributes and SM counts are per device
  if (C == cached_C[slot]) {
    *cfg = cached[slot];
    return cached_ok[slot];
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int nstage = ((small ? 44 : 56) * 1024) / stage_bytes;
  if (nstage > kPipeMaxStages) nstage = kPipeMaxStages;
  if (nstage < 3) nstage = 3;
  int smem = kPipeHeaderBytes + nstage * stage_bytes;
  bool ok = smem <= max_optin;
  int per_sm = 0;
  if (ok) {
    // the attribute is per kernel, not per launch: keep it at the larger of the two configurations
    int want = kPipeHeaderBytes + ((56 * 1024) / stage_bytes < 3 ? 3 : ((56 * 1024) / stage_bytes > kPipeMaxStages ? kPipeMaxStages : (56 * 1024) / stage_bytes)) * stage_bytes;
    if (want < smem) want = smem;
    if (want > max_optin) want = smem;
    ok = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want) == cudaSuccess;
  }
  if (ok) ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) == cudaSuccess && per_sm > 0;
  if (!ok) cudaGetLastError();
  cached_C[slot] = C;
  cached[slot] = PipeCfg{nstage, smem, per_sm * sms};
  cached_ok[slot] = ok;
  *cfg = cached[slot];
  return ok;
}

constexpr int kSmallGraphNodes = 6000;

__host__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pml

#define PML_CAT_(a, b) a##b
#define PML_CAT(a, b) PML_CAT_(a, b)

// test / profiling hook (capi.cu): PML_TP_SIMPLE=1 in the environment selects the non-pipelined kernels
extern "C" int pml_tp_force_simple(void);

namespace pml {

// ---- pipelined launches, one per path group ------------------------------------------------------------------------
template <int ID, int G = 0>
static bool groups_supported(int C) {
  if constexpr (G >= TPGroups<ID>::N) {
    return true;
  } else {
    return PipeLayout<TPGroup<ID, G>>::supported(C) && groups_supported<ID, G + 1>(C);
  }
}

template <class T, bool DEDICATED>
static bool launch_fwd_pipe(const float* h, const float* Y, const float* R, const int* gather, const int* rowptr,
                            const int* perm, float* out, int N, int C, int mm, cudaStream_t stream) {
  using PL = PipeLayout<T>;
  PipeCfg cfg;
  const int threads = PL::consumer_threads(C) + (DEDICATED ? 32 : 0);
  const int sbytes = PL::stage_floats(C) * 4;
  if (!pipe_config<tp_fwd_pipe_kernel<T, DEDICATED>>(threads, sbytes, C, N < kSmallGraphNodes, &cfg)) return false;
  // small graphs: one CTA per node (tile), more CTAs than slots: the hardware hands a free slot the next node, so
  // nodes of different degree balance themselves instead of being paired statically
  // (config 2, 1304 nodes on 740 slots: 65.6 -> 57.3 us, 57.5 -> 65.8 % of HBM peak)
  if (DEDICATED && N < kSmallGraphNodes) cfg.grid = N < 4 * cfg.grid ? N : 4 * cfg.grid;
  pml::launch_plain(tp_fwd_pipe_kernel<T, DEDICATED>, dim3(cfg.grid), dim3(threads), cfg.smem, stream, h, Y, R, gather, rowptr, perm, out, N, C, cfg.nstage, mm);
  return true;
}

template <int ID, int G = 0>
static int max_group_dout() {
  if constexpr (G >= TPGroups<ID>::N) {
    return 0;
  } else {
    const int rest = max_group_dout<ID, G + 1>();
    return TPGroup<ID, G>::DOUT > rest ? TPGroup<ID, G>::DOUT : rest;
  }
}

// all groups inside one CTA (one ring of full rows, TPGroups<ID>::N consumer warp sets)
template <int ID>
static bool intra_supported(int C) {
  using L = PipeLayout<TPCode<ID>>;
  // measured (profiles/r02_k4_path_group_experiments.log): with one consumer warp per group (C <= 64, NequIP) sharing the
  // staged rows wins clearly (forward 129 vs 184 us); with four warps per group (C = 256, MACE large) one launch per
  // group is faster in the forward (355 vs 420 us) and equal in the cotangent: those keep their own launches
  return TPGroups<ID>::N > 1 && L::rows_ok(C) && max_group_dout<ID>() <= 48 && L::consumer_threads(C) == 32 &&
         TPGroups<ID>::N * L::consumer_threads(C) + 32 <= 1024;
}

template <int ID>
static bool launch_fwd_intra(const float* h, const float* Y, const float* R, const int* gather, const int* rowptr,
                             const int* perm, float* out, int N, int C, int mm, cudaStream_t stream) {
  if constexpr (TPGroups<ID>::N > 1) {
    using L = PipeLayout<TPCode<ID>>;
    PipeCfg cfg;
    const int threads = TPGroups<ID>::N * L::consumer_threads(C) + 32;
    const int sbytes = L::stage_floats(C) * 4;
    if (!pipe_config<tp_fwd_pipe_g_kernel<ID, true>>(threads, sbytes, C, N < kSmallGraphNodes, &cfg)) return false;
    if (N < kSmallGraphNodes) cfg.grid = N < 4 * cfg.grid ? N : 4 * cfg.grid;
    pml::launch_plain(tp_fwd_pipe_g_kernel<ID, true>, dim3(cfg.grid), dim3(threads), cfg.smem, stream, h, Y, R, gather, rowptr, perm, out, N, C, cfg.nstage, mm);
    return true;
  } else {
    return false;
  }
}

template <int ID, bool H_, bool Y_, bool R_>
static bool launch_bwd_intra_one(const float* h, const float* Y, const float* R, const float* g, const int* gather,
                                 const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C, int mm,
                                 cudaStream_t stream) {
  if constexpr (TPGroups<ID>::N > 1) {
    using L = PipeLayout<TPCode<ID>>;
    PipeCfg cfg;
    const int threads = TPGroups<ID>::N * L::consumer_threads(C);
    const int sbytes = L::stage_floats(C) * 4;
    if (!pipe_config<tp_bwd_pipe_g_kernel<ID, false, H_, Y_, R_>>(threads, sbytes, C, N < kSmallGraphNodes, &cfg)) return false;
    pml::launch_plain(tp_bwd_pipe_g_kernel<ID, false, H_, Y_, R_>, dim3(cfg.grid), dim3(threads), cfg.smem, stream, h, Y, R, g, gather, rowptr, perm, dh, dY,
                                                                                         dR, N, C, cfg.nstage, mm);
    return true;
  } else {
    return false;
  }
}

template <int ID>
static bool launch_bwd_intra(int mask, const float* h, const float* Y, const float* R, const float* g, const int* gather,
                             const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C, int mm,
                             cudaStream_t stream) {
#define PML_TP_BWD_INTRA(M, H_, Y_, R_) \
  case M:                               \
    return launch_bwd_intra_one<ID, H_, Y_, R_>(h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream);
  switch (mask) {
    PML_TP_BWD_INTRA(1, true, false, false)
    PML_TP_BWD_INTRA(2, false, true, false)
    PML_TP_BWD_INTRA(3, true, true, false)
    PML_TP_BWD_INTRA(4, false, false, true)
    PML_TP_BWD_INTRA(5, true, false, true)
    PML_TP_BWD_INTRA(6, false, true, true)
    PML_TP_BWD_INTRA(7, true, true, true)
  }
#undef PML_TP_BWD_INTRA
  return false;
}

template <int ID, int G = 0>
static bool fwd_pipe_groups(bool dedicated, const float* h, const float* Y, const float* R, const int* gather,
                            const int* rowptr, const int* perm, float* out, int N, int C, int mm, cudaStream_t stream) {
  if constexpr (G >= TPGroups<ID>::N) {
    return true;
  } else {
    using T = TPGroup<ID, G>;
    bool ok;
    if constexpr (TPGroups<ID>::N == 1) {  // the producer-placement flip is a profiling aid of single-group signatures
      ok = dedicated ? launch_fwd_pipe<T, true>(h, Y, R, gather, rowptr, perm, out, N, C, mm, stream)
                     : launch_fwd_pipe<T, false>(h, Y, R, gather, rowptr, perm, out, N, C, mm, stream);
    } else {
      ok = launch_fwd_pipe<T, true>(h, Y, R, gather, rowptr, perm, out, N, C, mm, stream);
    }
    return ok && fwd_pipe_groups<ID, G + 1>(dedicated, h, Y, R, gather, rowptr, perm, out, N, C, mm, stream);
  }
}

template <class T, bool D, bool H_, bool Y_, bool R_>
static bool launch_bwd_pipe_one(const float* h, const float* Y, const float* R, const float* g, const int* gather,
                                const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C, int mm,
                                cudaStream_t stream) {
  using PL = PipeLayout<T>;
  PipeCfg cfg;
  const int threads = PL::consumer_threads(C) + (D ? 32 : 0);
  const int sbytes = PL::stage_floats(C) * 4;
  if (!pipe_config<tp_bwd_pipe_kernel<T, D, H_, Y_, R_>>(threads, sbytes, C, N < kSmallGraphNodes, &cfg)) return false;
  pml::launch_plain(tp_bwd_pipe_kernel<T, D, H_, Y_, R_>, dim3(cfg.grid), dim3(threads), cfg.smem, stream, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N,
                                                                                 C, cfg.nstage, mm);
  return true;
}

template <class T, bool D>
static bool launch_bwd_pipe(int mask, const float* h, const float* Y, const float* R, const float* g, const int* gather,
                            const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C, int mm,
                            cudaStream_t stream) {
#define PML_TP_BWD_PIPE(M, H_, Y_, R_) \
  case M:                              \
    return launch_bwd_pipe_one<T, D, H_, Y_, R_>(h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream);
  switch (mask) {
    PML_TP_BWD_PIPE(1, true, false, false)
    PML_TP_BWD_PIPE(2, false, true, false)
    PML_TP_BWD_PIPE(3, true, true, false)
    PML_TP_BWD_PIPE(4, false, false, true)
    PML_TP_BWD_PIPE(5, true, false, true)
    PML_TP_BWD_PIPE(6, false, true, true)
    PML_TP_BWD_PIPE(7, true, true, true)
  }
#undef PML_TP_BWD_PIPE
  return false;
}

template <int ID, int G = 0>
static bool bwd_pipe_groups(bool dedicated, int mask, const float* h, const float* Y, const float* R, const float* g,
                            const int* gather, const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N,
                            int C, int mm, cudaStream_t stream) {
  if constexpr (G >= TPGroups<ID>::N) {
    return true;
  } else {
    using T = TPGroup<ID, G>;
    bool ok;
    if constexpr (TPGroups<ID>::N == 1) {
      ok = dedicated ? launch_bwd_pipe<T, true>(mask, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream)
                     : launch_bwd_pipe<T, false>(mask, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream);
    } else {
      ok = launch_bwd_pipe<T, false>(mask, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream);
    }
    return ok && bwd_pipe_groups<ID, G + 1>(dedicated, mask, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream);
  }
}

}  // namespace pml

// Launchers (one pair per compiled signature; dispatched by id in capi.cu).  Streams are the caller's.
extern "C" int PML_CAT(pml_tp_fwd_launch_, PML_TP_ID)(const float* h, const float* Y, const float* R, const int* gather,
                                                       const int* rowptr, const int* perm, float* out, int N, int C,
                                                       int mm, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  // mode bit0: streaming kernels; bit1 / bit2: flip the fwd / bwd producer placement; bit3: one launch per path group
  // instead of all groups inside one CTA
  const int mode = pml_tp_force_simple();
  if (!(mode & (1 | 8)) && pml::intra_supported<PML_TP_ID>(C) && pml::aligned16(h) && pml::aligned16(R)) {
    if (pml::launch_fwd_intra<PML_TP_ID>(h, Y, R, gather, rowptr, perm, out, N, C, mm, stream)) {
      PML_CHECK_LAUNCH();
      return PML_OK;
    }
    cudaGetLastError();
  }
  if (pml::groups_supported<PML_TP_ID>(C) && pml::aligned16(h) && pml::aligned16(R) && !(mode & 1)) {
    if (pml::fwd_pipe_groups<PML_TP_ID>(!(mode & 2), h, Y, R, gather, rowptr, perm, out, N, C, mm, stream)) {
      PML_CHECK_LAUNCH();
      return PML_OK;
    }
    cudaGetLastError();
  }
  const int nslab = (C + 31) / 32;
  const int wpb = 4;
  const long long items = (long long)N * nslab;
  pml::launch_plain(pml::tp_fwd_kernel<PML_TP_ID>, dim3(pml::ceil_div(items, wpb)), dim3(wpb * 32), 0, stream, h, Y, R, gather, rowptr, perm, out, N, C, nslab, mm);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// dh and dY must be zero-initialised by the caller when C > 32 or the signature has several path groups (pml_tp_num_groups);
// dR is fully overwritten.
extern "C" int PML_CAT(pml_tp_bwd_launch_, PML_TP_ID)(const float* h, const float* Y, const float* R, const float* g,
                                                       const int* gather, const int* rowptr, const int* perm, float* dh,
                                                       float* dY, float* dR, int N, int C, int mm, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  const int mask = (dh ? 1 : 0) | (dY ? 2 : 0) | (dR ? 4 : 0);
  const int mode = pml_tp_force_simple();
  if (mask != 0 && !(mode & (1 | 8)) && pml::intra_supported<PML_TP_ID>(C) && pml::aligned16(h) && pml::aligned16(R)) {
    if (pml::launch_bwd_intra<PML_TP_ID>(mask, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream)) {
      PML_CHECK_LAUNCH();
      return PML_OK;
    }
    cudaGetLastError();
  }
  if (mask != 0 && pml::groups_supported<PML_TP_ID>(C) && pml::aligned16(h) && pml::aligned16(R) && !(mode & 1)) {
    if (pml::bwd_pipe_groups<PML_TP_ID>((mode & 4) != 0, mask, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, mm, stream)) {
      PML_CHECK_LAUNCH();
      return PML_OK;
    }
    cudaGetLastError();
  }
  const int nslab = (C + 31) / 32;
  const int wpb = 4;
  const long long items = (long long)N * nslab;
  const int grid = pml::ceil_div(items, wpb);
#define PML_TP_BWD_CASE(M, H_, Y_, R_)                                                                              \
  case M:                                                                                                           \
    pml::launch_plain(pml::tp_bwd_kernel<PML_TP_ID, H_, Y_, R_>, dim3(grid), dim3(wpb * 32), 0, stream, h, Y, R, g, gather, rowptr, perm, dh, dY, \
                                                                            dR, N, C, nslab, mm);                   \
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
