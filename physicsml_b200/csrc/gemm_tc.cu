// Tensor-core (tcgen05, TMEM accumulators) version of the grouped strided GEMM, with 3xTF32 error compensation so the
// result keeps fp32 accuracy (the parity bar is 1e-5 relative on energies, which single-pass TF32 cannot hold):
//     A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi ,   X_hi = tf32(X), X_lo = tf32(X - X_hi)
//
// Used for the radial MLP (e3nn nn.FullyConnectedNet; reference blocks.py:157-160,211, interaction_block.py:61-64,85)
// and for every other dense contraction large enough to fill 128-row MMA tiles.  Same problem table and semantics as
// gemm.cu (general strides, K-chunks, batches, split-K) so both kernels are interchangeable and cross-checked.
//
// Structure per CTA (one 128 x BN output tile, 256 threads):
//   * all threads stage A/B k-blocks: global (arbitrary strides) -> registers -> hi/lo split -> shared memory in the
//     canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices), 2-stage ring;
//   * one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) x 3 products per k-step, accumulating in TMEM — the
//     hi*hi products into one accumulator, the two small cross terms into a second one (see TcSmem / kPromoK);
//     tcgen05.commit arrives on the stage's mbarrier to hand the stage back to the loaders;
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> scale -> global (plain / accumulate / atomic).
// No TMA: operands must pass through registers anyway for the hi/lo split.
#include "gemm_problem.h"
#include "pml_common.cuh"

namespace pml {

constexpr int TC_BM = 128;  // the k-block BK (16 or 32) is a template parameter
// Threads per CTA.  The 32-wide k-block variant runs where a CTA has an SM to itself (grids of one wave) and takes 512
// threads: half the units per thread, four warps per scheduler instead of two.  Measured (scripts/bench_small_gemm.py,
// 1 304 x K x 128): K = 64 8.45 -> 8.16 us, K = 128 11.8 -> 11.3 us - a step still costs ~1.6 us per 32 k.
template <int BK> __host__ __device__ constexpr int tc_threads() { return BK == 32 ? 512 : 256; }

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

// x = hi + lo with hi = x truncated to tf32 (13 low mantissa bits cleared) and lo = x - hi, exact in fp32.  The tensor
// core reads the upper 19 bits of an operand word, so lo needs no conversion of its own; its truncation to tf32 costs
// 2^-21 relative, the same order as the A_lo B_lo product 3xTF32 drops anyway.  Two full-rate integer / fp32 ops per
// element instead of two cvt.rna.tf32 (a reduced-rate conversion pipe shared by the whole SM).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

// K-major, no swizzle: 16-byte unit (r, kc) of a [ROWS x BK] tile lives at  (r % 8) + 8 * kc + 2 BK * (r / 8)
//   => leading (K) byte offset = 128, stride (M/N) byte offset = 32 BK (512 at BK = 16, 1024 at BK = 32)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;        // leading_byte_offset
  d |= (uint64_t)(sbo_bytes >> 4) << 32;  // stride_byte_offset
  d |= (uint64_t)1 << 46;           // descriptor version (Blackwell)
  return d;                         // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// v[0..31] += 32 TMEM columns, eight at a time (a second 32-register staging array would cost the kernels a resident CTA)
__device__ __forceinline__ void tmem_add32(uint32_t taddr, float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + 8 * c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) v[8 * c + j] += __uint_as_float(r[j]);
  }
}

// Operand staging is split in two so the global loads of k-block i+1 are in flight while k-block i is converted,
// stored and multiplied (register double buffering):
//   fetch : element (row, k) at base + row*rs + k*cs -> registers ; rows >= nrows and k >= K read as 0
//   commit: registers -> hi/lo split -> canonical UMMA shared-memory layout
template <int ROWS, int BK>
struct TileRegs {
  static constexpr int UNITS = ROWS * (BK / 4);
  static constexpr int NT = tc_threads<BK>();
  static constexpr int PER_THREAD = (UNITS + NT - 1) / NT;
  float v[PER_THREAD][4];
};

template <int ROWS, int BK>
__device__ __forceinline__ void tile_coords(int idx, bool kfast, int& r, int& kc) {
  constexpr int KU = BK / 4;  // 16-byte k-units per row
  if (kfast) {
    // 8 consecutive lanes = 8 consecutive rows of one 16-byte k-unit = 8 consecutive units (128 contiguous bytes) of the
    // canonical layout: conflict-free 128-bit shared-memory accesses.  (Lanes along k hit units 8 apart = the same 4
    // banks: ncu counted 73 % of the kernel's shared-memory wavefronts as bank conflicts, which serialised the CTAs of
    // an SM on the shared-memory pipe.)  A warp still covers 8 rows x 64 contiguous bytes in global memory.
    static_assert(KU == 4 || KU == 8, "k-units per row");
    r = (idx & 7) + 8 * (idx / (8 * KU));
    kc = (idx >> 3) & (KU - 1);
  } else {
    r = idx % ROWS;
    kc = idx / ROWS;
  }
}

// What a thread stages of a [ROWS x 16] operand tile never changes during the kernel: which 16-byte units (row, kc), where
// they sit in the shared-memory image and where their row starts in global memory.  Computed once; a k-block then costs
// one 64-bit add and one load per unit.  (Recomputing coordinates, strides and bounds per element and k-block made the
// kernel instruction-bound: 3 500 instructions per warp for a 128 x 128 x 128 tile, 29 % issue utilisation with two
// warps per scheduler - profiles/r02_ncu_small_gemm.txt.)
template <int ROWS, int BK>
struct Lanes {
  static constexpr int UNITS = ROWS * (BK / 4);
  static constexpr int NT = tc_threads<BK>();
  static constexpr int PER_THREAD = (UNITS + NT - 1) / NT;
  long long off[PER_THREAD];  // row * row_stride + 4 kc * k_stride (elements, from the chunk's base)
  int unit[PER_THREAD];       // float4 index inside the stage image; -1: this thread has no such unit
  unsigned ok;                // bit i: unit i exists and its row is inside the matrix
};

template <int ROWS, int BK>
__device__ __forceinline__ void setup_lanes(bool kfast, int row0, int nrows, long long rs, long long cs, Lanes<ROWS, BK>& L) {
  L.ok = 0u;
#pragma unroll
  for (int it = 0; it < Lanes<ROWS, BK>::PER_THREAD; ++it) {
    const int idx = threadIdx.x + it * Lanes<ROWS, BK>::NT;
    L.unit[it] = -1;
    L.off[it] = 0;
    if (Lanes<ROWS, BK>::UNITS % Lanes<ROWS, BK>::NT != 0 && idx >= Lanes<ROWS, BK>::UNITS) continue;
    int r, kc;
    tile_coords<ROWS, BK>(idx, kfast, r, kc);
    L.unit[it] = (r & 7) + 8 * kc + 2 * BK * (r >> 3);
    const int gr = row0 + r;
    L.off[it] = (long long)gr * rs + (long long)(4 * kc) * cs;
    if (gr < nrows) L.ok |= 1u << it;
  }
}

// element (row, k) of the chunk at base + row*rs + k*cs ; rows >= nrows and k >= K read as 0
template <int ROWS, int BK>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ base, long long cs, int k0, int K, bool vec,
                                           const Lanes<ROWS, BK>& L, TileRegs<ROWS, BK>& t) {
  const bool full = k0 + BK <= K;  // uniform: only the last k-block of a chunk checks k per element
  const float* b0 = base + (long long)k0 * cs;
#pragma unroll
  for (int it = 0; it < Lanes<ROWS, BK>::PER_THREAD; ++it) {
    t.v[it][0] = t.v[it][1] = t.v[it][2] = t.v[it][3] = 0.f;
    if (!((L.ok >> it) & 1u)) continue;
    const float* p = b0 + L.off[it];
    if (full) {
      if (vec) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        t.v[it][0] = q.x; t.v[it][1] = q.y; t.v[it][2] = q.z; t.v[it][3] = q.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) t.v[it][j] = __ldg(p + (long long)j * cs);
      }
    } else {
      const int gk = k0 + 4 * ((L.unit[it] >> 3) & (BK / 4 - 1));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gk + j < K) t.v[it][j] = __ldg(p + (long long)j * cs);
    }
  }
}

template <int ROWS, int BK>
__device__ __forceinline__ void commit_tile(const TileRegs<ROWS, BK>& t, const Lanes<ROWS, BK>& L, float scale,
                                            float* __restrict__ s_hi, float* __restrict__ s_lo) {
  const bool scaled = scale != 1.f;  // uniform
#pragma unroll
  for (int it = 0; it < Lanes<ROWS, BK>::PER_THREAD; ++it) {
    if (Lanes<ROWS, BK>::UNITS % Lanes<ROWS, BK>::NT != 0 && L.unit[it] < 0) continue;
    float x0 = t.v[it][0], x1 = t.v[it][1], x2 = t.v[it][2], x3 = t.v[it][3];
    if (scaled) x0 *= scale, x1 *= scale, x2 *= scale, x3 *= scale;
    float4 hi, lo;
    split_tf32(x0, hi.x, lo.x);
    split_tf32(x1, hi.y, lo.y);
    split_tf32(x2, hi.z, lo.z);
    split_tf32(x3, hi.w, lo.w);
    reinterpret_cast<float4*>(s_hi)[L.unit[it]] = hi;
    reinterpret_cast<float4*>(s_lo)[L.unit[it]] = lo;
  }
}

constexpr int TC_STAGES = 2;

template <int BN, int BK>
struct __align__(128) TcSmem {
  float a[TC_STAGES][2][TC_BM * BK];
  float b[TC_STAGES][2][BN * BK];
  float stage_pad[(TC_BM * (BN + 4) > TC_STAGES * 2 * (TC_BM + BN) * BK) ? (TC_BM * (BN + 4) - TC_STAGES * 2 * (TC_BM + BN) * BK) : 4];
  unsigned long long bar[TC_STAGES];  // ring: the MMAs that read a stage have retired
  unsigned long long accbar[2];       // an accumulation period has retired into TMEM buffer 0 / 1
  uint32_t tmem_base;
  pml_gemm_problem prob;              // this CTA's problem descriptor (copied once: the table lives in global memory and
                                      // every "memory"-clobbering barrier op would otherwise force its fields to be re-read)
};

// tcgen05 accumulates in fp32 but aligns and TRUNCATES every partial sum, a bias of ~2^-24 of the running sum per MMA
// (1e-5 relative after the 480 MMAs of K = 1280).  With PROMOTE the K loop is cut into periods of kPromoK / BK
// k-blocks (K = 128): each period accumulates into one of two TMEM buffers from zero and is then added, with ordinary
// round-to-nearest fp32 adds, into per-thread register accumulators while the next period's MMAs run into the other
// buffer.  The result is deterministic and keeps ~1e-6 relative accuracy at any K.
//
// The two cross terms A_lo*B_hi + A_hi*B_lo are ~2^-11 of the hi*hi sum.  Added into the SAME accumulator they cost two of
// the three truncations per k-step; they go to a second TMEM accumulator instead (their own truncation is relative to
// their small magnitude, i.e. negligible) and the two are added in registers at the end: one truncation per k-step
// instead of three (K = 128: 16 instead of 48, ~1e-6 instead of ~3e-6 relative bias) for no extra MMA.
constexpr int kPromoK = 128;     // reduction length of one accumulation period
constexpr int kPlainMaxK = 256;  // longest reduction the non-promoting kernels take

template <int N32>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
#pragma unroll
  for (int c = 0; c < N32; ++c) {
    float t[32];
    tmem_ld32(taddr + 32 * c, t);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[32 * c + j] = t[j];
  }
}

// BK: reduction elements per pipeline step (shared-memory stage, barrier round trip, MMA issue).  Every step costs the same
// ~1 us of dependent instructions, fences and barriers whatever it multiplies (scripts/bench_small_gemm.py: 0.95 us per
// 16-wide k-block at N = 32 and at N = 128, unchanged by deeper prefetch or more stages), so grids of a few CTAs per
// SM - the node-level problems of a molecule batch - run BK = 32: half the steps.  Grids that fill the GPU keep BK = 16:
// 48 KB of shared memory and 80 registers let three CTAs share an SM and overlap each other's steps.
template <int BN, bool PROMOTE, int BK>
__global__ void __launch_bounds__(tc_threads<BK>(), (BK == 16 && !PROMOTE && BN <= 64) ? 3 : (BK == 16 ? 2 : 1))
gemm_grouped_tc_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                     float* __restrict__ C,
                                                                     const pml_gemm_problem* __restrict__ probs,
                                                                     int M_runtime, int splitk) {
  pdl_enter();  // before anything is read: the problem table itself may have been written by the previous kernel
  constexpr int NT = tc_threads<BK>();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kPromoBlocks = kPromoK / BK;
  TcSmem<BN, BK>& S = *reinterpret_cast<TcSmem<BN, BK>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const pml_gemm_problem& Pg = probs[blockIdx.y];
    const int Mg = Pg.M < 0 ? M_runtime : Pg.M;
    const int tiles_ng = (Pg.N + BN - 1) / BN;
    const int tmg = blockIdx.x / tiles_ng;
    if (tmg * TC_BM >= Mg || (int)(blockIdx.z / splitk) >= Pg.nbatch) return;  // uniform per CTA, before any barrier / allocation
    static_assert(sizeof(pml_gemm_problem) % 4 == 0, "descriptor is copied word by word");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&Pg);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S.prob);
    for (int i = tid; i < (int)(sizeof(pml_gemm_problem) / 4); i += NT) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const pml_gemm_problem& P = S.prob;
  const int M = P.M < 0 ? M_runtime : P.M;
  const int N = P.N;
  const int tiles_n = (N + BN - 1) / BN;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x - tm * tiles_n;
  int zb = blockIdx.z;
  const int ks = zb % splitk;
  zb /= splitk;
  // loop-invariant descriptor fields, held in registers
  const long long a_rs = P.a_rs, a_cs = P.a_cs, b_rs = P.b_rs, b_cs = P.b_cs;
  const int nchunks = P.nchunks;
  const long long a_boff = (long long)zb * P.a_bs, b_boff = (long long)zb * P.b_bs;
  const bool a_kfast = a_cs == 1, b_kfast = b_rs == 1;
  const bool a_vec_base = a_kfast && ((a_rs & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec_base = b_kfast && ((b_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  constexpr int NBUF = PROMOTE ? 2 : 1;
  constexpr uint32_t CORR = NBUF * BN;  // column offset of the cross-term accumulators (one per main buffer)
  constexpr uint32_t TMEM_COLS = (2 * NBUF * BN) < 32 ? 32 : (2 * NBUF * BN);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(smem_u32(&S.bar[s]), 1);
    mbar_init(smem_u32(&S.accbar[0]), 1);
    mbar_init(smem_u32(&S.accbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  // instruction descriptor: D=f32, A=B=tf32, both K-major, N = BN, M = 128
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  const int row0 = tm * TC_BM, col0 = tn * BN;

  // flattened (chunk, k0) iteration
  struct KPos {
    int ch, k0;
  };
  auto chunk_k = [&](int ch) { return P.kc[ch] < 0 ? M_runtime : P.kc[ch]; };
  auto first_pos = [&](int ch_from) {
    KPos q{ch_from, ks * BK};
    while (q.ch < nchunks && q.k0 >= chunk_k(q.ch)) ++q.ch;
    return q;
  };
  auto next_pos = [&](KPos q) {
    q.k0 += BK * splitk;
    if (q.k0 >= chunk_k(q.ch)) {
      KPos n = first_pos(q.ch + 1);
      return n;
    }
    return q;
  };
  Lanes<TC_BM, BK> LA;
  Lanes<BN, BK> LB;
  setup_lanes<TC_BM, BK>(a_kfast, row0, M, a_rs, a_cs, LA);
  // B operand as an [N x K] K-major tile: element (n, k) = B[k*b_rs + n*b_cs]
  setup_lanes<BN, BK>(b_kfast, col0, N, b_cs, b_rs, LB);
  auto fetch = [&](const KPos& q, TileRegs<TC_BM, BK>& ra, TileRegs<BN, BK>& rb) {
    const long long aoff = P.a_off[q.ch] + a_boff;
    const long long boff = P.b_off[q.ch] + b_boff;
    const int K = chunk_k(q.ch);
    fetch_tile<TC_BM, BK>(A + aoff, a_cs, q.k0, K, a_vec_base && ((aoff & 3) == 0), LA, ra);
    fetch_tile<BN, BK>(B + boff, b_rs, q.k0, K, b_vec_base && ((boff & 3) == 0), LB, rb);
  };

  // this thread's slice of the output tile: TMEM lane quadrant q (rows 32q + lane), column half `half`
  // warp w reads TMEM lane quadrant w % 4 (hardware rule); the warps of a quadrant split the columns between them
  constexpr int NHALF = (NT / 128) < (BN / 32) ? (NT / 128) : (BN / 32);
  constexpr int COLS_PER_WARP = BN / NHALF;
  const int q = warp & 3, half = warp >> 2;
  float acc[PROMOTE ? COLS_PER_WARP : 1];
  if (PROMOTE) {
#pragma unroll
    for (int j = 0; j < COLS_PER_WARP; ++j) acc[j] = 0.f;
  }
  auto promote = [&](int period) {  // add the retired period's TMEM buffer into the register accumulators
    const int buf = period & 1;
    mbar_wait(smem_u32(&S.accbar[buf]), (period >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (half < NHALF) {
      const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * COLS_PER_WARP);
      if (PROMOTE) {
#pragma unroll
        for (int c32 = 0; c32 < COLS_PER_WARP / 32; ++c32) {
          tmem_add32(t0 + CORR + 32 * c32, acc + (PROMOTE ? 32 * c32 : 0));  // the small cross terms first
          tmem_add32(t0 + 32 * c32, acc + (PROMOTE ? 32 * c32 : 0));
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  // two k-blocks of register prefetch: the loads of k-block i+2 are issued while block i is converted and multiplied
  TileRegs<TC_BM, BK> ra[2];
  TileRegs<BN, BK> rb[2];
  KPos pos[3];
  pos[0] = first_pos(0);
  pos[1] = pos[0].ch < nchunks ? next_pos(pos[0]) : pos[0];
  if (pos[0].ch < nchunks) fetch(pos[0], ra[0], rb[0]);
  if (pos[1].ch < nchunks) fetch(pos[1], ra[1], rb[1]);
  int it = 0;  // k-block counter (ring position)
  int to_promote = -1;
#pragma unroll 1
  while (pos[0].ch < nchunks) {
    pos[2] = pos[1].ch < nchunks ? next_pos(pos[1]) : pos[1];
    const int s = it % TC_STAGES;
    if (it >= TC_STAGES) mbar_wait(smem_u32(&S.bar[s]), ((it / TC_STAGES) - 1) & 1);  // MMAs that read stage s retired
    if ((it & 1) == 0) {  // statically indexed register sets
      commit_tile<TC_BM, BK>(ra[0], LA, P.cscale[pos[0].ch], S.a[s][0], S.a[s][1]);
      commit_tile<BN, BK>(rb[0], LB, 1.f, S.b[s][0], S.b[s][1]);
      if (pos[2].ch < nchunks) fetch(pos[2], ra[0], rb[0]);
    } else {
      commit_tile<TC_BM, BK>(ra[1], LA, P.cscale[pos[0].ch], S.a[s][0], S.a[s][1]);
      commit_tile<BN, BK>(rb[1], LB, 1.f, S.b[s][0], S.b[s][1]);
      if (pos[2].ch < nchunks) fetch(pos[2], ra[1], rb[1]);
    }
    if (PROMOTE && to_promote >= 0) {  // the period closed one k-block ago: its MMAs have had an iteration to retire
      promote(to_promote);
      to_promote = -1;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    const int period = PROMOTE ? it / kPromoBlocks : 0;
    const bool opens = PROMOTE ? (it % kPromoBlocks == 0) : (it == 0);
    const bool closes = PROMOTE ? ((it % kPromoBlocks == kPromoBlocks - 1) || pos[1].ch >= nchunks) : (pos[1].ch >= nchunks);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(S.a[s][0]), a_lo = smem_u32(S.a[s][1]);
      const uint32_t b_hi = smem_u32(S.b[s][0]), b_lo = smem_u32(S.b[s][1]);
      const uint32_t d = tmem + (uint32_t)((period & 1) * BN);
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk) {
        const uint32_t o = kk * 256;  // two 16-byte k-units = two core matrices of 128 B
        const uint32_t first = (!opens || kk > 0) ? 1u : 0u;
        constexpr uint32_t SBO = 32 * BK;
        umma_tf32(d + CORR, make_desc(a_lo + o, SBO), make_desc(b_hi + o, SBO), IDESC, first);
        umma_tf32(d + CORR, make_desc(a_hi + o, SBO), make_desc(b_lo + o, SBO), IDESC, 1u);
        umma_tf32(d, make_desc(a_hi + o, SBO), make_desc(b_hi + o, SBO), IDESC, first);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.bar[s]))
                   : "memory");
      if (closes)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&S.accbar[period & 1]))
                     : "memory");
    }
    if (PROMOTE && closes) to_promote = period;
    pos[0] = pos[1];
    pos[1] = pos[2];
    ++it;
  }

  if (it > 0) {
    float* Cb = C + P.c_off + (long long)zb * P.c_bs;
    const bool atomic = splitk > 1 || (P.c_bs == 0 && P.nbatch > 1);
    const bool direct = !atomic && P.c_cs == 1 && ((P.c_rs & 3) == 0) && (((P.c_off + (long long)zb * P.c_bs) & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (N % 4 == 0);
    if (PROMOTE) {
      if (to_promote >= 0) promote(to_promote);
    } else {
      mbar_wait(smem_u32(&S.accbar[0]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int r = row0 + q * 32 + lane;
    const float alpha = P.alpha;
    const bool accumulate = P.accumulate != 0;
    const long long c_rs = P.c_rs, c_cs = P.c_cs;
    if (direct) {
      // row-major C: the whole tile goes through shared memory once (pitch BN + 4 floats: conflict-free 128-bit accesses
      // both by rows, as written, and along a row, as read) and leaves as full rows, one 16-byte vector per lane = up to
      // 512 contiguous bytes per warp instruction.  (Storing straight from registers, 16 bytes per lane into 32
      // different rows, cost 32 L1 wavefronts per instruction: the epilogue alone took 110 of the 179 us of the
      // 41 570 x 64 -> 1280 forward.)
      constexpr int SP = BN + 4;
      float* stg = reinterpret_cast<float*>(S.a);  // [128][SP]: the operand ring (all MMAs retired) + stage_pad
      static_assert(sizeof(float) * TC_BM * SP <= sizeof(S.a) + sizeof(S.b) + sizeof(S.stage_pad), "staging tile does not fit");
      __syncthreads();  // nobody still converts into the ring
      if (half < NHALF) {
#pragma unroll 1
        for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
          float v[32];
          if (PROMOTE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = acc[PROMOTE ? (c0 + j) : 0];
          } else {
            const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * COLS_PER_WARP + c0);
            tmem_ld32(t0, v);
            tmem_add32(t0 + CORR, v);
          }
          float* srow = stg + (q * 32 + lane) * SP + half * COLS_PER_WARP + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(srow + j) = make_float4(alpha * v[j], alpha * v[j + 1], alpha * v[j + 2], alpha * v[j + 3]);
        }
      }
      __syncthreads();
      constexpr int VPR = BN / 4;              // 16-byte vectors per row
      constexpr int RPI = 32 / VPR > 0 ? 32 / VPR : 1;  // rows per warp instruction (BN = 128: 1, 64: 2, 32: 4)
      const int vr = lane / VPR, vc = lane % VPR;
      for (int rr = warp * RPI + vr; rr < TC_BM; rr += (NT / 32) * RPI) {
        const int gr = row0 + rr, gc = col0 + vc * 4;
        if (gr < M && gc < N) {
          float4 o = *reinterpret_cast<const float4*>(stg + rr * SP + vc * 4);
          float* p = Cb + (long long)gr * c_rs + gc;
          if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(p);
            o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
          }
          *reinterpret_cast<float4*>(p) = o;
        }
      }
    } else {
      // general strides / atomics: every 32-column group is transposed through shared memory (the operand ring is free
      // now) so that global stores / atomics run along the columns of a row
      constexpr int PITCH = 32 * NHALF + 1;
      float* stg = reinterpret_cast<float*>(S.a);  // [128][PITCH] floats <= 33.3 KB, inside the operand ring
      static_assert(sizeof(float) * TC_BM * PITCH <= sizeof(S.a) + sizeof(S.b), "staging tile does not fit");
      __syncthreads();  // every MMA has retired (waited above), nobody still converts into the ring
#pragma unroll 1
      for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
        if (half < NHALF) {
          float v[32];
          if (PROMOTE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = acc[PROMOTE ? (c0 + j) : 0];
          } else {
            const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * COLS_PER_WARP + c0);
            tmem_ld32(t0, v);
            tmem_add32(t0 + CORR, v);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[(q * 32 + lane) * PITCH + half * 32 + j] = v[j];
        }
        __syncthreads();
        for (int rr = warp; rr < TC_BM; rr += NT / 32) {
          const int gr = row0 + rr;
          if (gr >= M) break;
#pragma unroll
          for (int h = 0; h < NHALF; ++h) {
            const int gc = col0 + h * COLS_PER_WARP + c0 + lane;
            if (gc < N) {
              float* p = Cb + (long long)gr * c_rs + (long long)gc * c_cs;
              const float val = alpha * stg[rr * PITCH + h * 32 + lane];
              if (atomic) atomicAdd(p, val);
              else if (accumulate) *p += val;
              else *p = val;
            }
          }
        }
        __syncthreads();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, bool PROMOTE, int BK>
static int launch_tc_bk(const float* A, const float* B, float* C, const pml_gemm_problem* probs, int nprobs, int M_runtime,
                           dim3 grid, int splitk, cudaStream_t stream) {
  const size_t smem = sizeof(TcSmem<BN, BK>) + 128;
  static bool configured[kMaxDevices] = {};
  const int slot = device_slot();
  if (!configured[slot]) {
    if (cudaFuncSetAttribute(gemm_grouped_tc_kernel<BN, PROMOTE, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return PML_ERR_LAUNCH;
    configured[slot] = true;
  }
  launch_pdl(gemm_grouped_tc_kernel<BN, PROMOTE, BK>, grid, dim3(tc_threads<BK>()), smem, stream, A, B, C, probs, M_runtime, splitk);
  return PML_OK;
}

template <int BN, bool PROMOTE>
static int launch_tc(const float* A, const float* B, float* C, const pml_gemm_problem* probs, int nprobs, int M_runtime,
                     int max_M, int max_N, int max_batch, int splitk, cudaStream_t stream) {
  const int tiles_m = (max_M + TC_BM - 1) / TC_BM, tiles_n = (max_N + BN - 1) / BN;
  const long long gz = (long long)max_batch * splitk;
  if (gz > 65535 || nprobs > 65535) return PML_ERR_BAD_ARG;
  dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)nprobs, (unsigned)gz);
  static int sms[kMaxDevices] = {};
  const int slot = device_slot();
  if (sms[slot] == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms[slot], cudaDevAttrMultiProcessorCount, dev);
  }
  // BK = 32 runs 512 threads at up to 126 registers: one CTA per SM - used where the whole grid fits in one such wave
  if ((long long)grid.x * grid.y * grid.z <= (long long)sms[slot])
    return launch_tc_bk<BN, PROMOTE, 32>(A, B, C, probs, nprobs, M_runtime, grid, splitk, stream);
  return launch_tc_bk<BN, PROMOTE, 16>(A, B, C, probs, nprobs, M_runtime, grid, splitk, stream);
}

}  // namespace pml

// Same contract as pml_gemm_grouped (gemm.cu) plus max_k: an upper bound of the reduction length one CTA sees (sum of
// the chunk extents of a problem, divided by splitk); <= 0 means unknown.  Reductions longer than 256 run the kernel
// variant that promotes the TMEM accumulators into registers every 128 k (see kPromoK).
extern "C" int pml_gemm_grouped_tc(const float* A, const float* B, float* C, const pml_gemm_problem* probs, int nprobs,
                                   int M_runtime, int max_M, int max_N, int max_batch, int splitk, int max_k,
                                   cudaStream_t stream) {
  if (nprobs <= 0 || max_M <= 0 || max_N <= 0 || max_batch <= 0) return PML_OK;
  if (splitk < 1) splitk = 1;
  // With the cross terms in their own accumulator a plain reduction of 256 truncates the running sum 32 times - fewer than
  // the 48 of a 128-long reduction before that change - so up to 256 the faster plain kernels run (BN up to 128, three /
  // two CTAs per SM); beyond, the promoting variant.
  static int plain_max_k = 0;  // PML_GEMM_PLAIN_MAX_K in the environment overrides the limit (accuracy / speed A/B runs)
  if (plain_max_k == 0) {
    const char* e = getenv("PML_GEMM_PLAIN_MAX_K");
    plain_max_k = (e && atoi(e) > 0) ? atoi(e) : pml::kPlainMaxK;
  }
  const bool promote = max_k <= 0 || max_k > plain_max_k;
  int rc;
  if (promote) {
    if (max_N <= 32) rc = pml::launch_tc<32, true>(A, B, C, probs, nprobs, M_runtime, max_M, max_N, max_batch, splitk, stream);
    else rc = pml::launch_tc<64, true>(A, B, C, probs, nprobs, M_runtime, max_M, max_N, max_batch, splitk, stream);
  } else {
    if (max_N <= 32) rc = pml::launch_tc<32, false>(A, B, C, probs, nprobs, M_runtime, max_M, max_N, max_batch, splitk, stream);
    else if (max_N <= 64) rc = pml::launch_tc<64, false>(A, B, C, probs, nprobs, M_runtime, max_M, max_N, max_batch, splitk, stream);
    else rc = pml::launch_tc<128, false>(A, B, C, probs, nprobs, M_runtime, max_M, max_N, max_batch, splitk, stream);
  }
  if (rc != PML_OK) return rc;
  PML_CHECK_LAUNCH();
  return PML_OK;
}
