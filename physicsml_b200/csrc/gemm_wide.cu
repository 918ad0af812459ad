// Streaming tensor-core kernels for the one GEMM family that dominates the step: the output layer of the radial MLP
// (e3nn nn.FullyConnectedNet last layer; reference blocks.py:157-160,211 / interaction_block.py:61-64,85),
//     R[E, N] = alpha * H[E, K] @ W[K, N],      K = 64 hidden features,  N = paths x channels = 512 ... 1280,
// i.e. a tall operand with a short reduction and a wide output: 4*E*N bytes written, 4*E*K read — an HBM stream with a
// GEMM inside.  The general grouped kernel (gemm_tc.cu) re-stages both operands through registers for every 128 x 128
// output tile and runs at 1.6 TB/s on it; here
//   * W is split into its tf32 hi/lo parts ONCE per call by a tiny pack kernel that writes, per 64-column tile, the
//     exact shared-memory image the tensor core wants (canonical K-major no-swizzle UMMA layout), so a tile of B
//     arrives with a single 32 KB `cp.async.bulk` (TMA) — no register pass, no conversion in the main kernel;
//   * the CTA is persistent and A-stationary: the 128 x 64 A tile is loaded, split and stored once, then all N/64
//     column tiles of that row block are produced from it;
//   * warp-specialised: warp 5 streams B images through a 3-stage mbarrier ring, warp 4 issues tcgen05.mma
//     (kind::tf32, M128 N64 K8; 3xTF32: lo*hi + hi*lo + hi*hi) into one of two TMEM accumulators, warps 0-3 drain the
//     other one (tcgen05.ld -> per-warp shared staging -> 256-byte row segments to global), so the stores of tile j
//     overlap the MMAs of tile j+1 and the copy of tile j+2.
// fp32 accuracy: K <= 64 means at most 24 MMAs per accumulator, far below the length where the tensor core's
// truncating accumulation shows (see gemm_tc.cu, kPromoBlocks).
#include <cstdlib>

#include "pml_common.cuh"

namespace pml {
namespace wide {

constexpr int BM = 128, BN = 64, KP = 64;  // row tile, column tile, padded reduction length
constexpr int B_STAGES = 3;
constexpr int TILE_FLOATS = BN * KP;       // one hi (or lo) image of a B tile: 16 KB
constexpr int STG_PITCH = 32 + 4;  // the epilogue stages 32 columns at a time
constexpr int THREADS = 320;               // warps 0-7: A loaders + epilogue, warp 8: MMA issuer, warp 9: B producer

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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {  // see gemm_tc.cu
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
// 16-byte unit (r, kc) of a [rows x 64] K-major tile: (r % 8) + 8 * kc + 128 * (r / 8)
//   => core matrices 128 B apart along K (leading byte offset), 2048 B apart along M/N (stride byte offset)
__device__ __forceinline__ int unit_of(int r, int kc) { return (r & 7) + 8 * kc + 128 * (r >> 3); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(2048 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// One elected thread issues every MMA of the CTA, so its instruction stream is the critical path (measured: the
// per-k-step descriptor arithmetic and predicate set-up of separate asm statements cost more than the MMAs themselves
// take to execute).  A whole k-block of the 3xTF32 product goes out as one asm statement: the four descriptors advance
// by 256 bytes (16 in descriptor units) per 8-wide k-step inside it.
#define PML_MMA3(P)                                                    \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], al, bh, %5, " P ";\n\t"   \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bl, %5, pt;\n\t"      \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], ah, bh, %5, pt;\n\t"
#define PML_ADV "add.s64 ah, ah, 16;\n\tadd.s64 al, al, 16;\n\tadd.s64 bh, bh, 16;\n\tadd.s64 bl, bl, 16;\n\t"
#define PML_MMA_HEAD                                                                                               \
  "{\n\t.reg .pred p, pt;\n\t.reg .b64 ah, al, bh, bl;\n\t"                                                        \
  "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"                                                           \
  "mov.b64 ah, %1;\n\tmov.b64 al, %2;\n\tmov.b64 bh, %3;\n\tmov.b64 bl, %4;\n\t"
// D (+)= A B over 4 / 8 k-steps; `accumulate` = 0 clears D with the very first MMA
__device__ __forceinline__ void umma_3x_k32(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(PML_MMA_HEAD PML_MMA3("p") PML_ADV PML_MMA3("pt") PML_ADV PML_MMA3("pt") PML_ADV PML_MMA3("pt") "}"
               ::"r"(d), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void umma_3x_k64(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(PML_MMA_HEAD PML_MMA3("p") PML_ADV PML_MMA3("pt") PML_ADV PML_MMA3("pt") PML_ADV PML_MMA3("pt") PML_ADV
               PML_MMA3("pt") PML_ADV PML_MMA3("pt") PML_ADV PML_MMA3("pt") PML_ADV PML_MMA3("pt") "}"
               ::"r"(d), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
               : "memory");
}
#define PML_TMEM_LD32(r, taddr)                                                                                        \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                        \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                        \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),         \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),        \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                      \
      : "r"(taddr))
// 32 lanes x 64 consecutive columns: both loads in flight, one wait, then the values are read
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float* v) {
  uint32_t r0[32], r1[32];
  PML_TMEM_LD32(r0, taddr);
  PML_TMEM_LD32(r1, taddr + 32);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v[i] = __uint_as_float(r0[i]);
    v[32 + i] = __uint_as_float(r1[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// pack: W [K, N] (element (k, n) at W[k * w_rs + n * w_cs]) -> per 64-column tile t the shared-memory images
//   image[t][0] = hi, image[t][1] = lo, each [64 n][64 k] in the unit_of layout, k >= K zero-filled.
// TRANSPOSED = true packs W^T instead: tile t covers 64 consecutive k... (used by the dx kernel; see below)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void pack_w_kernel(const float* __restrict__ W, long long w_rs, long long w_cs, int K, int N,
                              float* __restrict__ image) {
  pdl_enter();
  // one thread per 16-byte unit of one tile: 64 rows x 16 k-units
  const int t = blockIdx.x;
  for (int idx = threadIdx.x; idx < BN * (KP / 4); idx += blockDim.x) {
    const int r = idx & 63, kc = idx >> 6;
    const int n = t * BN + r;
    float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), lo = hi;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * 4 + j;
      v[j] = (k < K && n < N) ? __ldg(W + (long long)k * w_rs + (long long)n * w_cs) : 0.f;
    }
    split_tf32(v[0], hi.x, lo.x);
    split_tf32(v[1], hi.y, lo.y);
    split_tf32(v[2], hi.z, lo.z);
    split_tf32(v[3], hi.w, lo.w);
    float* base = image + (size_t)t * 2 * TILE_FLOATS;
    reinterpret_cast<float4*>(base)[unit_of(r, kc)] = hi;
    reinterpret_cast<float4*>(base + TILE_FLOATS)[unit_of(r, kc)] = lo;
  }
}

struct __align__(128) FwdSmem {
  float a[2][BM * KP];                 // hi, lo                                   64 KB
  float b[B_STAGES][2 * TILE_FLOATS];  // hi, lo per stage                         96 KB
  float stg[8][32 * STG_PITCH];        // per epilogue warp                        36 KB
  unsigned long long b_full[B_STAGES], b_empty[B_STAGES], acc_full[2], acc_empty[2], a_full, a_empty;
  uint32_t tmem_base;
};

// C[M, N] (row stride ldc) = alpha * A[M, K] (row stride lda, K <= 64, K % 4 == 0) @ W ; W given as its packed image.
// Work items = (row tile, column range): item i -> row tile i / nsplit, column tiles [(i % nsplit) * per, ... + per).
__global__ void __launch_bounds__(THREADS, 1) wide_fwd_kernel(const float* __restrict__ A, long long lda, int K,
                                                              const float* __restrict__ image, float* __restrict__ C,
                                                              long long ldc, int M, int ntiles_n, int nsplit,
                                                              float alpha) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_m = (M + BM - 1) / BM;
  const int nitems = tiles_m * nsplit;
  const int per = (ntiles_n + nsplit - 1) / nsplit;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(smem_u32(&S.b_full[s]), 1);
      mbar_init(smem_u32(&S.b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&S.acc_full[s]), 1);
      mbar_init(smem_u32(&S.acc_empty[s]), 4);
    }
    mbar_init(smem_u32(&S.a_full), 8);
    mbar_init(smem_u32(&S.a_empty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 9) {
    // ---------------- B producer
    if (lane == 0) {
      int t = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int ns = item % nsplit;
        const int j0 = ns * per, j1 = min(ntiles_n, j0 + per);
        for (int j = j0; j < j1; ++j, ++t) {
          const int s = t % B_STAGES;
          mbar_wait(smem_u32(&S.b_empty[s]), ((t / B_STAGES) & 1) ^ 1);
          const uint32_t fb = smem_u32(&S.b_full[s]);
          mbar_expect_tx(fb, 2u * TILE_FLOATS * 4u);
          bulk_g2s(smem_u32(S.b[s]), image + (size_t)j * 2 * TILE_FLOATS, 2u * TILE_FLOATS * 4u, fb);
        }
      }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint64_t da_hi = make_desc(smem_u32(S.a[0])), da_lo = make_desc(smem_u32(S.a[1]));
      int t = 0, it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int ns = item % nsplit;
        const int j0 = ns * per, j1 = min(ntiles_n, j0 + per);
        mbar_wait(smem_u32(&S.a_full), it & 1);
        for (int j = j0; j < j1; ++j, ++t) {
          const int s = t % B_STAGES, buf = t & 1;
          mbar_wait(smem_u32(&S.b_full[s]), (t / B_STAGES) & 1);
          mbar_wait(smem_u32(&S.acc_empty[buf]), ((t >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_hi = smem_u32(S.b[s]), b_lo = b_hi + TILE_FLOATS * 4;
          const uint32_t d = tmem + (uint32_t)(buf * BN);
          umma_3x_k64(d, da_hi, da_lo, make_desc(b_hi), make_desc(b_lo), IDESC, 0u);
          umma_commit(smem_u32(&S.b_empty[s]));
          umma_commit(smem_u32(&S.acc_full[buf]));
        }
        umma_commit(smem_u32(&S.a_empty));  // every MMA that reads this item's A tile has retired
      }
    }
  } else {
    // ---------------- warps 0-7: A tile loaders, then two epilogue sets: warps 0-3 drain TMEM buffer 0 (even tiles),
    // warps 4-7 buffer 1 (odd tiles); warp w reads TMEM lane quadrant w % 4.  One set's tcgen05.ld -> staging -> store
    // chain (0.7 us per tile, latency-bound) overlaps the other set's.
    const int set = warp >> 2, quad = warp & 3;
    float* stg = S.stg[warp];
    constexpr int A_PER_THREAD = (BM * (KP / 4)) / 256;  // 8 float4 per thread
    float4 areg[A_PER_THREAD];
    auto load_a = [&](int item) {  // next item's A tile -> registers (consumed one item later: latency fully hidden)
      const int row0 = (item / nsplit) * BM;
#pragma unroll
      for (int i = 0; i < A_PER_THREAD; ++i) {
        const int idx = tid + 256 * i;
        const int r = (idx & 7) + 8 * (idx >> 7), kc = (idx >> 3) & 15;
        areg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < nitems && row0 + r < M && kc * 4 < K)
          areg[i] = __ldg(reinterpret_cast<const float4*>(A + (long long)(row0 + r) * lda + kc * 4));
      }
    };
    load_a(blockIdx.x);
    int t = 0, it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int mt = item / nsplit, ns = item % nsplit;
      const int j0 = ns * per, j1 = min(ntiles_n, j0 + per);
      const int row0 = mt * BM;
      mbar_wait(smem_u32(&S.a_empty), (it & 1) ^ 1);
      // 128 rows x 16 k-units over 256 threads; a warp instruction covers 8 rows x 64 contiguous bytes in global memory
      // and 512 contiguous bytes in shared memory (conflict-free)
#pragma unroll
      for (int i = 0; i < A_PER_THREAD; ++i) {
        const int idx = tid + 256 * i;
        const int r = (idx & 7) + 8 * (idx >> 7), kc = (idx >> 3) & 15;
        float4 hi, lo;
        split_tf32(areg[i].x, hi.x, lo.x);
        split_tf32(areg[i].y, hi.y, lo.y);
        split_tf32(areg[i].z, hi.z, lo.z);
        split_tf32(areg[i].w, hi.w, lo.w);
        const int u = unit_of(r, kc);
        reinterpret_cast<float4*>(S.a[0])[u] = hi;
        reinterpret_cast<float4*>(S.a[1])[u] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.a_full));
      load_a(item + gridDim.x);

      for (int j = j0; j < j1; ++j, ++t) {
        if ((t & 1) != set) continue;
        mbar_wait(smem_u32(&S.acc_full[set]), (t >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[BN];
        tmem_ld64(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(set * BN), v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&S.acc_empty[set]));
        // 32 columns at a time: own row -> staging (pitch 36 floats: conflict-free 128-bit stores), then 128-byte row
        // segments out, 4 rows per warp instruction
        const int sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float* srow = stg + lane * STG_PITCH;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(srow + c) = make_float4(alpha * v[32 * p + c], alpha * v[32 * p + c + 1],
                                                                alpha * v[32 * p + c + 2], alpha * v[32 * p + c + 3]);
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) {
            const int r = rr + sub;
            const int gr = row0 + quad * 32 + r;
            const float4 o = *reinterpret_cast<const float4*>(stg + r * STG_PITCH + c4);
            if (gr < M) *reinterpret_cast<float4*>(C + (long long)gr * ldc + (long long)j * BN + 32 * p + c4) = o;
          }
          __syncwarp();
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN) : "memory");
  }
}


// =====================================================================================================================
// dx of the same layer:  dA[M, K] = alpha * G[M, N] @ W^T   (K <= 64 outputs, reduction over the wide dimension N).
// The big operand G is now the streamed one: 8 loader warps pull 128 x 32 chunks through a register ring four chunks
// deep (the loads of chunk c+4 are issued when chunk c is converted), split them into tf32 hi/lo and store them in the
// UMMA layout of a 3-stage ring; W^T arrives pre-split, 16 KB per chunk, by TMA; warp 12 issues the MMAs; warps 0-3
// own the accumulators: every 128 k the tensor core switches TMEM buffer and they add the retired one into fp32
// registers (the tensor core truncates its running sum: see gemm_tc.cu), then write the 128 x K block.  The loaders run
// across row-block boundaries, so the stream never drains between items.
// =====================================================================================================================
constexpr int XC = 32;                      // reduction chunk
constexpr int X_A_STAGES = 3, X_B_STAGES = 4, X_PERIOD = 4;  // promotion every X_PERIOD chunks = 128 k
constexpr int X_A_FLOATS = BM * XC, X_B_FLOATS = KP * XC;    // one hi (or lo) image: 16 KB / 8 KB
constexpr int X_THREADS = 448;              // warps 0-3 accumulate, 4-11 load G, 12 MMA, 13 B producer
constexpr int X_DEPTH = 4;                  // chunks of register prefetch

__device__ __forceinline__ int unit32_of(int r, int kc) { return (r & 7) + 8 * kc + 64 * (r >> 3); }
__device__ __forceinline__ uint64_t make_desc32(uint32_t saddr) {  // [rows x 32] K-major tile: LBO 128 B, SBO 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// image[c][0 = hi, 1 = lo][unit32_of(k_out, kc)] = W[k_out, 32 c + 4 kc .. + 3] split; k_out >= K zero-filled
__global__ void pack_wt_kernel(const float* __restrict__ W, long long w_rs, long long w_cs, int K, int N,
                               float* __restrict__ image) {
  pdl_enter();
  const int c = blockIdx.x;
  for (int idx = threadIdx.x; idx < KP * (XC / 4); idx += blockDim.x) {
    const int kc = idx & 7, r = idx >> 3;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = c * XC + kc * 4 + j;
      v[j] = (r < K && n < N) ? __ldg(W + (long long)r * w_rs + (long long)n * w_cs) : 0.f;
    }
    float4 hi, lo;
    split_tf32(v[0], hi.x, lo.x);
    split_tf32(v[1], hi.y, lo.y);
    split_tf32(v[2], hi.z, lo.z);
    split_tf32(v[3], hi.w, lo.w);
    float* base = image + (size_t)c * 2 * X_B_FLOATS;
    reinterpret_cast<float4*>(base)[unit32_of(r, kc)] = hi;
    reinterpret_cast<float4*>(base + X_B_FLOATS)[unit32_of(r, kc)] = lo;
  }
}

struct __align__(128) BwdXSmem {
  float a[X_A_STAGES][2 * X_A_FLOATS];  // hi, lo per stage     96 KB
  float b[X_B_STAGES][2 * X_B_FLOATS];  // hi, lo per stage     64 KB
  unsigned long long a_full[X_A_STAGES], a_empty[X_A_STAGES], b_full[X_B_STAGES], b_empty[X_B_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(X_THREADS, 1) wide_bwd_x_kernel(const float* __restrict__ G, long long ldg, int N,
                                                                  const float* __restrict__ image, float* __restrict__ dA,
                                                                  long long lda, int K, int M, float alpha) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdXSmem& S = *reinterpret_cast<BwdXSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nitems = (M + BM - 1) / BM;
  const int nch = N / XC;
  const int my_items = ((int)blockIdx.x < nitems) ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_items * nch;  // chunks this CTA streams

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < X_A_STAGES; ++s) {
      mbar_init(smem_u32(&S.a_full[s]), 8);
      mbar_init(smem_u32(&S.a_empty[s]), 1);
    }
    for (int s = 0; s < X_B_STAGES; ++s) {
      mbar_init(smem_u32(&S.b_full[s]), 1);
      mbar_init(smem_u32(&S.b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&S.acc_full[s]), 1);
      mbar_init(smem_u32(&S.acc_empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 13) {
    // ---------------- B producer: chunk images repeat for every item
    if (lane == 0) {
      for (int g = 0; g < total; ++g) {
        const int c = g % nch, s = g % X_B_STAGES;
        mbar_wait(smem_u32(&S.b_empty[s]), ((g / X_B_STAGES) & 1) ^ 1);
        const uint32_t fb = smem_u32(&S.b_full[s]);
        mbar_expect_tx(fb, 2u * X_B_FLOATS * 4u);
        bulk_g2s(smem_u32(S.b[s]), image + (size_t)c * 2 * X_B_FLOATS, 2u * X_B_FLOATS * 4u, fb);
      }
    }
  } else if (warp == 12) {
    // ---------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int P = 0;  // accumulation periods issued so far
      for (int g = 0; g < total; ++g) {
        const int c = g % nch, sa = g % X_A_STAGES, sb = g % X_B_STAGES;
        const bool opens = (c % X_PERIOD) == 0, closes = (c % X_PERIOD) == X_PERIOD - 1 || c == nch - 1;
        const int buf = P & 1;
        if (opens) mbar_wait(smem_u32(&S.acc_empty[buf]), ((P >> 1) & 1) ^ 1);
        mbar_wait(smem_u32(&S.a_full[sa]), (g / X_A_STAGES) & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // loaders' generic-proxy stores -> tensor core
        mbar_wait(smem_u32(&S.b_full[sb]), (g / X_B_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(S.a[sa]), a_lo = a_hi + X_A_FLOATS * 4;
        const uint32_t b_hi = smem_u32(S.b[sb]), b_lo = b_hi + X_B_FLOATS * 4;
        const uint32_t d = tmem + (uint32_t)(buf * BN);
        umma_3x_k32(d, make_desc32(a_hi), make_desc32(a_lo), make_desc32(b_hi), make_desc32(b_lo), IDESC, opens ? 0u : 1u);
        umma_commit(smem_u32(&S.a_empty[sa]));
        umma_commit(smem_u32(&S.b_empty[sb]));
        if (closes) {
          umma_commit(smem_u32(&S.acc_full[buf]));
          ++P;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- G loaders: 256 threads, 4 float4 per thread and chunk
    const int lt = tid - 128;
    float4 q[X_DEPTH][4];
    auto load = [&](float4 (&dst)[4], int g) {
      if (g >= total) return;
      const int ik = g / nch, c = g - ik * nch;
      const int row0 = ((int)blockIdx.x + ik * (int)gridDim.x) * BM;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = lt + 256 * i;
        const int r = (idx & 7) + 8 * (idx >> 6), kc = (idx >> 3) & 7;
        dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) dst[i] = __ldcs(reinterpret_cast<const float4*>(G + (long long)(row0 + r) * ldg + c * XC + kc * 4));
      }
    };
    auto store = [&](const float4 (&src)[4], int g) {
      const int s = g % X_A_STAGES;
      mbar_wait(smem_u32(&S.a_empty[s]), ((g / X_A_STAGES) & 1) ^ 1);
      float* hi_base = S.a[s];
      float* lo_base = hi_base + X_A_FLOATS;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = lt + 256 * i;
        const int r = (idx & 7) + 8 * (idx >> 6), kc = (idx >> 3) & 7;
        float4 hi, lo;
        split_tf32(src[i].x, hi.x, lo.x);
        split_tf32(src[i].y, hi.y, lo.y);
        split_tf32(src[i].z, hi.z, lo.z);
        split_tf32(src[i].w, hi.w, lo.w);
        const int u = unit32_of(r, kc);
        reinterpret_cast<float4*>(hi_base)[u] = hi;
        reinterpret_cast<float4*>(lo_base)[u] = lo;
      }
      // no proxy fence here: it would wait for the prefetch loads this thread has in flight (fence.proxy.async is a
      // MEMBAR + FENCE.VIEW.ASYNC).  The MMA thread, which has none, fences after it has acquired a_full.
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.a_full[s]));
    };
#pragma unroll
    for (int u = 0; u < X_DEPTH; ++u) load(q[u], u);
    for (int g = 0; g < total; g += X_DEPTH) {
#pragma unroll
      for (int u = 0; u < X_DEPTH; ++u) {
        if (g + u < total) {
          store(q[u], g + u);
          load(q[u], g + u + X_DEPTH);
        }
      }
    }
  } else {
    // ---------------- warps 0-3: accumulator promotion (TMEM lane quadrant `warp`) and the output block
    int P = 0;
    for (int ik = 0; ik < my_items; ++ik) {
      const int row0 = ((int)blockIdx.x + ik * (int)gridDim.x) * BM;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      const int nper = (nch + X_PERIOD - 1) / X_PERIOD;
      for (int p = 0; p < nper; ++p, ++P) {
        const int buf = P & 1;
        mbar_wait(smem_u32(&S.acc_full[buf]), (P >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[BN];
        tmem_ld64(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN), v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&S.acc_empty[buf]));
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] += v[j];
      }
      const int gr = row0 + warp * 32 + lane;
      if (gr < M) {
        float* orow = dA + (long long)gr * lda;
#pragma unroll
        for (int j = 0; j < BN; j += 4)
          if (j < K) *reinterpret_cast<float4*>(orow + j) = make_float4(alpha * acc[j], alpha * acc[j + 1], alpha * acc[j + 2], alpha * acc[j + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * BN) : "memory");
  }
}


// =====================================================================================================================
// dx, second version (the one that runs): the streamed operand lives in TENSOR MEMORY.  A chunk of G goes: 16-byte
// cp.async -> raw [128 x 32] tile in an 8-stage shared-memory ring -> the thread that owns a row reads it back, splits
// it into tf32 hi/lo and writes both halves into TMEM (tcgen05.st 32x32b, lane = row, column = k); the MMAs take A
// from TMEM ([a_tmem] operand form) and only the W^T images come from shared memory.  Measured: moving the operand to
// TMEM did nothing by itself (the first version is not shared-memory-bandwidth-bound, as its 66 % l1tex utilisation had
// suggested) - what it frees is shared memory for a prefetch ring deep enough for DRAM latency (~100 KB in flight per
// SM, no registers held): 827 -> 540 us at 332 k rows.
// Two loader sets of four warps (one per TMEM lane quadrant) alternate chunks; 4 TMEM stages of 64 columns (hi | lo).
// =====================================================================================================================
constexpr int XT_A_STAGES = 4, XT_RAW_PITCH = 36, XT_RAW_STAGES = 4, XT_AHEAD = 3;

#define PML_MMA3_TS(P)                                                     \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [al], bh, %5, " P ";\n\t"     \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bl, %5, pt;\n\t"        \
  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah], bh, %5, pt;\n\t"
#define PML_ADV_TS "add.u32 ah, ah, 8;\n\tadd.u32 al, al, 8;\n\tadd.s64 bh, bh, 16;\n\tadd.s64 bl, bl, 16;\n\t"
__device__ __forceinline__ void umma_3x_k32_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, pt;\n\t.reg .b64 bh, bl;\n\t.reg .b32 ah, al;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 pt, %6, %6;\n\t"
      "mov.b32 ah, %1;\n\tmov.b32 al, %2;\n\tmov.b64 bh, %3;\n\tmov.b64 bl, %4;\n\t"
      PML_MMA3_TS("p") PML_ADV_TS PML_MMA3_TS("pt") PML_ADV_TS PML_MMA3_TS("pt") PML_ADV_TS PML_MMA3_TS("pt") "}"
      ::"r"(d), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
      : "memory");
}

struct __align__(128) BwdXTSmem {
  float raw[2][XT_RAW_STAGES][BM * XT_RAW_PITCH];  // [loader set][stage]: raw fp32 chunks, row pitch 36 floats  144 KB
  unsigned long long raw_full[2][XT_RAW_STAGES];
  float b[X_B_STAGES][2 * X_B_FLOATS];  // W^T chunk images (hi, lo)                                       64 KB
  unsigned long long a_full[XT_A_STAGES], a_empty[XT_A_STAGES], b_full[X_B_STAGES], b_empty[X_B_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(X_THREADS, 1) wide_bwd_x_tmem_kernel(const float* __restrict__ G, long long ldg, int N,
                                                                       const float* __restrict__ image,
                                                                       float* __restrict__ dA, long long lda, int K, int M,
                                                                       float alpha) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdXTSmem& S = *reinterpret_cast<BwdXTSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nitems = (M + BM - 1) / BM;
  const int nch = N / XC;
  const int my_items = ((int)blockIdx.x < nitems) ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = my_items * nch;
  constexpr uint32_t TCOLS = 512;            // 2 accumulators x 64 + 4 A stages x 64 = 384 -> next power of two
  constexpr uint32_t A_COL0 = 2 * BN;        // first A-stage column

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < XT_A_STAGES; ++s) {
      mbar_init(smem_u32(&S.a_full[s]), 4);
      mbar_init(smem_u32(&S.a_empty[s]), 1);
    }
    for (int s = 0; s < 2 * XT_RAW_STAGES; ++s) mbar_init(smem_u32(&S.raw_full[0][0] + s), 128);
    for (int s = 0; s < X_B_STAGES; ++s) {
      mbar_init(smem_u32(&S.b_full[s]), 1);
      mbar_init(smem_u32(&S.b_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&S.acc_full[s]), 1);
      mbar_init(smem_u32(&S.acc_empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 13) {
    if (lane == 0) {
      for (int g = 0; g < total; ++g) {
        const int c = g % nch, s = g % X_B_STAGES;
        mbar_wait(smem_u32(&S.b_empty[s]), ((g / X_B_STAGES) & 1) ^ 1);
        const uint32_t fb = smem_u32(&S.b_full[s]);
        mbar_expect_tx(fb, 2u * X_B_FLOATS * 4u);
        bulk_g2s(smem_u32(S.b[s]), image + (size_t)c * 2 * X_B_FLOATS, 2u * X_B_FLOATS * 4u, fb);
      }
    }
  } else if (warp == 12) {
    if (lane == 0) {
      constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int P = 0;
      for (int g = 0; g < total; ++g) {
        const int c = g % nch, sa = g % XT_A_STAGES, sb = g % X_B_STAGES;
        const bool opens = (c % X_PERIOD) == 0, closes = (c % X_PERIOD) == X_PERIOD - 1 || c == nch - 1;
        const int buf = P & 1;
        if (opens) mbar_wait(smem_u32(&S.acc_empty[buf]), ((P >> 1) & 1) ^ 1);
        mbar_wait(smem_u32(&S.a_full[sa]), (g / XT_A_STAGES) & 1);
        mbar_wait(smem_u32(&S.b_full[sb]), (g / X_B_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = tmem + A_COL0 + (uint32_t)(sa * 2 * XC), a_lo = a_hi + XC;
        const uint32_t b_hi = smem_u32(S.b[sb]), b_lo = b_hi + X_B_FLOATS * 4;
        umma_3x_k32_ts(tmem + (uint32_t)(buf * BN), a_hi, a_lo, make_desc32(b_hi), make_desc32(b_lo), IDESC, opens ? 0u : 1u);
        umma_commit(smem_u32(&S.a_empty[sa]));
        umma_commit(smem_u32(&S.b_empty[sb]));
        if (closes) {
          umma_commit(smem_u32(&S.acc_full[buf]));
          ++P;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- loader sets: set = chunk parity, quad = TMEM lane quadrant of this warp
    const int set = (warp - 4) >> 2, quad = (warp - 4) & 3;
    const int lt = tid - 128 - set * 128;  // 0..127 inside the set
    // The set's chunks (global chunk g = set + 2 n) go global -> shared by 16-byte cp.async, XT_AHEAD chunks ahead of
    // their conversion: ~100 KB in flight per SM without holding a single register (a register prefetch ring deep
    // enough for DRAM latency at this rate would need 128 registers per thread).
    auto issue = [&](int n) {
      const int g = set + 2 * n;
      if (g >= total) return;
      const int ik = g / nch, c = g - ik * nch;
      const int row0 = ((int)blockIdx.x + ik * (int)gridDim.x) * BM;
      float* rawb = S.raw[set][n % XT_RAW_STAGES];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = lt + 128 * i;
        const int r = (idx & 7) + 8 * (idx >> 6), kc = (idx >> 3) & 7;
        const bool ok = row0 + r < M;
        const float* src = G + (long long)(ok ? row0 + r : row0) * ldg + c * XC + kc * 4;
        const uint32_t dst = smem_u32(rawb + r * XT_RAW_PITCH + kc * 4);
        const uint32_t bytes = ok ? 16u : 0u;  // rows past M are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&S.raw_full[set][n % XT_RAW_STAGES]))
                   : "memory");
    };
    const int my_chunks = (total - set + 1) / 2;  // chunks g = set, set + 2, ... < total
#pragma unroll
    for (int n = 0; n < XT_AHEAD; ++n) issue(n);
    for (int n = 0; n < my_chunks; ++n) {
      const int g = set + 2 * n;
      const float* rawb = S.raw[set][n % XT_RAW_STAGES];
      mbar_wait(smem_u32(&S.raw_full[set][n % XT_RAW_STAGES]), (n / XT_RAW_STAGES) & 1);
      float hi[XC], lo[XC];
      const float* myrow = rawb + (quad * 32 + lane) * XT_RAW_PITCH;
#pragma unroll
      for (int j = 0; j < XC / 4; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(myrow + 4 * j);
        split_tf32(t.x, hi[4 * j], lo[4 * j]);
        split_tf32(t.y, hi[4 * j + 1], lo[4 * j + 1]);
        split_tf32(t.z, hi[4 * j + 2], lo[4 * j + 2]);
        split_tf32(t.w, hi[4 * j + 3], lo[4 * j + 3]);
      }
      // every thread of the set has read chunk n (hence n - 1): the stage of chunk n - 1 may be refilled
      asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
      issue(n + XT_AHEAD);
      const int s = g % XT_A_STAGES;
      mbar_wait(smem_u32(&S.a_empty[s]), ((g / XT_A_STAGES) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + A_COL0 + (uint32_t)(s * 2 * XC);
      tmem_st32(ta, hi);
      tmem_st32(ta + XC, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&S.a_full[s]));
    }
  } else {
    // ---------------- warps 0-3: accumulator promotion and the output block (as in wide_bwd_x_kernel)
    int P = 0;
    for (int ik = 0; ik < my_items; ++ik) {
      const int row0 = ((int)blockIdx.x + ik * (int)gridDim.x) * BM;
      float acc[BN];
#pragma unroll
      for (int j = 0; j < BN; ++j) acc[j] = 0.f;
      const int nper = (nch + X_PERIOD - 1) / X_PERIOD;
      for (int p = 0; p < nper; ++p, ++P) {
        const int buf = P & 1;
        mbar_wait(smem_u32(&S.acc_full[buf]), (P >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float v[BN];
        tmem_ld64(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN), v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&S.acc_empty[buf]));
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] += v[j];
      }
      const int gr = row0 + warp * 32 + lane;
      if (gr < M) {
        float* orow = dA + (long long)gr * lda;
#pragma unroll
        for (int j = 0; j < BN; j += 4)
          if (j < K) *reinterpret_cast<float4*>(orow + j) = make_float4(alpha * acc[j], alpha * acc[j + 1], alpha * acc[j + 2], alpha * acc[j + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
  }
}

}  // namespace wide
}  // namespace pml

// floats the packed image of a [K <= 64, N] weight needs (N a multiple of 64)
extern "C" long long pml_wide_image_floats(int N) { return (long long)((N + 63) / 64) * 2 * pml::wide::TILE_FLOATS; }

// image <- packed tf32 hi/lo tiles of W (element (k, n) at W[k * w_rs + n * w_cs]); K <= 64
extern "C" int pml_wide_pack_w(const float* W, long long w_rs, long long w_cs, int K, int N, float* image,
                               cudaStream_t stream) {
  if (K <= 0 || K > pml::wide::KP || N <= 0) return PML_ERR_BAD_ARG;
  pml::launch_plain(pml::wide::pack_w_kernel, dim3((N + 63) / 64), dim3(256), 0, stream, W, w_rs, w_cs, K, N, image);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// C[M, N] = alpha * A[M, K] @ W   (A rows lda apart, 16-byte aligned, K % 4 == 0, K <= 64; N % 64 == 0; C rows ldc apart)
extern "C" int pml_wide_fwd(const float* A, long long lda, int K, const float* image, float* C, long long ldc, int M, int N,
                            float alpha, cudaStream_t stream) {
  using namespace pml::wide;
  if (M <= 0) return PML_OK;
  if (K <= 0 || K > KP || (K & 3) || (N % BN) || (lda & 3) || (ldc & 3) || (reinterpret_cast<uintptr_t>(A) & 15) ||
      (reinterpret_cast<uintptr_t>(C) & 15) || (reinterpret_cast<uintptr_t>(image) & 15))
    return PML_ERR_UNSUPPORTED;
  static int sms_dev[pml::kMaxDevices] = {};
  const int slot = pml::device_slot();
  const int smem = (int)sizeof(FwdSmem) + 128;
  if (sms_dev[slot] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot);
    if (cudaFuncSetAttribute(wide_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PML_ERR_LAUNCH;
    sms_dev[slot] = n;
  }
  const int sms = sms_dev[slot];
  const int tiles_m = (M + BM - 1) / BM, ntiles_n = N / BN;
  // split the columns of a row block between CTAs when that evens out the rounds (325 row blocks on 148 SMs: 3 rounds
  // of whole blocks = 73 % busy; halves: 5 rounds of half blocks = 88 %)
  int nsplit = 1;
  double best = 1e30;
  for (int s = 1; s <= 4 && s <= ntiles_n; ++s) {
    if (ntiles_n % s) continue;
    const int rounds = (tiles_m * s + sms - 1) / sms;
    const double cost = rounds * ((double)ntiles_n / s + 1.5);  // +1.5 tile-times for loading A
    if (cost < best * 0.97) best = cost, nsplit = s;
  }
  const int grid = min(sms, tiles_m * nsplit);
  pml::launch_plain(wide_fwd_kernel, dim3(grid), dim3(THREADS), smem, stream, A, lda, K, image, C, ldc, M, ntiles_n, nsplit, alpha);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// image <- packed tf32 hi/lo tiles of W^T for the dx kernel: one [64 k_out x 32 n] tile per 32-wide chunk of N
// (2 * 64 * 32 floats per chunk: pml_wide_image_floats(N) floats in total, the same size as the forward image)
extern "C" int pml_wide_pack_wt(const float* W, long long w_rs, long long w_cs, int K, int N, float* image,
                                cudaStream_t stream) {
  if (K <= 0 || K > pml::wide::KP || N <= 0 || (N % pml::wide::XC)) return PML_ERR_BAD_ARG;
  pml::launch_plain(pml::wide::pack_wt_kernel, dim3(N / pml::wide::XC), dim3(256), 0, stream, W, w_rs, w_cs, K, N, image);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// dA[M, K] = alpha * G[M, N] @ W^T  (G rows ldg apart, dA rows lda apart, 16-byte aligned; K % 4 == 0, K <= 64,
// N % 32 == 0); image from pml_wide_pack_wt
extern "C" int pml_wide_bwd_x(const float* G, long long ldg, int N, const float* image, float* dA, long long lda, int K,
                              int M, float alpha, cudaStream_t stream) {
  using namespace pml::wide;
  if (M <= 0) return PML_OK;
  if (K <= 0 || K > KP || (K & 3) || (N % XC) || (ldg & 3) || (lda & 3) || (reinterpret_cast<uintptr_t>(G) & 15) ||
      (reinterpret_cast<uintptr_t>(dA) & 15) || (reinterpret_cast<uintptr_t>(image) & 15))
    return PML_ERR_UNSUPPORTED;
  static int sms_dev[pml::kMaxDevices] = {};
  const int slot = pml::device_slot();
  const int smem = (int)sizeof(BwdXSmem) + 128;
  if (sms_dev[slot] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot);
    if (cudaFuncSetAttribute(wide_bwd_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PML_ERR_LAUNCH;
    sms_dev[slot] = n;
  }
  const int sms = sms_dev[slot];
  const int nitems = (M + BM - 1) / BM;
  static int use_tmem = -1;  // PML_WIDE_X_TMEM=0 selects the first version (operands in shared memory) for A/B runs
  if (use_tmem < 0) use_tmem = getenv("PML_WIDE_X_TMEM") ? atoi(getenv("PML_WIDE_X_TMEM")) : 1;
  if (use_tmem) {
    static bool configured_t[pml::kMaxDevices] = {};
    const int smem_t = (int)sizeof(BwdXTSmem) + 128;
    if (!configured_t[slot]) {
      if (cudaFuncSetAttribute(wide_bwd_x_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t) != cudaSuccess)
        return PML_ERR_LAUNCH;
      configured_t[slot] = true;
    }
    pml::launch_plain(wide_bwd_x_tmem_kernel, dim3(min(sms, nitems)), dim3(X_THREADS), smem_t, stream, G, ldg, N, image, dA, lda, K, M, alpha);
    PML_CHECK_LAUNCH();
    return PML_OK;
  }
  pml::launch_plain(wide_bwd_x_kernel, dim3(min(sms, nitems)), dim3(X_THREADS), smem, stream, G, ldg, N, image, dA, lda, K, M, alpha);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
