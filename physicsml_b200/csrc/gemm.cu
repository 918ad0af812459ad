// Grouped, strided, multi-chunk fp32 GEMM used for every dense contraction on the path that is not the radial-MLP
// tensor-core kernel: per-irrep linear mixing (o3.Linear; reference blocks.py:139-144,176-181,65-70 and
// interaction_block.py:22-27,66-71), the readout layers, and all weight-gradient reductions.
//
// One launch executes a TABLE of problems.  Problem p computes, for every batch index z < nbatch,
//     C_z[r, c] (+)= alpha * sum_chunks sum_k A_z,chunk[r, k] * B_chunk[k, c]
// with fully general element strides, which is what the e3nn "mul-major" irreps layout needs:
//     y[n, off_o + w*d + m] = sum_u W[u, w] x[n, off_i + u*d + m]
// is the problem (rows = nodes, cols = w, k = u) with A row stride = row width, A col stride = d, batch stride = 1
// (the batch index is the m of the irrep), and several input blocks of the same irrep are K-chunks of one problem.
//
// fp32 FFMA, 64x64x16 tiles, 4x4 micro-tiles.  Split-K (atomic epilogue) is used for the tall-skinny weight-gradient
// shapes.  TODO(round 2): tcgen05 3xTF32 path for the large-M problems.
#include "pml_common.cuh"

#include "gemm_problem.h"

namespace pml {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

__global__ void __launch_bounds__(256) gemm_grouped_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ C, const pml_gemm_problem* __restrict__ probs,
                                                           int M_runtime, int splitk, int tiles_m_max) {
  pdl_enter();
  const pml_gemm_problem& P = probs[blockIdx.y];
  const int M = P.M < 0 ? M_runtime : P.M;
  const int N = P.N;
  const int tiles_n = (N + BN - 1) / BN;
  const int tile = blockIdx.x;
  const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
  if (tm * BM >= M) return;
  int zb = blockIdx.z;
  const int ks = zb % splitk;
  zb /= splitk;
  if (zb >= P.nbatch) return;

  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;  // 16 x 16 threads, each TM x TN outputs
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int row0 = tm * BM, col0 = tn * BN;
  // load mappings: pick the thread->element map that makes the unit-stride dimension contiguous across threads
  const bool a_kfast = (P.a_cs == 1);
  const bool b_nfast = (P.b_cs == 1);

  for (int ch = 0; ch < P.nchunks; ++ch) {
    const float* Ab = A + P.a_off[ch] + (long long)zb * P.a_bs;
    const float* Bb = B + P.b_off[ch] + (long long)zb * P.b_bs;
    const int K = P.kc[ch] < 0 ? M_runtime : P.kc[ch];
    const float cs = P.cscale[ch];
    // split-K: this CTA handles k-tiles kt with (kt % splitk) == ks
    for (int k0 = ks * BK; k0 < K; k0 += BK * splitk) {
#pragma unroll
      for (int it = 0; it < (BM * BK) / 256; ++it) {
        const int idx = tid + it * 256;
        int r, k;
        if (a_kfast) {
          k = idx % BK;
          r = idx / BK;
        } else {
          r = idx % BM;
          k = idx / BM;
        }
        const int gr = row0 + r, gk = k0 + k;
        float v = 0.f;
        if (gr < M && gk < K) v = cs * __ldg(Ab + (long long)gr * P.a_rs + (long long)gk * P.a_cs);
        As[k][r] = v;
      }
#pragma unroll
      for (int it = 0; it < (BN * BK) / 256; ++it) {
        const int idx = tid + it * 256;
        int c, k;
        if (b_nfast) {
          c = idx % BN;
          k = idx / BN;
        } else {
          k = idx % BK;
          c = idx / BK;
        }
        const int gc = col0 + c, gk = k0 + k;
        float v = 0.f;
        if (gc < N && gk < K) v = __ldg(Bb + (long long)gk * P.b_rs + (long long)gc * P.b_cs);
        Bs[k][c] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  float* Cb = C + P.c_off + (long long)zb * P.c_bs;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gr = row0 + ty * TM + i;
    if (gr >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gc = col0 + tx * TN + j;
      if (gc >= N) continue;
      float* p = Cb + (long long)gr * P.c_rs + (long long)gc * P.c_cs;
      const float v = P.alpha * acc[i][j];
      if (splitk > 1 || (P.c_bs == 0 && P.nbatch > 1))
        atomicAdd(p, v);
      else if (P.accumulate)
        *p += v;
      else
        *p = v;
    }
  }
}

}  // namespace pml

// `probs` is a DEVICE pointer to `nprobs` problems; max_M / max_N / max_batch bound the grid (host-known).
// With splitk > 1 (or c_bs == 0 with nbatch > 1: all batches reduce into one C) the epilogue is atomic and every C
// must be zero-initialised (or hold the value to accumulate onto) by the caller.
extern "C" int pml_gemm_grouped(const float* A, const float* B, float* C, const pml_gemm_problem* probs, int nprobs,
                                int M_runtime, int max_M, int max_N, int max_batch, int splitk, cudaStream_t stream) {
  if (nprobs <= 0 || max_M <= 0 || max_N <= 0 || max_batch <= 0) return PML_OK;
  if (splitk < 1) splitk = 1;
  const int tiles_m = (max_M + pml::BM - 1) / pml::BM, tiles_n = (max_N + pml::BN - 1) / pml::BN;
  const long long gz = (long long)max_batch * splitk;
  if (gz > 65535 || nprobs > 65535) return PML_ERR_BAD_ARG;
  dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)nprobs, (unsigned)gz);
  pml::launch_plain(pml::gemm_grouped_kernel, dim3(grid), dim3(256), 0, stream, A, B, C, probs, M_runtime, splitk, tiles_m);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
