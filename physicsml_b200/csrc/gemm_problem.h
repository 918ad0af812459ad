// Problem descriptor shared by the FFMA and the tcgen05 grouped GEMM kernels (mirrors include/physicsml_b200.h).
#pragma once
#define PML_GEMM_MAX_CHUNKS 8
extern "C" {
typedef struct {
  long long a_off[PML_GEMM_MAX_CHUNKS];  // element offset of chunk's A block (from A base)
  long long b_off[PML_GEMM_MAX_CHUNKS];  // element offset of chunk's B block (from B base)
  int kc[PML_GEMM_MAX_CHUNKS];           // K extent of each chunk (< 0: use the runtime extent passed to the launcher)
  float cscale[PML_GEMM_MAX_CHUNKS];     // per-chunk scale (multiplies alpha)
  int nchunks;
  long long c_off;
  long long a_rs, a_cs, a_bs;  // A element strides: row, k, batch
  long long b_rs, b_cs, b_bs;  // B element strides: k, col, batch
  long long c_rs, c_cs, c_bs;  // C element strides: row, col, batch
  int M, N;                    // M < 0: use the runtime extent passed to the launcher
  int nbatch;
  float alpha;
  int accumulate;  // 0: C = result ; 1: C += result (plain RMW, single writer) ; split-K always uses atomics
} pml_gemm_problem;
// how pml_segment_problems places a template problem on the row group of one element (mirrors include/physicsml_b200.h)
typedef struct {
  int v;                            // group (element) index into seg[]
  int mode;                         // 0: M = group size ; 1: kc[0] = group size
  long long a_mult, b_mult, c_mult; // offsets grow by seg[v] * mult (the row strides of the operands indexed by the group)
} pml_seg_meta;
}
