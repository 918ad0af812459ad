/* physicsml_b200 — C ABI of the B200-native MACE / NequIP message-passing hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + the caller's CUDA stream, enqueues its
 * work on that stream without synchronising or allocating, returns 0 on success and a negative code otherwise and
 * never throws.  The CALLER owns every buffer (the PyTorch host side allocates through torch's caching allocator so
 * CUDA-graph capture works); the library owns nothing but its code.  All floating-point data is fp32, row-major.
 *
 * Each function names the reference (Exscientia/physicsml, paths relative to src/physicsml/) operation it replaces;
 * INTEGRATION.md shows the ctypes / torch.library binding a maintainer adds on the reference side.
 *
 * "spec id" arguments select a build-time generated specialisation (physicsml_b200/csrc/gen/registry.json maps an
 * irreps signature to its id; physicsml_b200/codegen.py documents the signature strings).
 */
#ifndef PHYSICSML_B200_H
#define PHYSICSML_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pml_stream_t; /* == cudaStream_t */

#define PML_OK 0
#define PML_ERR_BAD_ARG (-1)
#define PML_ERR_UNSUPPORTED (-2)
#define PML_ERR_LAUNCH (-3)

int pml_abi_version(void);
int pml_num_tp_specs(void);
int pml_num_sym_specs(void);

/* ---- K2: edge featurisation ------------------------------------------------------------------------------------
 * Replaces compute_lengths_and_vectors (models/utils.py:71-95) + o3.SphericalHarmonics (models/mace/modules/mace.py:50-54,149;
 * models/nequip/modules/nequip.py:41-45,122) + RadialEmbeddingBlock (models/mace/modules/blocks.py:31-51, radial.py:6-49;
 * models/nequip/modules/radial.py:4-76).
 * pos [N,3]; snd/rcv [E] int32 (edge_index[0] / edge_index[1]); shift [E,3] or NULL; cell [3,3] or NULL;
 * bessel_w [nb].  B_n = prefactor * sin(bessel_w[n]*w_scale*d)/d * f_cut(d).  Y [E,(lmax+1)^2], B [E,nb], len [E] or NULL. */
int pml_edge_featurise_fwd(const float* pos, const int* snd, const int* rcv, const float* shift, const float* cell,
                           const float* bessel_w, float r_max, float prefactor, float w_scale, int p, int nb, int lmax,
                           float* Y, float* B, float* len, int E, pml_stream_t stream);
/* gpos [N,3] += J^T (gY, gB)  (caller zero-initialises gpos; gY / gB may be NULL) */
int pml_edge_featurise_bwd(const float* pos, const int* snd, const int* rcv, const float* shift, const float* cell,
                           const float* bessel_w, float r_max, float prefactor, float w_scale, int p, int nb, int lmax,
                           const float* gY, const float* gB, float* gpos, int E, pml_stream_t stream);
/* (tY, tB) = J t for a displacement field t [N,3]  (double backward of the force loss; lightning/module.py:28-48) */
int pml_edge_featurise_jvp(const float* pos, const int* snd, const int* rcv, const float* shift, const float* cell,
                           const float* bessel_w, float r_max, float prefactor, float w_scale, int p, int nb, int lmax,
                           const float* t, float* tY, float* tB, int E, pml_stream_t stream);
/* Gradient w.r.t. trainable Bessel frequencies (nequip/modules/radial.py:9-19, mace/modules/radial.py:17-20):
 * t == NULL: gw[n] += sum_e gB[e,n] dB_n/dw_n; t != NULL ([N,3]): gw[n] += sum_e gB[e,n] d2B_n/(dd dw_n) u.(t[r]-t[s])
 * (the parameter gradient of the force pass / of its tangent map).  gw [nb] zero-initialised by the caller; nb <= 16. */
int pml_bessel_weight_grad(const float* pos, const int* snd, const int* rcv, const float* shift, const float* cell,
                           const float* bessel_w, float r_max, float prefactor, float w_scale, int p, int nb,
                           const float* gB, const float* t, float* gw, int E, pml_stream_t stream);

/* ---- K4: fused channelwise tensor product + gather + segmented scatter-sum ----------------------------------------
 * Replaces w_h[sender] -> conv_tp -> scatter (models/mace/modules/blocks.py:214-226; models/nequip/modules/interaction_block.py:95-96).
 * h [N_src, C*DIN] (irreps layout), Y [E,NSH], R [E,NP*C], gather [E] int32, rowptr [N+1] int32 (CSR over the scatter
 * index), perm [E] int32 edge ids grouped by scatter index or NULL when edges are already grouped. out [N, C*DOUT]. */
int pml_tp_fwd(int spec, const float* h, const float* Y, const float* R, const int* gather, const int* rowptr,
               const int* perm, float* out, int N, int C, pml_stream_t stream);
/* Cotangents; any of dh [N_src,C*DIN] (zero-initialised by the caller, accumulated with fp32 atomics), dY [E,NSH]
 * (zero-initialised when C > 32), dR [E,NP*C] may be NULL to skip it. g [N, C*DOUT]. */
int pml_tp_bwd(int spec, const float* h, const float* Y, const float* R, const float* g, const int* gather,
               const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C, pml_stream_t stream);
/* Same kernels with a choice of layout inside every block of the output row `out` / its cotangent `g`:
 * out_layout bit 0: 0 = e3nn order u*(2l+1) + m (what pml_tp_fwd / pml_tp_bwd use), 1 = component-major m*C + u (packed
 * 8-byte stores; the per-irrep linear that consumes the row then reads its reduction index with unit stride).
 * out_layout bit 1 (pml_tp_bwd_ex only): dR is ADDED (L2 reductions) to the values already in the buffer instead of
 * overwritten — the second-order pass accumulates its two dR contributions without an R-sized add kernel. */
int pml_tp_fwd_ex(int spec, const float* h, const float* Y, const float* R, const int* gather, const int* rowptr,
                  const int* perm, float* out, int N, int C, int out_layout, pml_stream_t stream);
int pml_tp_bwd_ex(int spec, const float* h, const float* Y, const float* R, const float* g, const int* gather,
                  const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C, int out_layout,
                  pml_stream_t stream);
/* Kernel selection for K4 (diagnostics): by default the TMA-pipelined persistent kernels run whenever NP*C and DIN*C
 * are multiples of 4 floats (16-byte rows), h / R are 16-byte aligned and DOUT <= 48.  Bit 0 of v forces the
 * register-streaming kernels that also serve every other shape; bits 1 / 2 flip the producer placement of the
 * pipelined forward / cotangent kernel.  Returns the previous setting (initially env PML_TP_SIMPLE). */
int pml_tp_set_simple(int v);

/* ---- K6: symmetric contraction ------------------------------------------------------------------------------------
 * Replaces SymmetricContraction.forward (models/mace/modules/symmetric_contraction.py:75-77,218-239).
 * x [N,C,S]; attrs [N,Z] (one-hot rows take the fast path); out [N, C*DOUT] (hidden irreps layout). */
#define PML_SYM_MAX_OUT 3
#define PML_SYM_MAX_NU 3
typedef struct {
  const float* w[PML_SYM_MAX_OUT][PML_SYM_MAX_NU]; /* w[a][nu-1]: [Z, K[a][nu-1], C] */
  float* gw[PML_SYM_MAX_OUT][PML_SYM_MAX_NU];      /* gradient accumulators (zero-initialised) or NULL */
  int K[PML_SYM_MAX_OUT][PML_SYM_MAX_NU];
} pml_sym_weights;
int pml_sym_fwd(int spec, const float* x, const float* attrs, const pml_sym_weights* W, float* out, int N, int C, int Z,
                pml_stream_t stream);
int pml_sym_bwd(int spec, const float* x, const float* attrs, const pml_sym_weights* W, const float* g, float* dx, int N,
                int C, int Z, pml_stream_t stream);
/* directional second order: dg = J v, ddx = (g . H) v, W->gw += d/dW <g, J v> */
int pml_sym_dbl(int spec, const float* x, const float* v, const float* attrs, const pml_sym_weights* W, const float* g,
                float* dg, float* ddx, int N, int C, int Z, pml_stream_t stream);

/* ---- K5a / K3: grouped strided fp32 GEMM (o3.Linear, nn.FullyConnectedNet layers, weight gradients) ----------------
 * Replaces o3.Linear (models/mace/modules/blocks.py:139-144,176-181,65-70; mace.py:57-60; nequip interaction_block.py:22-27,66-71)
 * and the FullyConnectedNet matmuls (blocks.py:157-160; interaction_block.py:61-64). */
#define PML_GEMM_MAX_CHUNKS 8
typedef struct {
  long long a_off[PML_GEMM_MAX_CHUNKS];
  long long b_off[PML_GEMM_MAX_CHUNKS];
  int kc[PML_GEMM_MAX_CHUNKS];       /* < 0: runtime extent */
  float cscale[PML_GEMM_MAX_CHUNKS]; /* per-chunk scale */
  int nchunks;
  long long c_off;
  long long a_rs, a_cs, a_bs;
  long long b_rs, b_cs, b_bs;
  long long c_rs, c_cs, c_bs;
  int M, N; /* M < 0: runtime extent */
  int nbatch;
  float alpha;
  int accumulate;
} pml_gemm_problem;
/* probs: DEVICE pointer to nprobs problems. */
int pml_gemm_grouped(const float* A, const float* B, float* C, const pml_gemm_problem* probs, int nprobs, int M_runtime,
                     int max_M, int max_N, int max_batch, int splitk, pml_stream_t stream);

/* Same contract, executed on the 5th-gen tensor cores: tcgen05.mma kind::tf32 with TMEM accumulators and 3xTF32
 * error compensation (fp32-accurate).  This is the radial-MLP path (K3) and, above PML_TC_MIN_ROWS rows, every other
 * dense contraction.  max_k: upper bound of the reduction length one CTA sees (sum of a problem's chunk extents /
 * splitk; <= 0 unknown): reductions longer than 256 use the variant that promotes the TMEM accumulators into fp32
 * registers every 128 k, because the tensor core truncates its running sum at every MMA. */
int pml_gemm_grouped_tc(const float* A, const float* B, float* C, const pml_gemm_problem* probs, int nprobs, int M_runtime,
                        int max_M, int max_N, int max_batch, int splitk, int max_k, pml_stream_t stream);

/* ---- K3w: streaming kernels for the wide output layer of the radial MLP (reference blocks.py:157-160,211,
 * interaction_block.py:61-64,85: the last nn.FullyConnectedNet layer, hidden 64 -> paths x channels).
 * pml_wide_pack_w splits W [K <= 64, N] (element (k, n) at W[k*w_rs + n*w_cs]) into tf32 hi/lo parts laid out, per
 * 64-column tile, as the shared-memory image the tensor core reads (pml_wide_image_floats(N) floats);
 * pml_wide_fwd computes C[M, N] = alpha * A[M, K] @ W from that image with persistent A-stationary CTAs
 * (TMA-fed B ring, tcgen05 3xTF32, double-buffered TMEM accumulators).  Returns PML_ERR_UNSUPPORTED (-2) for shapes
 * outside K % 4 == 0, K <= 64, N % 64 == 0, 16-byte aligned rows: callers fall back to pml_gemm_grouped_tc. */
long long pml_wide_image_floats(int N);
int pml_wide_pack_w(const float* W, long long w_rs, long long w_cs, int K, int N, float* image, pml_stream_t stream);
int pml_wide_fwd(const float* A, long long lda, int K, const float* image, float* C, long long ldc, int M, int N,
                 float alpha, pml_stream_t stream);
/* dx of the same layer: dA[M, K] = alpha * G[M, N] @ W^T.  pml_wide_pack_wt writes one [64 x 32] hi/lo tile of W^T per
 * 32-wide chunk of N (same total size as the forward image); pml_wide_bwd_x streams G through a register-prefetched
 * shared-memory ring, promotes the TMEM accumulator into fp32 registers every 128 k.  N % 32 == 0. */
int pml_wide_pack_wt(const float* W, long long w_rs, long long w_cs, int K, int N, float* image, pml_stream_t stream);
int pml_wide_bwd_x(const float* G, long long ldg, int N, const float* image, float* dA, long long lda, int K, int M,
                   float alpha, pml_stream_t stream);

/* ---- K5b: element-indexed linear (FullyConnectedTensorProduct with one-hot node attributes) ------------------------
 * Replaces mix_layer / residual_connection_layer (models/mace/modules/blocks.py:183-188,72-78) and NequIP's
 * self_connection_layer (models/nequip/modules/interaction_block.py:73-82). */
#define PML_EL_MAX_PATHS 8
typedef struct {
  int npaths;
  int d[PML_EL_MAX_PATHS];
  int mul_in[PML_EL_MAX_PATHS];
  int mul_out[PML_EL_MAX_PATHS];
  long long w_off[PML_EL_MAX_PATHS];
  long long x_off[PML_EL_MAX_PATHS];
  long long o_off[PML_EL_MAX_PATHS];
  long long o_ws[PML_EL_MAX_PATHS];
  float alpha[PML_EL_MAX_PATHS];
  long long x_rs, o_rs;
} pml_el_plan;
int pml_element_linear_fwd(const float* x, const float* attrs, const float* W, const pml_el_plan* plan, float* out, int N,
                           int Z, int accumulate, pml_stream_t stream);
int pml_element_linear_bwd_x(const float* g, const float* attrs, const float* W, const pml_el_plan* plan, float* dx, int N,
                             int Z, int accumulate, pml_stream_t stream);
int pml_element_linear_bwd_w(const float* x, const float* g, const float* attrs, const pml_el_plan* plan, float* dW, int N,
                             int Z, int accumulate, pml_stream_t stream);

/* ---- small kernels ------------------------------------------------------------------------------------------------ */
/* kind 0 SiLU / 1 tanh; order 0: y=c f(x); 1: y=g c f'(x); 2: y=g v c f''(x)   (normalize2mom activations) */
int pml_act(const float* x, const float* g, const float* v, float c, int kind, int order, float* y, long long n,
            pml_stream_t stream);
/* ---- NequIP gate non-linearity -------------------------------------------------------------------------------------
 * Replaces Gate.forward (models/nequip/modules/_gate.py:131-152; built in convnet_layer.py:44-77): e3nn nn.Extract
 * slicing + Activation (normalize2mom SiLU / tanh) + o3.ElementwiseTensorProduct + concatenation in one pass over the
 * sorted convolution-output row.  mode 0: out0 = y [N,d_out]; mode 1: out0 = dx [N,d_in] given dy; mode 2: given v
 * (cotangent of dx) out0 = ddy [N,d_out], out1 = ddx [N,d_in]. */
#define PML_GATE_MAX_SEG 8
typedef struct {
  int n_scal;
  int s_in[PML_GATE_MAX_SEG], s_out[PML_GATE_MAX_SEG], s_mul[PML_GATE_MAX_SEG], s_kind[PML_GATE_MAX_SEG];
  float s_c[PML_GATE_MAX_SEG];
  int n_gated;
  int g_x[PML_GATE_MAX_SEG], g_gate[PML_GATE_MAX_SEG], g_out[PML_GATE_MAX_SEG], g_mul[PML_GATE_MAX_SEG];
  int g_d[PML_GATE_MAX_SEG], g_kind[PML_GATE_MAX_SEG];
  float g_c[PML_GATE_MAX_SEG];
  int d_in, d_out, items;
} pml_gate_plan;
int pml_gate(const float* x, const float* dy, const float* v, const pml_gate_plan* plan, int mode, float* out0,
             float* out1, int N, pml_stream_t stream);

/* ---- fused Adam (amsgrad, L2 weight decay) over all parameter tensors -------------------------------
 * Replaces torch.optim.Adam(..., amsgrad=True) as configured by the reference (models/mace/supervised/
 * default_configs.py:25-33).  tensors / chunks are DEVICE tables (chunks of 4096 elements); step is a device scalar that
 * the call increments before applying the update. */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* vmax;
  long long n;
} pml_adam_tensor;
typedef struct {
  int tensor;
  int pad;
  long long offset;
} pml_adam_chunk;
int pml_adam_amsgrad(const pml_adam_tensor* tensors, const pml_adam_chunk* chunks, int nchunks, float* step, float lr,
                     float b1, float b2, float one_minus_b1, float one_minus_b2, float eps, float wd, pml_stream_t stream);
/* Same update with per-tensor hyper-parameters and step counters: the reference builds one param group per parameter
 * with its own weight_decay (models/mace/supervised/mace_module.py:131-157).  gscale multiplies the gradient first
 * (1/world_size after a summed all-reduce).  steps: device table of the nsteps distinct counters, incremented first. */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* vmax;
  long long n;
  const float* step;
  float lr, b1, b2, omb1, omb2, eps, wd, gscale;
} pml_adam_tensor2;
int pml_adam_amsgrad_groups(const pml_adam_tensor2* tensors, const pml_adam_chunk* chunks, int nchunks, float* const* steps,
                            int nsteps, pml_stream_t stream);

/* Gradient packing for the data-parallel all-reduce (the reference delegates it to Lightning DDP,
 * lightning/model.py:159-163): flat[offset_t + i] = src_t ? src_t[i] : 0 for all tensors of a bucket in one launch
 * (chunks of 4096 elements as for pml_adam_amsgrad; tables in device memory). */
typedef struct {
  const float* src;
  long long offset;
  long long n;
} pml_pack_tensor;
int pml_pack_f32(const pml_pack_tensor* tensors, const pml_adam_chunk* chunks, int nchunks, float* flat, pml_stream_t stream);

/* Test utility: overwrite the shared memory of every SM with NaN (exposes kernels that read shared memory they never
 * wrote). */
int pml_debug_poison_smem(pml_stream_t stream);

/* per-graph pooled readout (models/mace/modules/mace.py:252-258): out[batch[n],f] += x[n,f]; out zero-initialised */
int pml_pool_sum(const float* x, const long long* batch, float* out, int N, int F, pml_stream_t stream);
int pml_pool_gather(const float* g, const long long* batch, float* out, int N, int F, pml_stream_t stream);
/* CSR row pointer over an int64 index vector (counts: zeroed scratch [N]) */
int pml_csr_rowptr(const long long* idx, int* counts, int* rowptr, int E, int N, pml_stream_t stream);
/* Edge ids grouped by an index vector, ascending inside every group (a stable counting sort: replaces the
 * torch.argsort(stable=True) a scatter over an unsorted index needs; blocks.py:221 / interaction_block.py:96 leave the
 * grouping to scatter_add's atomics).  cursor: zeroed scratch [N]; perm [E]. */
int pml_csr_group(const long long* idx, const int* rowptr, int* cursor, int* perm, int E, int N, pml_stream_t stream);
int pml_split_edge_index(const long long* edge_index, int* row0, int* row1, int E, pml_stream_t stream);
/* dst[n, j] = src[n, colmap[j]] (colmap: device int32 [dst_row]): component-major copies of o3.Linear operands
 * (blocks.py:139-144,176-181 store them in the e3nn order) so that the GEMM's reduction index has unit stride */
int pml_permute_columns(const float* src, long long src_row, const int* colmap, float* dst, int dst_row, long long N,
                        pml_stream_t stream);
/* ---- element-indexed linears on the tensor cores: nodes grouped by chemical element ------------------------------
 * o3.FullyConnectedTensorProduct(x, one-hot node_attrs) (mix_layer blocks.py:183-188,231-232; residual_connection_layer
 * blocks.py:72-78,88-92; NequIP self_connection_layer interaction_block.py:73-82,91-92) is one [mul_in, mul_out] matrix per
 * element: on nodes grouped by element every (element, irrep) pair is a dense GEMM over a contiguous row range.
 * pml_species_order: order[pos] = node ids grouped by element (ascending inside a group), seg[0..Z] = group starts,
 *   seg[Z+1] = 1 if some attrs row is not one-hot (its node is left out); Z <= 128, else PML_ERR_UNSUPPORTED.
 * pml_segment_problems: out[i] = tmpl[i] placed on group meta[i].v (all device pointers) - see pml_seg_meta; a batch that
 *   is not one-hot gets alpha = NaN in every problem.
 * pml_permute_rows_cols: scatter 0: dst[pos, j] = src[rows[pos], colmap[j]] ; scatter 1: dst[rows[pos], colmap[j]] = src[pos, j]
 *   for pos < N, j < ncols (rows / colmap: device int32). */
typedef struct {
  int v;    /* group (element) index into seg[] */
  int mode; /* 0: M = group size ; 1: kc[0] = group size (the reduction runs over the group) */
  long long a_mult, b_mult, c_mult; /* a_off / b_off / c_off grow by seg[v] * mult */
} pml_seg_meta;
int pml_species_order(const float* attrs, int N, int Z, int* order, int* seg, pml_stream_t stream);
int pml_segment_problems(const pml_gemm_problem* tmpl, const pml_seg_meta* meta, int nprob, const int* seg, int Z,
                         pml_gemm_problem* out, pml_stream_t stream);
int pml_permute_rows_cols(const float* src, long long src_row, const int* rows, const int* colmap, float* dst,
                          long long dst_row, int ncols, long long N, int scatter, pml_stream_t stream);
/* [N, C*S] irreps layout <-> [N, C, S] (models/mace/modules/irreps_tools.py:47-67) */
int pml_irreps_to_channel(const float* x, float* y, long long N, int C, int lmax, int inverse, pml_stream_t stream);

int pml_scan_i32(const int* counts, int* out, int n, pml_stream_t stream);
/* Fused force-matching loss (lightning/module.py:65-95, mace_module.py:159-185 with the stock MSE losses):
 * out[0] = w_e mean((e - e_t)^2) + w_f mean((-gpos - f_t)^2)   (forces = -gpos; ne / nf element counts); deterministic.
 * bwd: de = gl w_e 2 (e - e_t) / ne ; dgpos = gl w_f 2 (gpos + f_t) / nf   (gl: device scalar, the upstream gradient) */
int pml_ef_mse_fwd(const float* e, const float* e_t, int ne, const float* gpos, const float* f_t, int nf, float w_e, float w_f,
                   float* out, pml_stream_t stream);
int pml_ef_mse_bwd(const float* e, const float* e_t, int ne, const float* gpos, const float* f_t, int nf, float w_e, float w_f,
                   const float* gl, float* de, float* dgpos, pml_stream_t stream);
/* Verlet-skin test of a single-structure MD loop (plugins/openmm/openmm_graph.py:119-176 rebuilds the neighbour list on
 * every step): *out = max_i |pos_i - ref_i|^2 over N atoms; out: device float, zero-initialised by the caller. */
int pml_max_disp2(const float* pos, const float* ref, int N, float* out, pml_stream_t stream);

/* ---- K1: radius-graph neighbour list ------------------------------------------------------------------------------
 * Replaces compute_neighborlist_n2_no_cell / construct_edge_indices_and_attrs (lightning/graph_datasets/
 * neighbourhood_list_torch.py:33-96) and the vendored linked-cell list (torch_nl_vendored/neighbor_list.py:8-48,108-164,
 * linked_cell.py:129-355).  Two-phase (E is data dependent): *_count -> deg[N]; caller scans to rowptr[N+1], reads
 * E = rowptr[N], allocates; *_fill -> packed keys; sort_segments; emit.  Output sorted by (sender, receiver, shift);
 * PBC predicate d^2 < rc^2, open-boundary predicate d^2 <= rc^2; shifts refer to the unwrapped coordinates. */
/* Periodic path: the geometry of every structure of a (collated) batch is derived ON THE DEVICE by pml_nl_setup — cell
 * inverse in fp64, bins, stencil widths, a private bin range per structure — so there is no host round trip and a whole
 * batch of structures (one shared cell: cell_stride 0, the reference's batch contract graph_dataset.py:197-200,385; or
 * one cell each: cell_stride 9) is binned and walked in ONE pass.  graphs_out: opaque device table of
 * G * PML_NL_GRAPH_BYTES bytes; gptr [G+1] atom offsets; batch: int64 structure id per atom (sorted) or NULL when G == 1;
 * bin arrays sized by the caller with the bound 64*G + 4*N; *err (device, zeroed) gets bit 1 if a stencil would span
 * more than 127 images. */
#define PML_NL_GRAPH_BYTES 144
int pml_nl_setup(const float* cells, int cell_stride, const int* gptr, int G, float cutoff, void* graphs_out,
                 int* nbins_total, int* err, pml_stream_t stream);
int pml_nl_pbc_bin(const float* pos, const long long* batch, const void* graphs, float* wpos, int* fl, int* bin,
                   int* bincount, int N, pml_stream_t stream);
int pml_nl_pbc_fill_bins(const int* bin, const int* binptr, int* cursor, int* sorted, int N, pml_stream_t stream);
int pml_nl_pbc_count(const float* wpos, const long long* batch, const int* bin, const int* binptr, const int* sorted,
                     const void* graphs, float rc2, int self_interaction, int* deg, int N, pml_stream_t stream);
int pml_nl_pbc_fill(const float* wpos, const long long* batch, const int* bin, const int* binptr, const int* sorted,
                    const void* graphs, float rc2, int self_interaction, const int* rowptr, unsigned long long* keys, int N,
                    pml_stream_t stream);
int pml_nl_open_count(const float* pos, const long long* batch, const int* gptr, float rc2, int self_interaction, int* deg,
                      int N, pml_stream_t stream);
int pml_nl_open_fill(const float* pos, const long long* batch, const int* gptr, float rc2, int self_interaction,
                     const int* rowptr, unsigned long long* keys, int N, pml_stream_t stream);
int pml_nl_sort_segments(unsigned long long* keys, const int* rowptr, int N, int* err, pml_stream_t stream);
int pml_nl_emit(const unsigned long long* keys, const int* rowptr, const int* fl, long long* edge_index, float* shift,
                int* snd32, int* rcv32, int* perm, int N, int E, pml_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PHYSICSML_B200_H */
