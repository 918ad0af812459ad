// Radius-graph neighbour list on the GPU: PBC-aware cell list (general triclinic cell) and the non-PBC per-graph
// variant, both emitting a DETERMINISTIC edge set sorted by (sender, receiver, shift), plus the receiver-grouped
// permutation the fused edge kernels want (obtained from the symmetry (i,j,S) <-> (j,i,-S), no global sort).
//
// Replaces (reference paths under src/physicsml/lightning/graph_datasets):
//   compute_neighborlist_n2_no_cell            neighbourhood_list_torch.py:33-50   (dense N x N, d^2 <= rc^2, i != j)
//   construct_edge_indices_and_attrs           neighbourhood_list_torch.py:53-96   (wrap, linked cell, shift fix-up)
//   compute_neighborlist / linked_cell / strict_nl   torch_nl_vendored/neighbor_list.py:8-48,108-164, linked_cell.py:129-355
// The reference replicates every atom over (2r+1)^3 images, builds dense [nbins, max_per_bin] tables and loops over
// occupied bins on the host with several device syncs.  Here nothing is replicated: atoms are binned once in
// fractional coordinates and periodic images are generated on the fly from the stencil offset.
//
// Semantics kept bit-for-bit at the set level: PBC uses the strict predicate d^2 < rc^2, the open-boundary path
// d^2 <= rc^2; self-image edges (i -> i with a non-zero shift) are kept; shifts refer to the UNWRAPPED coordinates
// (r = pos[j] - pos[i] + S . cell), i.e. S = S_wrapped + floor_i - floor_j as in neighbourhood_list_torch.py:85-88.
#include "pml_common.cuh"

namespace pml {

// Geometry of ONE periodic structure of a batch, computed on the device (nl_setup_kernel): no host round trip for the
// cell inverse, and a collated batch of structures (each with its own cell, or all sharing one) is binned in one pass —
// structure g owns the bins [bin_off, bin_off + nb0*nb1*nb2), so atoms only ever meet atoms of their own structure.
struct NLCell {
  double inv[9];   // inverse cell, row-major: frac = pos . inv
  float cell[9];   // cell rows
  int nb[3];       // bins per axis (in fractional space)
  int R[3];        // stencil half-width per axis (in bins)
  int bin_off;     // first bin of this structure
  int pad;
};
static_assert(sizeof(NLCell) == 144, "pml_nl_graph layout");

// cells: [3,3] (cell_stride 0: shared by every structure — the reference's batch contract, graph_dataset.py:197-200) or
// [G,3,3] (cell_stride 9); gptr [G+1] atom offsets.  Same bin policy as the host version it replaces: bins at least
// cutoff wide (in perpendicular height), at most 128 per axis and max(64, 4 n_g) per structure, stencil half-width
// R = ceil(cutoff * nb / height).  err bit 1: a stencil spans more than 127 images (unsupported).
__global__ void nl_setup_kernel(const float* __restrict__ cells, int cell_stride, const int* __restrict__ gptr, int G,
                                float cutoff, NLCell* __restrict__ out, int* __restrict__ nbins_total, int* __restrict__ err) {
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float* cf = cells + (size_t)g * cell_stride;
    double c[9];
    for (int k = 0; k < 9; ++k) c[k] = (double)cf[k];
    const double det = c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) + c[2] * (c[3] * c[7] - c[4] * c[6]);
    const double id = 1.0 / det;
    NLCell r;
    r.inv[0] = (c[4] * c[8] - c[5] * c[7]) * id; r.inv[1] = (c[2] * c[7] - c[1] * c[8]) * id; r.inv[2] = (c[1] * c[5] - c[2] * c[4]) * id;
    r.inv[3] = (c[5] * c[6] - c[3] * c[8]) * id; r.inv[4] = (c[0] * c[8] - c[2] * c[6]) * id; r.inv[5] = (c[2] * c[3] - c[0] * c[5]) * id;
    r.inv[6] = (c[3] * c[7] - c[4] * c[6]) * id; r.inv[7] = (c[1] * c[6] - c[0] * c[7]) * id; r.inv[8] = (c[0] * c[4] - c[1] * c[3]) * id;
    for (int k = 0; k < 9; ++k) r.cell[k] = cf[k];
    double hgt[3];
    const int n_g = gptr[g + 1] - gptr[g];
    const int max_bins = max(64, 4 * n_g);
    for (int k = 0; k < 3; ++k) {  // perpendicular height along axis k = 1 / |column k of inv|
      hgt[k] = 1.0 / sqrt(r.inv[k] * r.inv[k] + r.inv[3 + k] * r.inv[3 + k] + r.inv[6 + k] * r.inv[6 + k]);
      r.nb[k] = max(1, min(128, (int)floor(hgt[k] / (double)cutoff)));
    }
    while ((long long)r.nb[0] * r.nb[1] * r.nb[2] > max_bins) {
      int k = 0;
      if (r.nb[1] > r.nb[k]) k = 1;
      if (r.nb[2] > r.nb[k]) k = 2;
      r.nb[k] = max(1, r.nb[k] / 2);
    }
    for (int k = 0; k < 3; ++k) {
      r.R[k] = (int)ceil((double)cutoff * r.nb[k] / hgt[k] - 1e-12);
      if (r.R[k] > 127) atomicOr(err, 2);
    }
    r.bin_off = 0;
    r.pad = 0;
    out[g] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int off = 0;
    for (int g = 0; g < G; ++g) {
      out[g].bin_off = off;
      off += out[g].nb[0] * out[g].nb[1] * out[g].nb[2];
    }
    *nbins_total = off;
  }
}

__device__ __forceinline__ unsigned long long pack_key(int j, int sx, int sy, int sz) {
  return ((unsigned long long)(unsigned)j << 32) | ((unsigned long long)(sx + 128) << 16) |
         ((unsigned long long)(sy + 128) << 8) | (unsigned long long)(sz + 128);
}

// ---- stage 1: wrap + bin ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nl_bin_kernel(const float* __restrict__ pos, const long long* __restrict__ batch,
                                                     const NLCell* __restrict__ graphs, float* __restrict__ wpos,
                                                     int* __restrict__ fl, int* __restrict__ bin, int* __restrict__ counts,
                                                     int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const NLCell& c = graphs[batch ? (int)batch[i] : 0];
  const double p[3] = {(double)pos[3 * i], (double)pos[3 * i + 1], (double)pos[3 * i + 2]};
  double f[3];
  int b[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    f[k] = p[0] * c.inv[k] + p[1] * c.inv[3 + k] + p[2] * c.inv[6 + k];
    const double fk = floor(f[k]);
    fl[3 * i + k] = (int)fk;
    f[k] -= fk;  // in [0, 1)
    int bk = (int)(f[k] * c.nb[k]);
    b[k] = bk >= c.nb[k] ? c.nb[k] - 1 : bk;
  }
  // wrapped cartesian position = frac . cell (fp64 accumulate, stored fp32)
#pragma unroll
  for (int a = 0; a < 3; ++a)
    wpos[3 * i + a] = (float)(f[0] * (double)c.cell[a] + f[1] * (double)c.cell[3 + a] + f[2] * (double)c.cell[6 + a]);
  const int bl = c.bin_off + (b[0] * c.nb[1] + b[1]) * c.nb[2] + b[2];
  bin[i] = bl;
  atomicAdd(counts + bl, 1);
}

__global__ void __launch_bounds__(256) nl_fill_bins_kernel(const int* __restrict__ bin, const int* __restrict__ binptr,
                                                           int* __restrict__ cursor, int* __restrict__ sorted, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int b = bin[i];
  sorted[binptr[b] + atomicAdd(cursor + b, 1)] = i;
}

// ---- stage 2: one warp per atom walks the stencil; FILL=false counts, FILL=true writes packed keys -------------------------
template <bool FILL>
__global__ void __launch_bounds__(256) nl_pbc_edges_kernel(const float* __restrict__ wpos, const long long* __restrict__ batch,
                                                           const int* __restrict__ bin, const int* __restrict__ binptr,
                                                           const int* __restrict__ sorted, const NLCell* __restrict__ graphs,
                                                           float rc2, int self_interaction, int* __restrict__ deg,
                                                           const int* __restrict__ rowptr, unsigned long long* __restrict__ keys,
                                                           int N) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= N) return;
  const NLCell& c = graphs[batch ? (int)batch[i] : 0];
  const float xi = wpos[3 * i], yi = wpos[3 * i + 1], zi = wpos[3 * i + 2];
  const int bl = bin[i] - c.bin_off;
  const int b2 = bl % c.nb[2], b1 = (bl / c.nb[2]) % c.nb[1], b0 = bl / (c.nb[2] * c.nb[1]);
  int n_out = 0;
  const int base = FILL ? rowptr[i] : 0;
  for (int d0 = -c.R[0]; d0 <= c.R[0]; ++d0) {
    const int t0 = b0 + d0;
    const int s0 = (t0 >= 0) ? t0 / c.nb[0] : -((-t0 + c.nb[0] - 1) / c.nb[0]);  // floor division
    const int q0 = t0 - s0 * c.nb[0];
    for (int d1 = -c.R[1]; d1 <= c.R[1]; ++d1) {
      const int t1 = b1 + d1;
      const int s1 = (t1 >= 0) ? t1 / c.nb[1] : -((-t1 + c.nb[1] - 1) / c.nb[1]);
      const int q1 = t1 - s1 * c.nb[1];
      for (int d2 = -c.R[2]; d2 <= c.R[2]; ++d2) {
        const int t2 = b2 + d2;
        const int s2 = (t2 >= 0) ? t2 / c.nb[2] : -((-t2 + c.nb[2] - 1) / c.nb[2]);
        const int q2 = t2 - s2 * c.nb[2];
        // image translation S . cell, accumulated in the reference's order without fused contraction
        float tr[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          tr[a] = __fadd_rn(__fadd_rn(__fmul_rn((float)s0, c.cell[a]), __fmul_rn((float)s1, c.cell[3 + a])),
                            __fmul_rn((float)s2, c.cell[6 + a]));
        const int nbin = c.bin_off + (q0 * c.nb[1] + q1) * c.nb[2] + q2;
        const int beg = binptr[nbin], end = binptr[nbin + 1];
        for (int q = beg; q < end; q += 32) {
          const int idx = q + lane;
          bool ok = false;
          int j = -1;
          if (idx < end) {
            j = sorted[idx];
            // strict_nl (neighbor_list.py:38-44): d2 = |pos_i - pos_j - S.cell|^2 , keep d2 < rc^2
            const float dx = __fsub_rn(__fsub_rn(xi, wpos[3 * j]), tr[0]);
            const float dy = __fsub_rn(__fsub_rn(yi, wpos[3 * j + 1]), tr[1]);
            const float dz = __fsub_rn(__fsub_rn(zi, wpos[3 * j + 2]), tr[2]);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            ok = d2 < rc2;
            if (j == i && s0 == 0 && s1 == 0 && s2 == 0 && !self_interaction) ok = false;
          }
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (FILL && ok) keys[base + n_out + __popc(m & ((1u << lane) - 1u))] = pack_key(j, s0, s1, s2);
          n_out += __popc(m);
        }
      }
    }
  }
  if (!FILL && lane == 0) deg[i] = n_out;
}

// ---- open boundaries: per-graph brute force, thread per atom, j ascending => already sorted by (i, j) -------------------
template <bool FILL>
__global__ void __launch_bounds__(128) nl_open_edges_kernel(const float* __restrict__ pos, const long long* __restrict__ batch,
                                                            const int* __restrict__ gptr, float rc2, int self_interaction,
                                                            int* __restrict__ deg, const int* __restrict__ rowptr,
                                                            unsigned long long* __restrict__ keys, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int g = batch ? (int)batch[i] : 0;
  const int beg = gptr[g], end = gptr[g + 1];
  const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
  int n_out = 0;
  const int base = FILL ? rowptr[i] : 0;
  for (int j = beg; j < end; ++j) {
    // neighbourhood_list_torch.py:38-39: ((pos_j - pos_i)^2).sum(-1) <= rc^2
    const float dx = __fsub_rn(pos[3 * j], xi), dy = __fsub_rn(pos[3 * j + 1], yi), dz = __fsub_rn(pos[3 * j + 2], zi);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (d2 <= rc2 && (j != i || self_interaction)) {
      if (FILL) keys[base + n_out] = pack_key(j, 0, 0, 0);
      ++n_out;
    }
  }
  if (!FILL) deg[i] = n_out;
}

// ---- per-sender segment sort (bitonic in shared memory, one CTA per atom, segments up to 2048 keys) ----------------------
constexpr int NL_SORT_MAX = 2048;
__global__ void __launch_bounds__(128) nl_sort_segments_kernel(unsigned long long* __restrict__ keys,
                                                               const int* __restrict__ rowptr, int N, int* __restrict__ err) {
  __shared__ unsigned long long s[NL_SORT_MAX];
  const int i = blockIdx.x;
  const int beg = rowptr[i], len = rowptr[i + 1] - beg;
  if (len <= 1) return;
  if (len > NL_SORT_MAX) {
    if (threadIdx.x == 0) atomicExch(err, 1);
    return;
  }
  int P = 2;
  while (P < len) P <<= 1;
  for (int t = threadIdx.x; t < P; t += blockDim.x) s[t] = t < len ? keys[beg + t] : ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < P; t += blockDim.x) {
        const int u = t ^ j;
        if (u > t) {
          const bool up = (t & k) == 0;
          const unsigned long long a = s[t], b = s[u];
          if ((a > b) == up) {
            s[t] = b;
            s[u] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < len; t += blockDim.x) keys[beg + t] = s[t];
}

// ---- emit: edge_index int64 [2,E], shift f32 [E,3] (unwrapped convention), int32 copies, receiver-grouped permutation ------
__global__ void __launch_bounds__(256) nl_emit_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ rowptr,
                                                      const int* __restrict__ fl, long long* __restrict__ edge_index,
                                                      float* __restrict__ shift, int* __restrict__ snd32,
                                                      int* __restrict__ rcv32, int* __restrict__ perm, int N, int E) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= N) return;
  const int beg = rowptr[i], end = rowptr[i + 1];
  for (int e = beg + lane; e < end; e += 32) {
    const unsigned long long k = keys[e];
    const int j = (int)(k >> 32);
    const int s0 = (int)((k >> 16) & 0xff) - 128, s1 = (int)((k >> 8) & 0xff) - 128, s2 = (int)(k & 0xff) - 128;
    edge_index[e] = i;
    edge_index[(size_t)E + e] = j;
    snd32[e] = i;
    rcv32[e] = j;
    if (shift != nullptr) {
      int f0 = 0, f1 = 0, f2 = 0;
      if (fl != nullptr) {
        f0 = fl[3 * i] - fl[3 * j];
        f1 = fl[3 * i + 1] - fl[3 * j + 1];
        f2 = fl[3 * i + 2] - fl[3 * j + 2];
      }
      shift[3 * (size_t)e] = (float)(s0 + f0);
      shift[3 * (size_t)e + 1] = (float)(s1 + f1);
      shift[3 * (size_t)e + 2] = (float)(s2 + f2);
    }
    if (perm != nullptr) {
      // reverse edge (j -> i, -S) lives in j's sorted segment: binary search. Slot (e - beg) of node i's INCOMING list.
      const unsigned long long want = pack_key(i, -s0, -s1, -s2);
      int lo = rowptr[j], hi = rowptr[j + 1] - 1, found = -1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned long long km = keys[mid];
        if (km == want) {
          found = mid;
          break;
        }
        if (km < want) lo = mid + 1;
        else hi = mid - 1;
      }
      perm[e] = found;  // -1 only if the edge set is not symmetric (cannot happen for a radius graph)
    }
  }
}

}  // namespace pml

// =========================================================================================================================
// C ABI.  Two-phase because E is data dependent:  *_count fills deg[N] and the caller builds rowptr (pml_scan_i32) and
// reads E = rowptr[N] on the host to allocate; *_fill then writes the sorted edge list.
// =========================================================================================================================
// graphs_out: device buffer of G * 144 bytes (pml_nl_graph, opaque to the caller); nbins_total / err: device ints
// (err zeroed by the caller; bit 1 = a stencil spans more than 127 images).  Bins per structure <= max(64, 4 n_g), so
// 64 G + 4 N is an upper bound the caller can size the bin arrays with before the kernel has run.
extern "C" int pml_nl_setup(const float* cells, int cell_stride, const int* gptr, int G, float cutoff, void* graphs_out,
                            int* nbins_total, int* err, cudaStream_t stream) {
  if (G <= 0) return PML_OK;
  if (cell_stride != 0 && cell_stride != 9) return PML_ERR_BAD_ARG;
  pml::nl_setup_kernel<<<1, 256, 0, stream>>>(cells, cell_stride, gptr, G, cutoff, reinterpret_cast<pml::NLCell*>(graphs_out),
                                              nbins_total, err);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// batch: int64 structure id per atom (sorted) or NULL (single structure); graphs from pml_nl_setup.
// wpos [N,3] f32, fl [N,3] i32, bin [N] i32, bincount [>= nbins] i32 (zeroed by the caller)
extern "C" int pml_nl_pbc_bin(const float* pos, const long long* batch, const void* graphs, float* wpos, int* fl, int* bin,
                              int* bincount, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_bin_kernel<<<pml::ceil_div(N, 256), 256, 0, stream>>>(pos, batch, reinterpret_cast<const pml::NLCell*>(graphs), wpos, fl,
                                                               bin, bincount, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// cursor [nbins] i32 zeroed by the caller; sorted [N] i32
extern "C" int pml_nl_pbc_fill_bins(const int* bin, const int* binptr, int* cursor, int* sorted, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_fill_bins_kernel<<<pml::ceil_div(N, 256), 256, 0, stream>>>(bin, binptr, cursor, sorted, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_nl_pbc_count(const float* wpos, const long long* batch, const int* bin, const int* binptr, const int* sorted,
                                const void* graphs, float rc2, int self_interaction, int* deg, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_pbc_edges_kernel<false><<<pml::ceil_div((long long)N * 32, 256), 256, 0, stream>>>(
      wpos, batch, bin, binptr, sorted, reinterpret_cast<const pml::NLCell*>(graphs), rc2, self_interaction, deg, nullptr, nullptr, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_nl_pbc_fill(const float* wpos, const long long* batch, const int* bin, const int* binptr, const int* sorted,
                               const void* graphs, float rc2, int self_interaction, const int* rowptr,
                               unsigned long long* keys, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_pbc_edges_kernel<true><<<pml::ceil_div((long long)N * 32, 256), 256, 0, stream>>>(
      wpos, batch, bin, binptr, sorted, reinterpret_cast<const pml::NLCell*>(graphs), rc2, self_interaction, nullptr, rowptr, keys, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// batch: int64 graph id per atom (sorted) or NULL (single structure); gptr: int32 [G+1] atom offsets per graph
extern "C" int pml_nl_open_count(const float* pos, const long long* batch, const int* gptr, float rc2, int self_interaction,
                                 int* deg, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_open_edges_kernel<false><<<pml::ceil_div(N, 128), 128, 0, stream>>>(pos, batch, gptr, rc2, self_interaction, deg,
                                                                             nullptr, nullptr, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_nl_open_fill(const float* pos, const long long* batch, const int* gptr, float rc2, int self_interaction,
                                const int* rowptr, unsigned long long* keys, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_open_edges_kernel<true><<<pml::ceil_div(N, 128), 128, 0, stream>>>(pos, batch, gptr, rc2, self_interaction, nullptr,
                                                                            rowptr, keys, N);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// sort each sender's segment by (receiver, shift); *err (device int, zeroed) is set if a segment exceeds 2048 keys
extern "C" int pml_nl_sort_segments(unsigned long long* keys, const int* rowptr, int N, int* err, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  pml::nl_sort_segments_kernel<<<N, 128, 0, stream>>>(keys, rowptr, N, err);
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// edge_index int64 [2,E]; shift f32 [E,3] or NULL; fl int32 [N,3] or NULL (open boundaries); snd32/rcv32 int32 [E];
// perm int32 [E] or NULL: perm[rowptr[n] + k] = id of the k-th edge ARRIVING at n (receiver-grouped order; rowptr is shared
// because in-degree == out-degree for a symmetric edge set)
extern "C" int pml_nl_emit(const unsigned long long* keys, const int* rowptr, const int* fl, long long* edge_index,
                           float* shift, int* snd32, int* rcv32, int* perm, int N, int E, cudaStream_t stream) {
  if (N <= 0 || E <= 0) return PML_OK;
  pml::nl_emit_kernel<<<pml::ceil_div((long long)N * 32, 256), 256, 0, stream>>>(keys, rowptr, fl, edge_index, shift, snd32,
                                                                                rcv32, perm, N, E);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
