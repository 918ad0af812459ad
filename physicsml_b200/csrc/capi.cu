// C-ABI dispatch layer (include/physicsml_b200.h): routes spec ids to the per-signature translation units.
#include <cuda_runtime.h>

#include "gen/pml_ids.h"
#include "pml_common.cuh"

struct pml_sym_weights;

#define DECL_TP(ID)                                                                                                      \
  extern "C" int pml_tp_fwd_launch_##ID(const float*, const float*, const float*, const int*, const int*, const int*,    \
                                        float*, int, int, int, cudaStream_t);                                            \
  extern "C" int pml_tp_bwd_launch_##ID(const float*, const float*, const float*, const float*, const int*, const int*,  \
                                        const int*, float*, float*, float*, int, int, int, cudaStream_t);
PML_TP_FOREACH(DECL_TP)
#undef DECL_TP

#define DECL_SYM(ID)                                                                                                     \
  extern "C" int pml_sym_fwd_launch_##ID(const float*, const float*, const pml_sym_weights*, float*, int, int, int,     \
                                         cudaStream_t);                                                                  \
  extern "C" int pml_sym_bwd_launch_##ID(const float*, const float*, const pml_sym_weights*, const float*, float*, int, \
                                         int, int, cudaStream_t);                                                        \
  extern "C" int pml_sym_dbl_launch_##ID(const float*, const float*, const float*, const pml_sym_weights*, const float*, \
                                         float*, float*, int, int, int, cudaStream_t);
PML_SYM_FOREACH(DECL_SYM)
#undef DECL_SYM

extern "C" int pml_abi_version(void) { return 1; }

// K4 kernel selection (A/B measurements, tests of the path taken when rows are not 16-byte multiples).  Bit 0: use the
// register-streaming kernels; bit 1 / bit 2: flip the producer placement of the pipelined forward / cotangent kernel
// (default: dedicated producer warp in the forward, warp 0 doubles as producer in the cotangent kernel); bit 3: signatures
// with several path groups run one launch per group instead of all groups inside one CTA.
// Initial value: env PML_TP_SIMPLE (integer).
#include <stdlib.h>
static int g_tp_mode = -1;
extern "C" int pml_tp_force_simple(void) {
  if (g_tp_mode < 0) {
    const char* e = getenv("PML_TP_SIMPLE");
    g_tp_mode = e ? atoi(e) & 15 : 0;
  }
  return g_tp_mode;
}
extern "C" int pml_tp_set_simple(int v) {
  const int old = pml_tp_force_simple();
  g_tp_mode = v & 15;
  return old;
}
extern "C" int pml_num_tp_specs(void) { return PML_TP_NUM_SPECS; }
extern "C" int pml_num_sym_specs(void) { return PML_SYM_NUM_SPECS; }

// out_layout bit 0: 0 = e3nn order inside every output block (u*(2l+1) + m), 1 = component-major (m*C + u);
// bit 1 (cotangent only): dR is ADDED to the values already in the buffer instead of overwritten
extern "C" int pml_tp_fwd_ex(int spec, const float* h, const float* Y, const float* R, const int* gather, const int* rowptr,
                             const int* perm, float* out, int N, int C, int out_layout, cudaStream_t stream) {
  switch (spec) {
#define CASE(ID) \
  case ID:       \
    return pml_tp_fwd_launch_##ID(h, Y, R, gather, rowptr, perm, out, N, C, out_layout, stream);
    PML_TP_FOREACH(CASE)
#undef CASE
    default:
      return PML_ERR_UNSUPPORTED;
  }
}

extern "C" int pml_tp_bwd_ex(int spec, const float* h, const float* Y, const float* R, const float* g, const int* gather,
                             const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C,
                             int out_layout, cudaStream_t stream) {
  switch (spec) {
#define CASE(ID) \
  case ID:       \
    return pml_tp_bwd_launch_##ID(h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, out_layout, stream);
    PML_TP_FOREACH(CASE)
#undef CASE
    default:
      return PML_ERR_UNSUPPORTED;
  }
}

extern "C" int pml_tp_fwd(int spec, const float* h, const float* Y, const float* R, const int* gather, const int* rowptr,
                          const int* perm, float* out, int N, int C, cudaStream_t stream) {
  return pml_tp_fwd_ex(spec, h, Y, R, gather, rowptr, perm, out, N, C, 0, stream);
}

extern "C" int pml_tp_bwd(int spec, const float* h, const float* Y, const float* R, const float* g, const int* gather,
                          const int* rowptr, const int* perm, float* dh, float* dY, float* dR, int N, int C,
                          cudaStream_t stream) {
  return pml_tp_bwd_ex(spec, h, Y, R, g, gather, rowptr, perm, dh, dY, dR, N, C, 0, stream);
}

extern "C" int pml_sym_fwd(int spec, const float* x, const float* attrs, const pml_sym_weights* W, float* out, int N,
                           int C, int Z, cudaStream_t stream) {
  switch (spec) {
#define CASE(ID) \
  case ID:       \
    return pml_sym_fwd_launch_##ID(x, attrs, W, out, N, C, Z, stream);
    PML_SYM_FOREACH(CASE)
#undef CASE
    default:
      return PML_ERR_UNSUPPORTED;
  }
}

extern "C" int pml_sym_bwd(int spec, const float* x, const float* attrs, const pml_sym_weights* W, const float* g,
                           float* dx, int N, int C, int Z, cudaStream_t stream) {
  switch (spec) {
#define CASE(ID) \
  case ID:       \
    return pml_sym_bwd_launch_##ID(x, attrs, W, g, dx, N, C, Z, stream);
    PML_SYM_FOREACH(CASE)
#undef CASE
    default:
      return PML_ERR_UNSUPPORTED;
  }
}

extern "C" int pml_sym_dbl(int spec, const float* x, const float* v, const float* attrs, const pml_sym_weights* W,
                           const float* g, float* dg, float* ddx, int N, int C, int Z, cudaStream_t stream) {
  switch (spec) {
#define CASE(ID) \
  case ID:       \
    return pml_sym_dbl_launch_##ID(x, v, attrs, W, g, dg, ddx, N, C, Z, stream);
    PML_SYM_FOREACH(CASE)
#undef CASE
    default:
      return PML_ERR_UNSUPPORTED;
  }
}
