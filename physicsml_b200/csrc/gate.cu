// NequIP equivariant gate non-linearity (reference models/nequip/modules/_gate.py:33-163, convnet_layer.py:44-77):
// the convolution output row is a sorted irreps row holding scalars, gate scalars and gated (l > 0) blocks;
//     y_scalar = c_k phi_k(x_scalar)        y_gated[u, m] = x_gated[u, m] * c_k phi_k(x_gate[u])
// with phi = SiLU for even and tanh for odd scalars, c = e3nn normalize2mom constant.  One kernel each for the value,
// the cotangent map (differentiable once more: force-matching training differentiates forces) and the derivative of
// the cotangent map.  Slicing (e3nn nn.Extract), activation (Activation), o3.ElementwiseTensorProduct and the final
// concatenation of the reference are all folded into the index arithmetic of the plan: no intermediate tensors.
//
// Work item = (node, item), items = all scalar channels followed by all gated channels (u of every gated block).
#include "pml_common.cuh"

#define PML_GATE_MAX_SEG 8
extern "C" {
typedef struct {
  int n_scal;                       // scalar segments (activation only)
  int s_in[PML_GATE_MAX_SEG];       // input offset
  int s_out[PML_GATE_MAX_SEG];      // output offset
  int s_mul[PML_GATE_MAX_SEG];
  int s_kind[PML_GATE_MAX_SEG];     // 0 SiLU, 1 tanh
  float s_c[PML_GATE_MAX_SEG];
  int n_gated;                      // gated segments
  int g_x[PML_GATE_MAX_SEG];        // input offset of the gated block [mul, d]
  int g_gate[PML_GATE_MAX_SEG];     // input offset of its gate scalars [mul]
  int g_out[PML_GATE_MAX_SEG];      // output offset of the block
  int g_mul[PML_GATE_MAX_SEG];
  int g_d[PML_GATE_MAX_SEG];        // 2l+1
  int g_kind[PML_GATE_MAX_SEG];
  float g_c[PML_GATE_MAX_SEG];
  int d_in, d_out, items;           // row widths and work items per node
} pml_gate_plan;
}

namespace pml {

// c*phi, c*phi', c*phi''
__device__ __forceinline__ void act3(int kind, float c, float t, float& f0, float& f1, float& f2) {
  if (kind == 0) {
    const float s = 1.f / (1.f + expf(-t));
    f0 = c * t * s;
    f1 = c * s * (1.f + t * (1.f - s));
    f2 = c * s * (1.f - s) * (2.f + t * (1.f - 2.f * s));
  } else {
    const float th = tanhf(t);
    f0 = c * th;
    f1 = c * (1.f - th * th);
    f2 = c * (-2.f * th * (1.f - th * th));
  }
}

struct GateItem {
  bool gated;
  int seg, u;
};

__device__ __forceinline__ GateItem gate_item(const pml_gate_plan& P, int it) {
  GateItem r{false, 0, 0};
  for (int s = 0; s < P.n_scal; ++s) {
    if (it < P.s_mul[s]) {
      r.seg = s, r.u = it;
      return r;
    }
    it -= P.s_mul[s];
  }
  r.gated = true;
  for (int s = 0; s < P.n_gated; ++s) {
    if (it < P.g_mul[s]) {
      r.seg = s, r.u = it;
      return r;
    }
    it -= P.g_mul[s];
  }
  return r;
}

// MODE 0: y = gate(x)
// MODE 1: dx = J(x)^T dy                                    (every input column is written exactly once)
// MODE 2: given v (cotangent of dx): ddy = d<v, dx>/d dy , ddx = d<v, dx>/d x
template <int MODE>
__global__ void __launch_bounds__(256) gate_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                   const float* __restrict__ v, float* __restrict__ o0,
                                                   float* __restrict__ o1, const pml_gate_plan P, long long total) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / P.items;
  const GateItem w = gate_item(P, (int)(i - n * P.items));
  const float* xr = x + n * P.d_in;
  if (!w.gated) {
    const int ci = P.s_in[w.seg] + w.u, co = P.s_out[w.seg] + w.u;
    float f0, f1, f2;
    act3(P.s_kind[w.seg], P.s_c[w.seg], xr[ci], f0, f1, f2);
    if (MODE == 0) {
      o0[n * P.d_out + co] = f0;
    } else if (MODE == 1) {
      o0[n * P.d_in + ci] = dy[n * P.d_out + co] * f1;
    } else {
      const float vv = v[n * P.d_in + ci];
      o0[n * P.d_out + co] = vv * f1;                          // ddy
      o1[n * P.d_in + ci] = vv * dy[n * P.d_out + co] * f2;    // ddx
    }
    return;
  }
  const int d = P.g_d[w.seg];
  const int cx = P.g_x[w.seg] + w.u * d, cg = P.g_gate[w.seg] + w.u, co = P.g_out[w.seg] + w.u * d;
  float f0, f1, f2;
  act3(P.g_kind[w.seg], P.g_c[w.seg], xr[cg], f0, f1, f2);
  if (MODE == 0) {
    for (int m = 0; m < d; ++m) o0[n * P.d_out + co + m] = xr[cx + m] * f0;
  } else if (MODE == 1) {
    float dot = 0.f;
    for (int m = 0; m < d; ++m) {
      const float g = dy[n * P.d_out + co + m];
      o0[n * P.d_in + cx + m] = g * f0;
      dot = fmaf(g, xr[cx + m], dot);
    }
    o0[n * P.d_in + cg] = f1 * dot;
  } else {
    const float vg = v[n * P.d_in + cg];
    float dot_gx = 0.f, dot_vg = 0.f;
    for (int m = 0; m < d; ++m) {
      const float g = dy[n * P.d_out + co + m], xv = xr[cx + m], vx = v[n * P.d_in + cx + m];
      o0[n * P.d_out + co + m] = vg * f1 * xv + vx * f0;  // ddy
      o1[n * P.d_in + cx + m] = vg * f1 * g;              // ddx (gated part)
      dot_gx = fmaf(g, xv, dot_gx);
      dot_vg = fmaf(vx, g, dot_vg);
    }
    o1[n * P.d_in + cg] = vg * f2 * dot_gx + f1 * dot_vg;  // ddx (gate scalar)
  }
}

}  // namespace pml

// mode 0: out0 = y [N, d_out].  mode 1: out0 = dx [N, d_in].  mode 2: out0 = ddy [N, d_out], out1 = ddx [N, d_in].
// Every column of the outputs is written (the plan covers the whole input and output rows).
extern "C" int pml_gate(const float* x, const float* dy, const float* v, const pml_gate_plan* plan, int mode, float* out0,
                        float* out1, int N, cudaStream_t stream) {
  if (N <= 0) return PML_OK;
  if (!plan || plan->n_scal > PML_GATE_MAX_SEG || plan->n_gated > PML_GATE_MAX_SEG) return PML_ERR_BAD_ARG;
  const long long total = (long long)N * plan->items;
  const int grid = pml::ceil_div(total, 256);
  if (mode == 0) pml::launch_plain(pml::gate_kernel<0>, dim3(grid), dim3(256), 0, stream, x, dy, v, out0, out1, *plan, total);
  else if (mode == 1) pml::launch_plain(pml::gate_kernel<1>, dim3(grid), dim3(256), 0, stream, x, dy, v, out0, out1, *plan, total);
  else if (mode == 2) pml::launch_plain(pml::gate_kernel<2>, dim3(grid), dim3(256), 0, stream, x, dy, v, out0, out1, *plan, total);
  else return PML_ERR_BAD_ARG;
  PML_CHECK_LAUNCH();
  return PML_OK;
}
