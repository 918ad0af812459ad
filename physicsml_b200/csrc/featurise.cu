// Edge featurisation: edge vectors (with PBC shifts), real spherical harmonics up to lmax <= 3 (e3nn basis,
// normalize=True, normalization="component") and the Bessel radial basis times the polynomial cutoff — one kernel,
// one thread per edge — plus its cotangent (forces) and tangent (double backward) kernels.
//
// Replaces, per edge (reference paths under src/physicsml/models):
//   compute_lengths_and_vectors            utils.py:71-95
//   o3.SphericalHarmonics(...)             mace/modules/mace.py:50-54,149 ; nequip/modules/nequip.py:41-45,122
//   RadialEmbeddingBlock                   mace/modules/blocks.py:31-51 + radial.py:6-49 ; nequip/modules/radial.py:4-76
#include "pml_common.cuh"

namespace pml {

struct RadialCfg {
  float r_max;
  float prefactor;  // MACE sqrt(2/r_max); NequIP 2/r_max
  float w_scale;    // bessel argument = bessel_weights[n] * w_scale * d   (MACE 1, NequIP 1/r_max)
  int p;            // polynomial cutoff order
  int nb;           // number of Bessel functions
};

// Polynomial part of the harmonics in the unit vector (x,y,z) and its gradient w.r.t. (x,y,z) treated as free.
template <int LMAX>
__device__ __forceinline__ void sh_poly(float x, float y, float z, float (&Y)[(LMAX + 1) * (LMAX + 1)],
                                        float (&G)[(LMAX + 1) * (LMAX + 1)][3]) {
  Y[0] = 1.f;
  G[0][0] = G[0][1] = G[0][2] = 0.f;
  if (LMAX >= 1) {
    const float s3 = 1.7320508075688772f;
    Y[1] = s3 * x; Y[2] = s3 * y; Y[3] = s3 * z;
    G[1][0] = s3; G[1][1] = 0.f; G[1][2] = 0.f;
    G[2][0] = 0.f; G[2][1] = s3; G[2][2] = 0.f;
    G[3][0] = 0.f; G[3][1] = 0.f; G[3][2] = s3;
  }
  if (LMAX >= 2) {
    const float a = 3.872983346207417f;   // sqrt(15)
    const float b = 2.23606797749979f;    // sqrt(5)
    Y[4] = a * x * z;                          G[4][0] = a * z;  G[4][1] = 0.f;       G[4][2] = a * x;
    Y[5] = a * x * y;                          G[5][0] = a * y;  G[5][1] = a * x;     G[5][2] = 0.f;
    Y[6] = b * (y * y - 0.5f * (x * x + z * z)); G[6][0] = -b * x; G[6][1] = 2.f * b * y; G[6][2] = -b * z;
    Y[7] = a * y * z;                          G[7][0] = 0.f;    G[7][1] = a * z;     G[7][2] = a * y;
    Y[8] = 0.5f * a * (z * z - x * x);         G[8][0] = -a * x; G[8][1] = 0.f;       G[8][2] = a * z;
  }
  if (LMAX >= 3) {
    const float c0 = 2.091650066335189f;   // sqrt(70)/4
    const float c1 = 10.246950765959598f;  // sqrt(105)
    const float c2 = 1.6201851746019651f;  // sqrt(42)/4
    const float c3 = 1.3228756555322954f;  // sqrt(7)/2
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    Y[9] = c0 * x * (3.f * z2 - x2);
    G[9][0] = c0 * 3.f * (z2 - x2); G[9][1] = 0.f; G[9][2] = c0 * 6.f * x * z;
    Y[10] = c1 * x * y * z;
    G[10][0] = c1 * y * z; G[10][1] = c1 * x * z; G[10][2] = c1 * x * y;
    Y[11] = c2 * x * (4.f * y2 - x2 - z2);
    G[11][0] = c2 * (4.f * y2 - 3.f * x2 - z2); G[11][1] = c2 * 8.f * x * y; G[11][2] = -c2 * 2.f * x * z;
    Y[12] = c3 * y * (2.f * y2 - 3.f * x2 - 3.f * z2);
    G[12][0] = -c3 * 6.f * x * y; G[12][1] = c3 * (6.f * y2 - 3.f * x2 - 3.f * z2); G[12][2] = -c3 * 6.f * y * z;
    Y[13] = c2 * z * (4.f * y2 - x2 - z2);
    G[13][0] = -c2 * 2.f * x * z; G[13][1] = c2 * 8.f * y * z; G[13][2] = c2 * (4.f * y2 - x2 - 3.f * z2);
    Y[14] = 0.5f * c1 * y * (z2 - x2);
    G[14][0] = -c1 * x * y; G[14][1] = 0.5f * c1 * (z2 - x2); G[14][2] = c1 * y * z;
    Y[15] = c0 * z * (z2 - 3.f * x2);
    G[15][0] = -c0 * 6.f * x * z; G[15][1] = 0.f; G[15][2] = c0 * 3.f * (z2 - x2);
  }
}

__device__ __forceinline__ float ipow(float x, int n) {
  float r = 1.f;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// polynomial cutoff f(d) and f'(d)
__device__ __forceinline__ void cutoff(float d, const RadialCfg& c, float& f, float& df) {
  if (!(d < c.r_max)) {
    f = 0.f;
    df = 0.f;
    return;
  }
  const float p = (float)c.p;
  const float x = d / c.r_max;
  const float xp1 = ipow(x, c.p - 1);  // x^(p-1)
  const float xp = xp1 * x;
  f = 1.f - ((p + 1.f) * (p + 2.f) * 0.5f) * xp + p * (p + 2.f) * xp * x - (p * (p + 1.f) * 0.5f) * xp * x * x;
  const float om = 1.f - x;
  df = -(p * (p + 1.f) * (p + 2.f) * 0.5f / c.r_max) * xp1 * om * om;
}

__device__ __forceinline__ void edge_vec(const float* __restrict__ pos, const int* __restrict__ snd,
                                         const int* __restrict__ rcv, const float* __restrict__ shift,
                                         const float* __restrict__ cell, int e, float (&v)[3], int& s, int& r) {
  s = __ldg(snd + e);
  r = __ldg(rcv + e);
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a] = __ldg(pos + 3 * (size_t)r + a) - __ldg(pos + 3 * (size_t)s + a);
  if (shift != nullptr && cell != nullptr) {
    const float s0 = __ldg(shift + 3 * (size_t)e), s1 = __ldg(shift + 3 * (size_t)e + 1), s2 = __ldg(shift + 3 * (size_t)e + 2);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // same association order as utils.py:85-91: ((r + s0*c0) + s1*c1) + s2*c2, no fused contraction
      v[a] = __fadd_rn(__fadd_rn(__fadd_rn(v[a], __fmul_rn(s0, __ldg(cell + a))), __fmul_rn(s1, __ldg(cell + 3 + a))),
                       __fmul_rn(s2, __ldg(cell + 6 + a)));
    }
  }
}

template <int LMAX>
__global__ void __launch_bounds__(128) featurise_fwd_kernel(const float* __restrict__ pos, const int* __restrict__ snd,
                                                            const int* __restrict__ rcv, const float* __restrict__ shift,
                                                            const float* __restrict__ cell,
                                                            const float* __restrict__ bessel_w, RadialCfg cfg,
                                                            float* __restrict__ Yout, float* __restrict__ Bout,
                                                            float* __restrict__ len_out, int E) {
  constexpr int NSH = (LMAX + 1) * (LMAX + 1);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float v[3];
  int s, r;
  edge_vec(pos, snd, rcv, shift, cell, e, v, s, r);
  const float d = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const float inv = 1.f / fmaxf(d, 1e-12f);
  float Y[NSH], G[NSH][3];
  sh_poly<LMAX>(v[0] * inv, v[1] * inv, v[2] * inv, Y, G);
#pragma unroll
  for (int j = 0; j < NSH; ++j) Yout[(size_t)e * NSH + j] = Y[j];
  const float dc = fmaxf(d, 1e-8f);
  float f, df;
  cutoff(dc, cfg, f, df);
  for (int n = 0; n < cfg.nb; ++n) {
    const float w = __ldg(bessel_w + n) * cfg.w_scale;
    Bout[(size_t)e * cfg.nb + n] = cfg.prefactor * (sinf(w * dc) / dc) * f;
  }
  if (len_out != nullptr) len_out[e] = dc;
}

// d(loss)/d(pos) from the cotangents of Y and B:  accumulates +gvec on the receiver and -gvec on the sender.
template <int LMAX>
__global__ void __launch_bounds__(128) featurise_bwd_kernel(const float* __restrict__ pos, const int* __restrict__ snd,
                                                            const int* __restrict__ rcv, const float* __restrict__ shift,
                                                            const float* __restrict__ cell,
                                                            const float* __restrict__ bessel_w, RadialCfg cfg,
                                                            const float* __restrict__ gY, const float* __restrict__ gB,
                                                            float* __restrict__ gpos, int E) {
  constexpr int NSH = (LMAX + 1) * (LMAX + 1);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float v[3];
  int s, r;
  edge_vec(pos, snd, rcv, shift, cell, e, v, s, r);
  const float d = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const float inv = 1.f / fmaxf(d, 1e-12f);
  const float u[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
  float gv[3] = {0.f, 0.f, 0.f};
  if (gY != nullptr) {
    float Y[NSH], G[NSH][3];
    sh_poly<LMAX>(u[0], u[1], u[2], Y, G);
    float gu[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 1; j < NSH; ++j) {
      const float g = __ldg(gY + (size_t)e * NSH + j);
      gu[0] = fmaf(g, G[j][0], gu[0]);
      gu[1] = fmaf(g, G[j][1], gu[1]);
      gu[2] = fmaf(g, G[j][2], gu[2]);
    }
    const float ud = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) gv[a] = (gu[a] - u[a] * ud) * inv;
  }
  if (gB != nullptr && d > 1e-8f) {
    float f, df;
    cutoff(d, cfg, f, df);
    float gd = 0.f;
    for (int n = 0; n < cfg.nb; ++n) {
      const float w = __ldg(bessel_w + n) * cfg.w_scale;
      float sn, cs;
      sincosf(w * d, &sn, &cs);
      const float dB = cfg.prefactor * (((w * cs * d - sn) / (d * d)) * f + (sn / d) * df);
      gd = fmaf(__ldg(gB + (size_t)e * cfg.nb + n), dB, gd);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) gv[a] = fmaf(gd, u[a], gv[a]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    atomicAdd(gpos + 3 * (size_t)r + a, gv[a]);
    atomicAdd(gpos + 3 * (size_t)s + a, -gv[a]);
  }
}

// Tangent map: given a displacement field t[N,3], the induced first-order change of Y and B per edge.  This is the
// transpose of featurise_bwd and is what the double-backward pass (force loss) propagates into the network.
template <int LMAX>
__global__ void __launch_bounds__(128) featurise_jvp_kernel(const float* __restrict__ pos, const int* __restrict__ snd,
                                                            const int* __restrict__ rcv, const float* __restrict__ shift,
                                                            const float* __restrict__ cell,
                                                            const float* __restrict__ bessel_w, RadialCfg cfg,
                                                            const float* __restrict__ t, float* __restrict__ tY,
                                                            float* __restrict__ tB, int E) {
  constexpr int NSH = (LMAX + 1) * (LMAX + 1);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float v[3];
  int s, r;
  edge_vec(pos, snd, rcv, shift, cell, e, v, s, r);
  const float d = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const float inv = 1.f / fmaxf(d, 1e-12f);
  const float u[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
  float dv[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) dv[a] = __ldg(t + 3 * (size_t)r + a) - __ldg(t + 3 * (size_t)s + a);
  const float udv = u[0] * dv[0] + u[1] * dv[1] + u[2] * dv[2];
  if (tY != nullptr) {
    float Y[NSH], G[NSH][3];
    sh_poly<LMAX>(u[0], u[1], u[2], Y, G);
    float du[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) du[a] = (dv[a] - u[a] * udv) * inv;
#pragma unroll
    for (int j = 0; j < NSH; ++j) tY[(size_t)e * NSH + j] = G[j][0] * du[0] + G[j][1] * du[1] + G[j][2] * du[2];
  }
  if (tB != nullptr) {
    float f = 0.f, df = 0.f;
    const bool ok = d > 1e-8f;
    if (ok) cutoff(d, cfg, f, df);
    for (int n = 0; n < cfg.nb; ++n) {
      float dB = 0.f;
      if (ok) {
        const float w = __ldg(bessel_w + n) * cfg.w_scale;
        float sn, cs;
        sincosf(w * d, &sn, &cs);
        dB = cfg.prefactor * (((w * cs * d - sn) / (d * d)) * f + (sn / d) * df);
      }
      tB[(size_t)e * cfg.nb + n] = dB * udv;
    }
  }
}

// Gradient w.r.t. the (trainable) Bessel frequencies bessel_w[n] (reference nequip/modules/radial.py:9-19, the NequIP
// default; mace/modules/radial.py:17-20).  With w = bessel_w[n] * w_scale and B_n = pref sin(w d)/d f(d):
//   t == null :  gw[n] += sum_e gB[e,n] * dB_n/dbw            = pref w_scale cos(w d) f
//   t != null :  gw[n] += sum_e gB[e,n] * d2B_n/(dd dbw) * (u . (t[r] - t[s]))  with
//                d2B_n/(dd dbw) = pref w_scale (cos(w d) f' - w sin(w d) f)
// The second form is the parameter gradient of the force pass (gpos is linear in dB/dd) and of its tangent map.
__global__ void __launch_bounds__(256) bessel_weight_grad_kernel(const float* __restrict__ pos, const int* __restrict__ snd,
                                                                 const int* __restrict__ rcv, const float* __restrict__ shift,
                                                                 const float* __restrict__ cell,
                                                                 const float* __restrict__ bessel_w, RadialCfg cfg,
                                                                 const float* __restrict__ gB, const float* __restrict__ t,
                                                                 float* __restrict__ gw, int E) {
  constexpr int MAXNB = 16;
  float acc[MAXNB];
#pragma unroll
  for (int n = 0; n < MAXNB; ++n) acc[n] = 0.f;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    float v[3];
    int s, r;
    edge_vec(pos, snd, rcv, shift, cell, e, v, s, r);
    const float d = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(d > 1e-8f)) continue;
    float f, df;
    cutoff(d, cfg, f, df);
    float udv = 1.f;
    if (t != nullptr) {
      const float inv = 1.f / d;
      udv = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) udv = fmaf(v[a] * inv, __ldg(t + 3 * (size_t)r + a) - __ldg(t + 3 * (size_t)s + a), udv);
    }
#pragma unroll
    for (int n = 0; n < MAXNB; ++n) {
      if (n < cfg.nb) {
        const float w = __ldg(bessel_w + n) * cfg.w_scale;
        float sn, cs;
        sincosf(w * d, &sn, &cs);
        const float k = (t == nullptr) ? cs * f : (cs * df - w * sn * f) * udv;
        acc[n] = fmaf(__ldg(gB + (size_t)e * cfg.nb + n), cfg.prefactor * cfg.w_scale * k, acc[n]);
      }
    }
  }
  __shared__ float red[MAXNB][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int n = 0; n < MAXNB; ++n) {
    const float x = warp_sum(acc[n]);
    if (lane == 0) red[n][wid] = x;
  }
  __syncthreads();
  if (threadIdx.x < cfg.nb) {
    float x = 0.f;
    for (int w = 0; w < 8; ++w) x += red[threadIdx.x][w];
    atomicAdd(gw + threadIdx.x, x);
  }
}

}  // namespace pml

#define PML_FEAT_DISPATCH(KERNEL, ...)                                                      \
  switch (lmax) {                                                                           \
    case 0: /* l=0 only: use the l<=1 code and let the caller slice */                      \
    case 1: pml::KERNEL<1><<<grid, 128, 0, stream>>>(__VA_ARGS__); break;                   \
    case 2: pml::KERNEL<2><<<grid, 128, 0, stream>>>(__VA_ARGS__); break;                   \
    case 3: pml::KERNEL<3><<<grid, 128, 0, stream>>>(__VA_ARGS__); break;                   \
    default: return PML_ERR_UNSUPPORTED;                                                    \
  }

static inline pml::RadialCfg make_cfg(float r_max, float prefactor, float w_scale, int p, int nb) {
  pml::RadialCfg c;
  c.r_max = r_max; c.prefactor = prefactor; c.w_scale = w_scale; c.p = p; c.nb = nb;
  return c;
}

// Y: [E, (lmax+1)^2], B: [E, nb], len (optional): [E].  shift/cell may be null (no PBC).  cell is row-major [3,3].
extern "C" int pml_edge_featurise_fwd(const float* pos, const int* snd, const int* rcv, const float* shift,
                                      const float* cell, const float* bessel_w, float r_max, float prefactor,
                                      float w_scale, int p, int nb, int lmax, float* Y, float* B, float* len, int E,
                                      cudaStream_t stream) {
  if (E <= 0) return PML_OK;
  if (lmax < 1) return PML_ERR_UNSUPPORTED;
  const int grid = pml::ceil_div(E, 128);
  const pml::RadialCfg cfg = make_cfg(r_max, prefactor, w_scale, p, nb);
  PML_FEAT_DISPATCH(featurise_fwd_kernel, pos, snd, rcv, shift, cell, bessel_w, cfg, Y, B, len, E)
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// gpos [N,3] must be zero-initialised (or hold a value to accumulate onto).  gY / gB may be null.
extern "C" int pml_edge_featurise_bwd(const float* pos, const int* snd, const int* rcv, const float* shift,
                                      const float* cell, const float* bessel_w, float r_max, float prefactor,
                                      float w_scale, int p, int nb, int lmax, const float* gY, const float* gB,
                                      float* gpos, int E, cudaStream_t stream) {
  if (E <= 0) return PML_OK;
  if (lmax < 1) return PML_ERR_UNSUPPORTED;
  const int grid = pml::ceil_div(E, 128);
  const pml::RadialCfg cfg = make_cfg(r_max, prefactor, w_scale, p, nb);
  PML_FEAT_DISPATCH(featurise_bwd_kernel, pos, snd, rcv, shift, cell, bessel_w, cfg, gY, gB, gpos, E)
  PML_CHECK_LAUNCH();
  return PML_OK;
}

extern "C" int pml_edge_featurise_jvp(const float* pos, const int* snd, const int* rcv, const float* shift,
                                      const float* cell, const float* bessel_w, float r_max, float prefactor,
                                      float w_scale, int p, int nb, int lmax, const float* t, float* tY, float* tB,
                                      int E, cudaStream_t stream) {
  if (E <= 0) return PML_OK;
  if (lmax < 1) return PML_ERR_UNSUPPORTED;
  const int grid = pml::ceil_div(E, 128);
  const pml::RadialCfg cfg = make_cfg(r_max, prefactor, w_scale, p, nb);
  PML_FEAT_DISPATCH(featurise_jvp_kernel, pos, snd, rcv, shift, cell, bessel_w, cfg, t, tY, tB, E)
  PML_CHECK_LAUNCH();
  return PML_OK;
}

// gw [nb] must be zero-initialised (or hold a value to accumulate onto); t [N,3] may be null (see the kernel).  nb <= 16.
extern "C" int pml_bessel_weight_grad(const float* pos, const int* snd, const int* rcv, const float* shift,
                                      const float* cell, const float* bessel_w, float r_max, float prefactor,
                                      float w_scale, int p, int nb, const float* gB, const float* t, float* gw, int E,
                                      cudaStream_t stream) {
  if (E <= 0) return PML_OK;
  if (nb > 16 || nb < 1) return PML_ERR_UNSUPPORTED;
  const pml::RadialCfg cfg = make_cfg(r_max, prefactor, w_scale, p, nb);
  int grid = pml::ceil_div(E, 256);
  if (grid > 1184) grid = 1184;  // 8 CTAs x 148 SMs: grid-stride beyond that, fewer atomics
  pml::bessel_weight_grad_kernel<<<grid, 256, 0, stream>>>(pos, snd, rcv, shift, cell, bessel_w, cfg, gB, t, gw, E);
  PML_CHECK_LAUNCH();
  return PML_OK;
}
