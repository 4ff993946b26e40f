// opz_device.cuh — device-side constants, activations and the per-column physics of the NDE
// right-hand side.  Shared by the FP32 (CUDA-core) and the tcgen05 (tensor-core) kernels: in
// both, ONE THREAD OWNS ONE OCEAN COLUMN for the physics, walks its Nz cells bottom-to-top and
// evaluates the two-point stencils from values it already holds.
//
// Reference semantics restated (not copied) from
//   wind_mixing/src/training_postprocessing.jl:105-153 (predict_flux!, NDE!)  and
//   wind_mixing/src/NDE_training.jl:46-54,83-165     (local_richardson, tanh_step, predict_flux, predict_NDE)
//   src/differentiation_operators.jl:6-29            (D^c, D^f become differences * Nz)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/opz.h"

namespace opz {

constexpr int kMaxLayers = 6;   // Dense layers per net
constexpr int kMaxNz = 128;     // cells per column supported by the per-thread physics (FP32 path; the tensor path is Nz = 32)

// Host-precomputed FP32 constants.  Every product below is formed in FP32 in the same
// left-to-right order as the oracle so that the two FP32 paths round alike.
struct DevParams {
  int nz;
  uint32_t flags;
  float nzf;            // float(Nz): both difference operators multiply by 1/Delta = Nz
  float cBz;            // ((H*g)*alpha)*sigma_T
  float s_u, s_v;       // sigma_u, sigma_v
  float mu_u, mu_v;
  float eps;            // 0 unless OPZ_FLAG_RI_EPS
  float nu0, nu_m, Ric, dRi, Pr, kappa;
  float c_u, c_v, c_T;  // sigma_x / sigma_flux / H
  float off[3];         // sub * scale_flux(0) = sub * (-mu_flux / sigma_flux)
  float A_u, A_v, A_T;  // (-tau/H) * sigma_flux / sigma_x
  float cor_u, cor_v;   // f*tau/sigma_u, f*tau/sigma_v
  float tau;
  float omega;          // 2*pi/diurnal_period
  float alpha_g;        // alpha*g
  float mu_wT, s_wT;
  float diurnal_off;    // scale_wT(0) if OPZ_FLAG_DIURNAL_MINUS_SCALE0 else 0
  float two_inv_dRi, inv_Pr;  // fast-math physics of the tensor-core path
};

inline DevParams make_dev_params(const opz_params& p) {
  DevParams d{};
  d.nz = p.nz;
  d.flags = p.flags;
  d.nzf = (float)p.nz;
  d.cBz = ((p.H * p.g) * p.alpha) * p.sigma[2];
  d.s_u = p.sigma[0];
  d.s_v = p.sigma[1];
  d.mu_u = p.mu[0];
  d.mu_v = p.mu[1];
  d.eps = (p.flags & OPZ_FLAG_RI_EPS) ? p.eps : 0.0f;
  d.nu0 = p.nu0; d.nu_m = p.nu_m; d.Ric = p.Ric; d.dRi = p.dRi; d.Pr = p.Pr; d.kappa = p.kappa;
  d.c_u = p.sigma[0] / p.sigma[3] / p.H;
  d.c_v = p.sigma[1] / p.sigma[4] / p.H;
  d.c_T = p.sigma[2] / p.sigma[5] / p.H;
  const float sub = (p.flags & OPZ_FLAG_BC_MINUS_SCALE0) ? 1.0f : 0.0f;
  for (int i = 0; i < 3; ++i) d.off[i] = sub * ((-p.mu[3 + i]) / p.sigma[3 + i]);
  const float A = -p.tau / p.H;
  d.A_u = A * p.sigma[3] / p.sigma[0];
  d.A_v = A * p.sigma[4] / p.sigma[1];
  d.A_T = A * p.sigma[5] / p.sigma[2];
  const float ft = p.f * p.tau;
  d.cor_u = ft / p.sigma[0];
  d.cor_v = ft / p.sigma[1];
  d.tau = p.tau;
  d.omega = (float)(2.0 * 3.14159265358979323846 / (double)p.diurnal_period);
  d.alpha_g = p.alpha * p.g;
  d.mu_wT = p.mu[5];
  d.s_wT = p.sigma[5];
  d.diurnal_off = (p.flags & OPZ_FLAG_DIURNAL_MINUS_SCALE0) ? ((-p.mu[5]) / p.sigma[5]) : 0.0f;
  d.two_inv_dRi = 2.0f / p.dRi;
  d.inv_Pr = 1.0f / p.Pr;
  return d;
}

// ---- activations (NNlib 0.7.20 definitions restated) ---------------------------------------
template <int ACT>
__device__ __forceinline__ float act_fn(float z) {
  if constexpr (ACT == OPZ_ACT_RELU) return fmaxf(z, 0.0f);
  else if constexpr (ACT == OPZ_ACT_TANH) return tanhf(z);
  else if constexpr (ACT == OPZ_ACT_MISH) {  // z * tanh(softplus(z)); softplus in the stable form
    const float sp = fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z)));
    return z * tanhf(sp);
  } else if constexpr (ACT == OPZ_ACT_SWISH) return z / (1.0f + expf(-z));
  else if constexpr (ACT == OPZ_ACT_LEAKYRELU) return fmaxf(0.01f * z, z);
  else return z;
}
// Fast variants for the reduced-precision tensor-core path (results feed a 16-bit operand): one MUFU.EX2 + one MUFU.RCP
// each, ~1e-6 relative, flush-to-zero forms (no denormal range fix-ups) and NO data-dependent select: nvcc turns
// `z > 20 ? z : f(z)` into a branch per element, which serialised the 16 elements a thread converts per chunk
// (measured: 150 cycles per element on the 50/20 mish net).  The clamps make the formulas exact in their tails instead.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float act_fast(float z) {
  constexpr float kLog2e = 1.4426950408889634f;
  if constexpr (ACT == OPZ_ACT_RELU) return fmaxf(z, 0.0f);
  else if constexpr (ACT == OPZ_ACT_TANH) {  // 1 - 2 / (1 + e^{2z}); the exponent is clamped so that e^{2z} stays finite
    return 1.0f - 2.0f * rcp_ftz(1.0f + ex2_ftz(fminf(2.0f * kLog2e * z, 120.0f)));
  } else if constexpr (ACT == OPZ_ACT_MISH) {  // tanh(log(1+n)) = w/(w+2), w = n(n+2), n = e^z; for z >= 20 the ratio is 1.0f exactly
    const float n = ex2_ftz(fminf(z, 20.0f) * kLog2e);
    const float w = n * (n + 2.0f);
    return z * (w * rcp_ftz(w + 2.0f));
  } else if constexpr (ACT == OPZ_ACT_SWISH) {  // z / (1 + e^{-z}); clamped exponent: for z <= -83 the result is z * 2^-120 ~ -0
    return z * rcp_ftz(1.0f + ex2_ftz(fminf(-kLog2e * z, 120.0f)));
  } else if constexpr (ACT == OPZ_ACT_LEAKYRELU) return fmaxf(0.01f * z, z);
  else return z;
}
__device__ __forceinline__ float act_dyn(float z, int act) {
  switch (act) {
    case OPZ_ACT_RELU: return act_fn<OPZ_ACT_RELU>(z);
    case OPZ_ACT_TANH: return act_fn<OPZ_ACT_TANH>(z);
    case OPZ_ACT_MISH: return act_fn<OPZ_ACT_MISH>(z);
    case OPZ_ACT_SWISH: return act_fn<OPZ_ACT_SWISH>(z);
    case OPZ_ACT_LEAKYRELU: return act_fn<OPZ_ACT_LEAKYRELU>(z);
    default: return z;
  }
}
// derivative act'(z) given the pre-activation z (used by the BPTT kernels)
__device__ __forceinline__ float act_grad_dyn(float z, int act) {
  switch (act) {
    case OPZ_ACT_RELU: return z > 0.0f ? 1.0f : 0.0f;
    case OPZ_ACT_TANH: { const float t = tanhf(z); return 1.0f - t * t; }
    case OPZ_ACT_MISH: {
      const float sp = fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z)));
      const float t = tanhf(sp);
      const float sg = 1.0f / (1.0f + expf(-z));
      return t + z * (1.0f - t * t) * sg;
    }
    case OPZ_ACT_SWISH: { const float sg = 1.0f / (1.0f + expf(-z)); return sg + z * sg * (1.0f - sg); }
    case OPZ_ACT_LEAKYRELU: return z > 0.0f ? 1.0f : 0.01f;
    default: return 1.0f;
  }
}

// ---- diffusivity on one interior face --------------------------------------------------------
struct FaceDiff {
  float nu, nuT;
};
// FAST = true (tensor-core path only): approximate division / exponential; S2 = 0 still gives
// +-Inf / NaN, and (1 - tanh z)/2 is evaluated as 1/(1 + e^{2z}), which keeps RELATIVE accuracy
// in the stable tail where nu -> nu0.
template <bool FAST = false>
__device__ __forceinline__ float richardson(const DevParams& P, float gu, float gv, float gT) {
  const float Bz = P.cBz * (gT + P.eps);
  const float a = P.s_u * (gu + P.eps);
  const float b = P.s_v * (gv + P.eps);
  const float S2 = a * a + b * b;
  if constexpr (FAST) return __fdividef(Bz, S2);
  else return Bz / S2;  // IEEE division: S2 = 0 gives +-Inf / NaN exactly like the reference
}
template <bool FAST = false>
__device__ __forceinline__ float nu_of_ri(const DevParams& P, float Ri) {
  if constexpr (FAST) return P.nu0 + __fdividef(P.nu_m, 1.0f + __expf((Ri - P.Ric) * P.two_inv_dRi));
  else return P.nu0 + P.nu_m * ((1.0f - tanhf((Ri - P.Ric) / P.dRi)) / 2.0f);
}
template <bool FAST = false>
__device__ __forceinline__ float nuT_of(const DevParams& P, float nu, float gu, float gT) {
  const float nuT = FAST ? nu * P.inv_Pr : nu / P.Pr;
  if (P.flags & OPZ_FLAG_CA) {
    const float sel = (P.flags & OPZ_FLAG_CA_SELECT_DUDZ) ? gu : gT;
    return sel > 0.0f ? nuT : P.kappa;
  }
  return nuT;
}

__device__ __forceinline__ float wT_top_value(const DevParams& P, const float* bc, float t) {
  if (P.flags & OPZ_FLAG_DIURNAL) {
    const float phys = bc[6] * sinf(P.omega * (t * P.tau)) / P.alpha_g;
    return (phys - P.mu_wT) / P.s_wT - P.diurnal_off;
  }
  return bc[5] - P.off[2];
}

// ---- the per-column right-hand side ----------------------------------------------------------
// X(i): scaled state, i in [0,3Nz) = [u; v; T], cell 0 at the bottom.
// Y(net, j): NN output of net (0 uw, 1 vw, 2 wT) for interior face j+1, j in [0, Nz-1).
// K(i, value): receives d x_i / d(t/tau).   FL(v, f, value): optional total flux on face f.
// NZ_CT > 0 fixes Nz at compile time so that the face loop unrolls and X/Y may live in registers.
// SMOOTH = false compiles the two box-filter branches out (the tensor-core kernel rejects those flags).
// FAST = true selects the approximate division / exponential (tensor-core path).
template <int NZ_CT = 0, bool SMOOTH = true, bool FAST = false, class XF, class YF, class KF, class FLF>
__device__ __forceinline__ void column_rhs(const DevParams& P, const float* bc, float t, XF X, YF Y,
                                           KF K, FLF FL) {
  const int Nz = NZ_CT > 0 ? NZ_CT : P.nz;
  const bool mpp = P.flags & OPZ_FLAG_MPP;
  const bool smooth_nn = SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_NN);
  const bool smooth_ri = SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_RI);

  // bottom boundary fluxes
  float Fp_u = bc[0] - P.off[0];
  float Fp_v = bc[2] - P.off[1];
  float Fp_T = bc[4] - P.off[2];
  FL(0, 0, Fp_u); FL(1, 0, Fp_v); FL(2, 0, Fp_T);
  const float Ft_u = bc[1] - P.off[0];
  const float Ft_v = bc[3] - P.off[1];
  const float Ft_T = wT_top_value(P, bc, t);

  float u_lo = X(0), v_lo = X(Nz), T_lo = X(2 * Nz);  // cell f-1
  // Ri on the face below / at / above, only maintained when smooth_Ri is on.
  float Ri_m = 0.f, Ri_c = 0.f;
  if (smooth_ri && mpp) {
    Ri_m = richardson(P, 0.f, 0.f, 0.f);  // boundary face 0: zero gradients (D^f first row)
    Ri_c = richardson(P, (X(1) - u_lo) * P.nzf, (X(Nz + 1) - v_lo) * P.nzf, (X(2 * Nz + 1) - T_lo) * P.nzf);
  }
#pragma unroll
  for (int f = 1; f <= (NZ_CT > 0 ? NZ_CT : Nz); ++f) {
    float F_u, F_v, F_T;
    float u_hi = 0.f, v_hi = 0.f, T_hi = 0.f;
    if (f < Nz) {
      u_hi = X(f); v_hi = X(Nz + f); T_hi = X(2 * Nz + f);
      const int j = f - 1;
      float yu = Y(0, j), yv = Y(1, j), yT = Y(2, j);
      if (smooth_nn) {  // edge-shrinking 3-point box filter over the Nz-1 interior outputs
        const int n = Nz - 1;
        if (j == 0) { yu = (yu + Y(0, 1)) * 0.5f; yv = (yv + Y(1, 1)) * 0.5f; yT = (yT + Y(2, 1)) * 0.5f; }
        else if (j == n - 1) { yu = (Y(0, j - 1) + yu) * 0.5f; yv = (Y(1, j - 1) + yv) * 0.5f; yT = (Y(2, j - 1) + yT) * 0.5f; }
        else {
          const float third = 1.0f / 3.0f;
          yu = (Y(0, j - 1) + yu + Y(0, j + 1)) * third;
          yv = (Y(1, j - 1) + yv + Y(1, j + 1)) * third;
          yT = (Y(2, j - 1) + yT + Y(2, j + 1)) * third;
        }
      }
      F_u = yu; F_v = yv; F_T = yT;
      if (mpp) {
        const float gu = (u_hi - u_lo) * P.nzf;
        const float gv = (v_hi - v_lo) * P.nzf;
        const float gT = (T_hi - T_lo) * P.nzf;
        float Ri;
        if (smooth_ri) {
          float Ri_p;
          if (f + 1 < Nz) Ri_p = richardson(P, (X(f + 1) - u_hi) * P.nzf, (X(Nz + f + 1) - v_hi) * P.nzf, (X(2 * Nz + f + 1) - T_hi) * P.nzf);
          else Ri_p = richardson(P, 0.f, 0.f, 0.f);  // top boundary face
          Ri = (Ri_m + Ri_c + Ri_p) * (1.0f / 3.0f);
          Ri_m = Ri_c; Ri_c = Ri_p;
        } else {
          Ri = richardson<FAST>(P, gu, gv, gT);
        }
        const float nu = nu_of_ri<FAST>(P, Ri);
        const float nuT = nuT_of<FAST>(P, nu, gu, gT);
        F_u -= P.c_u * nu * gu;
        F_v -= P.c_v * nu * gv;
        F_T -= P.c_T * nuT * gT;
      }
    } else {
      F_u = Ft_u; F_v = Ft_v; F_T = Ft_T;
    }
    FL(0, f, F_u); FL(1, f, F_v); FL(2, f, F_T);
    const int c = f - 1;
    K(c,          P.A_u * ((F_u - Fp_u) * P.nzf) + P.cor_u * (P.s_v * v_lo + P.mu_v));
    K(Nz + c,     P.A_v * ((F_v - Fp_v) * P.nzf) - P.cor_v * (P.s_u * u_lo + P.mu_u));
    K(2 * Nz + c, P.A_T * ((F_T - Fp_T) * P.nzf));
    Fp_u = F_u; Fp_v = F_v; Fp_T = F_T;
    u_lo = u_hi; v_lo = v_hi; T_lo = T_hi;
  }
}

struct NoFlux {
  __device__ __forceinline__ void operator()(int, int, float) const {}
};

// ---- the same right-hand side, one interior face at a time ---------------------------------------------------------
// face_flux: total flux (NN + diffusive) on interior face f (1 <= f < Nz), from scratch — the expressions, and their order, of the
// face loop of column_rhs, so that a column split over several threads gives bit-identical tendencies.
template <bool SMOOTH = true, bool FAST = false, class XF, class YF>
__device__ __forceinline__ void face_flux(const DevParams& P, int Nz, XF X, YF Y, int f, float& F_u, float& F_v, float& F_T) {
  const bool mpp = P.flags & OPZ_FLAG_MPP;
  const bool smooth_nn = SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_NN);
  const bool smooth_ri = SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_RI);
  const float u_lo = X(f - 1), v_lo = X(Nz + f - 1), T_lo = X(2 * Nz + f - 1);
  const float u_hi = X(f), v_hi = X(Nz + f), T_hi = X(2 * Nz + f);
  const int j = f - 1;
  float yu = Y(0, j), yv = Y(1, j), yT = Y(2, j);
  if (smooth_nn) {
    const int n = Nz - 1;
    if (j == 0) { yu = (yu + Y(0, 1)) * 0.5f; yv = (yv + Y(1, 1)) * 0.5f; yT = (yT + Y(2, 1)) * 0.5f; }
    else if (j == n - 1) { yu = (Y(0, j - 1) + yu) * 0.5f; yv = (Y(1, j - 1) + yv) * 0.5f; yT = (Y(2, j - 1) + yT) * 0.5f; }
    else {
      const float third = 1.0f / 3.0f;
      yu = (Y(0, j - 1) + yu + Y(0, j + 1)) * third;
      yv = (Y(1, j - 1) + yv + Y(1, j + 1)) * third;
      yT = (Y(2, j - 1) + yT + Y(2, j + 1)) * third;
    }
  }
  F_u = yu; F_v = yv; F_T = yT;
  if (mpp) {
    const float gu = (u_hi - u_lo) * P.nzf;
    const float gv = (v_hi - v_lo) * P.nzf;
    const float gT = (T_hi - T_lo) * P.nzf;
    float Ri;
    if (smooth_ri) {
      const float Ri_c = richardson(P, gu, gv, gT);
      const float Ri_m = f - 1 >= 1 ? richardson(P, (u_lo - X(f - 2)) * P.nzf, (v_lo - X(Nz + f - 2)) * P.nzf, (T_lo - X(2 * Nz + f - 2)) * P.nzf)
                                    : richardson(P, 0.f, 0.f, 0.f);
      const float Ri_p = f + 1 < Nz ? richardson(P, (X(f + 1) - u_hi) * P.nzf, (X(Nz + f + 1) - v_hi) * P.nzf, (X(2 * Nz + f + 1) - T_hi) * P.nzf)
                                    : richardson(P, 0.f, 0.f, 0.f);
      Ri = (Ri_m + Ri_c + Ri_p) * (1.0f / 3.0f);
    } else {
      Ri = richardson<FAST>(P, gu, gv, gT);
    }
    const float nu = nu_of_ri<FAST>(P, Ri);
    const float nuT = nuT_of<FAST>(P, nu, gu, gT);
    F_u -= P.c_u * nu * gu;
    F_v -= P.c_v * nu * gv;
    F_T -= P.c_T * nuT * gT;
  }
}
// tendencies of cells [c_begin, c_end) of one column (cell c lies between faces c and c + 1)
template <bool SMOOTH = true, bool FAST = false, class XF, class YF, class KF>
__device__ __forceinline__ void column_rhs_range(const DevParams& P, const float* bc, float t, XF X, YF Y, KF K, int c_begin, int c_end) {
  const int Nz = P.nz;
  float Fp_u, Fp_v, Fp_T;
  if (c_begin == 0) { Fp_u = bc[0] - P.off[0]; Fp_v = bc[2] - P.off[1]; Fp_T = bc[4] - P.off[2]; }
  else face_flux<SMOOTH, FAST>(P, Nz, X, Y, c_begin, Fp_u, Fp_v, Fp_T);
  for (int c = c_begin; c < c_end; ++c) {
    const int f = c + 1;
    float F_u, F_v, F_T;
    if (f < Nz) face_flux<SMOOTH, FAST>(P, Nz, X, Y, f, F_u, F_v, F_T);
    else { F_u = bc[1] - P.off[0]; F_v = bc[3] - P.off[1]; F_T = wT_top_value(P, bc, t); }
    K(c,          P.A_u * ((F_u - Fp_u) * P.nzf) + P.cor_u * (P.s_v * X(Nz + c) + P.mu_v));
    K(Nz + c,     P.A_v * ((F_v - Fp_v) * P.nzf) - P.cor_v * (P.s_u * X(c) + P.mu_u));
    K(2 * Nz + c, P.A_T * ((F_T - Fp_T) * P.nzf));
    Fp_u = F_u; Fp_v = F_v; Fp_T = F_T;
  }
}

// ---- model description on the device --------------------------------------------------------
struct NetDesc {
  int n_layers;               // Dense layers
  int sizes[kMaxLayers + 1];  // in, hidden..., out
  int act;
  int64_t P;                  // params per net
  int64_t w_off[kMaxLayers];  // offset of W_l inside one net's flat vector (Flux.destructure order)
  int64_t b_off[kMaxLayers];
};

}  // namespace opz
