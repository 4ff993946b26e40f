// opz_faces.cuh — one interior face of the right-hand side and of its vector-Jacobian product WITH the optional
// 3-point box filters of the reference's training path: smooth_NN on the Nz-1 NN outputs and smooth_Ri on the Nz+1
// Richardson numbers (wind_mixing/src/filtering_operators.jl:1-15, applied at NDE_training.jl:98-102,121-123).
//
// Used by the BPTT kernels (opz_bptt.cu, opz_bptt_cluster.cuh), which evaluate the physics with one thread per face from
// shared-memory copies of the stage state `xs` and of the tendency cotangent `kbar`, both laid out [entry][CP] for CP column
// slots.  A filter couples a face to its neighbours; instead of extra shared-memory passes and barriers inside those
// latency-bound sweeps, the thread of face f RE-EVALUATES what it needs of faces f-2 .. f+2 from xs / kbar (a few dozen
// FLOPs).  Both functions are out of line so that the unfiltered hot path keeps its registers.
//
// Filter S over n entries (edge-shrinking): row 0 = (a0 + a1)/2, row n-1 = (a[n-2] + a[n-1])/2, row i = (a[i-1] + a[i] + a[i+1])/3.
//   forward   y_s = S y (n = Nz-1 outputs);  Ri_s = S Ri (n = Nz+1 faces, boundary faces carry zero gradients; only the
//             interior faces 1..Nz-1 are consumed, all of them "thirds" rows)
//   backward  ybar = S^T Fbar;  Ribar_raw[f] = (1/3) sum over interior f' in {f-1, f, f+1} of Ribar_s[f']
// The recurrences without the filters are those of oracle/nde_adjoint.py; with them they are checked against the autograd
// twin (oracle/nde_oracle_torch.py, which carries both filters) in tests/test_gpu_bptt.py.
#pragma once
#include "opz_device.cuh"

namespace opz {
namespace faces {

struct View {          // shared-memory state of one column slot
  const float* xs;     // [3 nz][CP]
  const float* kb;     // [3 nz][CP] tendency cotangent (backward only)
  int CP, c, nz;
  __device__ __forceinline__ float x(int v, int cell) const { return xs[(v * nz + cell) * CP + c]; }
  __device__ __forceinline__ float k(int v, int cell) const { return kb[(v * nz + cell) * CP + c]; }
};

// gradient of variable v on face f (zero on the two boundary faces: first / last row of D^f)
__device__ __forceinline__ float grad(const DevParams& P, const View& V, int v, int f) {
  return (f > 0 && f < V.nz) ? (V.x(v, f) - V.x(v, f - 1)) * P.nzf : 0.0f;
}
__device__ __forceinline__ float ri_raw(const DevParams& P, const View& V, int f) {
  return richardson(P, grad(P, V, 0, f), grad(P, V, 1, f), grad(P, V, 2, f));
}
// Richardson number the diffusivity of interior face f sees
__device__ __forceinline__ float ri_seen(const DevParams& P, const View& V, int f) {
  if (P.flags & OPZ_FLAG_SMOOTH_RI) return (ri_raw(P, V, f - 1) + ri_raw(P, V, f) + ri_raw(P, V, f + 1)) * (1.0f / 3.0f);
  return ri_raw(P, V, f);
}
// cotangent of the total flux of variable v on interior face f
__device__ __forceinline__ float fbar(const DevParams& P, const View& V, int v, int f) {
  const float A = v == 0 ? P.A_u : (v == 1 ? P.A_v : P.A_T);
  return A * P.nzf * (V.k(v, f - 1) - V.k(v, f));
}
__device__ __forceinline__ float box_weight(int i, int n) { return (i == 0 || i == n - 1) ? 0.5f : (1.0f / 3.0f); }

// filtered NN output j of one net: y laid out [j][CP]
__device__ __forceinline__ float y_smooth(const float* y, int CP, int c, int j, int n) {
  if (j == 0) return (y[c] + y[CP + c]) * 0.5f;
  if (j == n - 1) return (y[(j - 1) * CP + c] + y[j * CP + c]) * 0.5f;
  return (y[(j - 1) * CP + c] + y[j * CP + c] + y[(j + 1) * CP + c]) * (1.0f / 3.0f);
}

// ---- forward: total fluxes (uw, vw, wT) on interior face f; yb = NN outputs [3][nout][CP] ----
__device__ __noinline__ void forward_filtered(const DevParams& P, const float* xs, const float* yb, int CP, int c, int f, float* F3) {
  const int nz = P.nz, nout = nz - 1, j = f - 1;
  const View V{xs, nullptr, CP, c, nz};
  const bool sn = P.flags & OPZ_FLAG_SMOOTH_NN;
  float F[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) F[v] = sn ? y_smooth(yb + v * nout * CP, CP, c, j, nout) : yb[(v * nout + j) * CP + c];
  if (P.flags & OPZ_FLAG_MPP) {
    const float gu = grad(P, V, 0, f), gv = grad(P, V, 1, f), gT = grad(P, V, 2, f);
    const float nu = nu_of_ri(P, ri_seen(P, V, f));
    const float nuT = nuT_of(P, nu, gu, gT);
    F[0] -= P.c_u * nu * gu;
    F[1] -= P.c_v * nu * gv;
    F[2] -= P.c_T * nuT * gT;
  }
  F3[0] = F[0]; F3[1] = F[1]; F3[2] = F[2];
}

// what back-propagation needs of one interior face f' : cotangent of nu, and of the Richardson number the face saw
struct NuBar {
  float nubar, ribar, th, ri, nu;
  bool keep;
};
__device__ __forceinline__ NuBar nu_bar(const DevParams& P, const View& V, int f) {
  NuBar r;
  const float gu = grad(P, V, 0, f), gv = grad(P, V, 1, f), gT = grad(P, V, 2, f);
  r.ri = ri_seen(P, V, f);
  r.th = tanhf((r.ri - P.Ric) / P.dRi);
  r.nu = P.nu0 + P.nu_m * ((1.0f - r.th) / 2.0f);
  r.keep = true;
  if (P.flags & OPZ_FLAG_CA) r.keep = ((P.flags & OPZ_FLAG_CA_SELECT_DUDZ) ? gu : gT) > 0.0f;
  r.nubar = -P.c_u * gu * fbar(P, V, 0, f) - P.c_v * gv * fbar(P, V, 1, f) - (r.keep ? P.c_T * gT * fbar(P, V, 2, f) / P.Pr : 0.0f);
  r.ribar = r.nubar * (-P.nu_m / (2.0f * P.dRi) * (1.0f - r.th * r.th));
  return r;
}

// ---- backward: interior face f.  dy3 = cotangent of NN outputs j = f-1 of the three nets; gb3 = cotangent of the three
//      face gradients; gp (optional) accumulates d loss / d (nu0, nu_m, dRi, Ric, Pr) ----
__device__ __noinline__ void backward_filtered(const DevParams& P, const float* xs, const float* kb, int CP, int c, int f,
                                               float* dy3, float* gb3, float* gp /* [5] or null */) {
  const int nz = P.nz, nout = nz - 1, j = f - 1;
  const View V{xs, kb, CP, c, nz};
  const float Fb[3] = {fbar(P, V, 0, f), fbar(P, V, 1, f), fbar(P, V, 2, f)};
  if (P.flags & OPZ_FLAG_SMOOTH_NN) {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      float s = box_weight(j, nout) * Fb[v];
      if (j > 0) s += box_weight(j - 1, nout) * fbar(P, V, v, f - 1);
      if (j < nout - 1) s += box_weight(j + 1, nout) * fbar(P, V, v, f + 1);
      dy3[v] = s;
    }
  } else {
    dy3[0] = Fb[0]; dy3[1] = Fb[1]; dy3[2] = Fb[2];
  }
  gb3[0] = gb3[1] = gb3[2] = 0.0f;
  if (!(P.flags & OPZ_FLAG_MPP)) return;
  const float gu = grad(P, V, 0, f), gv = grad(P, V, 1, f), gT = grad(P, V, 2, f);
  const NuBar me = nu_bar(P, V, f);
  float ribar = me.ribar;   // cotangent of the RAW Richardson number of face f
  if (P.flags & OPZ_FLAG_SMOOTH_RI) {
    float s = isfinite(me.ri) ? me.ribar : 0.0f;
    if (f - 1 >= 1) { const NuBar lo = nu_bar(P, V, f - 1); if (isfinite(lo.ri)) s += lo.ribar; }
    if (f + 1 <= nz - 1) { const NuBar hi = nu_bar(P, V, f + 1); if (isfinite(hi.ri)) s += hi.ribar; }
    ribar = s * (1.0f / 3.0f);
  }
  const float ea = P.s_u * (gu + P.eps), eb = P.s_v * (gv + P.eps);
  const float S2 = ea * ea + eb * eb;
  const float Ri = P.cBz * (gT + P.eps) / S2;    // raw Richardson number of this face
  const float dRi_dgT = P.cBz / S2;
  const bool ok = isfinite(Ri) && isfinite(dRi_dgT) && isfinite(me.ri);
  const float nuT = me.keep ? me.nu / P.Pr : P.kappa;
  gb3[0] = -P.c_u * me.nu * Fb[0] + (ok ? ribar * (-Ri * 2.0f * P.s_u * ea / S2) : 0.0f);
  gb3[1] = -P.c_v * me.nu * Fb[1] + (ok ? ribar * (-Ri * 2.0f * P.s_v * eb / S2) : 0.0f);
  gb3[2] = -P.c_T * nuT * Fb[2] + (ok ? ribar * dRi_dgT : 0.0f);
  if (gp) {
    const float sech2 = 1.0f - me.th * me.th;
    gp[0] += me.nubar;
    gp[1] += me.nubar * ((1.0f - me.th) / 2.0f);
    if (isfinite(me.ri)) {
      gp[2] += me.nubar * (P.nu_m / 2.0f * sech2 * (me.ri - P.Ric) / (P.dRi * P.dRi));
      gp[3] += me.nubar * (P.nu_m / (2.0f * P.dRi) * sech2);
    }
    if (me.keep) gp[4] += Fb[2] * (P.c_T * me.nu * gT / (P.Pr * P.Pr));
  }
}

}  // namespace faces
}  // namespace opz
