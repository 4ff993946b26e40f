// opz_fp32.cu — FP32 CUDA-core kernels: persistent fixed-step rollout, single RHS, batched MLP.
//
// One CTA owns a tile of M ocean columns for the WHOLE rollout: the state x_n, the RK
// accumulator and the stage input never leave shared memory between time steps; HBM sees the
// initial profiles once, the forcing once and the saved snapshots.  Per RHS the three nets run
// as batched FP32 GEMM chains (opz_simt.cuh) with the weights streamed from L2, then one thread
// per column evaluates the physics (opz_device.cuh::column_rhs) and folds the tendency straight
// into the Runge-Kutta registers.
//
// Replaces solve_NDE_mutating (wind_mixing/src/training_postprocessing.jl:55-159) with a
// fixed-step Euler / Heun / RK4 scheme (north_star replaces OrdinaryDiffEq's ROCK4).
#include <algorithm>
#include "opz_internal.h"
#include "opz_simt.cuh"

namespace opz {

struct Fp32Args {
  DevParams P;
  NetDesc nd;
  const float* w;   // [3P]
  const float* x0;
  const float* bc;
  float* out;
  float* aux;       // fluxes for the rhs kernel
  int64_t B;
  int scheme;
  float dt, t0;
  int n_steps, save_every, n_save;
  int max_hidden;
  int which_net;
  // profile diagnostics (opz_profile_diagnostics): rows are [column][snapshot]
  int rows_per_bc;   // > 1: forcing row = row / rows_per_bc, time = t0 + (row % rows_per_bc) * dt
  int skip_nn;       // 1: the NN outputs are taken as zero (the boundary + diffusive part of the fluxes)
};

template <int M>
struct Smem {
  float *xn, *kacc, *xs[2], *Y, *bufA, *bufB, *stage;
  __device__ Smem(float* base, int nz, int max_hidden) {
    const int S = 3 * nz * M;
    xn = base; base += S;
    kacc = base; base += S;
    xs[0] = base; base += S;
    xs[1] = base; base += S;
    Y = base; base += 3 * nz * M;
    bufA = base; base += max_hidden * M;
    bufB = base; base += max_hidden * M;
    stage = base;
  }
  static size_t bytes(int nz, int max_hidden) {
    return sizeof(float) * ((size_t)5 * 3 * nz * M + (size_t)2 * max_hidden * M + SimtCfg<M>::STAGE_TOTAL);
  }
};

template <int M>
__device__ __forceinline__ void run_three_nets(const Fp32Args& a, const Smem<M>& s, const float* in) {
  for (int net = 0; net < 3; ++net)
    mlp_chain_fp32<M>(a.nd, a.w + (size_t)net * a.nd.P, in, s.bufA, s.bufB, s.Y + (size_t)net * a.P.nz * M, s.stage);
}

template <int M>
__global__ void __launch_bounds__(kThreads, 1) rollout_fp32_kernel(const __grid_constant__ Fp32Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem<M> s(smem_f, a.P.nz, a.max_hidden);
  const int tid = threadIdx.x;
  const int nz = a.P.nz, S = 3 * nz;
  const int64_t n_tiles = (a.B + M - 1) / M;
  const float h = a.dt;
  const int stages = a.scheme == OPZ_RK4 ? 4 : (a.scheme == OPZ_RK2 ? 2 : 1);
  DevParams P_ex = a.P;   // OPZ_IMEX_EULER: the explicit part leaves the vertical diffusion out
  P_ex.flags &= ~(uint32_t)(OPZ_FLAG_MPP | OPZ_FLAG_CA | OPZ_FLAG_SMOOTH_RI);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t col0 = tile * M;
    // ---- load the initial profiles: global [col][i] -> shared [i][m] ----
    for (int e = tid; e < M * S; e += kThreads) {
      const int m = e / S, i = e - m * S;
      const int64_t col = col0 + m;
      const float v = col < a.B ? a.x0[col * S + i] : 0.0f;
      s.xn[i * M + m] = v;
      s.xs[0][i * M + m] = v;
    }
    float bc[OPZ_BC_STRIDE];
    if (tid < M) {
      const int64_t col = col0 + tid;
#pragma unroll
      for (int q = 0; q < OPZ_BC_STRIDE; ++q) bc[q] = col < a.B ? a.bc[col * OPZ_BC_STRIDE + q] : 0.0f;
    }
    __syncthreads();
    auto save = [&](int slot) {
      for (int e = tid; e < M * S; e += kThreads) {
        const int m = e / S, i = e - m * S;
        const int64_t col = col0 + m;
        if (col < a.B) a.out[(col * a.n_save + slot) * S + i] = s.xn[i * M + m];
      }
    };
    if (a.save_every > 0) save(0);

    int cur = 0;
    for (int n = 0; n < a.n_steps; ++n) {
      const float tn = a.t0 + (float)n * h;
      for (int st = 0; st < stages; ++st) {
        run_three_nets<M>(a, s, s.xs[cur]);  // ends with __syncthreads
        if (tid < M) {
          const float* xin = s.xs[cur];
          float* xnext = s.xs[cur ^ 1];
          float ts = tn;
          if (a.scheme == OPZ_RK4) ts = (st == 0) ? tn : (st == 3 ? tn + h : tn + h * 0.5f);
          else if (a.scheme == OPZ_RK2) ts = (st == 0) ? tn : tn + h;
          auto X = [&](int i) { return xin[i * M + tid]; };
          auto Y = [&](int net, int j) { return s.Y[(net * nz + j) * M + tid]; };
          const int scheme = a.scheme;
          auto K = [&](int i, float k) {
            const int o = i * M + tid;
            const float xo = s.xn[o];
            if (scheme == OPZ_RK4) {
              if (st == 0) { s.kacc[o] = k; xnext[o] = xo + (h * 0.5f) * k; }
              else if (st == 1) { s.kacc[o] = s.kacc[o] + 2.0f * k; xnext[o] = xo + (h * 0.5f) * k; }
              else if (st == 2) { s.kacc[o] = s.kacc[o] + 2.0f * k; xnext[o] = xo + h * k; }
              else { const float xv = xo + (h / 6.0f) * (s.kacc[o] + k); s.xn[o] = xv; xnext[o] = xv; }
            } else if (scheme == OPZ_RK2) {
              if (st == 0) { s.kacc[o] = k; xnext[o] = xo + h * k; }
              else { const float xv = xo + (h * 0.5f) * (s.kacc[o] + k); s.xn[o] = xv; xnext[o] = xv; }
            } else {
              const float xv = xo + h * k; s.xn[o] = xv; xnext[o] = xv;
            }
          };
          if (scheme == OPZ_IMEX_EULER) {
            // explicit part: everything but the vertical diffusion (flags masked), Euler update through K (the else branch)
            column_rhs(P_ex, bc, ts, X, Y, K, NoFlux());
            // implicit part: backward-Euler diffusion with the diffusivities of the updated state, one tridiagonal system per
            // variable solved by the Thomas algorithm, bottom to top (NDE_oceananigans.jl:61-101); s.kacc is free in this
            // scheme and holds the c' coefficients, the modified right-hand side overwrites the state in place
            if (a.P.flags & OPZ_FLAG_MPP) {
              const DevParams& P = a.P;
              const float coef_u = h * (-P.A_u) * P.c_u * P.nzf * P.nzf;   // = dt nu / dz^2 per unit nu
              const float coef_v = h * (-P.A_v) * P.c_v * P.nzf * P.nzf;
              const float coef_T = h * (-P.A_T) * P.c_T * P.nzf * P.nzf;
              auto XS = [&](int i) -> float& { return xnext[i * M + tid]; };
              auto CP = [&](int i) -> float& { return s.kacc[i * M + tid]; };
              float Dlo_u = 0.0f, Dlo_v = 0.0f, Dlo_T = 0.0f;     // D on the face below cell c
              float u_lo = XS(0), v_lo = XS(nz), T_lo = XS(2 * nz);   // ORIGINAL values of cell c (read before they are overwritten)
              // smooth_Ri: raw Richardson numbers of faces c and c + 1 roll along; face c + 2 is formed from cells c + 1, c + 2,
              // which the elimination has not touched yet
              const bool smooth_ri = P.flags & OPZ_FLAG_SMOOTH_RI;
              float Ri_m = 0.0f, Ri_c = 0.0f;
              if (smooth_ri) {
                Ri_m = richardson(P, 0.0f, 0.0f, 0.0f);
                Ri_c = nz > 1 ? richardson(P, (XS(1) - u_lo) * P.nzf, (XS(nz + 1) - v_lo) * P.nzf, (XS(2 * nz + 1) - T_lo) * P.nzf) : Ri_m;
              }
              for (int c = 0; c < nz; ++c) {
                float Dhi_u = 0.0f, Dhi_v = 0.0f, Dhi_T = 0.0f;   // face c + 1 (zero on the top boundary)
                float u_hi = 0.0f, v_hi = 0.0f, T_hi = 0.0f;
                if (c + 1 < nz) {
                  u_hi = XS(c + 1); v_hi = XS(nz + c + 1); T_hi = XS(2 * nz + c + 1);
                  const float gu = (u_hi - u_lo) * P.nzf, gv = (v_hi - v_lo) * P.nzf, gT = (T_hi - T_lo) * P.nzf;
                  float Ri = 0.0f;
                  if (smooth_ri) {
                    const float Ri_p = c + 2 < nz ? richardson(P, (XS(c + 2) - u_hi) * P.nzf, (XS(nz + c + 2) - v_hi) * P.nzf, (XS(2 * nz + c + 2) - T_hi) * P.nzf)
                                                  : richardson(P, 0.0f, 0.0f, 0.0f);
                    Ri = (Ri_m + Ri_c + Ri_p) * (1.0f / 3.0f);
                    Ri_m = Ri_c; Ri_c = Ri_p;
                  } else {
                    Ri = richardson(P, gu, gv, gT);
                  }
                  const float nu = nu_of_ri(P, Ri);
                  const float nuT = nuT_of(P, nu, gu, gT);
                  Dhi_u = coef_u * nu; Dhi_v = coef_v * nu; Dhi_T = coef_T * nuT;
                }
                // row c: -Dlo phi_{c-1} + (1 + Dlo + Dhi) phi_c - Dhi phi_{c+1} = phi*_c
                {
                  const float den = (1.0f + Dlo_u + Dhi_u) + Dlo_u * (c > 0 ? CP(c - 1) : 0.0f);
                  CP(c) = -Dhi_u / den;
                  XS(c) = (u_lo + Dlo_u * (c > 0 ? XS(c - 1) : 0.0f)) / den;
                }
                {
                  const float den = (1.0f + Dlo_v + Dhi_v) + Dlo_v * (c > 0 ? CP(nz + c - 1) : 0.0f);
                  CP(nz + c) = -Dhi_v / den;
                  XS(nz + c) = (v_lo + Dlo_v * (c > 0 ? XS(nz + c - 1) : 0.0f)) / den;
                }
                {
                  const float den = (1.0f + Dlo_T + Dhi_T) + Dlo_T * (c > 0 ? CP(2 * nz + c - 1) : 0.0f);
                  CP(2 * nz + c) = -Dhi_T / den;
                  XS(2 * nz + c) = (T_lo + Dlo_T * (c > 0 ? XS(2 * nz + c - 1) : 0.0f)) / den;
                }
                Dlo_u = Dhi_u; Dlo_v = Dhi_v; Dlo_T = Dhi_T;
                u_lo = u_hi; v_lo = v_hi; T_lo = T_hi;
              }
              for (int c = nz - 2; c >= 0; --c) {
                XS(c) = XS(c) - CP(c) * XS(c + 1);
                XS(nz + c) = XS(nz + c) - CP(nz + c) * XS(nz + c + 1);
                XS(2 * nz + c) = XS(2 * nz + c) - CP(2 * nz + c) * XS(2 * nz + c + 1);
              }
              for (int i = 0; i < 3 * nz; ++i) s.xn[i * M + tid] = XS(i);
            }
          } else {
            column_rhs(a.P, bc, ts, X, Y, K, NoFlux());
          }
        }
        cur ^= 1;
        __syncthreads();
      }
      if (a.save_every > 0 && (n + 1) % a.save_every == 0) save((n + 1) / a.save_every);
    }
    if (a.save_every <= 0) save(0);
    __syncthreads();
  }
}

template <int M>
__global__ void __launch_bounds__(kThreads, 1) rhs_fp32_kernel(const __grid_constant__ Fp32Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem<M> s(smem_f, a.P.nz, a.max_hidden);
  const int tid = threadIdx.x;
  const int nz = a.P.nz, S = 3 * nz;
  const int64_t n_tiles = (a.B + M - 1) / M;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t col0 = tile * M;
    for (int e = tid; e < M * S; e += kThreads) {
      const int m = e / S, i = e - m * S;
      const int64_t col = col0 + m;
      s.xs[0][i * M + m] = col < a.B ? a.x0[col * S + i] : 0.0f;
    }
    __syncthreads();
    if (!a.skip_nn) run_three_nets<M>(a, s, s.xs[0]);
    const int64_t col = col0 + tid;
    if (tid < M && col < a.B) {
      const int rpb = a.rows_per_bc > 1 ? a.rows_per_bc : 1;
      const int64_t brow = col / rpb;
      const float t = a.t0 + (float)(col - brow * rpb) * a.dt;
      const bool skip = a.skip_nn;
      float bc[OPZ_BC_STRIDE];
#pragma unroll
      for (int q = 0; q < OPZ_BC_STRIDE; ++q) bc[q] = a.bc[brow * OPZ_BC_STRIDE + q];
      auto X = [&](int i) { return s.xs[0][i * M + tid]; };
      auto Y = [&](int net, int j) { return skip ? 0.0f : s.Y[(net * nz + j) * M + tid]; };
      auto K = [&](int i, float k) { if (a.out) a.out[col * S + i] = k; };
      if (a.aux) {
        auto FL = [&](int v, int f, float val) { a.aux[(col * 3 + v) * (nz + 1) + f] = val; };
        column_rhs(a.P, bc, t, X, Y, K, FL);
      } else {
        column_rhs(a.P, bc, t, X, Y, K, NoFlux());
      }
    }
    __syncthreads();
  }
}

// Ri on the Nz + 1 faces of every row (local_richardson of D^f x: zero gradients on the two boundary faces, no epsilon;
// training_postprocessing.jl:428-431) and the per-snapshot mean squared error of u, v, T against a target (loss_per_tstep,
// loss.jl:44-46): one thread per (row, face) / (row, variable).
__global__ void diag_kernel(DevParams P, const float* __restrict__ x, const float* __restrict__ target, int64_t rows,
                            float* __restrict__ ri, float* __restrict__ loss_t) {
  const int nz = P.nz, S = 3 * nz;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ri && i < rows * (nz + 1)) {
    const int64_t r = i / (nz + 1);
    const int f = (int)(i - r * (nz + 1));
    const float* xr = x + r * S;
    float gu = 0.0f, gv = 0.0f, gT = 0.0f;
    if (f > 0 && f < nz) {
      gu = (xr[f] - xr[f - 1]) * P.nzf;
      gv = (xr[nz + f] - xr[nz + f - 1]) * P.nzf;
      gT = (xr[2 * nz + f] - xr[2 * nz + f - 1]) * P.nzf;
    }
    const float sa = P.s_u * gu, sb = P.s_v * gv;
    ri[i] = (P.cBz * gT) / (sa * sa + sb * sb);   // IEEE: 0/0 on the boundary faces is NaN, exactly like the reference
  }
  if (loss_t && target && i < rows * 3) {
    const int64_t r = i / 3;
    const int v = (int)(i - r * 3);
    const float* a = x + r * S + v * nz;
    const float* b = target + r * S + v * nz;
    float acc = 0.0f;
    for (int k = 0; k < nz; ++k) { const float d = a[k] - b[k]; acc = fmaf(d, d, acc); }
    loss_t[i] = acc / (float)nz;
  }
}

// predict(V, model) (src/predict.jl:12-34): x[Nt][3nz] -> y[Nt][nz-1] for one net.
template <int M>
__global__ void __launch_bounds__(kThreads, 1) mlp_fp32_kernel(const __grid_constant__ Fp32Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem<M> s(smem_f, a.P.nz, a.max_hidden);
  const int tid = threadIdx.x;
  const int nz = a.P.nz, S = 3 * nz, NO = a.nd.sizes[a.nd.n_layers];
  const int64_t n_tiles = (a.B + M - 1) / M;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t col0 = tile * M;
    for (int e = tid; e < M * S; e += kThreads) {
      const int m = e / S, i = e - m * S;
      const int64_t col = col0 + m;
      s.xs[0][i * M + m] = col < a.B ? a.x0[col * S + i] : 0.0f;
    }
    __syncthreads();
    mlp_chain_fp32<M>(a.nd, a.w + (size_t)a.which_net * a.nd.P, s.xs[0], s.bufA, s.bufB, s.Y, s.stage);
    for (int e = tid; e < M * NO; e += kThreads) {
      const int m = e / NO, j = e - m * NO;
      const int64_t col = col0 + m;
      if (col < a.B) a.out[col * NO + j] = s.Y[j * M + m];
    }
    __syncthreads();
  }
}

// ---- host launchers ---------------------------------------------------------------------------
static int pick_tile(const opz_ctx* ctx, const opz_model* m) {
  if ((int)Smem<64>::bytes(m->nz, m->max_hidden) <= ctx->smem_optin) return 64;
  if ((int)Smem<32>::bytes(m->nz, m->max_hidden) <= ctx->smem_optin) return 32;
  if ((int)Smem<16>::bytes(m->nz, m->max_hidden) <= ctx->smem_optin) return 16;   // Nz = 128 states (BASELINE configs[4]) with nets up to 256 wide
  return 0;
}

template <class KernelT>
static int launch(opz_ctx* ctx, KernelT kern, size_t smem, int64_t n_tiles, const Fp32Args& a) {
  OPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = n_tiles < (int64_t)ctx->sm_count ? n_tiles : (int64_t)ctx->sm_count;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(a);
  OPZ_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return OPZ_OK;
}

static Fp32Args base_args(const opz_model* m, const DevParams& P) {
  Fp32Args a{};
  a.P = P;
  a.nd = m->nd;
  a.w = m->w_dev;
  a.max_hidden = m->max_hidden;
  return a;
}

int fp32_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c) {
  const int M = pick_tile(ctx, m);
  if (!M) { set_error("fp32 path: hidden width / nz too large for shared memory"); return OPZ_EUNSUP; }
  Fp32Args a = base_args(m, c.P);
  a.x0 = c.x0; a.bc = c.bc; a.out = c.out; a.B = c.B; a.scheme = c.scheme; a.dt = c.dt; a.t0 = c.t0;
  a.n_steps = c.n_steps; a.save_every = c.save_every; a.n_save = c.n_save;
  if (M == 64) return launch(ctx, rollout_fp32_kernel<64>, Smem<64>::bytes(m->nz, m->max_hidden), (c.B + 63) / 64, a);
  if (M == 32) return launch(ctx, rollout_fp32_kernel<32>, Smem<32>::bytes(m->nz, m->max_hidden), (c.B + 31) / 32, a);
  return launch(ctx, rollout_fp32_kernel<16>, Smem<16>::bytes(m->nz, m->max_hidden), (c.B + 15) / 16, a);
}

int fp32_rhs(opz_ctx* ctx, const opz_model* m, const DevParams& P, const float* x, const float* bc, float t,
             int64_t B, float* dxdt, float* fluxes) {
  const int M = pick_tile(ctx, m);
  if (!M) { set_error("fp32 path: hidden width / nz too large for shared memory"); return OPZ_EUNSUP; }
  Fp32Args a = base_args(m, P);
  a.x0 = x; a.bc = bc; a.out = dxdt; a.aux = fluxes; a.B = B; a.t0 = t;
  if (M == 64) return launch(ctx, rhs_fp32_kernel<64>, Smem<64>::bytes(m->nz, m->max_hidden), (B + 63) / 64, a);
  if (M == 32) return launch(ctx, rhs_fp32_kernel<32>, Smem<32>::bytes(m->nz, m->max_hidden), (B + 31) / 32, a);
  return launch(ctx, rhs_fp32_kernel<16>, Smem<16>::bytes(m->nz, m->max_hidden), (B + 15) / 16, a);
}

int fp32_diagnostics(opz_ctx* ctx, const opz_model* m, const DevParams& P, const float* sol, const float* bc, int64_t B, int n_save,
                     float t0, float dt_save, const float* target, float* fluxes, float* fluxes_mpp, float* ri, float* loss_t) {
  const int M = pick_tile(ctx, m);
  if (!M) { set_error("fp32 path: hidden width / nz too large for shared memory"); return OPZ_EUNSUP; }
  const int64_t rows = B * n_save;
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? fluxes : fluxes_mpp;
    if (!dst) continue;
    Fp32Args a = base_args(m, P);
    a.x0 = sol; a.bc = bc; a.out = nullptr; a.aux = dst; a.B = rows; a.t0 = t0; a.dt = dt_save;
    a.rows_per_bc = n_save; a.skip_nn = pass;
    int rc;
    if (M == 64) rc = launch(ctx, rhs_fp32_kernel<64>, Smem<64>::bytes(m->nz, m->max_hidden), (rows + 63) / 64, a);
    else if (M == 32) rc = launch(ctx, rhs_fp32_kernel<32>, Smem<32>::bytes(m->nz, m->max_hidden), (rows + 31) / 32, a);
    else rc = launch(ctx, rhs_fp32_kernel<16>, Smem<16>::bytes(m->nz, m->max_hidden), (rows + 15) / 16, a);
    if (rc) return rc;
  }
  if (ri || (loss_t && target)) {
    const int64_t n = std::max<int64_t>(rows * (m->nz + 1), rows * 3);
    diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(P, sol, target, rows, ri, loss_t);
    OPZ_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  return OPZ_OK;
}

int fp32_mlp(opz_ctx* ctx, const opz_model* m, int which_net, const float* x, int64_t Nt, float* y) {
  const int M = pick_tile(ctx, m);
  if (!M) { set_error("fp32 path: hidden width / nz too large for shared memory"); return OPZ_EUNSUP; }
  DevParams P{};
  P.nz = m->nz;
  Fp32Args a = base_args(m, P);
  a.x0 = x; a.out = y; a.B = Nt; a.which_net = which_net;
  if (M == 64) return launch(ctx, mlp_fp32_kernel<64>, Smem<64>::bytes(m->nz, m->max_hidden), (Nt + 63) / 64, a);
  if (M == 32) return launch(ctx, mlp_fp32_kernel<32>, Smem<32>::bytes(m->nz, m->max_hidden), (Nt + 31) / 32, a);
  return launch(ctx, mlp_fp32_kernel<16>, Smem<16>::bytes(m->nz, m->max_hidden), (Nt + 15) / 16, a);
}

}  // namespace opz
