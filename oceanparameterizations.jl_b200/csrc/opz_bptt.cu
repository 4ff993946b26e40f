// opz_bptt.cu — loss of the fixed-step NDE solve and its exact discrete adjoint (BPTT) w.r.t. the
// NN weights, FP32.
//
// Replaces loss_NDE / loss_gradient_NDE + Zygote.gradient through
// solve(...; sensealg = InterpolatingAdjoint(autojacvec = ZygoteVJP()))
// (wind_mixing/src/NDE_training.jl:290-333; loss: wind_mixing/src/loss.jl:1-42).
//
// The training batch of the reference is 18 columns (train_NDE_args.jl:39-60), far too few to fill
// 148 SMs column-by-column, and the rollout is strictly sequential in time.  So this path is laid out
// the other way round from the forward ensemble kernels: ONE right-hand side is spread over the whole
// GPU, layer by layer (every CTA owns a slice of a layer's units and all B columns), and the
// launch-bound chain of small kernels is captured once into CUDA graphs that are replayed over the
// time steps; a device-side cursor carries the step index so that the graphs never need re-parametrising.
//
//   forward sweep     x_n -> x_{n+1}: per stage  [hidden layers ...] -> [last layer + physics + RK update]
//                     every stage's input, hidden activations h_l and act'(z_l) are RECORDED in HBM
//   backward sweep    per stage  [physics VJP + last layer^T] -> [hidden layers^T ...] -> [first layer^T + RK adjoint]
//                     act'(z_l) records are overwritten in place by dz_l
//   weight gradient   dW_l = sum over all recorded right-hand sides of h_{l-1}^T dz_l : ONE large
//                     reduction GEMM per layer over (columns x stages x steps) rows, instead of a
//                     read-modify-write of the 2.5 MB gradient per right-hand side.
// If the records of the whole rollout do not fit the memory budget, the rollout is cut into time
// segments (state checkpoints x_n are kept for every step): the forward sweep runs without records,
// then each segment is recomputed with records right before its backward sweep.
//
// The recurrences are the ones restated and checked against autograd in oracle/nde_adjoint.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "opz_internal.h"
#include "opz_faces.cuh"

namespace opz {
namespace bptt {

constexpr int kCT = 20;        // columns per GEMM column tile (18 training cases + padding)
constexpr int kNT = 16;        // units per CTA
constexpr int kKS = 16;        // split of the reduction dimension inside a CTA
constexpr int kGemmThreads = kNT * kKS;
constexpr int kMaxStages = 4;
constexpr int kWR = 32;                  // weight registers per thread and chunk
constexpr int kFW = 40;                  // last-layer weight registers per thread (final_fwd_kernel)
constexpr int kChunk = kWR * kKS;      // inputs per staged chunk (512)

struct Cursor {
  int n;        // base time step of the graph launch in flight
  int seg_n0;   // first step of the recorded segment
};

struct Args {
  DevParams P;
  NetDesc nd;
  const float* w;        // [3P] Flux order
  const float* wT;       // transposed copies of layers 0..L-2, per net
  int64_t wT_off[kMaxLayers];   // offset of layer l inside one net's block
  int64_t wT_net;        // floats per net
  const float* bc;       // [B][8]
  const float* target;   // [B][n_save][3nz]
  float* XN;             // [n_steps+1][B][3nz] state checkpoints
  float* X;              // [Rcap][B][3nz] stage inputs
  float* H[kMaxLayers];  // [3][Rcap*B][n_l] hidden activations
  float* G[kMaxLayers];  // [3][Rcap*B][n_l] act'(z_l), then dz_l
  float* DY;             // [3][Rcap*B][nout]
  float* KACC;           // [B][3nz] Runge-Kutta accumulator (forward)
  float* LAM;            // [B][3nz] dL/dx_{n+1} -> dL/dx_n
  float* KBAR;           // [B][3nz] cotangent of the current stage's tendency
  float* XSUM;           // [B][3nz] sum of the stage-input cotangents of the step
  float* XPHYS;          // [B][3nz] physics part of the current stage-input cotangent
  float* XSTAR;          // OPZ_IMEX_EULER: [n_steps][B][3nz] the explicit half-step results x* (null otherwise)
  uint32_t imex_flags;   // OPZ_IMEX_EULER: the caller's full flag set (P.flags then holds the explicit part's: no diffusion)
  double* loss_part;     // [B][6]
  float* loss_out;       // [7]
  float* grad;           // [3P]
  float* gphys;          // [5] or null: gradient w.r.t. the modified Pacanowski-Philander parameters (nu0, nu_m, dRi, Ric, Pr)
  const Cursor* cur;
  int64_t B, B_total;
  int64_t rows_cap;      // Rcap * B
  int S3, nout, L;
  int scheme, stages;
  int n_steps, save_every, n_save;
  int record;            // 1: record index = (n - seg_n0)*stages + s ; 0: ring of `stages` records
  int rcap;              // Rcap
  float h, t0;
  float rk_b[kMaxStages], rk_a[kMaxStages], rk_c[kMaxStages];
  float loss_w[6];
};

__device__ __forceinline__ int rec_index(const Args& a, int n, int s) {
  return a.record ? (n - a.cur->seg_n0) * a.stages + s : s;
}

// ---- cursor -----------------------------------------------------------------------------------------
__global__ void set_cursor_kernel(Cursor* c, int n, int seg_n0) { c->n = n; c->seg_n0 = seg_n0; }
__global__ void bump_cursor_kernel(Cursor* c, int dn) { c->n += dn; }

// ---- layer-parallel skinny GEMM ------------------------------------------------------------------------
// out[c][j] = sum_i A[c][i] * M[i*J + j] for a tile of kCT columns x kNT outputs, reduction split kKS ways.
enum { MODE_FWD = 0, MODE_BWD_HID = 1, MODE_BWD_IN = 2 };

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads) skinny_kernel(const __grid_constant__ Args a, int layer, int stage, int dn) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int nl = tid % kNT, ks = tid / kNT;
  const int net0 = (MODE == MODE_BWD_IN) ? 0 : blockIdx.z;
  const int nsum = (MODE == MODE_BWD_IN) ? 3 : 1;
  int I, J;
  if (MODE == MODE_FWD) { I = a.nd.sizes[layer]; J = a.nd.sizes[layer + 1]; }
  else if (MODE == MODE_BWD_HID) { I = a.nd.sizes[layer + 1]; J = a.nd.sizes[layer]; }   // dz_layer -> dz_{layer-1}
  else { I = a.nd.sizes[1]; J = a.S3; }
  const int Ic_max = min(I, kChunk);
  float* As = sm;                          // [min(I, kChunk)][kCT]
  float* red = sm + (size_t)Ic_max * kCT;  // [kKS][kCT][kNT]
  const int j = blockIdx.x * kNT + nl;

  float acc[kCT];
#pragma unroll
  for (int c = 0; c < kCT; ++c) acc[c] = 0.0f;

  // The weights do not depend on the step: their loads are issued first so that their L2 latency
  // overlaps the cursor read and the staging of the activations.
  float wreg[kWR];
  auto weights_of = [&](int net) -> const float* {
    if (MODE == MODE_FWD) return a.w + (int64_t)net * a.nd.P + a.nd.w_off[layer];
    return a.wT + (int64_t)net * a.wT_net + a.wT_off[MODE == MODE_BWD_HID ? layer : 0];
  };
  auto fetch_w = [&](int net, int i0) {
    const float* M = weights_of(net);
#pragma unroll
    for (int r = 0; r < kWR; ++r) {
      const int i = i0 + ks + r * kKS;
      wreg[r] = (i < I && j < J) ? __ldg(M + (size_t)i * J + j) : 0.0f;
    }
  };
  fetch_w(net0, 0);

  const int n = a.cur->n + dn;
  const int q = rec_index(a, n, stage);
  const int64_t col0 = (int64_t)blockIdx.y * kCT;
  const int ncols = (int)min((int64_t)kCT, a.B - col0);
  const int64_t r0 = (int64_t)q * a.B + col0;

  bool first = true;
  for (int ns = 0; ns < nsum; ++ns) {
    const int net = net0 + ns;
    const float* A;
    int lda;
    if (MODE == MODE_FWD) {
      if (layer == 0) { A = a.X + r0 * a.S3; lda = a.S3; }
      else { lda = a.nd.sizes[layer]; A = a.H[layer - 1] + ((int64_t)net * a.rows_cap + r0) * lda; }
    } else if (MODE == MODE_BWD_HID) {
      lda = I;
      A = a.G[layer] + ((int64_t)net * a.rows_cap + r0) * lda;
    } else {
      lda = I;
      A = a.G[0] + ((int64_t)net * a.rows_cap + r0) * lda;
    }
    for (int i0 = 0; i0 < I; i0 += kChunk) {
      const int Ic = min(kChunk, I - i0);
      if (!first) {
        __syncthreads();  // the previous chunk's As has been consumed
        fetch_w(net, i0);
      }
      first = false;
#pragma unroll 8
      for (int e = tid; e < Ic * kCT; e += kGemmThreads) {
        const int c = e / Ic, i = e - c * Ic;
        As[i * kCT + c] = c < ncols ? A[(int64_t)c * lda + i0 + i] : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kWR; ++r) {
        const int il = ks + r * kKS;
        if (il < Ic) {
          const float wv = wreg[r];
          const float4* ap = reinterpret_cast<const float4*>(As + il * kCT);
#pragma unroll
          for (int v = 0; v < kCT / 4; ++v) {
            const float4 a4 = ap[v];
            acc[4 * v + 0] = fmaf(wv, a4.x, acc[4 * v + 0]);
            acc[4 * v + 1] = fmaf(wv, a4.y, acc[4 * v + 1]);
            acc[4 * v + 2] = fmaf(wv, a4.z, acc[4 * v + 2]);
            acc[4 * v + 3] = fmaf(wv, a4.w, acc[4 * v + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kCT; ++c) red[(ks * kCT + c) * kNT + nl] = acc[c];
  __syncthreads();
  for (int e = tid; e < kCT * kNT; e += kGemmThreads) {
    const int c = e / kNT, jl = e - c * kNT;
    const int jj = blockIdx.x * kNT + jl;
    if (c >= ncols || jj >= J) continue;
    float sum = 0.0f;
#pragma unroll
    for (int s2 = 0; s2 < kKS; ++s2) sum += red[(s2 * kCT + c) * kNT + jl];
    const int64_t row = r0 + c;
    if (MODE == MODE_FWD) {
      const int net = net0;
      const float z = sum + __ldg(a.w + (int64_t)net * a.nd.P + a.nd.b_off[layer] + jj);
      const int64_t o = ((int64_t)net * a.rows_cap + row) * J + jj;
      a.H[layer][o] = act_dyn(z, a.nd.act);
      a.G[layer][o] = act_grad_dyn(z, a.nd.act);
    } else if (MODE == MODE_BWD_HID) {
      const int64_t o = ((int64_t)net0 * a.rows_cap + row) * J + jj;
      a.G[layer - 1][o] = sum * a.G[layer - 1][o];
    } else {
      // stage-input cotangent complete: Runge-Kutta adjoint bookkeeping (oracle/nde_adjoint.py)
      const int64_t o = (col0 + c) * a.S3 + jj;
      const float xbar = sum + a.XPHYS[o];
      const float lam = a.LAM[o];
      const float xsum = (stage == a.stages - 1 ? 0.0f : a.XSUM[o]) + xbar;
      if (stage > 0) {
        a.KBAR[o] = a.rk_b[stage - 1] * lam + a.rk_a[stage - 1] * xbar;
        a.XSUM[o] = xsum;
      } else {
        float ln = lam + xsum;
        if (n % a.save_every == 0) {
          // dL/dx_n of the profile loss at this save point (loss.jl:1-9; NDE_training.jl:290-317)
          const int nz = a.P.nz, v = jj / nz, cz = jj - v * nz;
          const int snap = n / a.save_every;
          const float* xs = a.XN + ((int64_t)n * a.B + col0 + c) * a.S3 + v * nz;
          const float* tg = a.target + (((col0 + c) * a.n_save + snap) * (int64_t)a.S3) + v * nz;
          const float inv_p = 1.0f / ((float)nz * (float)a.n_save * (float)a.B_total);
          const float inv_g = 1.0f / ((float)(nz + 1) * (float)a.n_save * (float)a.B_total);
          const float d0 = xs[cz] - tg[cz];
          float add = a.loss_w[v] * 2.0f * d0 * inv_p;
          const float sc = a.loss_w[3 + v] * 2.0f * a.P.nzf * inv_g;
          if (sc != 0.0f) {
            if (cz > 0) add += sc * ((d0 - (xs[cz - 1] - tg[cz - 1])) * a.P.nzf);
            if (cz < nz - 1) add -= sc * (((xs[cz + 1] - tg[cz + 1]) - d0) * a.P.nzf);
          }
          ln += add;
        }
        a.LAM[o] = ln;
        a.KBAR[o] = a.rk_b[a.stages - 1] * ln;
      }
    }
  }
}

// ---- last Dense layer + physics + Runge-Kutta update: one CTA per column -----------------------------------
__global__ void __launch_bounds__(1024) final_fwd_kernel(const __grid_constant__ Args a, int stage, int dn) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nz = a.P.nz, S3 = a.S3, nout = a.nout, L = a.L;
  const int nh = a.nd.sizes[L - 1];          // width of the last hidden layer
  const int OP = (3 * nout + 31) / 32 * 32;  // padded number of outputs
  const int KSL = nthr / OP;                 // reduction slices
  float* hs = sm;                            // [3][nh]
  float* ypart = hs + 3 * nh;                // [KSL][OP]
  float* xs = ypart + KSL * OP;              // [S3]
  float* Fs = xs + S3;                       // [3][nz+1]
  const int64_t col = blockIdx.x;
  // last-layer weights first (independent of the step), then the cursor and the activations
  const int o = tid % OP, ksl = tid / OP;
  const bool live = ksl < KSL && o < 3 * nout;
  const int onet = live ? o / nout : 0, oj = live ? o - onet * nout : 0;
  const float* Wl = a.w + (int64_t)onet * a.nd.P + a.nd.w_off[L - 1] + oj;
  float wreg[kFW];
#pragma unroll
  for (int r = 0; r < kFW; ++r) {
    const int k = ksl + r * KSL;
    wreg[r] = (live && k < nh) ? __ldg(Wl + (size_t)k * nout) : 0.0f;
  }
  const int n = a.cur->n + dn;
  const int q = rec_index(a, n, stage);
  const int64_t row = (int64_t)q * a.B + col;

#pragma unroll 4
  for (int e = tid; e < 3 * nh; e += nthr) {
    const int net = e / nh, k = e - net * nh;
    hs[e] = a.H[L - 2][((int64_t)net * a.rows_cap + row) * nh + k];
  }
  for (int i = tid; i < S3; i += nthr) xs[i] = a.X[row * S3 + i];
  __syncthreads();
  {
    float part = 0.0f;
    const float* hh = hs + onet * nh;
#pragma unroll
    for (int r = 0; r < kFW; ++r) {
      const int k = ksl + r * KSL;
      if (k < nh) part = fmaf(hh[k], wreg[r], part);
    }
    for (int k = ksl + kFW * KSL; live && k < nh; k += KSL) part = fmaf(hh[k], __ldg(Wl + (size_t)k * nout), part);
    if (ksl < KSL) ypart[ksl * OP + o] = part;
  }
  __syncthreads();
  if (tid < 3 * nout) {
    const int net = tid / nout, j = tid - net * nout;
    float y = 0.0f;
    for (int s2 = 0; s2 < KSL; ++s2) y += ypart[s2 * OP + tid];
    y += __ldg(a.w + (int64_t)net * a.nd.P + a.nd.b_off[L - 1] + j);
    ypart[tid] = y;  // slice 0 now holds the outputs
  }
  __syncthreads();
  const float* bc = a.bc + col * OPZ_BC_STRIDE;
  const float tn = a.t0 + (float)n * a.h;
  const float ts = tn + a.rk_c[stage] * a.h;
  const DevParams& P = a.P;
  if (tid <= nz) {  // one thread per face
    const int f = tid;
    float Fu, Fv, FT;
    if (f == 0) { Fu = bc[0] - P.off[0]; Fv = bc[2] - P.off[1]; FT = bc[4] - P.off[2]; }
    else if (f == nz) { Fu = bc[1] - P.off[0]; Fv = bc[3] - P.off[1]; FT = wT_top_value(P, bc, ts); }
    else if (P.flags & (OPZ_FLAG_SMOOTH_NN | OPZ_FLAG_SMOOTH_RI)) {   // box-filtered NN outputs / Richardson numbers (opz_faces.cuh)
      float F3[3];
      faces::forward_filtered(P, xs, ypart, 1, 0, f, F3);
      Fu = F3[0]; Fv = F3[1]; FT = F3[2];
    } else {
      Fu = ypart[f - 1]; Fv = ypart[nout + f - 1]; FT = ypart[2 * nout + f - 1];
      if (P.flags & OPZ_FLAG_MPP) {
        const float gu = (xs[f] - xs[f - 1]) * P.nzf;
        const float gv = (xs[nz + f] - xs[nz + f - 1]) * P.nzf;
        const float gT = (xs[2 * nz + f] - xs[2 * nz + f - 1]) * P.nzf;
        const float Ri = richardson(P, gu, gv, gT);
        const float nu = nu_of_ri(P, Ri);
        const float nuT = nuT_of(P, nu, gu, gT);
        Fu -= P.c_u * nu * gu;
        Fv -= P.c_v * nu * gv;
        FT -= P.c_T * nuT * gT;
      }
    }
    Fs[f] = Fu; Fs[(nz + 1) + f] = Fv; Fs[2 * (nz + 1) + f] = FT;
  }
  __syncthreads();
  if (tid < S3) {
    const int v = tid / nz, c = tid - v * nz;
    const float* F = Fs + v * (nz + 1);
    float k;
    if (v == 0) k = P.A_u * ((F[c + 1] - F[c]) * P.nzf) + P.cor_u * (P.s_v * xs[nz + c] + P.mu_v);
    else if (v == 1) k = P.A_v * ((F[c + 1] - F[c]) * P.nzf) - P.cor_v * (P.s_u * xs[c] + P.mu_u);
    else k = P.A_T * ((F[c + 1] - F[c]) * P.nzf);
    const int64_t o = col * S3 + tid;
    const float xn = a.XN[((int64_t)n * a.B + col) * S3 + tid];
    const float hh = a.h;
    float xnext;
    bool last = (stage == a.stages - 1);
    if (a.scheme == OPZ_RK4) {
      if (stage == 0) { a.KACC[o] = k; xnext = xn + (hh * 0.5f) * k; }
      else if (stage == 1) { a.KACC[o] = a.KACC[o] + 2.0f * k; xnext = xn + (hh * 0.5f) * k; }
      else if (stage == 2) { a.KACC[o] = a.KACC[o] + 2.0f * k; xnext = xn + hh * k; }
      else xnext = xn + (hh / 6.0f) * (a.KACC[o] + k);
    } else if (a.scheme == OPZ_RK2) {
      if (stage == 0) { a.KACC[o] = k; xnext = xn + hh * k; }
      else xnext = xn + (hh * 0.5f) * (a.KACC[o] + k);
    } else {
      xnext = xn + hh * k;
    }
    if (last) a.XN[((int64_t)(n + 1) * a.B + col) * S3 + tid] = xnext;
    int qn;
    if (a.record) qn = q + 1 < a.rcap ? q + 1 : -1;
    else qn = last ? 0 : stage + 1;
    if (qn >= 0) a.X[((int64_t)qn * a.B + col) * S3 + tid] = xnext;
  }
}

// ---- physics VJP + transpose of the last Dense layer: one CTA per column --------------------------------
// ---- cotangents of the five modified Pacanowski-Philander parameters from one interior face ----------------------
// (diffusivity_parameter_optimisation.jl:1-33: p = (nu0, nu_m, dRi, Ric, Pr) of nu = nu0 + nu_m (1 - tanh((Ri - Ric)/dRi))/2,
//  nu_T = nu / Pr).  nubar = cotangent of nu from all three fluxes, FbT = cotangent of the total T flux, th = tanh(...).
__device__ __forceinline__ void mpp_param_cotangents(const DevParams& P, float nubar, float FbT, float th, float Ri, float nu,
                                                     float gT, bool keep, bool ok, float (&gp)[5]) {
  const float sech2 = 1.0f - th * th;
  gp[0] += nubar;
  gp[1] += nubar * ((1.0f - th) / 2.0f);
  if (ok) {
    gp[2] += nubar * (P.nu_m / 2.0f * sech2 * (Ri - P.Ric) / (P.dRi * P.dRi));
    gp[3] += nubar * (P.nu_m / (2.0f * P.dRi) * sech2);
  }
  if (keep) gp[4] += FbT * (P.c_T * nu * gT / (P.Pr * P.Pr));
}
// sum gp over the warp; lane 0 adds the result to out[5]
__device__ __forceinline__ void mpp_param_flush(float (&gp)[5], float* out) {
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    float v = gp[i];
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
    if ((threadIdx.x & 31) == 0 && v != 0.0f) atomicAdd(out + i, v);
  }
}

__global__ void __launch_bounds__(1024) first_bwd_kernel(const __grid_constant__ Args a, int stage, int dn) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nz = a.P.nz, S3 = a.S3, nout = a.nout, L = a.L;
  const int nh = a.nd.sizes[L - 1];
  float* xs = sm;              // [S3]
  float* kb = xs + S3;         // [S3]
  float* dy = kb + S3;         // [3][nout]
  float* gb = dy + 3 * nout;   // [3][nz+1] cotangents of the face gradients
  const int64_t col = blockIdx.x;
  const int n = a.cur->n + dn;
  const int q = rec_index(a, n, stage);
  const int64_t row = (int64_t)q * a.B + col;
  const DevParams& P = a.P;
  for (int i = tid; i < S3; i += nthr) {
    xs[i] = a.X[row * S3 + i];
    kb[i] = a.KBAR[col * S3 + i];
  }
  __syncthreads();
  float gp[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (tid <= nz) {
    const int f = tid;
    float gbu = 0.0f, gbv = 0.0f, gbT = 0.0f;
    if (f > 0 && f < nz && (P.flags & (OPZ_FLAG_SMOOTH_NN | OPZ_FLAG_SMOOTH_RI))) {
      float dy3[3], gb3[3];
      faces::backward_filtered(P, xs, kb, 1, 0, f, dy3, gb3, a.gphys ? gp : nullptr);
      dy[f - 1] = dy3[0]; dy[nout + f - 1] = dy3[1]; dy[2 * nout + f - 1] = dy3[2];
      a.DY[((int64_t)0 * a.rows_cap + row) * nout + f - 1] = dy3[0];
      a.DY[((int64_t)1 * a.rows_cap + row) * nout + f - 1] = dy3[1];
      a.DY[((int64_t)2 * a.rows_cap + row) * nout + f - 1] = dy3[2];
      gbu = gb3[0]; gbv = gb3[1]; gbT = gb3[2];
    } else if (f > 0 && f < nz) {
      const float Fbu = P.A_u * P.nzf * (kb[f - 1] - kb[f]);
      const float Fbv = P.A_v * P.nzf * (kb[nz + f - 1] - kb[nz + f]);
      const float FbT = P.A_T * P.nzf * (kb[2 * nz + f - 1] - kb[2 * nz + f]);
      dy[f - 1] = Fbu; dy[nout + f - 1] = Fbv; dy[2 * nout + f - 1] = FbT;
      a.DY[((int64_t)0 * a.rows_cap + row) * nout + f - 1] = Fbu;
      a.DY[((int64_t)1 * a.rows_cap + row) * nout + f - 1] = Fbv;
      a.DY[((int64_t)2 * a.rows_cap + row) * nout + f - 1] = FbT;
      if (P.flags & OPZ_FLAG_MPP) {
        const float gu = (xs[f] - xs[f - 1]) * P.nzf;
        const float gv = (xs[nz + f] - xs[nz + f - 1]) * P.nzf;
        const float gT = (xs[2 * nz + f] - xs[2 * nz + f - 1]) * P.nzf;
        const float ea = P.s_u * (gu + P.eps), eb = P.s_v * (gv + P.eps);
        const float S2 = ea * ea + eb * eb;
        const float Ri = P.cBz * (gT + P.eps) / S2;
        const float th = tanhf((Ri - P.Ric) / P.dRi);
        const float nu = P.nu0 + P.nu_m * ((1.0f - th) / 2.0f);
        bool keep = true;
        float nuT = nu / P.Pr;
        if (P.flags & OPZ_FLAG_CA) {
          const float sel = (P.flags & OPZ_FLAG_CA_SELECT_DUDZ) ? gu : gT;
          keep = sel > 0.0f;
          if (!keep) nuT = P.kappa;
        }
        const float nubar = -P.c_u * gu * Fbu - P.c_v * gv * Fbv - (keep ? P.c_T * gT * FbT / P.Pr : 0.0f);
        const float Ribar = nubar * (-P.nu_m / (2.0f * P.dRi) * (1.0f - th * th));
        const float dRi_dgT = P.cBz / S2;
        const bool ok = isfinite(Ri) && isfinite(dRi_dgT);
        gbu = -P.c_u * nu * Fbu + (ok ? Ribar * (-Ri * 2.0f * P.s_u * ea / S2) : 0.0f);
        gbv = -P.c_v * nu * Fbv + (ok ? Ribar * (-Ri * 2.0f * P.s_v * eb / S2) : 0.0f);
        gbT = -P.c_T * nuT * FbT + (ok ? Ribar * dRi_dgT : 0.0f);
        if (a.gphys) mpp_param_cotangents(P, nubar, FbT, th, Ri, nu, gT, keep, ok, gp);
      }
    }
    gb[f] = gbu; gb[(nz + 1) + f] = gbv; gb[2 * (nz + 1) + f] = gbT;
  }
  if (a.gphys && tid < ((nz + 1 + 31) / 32) * 32) mpp_param_flush(gp, a.gphys);   // whole warps that hold faces
  __syncthreads();
  if (tid < S3) {
    const int v = tid / nz, c = tid - v * nz;
    const float* g = gb + v * (nz + 1);
    float xb = P.nzf * (g[c] - g[c + 1]);  // cell c is the upper cell of face c and the lower cell of face c+1
    if (v == 0) xb -= P.cor_v * P.s_u * kb[nz + c];
    else if (v == 1) xb += P.cor_u * P.s_v * kb[c];
    a.XPHYS[col * S3 + tid] = xb;
  }
  // dz_{L-2} = (W_L dy) * act'(z_{L-2})   (in place over the act' record): one warp per weight row,
  // lanes along the nout contiguous outputs, shuffle reduction
  {
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    for (int e0 = warp; e0 < 3 * nh; e0 += nwarp * 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e0 + u * nwarp;
        s[u] = 0.0f;
        if (e < 3 * nh) {
          const int net = e / nh, k = e - net * nh;
          const float* W = a.w + (int64_t)net * a.nd.P + a.nd.w_off[L - 1] + (size_t)k * nout;
          for (int j = lane; j < nout; j += 32) s[u] = fmaf(__ldg(W + j), dy[net * nout + j], s[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], sh);
        const int e = e0 + u * nwarp;
        if (lane == 0 && e < 3 * nh) {
          const int net = e / nh, k = e - net * nh;
          const int64_t o = ((int64_t)net * a.rows_cap + row) * nh + k;
          a.G[L - 2][o] = s[u] * a.G[L - 2][o];
        }
      }
    }
  }
}

// ---- split implicit step (OPZ_IMEX_EULER): x* = x_n + h k_ex(x_n), then (I + L(nu(x*))) x_{n+1} = x* per variable ------
// (NDE_oceananigans.jl:61-101; forward twin: rollout_fp32_kernel, oracle step_imex).  L is the tridiagonal diffusion operator
// with D_f = coef nu_f(x*) on the interior faces: symmetric, so the adjoint solve is the same Thomas solve.
//   forward   x* (written by final_fwd_kernel as the Euler result) -> XSTAR[n]; solve -> XN[n+1] and the next stage input
//   backward  lam = dL/dx_{n+1}:  mu_v = M_v^{-1} lam_v;  dL/dD_f = -(mu_f - mu_{f-1}) (x_f - x_{f-1})  (x = x_{n+1});
//             nubar_f = coef (dD^u + dD^v + [keep] dD^T / Pr);  Ribar -> gradient cotangents -> x*bar = mu + D^f-transpose terms;
//             LAM <- x*bar, KBAR <- h x*bar, after which the explicit Euler adjoint of the step runs unchanged.
// One thread per column (the training batch is 18 columns and this is a serial recurrence over the Nz cells).
struct ImexFaces {
  float ri_raw[kMaxNz + 1], ri_s[kMaxNz + 1], th[kMaxNz + 1], nu[kMaxNz + 1];
  bool keep[kMaxNz + 1];
};
__device__ void imex_diffusivities(const DevParams& P, const float* xs, ImexFaces& F) {
  const int nz = P.nz;
  for (int f = 0; f <= nz; ++f) {
    const bool in = f > 0 && f < nz;
    const float gu = in ? (xs[f] - xs[f - 1]) * P.nzf : 0.0f;
    const float gv = in ? (xs[nz + f] - xs[nz + f - 1]) * P.nzf : 0.0f;
    const float gT = in ? (xs[2 * nz + f] - xs[2 * nz + f - 1]) * P.nzf : 0.0f;
    F.ri_raw[f] = richardson(P, gu, gv, gT);
    F.keep[f] = !(P.flags & OPZ_FLAG_CA) || (((P.flags & OPZ_FLAG_CA_SELECT_DUDZ) ? gu : gT) > 0.0f);
  }
  for (int f = 1; f < nz; ++f) {
    F.ri_s[f] = (P.flags & OPZ_FLAG_SMOOTH_RI) ? (F.ri_raw[f - 1] + F.ri_raw[f] + F.ri_raw[f + 1]) * (1.0f / 3.0f) : F.ri_raw[f];
    F.th[f] = tanhf((F.ri_s[f] - P.Ric) / P.dRi);
    F.nu[f] = P.nu0 + P.nu_m * ((1.0f - F.th[f]) / 2.0f);
  }
}
// (I + L) sol = rhs with D[f] on faces 0..nz (D[0] = D[nz] = 0); cp is scratch
__device__ void imex_thomas(const float* D, const float* rhs, float* sol, float* cp, int nz) {
  float den = 1.0f + D[0] + D[1];
  cp[0] = -D[1] / den;
  sol[0] = rhs[0] / den;
  for (int c = 1; c < nz; ++c) {
    den = (1.0f + D[c] + D[c + 1]) + D[c] * cp[c - 1];
    cp[c] = -D[c + 1] / den;
    sol[c] = (rhs[c] + D[c] * sol[c - 1]) / den;
  }
  for (int c = nz - 2; c >= 0; --c) sol[c] = sol[c] - cp[c] * sol[c + 1];
}
__device__ __forceinline__ float imex_coef(const DevParams& P, float h, int v) {   // dt nu / dz^2 per unit nu, as in rollout_fp32_kernel
  return v == 0 ? h * (-P.A_u) * P.c_u * P.nzf * P.nzf : (v == 1 ? h * (-P.A_v) * P.c_v * P.nzf * P.nzf : h * (-P.A_T) * P.c_T * P.nzf * P.nzf);
}

__global__ void __launch_bounds__(32) imex_fwd_kernel(const __grid_constant__ Args a, int dn) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.B) return;
  DevParams P = a.P;
  P.flags = a.imex_flags;
  const int nz = P.nz, S3 = a.S3;
  const int n = a.cur->n + dn;
  float* xnew = a.XN + ((int64_t)(n + 1) * a.B + col) * S3;      // holds x* on entry
  float* xstar = a.XSTAR + ((int64_t)n * a.B + col) * S3;
  float xs[3 * kMaxNz];
  for (int i = 0; i < S3; ++i) { xs[i] = xnew[i]; xstar[i] = xs[i]; }
  ImexFaces F;
  imex_diffusivities(P, xs, F);
  float D[kMaxNz + 1], cp[kMaxNz], sol[kMaxNz];
  const int q = rec_index(a, n, 0);
  int qn;
  if (a.record) qn = q + 1 < a.rcap ? q + 1 : -1;
  else qn = 0;
  for (int v = 0; v < 3; ++v) {
    const float coef = imex_coef(P, a.h, v);
    D[0] = 0.0f; D[nz] = 0.0f;
    for (int f = 1; f < nz; ++f) D[f] = coef * (v < 2 ? F.nu[f] : (F.keep[f] ? F.nu[f] / P.Pr : P.kappa));
    imex_thomas(D, xs + v * nz, sol, cp, nz);
    for (int c = 0; c < nz; ++c) {
      xnew[v * nz + c] = sol[c];
      if (qn >= 0) a.X[((int64_t)qn * a.B + col) * S3 + v * nz + c] = sol[c];
    }
  }
}

__global__ void __launch_bounds__(32) imex_bwd_kernel(const __grid_constant__ Args a, int dn) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float gp[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  if (col < a.B) {
    DevParams P = a.P;
    P.flags = a.imex_flags;
    const int nz = P.nz, S3 = a.S3;
    const int n = a.cur->n + dn;
    const float* xnew = a.XN + ((int64_t)(n + 1) * a.B + col) * S3;
    const float* xstar = a.XSTAR + ((int64_t)n * a.B + col) * S3;
    float xs[3 * kMaxNz];
    for (int i = 0; i < S3; ++i) xs[i] = xstar[i];
    ImexFaces F;
    imex_diffusivities(P, xs, F);
    float D[kMaxNz + 1], cp[kMaxNz], mu[kMaxNz], lamv[kMaxNz], nubar[kMaxNz + 1], xbar[3 * kMaxNz];
    for (int f = 0; f <= nz; ++f) nubar[f] = 0.0f;
    for (int v = 0; v < 3; ++v) {
      const float coef = imex_coef(P, a.h, v);
      D[0] = 0.0f; D[nz] = 0.0f;
      for (int f = 1; f < nz; ++f) D[f] = coef * (v < 2 ? F.nu[f] : (F.keep[f] ? F.nu[f] / P.Pr : P.kappa));
      for (int c = 0; c < nz; ++c) lamv[c] = a.LAM[col * S3 + v * nz + c];
      imex_thomas(D, lamv, mu, cp, nz);
      for (int c = 0; c < nz; ++c) xbar[v * nz + c] = mu[c];
      for (int f = 1; f < nz; ++f) {
        const float dD = -(mu[f] - mu[f - 1]) * (xnew[v * nz + f] - xnew[v * nz + f - 1]);
        if (v < 2) nubar[f] += coef * dD;
        else if (F.keep[f]) {
          nubar[f] += coef * dD / P.Pr;
          gp[4] += coef * dD * (-F.nu[f] / (P.Pr * P.Pr));
        }
      }
    }
    // nu -> Richardson number the face saw -> raw Richardson numbers -> face gradients of x*
    float ribar_s[kMaxNz + 1];
    ribar_s[0] = 0.0f; ribar_s[nz] = 0.0f;
    for (int f = 1; f < nz; ++f) {
      const float sech2 = 1.0f - F.th[f] * F.th[f];
      const bool fin = isfinite(F.ri_s[f]);
      ribar_s[f] = fin ? nubar[f] * (-P.nu_m / (2.0f * P.dRi) * sech2) : 0.0f;
      gp[0] += nubar[f];
      gp[1] += nubar[f] * ((1.0f - F.th[f]) / 2.0f);
      if (fin) {
        gp[2] += nubar[f] * (P.nu_m / 2.0f * sech2 * (F.ri_s[f] - P.Ric) / (P.dRi * P.dRi));
        gp[3] += nubar[f] * (P.nu_m / (2.0f * P.dRi) * sech2);
      }
    }
    float gb[3][2];   // cotangents of the three gradients on faces f (slot 1) and f - 1 (slot 0), rolling
    gb[0][0] = gb[1][0] = gb[2][0] = 0.0f;
    for (int f = 1; f <= nz; ++f) {
      float g3[3] = {0.0f, 0.0f, 0.0f};
      if (f < nz) {
        float rb = ribar_s[f];
        if (P.flags & OPZ_FLAG_SMOOTH_RI) rb = (ribar_s[f - 1] + ribar_s[f] + ribar_s[f + 1]) * (1.0f / 3.0f);
        const float gu = (xs[f] - xs[f - 1]) * P.nzf, gv = (xs[nz + f] - xs[nz + f - 1]) * P.nzf, gT = (xs[2 * nz + f] - xs[2 * nz + f - 1]) * P.nzf;
        const float ea = P.s_u * (gu + P.eps), eb = P.s_v * (gv + P.eps);
        const float S2 = ea * ea + eb * eb;
        const float Ri = P.cBz * (gT + P.eps) / S2;
        const float dRi_dgT = P.cBz / S2;
        if (isfinite(Ri) && isfinite(dRi_dgT)) {
          g3[0] = rb * (-Ri * 2.0f * P.s_u * ea / S2);
          g3[1] = rb * (-Ri * 2.0f * P.s_v * eb / S2);
          g3[2] = rb * dRi_dgT;
        }
      }
      for (int v = 0; v < 3; ++v) {
        gb[v][1] = g3[v];
        xbar[v * nz + f - 1] += P.nzf * (gb[v][0] - gb[v][1]);   // cell f - 1 is the upper cell of face f - 1, the lower of face f
        gb[v][0] = gb[v][1];
      }
    }
    for (int i = 0; i < S3; ++i) {
      a.LAM[col * S3 + i] = xbar[i];
      a.KBAR[col * S3 + i] = a.rk_b[0] * xbar[i];
    }
  }
  if (a.gphys) mpp_param_flush(gp, a.gphys);
}

// ---- lambda_N and the first tendency cotangent ----------------------------------------------------------
__global__ void init_lambda_kernel(const __grid_constant__ Args a) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.B * a.S3) return;
  const int64_t col = e / a.S3;
  const int jj = (int)(e - col * a.S3);
  const int n = a.n_steps;
  const int nz = a.P.nz, v = jj / nz, cz = jj - v * nz;
  const int snap = n / a.save_every;
  const float* xs = a.XN + ((int64_t)n * a.B + col) * a.S3 + v * nz;
  const float* tg = a.target + ((col * a.n_save + snap) * (int64_t)a.S3) + v * nz;
  const float inv_p = 1.0f / ((float)nz * (float)a.n_save * (float)a.B_total);
  const float inv_g = 1.0f / ((float)(nz + 1) * (float)a.n_save * (float)a.B_total);
  const float d0 = xs[cz] - tg[cz];
  float add = a.loss_w[v] * 2.0f * d0 * inv_p;
  const float sc = a.loss_w[3 + v] * 2.0f * a.P.nzf * inv_g;
  if (sc != 0.0f) {
    if (cz > 0) add += sc * ((d0 - (xs[cz - 1] - tg[cz - 1])) * a.P.nzf);
    if (cz < nz - 1) add -= sc * (((xs[cz + 1] - tg[cz + 1]) - d0) * a.P.nzf);
  }
  a.LAM[e] = add;
  a.KBAR[e] = a.rk_b[a.stages - 1] * add;
}

// ---- loss --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) loss_partial_kernel(const __grid_constant__ Args a) {
  __shared__ double red[6][128];
  const int64_t col = blockIdx.x;
  const int nz = a.P.nz;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int e = threadIdx.x; e < a.n_save * a.S3; e += blockDim.x) {
    const int snap = e / a.S3, i = e - snap * a.S3;
    const int v = i / nz, cz = i - v * nz;
    const int64_t n = (int64_t)snap * a.save_every;
    const float* xs = a.XN + (n * a.B + col) * a.S3 + v * nz;
    const float* tg = a.target + ((col * a.n_save + snap) * (int64_t)a.S3) + v * nz;
    const float d0 = xs[cz] - tg[cz];
    s[v] += (double)d0 * d0;
    if (cz > 0) {
      const float dg = (d0 - (xs[cz - 1] - tg[cz - 1])) * a.P.nzf;
      s[3 + v] += (double)dg * dg;
    }
  }
  for (int c = 0; c < 6; ++c) red[c][threadIdx.x] = s[c];
  __syncthreads();
  if (threadIdx.x < 6) {
    double t = 0;
    for (int i = 0; i < (int)blockDim.x; ++i) t += red[threadIdx.x][i];
    a.loss_part[col * 6 + threadIdx.x] = t;
  }
}

__global__ void loss_final_kernel(const __grid_constant__ Args a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t c = 0; c < a.B; ++c)
    for (int k = 0; k < 6; ++k) s[k] += a.loss_part[c * 6 + k];
  const double nz = a.P.nz;
  double tot = 0;
  for (int k = 0; k < 6; ++k) {
    const double den = (k < 3 ? nz : nz + 1.0) * (double)a.n_save * (double)a.B_total;
    const double v = (double)a.loss_w[k] * s[k] / den;
    a.loss_out[1 + k] = (float)v;
    tot += v;
  }
  a.loss_out[0] = (float)tot;
}

// ---- weight-gradient reduction GEMM: dW[k][n] += sum_r A[r][k] D[r][n] ---------------------------------------
constexpr int kTK = 64, kTN = 64, kRC = 16;
__global__ void __launch_bounds__(256) dw_kernel(const float* __restrict__ A, int64_t a_net, int lda,
                                                 const float* __restrict__ D, int64_t d_net, int ldd, int K, int N,
                                                 int64_t R, int split, float* __restrict__ gW, float* __restrict__ gb,
                                                 int64_t g_net) {
  __shared__ __align__(16) float As[kRC][kTK];
  __shared__ __align__(16) float Ds[kRC][kTN];
  const int tid = threadIdx.x;
  const int tk = tid / 16, tn = tid % 16;
  const int net = blockIdx.z / split, sp = blockIdx.z % split;
  const int k0 = blockIdx.x * kTK, n0 = blockIdx.y * kTN;
  const int64_t per = (R + split - 1) / split;
  const int64_t rbeg = per * sp, rend = min(R, per * (sp + 1));
  A += net * a_net;
  D += net * d_net;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = (blockIdx.x == 0 && tk == 0);
  for (int64_t r0 = rbeg; r0 < rend; r0 += kRC) {
#pragma unroll
    for (int i = 0; i < (kRC * kTK) / 256; ++i) {
      const int e = tid + i * 256;
      const int rr = e / kTK, kk = e % kTK;
      const int64_t r = r0 + rr;
      As[rr][kk] = (r < rend && k0 + kk < K) ? A[r * lda + k0 + kk] : 0.0f;
      Ds[rr][kk] = (r < rend && n0 + kk < N) ? D[r * ldd + n0 + kk] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kRC; ++rr) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[rr][tk * 4]);
      const float4 d4 = *reinterpret_cast<const float4*>(&Ds[rr][tn * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bacc[j] += dv[j];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + tk * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tn * 4 + j;
      if (nn < N) atomicAdd(gW + net * g_net + (size_t)k * N + nn, acc[i][j]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tn * 4 + j;
      if (nn < N) atomicAdd(gb + net * g_net + nn, bacc[j]);
    }
  }
}

// W (K x N, element (k,n) at k*N+n) -> WT (element (n,k) at n*K+k)
__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ WT, int K, int N) {
  __shared__ float t[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, n = n0 + threadIdx.x;
    t[r][threadIdx.x] = (k < K && n < N) ? W[(size_t)k * N + n] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, k = k0 + threadIdx.x;
    if (n < N && k < K) WT[(size_t)n * K + k] = t[threadIdx.x][r];
  }
}

#include "opz_bptt_cluster.cuh"

}  // namespace bptt

// ---- host orchestration ----------------------------------------------------------------------------------------
namespace {

struct GraphPair {
  cudaGraphExec_t exec = nullptr;
  int steps = 0;
  int64_t kernels = 0;
};

void destroy(GraphPair& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  g.exec = nullptr;
}

// The instantiated graphs depend on every by-value kernel argument; a training loop repeats the same
// call (same scratch buffers, same shapes), so they are cached on the context and rebuilt on any change.
struct BpttCache {
  bptt::Args key;
  int U = 0;
  bool valid = false;
  GraphPair fwd_rec_U, fwd_rec_1, fwd_ring_U, fwd_ring_1, bwd_U, bwd_1;
  void clear() {
    destroy(fwd_rec_U); destroy(fwd_rec_1); destroy(fwd_ring_U); destroy(fwd_ring_1); destroy(bwd_U); destroy(bwd_1);
    valid = false;
  }
};

}  // namespace

void bptt_release(opz_ctx* ctx) {
  BpttCache* bc = static_cast<BpttCache*>(ctx->bptt_cache);
  if (bc) { bc->clear(); delete bc; }
  ctx->bptt_cache = nullptr;
  if (ctx->capture_stream) cudaStreamDestroy(ctx->capture_stream);
  ctx->capture_stream = nullptr;
}

namespace {

// Shared-memory plan of the cluster kernel for CP column slots per cluster; returns false if it cannot run.
bool plan_cluster(const opz_ctx* ctx, const NetDesc& nd, int nz, int CP, int CS, bptt::cl::Layout& lay) {
  using namespace bptt;
  const int L = nd.n_layers, S3 = 3 * nz, nout = nz - 1, NO3 = 3 * nout;
  std::memset(&lay, 0, sizeof(lay));
  if (S3 * CP > cl::kThreads || NO3 * CP > cl::kThreads) return false;
  int off = 0, ulmax = 0;
  auto take = [&](int n) { const int o = off; off += (n + 3) / 4 * 4; return o; };
  for (int l = 0; l < L - 1; ++l) {
    const int UL = (nd.sizes[l + 1] + CS - 1) / CS;
    if (3 * UL * CP > cl::kThreads) return false;
    lay.UL[l] = UL;
    ulmax = std::max(ulmax, UL);
    lay.W[l] = take(3 * nd.sizes[l] * UL);
    lay.bias[l] = take(3 * UL);
    lay.gs[l] = take(3 * UL * CP);
    if (l <= L - 3) lay.hfull[l] = take(3 * nd.sizes[l + 1] * CP);
  }
  lay.W[L - 1] = take(3 * lay.UL[L - 2] * nout);
  lay.bias[L - 1] = take(NO3);
  lay.hs = take(3 * ulmax * CP);
  lay.dzs = take(3 * ulmax * CP);
  lay.part = take(cl::kThreads * CP);
  lay.y = take(std::max(NO3, S3) * CP);
  lay.dy = take(NO3 * CP);
  lay.xn = take(S3 * CP); lay.xs = take(S3 * CP); lay.kacc = take(S3 * CP); lay.lam = take(S3 * CP);
  lay.kbar = take(S3 * CP); lay.xsum = take(S3 * CP); lay.xphys = take(S3 * CP);
  lay.Fs = take(3 * (nz + 1) * CP);
  lay.recv_big_size = (CS * 3 * ulmax * CP + 3) / 4 * 4;
  off = (off + 3) / 4 * 4;   // 16-byte aligned: the reduce-scatter stores CP columns per transaction
  lay.recv_big = take(lay.recv_big_size * (L >= 4 ? 2 : 1));
  lay.SL = (std::max(NO3, S3) * CP + CS - 1) / CS;
  lay.recv_small = take(CS * lay.SL);
  lay.rbase = take(CS);
  lay.mbar = take(2 * cl::kNumXB);
  lay.total = off;
  return (size_t)off * sizeof(float) <= (size_t)ctx->smem_optin;
}

// groups == 0: only queries how many clusters can be resident at once
template <int CP, int CS>
int launch_cluster(opz_ctx* ctx, const bptt::Args& a, const bptt::cl::Layout& lay, int cg, int64_t groups, int* max_active) {
  using namespace bptt;
  auto kern = cl::sweep_kernel<CP, CS>;
  const size_t smem = (size_t)lay.total * sizeof(float);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    (void)cudaGetLastError();
    return OPZ_EUNSUP;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(std::max<int64_t>(groups, 1) * CS));
  cfg.blockDim = dim3(cl::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
    (void)cudaGetLastError();
    return OPZ_EUNSUP;
  }
  *max_active = n;
  if (groups == 0) return OPZ_OK;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, lay, cg);
  if (e != cudaSuccess) return cuda_fail(e, "cluster sweep launch");
  ctx->launches += 1;
  return OPZ_OK;
}

int slots_for(int cg) { return cg <= 1 ? 1 : (cg <= 2 ? 2 : 4); }

int dispatch_cluster(opz_ctx* ctx, const bptt::Args& a, const bptt::cl::Layout& lay, int cs, int cg, int64_t groups, int* max_active) {
  if (cs == 1) {
    switch (slots_for(cg)) {
      case 1: return launch_cluster<1, 1>(ctx, a, lay, cg, groups, max_active);
      case 2: return launch_cluster<2, 1>(ctx, a, lay, cg, groups, max_active);
      default: return launch_cluster<4, 1>(ctx, a, lay, cg, groups, max_active);
    }
  }
  switch (slots_for(cg)) {
    case 1: return launch_cluster<1, bptt::cl::kCS>(ctx, a, lay, cg, groups, max_active);
    case 2: return launch_cluster<2, bptt::cl::kCS>(ctx, a, lay, cg, groups, max_active);
    default: return launch_cluster<4, bptt::cl::kCS>(ctx, a, lay, cg, groups, max_active);
  }
}

}  // namespace

int fp32_loss_grad(opz_ctx* ctx, const opz_model* m, const LossGradCall& c) {
  using namespace bptt;
  const NetDesc& nd = m->nd;
  const int L = nd.n_layers;
  const int nz = m->nz, S3 = 3 * nz, nout = nz - 1;
  const int64_t B = c.r.B;
  const size_t P3 = 3 * (size_t)nd.P;
  cudaStream_t st = ctx->stream;
  if (L < 2) { set_error("opz_loss_grad: the nets need at least one hidden layer"); return OPZ_EUNSUP; }
  OPZ_CUDA(cudaMemsetAsync(c.grad_out, 0, sizeof(float) * P3, st));
  OPZ_CUDA(cudaMemsetAsync(c.loss_out, 0, sizeof(float) * 7, st));
  if (c.gphys_out) OPZ_CUDA(cudaMemsetAsync(c.gphys_out, 0, sizeof(float) * 5, st));
  if (B == 0) return OPZ_OK;

  const int stages = c.r.scheme == OPZ_RK4 ? 4 : (c.r.scheme == OPZ_RK2 ? 2 : 1);
  const int n_steps = c.r.n_steps;
  // split implicit step: an Euler step of the explicit part followed by the tridiagonal solve; only the graph path carries it
  const bool imex = c.r.scheme == OPZ_IMEX_EULER && (c.r.P.flags & OPZ_FLAG_MPP);

  // ---- memory plan ----
  size_t rec_floats = S3 + 3 * (size_t)nout;  // per (column, right-hand side)
  int max_in = S3;
  for (int l = 0; l < L - 1; ++l) { rec_floats += 6 * (size_t)nd.sizes[l + 1]; max_in = std::max(max_in, nd.sizes[l + 1]); }
  size_t budget = (size_t)24 << 30;
  if (const char* e = getenv("OPZ_BPTT_RECORD_BYTES")) budget = (size_t)atoll(e);
  size_t free_b = 0, total_b = 0;
  OPZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->scratch_bytes[4] + ctx->scratch_bytes[5];  // our own grow-only buffers can be reused
  const size_t xn_only = sizeof(float) * (size_t)(n_steps + 1) * B * S3;
  const size_t xn_bytes = xn_only * (imex ? 2 : 1);   // IMEX keeps the explicit half-step results x* beside the checkpoints
  if (xn_bytes + ((size_t)1 << 30) > free_b) {
    set_error("opz_loss_grad: state checkpoints of " + std::to_string(xn_bytes) + " bytes do not fit device memory");
    return OPZ_ENOMEM;
  }
  budget = std::min(budget, (free_b - xn_bytes) / 2);
  const size_t step_bytes = sizeof(float) * rec_floats * (size_t)B * stages;
  int64_t seg_steps = (int64_t)(budget / step_bytes);
  if (seg_steps < 1) { set_error("opz_loss_grad: one time step of records does not fit the memory budget"); return OPZ_ENOMEM; }
  if (seg_steps > n_steps) seg_steps = n_steps;
  const bool single = (seg_steps >= n_steps);
  const int rcap = (int)std::max<int64_t>(seg_steps * stages, stages);
  const int64_t rows_cap = (int64_t)rcap * B;

  float *XN, *REC, *SMALL;
  int rc;
  if ((rc = ensure_scratch(ctx, 4, xn_bytes, (void**)&XN))) return rc;
  if ((rc = ensure_scratch(ctx, 5, sizeof(float) * rec_floats * (size_t)rows_cap, (void**)&REC))) return rc;
  size_t wT_net = 0;
  int64_t wT_off[kMaxLayers] = {0};
  for (int l = 0; l < L - 1; ++l) { wT_off[l] = (int64_t)wT_net; wT_net += (size_t)nd.sizes[l] * nd.sizes[l + 1]; }
  const size_t small_floats = 5 * (size_t)B * S3 + 3 * wT_net + 64 + (size_t)B * 12 /* 6 doubles */;
  if ((rc = ensure_scratch(ctx, 6, sizeof(float) * small_floats, (void**)&SMALL))) return rc;

  Args a;
  std::memset(&a, 0, sizeof(a));  // the struct doubles as the graph-cache key (padding included)
  a.P = c.r.P;
  a.nd = nd;
  a.w = m->w_dev;
  a.bc = c.r.bc;
  a.target = c.target;
  a.XN = XN;
  a.XSTAR = imex ? XN + (size_t)(n_steps + 1) * B * S3 : nullptr;
  a.imex_flags = c.r.P.flags;
  if (c.r.scheme == OPZ_IMEX_EULER) a.P.flags &= ~(uint32_t)(OPZ_FLAG_MPP | OPZ_FLAG_CA | OPZ_FLAG_SMOOTH_RI);   // the explicit part
  {
    float* p = REC;
    a.X = p; p += (size_t)rows_cap * S3;
    for (int l = 0; l < L - 1; ++l) {
      a.H[l] = p; p += 3 * (size_t)rows_cap * nd.sizes[l + 1];
      a.G[l] = p; p += 3 * (size_t)rows_cap * nd.sizes[l + 1];
    }
    a.DY = p;
  }
  {
    float* p = SMALL;
    a.loss_part = reinterpret_cast<double*>(p); p += (size_t)B * 12;
    a.KACC = p; p += (size_t)B * S3;
    a.LAM = p; p += (size_t)B * S3;
    a.KBAR = p; p += (size_t)B * S3;
    a.XSUM = p; p += (size_t)B * S3;
    a.XPHYS = p; p += (size_t)B * S3;
    a.cur = reinterpret_cast<Cursor*>(p); p += 64;
    a.wT = p;
  }
  float* wT_dev = const_cast<float*>(a.wT);
  Cursor* cur_dev = const_cast<Cursor*>(a.cur);
  for (int l = 0; l < kMaxLayers; ++l) a.wT_off[l] = wT_off[l];
  a.wT_net = (int64_t)wT_net;
  a.loss_out = c.loss_out;
  a.grad = c.grad_out;
  a.gphys = c.gphys_out;
  a.B = B; a.B_total = c.B_total; a.rows_cap = rows_cap;
  a.S3 = S3; a.nout = nout; a.L = L;
  a.scheme = c.r.scheme; a.stages = stages;
  a.n_steps = n_steps; a.save_every = c.r.save_every; a.n_save = c.r.n_save;
  a.rcap = rcap;
  a.h = c.r.dt; a.t0 = c.r.t0;
  {
    const float h = c.r.dt;
    if (c.r.scheme == OPZ_RK4) {
      const float b[4] = {h / 6.0f, h / 3.0f, h / 3.0f, h / 6.0f}, aa[4] = {h * 0.5f, h * 0.5f, h, 0.f}, cc[4] = {0.f, 0.5f, 0.5f, 1.0f};
      for (int i = 0; i < 4; ++i) { a.rk_b[i] = b[i]; a.rk_a[i] = aa[i]; a.rk_c[i] = cc[i]; }
    } else if (c.r.scheme == OPZ_RK2) {
      a.rk_b[0] = a.rk_b[1] = h * 0.5f; a.rk_a[0] = h; a.rk_c[0] = 0.f; a.rk_c[1] = 1.0f;
    } else {
      a.rk_b[0] = h; a.rk_c[0] = 0.f;
    }
  }
  for (int i = 0; i < 6; ++i) a.loss_w[i] = c.loss_w[i];

  // weight gradient: one reduction GEMM per layer over all recorded right-hand sides
  auto dw_all = [&](int64_t rows) -> int {
    for (int l = 0; l < L; ++l) {
      const int K = nd.sizes[l], N = nd.sizes[l + 1];
      const float* Aop = l == 0 ? a.X : a.H[l - 1];
      const int64_t a_net = l == 0 ? 0 : rows_cap * (int64_t)K;
      const float* Dop = l == L - 1 ? a.DY : a.G[l];
      const int64_t d_net = rows_cap * (int64_t)N;
      const int tiles = ((K + kTK - 1) / kTK) * ((N + kTN - 1) / kTN) * 3;
      int split = (4 * ctx->sm_count + tiles - 1) / tiles;
      split = (int)std::max<int64_t>(1, std::min<int64_t>(split, rows / 256));
      dim3 g((K + kTK - 1) / kTK, (N + kTN - 1) / kTN, 3 * split);
      dw_kernel<<<g, 256, 0, st>>>(Aop, a_net, K, Dop, d_net, N, K, N, rows, split, c.grad_out + nd.w_off[l],
                                   c.grad_out + nd.b_off[l], nd.P);
      ctx->launches += 1;
    }
    OPZ_CUDA(cudaGetLastError());
    return OPZ_OK;
  };
  const bool prof = getenv("OPZ_BPTT_PROFILE") != nullptr;

  // ---- cluster path: the whole sweep of a column group inside one 16-CTA cluster (weights in shared memory) ----
  if (single && n_steps > 0 && !imex && !getenv("OPZ_BPTT_NO_CLUSTER")) {
    cl::Layout lay;
    int CG = 0, CS = 0, max_active = 0;
    // fewest columns per cluster that still lets all groups run concurrently; else the widest that fits.  Nets whose FP32
    // weights fit ONE SM's shared memory (the 50/20 production net: 78 KB) run with a "cluster" of one CTA: no DSMEM
    // exchange, __syncthreads instead of cluster barriers, one CTA per column group.
    auto try_cg = [&](int cs, int cg) -> bool {
      cl::Layout tmp;
      int ma = 0;
      if (!plan_cluster(ctx, nd, nz, slots_for(cg), cs, tmp)) return false;
      if (dispatch_cluster(ctx, a, tmp, cs, cg, 0, &ma) != OPZ_OK) return false;
      CG = cg; CS = cs; lay = tmp; max_active = ma;
      return true;
    };
    const bool allow_single = !getenv("OPZ_BPTT_NO_SINGLE_CTA");
    bool chosen = false;
    if (allow_single)
      for (int cg = 1; cg <= 4 && !chosen; ++cg)
        if (try_cg(1, cg) && (B + cg - 1) / cg <= max_active) chosen = true;
    if (!chosen) {
      CG = 0;
      for (int cg = 1; cg <= 4; ++cg)
        if (try_cg(cl::kCS, cg) && (B + cg - 1) / cg <= max_active) break;
    }
    if (const char* e = getenv("OPZ_BPTT_CLUSTER_COLUMNS")) {
      const int cg = atoi(e);
      if (cg >= 1 && cg <= 4) try_cg(CS ? CS : cl::kCS, cg);
    }
    if (CG > 0) {
      cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
      if (prof) { for (auto& e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], st); }
      OPZ_CUDA(cudaMemcpyAsync(XN, c.r.x0, sizeof(float) * B * S3, cudaMemcpyDeviceToDevice, st));
      Args ac = a;
      ac.record = 1;
      const int64_t groups = (B + CG - 1) / CG;
      int ma = 0;
      if ((rc = dispatch_cluster(ctx, ac, lay, CS, CG, groups, &ma))) return rc;
      if (prof) cudaEventRecord(ev[1], st);
      loss_partial_kernel<<<(unsigned)B, 128, 0, st>>>(a);
      loss_final_kernel<<<1, 32, 0, st>>>(a);
      ctx->launches += 2;
      if ((rc = dw_all((int64_t)n_steps * stages * B))) return rc;
      if (prof) {
        cudaEventRecord(ev[2], st);
        cudaEventSynchronize(ev[2]);
        float f = 0, g = 0;
        cudaEventElapsedTime(&f, ev[0], ev[1]);
        cudaEventElapsedTime(&g, ev[1], ev[2]);
        fprintf(stderr, "opz bptt profile (cluster path, %d CTA(s) per cluster): sweep %.2f ms (%d columns/cluster, %lld clusters, %d resident), "
                "loss + dW %.2f ms, shared memory %.1f KB, records %.2f GB\n", CS, f, CG, (long long)groups, max_active, g,
                lay.total * 4 / 1024.0, (double)(sizeof(float) * rec_floats * (size_t)rows_cap) / 1e9);
        for (auto& e : ev) cudaEventDestroy(e);
        unsigned long long ph[24];
        cudaMemcpyFromSymbol(ph, cl::g_phase_cycles, sizeof(ph));
        const char* names[24] = {"fwd slice matvec", "fwd epilogue+gather", "fwd last-layer partial", "fwd y all-reduce", "fwd physics faces",
                                 "fwd cells+RK", "", "", "bwd land+prefetch", "bwd physics VJP", "bwd dz last", "bwd dh partial+send",
                                 "bwd reduce-scatter wait", "bwd dz hidden", "bwd dx partial", "bwd dx all-reduce", "bwd RK adjoint", "  fwd matvec: loop", "  fwd matvec: shuffle+store", "  fwd matvec: sync", "  fwd matvec: plane sum", "", "", ""};
        const double per = 1.0 / ((double)n_steps * stages);
        for (int i = 0; i < 24; ++i)
          if (ph[i]) fprintf(stderr, "    phase %-26s %8.0f cycles / right-hand side\n", names[i], ph[i] * per);
      }
      return OPZ_OK;
    }
  }

  // ---- transposed weight copies for the backward GEMMs ----
  for (int net = 0; net < 3; ++net)
    for (int l = 0; l < L - 1; ++l) {
      const int K = nd.sizes[l], N = nd.sizes[l + 1];
      dim3 g((K + 31) / 32, (N + 31) / 32), b(32, 8);
      transpose_kernel<<<g, b, 0, st>>>(m->w_dev + (size_t)net * nd.P + nd.w_off[l], wT_dev + (size_t)net * wT_net + wT_off[l], K, N);
      ctx->launches += 1;
    }
  OPZ_CUDA(cudaGetLastError());

  // ---- launch configuration ----
  const int n_ct = (int)((B + kCT - 1) / kCT);
  const size_t gemm_smem = sizeof(float) * ((size_t)std::min(max_in, kChunk) * kCT + (size_t)kKS * kCT * kNT);
  if ((int)gemm_smem > ctx->smem_optin) { set_error("opz_loss_grad: layer too wide for the GEMM tile"); return OPZ_EUNSUP; }
  OPZ_CUDA(cudaFuncSetAttribute(skinny_kernel<MODE_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  OPZ_CUDA(cudaFuncSetAttribute(skinny_kernel<MODE_BWD_HID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  OPZ_CUDA(cudaFuncSetAttribute(skinny_kernel<MODE_BWD_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
  const int OP = (3 * nout + 31) / 32 * 32;
  int ff_threads = std::max(OP, (1024 / OP) * OP);
  ff_threads = std::max(ff_threads, (S3 + 31) / 32 * 32);
  if (ff_threads > 1024) { set_error("opz_loss_grad: nz too large"); return OPZ_EUNSUP; }
  const int nh = nd.sizes[L - 1];
  const size_t ff_smem = sizeof(float) * ((size_t)3 * nh + (size_t)(ff_threads / OP) * OP + S3 + 3 * (size_t)(nz + 1));
  const size_t fb_smem = sizeof(float) * ((size_t)2 * S3 + 3 * (size_t)nout + 3 * (size_t)(nz + 1));
  OPZ_CUDA(cudaFuncSetAttribute(final_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_smem));
  const int fb_threads = 1024;

  int64_t step_kernels_fwd = 0, step_kernels_bwd = 0;
  auto enqueue_fwd_step = [&](const Args& aa, int dn, cudaStream_t s) {
    for (int sg = 0; sg < stages; ++sg) {
      for (int l = 0; l < L - 1; ++l) {
        dim3 g((nd.sizes[l + 1] + kNT - 1) / kNT, n_ct, 3);
        skinny_kernel<MODE_FWD><<<g, kGemmThreads, gemm_smem, s>>>(aa, l, sg, dn);
      }
      final_fwd_kernel<<<(unsigned)B, ff_threads, ff_smem, s>>>(aa, sg, dn);
    }
    if (imex) imex_fwd_kernel<<<(unsigned)((B + 31) / 32), 32, 0, s>>>(aa, dn);
  };
  step_kernels_fwd = (int64_t)stages * L + (imex ? 1 : 0);
  auto enqueue_bwd_step = [&](const Args& aa, int dn, cudaStream_t s) {
    if (imex) imex_bwd_kernel<<<(unsigned)((B + 31) / 32), 32, 0, s>>>(aa, dn);
    for (int sg = stages - 1; sg >= 0; --sg) {
      first_bwd_kernel<<<(unsigned)B, fb_threads, fb_smem, s>>>(aa, sg, dn);
      for (int l = L - 2; l >= 1; --l) {
        dim3 g((nd.sizes[l] + kNT - 1) / kNT, n_ct, 3);
        skinny_kernel<MODE_BWD_HID><<<g, kGemmThreads, gemm_smem, s>>>(aa, l, sg, dn);
      }
      dim3 g((S3 + kNT - 1) / kNT, n_ct, 1);
      skinny_kernel<MODE_BWD_IN><<<g, kGemmThreads, gemm_smem, s>>>(aa, 0, sg, dn);
    }
  };
  step_kernels_bwd = (int64_t)stages * L + (imex ? 1 : 0);

  // ---- graphs: U unrolled steps + cursor bump, captured on a private stream ----
  if (!ctx->capture_stream) OPZ_CUDA(cudaStreamCreateWithFlags(&ctx->capture_stream, cudaStreamNonBlocking));
  cudaStream_t cs = ctx->capture_stream;
  auto build = [&](bool forward, const Args& aa, int U, GraphPair& out) -> int {
    cudaGraph_t graph = nullptr;
    OPZ_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    for (int u = 0; u < U; ++u) {
      if (forward) enqueue_fwd_step(aa, u, cs);
      else enqueue_bwd_step(aa, -u, cs);
    }
    bump_cursor_kernel<<<1, 1, 0, cs>>>(cur_dev, forward ? U : -U);
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&out.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
    out.steps = U;
    out.kernels = (forward ? step_kernels_fwd : step_kernels_bwd) * U + 1;
    return OPZ_OK;
  };
  int U = 16;
  if (const char* e = getenv("OPZ_BPTT_GRAPH_STEPS")) U = std::max(1, atoi(e));
  U = std::min(U, std::max(1, n_steps));
  BpttCache* cache = static_cast<BpttCache*>(ctx->bptt_cache);
  if (!cache) { cache = new BpttCache(); ctx->bptt_cache = cache; }
  GraphPair &gf_rec_U = cache->fwd_rec_U, &gf_rec_1 = cache->fwd_rec_1, &gf_ring_U = cache->fwd_ring_U,
            &gf_ring_1 = cache->fwd_ring_1, &gb_U = cache->bwd_U, &gb_1 = cache->bwd_1;
  auto cleanup = [&]() { cudaStreamSynchronize(st); cache->clear(); };
  Args a_rec = a; a_rec.record = 1;
  Args a_ring = a; a_ring.record = 0;
  if (n_steps > 0 && !(cache->valid && cache->U == U && std::memcmp(&cache->key, &a, sizeof(Args)) == 0)) {
    OPZ_CUDA(cudaStreamSynchronize(st));  // graphs of an earlier call may still be in flight
    cache->clear();
    if ((rc = build(true, a_rec, U, gf_rec_U)) || (rc = build(true, a_rec, 1, gf_rec_1)) ||
        (rc = build(false, a_rec, U, gb_U)) || (rc = build(false, a_rec, 1, gb_1))) { cleanup(); return rc; }
    if (!single && ((rc = build(true, a_ring, U, gf_ring_U)) || (rc = build(true, a_ring, 1, gf_ring_1)))) { cleanup(); return rc; }
    std::memcpy(&cache->key, &a, sizeof(Args));
    cache->U = U;
    cache->valid = true;
  }
  auto run_steps = [&](GraphPair& gU, GraphPair& g1, int count) -> int {
    while (count > 0) {
      GraphPair& g = count >= gU.steps ? gU : g1;
      cudaError_t e = cudaGraphLaunch(g.exec, st);
      if (e != cudaSuccess) return cuda_fail(e, "cudaGraphLaunch");
      ctx->launches += g.kernels;
      count -= g.steps;
    }
    return OPZ_OK;
  };
  auto set_cursor = [&](int n, int seg_n0) {
    set_cursor_kernel<<<1, 1, 0, st>>>(cur_dev, n, seg_n0);
    ctx->launches += 1;
  };

  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (prof) for (auto& e : ev) cudaEventCreate(&e);
  if (prof) cudaEventRecord(ev[0], st);
  // ---- forward sweep ----
  OPZ_CUDA(cudaMemcpyAsync(XN, c.r.x0, sizeof(float) * B * S3, cudaMemcpyDeviceToDevice, st));
  OPZ_CUDA(cudaMemcpyAsync(a.X, c.r.x0, sizeof(float) * B * S3, cudaMemcpyDeviceToDevice, st));
  set_cursor(0, 0);
  if (single) rc = run_steps(gf_rec_U, gf_rec_1, n_steps);
  else rc = run_steps(gf_ring_U, gf_ring_1, n_steps);
  if (rc) { cleanup(); return rc; }

  if (prof) cudaEventRecord(ev[1], st);
  // ---- loss ----
  loss_partial_kernel<<<(unsigned)B, 128, 0, st>>>(a);
  loss_final_kernel<<<1, 32, 0, st>>>(a);
  init_lambda_kernel<<<(unsigned)((B * S3 + 255) / 256), 256, 0, st>>>(a);
  ctx->launches += 3;

  // ---- backward sweep, segment by segment ----
  for (int64_t seg_end = n_steps; seg_end > 0;) {
    const int64_t seg_n0 = std::max<int64_t>(0, seg_end - seg_steps);
    const int len = (int)(seg_end - seg_n0);
    if (!single) {  // recompute this segment's forward with records
      OPZ_CUDA(cudaMemcpyAsync(a.X, XN + (size_t)seg_n0 * B * S3, sizeof(float) * B * S3, cudaMemcpyDeviceToDevice, st));
      set_cursor((int)seg_n0, (int)seg_n0);
      if ((rc = run_steps(gf_rec_U, gf_rec_1, len))) { cleanup(); return rc; }
    }
    set_cursor((int)seg_end - 1, (int)seg_n0);
    if ((rc = run_steps(gb_U, gb_1, len))) { cleanup(); return rc; }
    if ((rc = dw_all((int64_t)len * stages * B))) { cleanup(); return rc; }
    seg_end = seg_n0;
  }
  OPZ_CUDA(cudaGetLastError());
  if (prof) {
    cudaEventRecord(ev[3], st);
    cudaEventSynchronize(ev[3]);
    float f = 0, t = 0;
    cudaEventElapsedTime(&f, ev[0], ev[1]);
    cudaEventElapsedTime(&t, ev[0], ev[3]);
    fprintf(stderr, "opz bptt profile: forward sweep %.2f ms, loss + backward + dW %.2f ms, segments %s, records %.2f GB\n", f,
            t - f, single ? "1" : ">1", (double)(sizeof(float) * rec_floats * (size_t)rows_cap) / 1e9);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return OPZ_OK;
}

// ================================================================================================================
// Supervised flux pre-training of ONE net: train_NN / NN_loss (wind_mixing/src/NN_training.jl:207-249),
// predict_uw / predict_vw / predict_wT (:25-169).  Per sample (one profile snapshot):
//     loss = mse(F_pred, F_data) + gradient_scaling * mse(D^c F_data, D^c F_pred)
// with F_pred the total flux of that net on the Nz + 1 faces (what opz_rhs reports) — only its interior entries depend on
// the weights, through the NN output alone, so the gradient is a plain MLP back-propagation of d loss / d y.
// The batch value is the mean over the samples (total_loss, :230-232).
namespace bptt {

// one CTA per sample: forward of the chosen net with records, loss terms, cotangent chain down to dz_0
__global__ void __launch_bounds__(256) flux_sample_kernel(NetDesc nd, DevParams P, const float* __restrict__ w, int which,
                                                          const float* __restrict__ x, const float* __restrict__ Fpred,
                                                          const float* __restrict__ Ftgt, float gscale, int64_t N, int max_w,
                                                          float* __restrict__ H, float* __restrict__ G, int64_t rec_stride,
                                                          float* __restrict__ DY, double* __restrict__ loss_part) {
  extern __shared__ float sm[];
  const int nz = P.nz, S3 = 3 * nz, nout = nz - 1, L = nd.n_layers;
  float* hA = sm;               // [max_w]
  float* hB = hA + max_w;       // [max_w]
  float* res = hB + max_w;      // [nz + 1]
  float* dcs = res + nz + 1;    // [nz]
  const int tid = threadIdx.x;
  const int64_t r = blockIdx.x;
  for (int i = tid; i < S3; i += blockDim.x) hA[i] = x[r * S3 + i];
  __syncthreads();
  // ---- forward with records: H_l = act(z_l), G_l = act'(z_l) for the hidden layers ----
  float* in = hA;
  float* out = hB;
  int64_t roff = 0;
  for (int l = 0; l < L - 1; ++l) {
    const int K = nd.sizes[l], Nl = nd.sizes[l + 1];
    const float* W = w + nd.w_off[l];
    const float* b = w + nd.b_off[l];
    for (int n = tid; n < Nl; n += blockDim.x) {
      float z = b[n];
      for (int k = 0; k < K; ++k) z = fmaf(in[k], W[(size_t)k * Nl + n], z);
      const float hv = act_dyn(z, nd.act);
      out[n] = hv;
      H[roff + r * Nl + n] = hv;
      G[roff + r * Nl + n] = act_grad_dyn(z, nd.act);
    }
    __syncthreads();
    float* t = in; in = out; out = t;
    roff += rec_stride * Nl;
  }
  // ---- loss terms and d loss / d y ----
  const float* Fp = Fpred + (r * 3 + which) * (nz + 1);
  const float* Ft = Ftgt + r * (nz + 1);
  for (int f = tid; f <= nz; f += blockDim.x) res[f] = Fp[f] - Ft[f];
  __syncthreads();
  for (int c = tid; c < nz; c += blockDim.x) dcs[c] = (res[c + 1] - res[c]) * P.nzf;   // D^c of the residual
  __syncthreads();
  if (tid == 0) {
    double l1 = 0.0, l2 = 0.0;
    for (int f = 0; f <= nz; ++f) l1 += (double)res[f] * res[f];
    for (int c = 0; c < nz; ++c) l2 += (double)dcs[c] * dcs[c];
    loss_part[2 * r] = l1 / (nz + 1);
    loss_part[2 * r + 1] = l2 / nz;
  }
  float* dz = out;   // free buffer
  const float invN = 1.0f / (float)N;
  for (int j = tid; j < nout; j += blockDim.x) {
    const int f = j + 1;   // interior face
    const float g = 2.0f * res[f] / (float)(nz + 1) + gscale * (2.0f / (float)nz) * P.nzf * (dcs[f - 1] - dcs[f]);
    dz[j] = g * invN;    // cotangent of the total flux on interior face f
  }
  __syncthreads();
  if (P.flags & OPZ_FLAG_SMOOTH_NN) {   // the flux holds the box-filtered NN output: d loss / d y = S^T (d loss / d F)
    float sv[4];                        // (nout <= 127 < 4 * blockDim.x)
    int cnt = 0;
    for (int j = tid; j < nout; j += blockDim.x, ++cnt) {
      float s2 = faces::box_weight(j, nout) * dz[j];
      if (j > 0) s2 += faces::box_weight(j - 1, nout) * dz[j - 1];
      if (j < nout - 1) s2 += faces::box_weight(j + 1, nout) * dz[j + 1];
      sv[cnt] = s2;
    }
    __syncthreads();
    cnt = 0;
    for (int j = tid; j < nout; j += blockDim.x, ++cnt) dz[j] = sv[cnt];
    __syncthreads();
  }
  for (int j = tid; j < nout; j += blockDim.x) DY[r * nout + j] = dz[j];
  // ---- back through the layers: dz_{l-1} = (W_l dz_l) * act'(z_{l-1}), in place over the act' records ----
  float* cur = dz;
  float* nxt = in;    // the last hidden activation is already recorded
  for (int l = L - 1; l >= 1; --l) {
    const int K = nd.sizes[l], Nl = nd.sizes[l + 1];
    const float* W = w + nd.w_off[l];
    roff -= rec_stride * K;
    for (int k = tid; k < K; k += blockDim.x) {
      const float* Wk = W + (size_t)k * Nl;
      float sacc = 0.0f;
      for (int n = 0; n < Nl; ++n) sacc = fmaf(Wk[n], cur[n], sacc);
      const int64_t o = roff + r * K + k;
      const float v = sacc * G[o];
      G[o] = v;
      nxt[k] = v;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace bptt

int fp32_flux_loss_grad(opz_ctx* ctx, const opz_model* m, const DevParams& P, int which_net, const float* x, const float* bc,
                        const float* flux_target, int64_t N, float gradient_scaling, float* loss_out_host, float* grad_out) {
  using namespace bptt;
  const NetDesc& nd = m->nd;
  const int L = nd.n_layers, nz = m->nz, S3 = 3 * nz, nout = nz - 1;
  cudaStream_t st = ctx->stream;
  OPZ_CUDA(cudaMemsetAsync(grad_out, 0, sizeof(float) * (size_t)nd.P, st));
  if (N == 0) { loss_out_host[0] = loss_out_host[1] = loss_out_host[2] = 0.0f; return OPZ_OK; }
  // records: per hidden layer [N][n_l] of H and G, then DY [N][nout], fluxes [N][3][nz+1], loss partials [N][2] doubles
  size_t hid = 0;
  int max_w = std::max(S3, nout);
  for (int l = 0; l < L - 1; ++l) { hid += (size_t)nd.sizes[l + 1]; max_w = std::max(max_w, nd.sizes[l + 1]); }
  const size_t floats = (size_t)N * (2 * hid + nout + 3 * (size_t)(nz + 1) + 4) + S3 * (size_t)N + 64;
  float* base;
  int rc;
  if ((rc = ensure_scratch(ctx, 5, sizeof(float) * floats, (void**)&base))) return rc;
  double* lp = reinterpret_cast<double*>(base);   // [N][2] (the scratch block is 256-byte aligned)
  float* H = base + 4 * (size_t)N;
  float* G = H + (size_t)N * hid;
  float* DY = G + (size_t)N * hid;
  float* Fp = DY + (size_t)N * nout;
  float* dk = Fp + (size_t)N * 3 * (nz + 1);       // tendencies of the RHS launch (unused)
  if ((rc = fp32_rhs(ctx, m, P, x, bc, 0.0f, N, dk, Fp))) return rc;
  const size_t smem = sizeof(float) * (2 * (size_t)max_w + 2 * (size_t)nz + 1);
  flux_sample_kernel<<<(unsigned)N, 256, smem, st>>>(nd, P, m->w_dev + (size_t)which_net * nd.P, which_net, x, Fp, flux_target,
                                                      gradient_scaling, N, max_w, H, G, N, DY, lp);
  OPZ_CUDA(cudaGetLastError());
  ctx->launches += 1;
  // dW_l = A_l^T D_l over the N samples: A_0 = x, A_l = H_{l-1}; D_l = dz_l (G records), D_{L-1} = DY
  size_t aoff = 0, doff = 0;
  for (int l = 0; l < L; ++l) {
    const int K = nd.sizes[l], Nl = nd.sizes[l + 1];
    const float* A = l == 0 ? x : H + aoff;
    const float* D = l == L - 1 ? DY : G + doff;
    const int split = (int)std::min<int64_t>(std::max<int64_t>(N / 256, 1), 32);
    dim3 grid((K + kTK - 1) / kTK, (Nl + kTN - 1) / kTN, split);
    dw_kernel<<<grid, 256, 0, st>>>(A, 0, K, D, 0, Nl, K, Nl, N, split, grad_out + nd.w_off[l], grad_out + nd.b_off[l], 0);
    OPZ_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (l >= 1) aoff += (size_t)N * K;
    if (l < L - 1) doff += (size_t)N * Nl;
  }
  std::vector<double> hl(2 * (size_t)N);
  OPZ_CUDA(cudaMemcpyAsync(hl.data(), lp, sizeof(double) * hl.size(), cudaMemcpyDeviceToHost, st));
  OPZ_CUDA(cudaStreamSynchronize(st));
  double l1 = 0.0, l2 = 0.0;
  for (int64_t i = 0; i < N; ++i) { l1 += hl[2 * i]; l2 += hl[2 * i + 1]; }
  l1 /= (double)N; l2 /= (double)N;
  loss_out_host[0] = (float)(l1 + (double)gradient_scaling * l2);
  loss_out_host[1] = (float)l1;
  loss_out_host[2] = (float)l2;
  return OPZ_OK;
}

}  // namespace opz
