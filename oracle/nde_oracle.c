/* nde_oracle.c — plain-C CPU restatement of the reference's NDE solve (TEST INFRASTRUCTURE).
 *
 * Same algorithm as oracle/nde_oracle.py, one column per OpenMP thread, allocation-free inner loop in the
 * shape of solve_NDE_mutating's in-place RHS (wind_mixing/src/training_postprocessing.jl:55-159): three Flux
 * Chains evaluated as dense mat-vecs (W is out x in column-major, NDE_training.jl:11-13), two-point stencils
 * for D^f / D^c (src/differentiation_operators.jl:6-29), local_richardson + tanh_step diffusivity
 * (NDE_training.jl:46-54), optional convective adjustment (tp.jl:118-121), 3-point box filters
 * (filtering_operators.jl:1-15), fixed-step Euler / Heun / RK4 instead of ROCK4 (north_star).
 *
 * The body is compiled twice: REAL = float (orc_rollout: the reference's Float32 arithmetic, the CPU baseline
 * that bench.py times) and REAL = double (orc_rollout_f64: the same Float32-specified problem — Float32
 * parameters, weights, forcing and initial state — solved in double, i.e. the noise floor of any Float32
 * comparison).
 *
 * Used by tests/ (cross-check against the NumPy oracle, long-rollout parity of the CUDA path) and by bench.py's
 * cpu_baseline / --impl reference legs ("kind": "port": Julia is not installed in this image, so the reference
 * itself cannot be timed).  PARITY UNPINNED: see the header of nde_oracle.py.
 * Never linked into or called from the product library.
 */
#ifndef ORC_BODY
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define F_MPP (1u << 0)
#define F_CA (1u << 1)
#define F_CA_DUDZ (1u << 2)
#define F_BC_MINUS (1u << 3)
#define F_DIURNAL (1u << 4)
#define F_DIURNAL_MINUS (1u << 5)
#define F_RI_EPS (1u << 6)
#define F_SMOOTH_NN (1u << 7)
#define F_SMOOTH_RI (1u << 8)

typedef struct {
  int32_t nz;
  float H, tau, f, g, alpha, nu0, nu_m, Ric, dRi, Pr, kappa, eps;
  float mu[6], sigma[6];
  uint32_t flags;
  float diurnal_period;
  float reserved[10];
} orc_params;

#define MAXNZ 128
#define MAXW 2048

#define ORC_BODY
#define REAL float
#define FN(x) x##_f32
#define R(x) x##f
#define M_TANH tanhf
#define M_LOG1P log1pf
#define M_EXP expf
#define M_FABS fabsf
#define M_SIN sinf
#include "nde_oracle.c"
#undef REAL
#undef FN
#undef R
#undef M_TANH
#undef M_LOG1P
#undef M_EXP
#undef M_FABS
#undef M_SIN
#define REAL double
#define FN(x) x##_f64
#define R(x) x
#define M_TANH tanh
#define M_LOG1P log1p
#define M_EXP exp
#define M_FABS fabs
#define M_SIN sin
#include "nde_oracle.c"

/* out[B][n_save][3nz] Float32 (save_every <= 0: final state only); returns 0 on success */
int orc_rollout(const orc_params* p, const int* sizes, int n_sizes, int act, const float* w, const float* x0,
                const float* bc, int64_t B, int scheme, float dt, float t0, int n_steps, int save_every,
                float* out, int n_threads) {
  return rollout_f32(p, sizes, n_sizes, act, w, x0, bc, B, scheme, dt, t0, n_steps, save_every, out, n_threads);
}

/* the same with every operation in double (dt and t0 are the Float32 values a caller passes through the ABI) */
int orc_rollout_f64(const orc_params* p, const int* sizes, int n_sizes, int act, const float* w, const float* x0,
                    const float* bc, int64_t B, int scheme, float dt, float t0, int n_steps, int save_every,
                    double* out, int n_threads) {
  return rollout_f64(p, sizes, n_sizes, act, w, x0, bc, B, scheme, dt, t0, n_steps, save_every, out, n_threads);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* cores of the machine, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers) */
int orc_num_procs(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

#else /* ORC_BODY: the generic part, included once per REAL */

static REAL FN(act)(REAL z, int act) {
  switch (act) {
    case 1: return z > R(0.0) ? z : R(0.0);
    case 2: return M_TANH(z);
    case 3: { REAL sp = (z > R(0.0) ? z : R(0.0)) + M_LOG1P(M_EXP(-M_FABS(z))); return z * M_TANH(sp); }
    case 4: return z / (R(1.0) + M_EXP(-z));
    case 5: { REAL a = R(0.01) * z; return a > z ? a : z; }
    default: return z;
  }
}

/* y = W x + b ; W out x in column-major */
static void FN(dense)(const float* W, const float* b, int nin, int nout, const REAL* x, REAL* y, int act) {
  for (int o = 0; o < nout; ++o) y[o] = R(0.0);
  for (int k = 0; k < nin; ++k) {
    const REAL xk = x[k];
    const float* col = W + (size_t)k * nout;
    for (int o = 0; o < nout; ++o) y[o] += (REAL)col[o] * xk;
  }
  for (int o = 0; o < nout; ++o) y[o] = FN(act)(y[o] + (REAL)b[o], act);
}

static void FN(chain)(const float* w, const int* sizes, int n_sizes, int act, const REAL* x, REAL* out, REAL* t0,
                      REAL* t1) {
  const REAL* cur = x;
  size_t off = 0;
  for (int l = 0; l + 1 < n_sizes; ++l) {
    const int nin = sizes[l], nout = sizes[l + 1];
    const int last = (l + 2 == n_sizes);
    REAL* dst = last ? out : ((l & 1) ? t1 : t0);
    FN(dense)(w + off, w + off + (size_t)nin * nout, nin, nout, cur, dst, last ? 0 : act);
    off += (size_t)nin * nout + nout;
    cur = dst;
  }
}

static void FN(box3)(const REAL* a, REAL* o, int n) {
  o[0] = (a[0] + a[1]) * R(0.5);
  o[n - 1] = (a[n - 2] + a[n - 1]) * R(0.5);
  for (int i = 1; i < n - 1; ++i) o[i] = (a[i - 1] + a[i] + a[i + 1]) * (R(1.0) / R(3.0));
}

typedef struct {
  REAL y[3][MAXNZ], F[3][MAXNZ + 1], g[3][MAXNZ + 1], Ri[MAXNZ + 1], tmp[MAXNZ + 1], t0[MAXW], t1[MAXW];
} FN(work_t);

static void FN(rhs_col)(const orc_params* p, const int* sizes, int n_sizes, int act, const float* w, size_t P,
                        const REAL* x, const float* bc, REAL t, REAL* k, FN(work_t)* wk) {
  const int Nz = p->nz;
  const REAL nzf = (REAL)Nz;
  const REAL *u = x, *v = x + Nz, *T = x + 2 * Nz;
  const REAL s_u = p->sigma[0], s_v = p->sigma[1], s_T = p->sigma[2];
  const REAL s_f[3] = {p->sigma[3], p->sigma[4], p->sigma[5]}, mu_f[3] = {p->mu[3], p->mu[4], p->mu[5]};
  const REAL H = p->H, tau = p->tau, grav = p->g, alpha = p->alpha;
  for (int n = 0; n < 3; ++n) {
    FN(chain)(w + n * P, sizes, n_sizes, act, x, wk->y[n], wk->t0, wk->t1);
    if (p->flags & F_SMOOTH_NN) { FN(box3)(wk->y[n], wk->tmp, Nz - 1); memcpy(wk->y[n], wk->tmp, sizeof(REAL) * (Nz - 1)); }
  }
  const REAL sub = (p->flags & F_BC_MINUS) ? R(1.0) : R(0.0);
  for (int n = 0; n < 3; ++n) {
    const REAL off = sub * ((-mu_f[n]) / s_f[n]);
    wk->F[n][0] = (REAL)bc[2 * n] - off;
    wk->F[n][Nz] = (REAL)bc[2 * n + 1] - off;
    for (int j = 0; j < Nz - 1; ++j) wk->F[n][j + 1] = wk->y[n][j];
  }
  if (p->flags & F_DIURNAL) {
    const REAL omega = (REAL)(2.0 * 3.14159265358979323846 / (double)p->diurnal_period);
    const REAL phys = (REAL)bc[6] * M_SIN(omega * (t * tau)) / (alpha * grav);
    REAL top = (phys - mu_f[2]) / s_f[2];
    if (p->flags & F_DIURNAL_MINUS) top -= (-mu_f[2]) / s_f[2];
    wk->F[2][Nz] = top;
  }
  if (p->flags & F_MPP) {
    const REAL* c[3] = {u, v, T};
    for (int n = 0; n < 3; ++n) {
      wk->g[n][0] = R(0.0); wk->g[n][Nz] = R(0.0);
      for (int f = 1; f < Nz; ++f) wk->g[n][f] = (c[n][f] - c[n][f - 1]) * nzf;
    }
    const REAL e = (p->flags & F_RI_EPS) ? (REAL)p->eps : R(0.0);
    const REAL cBz = ((H * grav) * alpha) * s_T;
    for (int f = 0; f <= Nz; ++f) {
      const REAL a = s_u * (wk->g[0][f] + e), b = s_v * (wk->g[1][f] + e);
      wk->Ri[f] = (cBz * (wk->g[2][f] + e)) / (a * a + b * b);
    }
    if (p->flags & F_SMOOTH_RI) { FN(box3)(wk->Ri, wk->tmp, Nz + 1); memcpy(wk->Ri, wk->tmp, sizeof(REAL) * (Nz + 1)); }
    const REAL c_u = s_u / s_f[0] / H, c_v = s_v / s_f[1] / H, c_T = s_T / s_f[2] / H;
    const REAL nu0 = p->nu0, nu_m = p->nu_m, Ric = p->Ric, dRi = p->dRi, Pr = p->Pr, kappa = p->kappa;
    for (int f = 1; f < Nz; ++f) {
      const REAL nu = nu0 + nu_m * ((R(1.0) - M_TANH((wk->Ri[f] - Ric) / dRi)) / R(2.0));
      REAL nuT = nu / Pr;
      if (p->flags & F_CA) {
        const REAL sel = (p->flags & F_CA_DUDZ) ? wk->g[0][f] : wk->g[2][f];
        nuT = sel > R(0.0) ? nu / Pr : kappa;
      }
      wk->F[0][f] -= c_u * nu * wk->g[0][f];
      wk->F[1][f] -= c_v * nu * wk->g[1][f];
      wk->F[2][f] -= c_T * nuT * wk->g[2][f];
    }
  }
  const REAL A = -tau / H, ft = (REAL)p->f * tau;
  const REAL A_u = A * s_f[0] / s_u, A_v = A * s_f[1] / s_v, A_T = A * s_f[2] / s_T;
  const REAL cor_u = ft / s_u, cor_v = ft / s_v;
  const REAL mu_u = p->mu[0], mu_v = p->mu[1];
  for (int c = 0; c < Nz; ++c) {
    k[c] = A_u * ((wk->F[0][c + 1] - wk->F[0][c]) * nzf) + cor_u * (s_v * v[c] + mu_v);
    k[Nz + c] = A_v * ((wk->F[1][c + 1] - wk->F[1][c]) * nzf) - cor_v * (s_u * u[c] + mu_u);
    k[2 * Nz + c] = A_T * ((wk->F[2][c + 1] - wk->F[2][c]) * nzf);
  }
}

static int FN(rollout)(const orc_params* p, const int* sizes, int n_sizes, int act, const float* w, const float* x0,
                       const float* bc, int64_t B, int scheme, float dt_f, float t0_f, int n_steps, int save_every,
                       REAL* out, int n_threads) {
  const int Nz = p->nz, S = 3 * Nz;
  if (Nz > MAXNZ) return -1;
  size_t P = 0;
  for (int l = 0; l + 1 < n_sizes; ++l) {
    if (sizes[l + 1] > MAXW) return -1;
    P += (size_t)sizes[l] * sizes[l + 1] + sizes[l + 1];
  }
  const int n_save = save_every > 0 ? n_steps / save_every + 1 : 1;
  const REAL dt = dt_f, t0 = t0_f;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    FN(work_t)* wk = (FN(work_t)*)malloc(sizeof(FN(work_t)));
    REAL x[3 * MAXNZ], xs[3 * MAXNZ], k[3 * MAXNZ], acc[3 * MAXNZ];
#pragma omp for schedule(dynamic, 1)
    for (int64_t col = 0; col < B; ++col) {
      const float* b = bc + col * 8;
      for (int i = 0; i < S; ++i) x[i] = (REAL)x0[col * S + i];
      REAL* o = out + col * (size_t)n_save * S;
      if (save_every > 0) memcpy(o, x, sizeof(REAL) * S);
      for (int n = 0; n < n_steps; ++n) {
        const REAL t = t0 + (REAL)n * dt;
        if (scheme == 0) {
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, x, b, t, k, wk);
          for (int i = 0; i < S; ++i) x[i] = x[i] + dt * k[i];
        } else if (scheme == 1) {
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, x, b, t, k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = k[i]; xs[i] = x[i] + dt * k[i]; }
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, xs, b, t + dt, k, wk);
          for (int i = 0; i < S; ++i) x[i] = x[i] + (dt * R(0.5)) * (acc[i] + k[i]);
        } else {
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, x, b, t, k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = k[i]; xs[i] = x[i] + (dt * R(0.5)) * k[i]; }
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, xs, b, t + dt * R(0.5), k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = acc[i] + R(2.0) * k[i]; xs[i] = x[i] + (dt * R(0.5)) * k[i]; }
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, xs, b, t + dt * R(0.5), k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = acc[i] + R(2.0) * k[i]; xs[i] = x[i] + dt * k[i]; }
          FN(rhs_col)(p, sizes, n_sizes, act, w, P, xs, b, t + dt, k, wk);
          for (int i = 0; i < S; ++i) x[i] = x[i] + (dt / R(6.0)) * (acc[i] + k[i]);
        }
        if (save_every > 0 && (n + 1) % save_every == 0) memcpy(o + (size_t)((n + 1) / save_every) * S, x, sizeof(REAL) * S);
      }
      if (save_every <= 0) memcpy(o, x, sizeof(REAL) * S);
    }
    free(wk);
  }
  return 0;
}

#endif /* ORC_BODY */
