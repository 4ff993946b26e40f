/* nde_oracle.c — plain-C CPU restatement of the reference's NDE solve (TEST INFRASTRUCTURE).
 *
 * Same algorithm as oracle/nde_oracle.py, one column per OpenMP thread, Float32, allocation-free
 * inner loop in the shape of solve_NDE_mutating's in-place RHS
 * (wind_mixing/src/training_postprocessing.jl:55-159): three Flux Chains evaluated as dense
 * mat-vecs (W is out x in column-major, NDE_training.jl:11-13), two-point stencils for D^f / D^c
 * (src/differentiation_operators.jl:6-29), local_richardson + tanh_step diffusivity
 * (NDE_training.jl:46-54), optional convective adjustment (tp.jl:118-121), fixed-step
 * Euler / Heun / RK4 instead of ROCK4 (north_star).
 *
 * Used by tests/ (cross-check against the NumPy oracle) and by bench.py's cpu_baseline /
 * --impl reference legs ("kind": "port": Julia is not installed in this image, so the reference
 * itself cannot be timed).  PARITY UNPINNED: see the header of nde_oracle.py.
 * Never linked into or called from the product library.
 */
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

static float act_f(float z, int act) {
  switch (act) {
    case 1: return z > 0.f ? z : 0.f;
    case 2: return tanhf(z);
    case 3: { float sp = (z > 0.f ? z : 0.f) + log1pf(expf(-fabsf(z))); return z * tanhf(sp); }
    case 4: return z / (1.f + expf(-z));
    case 5: { float a = 0.01f * z; return a > z ? a : z; }
    default: return z;
  }
}

/* y = W x + b ; W out x in column-major */
static void dense(const float* W, const float* b, int nin, int nout, const float* x, float* y, int act) {
  for (int o = 0; o < nout; ++o) y[o] = 0.f;
  for (int k = 0; k < nin; ++k) {
    const float xk = x[k];
    const float* col = W + (size_t)k * nout;
    for (int o = 0; o < nout; ++o) y[o] += col[o] * xk;
  }
  for (int o = 0; o < nout; ++o) y[o] = act_f(y[o] + b[o], act);
}

static void chain(const float* w, const int* sizes, int n_sizes, int act, const float* x, float* out, float* t0,
                  float* t1) {
  const float* cur = x;
  size_t off = 0;
  for (int l = 0; l + 1 < n_sizes; ++l) {
    const int nin = sizes[l], nout = sizes[l + 1];
    const int last = (l + 2 == n_sizes);
    float* dst = last ? out : ((l & 1) ? t1 : t0);
    dense(w + off, w + off + (size_t)nin * nout, nin, nout, cur, dst, last ? 0 : act);
    off += (size_t)nin * nout + nout;
    cur = dst;
  }
}

static void box3(const float* a, float* o, int n) {
  o[0] = (a[0] + a[1]) * 0.5f;
  o[n - 1] = (a[n - 2] + a[n - 1]) * 0.5f;
  for (int i = 1; i < n - 1; ++i) o[i] = (a[i - 1] + a[i] + a[i + 1]) * (1.0f / 3.0f);
}

typedef struct {
  float y[3][MAXNZ], F[3][MAXNZ + 1], g[3][MAXNZ + 1], Ri[MAXNZ + 1], tmp[MAXNZ + 1], t0[MAXW], t1[MAXW];
} work_t;

static void rhs_col(const orc_params* p, const int* sizes, int n_sizes, int act, const float* w, size_t P,
                    const float* x, const float* bc, float t, float* k, work_t* wk) {
  const int Nz = p->nz;
  const float nzf = (float)Nz;
  const float *u = x, *v = x + Nz, *T = x + 2 * Nz;
  const float s_u = p->sigma[0], s_v = p->sigma[1], s_T = p->sigma[2];
  for (int n = 0; n < 3; ++n) {
    chain(w + n * P, sizes, n_sizes, act, x, wk->y[n], wk->t0, wk->t1);
    if (p->flags & F_SMOOTH_NN) { box3(wk->y[n], wk->tmp, Nz - 1); memcpy(wk->y[n], wk->tmp, sizeof(float) * (Nz - 1)); }
  }
  const float sub = (p->flags & F_BC_MINUS) ? 1.f : 0.f;
  for (int n = 0; n < 3; ++n) {
    const float off = sub * ((-p->mu[3 + n]) / p->sigma[3 + n]);
    wk->F[n][0] = bc[2 * n] - off;
    wk->F[n][Nz] = bc[2 * n + 1] - off;
    for (int j = 0; j < Nz - 1; ++j) wk->F[n][j + 1] = wk->y[n][j];
  }
  if (p->flags & F_DIURNAL) {
    const float omega = (float)(2.0 * 3.14159265358979323846 / (double)p->diurnal_period);
    const float phys = bc[6] * sinf(omega * (t * p->tau)) / (p->alpha * p->g);
    float top = (phys - p->mu[5]) / p->sigma[5];
    if (p->flags & F_DIURNAL_MINUS) top -= (-p->mu[5]) / p->sigma[5];
    wk->F[2][Nz] = top;
  }
  if (p->flags & F_MPP) {
    const float* c[3] = {u, v, T};
    for (int n = 0; n < 3; ++n) {
      wk->g[n][0] = 0.f; wk->g[n][Nz] = 0.f;
      for (int f = 1; f < Nz; ++f) wk->g[n][f] = (c[n][f] - c[n][f - 1]) * nzf;
    }
    const float e = (p->flags & F_RI_EPS) ? p->eps : 0.f;
    const float cBz = ((p->H * p->g) * p->alpha) * s_T;
    for (int f = 0; f <= Nz; ++f) {
      const float a = s_u * (wk->g[0][f] + e), b = s_v * (wk->g[1][f] + e);
      wk->Ri[f] = (cBz * (wk->g[2][f] + e)) / (a * a + b * b);
    }
    if (p->flags & F_SMOOTH_RI) { box3(wk->Ri, wk->tmp, Nz + 1); memcpy(wk->Ri, wk->tmp, sizeof(float) * (Nz + 1)); }
    const float c_u = s_u / p->sigma[3] / p->H, c_v = s_v / p->sigma[4] / p->H, c_T = s_T / p->sigma[5] / p->H;
    for (int f = 1; f < Nz; ++f) {
      const float nu = p->nu0 + p->nu_m * ((1.f - tanhf((wk->Ri[f] - p->Ric) / p->dRi)) / 2.f);
      float nuT = nu / p->Pr;
      if (p->flags & F_CA) {
        const float sel = (p->flags & F_CA_DUDZ) ? wk->g[0][f] : wk->g[2][f];
        nuT = sel > 0.f ? nu / p->Pr : p->kappa;
      }
      wk->F[0][f] -= c_u * nu * wk->g[0][f];
      wk->F[1][f] -= c_v * nu * wk->g[1][f];
      wk->F[2][f] -= c_T * nuT * wk->g[2][f];
    }
  }
  const float A = -p->tau / p->H, ft = p->f * p->tau;
  const float A_u = A * p->sigma[3] / s_u, A_v = A * p->sigma[4] / s_v, A_T = A * p->sigma[5] / s_T;
  const float cor_u = ft / s_u, cor_v = ft / s_v;
  for (int c = 0; c < Nz; ++c) {
    k[c] = A_u * ((wk->F[0][c + 1] - wk->F[0][c]) * nzf) + cor_u * (s_v * v[c] + p->mu[1]);
    k[Nz + c] = A_v * ((wk->F[1][c + 1] - wk->F[1][c]) * nzf) - cor_v * (s_u * u[c] + p->mu[0]);
    k[2 * Nz + c] = A_T * ((wk->F[2][c + 1] - wk->F[2][c]) * nzf);
  }
}

/* returns 0 on success; out[B][n_save][3nz] (save_every <= 0: final state only) */
int orc_rollout(const orc_params* p, const int* sizes, int n_sizes, int act, const float* w, const float* x0,
                const float* bc, int64_t B, int scheme, float dt, float t0, int n_steps, int save_every,
                float* out, int n_threads) {
  const int Nz = p->nz, S = 3 * Nz;
  if (Nz > MAXNZ) return -1;
  size_t P = 0;
  for (int l = 0; l + 1 < n_sizes; ++l) {
    if (sizes[l + 1] > MAXW) return -1;
    P += (size_t)sizes[l] * sizes[l + 1] + sizes[l + 1];
  }
  const int n_save = save_every > 0 ? n_steps / save_every + 1 : 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    work_t* wk = (work_t*)malloc(sizeof(work_t));
    float x[3 * MAXNZ], xs[3 * MAXNZ], k[3 * MAXNZ], acc[3 * MAXNZ];
#pragma omp for schedule(dynamic, 1)
    for (int64_t col = 0; col < B; ++col) {
      const float* b = bc + col * 8;
      memcpy(x, x0 + col * S, sizeof(float) * S);
      float* o = out + col * (size_t)n_save * S;
      if (save_every > 0) memcpy(o, x, sizeof(float) * S);
      for (int n = 0; n < n_steps; ++n) {
        const float t = t0 + (float)n * dt;
        if (scheme == 0) {
          rhs_col(p, sizes, n_sizes, act, w, P, x, b, t, k, wk);
          for (int i = 0; i < S; ++i) x[i] = x[i] + dt * k[i];
        } else if (scheme == 1) {
          rhs_col(p, sizes, n_sizes, act, w, P, x, b, t, k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = k[i]; xs[i] = x[i] + dt * k[i]; }
          rhs_col(p, sizes, n_sizes, act, w, P, xs, b, t + dt, k, wk);
          for (int i = 0; i < S; ++i) x[i] = x[i] + (dt * 0.5f) * (acc[i] + k[i]);
        } else {
          rhs_col(p, sizes, n_sizes, act, w, P, x, b, t, k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = k[i]; xs[i] = x[i] + (dt * 0.5f) * k[i]; }
          rhs_col(p, sizes, n_sizes, act, w, P, xs, b, t + dt * 0.5f, k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = acc[i] + 2.f * k[i]; xs[i] = x[i] + (dt * 0.5f) * k[i]; }
          rhs_col(p, sizes, n_sizes, act, w, P, xs, b, t + dt * 0.5f, k, wk);
          for (int i = 0; i < S; ++i) { acc[i] = acc[i] + 2.f * k[i]; xs[i] = x[i] + dt * k[i]; }
          rhs_col(p, sizes, n_sizes, act, w, P, xs, b, t + dt, k, wk);
          for (int i = 0; i < S; ++i) x[i] = x[i] + (dt / 6.f) * (acc[i] + k[i]);
        }
        if (save_every > 0 && (n + 1) % save_every == 0) memcpy(o + (size_t)((n + 1) / save_every) * S, x, sizeof(float) * S);
      }
      if (save_every <= 0) memcpy(o, x, sizeof(float) * S);
    }
    free(wk);
  }
  return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
