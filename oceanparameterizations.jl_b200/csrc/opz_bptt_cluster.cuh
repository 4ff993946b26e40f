// opz_bptt_cluster.cuh — the whole BPTT sweep of one group of columns inside ONE thread-block cluster.
// Included by opz_bptt.cu (namespace opz::bptt; uses Args, rec-free indexing q = n*stages + s).
//
// 16 CTAs form a cluster; CTA r keeps, for the WHOLE rollout, its 1/16 slice of every Dense layer of
// the three nets resident in shared memory in FP32 (construct_NN: 2.54 MB over 16 SMs = 158 KB each),
// so the weights are read from HBM/L2 exactly once per gradient step.  A right-hand side is then a
// sequence of slice-local matrix-vector products separated by exchanges through distributed shared
// memory (st.shared::cluster) and hardware cluster barriers:
//     forward   h_l slice -> all-gather (every CTA needs the full h_l for its slice of layer l+1)
//               last layer: split-K over the CTA's slice of the last hidden layer -> two-phase
//               all-reduce of the 3(Nz-1) outputs -> physics + Runge-Kutta update, replicated in every CTA
//     backward  physics VJP replicated; dz of the last hidden layer is slice-local;
//               dh_{l-1} = W_l^T dz_l: slice-local partial over ALL inputs -> reduce-scatter to the owners;
//               first layer: partial input cotangent -> two-phase all-reduce -> RK adjoint, replicated.
// Clusters are independent (a group of <= CG columns each), so there is no grid-wide synchronisation
// and no co-scheduling requirement.  Stage inputs, h_l and act'(z_l) -> dz_l are recorded in HBM for the
// weight-gradient reduction GEMMs (dw_kernel), exactly as in the graph path.
#pragma once

namespace cl {

constexpr int kCS = 16;         // CTAs per cluster
constexpr int kThreads = 512;

struct Layout {                 // float offsets into dynamic shared memory (identical in every CTA)
  int W[kMaxLayers];            // hidden l: [3][K_l][UL_l] ; last: [3][UL_{L-2}][nout]
  int bias[kMaxLayers];         // hidden l: [3][UL_l]      ; last: [3][nout]
  int UL[kMaxLayers];           // units per CTA of hidden layer l
  int hfull[kMaxLayers];        // gathered activations of hidden layer l <= L-3: [3][N_l][CG]
  int hs, dzs;                  // own slice of the last hidden layer / of the current dz: [3][ULmax][CG]
  int part;                     // split-K partials: kThreads * CG
  int y;                        // [max(3 nout, S3)][CG]: NN outputs (fwd), dy then the summed input cotangent (bwd)
  int dy;                       // [3 nout][CG]
  int xn, xs, kacc, lam, kbar, xsum, xphys;  // [S3][CG]
  int Fs;                       // [3][nz+1][CG]
  int recv_big, recv_big_size;  // [kCS][3 ULmax][CG] (x2 when L >= 4)
  int recv_small;               // [kCS][SL]
  int SL;
  int rbase;                    // 16 x uint32: shared::cluster base address of every rank
  int total;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ int g_dbg;
__device__ __forceinline__ void st_remote(uint32_t addr, float v) {
  if (g_dbg & 4) return;
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  if (g_dbg & 2) { __syncthreads(); return; }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__global__ void __launch_bounds__(kThreads, 1) sweep_kernel(const __grid_constant__ Args a, const __grid_constant__ Layout lay) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  const int rank = (int)cluster_rank();
  const int64_t group = blockIdx.x / kCS;
  const int64_t col0 = group * CG;
  const int ncol = (int)min((int64_t)CG, a.B - col0);
  const DevParams& P = a.P;
  const int nz = P.nz, S3 = a.S3, nout = a.nout, L = a.L, stages = a.stages;
  const int NO3 = 3 * nout;

  float* xn = sm + lay.xn;
  float* xs = sm + lay.xs;
  float* kacc = sm + lay.kacc;
  float* lam = sm + lay.lam;
  float* kbar = sm + lay.kbar;
  float* xsum = sm + lay.xsum;
  float* xphys = sm + lay.xphys;
  float* hs = sm + lay.hs;
  float* dzs = sm + lay.dzs;
  float* part = sm + lay.part;
  float* yb = sm + lay.y;
  float* dy = sm + lay.dy;
  float* Fs = sm + lay.Fs;
  float* recv_small = sm + lay.recv_small;
  uint32_t* rbase = reinterpret_cast<uint32_t*>(sm + lay.rbase);
  const uint32_t my_base = (uint32_t)__cvta_generic_to_shared(sm);

  // ---- one-time setup: remote bases, weight slices, initial state ----
  if (t < kCS) rbase[t] = map_to_rank(my_base, (uint32_t)t);
  for (int l = 0; l < L - 1; ++l) {
    const int K = a.nd.sizes[l], N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
    float* W = sm + lay.W[l];
    for (int e = t; e < 3 * K * UL; e += kThreads) {
      const int u = e % UL, k = (e / UL) % K, net = e / (UL * K);
      W[e] = (u0 + u < N) ? a.w[(int64_t)net * a.nd.P + a.nd.w_off[l] + (int64_t)k * N + u0 + u] : 0.0f;
    }
    float* bsl = sm + lay.bias[l];
    for (int e = t; e < 3 * UL; e += kThreads) {
      const int u = e % UL, net = e / UL;
      bsl[e] = (u0 + u < N) ? a.w[(int64_t)net * a.nd.P + a.nd.b_off[l] + u0 + u] : 0.0f;
    }
  }
  {
    const int l = L - 1, K = a.nd.sizes[l], UL = lay.UL[L - 2], u0 = rank * UL;  // K = width of the last hidden layer
    float* W = sm + lay.W[l];
    for (int e = t; e < 3 * UL * nout; e += kThreads) {
      const int j = e % nout, u = (e / nout) % UL, net = e / (nout * UL);
      W[e] = (u0 + u < K) ? a.w[(int64_t)net * a.nd.P + a.nd.w_off[l] + (int64_t)(u0 + u) * nout + j] : 0.0f;
    }
    float* bsl = sm + lay.bias[l];
    for (int e = t; e < NO3; e += kThreads) bsl[e] = a.w[(int64_t)(e / nout) * a.nd.P + a.nd.b_off[l] + e % nout];
  }
  for (int e = t; e < S3 * CG; e += kThreads) {
    const int i = e / CG, c = e % CG;
    const float v = c < ncol ? a.XN[(col0 + c) * S3 + i] : 0.0f;  // XN[0] was filled by the host
    xn[e] = v;
    xs[e] = v;
  }
  cluster_sync_all();  // rbase of every CTA is mapped; nobody writes remotely before everyone is here

  // two-phase all-reduce of `count` values held as per-CTA partials in registers of threads t < count:
  // reduce-scatter to owner = e % 16, the owner adds `bias_of(e)` and broadcasts into dst[e] of every CTA
  auto all_reduce = [&](float value, int count, float* dst, const float* bias_by_elem) {
    if (t < count) {
      const int owner = t % kCS, slot = t / kCS;
      st_remote(rbase[owner] + 4u * (uint32_t)(lay.recv_small + rank * lay.SL + slot), value);
    }
    cluster_sync_all();
    if (t < lay.SL) {
      const int e = t * kCS + rank;
      if (e < count) {
        float s = 0.0f;
#pragma unroll
        for (int src = 0; src < kCS; ++src) s += recv_small[src * lay.SL + t];
        if (bias_by_elem) s += bias_by_elem[e / CG];
        const uint32_t off = 4u * (uint32_t)((int)(dst - sm) + e);
#pragma unroll
        for (int d = 0; d < kCS; ++d) st_remote(rbase[d] + off, s);
      }
    }
    cluster_sync_all();
  };

  // hidden layer l, forward, for this CTA's slice; input activations `in` laid out [net][K][CG]
  // (net_stride 0 for the shared state input).  Leaves z in registers of threads t < 3 UL CG.
  auto slice_forward = [&](int l, const float* in, int net_stride) -> float {
    const int K = a.nd.sizes[l], UL = lay.UL[l], NP = 3 * UL;
    const int KSL = max(1, min(kThreads / NP, K));
    const float* W = sm + lay.W[l];
    if (t < NP * KSL) {
      const int p = t % NP, ks = t / NP, net = p / UL, u = p - net * UL;
      float acc[CG];
#pragma unroll
      for (int c = 0; c < CG; ++c) acc[c] = 0.0f;
      const float* Wp = W + (size_t)net * K * UL + u;
      const float* ip = in + (size_t)net * net_stride;
#pragma unroll 4
      for (int k = ks; k < K; k += KSL) {
        const float wv = Wp[k * UL];
#pragma unroll
        for (int c = 0; c < CG; ++c) acc[c] = fmaf(wv, ip[k * CG + c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < CG; ++c) part[(ks * NP + p) * CG + c] = acc[c];
    }
    __syncthreads();
    float z = 0.0f;
    if (t < NP * CG) {
      const int p = t / CG;
      for (int ks = 0; ks < KSL; ++ks) z += part[ks * NP * CG + t];
      z += (sm + lay.bias[l])[p];
    }
    return z;
  };

  // ================================ forward sweep ================================
  for (int n = 0; n < a.n_steps; ++n) {
    const float tn = a.t0 + (float)n * a.h;
    for (int s = 0; s < stages; ++s) {
      const int q = n * stages + s;
      const int64_t row0 = (int64_t)q * a.B + col0;
      if (rank == 0) {
        for (int e = t; e < S3 * CG; e += kThreads) {
          const int i = e / CG, c = e % CG;
          if (c < ncol) a.X[(row0 + c) * S3 + i] = xs[e];
        }
      }
      for (int l = 0; l < L - 1; ++l) {
        const int N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
        const float z = slice_forward(l, l == 0 ? xs : sm + lay.hfull[l - 1], l == 0 ? 0 : a.nd.sizes[l] * CG);
        if (t < 3 * UL * CG) {
          const int p = t / CG, c = t % CG, net = p / UL, u = p - net * UL;
          if (u0 + u < N) {
            const float hv = act_dyn(z, a.nd.act);
            if (c < ncol && !(g_dbg & 1)) {
              const int64_t o = ((int64_t)net * a.rows_cap + row0 + c) * N + u0 + u;
              a.H[l][o] = hv;
              a.G[l][o] = act_grad_dyn(z, a.nd.act);
            }
            if (l < L - 2) {
              const uint32_t off = 4u * (uint32_t)(lay.hfull[l] + (net * N + u0 + u) * CG + c);
#pragma unroll
              for (int d = 0; d < kCS; ++d) st_remote(rbase[d] + off, hv);
            } else {
              hs[t] = hv;
            }
          } else if (l == L - 2) {
            hs[t] = 0.0f;
          }
        }
        if (l < L - 2) cluster_sync_all();
        else __syncthreads();
      }
      // last layer: split-K over this CTA's slice of the last hidden layer, then all-reduce
      {
        const int UL = lay.UL[L - 2];
        const float* W = sm + lay.W[L - 1];
        float yp = 0.0f;
        if (t < NO3 * CG) {
          const int o = t / CG, c = t % CG, net = o / nout, j = o - net * nout;
          const float* Wp = W + (size_t)net * UL * nout + j;
          const float* hp = hs + net * UL * CG + c;
#pragma unroll 5
          for (int u = 0; u < UL; ++u) yp = fmaf(hp[u * CG], Wp[u * nout], yp);
        }
        all_reduce(yp, NO3 * CG, yb, sm + lay.bias[L - 1]);
      }
      // physics + Runge-Kutta update, replicated in every CTA (opz_device.cuh formulas, one thread per face / cell)
      const float ts = tn + a.rk_c[s] * a.h;
      if (t < (nz + 1) * CG) {
        const int f = t / CG, c = t % CG;
        const float* bc = a.bc + (col0 + min(c, ncol - 1)) * OPZ_BC_STRIDE;
        float Fu, Fv, FT;
        if (f == 0) { Fu = bc[0] - P.off[0]; Fv = bc[2] - P.off[1]; FT = bc[4] - P.off[2]; }
        else if (f == nz) { Fu = bc[1] - P.off[0]; Fv = bc[3] - P.off[1]; FT = wT_top_value(P, bc, ts); }
        else {
          Fu = yb[(f - 1) * CG + c]; Fv = yb[(nout + f - 1) * CG + c]; FT = yb[(2 * nout + f - 1) * CG + c];
          if (P.flags & OPZ_FLAG_MPP) {
            const float gu = (xs[f * CG + c] - xs[(f - 1) * CG + c]) * P.nzf;
            const float gv = (xs[(nz + f) * CG + c] - xs[(nz + f - 1) * CG + c]) * P.nzf;
            const float gT = (xs[(2 * nz + f) * CG + c] - xs[(2 * nz + f - 1) * CG + c]) * P.nzf;
            const float Ri = richardson(P, gu, gv, gT);
            const float nu = nu_of_ri(P, Ri);
            const float nuT = nuT_of(P, nu, gu, gT);
            Fu -= P.c_u * nu * gu;
            Fv -= P.c_v * nu * gv;
            FT -= P.c_T * nuT * gT;
          }
        }
        Fs[f * CG + c] = Fu; Fs[((nz + 1) + f) * CG + c] = Fv; Fs[(2 * (nz + 1) + f) * CG + c] = FT;
      }
      __syncthreads();
      if (t < S3 * CG) {
        const int i = t / CG, c = t % CG, v = i / nz, cz = i - v * nz;
        const float* F = Fs + v * (nz + 1) * CG + c;
        float k;
        if (v == 0) k = P.A_u * ((F[(cz + 1) * CG] - F[cz * CG]) * P.nzf) + P.cor_u * (P.s_v * xs[(nz + cz) * CG + c] + P.mu_v);
        else if (v == 1) k = P.A_v * ((F[(cz + 1) * CG] - F[cz * CG]) * P.nzf) - P.cor_v * (P.s_u * xs[cz * CG + c] + P.mu_u);
        else k = P.A_T * ((F[(cz + 1) * CG] - F[cz * CG]) * P.nzf);
        const float x0v = xn[t], hh = a.h;
        const bool last = (s == stages - 1);
        float xnext;
        if (a.scheme == OPZ_RK4) {
          if (s == 0) { kacc[t] = k; xnext = x0v + (hh * 0.5f) * k; }
          else if (s == 1) { kacc[t] = kacc[t] + 2.0f * k; xnext = x0v + (hh * 0.5f) * k; }
          else if (s == 2) { kacc[t] = kacc[t] + 2.0f * k; xnext = x0v + hh * k; }
          else xnext = x0v + (hh / 6.0f) * (kacc[t] + k);
        } else if (a.scheme == OPZ_RK2) {
          if (s == 0) { kacc[t] = k; xnext = x0v + hh * k; }
          else xnext = x0v + (hh * 0.5f) * (kacc[t] + k);
        } else {
          xnext = x0v + hh * k;
        }
        xs[t] = xnext;   // every thread touches only its own element of xs / xn here
        if (last) {
          xn[t] = xnext;
          if (rank == 0 && c < ncol) a.XN[((int64_t)(n + 1) * a.B + col0 + c) * S3 + i] = xnext;
        }
      }
      __syncthreads();
    }
  }

  // ================================ backward sweep ================================
  // dL/dx_N and the first tendency cotangent (replicated)
  auto loss_cotangent = [&](int n, int i, int c) -> float {
    const int v = i / nz, cz = i - v * nz;
    const int snap = n / a.save_every;
    const float* xv = a.XN + ((int64_t)n * a.B + col0 + c) * S3 + v * nz;
    const float* tg = a.target + (((col0 + c) * a.n_save + snap) * (int64_t)S3) + v * nz;
    const float inv_p = 1.0f / ((float)nz * (float)a.n_save * (float)a.B_total);
    const float inv_g = 1.0f / ((float)(nz + 1) * (float)a.n_save * (float)a.B_total);
    const float d0 = __ldcg(xv + cz) - tg[cz];
    float add = a.loss_w[v] * 2.0f * d0 * inv_p;
    const float sc = a.loss_w[3 + v] * 2.0f * P.nzf * inv_g;
    if (sc != 0.0f) {
      if (cz > 0) add += sc * ((d0 - (__ldcg(xv + cz - 1) - tg[cz - 1])) * P.nzf);
      if (cz < nz - 1) add -= sc * (((__ldcg(xv + cz + 1) - tg[cz + 1]) - d0) * P.nzf);
    }
    return add;
  };
  cluster_sync_all();  // rank 0's checkpoint stores are visible to the whole cluster
  if (t < S3 * CG) {
    const int i = t / CG, c = t % CG;
    const float l0 = c < ncol ? loss_cotangent(a.n_steps, i, c) : 0.0f;
    lam[t] = l0;
    kbar[t] = a.rk_b[stages - 1] * l0;
  }
  __syncthreads();

  int rs_parity = 0;
  for (int n = (g_dbg & 8) ? -1 : a.n_steps - 1; n >= 0; --n) {
    for (int s = stages - 1; s >= 0; --s) {
      const int q = n * stages + s;
      const int64_t row0 = (int64_t)q * a.B + col0;
      // prefetch this CTA's act'(z_l) slices (written by this CTA during the forward sweep)
      float greg[kMaxLayers];
#pragma unroll
      for (int l = 0; l < kMaxLayers; ++l) {
        greg[l] = 0.0f;
        if (l < L - 1) {
          const int N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
          if (t < 3 * UL * CG) {
            const int p = t / CG, c = t % CG, net = p / UL, u = p - net * UL;
            if (u0 + u < N && c < ncol) greg[l] = __ldcg(a.G[l] + ((int64_t)net * a.rows_cap + row0 + c) * N + u0 + u);
          }
        }
      }
      if (t < S3 * CG) {
        const int i = t / CG, c = t % CG;
        xs[t] = c < ncol ? __ldcg(a.X + (row0 + c) * S3 + i) : 0.0f;
      }
      __syncthreads();
      // ---- physics VJP (replicated) ----
      if (t < (nz + 1) * CG) {
        const int f = t / CG, c = t % CG;
        float gbu = 0.0f, gbv = 0.0f, gbT = 0.0f;
        if (f > 0 && f < nz) {
          const float Fbu = P.A_u * P.nzf * (kbar[(f - 1) * CG + c] - kbar[f * CG + c]);
          const float Fbv = P.A_v * P.nzf * (kbar[(nz + f - 1) * CG + c] - kbar[(nz + f) * CG + c]);
          const float FbT = P.A_T * P.nzf * (kbar[(2 * nz + f - 1) * CG + c] - kbar[(2 * nz + f) * CG + c]);
          dy[(f - 1) * CG + c] = Fbu; dy[(nout + f - 1) * CG + c] = Fbv; dy[(2 * nout + f - 1) * CG + c] = FbT;
          if (rank == 0 && c < ncol) {
            a.DY[((int64_t)0 * a.rows_cap + row0 + c) * nout + f - 1] = Fbu;
            a.DY[((int64_t)1 * a.rows_cap + row0 + c) * nout + f - 1] = Fbv;
            a.DY[((int64_t)2 * a.rows_cap + row0 + c) * nout + f - 1] = FbT;
          }
          if (P.flags & OPZ_FLAG_MPP) {
            const float gu = (xs[f * CG + c] - xs[(f - 1) * CG + c]) * P.nzf;
            const float gv = (xs[(nz + f) * CG + c] - xs[(nz + f - 1) * CG + c]) * P.nzf;
            const float gT = (xs[(2 * nz + f) * CG + c] - xs[(2 * nz + f - 1) * CG + c]) * P.nzf;
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
          }
        }
        Fs[f * CG + c] = gbu; Fs[((nz + 1) + f) * CG + c] = gbv; Fs[(2 * (nz + 1) + f) * CG + c] = gbT;
      }
      __syncthreads();
      if (t < S3 * CG) {
        const int i = t / CG, c = t % CG, v = i / nz, cz = i - v * nz;
        const float* g = Fs + v * (nz + 1) * CG + c;
        float xb = P.nzf * (g[cz * CG] - g[(cz + 1) * CG]);
        if (v == 0) xb -= P.cor_v * P.s_u * kbar[(nz + cz) * CG + c];
        else if (v == 1) xb += P.cor_u * P.s_v * kbar[cz * CG + c];
        xphys[t] = xb;
      }
      // ---- dz of the last hidden layer: slice-local ----
      {
        const int l = L - 2, N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
        const float* W = sm + lay.W[L - 1];
        if (t < 3 * UL * CG) {
          const int p = t / CG, c = t % CG, net = p / UL, u = p - net * UL;
          const float* Wp = W + (size_t)(net * UL + u) * nout;
          const float* dp = dy + net * nout * CG + c;
          float sacc = 0.0f;
#pragma unroll 4
          for (int j = 0; j < nout; ++j) sacc = fmaf(Wp[j], dp[j * CG], sacc);
          const float dz = sacc * greg[l];
          dzs[t] = dz;
          if (u0 + u < N && c < ncol) a.G[l][((int64_t)net * a.rows_cap + row0 + c) * N + u0 + u] = dz;
        }
      }
      __syncthreads();
      // ---- hidden layers: partial dh_{l-1} over all inputs of layer l -> reduce-scatter to the owners ----
      for (int l = L - 2; l >= 1; --l) {
        const int K = a.nd.sizes[l], UL = lay.UL[l];        // layer l: K inputs (= width of hidden l-1)
        const int ULp = lay.UL[l - 1], u0p = rank * ULp;     // slicing of hidden layer l-1
        const float* W = sm + lay.W[l];
        float* rb = sm + lay.recv_big + rs_parity * lay.recv_big_size;
        for (int it = t; it < 3 * K; it += kThreads) {
          const int net = it / K, k = it - net * K;
          const float* Wp = W + ((size_t)net * K + k) * UL;
          const float* dp = dzs + net * UL * CG;
          float acc[CG];
#pragma unroll
          for (int c = 0; c < CG; ++c) acc[c] = 0.0f;
#pragma unroll 5
          for (int u = 0; u < UL; ++u) {
            const float wv = Wp[u];
#pragma unroll
            for (int c = 0; c < CG; ++c) acc[c] = fmaf(wv, dp[u * CG + c], acc[c]);
          }
          const int owner = k / ULp, kl = k - owner * ULp;
          const uint32_t off = 4u * (uint32_t)((int)(rb - sm) + ((rank * 3 + net) * ULp + kl) * CG);
#pragma unroll
          for (int c = 0; c < CG; ++c) st_remote(rbase[owner] + off + 4u * c, acc[c]);
        }
        cluster_sync_all();
        if (t < 3 * ULp * CG) {
          const int p = t / CG, c = t % CG, net = p / ULp, u = p - net * ULp;
          float sacc = 0.0f;
#pragma unroll
          for (int src = 0; src < kCS; ++src) sacc += rb[(src * 3 * ULp) * CG + t];
          const float dz = sacc * greg[l - 1];
          dzs[t] = dz;
          const int N = a.nd.sizes[l];
          if (u0p + u < N && c < ncol) a.G[l - 1][((int64_t)net * a.rows_cap + row0 + c) * N + u0p + u] = dz;
        }
        if (L >= 4) rs_parity ^= 1;
        __syncthreads();
      }
      // ---- first layer: partial input cotangent over this CTA's units, all-reduce ----
      {
        const int UL = lay.UL[0];
        const float* W = sm + lay.W[0];
        float xp = 0.0f;
        if (t < S3 * CG) {
          const int i = t / CG, c = t % CG;
          for (int net = 0; net < 3; ++net) {
            const float* Wp = W + ((size_t)net * S3 + i) * UL;
            const float* dp = dzs + net * UL * CG + c;
#pragma unroll 5
            for (int u = 0; u < UL; ++u) xp = fmaf(Wp[u], dp[u * CG], xp);
          }
        }
        all_reduce(xp, S3 * CG, yb, nullptr);
      }
      // ---- Runge-Kutta adjoint bookkeeping (replicated) ----
      if (t < S3 * CG) {
        const int i = t / CG, c = t % CG;
        const float xbar = yb[t] + xphys[t];
        const float lv = lam[t];
        const float xsv = (s == stages - 1 ? 0.0f : xsum[t]) + xbar;
        if (s > 0) {
          kbar[t] = a.rk_b[s - 1] * lv + a.rk_a[s - 1] * xbar;
          xsum[t] = xsv;
        } else {
          float ln = lv + xsv;
          if (n % a.save_every == 0 && c < ncol) ln += loss_cotangent(n, i, c);
          lam[t] = ln;
          kbar[t] = a.rk_b[stages - 1] * ln;
        }
      }
      __syncthreads();
    }
  }
  cluster_sync_all();  // no CTA exits while a peer may still address its shared memory
}

}  // namespace cl
