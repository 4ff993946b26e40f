// opz_simt.cuh — FP32 CUDA-core batched Dense layer for a tile of M ocean columns.
//
// This is the parity path (OPZ_PREC_FP32): out[n][m] = act(sum_k W[k][n] * in[k][m] + b[n]).
// Activations live in shared memory FEATURE-MAJOR ([feature][column], column contiguous) so that
//   * the Flux weight matrix (out x in, column-major == [k][n] with n contiguous) is consumed
//     as stored — no transposition anywhere (NDE_training.jl:11-13 layout),
//   * each thread owns a TM x TN = 4 columns x 8 units register tile (32 FMA per 3 LDS.128),
//   * the W k-slabs stream global/L2 -> registers -> shared, double buffered, one
//     __syncthreads per slab.
// Narrow layers (the 31-unit output layer) are split over K across the idle unit-groups and
// reduced through the (then dead) staging buffer.
#pragma once
#include "opz_device.cuh"

namespace opz {

constexpr int kThreads = 256;
constexpr int kTM = 4;
constexpr int kTN = 8;
constexpr int kKC = 16;  // k-rows per staged slab (per K-slice)

template <int M>
struct SimtCfg {
  static constexpr int MG = M / kTM;          // column groups
  static constexpr int NG = kThreads / MG;    // unit groups per pass
  static constexpr int NP = NG * kTN;         // units per pass
  static constexpr int STAGE_FLOATS = kKC * NP;  // one staging buffer
  // two staging buffers, also reused as the split-K reduction scratch [KS][NP/KS][M] = NP*M floats
  static constexpr int STAGE_TOTAL = (2 * STAGE_FLOATS > NP * M) ? 2 * STAGE_FLOATS : NP * M;
};

__device__ __forceinline__ int ks_for(int N, int NG) {
  const int need = (N + kTN - 1) / kTN;
  int ks = 1;
  while (ks * 2 * need <= NG) ks *= 2;
  return ks;
}

// in  : shared [K][M]        out : shared [N][M]
// W   : global, element (k, n) at W[k*N + n];  b : global [N]
// stage : shared, SimtCfg<M>::STAGE_TOTAL floats
template <int M>
__device__ void dense_layer_fp32(const float* __restrict__ W, const float* __restrict__ b, int K, int N,
                                 int act, const float* __restrict__ in, float* __restrict__ out,
                                 float* __restrict__ stage) {
  using C = SimtCfg<M>;
  const int tid = threadIdx.x;
  const int mg = tid % C::MG;
  const int ng = tid / C::MG;
  const int KS = ks_for(N, C::NG);        // K-slices (power of two)
  const int NGs = C::NG / KS;             // unit groups per slice
  const int NPs = NGs * kTN;              // units per pass
  const int ks = ng / NGs;
  const int ngl = ng % NGs;
  const int rows = kKC * KS;              // k-rows per staged slab
  const int n_slabs = (K + rows - 1) / rows;
  const int n_pass = (N + NPs - 1) / NPs;
  constexpr int LD = C::STAGE_FLOATS / kThreads;  // floats each thread stages per slab (16)

  for (int pass = 0; pass < n_pass; ++pass) {
    const int nbase = pass * NPs;
    float acc[kTN][kTM];
#pragma unroll
    for (int j = 0; j < kTN; ++j)
#pragma unroll
      for (int i = 0; i < kTM; ++i) acc[j][i] = 0.0f;

    float regs[LD];
    // slab element e -> (row r = e / NPs, unit nn = e % NPs); thread stages e = tid + i*256
    auto fetch = [&](int slab) {
      const int k0 = slab * rows;
#pragma unroll
      for (int i = 0; i < LD; ++i) {
        const int e = tid + i * kThreads;
        const int r = e / NPs, nn = e - r * NPs;
        const int k = k0 + r, n = nbase + nn;
        regs[i] = (k < K && n < N) ? __ldg(W + (size_t)k * N + n) : 0.0f;
      }
    };
    auto commit = [&](int buf) {
      float* s = stage + buf * C::STAGE_FLOATS;
#pragma unroll
      for (int i = 0; i < LD; ++i) s[tid + i * kThreads] = regs[i];
    };
    fetch(0);
    commit(0);
    __syncthreads();
    for (int slab = 0; slab < n_slabs; ++slab) {
      if (slab + 1 < n_slabs) fetch(slab + 1);
      const float* s = stage + (slab & 1) * C::STAGE_FLOATS + (ks * kKC) * NPs + ngl * kTN;
      const int kbase = slab * rows + ks * kKC;
      const float* a = in + (size_t)kbase * M + mg * kTM;
      int kmax = K - kbase;
      kmax = kmax < 0 ? 0 : (kmax > kKC ? kKC : kmax);
      if (kmax == kKC) {
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
          const float4 av = *reinterpret_cast<const float4*>(a + kk * M);
          const float4 w0 = *reinterpret_cast<const float4*>(s + kk * NPs);
          const float4 w1 = *reinterpret_cast<const float4*>(s + kk * NPs + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float ax[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
          for (int j = 0; j < kTN; ++j)
#pragma unroll
            for (int i = 0; i < kTM; ++i) acc[j][i] = fmaf(wv[j], ax[i], acc[j][i]);
        }
      } else {
        for (int kk = 0; kk < kmax; ++kk) {
          const float4 av = *reinterpret_cast<const float4*>(a + kk * M);
          const float4 w0 = *reinterpret_cast<const float4*>(s + kk * NPs);
          const float4 w1 = *reinterpret_cast<const float4*>(s + kk * NPs + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float ax[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
          for (int j = 0; j < kTN; ++j)
#pragma unroll
            for (int i = 0; i < kTM; ++i) acc[j][i] = fmaf(wv[j], ax[i], acc[j][i]);
        }
      }
      if (slab + 1 < n_slabs) commit((slab + 1) & 1);
      __syncthreads();
    }
    // ---- epilogue ----
    if (KS == 1) {
#pragma unroll
      for (int j = 0; j < kTN; ++j) {
        const int n = nbase + ngl * kTN + j;
        if (n < N) {
          const float bb = __ldg(b + n);
          float4 o;
          o.x = act_dyn(acc[j][0] + bb, act);
          o.y = act_dyn(acc[j][1] + bb, act);
          o.z = act_dyn(acc[j][2] + bb, act);
          o.w = act_dyn(acc[j][3] + bb, act);
          *reinterpret_cast<float4*>(out + (size_t)n * M + mg * kTM) = o;
        }
      }
      // `out` is not read and `stage` not rewritten before the next __syncthreads of the caller
      // or of the next pass's first commit+sync; passes write disjoint rows of `out`.
      __syncthreads();
    } else {
      // partial sums -> stage[ks][nn][m] (KS*NPs*M = NP*M floats <= STAGE_TOTAL)
      float* red = stage;
#pragma unroll
      for (int j = 0; j < kTN; ++j) {
        const int nn = ngl * kTN + j;
        float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        *reinterpret_cast<float4*>(red + ((size_t)ks * NPs + nn) * M + mg * kTM) = o;
      }
      __syncthreads();
      for (int e = tid; e < NPs * M; e += kThreads) {
        const int nn = e / M, m = e - nn * M;
        const int n = nbase + nn;
        if (n < N) {
          float sum = red[(size_t)nn * M + m];
          for (int s2 = 1; s2 < KS; ++s2) sum += red[((size_t)s2 * NPs + nn) * M + m];
          out[(size_t)n * M + m] = act_dyn(sum + __ldg(b + n), act);
        }
      }
      __syncthreads();
    }
  }
}

// Runs one net's Chain: in (shared [sizes[0]][M]) -> y (shared [sizes[L]][M]).
// bufA/bufB: shared ping-pong buffers of at least max_hidden*M floats each.
template <int M>
__device__ void mlp_chain_fp32(const NetDesc& nd, const float* __restrict__ wnet, const float* __restrict__ in,
                               float* bufA, float* bufB, float* y, float* stage) {
  const float* cur = in;
  for (int l = 0; l < nd.n_layers; ++l) {
    const bool last = (l == nd.n_layers - 1);
    float* dst = last ? y : ((l & 1) ? bufB : bufA);
    dense_layer_fp32<M>(wnet + nd.w_off[l], wnet + nd.b_off[l], nd.sizes[l], nd.sizes[l + 1],
                        last ? OPZ_ACT_IDENTITY : nd.act, cur, dst, stage);
    cur = dst;
  }
}

}  // namespace opz
