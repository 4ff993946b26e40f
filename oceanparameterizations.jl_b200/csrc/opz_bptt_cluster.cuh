// opz_bptt_cluster.cuh — the whole BPTT sweep of one group of columns inside ONE thread-block cluster.
// Included by opz_bptt.cu (namespace opz::bptt; uses Args; record index q = n*stages + s).
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
// Clusters are independent (a group of <= CP columns each), so there is no grid-wide synchronisation
// and no co-scheduling requirement.  Stage inputs, h_l and act'(z_l) -> dz_l are recorded in HBM for the
// weight-gradient reduction GEMMs (dw_kernel), exactly as in the graph path.
//
// Latency notes (the sweep is a chain of ~6 exchanges per right-hand side): activations are laid out
// [feature][CP] with CP = 1, 2 or 4 column slots so that one broadcast LDS.32/64/128 feeds CP FMAs;
// record stores to HBM are issued between barrier.cluster.arrive and .wait so that no barrier's release
// has to drain them; the backward sweep prefetches the next right-hand side's records one step ahead;
// the CTA that writes the replicated records rotates with the record index.
#pragma once

namespace cl {

constexpr int kCS = 16;         // CTAs per cluster for nets that do not fit one SM's shared memory (CS = 1 otherwise)
constexpr int kThreads = 512;
constexpr int kMaxKSL = 8;      // split of a slice's reduction dimension over threads

struct Layout {                 // float offsets into dynamic shared memory (identical in every CTA)
  int W[kMaxLayers];            // hidden l: [3][K_l][UL_l] ; last: [3][UL_{L-2}][nout]
  int bias[kMaxLayers];         // hidden l: [3][UL_l]      ; last: [3][nout]
  int UL[kMaxLayers];           // units per CTA of hidden layer l
  int hfull[kMaxLayers];        // gathered activations of hidden layer l <= L-3: [3][N_l][CP]
  int gs[kMaxLayers];           // this CTA's act'(z_l) slice of the right-hand side being back-propagated: [3][UL_l][CP]
  int hs, dzs;                  // own slice of the last hidden layer / of the current dz: [3][ULmax][CP]
  int part;                     // split-K partials: kThreads * CP
  int y;                        // [max(3 nout, S3)][CP]: NN outputs (fwd), summed input cotangent (bwd)
  int dy;                       // [3 nout][CP]
  int xn, xs, kacc, lam, kbar, xsum, xphys;  // [S3][CP]
  int Fs;                       // [3][nz+1][CP]
  int recv_big, recv_big_size;  // [kCS][3 ULmax][CP] (x2 when L >= 4)
  int recv_small;               // [kCS][SL]
  int SL;
  int rbase;                    // 16 x uint32: shared::cluster base address of every rank
  int mbar;                     // kNumXB x uint64 (8-byte aligned): one mbarrier per exchange channel, see XB_*
  int total;
};

// Exchange channels (CS > 1).  Every exchange of the sweep is data written into a peer's shared memory with
// st.async ... mbarrier::complete_tx::bytes: the store itself signals the RECEIVER's mbarrier, and a CTA proceeds as soon as the
// bytes IT expects have landed.  No cluster-wide barrier, no release fence (barrier.cluster.arrive.release costs MEMBAR.ALL.GPU,
// which also had to drain the HBM record stores): an exchange costs the DSMEM transit (~215 cycles) plus its bytes.
// Single-buffered receive areas are safe because every reuse of a buffer lies a full all-to-all dependency later: a peer can
// only be one exchange ahead, and it gets there only with data this CTA produced AFTER its last read of the buffer.
enum : int { XB_GATHER = 0,                 // [kMaxLayers] all-gather of h_l (forward)
             XB_AR1F = XB_GATHER + kMaxLayers, XB_AR2F,   // forward all-reduce of the NN outputs: reduce-scatter / broadcast
             XB_AR1B, XB_AR2B,              // backward all-reduce of the input cotangent
             XB_RS,                         // [2] reduce-scatter of W^T dz (backward), double-buffered when L >= 4
             kNumXB = XB_RS + 2 };

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
__device__ __forceinline__ void st_remote(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// CP consecutive floats in one transaction (the address is CP * 4 bytes aligned: every slot index is a multiple of CP)
template <int CP>
__device__ __forceinline__ void st_remote_cols(uint32_t addr, const float (&v)[CP]) {
  if constexpr (CP == 2) asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v[0]), "f"(v[1]) : "memory");
  else if constexpr (CP == 4)
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  else {
#pragma unroll
    for (int c = 0; c < CP; ++c) st_remote(addr + 4u * c, v[c]);
  }
}
// store + signal: 4 (or 4 CP) bytes into a peer's shared memory, counted on the peer's mbarrier `bar` (both shared::cluster addresses)
__device__ __forceinline__ void st_async(uint32_t addr, float v, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(__float_as_uint(v)), "r"(bar) : "memory");
}
template <int CP>
__device__ __forceinline__ void st_async_cols(uint32_t addr, const float (&v)[CP], uint32_t bar) {
  if constexpr (CP == 2)
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(addr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(bar) : "memory");
  else if constexpr (CP == 4)
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(bar) : "memory");
  else {
#pragma unroll
    for (int c = 0; c < CP; ++c) st_async(addr + 4u * c, v[c], bar);
  }
}
__device__ __forceinline__ void xb_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); }
// the receiver's one arrival of an exchange instance, carrying the bytes it expects (may come after the first bytes have landed)
__device__ __forceinline__ void xb_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xb_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// CS = 1 (the whole net in one CTA): the "exchanges" are local shared-memory stores and the barrier is __syncthreads — none of the
// MEMBAR.ALL.GPU / CCTL.IVALL that the release / acquire cluster barrier costs
template <int CS>
__device__ __forceinline__ void cluster_arrive() {
  if constexpr (CS > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
template <int CS>
__device__ __forceinline__ void cluster_wait() {
  if constexpr (CS > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
}
template <int CS>
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive<CS>(); cluster_wait<CS>(); }

// acc[0..CP) += w * p[0..CP)  with one broadcast vector load
template <int CP>
__device__ __forceinline__ void fma_cols(float (&acc)[CP], float w, const float* p) {
  if constexpr (CP == 1) {
    acc[0] = fmaf(w, p[0], acc[0]);
  } else if constexpr (CP == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
  } else {
    const float4 v = *reinterpret_cast<const float4*>(p);
    acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
    acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
  }
}

// v[0..CP) = p[0..CP) with one vector load
template <int CP>
__device__ __forceinline__ void load_cols(float (&v)[CP], const float* p) {
  if constexpr (CP == 1) v[0] = p[0];
  else if constexpr (CP == 2) { const float2 q = *reinterpret_cast<const float2*>(p); v[0] = q.x; v[1] = q.y; }
  else { const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
}
template <int CP>
__device__ __forceinline__ void store_cols(float* p, const float (&v)[CP]) {
  if constexpr (CP == 1) p[0] = v[0];
  else if constexpr (CP == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  else *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
constexpr int kRB = 5;   // register blocking of the slice mat-vecs: one vector load of the CP activation columns feeds kRB x CP FMAs.
                         // These loops are bound by shared-memory instruction passes (an LDS.128 costs 4 even as a broadcast, an
                         // LDS.32 one): 1 + 4 passes per CP FMAs unblocked, kRB + 4 per kRB x CP blocked.

// per-phase cycle counters of CTA 0 (thread 0), dumped by the host when OPZ_BPTT_PROFILE is set
__device__ unsigned long long g_phase_cycles[24];
#define OPZ_PH(i)                                                     \
  do {                                                                \
    if (timing) { const long long now_ = clock64(); ph[i] += now_ - tlast; tlast = now_; } \
  } while (0)

template <int CP, int CS>
__global__ void __launch_bounds__(kThreads, 1) sweep_kernel(const __grid_constant__ Args a, const __grid_constant__ Layout lay, int cg) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  const int tc = t % CP, tq = t / CP;   // column slot / item index of the "one thread per (item, column)" phases
  const int rank = CS > 1 ? (int)cluster_rank() : 0;
  const int64_t group = blockIdx.x / CS;
  const int64_t col0 = group * cg;
  const int ncol = (int)min((int64_t)cg, a.B - col0);
  const bool col_ok = tc < ncol;
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
  const bool timing = (blockIdx.x == 1 && t == 0);
  long long ph[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) ph[i] = 0;
  long long tlast = clock64();

  // ---- one-time setup: remote bases, exchange barriers, weight slices, initial state ----
  if (t < CS) rbase[t] = map_to_rank(my_base, (uint32_t)t);
  const uint32_t xb0 = my_base + 4u * (uint32_t)lay.mbar;     // this CTA's exchange barriers (shared::cta address)
  const uint32_t xb_off = 4u * (uint32_t)lay.mbar;            // ... and their offset from a peer's base
  uint32_t xb_par = 0;                                         // parity of the next completion of each channel (bit = channel)
  // bytes this CTA receives per instance of each channel
  auto share_of = [&](int count, int r) { return count > r ? (count - r + CS - 1) / CS : 0; };   // entries e < count with e % CS == r
  auto xb_bytes = [&](int ch) -> uint32_t {
    if (ch < XB_GATHER + kMaxLayers) { const int l = ch - XB_GATHER; return l <= L - 3 ? 12u * (uint32_t)(a.nd.sizes[l + 1] * CP) : 0u; }
    if (ch == XB_AR1F) return 4u * (uint32_t)(CS * share_of(NO3 * CP, rank));
    if (ch == XB_AR2F) return 4u * (uint32_t)(NO3 * CP);
    if (ch == XB_AR1B) return 4u * (uint32_t)(CS * share_of(S3 * CP, rank));
    if (ch == XB_AR2B) return 4u * (uint32_t)(S3 * CP);
    return 0u;   // XB_RS: depends on the layer, posted per use
  };
  auto rs_bytes = [&](int l) -> uint32_t {   // reduce-scatter of layer l's input cotangent: this CTA owns ULp units of hidden layer l - 1
    const int K = a.nd.sizes[l], ULp = lay.UL[l - 1];
    const int cnt = max(0, min(ULp, K - rank * ULp));
    return 4u * (uint32_t)(CS * 3 * cnt * CP);
  };
  if constexpr (CS > 1) {
    if (t == 0) {
      for (int ch = 0; ch < kNumXB; ++ch) xb_init(xb0 + 8u * ch);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      for (int ch = 0; ch < XB_RS; ++ch) xb_expect(xb0 + 8u * ch, xb_bytes(ch));
    }
  }
  // wait for the current instance of channel ch, then post the expectation of the next one (same byte count every time)
  auto xb_wait_repost = [&](int ch, uint32_t next_bytes) {
    if constexpr (CS > 1) {
      xb_wait(xb0 + 8u * ch, (xb_par >> ch) & 1u);
      xb_par ^= 1u << ch;
      __syncthreads();   // the barrier this replaces also ordered the CTA's own threads (reuse of part / hs / dzs ...)
      if (t == 0) xb_expect(xb0 + 8u * ch, next_bytes);
    } else {
      __syncthreads();
    }
  };
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
  if (t < S3 * CP) {
    const float v = col_ok ? a.XN[(col0 + tc) * S3 + tq] : 0.0f;  // XN[0] was filled by the host
    xn[t] = v;
    xs[t] = v;
  }
  cluster_sync_all<CS>();  // rbase of every CTA is mapped; nobody writes remotely before everyone is here

  // two-phase all-reduce of `count` (<= kThreads) per-CTA partials held by threads t < count:
  // reduce-scatter to owner = e % 16; the owner adds bias[e / CP] and broadcasts into dst[e] of every CTA.
  // `between()` runs between the first arrive and wait (HBM record stores go there).
  auto all_reduce = [&](float value, int count, float* dst, const float* bias_by_item, int ch1, int ch2, auto&& between) {
    if constexpr (CS > 1) {
      if (t < count) {
        const int owner = t % CS, slot = t / CS;
        st_async(rbase[owner] + 4u * (uint32_t)(lay.recv_small + rank * lay.SL + slot), value, rbase[owner] + xb_off + 8u * ch1);
      }
      between();
      xb_wait_repost(ch1, xb_bytes(ch1));
      if (t < lay.SL * CS) {   // CS threads per owned entry: each forms the sum and sends it to ONE peer
        const int sl = t / CS, d = t % CS;
        const int e = sl * CS + rank;
        if (e < count) {
          float s = 0.0f;
#pragma unroll
          for (int src = 0; src < CS; ++src) s += recv_small[src * lay.SL + sl];
          if (bias_by_item) s += bias_by_item[e / CP];
          st_async(rbase[d] + 4u * (uint32_t)((int)(dst - sm) + e), s, rbase[d] + xb_off + 8u * ch2);
        }
      }
      xb_wait_repost(ch2, xb_bytes(ch2));
    } else {
      if (t < count) recv_small[t] = value;   // (SL = count when CS = 1)
      between();
      __syncthreads();
      if (t < count) dst[t] = recv_small[t] + (bias_by_item ? bias_by_item[t / CP] : 0.0f);
      __syncthreads();
    }
  };

  // hidden layer l, forward, this CTA's slice; `in` is laid out [net][K][CP] (net_stride 0 for the shared
  // state input).  Returns z for threads t < 3 UL CP (item p = t / CP = net*UL + u, column slot t % CP).
  auto slice_forward = [&](int l, const float* in, int net_stride) -> float {
    const int K = a.nd.sizes[l], UL = lay.UL[l], NP = 3 * UL;
    const float* W = sm + lay.W[l];
    int KSL;
    constexpr int kKS = 32, kPL = 16;   // k-split of the register-blocked form = one warp per (net, unit group); partial-sum planes kept
    if (CS > 1 && UL % kRB == 0 && 3 * (UL / kRB) * kKS <= kThreads && kPL * NP * CP <= lay.recv_big_size) {
      // Register-blocked: one thread = kRB consecutive units of one net x all CP columns over every 32nd input; the 32 lanes of
      // a warp are the 32 k-splits of ONE unit group, so the weights of a load are UL words apart (UL odd: 32 distinct banks —
      // with two unit groups per warp the half-warps collided, 1.9 wavefronts per load in the ncu source view) and the input
      // columns of 32 consecutive k are one contiguous vector load.  Lanes 16..31 hand their partial sums to lanes 0..15
      // (one shuffle each); the 16 planes left go through the (forward-idle) reduce-scatter buffer.
      const int gpn = UL / kRB;
      float* scratch = sm + lay.recv_big;
      const int grp = t / kKS, ks = t % kKS;
      if (grp < 3 * gpn) {   // whole warps
        const int net = grp / gpn, ug = (grp - net * gpn) * kRB;
        float acc[kRB][CP];
#pragma unroll
        for (int j = 0; j < kRB; ++j)
#pragma unroll
          for (int c = 0; c < CP; ++c) acc[j][c] = 0.0f;
        const float* Wp = W + (size_t)net * K * UL + ug;
        const float* ip = in + (size_t)net * net_stride;
#pragma unroll 4
        for (int k = ks; k < K; k += kKS) {
          float xv[CP];
          load_cols<CP>(xv, ip + k * CP);
#pragma unroll
          for (int j = 0; j < kRB; ++j) {
            const float w = Wp[k * UL + j];
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[j][c] = fmaf(w, xv[c], acc[j][c]);
          }
        }
        OPZ_PH(17);
#pragma unroll
        for (int j = 0; j < kRB; ++j)
#pragma unroll
          for (int c = 0; c < CP; ++c) acc[j][c] += __shfl_down_sync(0xffffffffu, acc[j][c], kPL);
        if (ks < kPL) {
          float* sp = scratch + (ks * NP + net * UL + ug) * CP;   // vector stores: 16 B slots of a quarter warp (ks = 0..7) are 1200 B apart: conflict-free
#pragma unroll
          for (int j = 0; j < kRB; ++j) store_cols<CP>(sp + j * CP, acc[j]);
        }
      }
      OPZ_PH(18);
      __syncthreads();
      OPZ_PH(19);
      float z = 0.0f;
      if (t < NP * CP) {
#pragma unroll
        for (int ks2 = 0; ks2 < kPL; ++ks2) z += scratch[ks2 * NP * CP + t];
        z += (sm + lay.bias[l])[tq];
      }
      OPZ_PH(20);
      return z;
    }
    if (UL <= 32) {
      // One (net, k-split) task per group of LT lanes, lane = unit: the LT weights of a k are consecutive words (one conflict-free
      // wavefront) and the CP input columns of that k are ONE broadcast address for the whole group.  (With items laid over
      // threads modulo NP a warp straddled nets and k-splits: 2-way conflicts on the weights and 2 - 3 distinct input
      // addresses per load; the slice mat-vec is bound by shared-memory wavefronts.)
      const int LT = UL <= 8 ? 8 : (UL <= 16 ? 16 : 32);
      const int n_tasks = kThreads / LT;
      KSL = max(1, min(min(n_tasks / 3, K), kMaxKSL));
      const int task = t / LT, u = t % LT;
      if (task < 3 * KSL && u < UL) {
        const int net = task / KSL, ks = task - net * KSL;
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
        const float* Wp = W + (size_t)net * K * UL + u;
        const float* ip = in + (size_t)net * net_stride;
#pragma unroll 8
        for (int k = ks; k < K; k += KSL) fma_cols<CP>(acc, Wp[k * UL], ip + k * CP);
#pragma unroll
        for (int c = 0; c < CP; ++c) part[(ks * NP + net * UL + u) * CP + c] = acc[c];
      }
    } else {
      KSL = max(1, min(min(kThreads / NP, K), kMaxKSL));
      if (t < NP * KSL) {
        const int p = t % NP, ks = t / NP, net = p / UL, u = p - net * UL;
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
        const float* Wp = W + (size_t)net * K * UL + u;
        const float* ip = in + (size_t)net * net_stride;
#pragma unroll 8
        for (int k = ks; k < K; k += KSL) fma_cols<CP>(acc, Wp[k * UL], ip + k * CP);
#pragma unroll
        for (int c = 0; c < CP; ++c) part[(ks * NP + p) * CP + c] = acc[c];
      }
    }
    __syncthreads();
    float z = 0.0f;
    if (t < NP * CP) {
#pragma unroll
      for (int ks = 0; ks < kMaxKSL; ++ks)
        if (ks < KSL) z += part[ks * NP * CP + t];
      z += (sm + lay.bias[l])[tq];
    }
    return z;
  };

  // ================================ forward sweep ================================
  for (int n = 0; n < a.n_steps; ++n) {
    const float tn = a.t0 + (float)n * a.h;
    for (int s = 0; s < stages; ++s) {
      const int q = n * stages + s;
      const int64_t row0 = (int64_t)q * a.B + col0;
      const bool recorder = (rank == (q & (CS - 1)));   // the replicated records are written by one CTA, in rotation
      for (int l = 0; l < L - 1; ++l) {
        const int N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
        const float z = slice_forward(l, l == 0 ? xs : sm + lay.hfull[l - 1], l == 0 ? 0 : a.nd.sizes[l] * CP);
        OPZ_PH(0);
        const bool mine = t < 3 * UL * CP;
        const int net = mine ? tq / UL : 0, u = tq - net * UL;
        const bool live = mine && (u0 + u < N);
        const float hv = act_dyn(z, a.nd.act);
        if (l < L - 2) {
          if (live) {
            const uint32_t off = 4u * (uint32_t)(lay.hfull[l] + (net * N + u0 + u) * CP + tc);
            if constexpr (CS > 1) {
              // (one vector store of the CP columns per (item, peer), staged through shared memory, measured slower)
#pragma unroll
              for (int d = 0; d < CS; ++d) st_async(rbase[d] + off, hv, rbase[d] + xb_off + 8u * (uint32_t)(XB_GATHER + l));
            } else {
              sm[lay.hfull[l] + (net * N + u0 + u) * CP + tc] = hv;
            }
          }
        } else if (mine) {
          hs[t] = live ? hv : 0.0f;
        }
        // HBM records (consumed by the backward sweep and the weight-gradient GEMMs)
        if (live && col_ok) {
          const int64_t o = ((int64_t)net * a.rows_cap + row0 + tc) * N + u0 + u;
          a.H[l][o] = hv;
          a.G[l][o] = act_grad_dyn(z, a.nd.act);
        }
        if (l == 0 && recorder && t < S3 * CP && col_ok) a.X[(row0 + tc) * S3 + tq] = xs[t];
        if (l < L - 2) xb_wait_repost(XB_GATHER + l, xb_bytes(XB_GATHER + l));
        else __syncthreads();
        OPZ_PH(1);
      }
      // last layer: split-K over this CTA's slice of the last hidden layer, then all-reduce
      {
        const int UL = lay.UL[L - 2];
        const float* W = sm + lay.W[L - 1];
        float yp = 0.0f;
        if (t < NO3 * CP) {
          const int net = tq / nout, j = tq - net * nout;
          const float* Wp = W + (size_t)net * UL * nout + j;
          const float* hp = hs + net * UL * CP + tc;
          float yq = 0.0f;   // two independent chains: the loop is bound by the FMA / LDS latency, not by throughput
          int u = 0;
#pragma unroll 4
          for (; u + 1 < UL; u += 2) {
            yp = fmaf(hp[u * CP], Wp[u * nout], yp);
            yq = fmaf(hp[(u + 1) * CP], Wp[(u + 1) * nout], yq);
          }
          if (u < UL) yp = fmaf(hp[u * CP], Wp[u * nout], yp);
          yp += yq;
        }
        OPZ_PH(2);
        all_reduce(yp, NO3 * CP, yb, sm + lay.bias[L - 1], XB_AR1F, XB_AR2F, [] {});
        OPZ_PH(3);
      }
      // physics + Runge-Kutta update, replicated in every CTA (opz_device.cuh formulas, one thread per face / cell)
      const float ts = tn + a.rk_c[s] * a.h;
      if (t < (nz + 1) * CP) {
        const int f = tq, c = tc;
        const float* bc = a.bc + (col0 + min(c, ncol - 1)) * OPZ_BC_STRIDE;
        float Fu, Fv, FT;
        if (f == 0) { Fu = bc[0] - P.off[0]; Fv = bc[2] - P.off[1]; FT = bc[4] - P.off[2]; }
        else if (f == nz) { Fu = bc[1] - P.off[0]; Fv = bc[3] - P.off[1]; FT = wT_top_value(P, bc, ts); }
        else if (P.flags & (OPZ_FLAG_SMOOTH_NN | OPZ_FLAG_SMOOTH_RI)) {   // box-filtered NN outputs / Richardson numbers (opz_faces.cuh)
          float F3[3];
          faces::forward_filtered(P, xs, yb, CP, c, f, F3);
          Fu = F3[0]; Fv = F3[1]; FT = F3[2];
        } else {
          Fu = yb[(f - 1) * CP + c]; Fv = yb[(nout + f - 1) * CP + c]; FT = yb[(2 * nout + f - 1) * CP + c];
          if (P.flags & OPZ_FLAG_MPP) {
            const float gu = (xs[f * CP + c] - xs[(f - 1) * CP + c]) * P.nzf;
            const float gv = (xs[(nz + f) * CP + c] - xs[(nz + f - 1) * CP + c]) * P.nzf;
            const float gT = (xs[(2 * nz + f) * CP + c] - xs[(2 * nz + f - 1) * CP + c]) * P.nzf;
            const float Ri = richardson(P, gu, gv, gT);
            const float nu = nu_of_ri(P, Ri);
            const float nuT = nuT_of(P, nu, gu, gT);
            Fu -= P.c_u * nu * gu;
            Fv -= P.c_v * nu * gv;
            FT -= P.c_T * nuT * gT;
          }
        }
        Fs[f * CP + c] = Fu; Fs[((nz + 1) + f) * CP + c] = Fv; Fs[(2 * (nz + 1) + f) * CP + c] = FT;
      }
      __syncthreads();
      OPZ_PH(4);
      float xnext = 0.0f;
      const bool last = (s == stages - 1);
      if (t < S3 * CP) {
        const int i = tq, c = tc, v = i / nz, cz = i - v * nz;
        const float* F = Fs + v * (nz + 1) * CP + c;
        float k;
        if (v == 0) k = P.A_u * ((F[(cz + 1) * CP] - F[cz * CP]) * P.nzf) + P.cor_u * (P.s_v * xs[(nz + cz) * CP + c] + P.mu_v);
        else if (v == 1) k = P.A_v * ((F[(cz + 1) * CP] - F[cz * CP]) * P.nzf) - P.cor_v * (P.s_u * xs[cz * CP + c] + P.mu_u);
        else k = P.A_T * ((F[(cz + 1) * CP] - F[cz * CP]) * P.nzf);
        const float x0v = xn[t], hh = a.h;
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
      }
      __syncthreads();  // the Coriolis terms read OTHER cells of xs: nobody overwrites xs before everybody has read it
      if (t < S3 * CP) {
        xs[t] = xnext;
        if (last) {
          xn[t] = xnext;
          if (recorder && col_ok) a.XN[((int64_t)(n + 1) * a.B + col0 + tc) * S3 + tq] = xnext;
        }
      }
      __syncthreads();
      OPZ_PH(5);
    }
  }

  // ================================ backward sweep ================================
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
  // records of right-hand side q needed by its back-propagation: this CTA's act'(z_l) slices and the stage input
  float g_next[kMaxLayers - 1], x_next = 0.0f;
  auto prefetch = [&](int q) {
    const int64_t row0 = (int64_t)q * a.B + col0;
#pragma unroll
    for (int l = 0; l < kMaxLayers - 1; ++l) {
      g_next[l] = 0.0f;
      if (l < L - 1 && q >= 0) {
        const int N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
        if (t < 3 * UL * CP) {
          const int net = tq / UL, u = tq - net * UL;
          if (u0 + u < N && col_ok) g_next[l] = __ldcg(a.G[l] + ((int64_t)net * a.rows_cap + row0 + tc) * N + u0 + u);
        }
      }
    }
    x_next = (q >= 0 && t < S3 * CP && col_ok) ? __ldcg(a.X + (row0 + tc) * S3 + tq) : 0.0f;
  };
  cluster_sync_all<CS>();  // the recorders' checkpoint / record stores are visible to the whole cluster
  if (t < S3 * CP) {
    const float l0 = col_ok ? loss_cotangent(a.n_steps, tq, tc) : 0.0f;
    lam[t] = l0;
    kbar[t] = a.rk_b[stages - 1] * l0;
  }
  prefetch(a.n_steps * stages - 1);
  __syncthreads();

  float gp[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // this thread's share of d loss / d (nu0, nu_m, dRi, Ric, Pr) (rank 0 only: the physics is replicated)
  int rs_parity = 0;
  for (int n = a.n_steps - 1; n >= 0; --n) {
    for (int s = stages - 1; s >= 0; --s) {
      const int q = n * stages + s;
      const int64_t row0 = (int64_t)q * a.B + col0;
      const bool recorder = (rank == (q & (CS - 1)));
      // land the prefetched records in shared memory, then prefetch the next right-hand side's
#pragma unroll
      for (int l = 0; l < kMaxLayers - 1; ++l)
        if (l < L - 1 && t < 3 * lay.UL[l] * CP) (sm + lay.gs[l])[t] = g_next[l];
      if (t < S3 * CP) xs[t] = x_next;
      prefetch(q - 1);
      __syncthreads();
      OPZ_PH(8);
      // ---- physics VJP (replicated) ----
      float dy_u = 0.0f, dy_v = 0.0f, dy_T = 0.0f;
      if (t < (nz + 1) * CP) {
        const int f = tq, c = tc;
        float gbu = 0.0f, gbv = 0.0f, gbT = 0.0f;
        if (f > 0 && f < nz && (P.flags & (OPZ_FLAG_SMOOTH_NN | OPZ_FLAG_SMOOTH_RI))) {
          float dy3[3], gb3[3];
          faces::backward_filtered(P, xs, kbar, CP, c, f, dy3, gb3, (a.gphys && rank == 0 && col_ok) ? gp : nullptr);
          dy[(f - 1) * CP + c] = dy3[0]; dy[(nout + f - 1) * CP + c] = dy3[1]; dy[(2 * nout + f - 1) * CP + c] = dy3[2];
          dy_u = dy3[0]; dy_v = dy3[1]; dy_T = dy3[2];
          gbu = gb3[0]; gbv = gb3[1]; gbT = gb3[2];
        } else if (f > 0 && f < nz) {
          const float Fbu = P.A_u * P.nzf * (kbar[(f - 1) * CP + c] - kbar[f * CP + c]);
          const float Fbv = P.A_v * P.nzf * (kbar[(nz + f - 1) * CP + c] - kbar[(nz + f) * CP + c]);
          const float FbT = P.A_T * P.nzf * (kbar[(2 * nz + f - 1) * CP + c] - kbar[(2 * nz + f) * CP + c]);
          dy[(f - 1) * CP + c] = Fbu; dy[(nout + f - 1) * CP + c] = Fbv; dy[(2 * nout + f - 1) * CP + c] = FbT;
          dy_u = Fbu; dy_v = Fbv; dy_T = FbT;
          if (P.flags & OPZ_FLAG_MPP) {
            const float gu = (xs[f * CP + c] - xs[(f - 1) * CP + c]) * P.nzf;
            const float gv = (xs[(nz + f) * CP + c] - xs[(nz + f - 1) * CP + c]) * P.nzf;
            const float gT = (xs[(2 * nz + f) * CP + c] - xs[(2 * nz + f - 1) * CP + c]) * P.nzf;
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
            if (a.gphys && rank == 0 && col_ok) mpp_param_cotangents(P, nubar, FbT, th, Ri, nu, gT, keep, ok, gp);   // not the padding slots
          }
        }
        Fs[f * CP + c] = gbu; Fs[((nz + 1) + f) * CP + c] = gbv; Fs[(2 * (nz + 1) + f) * CP + c] = gbT;
      }
      __syncthreads();
      if (t < S3 * CP) {
        const int i = tq, c = tc, v = i / nz, cz = i - v * nz;
        const float* g = Fs + v * (nz + 1) * CP + c;
        float xb = P.nzf * (g[cz * CP] - g[(cz + 1) * CP]);
        if (v == 0) xb -= P.cor_v * P.s_u * kbar[(nz + cz) * CP + c];
        else if (v == 1) xb += P.cor_u * P.s_v * kbar[cz * CP + c];
        xphys[t] = xb;
      }
      OPZ_PH(9);
      // ---- dz of the last hidden layer: slice-local ----
      float pend = 0.0f;          // dz value whose HBM record store is deferred past the next barrier arrive
      int64_t pend_off = -1;
      int pend_layer = 0;
      {
        const int l = L - 2, N = a.nd.sizes[l + 1], UL = lay.UL[l], u0 = rank * UL;
        const float* W = sm + lay.W[L - 1];
        if (t < 3 * UL * CP) {
          const int net = tq / UL, u = tq - net * UL;
          const float* Wp = W + (size_t)(net * UL + u) * nout;
          const float* dp = dy + net * nout * CP + tc;
          float sacc = 0.0f, sacc2 = 0.0f;
          int j = 0;
#pragma unroll 4
          for (; j + 1 < nout; j += 2) {
            sacc = fmaf(Wp[j], dp[j * CP], sacc);
            sacc2 = fmaf(Wp[j + 1], dp[(j + 1) * CP], sacc2);
          }
          if (j < nout) sacc = fmaf(Wp[j], dp[j * CP], sacc);
          sacc += sacc2;
          const float dz = sacc * (sm + lay.gs[l])[t];
          dzs[t] = dz;
          if (u0 + u < N && col_ok) { pend = dz; pend_layer = l; pend_off = ((int64_t)net * a.rows_cap + row0 + tc) * N + u0 + u; }
        }
      }
      auto flush = [&]() {
        if (pend_off >= 0) { a.G[pend_layer][pend_off] = pend; pend_off = -1; }
      };
      __syncthreads();
      OPZ_PH(10);
      // ---- hidden layers: partial dh_{l-1} over all inputs of layer l -> reduce-scatter to the owners ----
      for (int l = L - 2; l >= 1; --l) {
        const int K = a.nd.sizes[l], UL = lay.UL[l];        // layer l: K inputs (= width of hidden l-1)
        const int ULp = lay.UL[l - 1], u0p = rank * ULp;     // slicing of hidden layer l-1
        const float* W = sm + lay.W[l];
        float* rb = sm + lay.recv_big + rs_parity * lay.recv_big_size;
        const int rs_ch = XB_RS + rs_parity;
        if constexpr (CS > 1) { if (t == 0) xb_expect(xb0 + 8u * rs_ch, rs_bytes(l)); }   // (every sender's bytes may already be here)
        auto send_dh = [&](int net, int k, const float (&v)[CP]) {
          const int owner = k / ULp, kl = k - owner * ULp;
          const int idx = ((rank * 3 + net) * ULp + kl) * CP;
          if constexpr (CS > 1) {
            st_async_cols<CP>(rbase[owner] + 4u * (uint32_t)((int)(rb - sm) + idx), v, rbase[owner] + xb_off + 8u * (uint32_t)rs_ch);
          } else {
#pragma unroll
            for (int c = 0; c < CP; ++c) rb[idx + c] = v[c];
          }
        };
        if (K % kRB == 0) {
          // one thread = kRB consecutive inputs k of one net x all CP columns
          const int gpn = K / kRB;
          for (int it = t; it < 3 * gpn; it += kThreads) {
            const int net = it / gpn, k0 = (it - net * gpn) * kRB;
            const float* Wp = W + ((size_t)net * K + k0) * UL;
            const float* dp = dzs + net * UL * CP;
            float acc[kRB][CP];
#pragma unroll
            for (int j = 0; j < kRB; ++j)
#pragma unroll
              for (int c = 0; c < CP; ++c) acc[j][c] = 0.0f;
#pragma unroll 5
            for (int u = 0; u < UL; ++u) {
              float dv[CP];
              load_cols<CP>(dv, dp + u * CP);
#pragma unroll
              for (int j = 0; j < kRB; ++j) {
                const float w = Wp[j * UL + u];
#pragma unroll
                for (int c = 0; c < CP; ++c) acc[j][c] = fmaf(w, dv[c], acc[j][c]);
              }
            }
#pragma unroll
            for (int j = 0; j < kRB; ++j) send_dh(net, k0 + j, acc[j]);
          }
        } else {
          for (int it = t; it < 3 * K; it += kThreads) {
            const int net = it / K, k = it - net * K;
            const float* Wp = W + ((size_t)net * K + k) * UL;
            const float* dp = dzs + net * UL * CP;
            float acc[CP];
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
#pragma unroll 5
            for (int u = 0; u < UL; ++u) fma_cols<CP>(acc, Wp[u], dp + u * CP);
            send_dh(net, k, acc);
          }
        }
        OPZ_PH(11);
        flush();
        if constexpr (CS > 1) {
          xb_wait(xb0 + 8u * rs_ch, (xb_par >> rs_ch) & 1u);
          xb_par ^= 1u << rs_ch;
        }
        __syncthreads();
        OPZ_PH(12);
        if (t < 3 * ULp * CP) {
          const int net = tq / ULp, u = tq - net * ULp;
          float sacc = 0.0f;
#pragma unroll
          for (int src = 0; src < CS; ++src) sacc += rb[(src * 3 * ULp) * CP + t];
          const float dz = sacc * (sm + lay.gs[l - 1])[t];
          dzs[t] = dz;
          const int N = a.nd.sizes[l];
          if (u0p + u < N && col_ok) { pend = dz; pend_layer = l - 1; pend_off = ((int64_t)net * a.rows_cap + row0 + tc) * N + u0p + u; }
        }
        if (L >= 4) rs_parity ^= 1;
        __syncthreads();
        OPZ_PH(13);
      }
      // ---- first layer: partial input cotangent over this CTA's units, all-reduce ----
      {
        const int UL = lay.UL[0];
        const float* W = sm + lay.W[0];
        float xp = 0.0f;
        constexpr int kIB = 4, kUS = 4;   // register-blocked form: kIB state entries x CP columns per thread, units split kUS ways
        const int pstride = S3 * CP + 4;   // partial-sum planes 16 B out of phase: the vector stores of a quarter warp do not collide
        if (CS > 1 && S3 % kIB == 0 && 3 * (S3 / kIB) * kUS <= kThreads && 3 * kUS * pstride <= lay.recv_big_size) {
          float* scratch = sm + lay.recv_big;   // this right-hand side's reduce-scatter data has been consumed
          const int ngi = S3 / kIB;
          if (t < 3 * ngi * kUS) {
            const int us = t % kUS, ig = (t / kUS) % ngi, net = t / (kUS * ngi);
            const int ulen = (UL + kUS - 1) / kUS, ua = us * ulen, ub = min(UL, ua + ulen);
            float acc[kIB][CP];
#pragma unroll
            for (int j = 0; j < kIB; ++j)
#pragma unroll
              for (int c = 0; c < CP; ++c) acc[j][c] = 0.0f;
            const float* Wp = W + ((size_t)net * S3 + ig * kIB) * UL;
            const float* dp = dzs + net * UL * CP;
            for (int u = ua; u < ub; ++u) {
              float dv[CP];
              load_cols<CP>(dv, dp + u * CP);
#pragma unroll
              for (int j = 0; j < kIB; ++j) {
                const float w = Wp[j * UL + u];
#pragma unroll
                for (int c = 0; c < CP; ++c) acc[j][c] = fmaf(w, dv[c], acc[j][c]);
              }
            }
            float* sp = scratch + (net * kUS + us) * pstride + ig * kIB * CP;
#pragma unroll
            for (int j = 0; j < kIB; ++j) store_cols<CP>(sp + j * CP, acc[j]);
          }
          __syncthreads();
          if (t < S3 * CP) {
#pragma unroll
            for (int q2 = 0; q2 < 3 * kUS; ++q2) xp += scratch[q2 * pstride + t];
          }
        } else if (t < S3 * CP) {
          float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;   // the three nets as three independent chains (latency-bound loop)
          const float* Wp = W + (size_t)tq * UL;
          const float* dp = dzs + tc;
          const int wn = S3 * UL, dn = UL * CP;
#pragma unroll 5
          for (int u = 0; u < UL; ++u) {
            x0 = fmaf(Wp[u], dp[u * CP], x0);
            x1 = fmaf(Wp[wn + u], dp[dn + u * CP], x1);
            x2 = fmaf(Wp[2 * wn + u], dp[2 * dn + u * CP], x2);
          }
          xp = (x0 + x1) + x2;
        }
        OPZ_PH(14);
        all_reduce(xp, S3 * CP, yb, nullptr, XB_AR1B, XB_AR2B, [&] {
          flush();
          if (recorder && t < (nz + 1) * CP && tq > 0 && tq < nz && col_ok) {
            a.DY[((int64_t)0 * a.rows_cap + row0 + tc) * nout + tq - 1] = dy_u;
            a.DY[((int64_t)1 * a.rows_cap + row0 + tc) * nout + tq - 1] = dy_v;
            a.DY[((int64_t)2 * a.rows_cap + row0 + tc) * nout + tq - 1] = dy_T;
          }
        });
        OPZ_PH(15);
      }
      // ---- Runge-Kutta adjoint bookkeeping (replicated) ----
      if (t < S3 * CP) {
        const float xbar = yb[t] + xphys[t];
        const float lv = lam[t];
        const float xsv = (s == stages - 1 ? 0.0f : xsum[t]) + xbar;
        if (s > 0) {
          kbar[t] = a.rk_b[s - 1] * lv + a.rk_a[s - 1] * xbar;
          xsum[t] = xsv;
        } else {
          float ln = lv + xsv;
          if (n % a.save_every == 0 && col_ok) ln += loss_cotangent(n, tq, tc);
          lam[t] = ln;
          kbar[t] = a.rk_b[stages - 1] * ln;
        }
      }
      __syncthreads();
      OPZ_PH(16);
    }
  }
  if (a.gphys && rank == 0) mpp_param_flush(gp, a.gphys);   // every thread of the CTA takes part (zeros outside the face threads)
  if (timing)
    for (int i = 0; i < 24; ++i) g_phase_cycles[i] = (unsigned long long)ph[i];
  cluster_sync_all<CS>();  // no CTA exits while a peer may still address its shared memory
}

}  // namespace cl
