// opz_tc_probe.cu — one-tile tcgen05 GEMM used by the tests to pin the descriptor / TMEM layout
// conventions the fused rollout kernel relies on:  D[128 x N] = A[128 x K] * B[N x K]^T with BF16
// operands, FP32 accumulation in tensor memory.
//   mode 0: A and B from shared memory (K-major, no swizzle, bulk-async-copied images)
//   mode 1: A from tensor memory (written with tcgen05.st as packed BF16 pairs), B from shared memory
// Exported as opz_debug_umma (a debug entry point, not part of include/opz.h).
#include <cuda_bf16.h>
#include <vector>
#include "opz_internal.h"
#include "opz_ptx.cuh"

namespace opz {

struct ProbeArgs {
  const uint8_t* a_img;   // [K/8][128][8] bf16
  const uint8_t* b_img;   // [K/8][N][8] bf16
  const uint16_t* a_rows; // [128][K] bf16 row-major (mode 1)
  float* d;               // [128][N]
  int N, K, mode;
};

__global__ void __launch_bounds__(192, 1) umma_probe_kernel(const __grid_constant__ ProbeArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t bytesA = 128u * a.K * 2u, bytesB = (uint32_t)a.N * a.K * 2u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + bytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + bytesA + bytesB);
  uint64_t* bar_full = bars + 0;
  uint64_t* bar_done = bars + 1;
  uint64_t* bar_aready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(bar_full, 1);
      ptx::mbar_init(bar_done, 1);
      ptx::mbar_init(bar_aready, 128);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_a = tmem + 256;  // columns [256, 256 + K/2) hold A in mode 1

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(bar_full, bytesA + bytesB);
      ptx::bulk_g2s(sA, a.a_img, bytesA, bar_full);
      ptx::bulk_g2s(sB, a.b_img, bytesB, bar_full);
    }
  } else if (warp == 1) {
    ptx::mbar_wait(bar_full, 0);
    if (a.mode == 1) ptx::mbar_wait(bar_aready, 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16(128, a.N);
      const uint32_t lboA = 128 * 16, lboB = (uint32_t)a.N * 16;
      for (int ks = 0; ks < a.K / 16; ++ks) {
        const uint64_t bd = ptx::smem_desc(ptx::smem_u32(sB) + ks * 2 * lboB, lboB, 128);
        if (a.mode == 0) {
          const uint64_t ad = ptx::smem_desc(ptx::smem_u32(sA) + ks * 2 * lboA, lboA, 128);
          ptx::mma_ss(tmem, ad, bd, idesc, ks > 0);
        } else {
          ptx::mma_ts(tmem, tmem_a + ks * 8, bd, idesc, ks > 0);
        }
      }
      ptx::tc_commit(bar_done);
    }
    __syncwarp();
  } else {
    const int q = warp % 4;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    if (a.mode == 1) {
      for (int c = 0; c < a.K / 2; c += 8) {
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t lo = a.a_rows[(size_t)row * a.K + 2 * (c + j)];
          const uint32_t hi = a.a_rows[(size_t)row * a.K + 2 * (c + j) + 1];
          r[j] = lo | (hi << 16);
        }
        ptx::tmem_st8(tmem_a + lane_addr + c, r);
      }
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_aready);
    }
    ptx::mbar_wait(bar_done, 0);
    ptx::tc_fence_after();
    for (int c = 0; c < a.N; c += 16) {
      uint32_t r[16];
      ptx::tmem_ld16(tmem + lane_addr + c, r);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) a.d[(size_t)row * a.N + c + j] = __uint_as_float(r[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem);
}

}  // namespace opz

using namespace opz;

// A: [128][K] bf16 bits row-major, B: [N][K] bf16 bits, D: [128][N] float.  HOST pointers.
extern "C" int opz_debug_umma(int N, int K, int mode, const uint16_t* A, const uint16_t* B, float* D) {
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 512) { set_error("probe: bad N/K"); return OPZ_EINVAL; }
  std::vector<uint16_t> ai((size_t)128 * K), bi((size_t)N * K);
  for (int k = 0; k < K; ++k)
    for (int r = 0; r < 128; ++r) ai[((size_t)(k / 8) * 128 + r) * 8 + (k % 8)] = A[(size_t)r * K + k];
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) bi[((size_t)(k / 8) * N + n) * 8 + (k % 8)] = B[(size_t)n * K + k];
  uint8_t *da, *db;
  uint16_t* dar;
  float* dd;
  OPZ_CUDA(cudaMalloc(&da, ai.size() * 2));
  OPZ_CUDA(cudaMalloc(&db, bi.size() * 2));
  OPZ_CUDA(cudaMalloc(&dar, ai.size() * 2));
  OPZ_CUDA(cudaMalloc(&dd, sizeof(float) * 128 * N));
  OPZ_CUDA(cudaMemcpy(da, ai.data(), ai.size() * 2, cudaMemcpyHostToDevice));
  OPZ_CUDA(cudaMemcpy(db, bi.data(), bi.size() * 2, cudaMemcpyHostToDevice));
  OPZ_CUDA(cudaMemcpy(dar, A, ai.size() * 2, cudaMemcpyHostToDevice));
  OPZ_CUDA(cudaMemset(dd, 0, sizeof(float) * 128 * N));
  ProbeArgs pa{da, db, dar, dd, N, K, mode};
  const size_t smem = (size_t)128 * K * 2 + (size_t)N * K * 2 + 64;
  OPZ_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 192, smem>>>(pa);
  OPZ_CUDA(cudaGetLastError());
  OPZ_CUDA(cudaDeviceSynchronize());
  OPZ_CUDA(cudaMemcpy(D, dd, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dar); cudaFree(dd);
  return OPZ_OK;
}

// ---- issue-rate probe: how many cycles does one tcgen05.mma (M=128, N, K=16, kind::f16) cost when a single
// thread issues `per_commit` of them back to back and then waits for the commit, `reps` times?
namespace opz {
__global__ void __launch_bounds__(192, 1) umma_rate_kernel(int N, int per_commit, int reps, int mode, int same_d, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    if (threadIdx.x == 0) { ptx::mbar_init(bars, 1); ptx::mbar_init(bars + 1, 1); ptx::fence_barrier_init(); tmem_slot[1] = 0u; }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = ptx::idesc_f16kind(128, N, true);
    const uint32_t sb = ptx::smem_u32(smem), sa = ptx::smem_u32(smem + 48 * 1024);
    uint32_t parity = 0;
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      long long t_issue = 0;
      for (int r = 0; r < reps; ++r) {
        const long long ti = clock64();
        for (int i = 0; i < per_commit; ++i) {
          const uint64_t bd = ptx::smem_desc(sb + (uint32_t)(i % 8) * 2u * (uint32_t)N * 16u, (uint32_t)N * 16u, 128);
          const uint32_t d = tmem + (same_d >= 20 ? (uint32_t)((i & 1) * 128) : ((same_d & 1) || same_d >= 10 ? 0u : (uint32_t)((i & 1) * N)));  // >= 20: streaming, alternating accumulators
          if ((mode & 1) == 0) ptx::mma_ss(d, ptx::smem_desc(sa + (uint32_t)(i % 8) * 4096u, 2048, 128), bd, idesc, i > 0);
          else ptx::mma_ts(d, tmem + 256 + (uint32_t)(i % 16) * 8u, bd, idesc, i > 0);
        }
        if (same_d >= 10) {  // streaming variant: (same_d % 10) commits per batch, no completion wait inside
          for (int cc = 0; cc < same_d % 10; ++cc) ptx::tc_commit(bars + 1);
          t_issue += clock64() - ti;
          continue;
        }
        ptx::tc_commit(bars);
        t_issue += clock64() - ti;
        ptx::mbar_wait(bars, parity);
        parity ^= 1;
      }
      if (same_d >= 10) { ptx::tc_commit(bars); ptx::mbar_wait(bars, 0); }
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
      out[gridDim.x + blockIdx.x] = t_issue;
      *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1) = 1u;  // stop the load warps
    }
    __syncwarp();
  } else if (warp >= 2 && (mode & 6)) {
    // background TMEM traffic like the rollout kernel's epilogue: mode bit 1 = tcgen05.ld of 64 columns,
    // bit 2 = tcgen05.st of 32 columns, on columns the MMAs do not touch
    const uint32_t lane_addr = (uint32_t)((warp % 4) * 32) << 16;
    volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(tmem_slot + 1);
    uint32_t sink = 0;
    while (*stop == 0u) {
      if (mode & 2) {
        uint32_t r[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(tmem + lane_addr + 384 + g * 16, r[g]);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) sink ^= r[g][3];
      }
      if (mode & 4) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = sink + j;
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_st8(tmem + lane_addr + 448 + g * 8, pk);
        ptx::tmem_wait_st();
      }
    }
    if (sink == 0x12345u) out[0] = 0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}
}  // namespace opz

extern "C" int opz_debug_umma_rate(int N, int per_commit, int reps, int mode, int same_d, int grid, double* cycles_per_mma) {
  long long* d;
  OPZ_CUDA(cudaMalloc(&d, sizeof(long long) * 2 * grid));
  const size_t smem = 96 * 1024 + 64;
  OPZ_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_kernel<<<grid, 192, smem>>>(N, per_commit, reps, mode, same_d, d);
  OPZ_CUDA(cudaGetLastError());
  OPZ_CUDA(cudaDeviceSynchronize());
  std::vector<long long> h(2 * grid);
  OPZ_CUDA(cudaMemcpy(h.data(), d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
  cudaFree(d);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  cycles_per_mma[0] = (double)mx / ((double)per_commit * reps);
  cycles_per_mma[1] = (double)h[grid] / ((double)per_commit * reps);  // issue loop only (CTA 0)
  return OPZ_OK;
}

// ---- static issue probe: 25 tcgen05.mma (M=128, TS) per batch with immediate operand offsets, as the rollout kernel
// issues them; reports cycles per MMA for the whole run (issue + drain) and for the issue sequence alone.
namespace opz {
__device__ __forceinline__ void probe_mma_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
template <int N>
__global__ void __launch_bounds__(64, 1) umma_static_kernel(int reps, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    if (threadIdx.x == 0) { ptx::mbar_init(bars, 1); ptx::fence_barrier_init(); }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = ptx::idesc_f16kind(128, N, true);
    const uint64_t bd0 = ptx::smem_desc(ptx::smem_u32(smem), (uint32_t)N * 16u, 128);
    const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
    if (ptx::elect_one()) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 25; ++k) probe_mma_ts_lohi(tmem, tmem + 256u + (uint32_t)k * 8u, lo + (uint32_t)(k * N), hi, idesc, k > 0);
      }
      const long long t1 = clock64();
      ptx::tc_commit(bars);
      ptx::mbar_wait(bars, 0);
      const long long t2 = clock64();
      out[0] = t2 - t0;
      out[1] = t1 - t0;
    }
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}
}  // namespace opz

extern "C" int opz_debug_umma_static(int N, int reps, double* cycles_per_mma) {
  long long* d;
  OPZ_CUDA(cudaMalloc(&d, sizeof(long long) * 2));
  const size_t smem = 96 * 1024 + 64;
  void (*k)(int, long long*) = N == 8 ? opz::umma_static_kernel<8> : N == 16 ? opz::umma_static_kernel<16> : N == 32 ? opz::umma_static_kernel<32>
                               : N == 64 ? opz::umma_static_kernel<64> : N == 128 ? opz::umma_static_kernel<128> : opz::umma_static_kernel<256>;
  OPZ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<1, 64, smem>>>(reps, d);
  OPZ_CUDA(cudaGetLastError());
  OPZ_CUDA(cudaDeviceSynchronize());
  long long h[2];
  OPZ_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  cycles_per_mma[0] = (double)h[0] / (25.0 * reps);
  cycles_per_mma[1] = (double)h[1] / (25.0 * reps);
  return OPZ_OK;
}

// ---- tensor-memory read throughput: W warps each read 64 columns x 32 lanes (tcgen05.ld 32x32b.x16 x 4 + wait) `iters` times;
// with_mma = 1 keeps one thread issuing N = 64 TS MMAs into other columns meanwhile.
namespace opz {
__global__ void __launch_bounds__(576, 1) tmem_read_kernel(int n_warps, int iters, int with_mma, int x32, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    if (threadIdx.x == 0) { ptx::mbar_init(bars, 1); ptx::fence_barrier_init(); tmem_slot[1] = 0u; }
    __syncwarp();
    ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(tmem_slot + 1);
    if (with_mma) {
      const uint32_t idesc = ptx::idesc_f16kind(128, 64, true);
      const uint64_t bd0 = ptx::smem_desc(ptx::smem_u32(smem), 64u * 16u, 128);
      const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
      if (ptx::elect_one()) {
        long long n = 0;
        const long long t0 = clock64();
        while (*stop < (uint32_t)n_warps) {
#pragma unroll
          for (int k = 0; k < 25; ++k) probe_mma_ts_lohi(tmem, tmem + 256u + (uint32_t)k * 8u, lo + (uint32_t)(k * 64), hi, idesc, k > 0);
          n += 25;
        }
        const long long t1 = clock64();
        ptx::tc_commit(bars);
        ptx::mbar_wait(bars, 0);
        out[2] = t1 - t0;
        out[3] = n;
      }
      __syncwarp();
    }
  } else if (warp >= 2 && warp < 2 + n_warps) {
    const uint32_t lane_addr = (uint32_t)((warp % 4) * 32) << 16;
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (x32) {
        uint32_t r[2][32];
        ptx::tmem_ld32(tmem + lane_addr + 128, r[0]);
        ptx::tmem_ld32(tmem + lane_addr + 160, r[1]);
        ptx::tmem_wait_ld();
        sink ^= r[0][5] ^ r[1][7];
      } else {
        uint32_t r[4][16];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(tmem + lane_addr + 128 + g * 16, r[g]);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) sink ^= r[g][3];
      }
    }
    const long long t1 = clock64();
    if (sink == 0x12345u) out[7] = 0;
    if (threadIdx.x % 32 == 0) {
      if (warp == 2) { out[0] = t1 - t0; out[1] = iters; }
      atomicAdd(reinterpret_cast<unsigned int*>(tmem_slot + 1), 1u);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}
}  // namespace opz

extern "C" int opz_debug_tmem_read(int n_warps, int iters, int with_mma, int x32, double* res) {
  long long* d;
  OPZ_CUDA(cudaMalloc(&d, sizeof(long long) * 8));
  OPZ_CUDA(cudaMemset(d, 0, sizeof(long long) * 8));
  const size_t smem = 96 * 1024 + 64;
  OPZ_CUDA(cudaFuncSetAttribute(opz::tmem_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  opz::tmem_read_kernel<<<1, 576, smem>>>(n_warps, iters, with_mma, x32, d);
  OPZ_CUDA(cudaGetLastError());
  OPZ_CUDA(cudaDeviceSynchronize());
  long long h[8];
  OPZ_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  res[0] = (double)h[0] / (double)h[1];                 // cycles per 64-column read of one warp
  res[1] = h[3] ? (double)h[2] / (double)h[3] : 0.0;    // cycles per MMA meanwhile
  return OPZ_OK;
}

// ---- the same for a CTA pair: cta_group::2, M = 256, each CTA holds N / 2 rows of B ----
namespace opz {
template <int N>
__global__ void __launch_bounds__(64, 1) umma_static2_kernel(int reps, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x / 32;
  const uint32_t rank = ptx::cluster_ctarank();
  if (warp == 0) {
    if (threadIdx.x == 0) { ptx::mbar_init(bars, 1); ptx::fence_barrier_init(); }
    __syncwarp();
    ptx::tmem_alloc2<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = ptx::idesc_f16kind(256, N, true);
    const uint64_t bd0 = ptx::smem_desc(ptx::smem_u32(smem), (uint32_t)(N / 2) * 16u, 128);
    const uint32_t lo = (uint32_t)bd0, hi = (uint32_t)(bd0 >> 32);
    if (ptx::elect_one()) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 25; ++k) ptx::mma_ts2_lohi(tmem, tmem + 256u + (uint32_t)k * 8u, lo + (uint32_t)(k * N), hi, idesc, k > 0);
      }
      const long long t1 = clock64();
      ptx::tc_commit2(bars);
      ptx::mbar_wait(bars, 0);
      const long long t2 = clock64();
      out[0] = t2 - t0;
      out[1] = t1 - t0;
    }
    __syncwarp();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 0) ptx::tmem_dealloc2<512>(tmem);
}
}  // namespace opz

extern "C" int opz_debug_umma_static2(int N, int reps, double* cycles_per_mma) {
  long long* d;
  OPZ_CUDA(cudaMalloc(&d, sizeof(long long) * 2));
  const size_t smem = 96 * 1024 + 64;
  void (*k)(int, long long*) = N == 16 ? opz::umma_static2_kernel<16> : N == 32 ? opz::umma_static2_kernel<32>
                               : N == 64 ? opz::umma_static2_kernel<64> : N == 128 ? opz::umma_static2_kernel<128> : opz::umma_static2_kernel<256>;
  OPZ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OPZ_CUDA(cudaLaunchKernelEx(&cfg, k, reps, d));
  OPZ_CUDA(cudaGetLastError());
  OPZ_CUDA(cudaDeviceSynchronize());
  long long h[2];
  OPZ_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  cycles_per_mma[0] = (double)h[0] / (25.0 * reps);
  cycles_per_mma[1] = (double)h[1] / (25.0 * reps);
  return OPZ_OK;
}
