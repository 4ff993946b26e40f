// opz_tc2.cu — the fused persistent tcgen05 / TMEM rollout kernel: CTA pairs, chunk-sized weight stages, a host-built schedule
// table, eight epilogue warps and level-parallel physics.
//
// Shape of the kernel (what the first-generation kernel of round 1 taught; that kernel is gone):
//   * a single MMA-issuing warp spending ~570 cycles of bookkeeping per 6-MMA weight stage was the limiter.  Here a weight stage
//     is a WHOLE accumulator chunk (up to 25 k-steps + the 4 output-layer k-steps of the fused chunk two stages back),
//     43 stages per right-hand side, each issued from one unrolled, immediate-offset sequence by one of two issuing warps;
//   * that needs 27.6 KB per stage, which only fits because the kernel always runs as a CTA PAIR
//     (tcgen05 cta_group::2, M = 256): each CTA stages HALF of every weight block and the hardware shares the
//     halves, so L2 -> SM weight traffic per SM halves too.  The ring is a byte FIFO (host-computed offsets and
//     "stages that may stay in flight" per stage) rather than fixed slots;
//   * the schedule of a right-hand side is the same every time: build_plan() lays it out once as a table of ChunkRec records
//     (kernel parameters), and issuers and epilogue walk that table — everything that does not depend on a barrier is computed
//     before the waits, so a hand-over costs the barrier, a fence and the tcgen05 instructions themselves;
//   * 8 epilogue warps (two per TMEM lane quadrant, 32 accumulator columns each) instead of 4 halve the drain
//     latency per chunk, and the physics runs LEVEL-PARALLEL: two threads per ocean column, 16 cells each, with
//     the NN-independent part (gradients, Ri, diffusivities, Coriolis) evaluated BEFORE the last output-layer
//     MMAs have landed;
//   * hidden-layer biases ride in the weight tape as one extra k-step against a constant-one operand: the drain has no bias
//     loads or adds.
//
// The physics evaluates the same FP32 expressions per face as opz_device.cuh::column_rhs<FAST>.
//
// Numerics of the 16-bit operands (measured: tests/test_gpu_long.py, scripts/precision_study.py).  Over the
// 23 040 steps of the 8-day rollout the error of a plainly rounded FP16 path is NOT the per-step rounding noise but the part of
// the rounding that is CONSTANT in time: the weights (rounded once: a slightly different model, 1.6e-2 after 8 days) and the
// slowly varying state image.  Both are therefore dithered in time, at no cost in tensor work:
//   * the weight tape exists in kTapes = 16 versions; version t holds, per weight, the FP16 neighbour below or above it such that
//     the MEAN over the 16 versions reproduces the FP32 weight to 4 more bits (ordered dither, thresholds in bit-reversed order,
//     phase rotated per weight); right-hand side (step n, stage s) streams version (n + s) mod 16, so every version visits every
//     Runge-Kutta stage slot equally often;
//   * the 16-bit image of the stage state is rounded with a Weyl-sequence threshold added to the magnitude bits before truncation
//     (a different threshold every right-hand side): its time average carries the FP32 value.
// Hidden activations are rounded to nearest as before (their rounding is not persistent).
//
// Replaces the batched form of NDE! / predict_flux! (wind_mixing/src/training_postprocessing.jl:105-153)
// inside a fixed-step solve (solve_NDE_mutating, :55-159).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "opz_internal.h"
#include "opz_ptx.cuh"

namespace opz {
namespace tc2 {

constexpr int kNz = 32;
constexpr int kS = 3 * kNz;          // 96 state entries per column
constexpr int kM = 128;              // columns per tile (one tile per CTA, two per pair)
constexpr int kCW = 64;              // accumulator chunk width (TMEM columns)
constexpr int kMaxHidden = 400;      // HF region: 200 TMEM columns
constexpr int kNB = 8;               // stage barrier pairs (stages in flight <= kNB - 1)
constexpr int kMaxStages = 72;       // stages per right-hand side
constexpr int kSplitK = 12;          // a 64-row x 25-k-step chunk is staged as two weight stages: k-steps [0, 12) and [12, 25) + bias + W_out
constexpr int kRingBytes = 84480;    // 3 x (32 rows x (400 + 8) k + 16 rows x 64 k) x 2 B
constexpr int kMaxKs = 28;           // 7 HF chunks x 4 k-steps
constexpr int kTapes = 16;           // temporally dithered versions of the weight tape (power of two)

constexpr uint32_t kColD0 = 0, kColY = 128, kColX = 224, kColH2 = 272, kColHF = 304, kColOne = 504;
// kColOne: one k-step (8 columns) holding [1, 0, ..., 0] per lane: the A operand of the bias MMAs

enum : int {
  B_FULL = 0,                   // [kNB] stage bytes landed               (producer expect_tx + peer forward)
  B_EMPTY = B_FULL + kNB,       // [kNB] stage consumed                   (tcgen05.commit, both CTAs)
  B_DFULL = B_EMPTY + kNB,      // [2] accumulator chunk complete         (tcgen05.commit, both CTAs)
  B_DEMPTY = B_DFULL + 2,       // [2] accumulator chunk drained          (epilogue warps of both CTAs)
  B_H1 = B_DEMPTY + 2,          // [7] chunk c of HF written
  B_H2FULL = B_H1 + 7,          // (unused: "H2c written" is signalled by the chunk's D-empty arrival)
  B_H2EMPTY = B_H2FULL + 1,     // H2c consumed by the output-layer MMAs  (tcgen05.commit)
  B_YFULL = B_H2EMPTY + 1,      // Y complete                             (tcgen05.commit)
  B_XFULL = B_YFULL + 1,        // X written, Y consumed
  B_COUNT = B_XFULL + 1
};
static_assert(B_COUNT <= 32, "phase bits are tracked in one 32-bit mask");

constexpr int kMaxChunks = 48;       // accumulator chunks per right-hand side

// One accumulator chunk of the per-right-hand-side schedule.  The schedule is the same for every right-hand side, so the host
// lays it out once (build_plan) and both the MMA issuers and the epilogue warps walk this table instead of re-deriving the
// nested layer / chunk loops: what is left on their per-chunk critical path is the barrier wait, the fence and the
// tcgen05 instructions themselves (the wait profile had ~380 cycles of address arithmetic between an issuer's wake-up and its first MMA).
enum : uint8_t { CH_SPLIT = 1, CH_FUSED = 2, CH_WAIT_H1 = 4 };
struct ChunkRec {
  uint32_t offA, offB;     // ring byte offsets of the chunk's weight stage (and of its second stage when CH_SPLIT)
  uint32_t out_off;        // ring byte offset of the W_out block riding in this chunk's (last) stage, when out_ks > 0
  uint16_t a_col;          // tensor-memory column of the A operand (X, or the net's HF image)
  uint16_t dst_col;        // tensor-memory column the drained chunk goes to (its place in HF, or H2c)
  uint8_t rows;            // accumulator columns of the chunk = N of its MMAs (multiple of 16, <= kCW)
  uint8_t kin;             // k-steps of the layer input
  uint8_t stage;           // index of the chunk's first weight stage within the right-hand side
  uint8_t flags;           // CH_*
  uint8_t h1_bar;          // H1 barrier the drain of a non-fused chunk signals
  uint8_t h1_base;         // first H1 barrier of the layer's input (CH_WAIT_H1)
  uint8_t h1_seq;          // completions of those barriers earlier in the same right-hand side
  uint8_t out_ks;          // output-layer k-steps of the fused chunk two stages back, issued at the head of this stage (0: none)
  uint8_t out_net;         // ... and the net (Y accumulator) they belong to
  uint8_t pad[3];
};
static_assert(sizeof(ChunkRec) == 28, "ChunkRec layout");

struct Plan {                     // host-side, cached on the model
  int L = 0;
  int hpad[2] = {0, 0};
  int act = 0;
  bool fp16 = false;
  bool interleave = false;
  int n_stages = 0;
  int n_chunks = 0;
  ChunkRec ch[kMaxChunks];
  uint32_t tail_off = 0;          // ring offset of the tail stage (the W_out blocks of the last two chunks)
  uint8_t tail_ks[2] = {0, 0}, tail_net[2] = {0, 0};   // [0]: chunk n - 2 (ks 0: not a fused one), [1]: chunk n - 1
  uint32_t st_off[kMaxStages];    // ring offset of the stage (bytes)
  uint32_t st_bytes[kMaxStages];  // bytes PER CTA
  uint32_t st_src[kMaxStages];    // tape offset of the stage (leader half first, then the peer half)
  uint8_t st_keep[kMaxStages];    // earlier stages that may still be in flight when this one is filled
  size_t tape_bytes = 0;          // one version
  size_t tape_stride = 0;         // distance between versions (bytes)
  int n_tapes = 1;
  uint8_t* tape_dev = nullptr;
  float b3[96];                   // output-layer biases [net][32]
};

struct Args {
  DevParams P;
  const uint8_t* tape;
  const float* x0;
  const float* bc;
  float* out;
  int64_t B;
  int scheme;
  float dt, t0;
  int n_steps, save_every, n_save;
  int L, n_stages;
  int n_mma_warps;   // 1 or 2 issuing warps in the leader CTA
  int interleave;    // 1: layer-major schedule (all three nets' first layers, then their second layers), see build_plan
  int nc1;           // accumulator chunks of the first hidden layer
  int n_chunks;      // accumulator chunks per right-hand side
  int h1_mult;       // completions of an H1 barrier per right-hand side (3 net-major, 1 layer-major)
  uint32_t tail_off;
  uint8_t tail_ks[2], tail_net[2];
  ChunkRec ch[kMaxChunks];
  int tape_mask;     // versions of the tape - 1 (0: a single, plainly rounded tape)
  uint32_t tape_stride;
  int n0;            // global index of the first time step of this call (t0 / dt): the dither phases continue across calls
  int dither_x;      // 1: Weyl-dithered state image
  int hpad[2];
  float b3[96];   // output-layer biases [net][32] (the hidden-layer biases ride in the weight tape)
  float pk[18];   // products of the physics constants folded on the host (see PK_* below): one FFMA per use in the kernel
  uint32_t st_off[kMaxStages];
  uint32_t st_bytes[kMaxStages];
  uint32_t st_src[kMaxStages];
  uint8_t st_keep[kMaxStages];
};

struct SmemLayout {
  static constexpr size_t off_xn = 0;
  static constexpr size_t off_acc = off_xn + (size_t)kS * kM * 4;
  static constexpr size_t off_xs = off_acc + (size_t)kS * kM * 4;
  static constexpr size_t off_ring = (off_xs + (size_t)kS * kM * 4 + 127) / 128 * 128;
  static constexpr size_t off_bars = off_ring + (size_t)kRingBytes;
  static constexpr size_t off_tmem = off_bars + 8 * B_COUNT;
  static constexpr size_t bytes = off_tmem + 16;
};

// s_xn is stored with a rotation so that both the per-column accesses (threads = consecutive columns) and the
// transposing snapshot copies (threads = consecutive state entries) are bank-conflict free without padding
__device__ __forceinline__ int xn_idx(int i, int m) { return i * kM + ((m + i) & (kM - 1)); }

__device__ __forceinline__ void bar_wait(uint64_t* bars, int id, uint32_t& mask) {
  ptx::mbar_wait(bars + id, (mask >> id) & 1u);
  mask ^= 1u << id;
}

// wait-time profile (OPZ_TC_PROFILE=1, relu kernels only): [0..7] MMA warp, [8..15] first epilogue warp of CTA 0
__device__ unsigned long long g_tc2_wait[16];
enum { W_X = 0, W_DEMPTY = 1, W_H1 = 2, W_FULL = 3, W_H2FULL = 4, W_TOTAL = 5, W_DECODE = 6, W_ISSUE = 7,
       E_DFULL = 8, E_H2EMPTY = 9, E_YFULL = 10, E_PHYS_A = 11, E_TOTAL = 12, E_PHYS_B = 13, E_LD = 14, E_ST = 15 };
template <bool PROF>
__device__ __forceinline__ void bar_wait_t(uint64_t* bars, int id, uint32_t& mask, long long* acc, int slot) {
  if constexpr (PROF) {
    const long long t0 = clock64();
    bar_wait(bars, id, mask);
    acc[slot] += clock64() - t0;
  } else {
    bar_wait(bars, id, mask);
  }
}

// wait for completion number `phase` (0-based) of a barrier; callers never ask for anything older than the last completion
template <bool PROF>
__device__ __forceinline__ void wait_phase(uint64_t* bars, int id, uint32_t phase, long long* acc, int slot) {
  if constexpr (PROF) {
    const long long t0 = clock64();
    ptx::mbar_wait(bars + id, phase & 1u);
    acc[slot] += clock64() - t0;
  } else {
    ptx::mbar_wait(bars + id, phase & 1u);
  }
}

// Args::pk: the physics of this kernel works on raw cell differences d = x[i+1] - x[i]; the 1/Delta = Nz factor and the
// feature scalings are folded into these constants (the FP32 parity path keeps the oracle's operation order)
enum { PK_KBZ = 0, PK_EBZ, PK_KSA, PK_ESA, PK_KSB, PK_ESB, PK_C2, PK_R2, PK_CUN, PK_CVN, PK_CTN, PK_AU, PK_AV, PK_AT,
       PK_CCU, PK_DCU, PK_CCV, PK_DCV };

template <bool FP16, int ACT>
__device__ __forceinline__ uint32_t pack_act(float a, float b) {
  if constexpr (ACT == OPZ_ACT_RELU) return FP16 ? ptx::pack_f16x2_relu(a, b) : ptx::pack_bf16x2_relu(a, b);
  a = act_fast<ACT>(a);
  b = act_fast<ACT>(b);
  return FP16 ? ptx::pack_f16x2(a, b) : ptx::pack_bf16x2(a, b);
}

// k-steps [K0, K1) of a ROWS-row weight block against the A operand: one straight run of tcgen05.mma whose
// descriptor / address offsets are immediates (issued by the elected lane); the first k-step of a chunk overwrites D
template <int ROWS, int K0, int K1>
__device__ __forceinline__ void issue_range(uint32_t d, uint32_t a0, uint32_t lo, uint32_t hi, uint32_t idesc) {
#pragma unroll
  for (int k = K0; k < K1; ++k) ptx::mma_ts2_lohi(d, a0 + (uint32_t)k * 8u, lo + (uint32_t)(k * ROWS), hi, idesc, k > 0);
}
// HF chunks [3g', ...) of a 25-k-step layer in three groups: k-steps [0,12), [12,20), [20,25)
template <int ROWS>
__device__ __forceinline__ void issue_hf_group(int grp, uint32_t d, uint32_t a0, uint32_t lo, uint32_t hi, uint32_t idesc) {
  if (grp == 0) issue_range<ROWS, 0, 12>(d, a0, lo, hi, idesc);
  else if (grp == 1) issue_range<ROWS, 12, 20>(d, a0, lo, hi, idesc);
  else issue_range<ROWS, 20, 25>(d, a0, lo, hi, idesc);
}

// E epilogue warps per CTA: E / 4 threads share one ocean column (TMEM lane); each owns LV = 128 / E cells.
// SMOOTH compiles in the 3-point box filters on the NN outputs and on the Richardson numbers (flags OPZ_FLAG_SMOOTH_NN /
// OPZ_FLAG_SMOOTH_RI; filtering_operators.jl:1-15 at NDE_training.jl:98-102,121-123): a separate instantiation, so that the
// unfiltered kernel keeps its registers and code size.
template <bool FP16, int ACT, int E, bool PROF, bool SMOOTH = false>
__global__ void __launch_bounds__((3 + E) * 32, 1) rollout_tc2_kernel(const __grid_constant__ Args a) {
  constexpr int kThreads = (3 + E) * 32;
  constexpr int kMma2Warp = 2 + E;     // second MMA-issuing warp (the epilogue warps keep indices 2 .. 1 + E: warp % 4 = TMEM quadrant)
  constexpr int kEpiThreads = E * 32;
  constexpr int PS = E / 4;            // threads per column
  constexpr int LV = kNz / PS;         // cells per thread
  constexpr int CPW = kCW / PS;        // accumulator columns per warp and chunk
  static_assert(E == 8 || E == 16, "8 or 16 epilogue warps");
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_xn = reinterpret_cast<float*>(smem + SmemLayout::off_xn);
  float* s_acc = reinterpret_cast<float*>(smem + SmemLayout::off_acc);
  float* s_xs = reinterpret_cast<float*>(smem + SmemLayout::off_xs);
  uint8_t* s_ring = smem + SmemLayout::off_ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SmemLayout::off_tmem);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t n_tiles = (a.B + kM - 1) / kM;
  const int stages = a.scheme == OPZ_RK4 ? 4 : (a.scheme == OPZ_RK2 ? 2 : 1);
  const int n_rhs = a.n_steps * stages;
  const uint32_t rank = ptx::cluster_ctarank();
  const int64_t n_items = (n_tiles + 1) / 2;      // one pair of adjacent tiles per cluster and iteration
  const int64_t item0 = blockIdx.x / 2, item_stride = gridDim.x / 2;
  const uint32_t leader_bars = ptx::mapa(ptx::smem_u32(bars), 0u);
  auto arrive_issuer = [&](int id) {              // barriers the MMA issuer waits on live in the leader CTA
    if (rank != 0) ptx::mbar_arrive_remote_relaxed(leader_bars + 8u * (uint32_t)id);
    else ptx::mbar_arrive_relaxed(bars + id);
  };
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kNB; ++s) { ptx::mbar_init(bars + B_FULL + s, rank == 0 ? 2 : 1); ptx::mbar_init(bars + B_EMPTY + s, 1); }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(bars + B_DFULL + b, 1); ptx::mbar_init(bars + B_DEMPTY + b, 2 * E); }
      for (int c = 0; c < 7; ++c) ptx::mbar_init(bars + B_H1 + c, 2 * E);
      ptx::mbar_init(bars + B_H2FULL, 2 * E);
      ptx::mbar_init(bars + B_H2EMPTY, 1);
      ptx::mbar_init(bars + B_YFULL, 1);
      ptx::mbar_init(bars + B_XFULL, 2 * E);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc2<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // both CTAs' barriers exist before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= weight-tape producer (one per CTA: its half of every stage) =================
    if (lane == 0 && n_rhs > 0) {
      uint32_t gs = 0, c = 0;   // stages issued / stages known to be consumed
      for (int64_t item = item0; item < n_items; item += item_stride) {
        for (int r = 0; r < n_rhs; ++r) {
          // version of the weight tape of right-hand side (step n, stage st): (n + st) mod kTapes
          const int n_loc = r / stages, st_loc = r - n_loc * stages;
          const uint8_t* tape = a.tape + (size_t)((a.n0 + n_loc + st_loc) & a.tape_mask) * a.tape_stride;
          for (int s = 0; s < a.n_stages; ++s, ++gs) {
            const uint32_t keep = a.st_keep[s];
            while (c + keep < gs) {  // stage c must have been consumed: its bytes (or its barrier pair) are reused
              ptx::mbar_wait(bars + B_EMPTY + (c % kNB), (c / kNB) & 1u);
              ++c;
            }
            const uint32_t bytes = a.st_bytes[s];
            uint64_t* full = bars + B_FULL + (gs % kNB);
            ptx::mbar_arrive_expect_tx(full, bytes);
            ptx::bulk_g2s(s_ring + a.st_off[s], tape + a.st_src[s] + rank * bytes, bytes, full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank != 0) {  // (warp 2 + E of the peer CTA idles)
    // ================= peer CTA: forward "my half of the stage has landed" to the leader =================
    if (lane == 0 && n_rhs > 0) {
      uint32_t gs = 0;
      for (int64_t item = item0; item < n_items; item += item_stride)
        for (int r = 0; r < n_rhs; ++r)
          for (int s = 0; s < a.n_stages; ++s, ++gs) {
            ptx::mbar_wait(bars + B_FULL + (gs % kNB), (gs / kNB) & 1u);
            ptx::mbar_arrive_remote_relaxed(leader_bars + 8u * (uint32_t)(B_FULL + (gs % kNB)));
          }
    }
    __syncwarp();
  } else if (warp == 1 || warp == kMma2Warp) {
    // ================= MMA issuer(s) (leader CTA) =================
    // With two issuing warps, warp 1 owns the even stages and warp 2 + E the odd ones: both walk the same schedule with the
    // same counters, every barrier phase is computed from those counters (no per-warp toggling), and a warp simply skips
    // the stages of the other.  What orders the two warps' MMAs where it matters: accumulator buffers and H2c go through the
    // epilogue's barriers; the output-layer MMAs all ACCUMULATE into a Y the epilogue zeroed, so their order is free.
    // The whole warp walks the schedule (the loops build_plan() used to lay out the tape) with warp-uniform
    // counters; one elected lane issues the tcgen05.mma / tcgen05.commit instructions.
    const uint32_t me = warp == 1 ? 0u : 1u;
    const uint32_t nw = (uint32_t)a.n_mma_warps;
    if (n_rhs > 0 && me < nw && rank == 0) {
      uint32_t rhs_ctr = 0;   // right-hand sides done: chunk g of right-hand side r has the global index r * n_chunks + g (likewise the stages)
      long long wacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_begin = PROF ? clock64() : 0;
      const uint32_t ring_u32 = ptx::smem_u32(s_ring);
      const uint32_t idesc0 = ptx::idesc_f16kind(2 * kM, 0, FP16);
      const uint32_t idesc_out = idesc0 | ((32u >> 3) << 17);
      const uint32_t nch = (uint32_t)a.n_chunks, nst = (uint32_t)a.n_stages;
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);   // upper half of every B descriptor: SBO = 128 B, version 1
      auto wait_full = [&](uint32_t gs) {
        if constexpr (PROF) {
          const long long t0 = clock64();
          ptx::mbar_wait(bars + B_FULL + (gs % kNB), (gs / kNB) & 1u);
          wacc[W_FULL] += clock64() - t0;
        } else {
          ptx::mbar_wait(bars + B_FULL + (gs % kNB), (gs / kNB) & 1u);
        }
      };
      // issue (elected lane): ks output-layer MMAs Y_net += H2c * W_out[:, chunk]^T, B block of 16 rows per CTA at ring offset boff
      auto out_mmas = [&](uint32_t boff, uint32_t net, uint32_t ks) {
        const uint32_t lo = (((ring_u32 + boff) >> 4) & 0x3FFFu) | (16u << 16);
        const uint32_t dy = tmem + kColY + 32u * net;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < (int)ks) ptx::mma_ts2_lohi(dy, tmem + kColH2 + (uint32_t)k * 8u, lo + (uint32_t)k * 32u, kDescHi, idesc_out, true);   // Y was zeroed by the epilogue
        ptx::tc_commit2(bars + B_H2EMPTY);
      };
      for (int64_t item = item0; item < n_items; item += item_stride) {
#pragma unroll 1
        for (int r = 0; r < n_rhs; ++r, ++rhs_ctr) {
          // Output-layer MMAs of a fused chunk g (Y_net (+)= H2c * W_out[:, chunk]^T) are issued TWO stages later, at the
          // head of the stage of chunk g + 2: that stage waits for "accumulator buffer of chunk g drained" anyway, which is
          // the moment H2c(g) has been written, so they cost no extra blocking wait and no extra issue region.
          const uint32_t g0 = rhs_ctr * nch, gs0 = rhs_ctr * nst;
          wait_phase<PROF>(bars, B_XFULL, rhs_ctr, wacc, W_X);
#pragma unroll 1
          for (uint32_t ci = 0; ci < nch; ++ci) {
            const uint32_t gi = g0 + ci;                     // global index of this accumulator chunk
            if (nw == 2u && (gi & 1u) != me) continue;       // the other issuing warp's chunk (a warp owns one accumulator buffer)
            // ---- everything that does not depend on a barrier: before the waits ----
            const long long t_top = PROF ? clock64() : 0;
            const ChunkRec& c = a.ch[ci];
            const uint32_t buf = gi & 1u;
            const uint32_t rows = c.rows, kin = c.kin;
            const uint32_t gs = gs0 + c.stage;
            const bool split = c.flags & CH_SPLIT;           // two weight stages (see build_plan)
            const bool wait_h1 = c.flags & CH_WAIT_H1;
            const uint32_t more = c.out_ks;                  // output-layer MMAs of chunk g - 2 lead this stage
            const uint32_t out_net = c.out_net, out_off = c.out_off;
            const uint32_t idesc = idesc0 | ((rows >> 3) << 17);
            const uint32_t d = tmem + kColD0 + buf * kCW;
            const uint32_t a0 = tmem + c.a_col;
            const uint32_t a_one = tmem + kColOne;
            // B descriptor of the stage: this CTA holds rows / 2 rows as [k/8][rows/2][8] (LBO = rows / 2 * 16 B)
            const uint32_t blo = (((ring_u32 + c.offA) >> 4) & 0x3FFFu) | ((rows >> 1) << 16);
            const uint32_t bloB = (((ring_u32 + c.offB) >> 4) & 0x3FFFu) | ((uint32_t)(kCW / 2) << 16);
            // bias MMA: D += [1, 0, ..] x [b; don't care]: the block of rows / 2 x 8 halves right behind the weights; its
            // second k-group reads what follows inside the stage (zeros or finite weights), multiplied by the zeros of A
            const uint32_t bias_lo = blo + (rows >> 1) * kin * 2u;   // descriptor address units of 16 B
            const int h1_first = B_H1 + (int)c.h1_base;
            const uint32_t h1_phase = rhs_ctr * (uint32_t)a.h1_mult + c.h1_seq;   // completions of these H1 barriers so far
            const bool static_shape = (rows == (uint32_t)kCW || rows == 16u) && (kin == 25u || kin == 6u);
            // ---- waits: the weights, then "buffer drained" of chunk g - 2 = completion (g >> 1) - 1 of this buffer's barrier
            //      (for a fused chunk also: its H2c is written) ----
            if constexpr (PROF) wacc[W_DECODE] += clock64() - t_top;
            wait_full(gs);
            if (gi >= 2u) wait_phase<PROF>(bars, B_DEMPTY + (int)buf, (gi >> 1) - 1u, wacc, W_DEMPTY);
            const long long t_iss = PROF ? clock64() : 0;
            const long long w_before = PROF ? wacc[W_H1] + wacc[W_FULL] : 0;
            // Only the first chunk of a layer fed by HF paces itself on the H1 barriers (CH_WAIT_H1).  The second chunk (possibly
            // the other issuing warp's) is covered by its D-empty wait: chunk g - 2 is then the LAST first-layer chunk, whose
            // drained arrival comes after its tcgen05.st and after all earlier drains — so D-empty must keep being signalled
            // after the store, not right after the accumulator load.
            if (split) {
              // ---- 64 rows x 25 k-steps in two weight stages, so that the first 12 KB of the ring are free again after
              //      12 MMAs: stage A = k-steps [0, 12) (HF chunks 0..2), stage B = [12, 25) + bias + W_out of chunk g - 2
              if (wait_h1)
                for (int h = 0; h < 3; ++h) wait_phase<PROF>(bars, h1_first + h, h1_phase, wacc, W_H1);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                issue_range<kCW, 0, kSplitK>(d, a0, blo, kDescHi, idesc);
                ptx::tc_commit2(bars + B_EMPTY + (gs % kNB));
              }
              __syncwarp();
              wait_full(gs + 1u);
              if (wait_h1)
                for (int h = 3; h < 7; ++h) wait_phase<PROF>(bars, h1_first + h, h1_phase, wacc, W_H1);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                if (more) out_mmas(out_off, out_net, more);
                issue_range<kCW, kSplitK, 25>(d, a0, bloB - (uint32_t)(kSplitK * kCW), kDescHi, idesc);   // B offsets restart at the stage
                ptx::mma_ts2_lohi(d, a_one, bloB + (kCW / 2) * (uint32_t)(25 - kSplitK) * 2u, kDescHi, idesc, true);
                ptx::tc_commit2(bars + B_DFULL + buf);
                ptx::tc_commit2(bars + B_EMPTY + ((gs + 1u) % kNB));
              }
              __syncwarp();
            } else if (static_shape && !wait_h1) {
              // hot path: the A operand is already complete; one straight immediate-offset run.  (A wider dispatch that also
              // covered the narrow layers' shapes, 12 instantiations in this region, cost construct_NN 3 % and gained the
              // narrow nets nothing: they are bound by the epilogue.)
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                if (more) out_mmas(out_off, out_net, more);
                if (rows == (uint32_t)kCW) {
                  if (kin == 25u) issue_range<kCW, 0, 25>(d, a0, blo, kDescHi, idesc);
                  else issue_range<kCW, 0, 6>(d, a0, blo, kDescHi, idesc);
                } else {
                  if (kin == 25u) issue_range<16, 0, 25>(d, a0, blo, kDescHi, idesc);
                  else issue_range<16, 0, 6>(d, a0, blo, kDescHi, idesc);
                }
                ptx::mma_ts2_lohi(d, a_one, bias_lo, kDescHi, idesc, true);
                ptx::tc_commit2(bars + B_DFULL + buf);
                ptx::tc_commit2(bars + B_EMPTY + (gs % kNB));
              }
              __syncwarp();
            } else if (static_shape && kin == 25u) {
              // first chunk of a 400-wide layer fed by HF: follow the epilogue in three groups of HF chunks
              // (k-steps [0,12) need HF chunks 0..2, [12,20) chunks 3..4, [20,25) chunks 5..6)
#pragma unroll 1
              for (int grp = 0; grp < 3; ++grp) {
                const int h0 = grp == 0 ? 0 : (grp == 1 ? 3 : 5), h1 = grp == 0 ? 3 : (grp == 1 ? 5 : 7);
                for (int h = h0; h < h1; ++h) wait_phase<PROF>(bars, h1_first + h, h1_phase, wacc, W_H1);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                  if (grp == 0 && more) out_mmas(out_off, out_net, more);
                  if (rows == (uint32_t)kCW) issue_hf_group<kCW>(grp, d, a0, blo, kDescHi, idesc);
                  else issue_hf_group<16>(grp, d, a0, blo, kDescHi, idesc);
                  if (grp == 2) {
                    ptx::mma_ts2_lohi(d, a_one, bias_lo, kDescHi, idesc, true);
                    ptx::tc_commit2(bars + B_DFULL + buf);
                    ptx::tc_commit2(bars + B_EMPTY + (gs % kNB));
                  }
                }
                __syncwarp();
              }
            } else {
              if (more) {
                ptx::tc_fence_after();
                if (ptx::elect_one()) out_mmas(out_off, out_net, more);
                __syncwarp();
              }
#pragma unroll
              for (int h = 0; h < kMaxKs / 4; ++h) {
                if (h * 4 < (int)kin) {
                  if (wait_h1) wait_phase<PROF>(bars, h1_first + h, h1_phase, wacc, W_H1);
                  ptx::tc_fence_after();  // orders the MMAs after the epilogue's tcgen05.st seen through the barriers above
                  if (ptx::elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                      const int k = h * 4 + kk;
                      if (k < (int)kin) ptx::mma_ts2_lohi(d, a0 + (uint32_t)k * 8u, blo + (uint32_t)k * rows, kDescHi, idesc, k > 0);
                    }
                  }
                  __syncwarp();
                }
              }
              if (ptx::elect_one()) {
                ptx::mma_ts2_lohi(d, a_one, bias_lo, kDescHi, idesc, true);
                ptx::tc_commit2(bars + B_DFULL + buf);
                ptx::tc_commit2(bars + B_EMPTY + (gs % kNB));
              }
              __syncwarp();
            }
            if constexpr (PROF) wacc[W_ISSUE] += (clock64() - t_iss) - (wacc[W_H1] + wacc[W_FULL] - w_before);   // issue region without the waits inside it
          }
          // tail stage: the output-layer MMAs of the last two chunks, each behind the "drained" arrival of its buffer
          const uint32_t g = g0 + nch;
          if (nw == 1u || (g & 1u) == me) {
            const uint32_t gs = gs0 + nst - 1u;
            wait_full(gs);
            uint32_t off = a.tail_off;
            wait_phase<PROF>(bars, B_DEMPTY + (int)((g - 2u) & 1u), (g - 2u) >> 1, wacc, W_H2FULL);   // chunk g - 2
            if (a.tail_ks[0] > 0) {
              ptx::tc_fence_after();
              if (ptx::elect_one()) out_mmas(off, a.tail_net[0], a.tail_ks[0]);
              __syncwarp();
              off += (uint32_t)a.tail_ks[0] * 512u;
            }
            wait_phase<PROF>(bars, B_DEMPTY + (int)((g - 1u) & 1u), (g - 1u) >> 1, wacc, W_H2FULL);   // chunk g - 1
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              out_mmas(off, a.tail_net[1], a.tail_ks[1]);
              ptx::tc_commit2(bars + B_YFULL);
              ptx::tc_commit2(bars + B_EMPTY + (gs % kNB));
            }
            __syncwarp();
          }
        }
      }
      if (PROF && blockIdx.x == 0 && lane == 0 && me == 0u) {
        wacc[W_TOTAL] = clock64() - t_begin;
        for (int i = 0; i < 8; ++i) g_tc2_wait[i] = (unsigned long long)wacc[i];
      }
    }
  } else {
    // ================= epilogue + physics =================
    // warp -> TMEM lane quadrant q (hardware: warp % 4) and sub-index ps; thread (q, lane) = ocean column m;
    // the PS threads of a column split the accumulator columns of every chunk and the cells of the physics.
    const int e = warp - 2;
    const int q = warp % 4;
    const int ps = e / 4;
    const int m = q * 32 + lane;
    const int etid = e * 32 + lane;
    const int c0 = ps * LV;            // first cell this thread owns
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t mask = 1u << B_H2EMPTY;   // released state
    uint32_t g = 0;
    const float h = a.dt;
    const DevParams& P = a.P;
    const bool prof_me = PROF && blockIdx.x == 0 && warp == 2;
    long long eacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = PROF ? clock64() : 0;
    auto epi_bar_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); };

    if (ps == 0) {  // the "ones" k-step of this thread's TMEM lane: [1, 0, ..., 0] as 16-bit pairs (ordered before the MMAs by X-full)
      const uint32_t one[8] = {FP16 ? 0x00003C00u : 0x00003F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      ptx::tmem_st8(tmem + lane_addr + kColOne, one);
      ptx::tmem_wait_st();
    }
    for (int64_t item = item0; item < n_items; item += item_stride) {
      const int64_t tile = item * 2 + rank;   // may be one past the last tile for the peer: a padding tile, never stored
      const int64_t col0 = tile * kM;
      const int64_t col = col0 + m;
      // ---- load the initial profiles (coalesced through the padded state array) ----
      epi_bar_sync();  // previous tile's snapshot readers are done with s_xn
      for (int ee = etid; ee < kM * kS; ee += kEpiThreads) {
        const int mm = ee / kS, i = ee - mm * kS;
        const int64_t c = col0 + mm;
        s_xn[xn_idx(i, mm)] = c < a.B ? a.x0[c * kS + i] : 0.0f;
      }
      float bc[OPZ_BC_STRIDE];
#pragma unroll
      for (int qq = 0; qq < OPZ_BC_STRIDE; ++qq) bc[qq] = col < a.B ? a.bc[col * OPZ_BC_STRIDE + qq] : 0.0f;
      epi_bar_sync();
      auto save = [&](int slot_idx) {
        epi_bar_sync();  // every column's s_xn is final
        for (int ee = etid; ee < kM * kS; ee += kEpiThreads) {
          const int mm = ee / kS, i = ee - mm * kS;
          const int64_t c = col0 + mm;
          if (c < a.B) a.out[(c * a.n_save + slot_idx) * kS + i] = s_xn[xn_idx(i, mm)];
        }
      };
      if (a.save_every > 0) save(0);
      // 16-bit image of this thread's cells of the stage state -> tensor memory (A operand of the first layers)
      // `t_rhs` = global index of the right-hand side that will consume the image: the phase of the rounding dither
      auto publish_x = [&](const float (&xv)[3][LV], uint32_t t_rhs) {
        // Weyl-sequence thresholds on the dropped mantissa bits (13 for IEEE half, 16 for BF16), eight per right-hand side;
        // magnitude bits + threshold, then truncation = a rounding whose mean over consecutive right-hand sides is the FP32 value
        uint32_t thr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) thr[j] = a.dither_x ? (((t_rhs * 5061u + (uint32_t)j * 1021u) & 8191u) << (FP16 ? 0 : 3)) : (FP16 ? 4096u : 32768u);
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          uint32_t pk[LV / 2];
#pragma unroll
          for (int j = 0; j < LV / 2; ++j) {
            const float e0 = __uint_as_float(__float_as_uint(xv[v][2 * j]) + thr[(2 * j) & 7]);
            const float e1 = __uint_as_float(__float_as_uint(xv[v][2 * j + 1]) + thr[(2 * j + 1) & 7]);
            pk[j] = FP16 ? ptx::pack_f16x2_rz(e0, e1) : ptx::pack_bf16x2_rz(e0, e1);
          }
          const uint32_t dst = tmem + lane_addr + kColX + (uint32_t)(v * (kNz / 2) + c0 / 2);
          if constexpr (LV == 16) ptx::tmem_st8(dst, pk);
          else ptx::tmem_st4(dst, pk);
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_issuer(B_XFULL);
      };
      {
        float xv[3][LV];
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
          for (int i = 0; i < LV; ++i) {
            const int idx = v * kNz + c0 + i;
            xv[v][i] = s_xn[xn_idx(idx, m)];
            s_xs[idx * kM + m] = xv[v][i];
          }
        if (n_rhs > 0) publish_x(xv, (uint32_t)(a.n0 * stages));
      }

      int rhs_idx = 0;
#pragma unroll 1
      for (int n = 0; n < a.n_steps; ++n) {
        const float tn = a.t0 + (float)n * h;
#pragma unroll 1
        for (int st = 0; st < stages; ++st, ++rhs_idx) {
          // ---- drain the accumulator chunks of the hidden layers (the schedule table, see ChunkRec) ----
#pragma unroll 1
          for (int ci = 0; ci < a.n_chunks; ++ci) {
            const ChunkRec& c = a.ch[ci];
            const uint32_t buf = g & 1u;
            ++g;
            const bool fused = c.flags & CH_FUSED;
            const int ncols = (int)c.rows - ps * CPW;   // columns of this chunk that belong to this warp (<= 0: none)
            const uint32_t src = tmem + lane_addr + kColD0 + buf * kCW + (uint32_t)(ps * CPW);
            const uint32_t dst = tmem + lane_addr + (uint32_t)c.dst_col + (uint32_t)(ps * CPW / 2);
            const int h1_bar = B_H1 + (int)c.h1_bar;
            if (prof_me) bar_wait_t<PROF>(bars, B_DFULL + buf, mask, eacc, E_DFULL - 8);
            else bar_wait(bars, B_DFULL + buf, mask);
            ptx::tc_fence_after();
            const long long t_ld = prof_me ? clock64() : 0;
            uint32_t rr[CPW / 16][16];
            uint32_t pk[CPW / 16][8];
            if (ncols >= CPW) {
              // the usual case, free of predicates: this warp's full share of a kCW-column chunk
#pragma unroll
              for (int grp = 0; grp < CPW / 16; ++grp) ptx::tmem_ld16(src + grp * 16, rr[grp]);
              ptx::tmem_wait_ld();
              if (prof_me) eacc[E_LD - 8] += clock64() - t_ld;
#pragma unroll
              for (int grp = 0; grp < CPW / 16; ++grp)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)   // the bias is already in the accumulator (bias MMA)
                  pk[grp][jj] = pack_act<FP16, ACT>(__uint_as_float(rr[grp][2 * jj]), __uint_as_float(rr[grp][2 * jj + 1]));
              if (fused) {  // the output-layer MMAs of the previous chunk must have consumed H2c
                if (prof_me) bar_wait_t<PROF>(bars, B_H2EMPTY, mask, eacc, E_H2EMPTY - 8);
                else bar_wait(bars, B_H2EMPTY, mask);
                ptx::tc_fence_after();
              }
#pragma unroll
              for (int grp = 0; grp < CPW / 16; ++grp) ptx::tmem_st8(dst + grp * 8, pk[grp]);
            } else {
              // narrow last chunk of a layer: groups of 16 columns, predicated
#pragma unroll
              for (int grp = 0; grp < CPW / 16; ++grp)
                if (grp * 16 < ncols) ptx::tmem_ld16(src + grp * 16, rr[grp]);
              ptx::tmem_wait_ld();
              if (prof_me) eacc[E_LD - 8] += clock64() - t_ld;
#pragma unroll
              for (int grp = 0; grp < CPW / 16; ++grp) {
                if (grp * 16 < ncols) {
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj)
                    pk[grp][jj] = pack_act<FP16, ACT>(__uint_as_float(rr[grp][2 * jj]), __uint_as_float(rr[grp][2 * jj + 1]));
                }
              }
              if (fused) {
                if (prof_me) bar_wait_t<PROF>(bars, B_H2EMPTY, mask, eacc, E_H2EMPTY - 8);
                else bar_wait(bars, B_H2EMPTY, mask);
                ptx::tc_fence_after();
              }
#pragma unroll
              for (int grp = 0; grp < CPW / 16; ++grp)
                if (grp * 16 < ncols) ptx::tmem_st8(dst + grp * 8, pk[grp]);
            }
            const long long t_st = prof_me ? clock64() : 0;
            if (ci == 0) {
              // first chunk of a right-hand side: zero this thread's share of Y (every output-layer MMA accumulates).  All
              // threads have read the previous Y (this chunk's MMAs were issued behind the X-full barrier), and the first
              // output-layer MMA waits for a LATER chunk's drained arrival of these same warps.
              const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
              for (int z = 0; z < 96 / PS; z += 8) ptx::tmem_st8(tmem + lane_addr + kColY + (uint32_t)(ps * (96 / PS) + z), zero);
            }
            ptx::tmem_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              arrive_issuer(B_DEMPTY + buf);          // accumulator drained AND (fused chunk) H2c written
              if (!fused) arrive_issuer(h1_bar);
            }
            if (prof_me) eacc[E_ST - 8] += clock64() - t_st;
          }
          // ---- physics, part A: everything that does not need the NN outputs (runs while the last output-layer
          //      MMAs are still in flight).  Index i of the x arrays is cell c0 - 1 + i. ----
          epi_bar_sync();  // the stage state written by the other warps of this tile (previous right-hand side) is visible
          const long long t_a = PROF ? clock64() : 0;
          float Du[LV + 1], Dv[LV + 1], DT[LV + 1];   // diffusive part of the flux on faces c0 .. c0 + LV
          float cu[LV], cv[LV];                        // Coriolis terms of the owned cells
          // one copy per first-cell index when a column is split over two threads (c0 = 0 or 16): cell clamps, bias indices
          // and shared-memory offsets become immediates
          auto phase_a = [&](auto c0c) {
            const int c0s = decltype(c0c)::value >= 0 ? decltype(c0c)::value : c0;
            // branch-free on purpose (clamped loads, selects): the LV + 1 faces are independent dependency chains that the
            // scheduler can only interleave inside one basic block; values computed for a boundary face are discarded
            // x index i is cell c0 - XO + i (clamped into the column: the differences across a boundary face come out zero,
            // which is what the first / last row of D^f gives)
            constexpr int XO = SMOOTH ? 2 : 1;
            float xu[LV + 2 * XO], xw[LV + 2 * XO], xT[LV + 2 * XO];
#pragma unroll
            for (int i = 0; i < LV + 2 * XO; ++i) {
              const int cell = min(max(c0s - XO + i, 0), kNz - 1);
              xu[i] = s_xs[cell * kM + m];
              xw[i] = s_xs[(kNz + cell) * kM + m];
              xT[i] = s_xs[(2 * kNz + cell) * kM + m];
            }
            const bool mpp = P.flags & OPZ_FLAG_MPP;
            const bool ca = mpp && (P.flags & OPZ_FLAG_CA);
            const bool ca_du = P.flags & OPZ_FLAG_CA_SELECT_DUDZ;
            // smooth_Ri: raw Richardson numbers of faces c0 - 1 .. c0 + LV + 1 (index g is face c0 - 1 + g), filtered below
            float Rr[SMOOTH ? LV + 3 : 1];
            const bool smooth_ri = SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_RI);
            if constexpr (SMOOTH) {
#pragma unroll
              for (int g2 = 0; g2 < LV + 3; ++g2) {
                const float Bz = fmaf(a.pk[PK_KBZ], xT[g2 + 1] - xT[g2], a.pk[PK_EBZ]);
                const float sa = fmaf(a.pk[PK_KSA], xu[g2 + 1] - xu[g2], a.pk[PK_ESA]);
                const float sb = fmaf(a.pk[PK_KSB], xw[g2 + 1] - xw[g2], a.pk[PK_ESB]);
                Rr[g2] = Bz * rcp_ftz(sa * sa + sb * sb);
              }
            }
#pragma unroll
            for (int i = 0; i <= LV; ++i) {
              const int f = c0s + i;
              const int jb = max(f - 1, 0);   // NN output (and its bias) on this face
              const float du = xu[i + XO] - xu[i + XO - 1];
              const float dv = xw[i + XO] - xw[i + XO - 1];
              const float dT = xT[i + XO] - xT[i + XO - 1];
              // = richardson<true>, nu_of_ri<true>, nuT_of<true> of opz_device.cuh with the flag tests hoisted and the constant
              // factors folded (Args::pk): Bz = cBz (Nz dT + eps), sa = s_u (Nz du + eps), sb likewise
              const float Bz = fmaf(a.pk[PK_KBZ], dT, a.pk[PK_EBZ]);
              const float sa = fmaf(a.pk[PK_KSA], du, a.pk[PK_ESA]);
              const float sb = fmaf(a.pk[PK_KSB], dv, a.pk[PK_ESB]);
              // flush-to-zero approximations without range fix-ups: 0 / 0 stays NaN (0 * inf), the exponent is clamped so that
              // e^{..} stays finite (nu -> nu0 in that tail to FP32 precision anyway)
              float Ri = Bz * rcp_ftz(sa * sa + sb * sb);
              if constexpr (SMOOTH) { if (smooth_ri) Ri = (Rr[i] + Rr[i + 1] + Rr[i + 2]) * (1.0f / 3.0f); }
              const float nu_raw = P.nu0 + P.nu_m * rcp_ftz(1.0f + ex2_ftz(fminf(fmaf(Ri, a.pk[PK_C2], a.pk[PK_R2]), 120.0f)));
              const float nu = mpp ? nu_raw : 0.0f;   // a select, not a factor: a NaN Ri must not leak when mPP is off
              const float sel = ca_du ? du : dT;   // Nz > 0: the raw difference has the gradient's sign
              const float nuT = (ca && !(sel > 0.0f)) ? P.kappa : nu * P.inv_Pr;
              // diffusive flux minus the output-layer bias: the total flux of an interior face is then y_raw - D (the values
              // computed for a boundary face are discarded by the flux select below).  With smooth_NN the bias is added to the
              // raw outputs before the filter instead (part B).
              const bool fold_bias = !(SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_NN));
              Du[i] = fmaf(a.pk[PK_CUN] * nu, du, fold_bias ? -a.b3[jb] : 0.0f);
              Dv[i] = fmaf(a.pk[PK_CVN] * nu, dv, fold_bias ? -a.b3[32 + jb] : 0.0f);
              DT[i] = fmaf(a.pk[PK_CTN] * nuT, dT, fold_bias ? -a.b3[64 + jb] : 0.0f);
            }
#pragma unroll
            for (int i = 0; i < LV; ++i) {
              cu[i] = fmaf(a.pk[PK_CCU], xw[i + XO], a.pk[PK_DCU]);   //  cor_u (s_v v + mu_v)
              cv[i] = fmaf(a.pk[PK_CCV], xu[i + XO], a.pk[PK_DCV]);   // -cor_v (s_u u + mu_u)
            }
                    };
          if constexpr (LV == 16) {
            if (c0 == 0) phase_a(std::integral_constant<int, 0>{});
            else phase_a(std::integral_constant<int, 16>{});
          } else {
            phase_a(std::integral_constant<int, -1>{});
          }
          float ts = tn;
          if (a.scheme == OPZ_RK4) ts = (st == 0) ? tn : (st == 3 ? tn + h : tn + h * 0.5f);
          else if (a.scheme == OPZ_RK2) ts = (st == 0) ? tn : tn + h;
          const float Fb[3] = {bc[0] - P.off[0], bc[2] - P.off[1], bc[4] - P.off[2]};
          const float Ft[3] = {bc[1] - P.off[0], bc[3] - P.off[1], wT_top_value(P, bc, ts)};
          // ---- NN outputs ----
          if (prof_me) { eacc[E_PHYS_A - 8] += clock64() - t_a; bar_wait_t<PROF>(bars, B_YFULL, mask, eacc, E_YFULL - 8); }
          else bar_wait(bars, B_YFULL, mask);
          ptx::tc_fence_after();
          const long long t_b = PROF ? clock64() : 0;
          // Runge-Kutta bookkeeping in the oracle's operation order: acc = k1 (+ 2 k2 + 2 k3);
          // stage input xs = xn + cs k; last stage xn += cf (acc + k)
          const bool first = (st == 0), last = (st == stages - 1);
          const float wk = (a.scheme == OPZ_RK4 && !first && !last) ? 2.0f : 1.0f;
          const float cs = (a.scheme == OPZ_RK4 && st < 2) ? h * 0.5f : h;
          const float cf = a.scheme == OPZ_RK4 ? h / 6.0f : (a.scheme == OPZ_RK2 ? h * 0.5f : h);
          const float cx = last ? cf : cs;
          float xnew[3][LV];
          epi_bar_sync();  // every thread has read the old stage state of its neighbours' cells
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            // total flux on faces c0 .. c0 + LV: face f > 0 interior takes NN output j = f - 1 (TMEM column j of Y_v)
            float F[LV + 1];
            {
              const uint32_t ybase = tmem + lane_addr + kColY + (uint32_t)(v * 32);
              uint32_t ylo = 0u, yhi[LV];
              if (c0 > 0) ylo = ptx::tmem_ld1(ybase + (uint32_t)(c0 - 1));   // output c0 - 1 (face c0)
              if constexpr (LV == 16) ptx::tmem_ld16(ybase + (uint32_t)c0, yhi);
              else ptx::tmem_ld8(ybase + (uint32_t)c0, yhi);
              uint32_t ylo2 = 0u, ytop = 0u;   // smooth_NN: one more output on either side (j = c0 - 2 and j = c0 + LV)
              if constexpr (SMOOTH) {
                if (c0 > 1) ylo2 = ptx::tmem_ld1(ybase + (uint32_t)(c0 - 2));
                if (c0 + LV < kNz) ytop = ptx::tmem_ld1(ybase + (uint32_t)(c0 + LV));
              }
              ptx::tmem_wait_ld();
              const float* D = v == 0 ? Du : (v == 1 ? Dv : DT);
              if (SMOOTH && (P.flags & OPZ_FLAG_SMOOTH_NN)) {
                // yy[q] = raw output j = c0 - 2 + q plus its bias, q = 0 .. LV + 2; face c0 + i takes the filtered output j = c0 + i - 1
                float yy[LV + 3];
                constexpr int n = kNz - 1;
#pragma unroll
                for (int q2 = 0; q2 < LV + 3; ++q2) {
                  const int j = c0 - 2 + q2;
                  const uint32_t raw = q2 == 0 ? ylo2 : (q2 == 1 ? ylo : (q2 == LV + 2 ? ytop : yhi[q2 >= 2 && q2 < LV + 2 ? q2 - 2 : 0]));
                  yy[q2] = (j >= 0 && j < n) ? __uint_as_float(raw) + a.b3[v * 32 + min(max(j, 0), n - 1)] : 0.0f;
                }
#pragma unroll
                for (int i = 0; i <= LV; ++i) {
                  const int j = c0 + i - 1;
                  float y;
                  if (j <= 0) y = (yy[i + 1] + yy[i + 2]) * 0.5f;
                  else if (j >= n - 1) y = (yy[i] + yy[i + 1]) * 0.5f;
                  else y = (yy[i] + yy[i + 1] + yy[i + 2]) * (1.0f / 3.0f);
                  F[i] = (i == 0 && c0 == 0) ? Fb[v] : ((i == LV && c0 + LV == kNz) ? Ft[v] : y - D[i]);
                }
              } else {
#pragma unroll
                for (int i = 0; i <= LV; ++i) {
                  const float y = i == 0 ? __uint_as_float(ylo) : __uint_as_float(yhi[i == 0 ? 0 : i - 1]);
                  F[i] = (i == 0 && c0 == 0) ? Fb[v] : ((i == LV && c0 + LV == kNz) ? Ft[v] : y - D[i]);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < LV; ++i) {
              const float dF = F[i + 1] - F[i];
              const float k = v == 0 ? fmaf(a.pk[PK_AU], dF, cu[i]) : (v == 1 ? fmaf(a.pk[PK_AV], dF, cv[i]) : a.pk[PK_AT] * dF);
              const int idx = v * kNz + c0 + i;
              const int o = idx * kM + m, on = xn_idx(idx, m);
              const float xo = s_xn[on];
              const float acc_in = first ? 0.0f : s_acc[o];
              const float sum = acc_in + wk * k;       // first stage: 0 + 1*k == k exactly
              const float xv = xo + cx * (last ? sum : k);
              s_acc[o] = sum;
              s_xs[o] = xv;
              if (last) s_xn[on] = xv;
              xnew[v][i] = xv;
            }
          }
          if (rhs_idx + 1 < n_rhs) publish_x(xnew, (uint32_t)(a.n0 * stages + rhs_idx + 1));  // also tells the MMA warp that Y has been consumed
          if (prof_me) eacc[E_PHYS_B - 8] += clock64() - t_b;
        }
        if (a.save_every > 0 && (n + 1) % a.save_every == 0) save((n + 1) / a.save_every);
      }
      if (a.save_every <= 0) save(0);
    }
    if (prof_me && lane == 0) {
      eacc[E_TOTAL - 8] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) g_tc2_wait[8 + i] = (unsigned long long)eacc[i];
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // the peer's tensor / shared memory stay valid until both CTAs are done
  if (warp == 1) ptx::tmem_dealloc2<512>(tmem);
}

// ---- host: weight tape in stage order -----------------------------------------------------------------
static uint16_t f16_rne(float f) {  // IEEE binary16, round to nearest even, saturating to +-65504
  uint32_t u;
  std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t absu = u & 0x7FFFFFFFu;
  if (absu > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);
  if (absu >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);
  if (absu < 0x33000001u) return (uint16_t)sign;
  const int e = (int)(absu >> 23) - 127;
  uint32_t mant = (absu & 0x7FFFFFu) | 0x800000u;
  int shift = e >= -14 ? 13 : 13 + (-14 - e);
  uint32_t half_m = mant >> shift;
  const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_m & 1u))) ++half_m;
  uint32_t out;
  if (e >= -14) out = ((uint32_t)(e + 15) << 10) + (half_m - 0x400u);
  else out = half_m;
  return (uint16_t)(sign | out);
}
static uint16_t bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static float f16_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  uint32_t u;
  if (e == 0) {
    if (m == 0) u = sign;
    else { int sh = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; ++sh; } u = sign | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3FFu) << 13); }
  } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
  else u = sign | ((e + 112u) << 23) | (m << 13);
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static float bf16_to_float(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
// neighbours of a 16-bit pattern on the real line (sign-magnitude formats)
static uint16_t next_below(uint16_t b) { return b == 0x0000u ? 0x8001u : ((b & 0x8000u) ? (uint16_t)(b + 1) : (uint16_t)(b - 1)); }
static uint16_t next_above(uint16_t b) { return b == 0x8000u ? 0x0001u : ((b & 0x8000u) ? (uint16_t)(b - 1) : (uint16_t)(b + 1)); }

// Temporally dithered rounding of one weight (see the header): version `tape` of `n_tapes` takes the 16-bit neighbour above the
// weight when the weight's position between its two neighbours exceeds that version's threshold.  `key` identifies the weight
// (net, layer, input, output) and rotates the threshold order, so that the weights of one dot product do not switch in phase.
// The same function is restated in oracle/nde_oracle.py (TensorPathEmulation) for the emulating parity tests.
static uint32_t dither_key(int net, int layer, int k, int n) {
  return ((uint32_t)k * 2654435761u) ^ ((uint32_t)n * 2246822519u) ^ ((uint32_t)(net * 8 + layer) * 3266489917u);
}
static uint16_t quant16(float v, bool fp16, int tape, int n_tapes, uint32_t key) {
  const uint16_t q = fp16 ? f16_rne(v) : bf16_rne(v);
  if (n_tapes <= 1 || !std::isfinite(v)) return q;
  const uint16_t expo = fp16 ? (q & 0x7C00u) : (q & 0x7F80u);
  if (expo == (fp16 ? 0x7C00u : 0x7F80u) || (fp16 && std::fabs(v) >= 65504.0f)) return q;   // saturated / non-finite
  auto val = [&](uint16_t b) { return (double)(fp16 ? f16_to_float(b) : bf16_to_float(b)); };
  uint16_t lo = q;
  if (val(q) > (double)v) lo = next_below(q);
  const uint16_t hi = next_above(lo);
  const double dlo = val(lo), dhi = val(hi);
  if (!(dhi > dlo) || !std::isfinite(dhi)) return q;
  const double frac = ((double)v - dlo) / (dhi - dlo);
  int bits = 0;
  while ((1 << bits) < n_tapes) ++bits;
  const uint32_t rot = (key >> 13) & (uint32_t)(n_tapes - 1);
  const uint32_t idx = ((uint32_t)tape + rot) & (uint32_t)(n_tapes - 1);
  uint32_t rev = 0;
  for (int b = 0; b < bits; ++b) rev |= ((idx >> b) & 1u) << (bits - 1 - b);
  const double thr = ((double)rev + 0.5) / (double)n_tapes;
  return frac > thr ? hi : lo;
}

struct TapeQuant {   // what append_half needs to know about the version being built
  bool fp16;
  int tape, n_tapes, net, layer;
};

// One CTA's half (rows [r0, r0 + hrows)) of the K-major no-swizzle block of W (Flux layout: element (k_in, n_out) at
// w[k * nout + n]) over inputs [k0, k0 + 16 * ksteps): [k/8][row][k%8] 16-bit, zero padded; appended to `dst`.
static void append_half(std::vector<uint16_t>& dst, const TapeQuant& tq, const float* w, int nin, int nout, int r0, int hrows, int k0, int ksteps) {
  const size_t base = dst.size();
  dst.resize(base + (size_t)hrows * ksteps * 16, 0);
  for (int kk = 0; kk < ksteps * 16; ++kk) {
    const int k = k0 + kk;
    if (k >= nin) continue;
    for (int r = 0; r < hrows; ++r) {
      const int n = r0 + r;
      if (n >= nout) continue;
      const float v = w[(size_t)k * nout + n];
      dst[base + ((size_t)(kk / 8) * hrows + r) * 8 + (kk % 8)] = quant16(v, tq.fp16, tq.tape, tq.n_tapes, dither_key(tq.net, tq.layer, k, n));
    }
  }
}

// builds version `tape_idx` of `n_tapes` of the weight tape (the schedule, offsets and sizes are the same for every version)
static int build_plan(const opz_model* m, const std::vector<float>& w, bool fp16, int tape_idx, int n_tapes, Plan& pl, std::vector<uint16_t>& tape) {
  const NetDesc& nd = m->nd;
  const int L = nd.n_layers;
  if (m->nz != kNz) { set_error("tensor-core path: nz must be 32"); return OPZ_EUNSUP; }
  if (L != 2 && L != 3) { set_error("tensor-core path: 2 or 3 Dense layers per net"); return OPZ_EUNSUP; }
  pl.L = L;
  pl.act = nd.act;
  pl.fp16 = fp16;
  for (int j = 0; j < L - 1; ++j) {
    pl.hpad[j] = (nd.sizes[j + 1] + 15) / 16 * 16;
    if (pl.hpad[j] > kMaxHidden) { set_error("tensor-core path: hidden width above 400"); return OPZ_EUNSUP; }
  }
  for (int i = 0; i < 96; ++i) pl.b3[i] = 0.0f;
  tape.clear();
  pl.n_stages = 0;
  pl.n_chunks = 0;
  std::vector<uint16_t> half[2];
  // output-layer blocks ride two stages behind their fused chunk (see the MMA issuer): q1 goes into the coming stage
  struct Pending { bool valid; int net, c, cols; } q1{false, 0, 0, 0}, q0{false, 0, 0, 0};
  bool overflow = false;
  auto out_halves = [&](const Pending& p) {   // 16 rows per CTA of W_out over the 64 (or fewer) inputs of chunk p.c
    const float* wn = w.data() + (size_t)p.net * nd.P;
    const TapeQuant tq{fp16, tape_idx, n_tapes, p.net, L - 1};
    for (int hf = 0; hf < 2; ++hf)
      append_half(half[hf], tq, wn + nd.w_off[L - 1], nd.sizes[L - 1], nd.sizes[L], hf * 16, 16, p.c * kCW, p.cols / 16);
  };
  auto close_stage = [&]() {
    if (pl.n_stages >= kMaxStages) { overflow = true; half[0].clear(); half[1].clear(); return; }
    const int s = pl.n_stages++;
    pl.st_bytes[s] = (uint32_t)(half[0].size() * 2);
    pl.st_src[s] = (uint32_t)(tape.size() * 2);
    tape.insert(tape.end(), half[0].begin(), half[0].end());
    tape.insert(tape.end(), half[1].begin(), half[1].end());
    half[0].clear();
    half[1].clear();
  };
  // narrow three-layer nets: layer-major schedule (three first-layer HF images of <= 64 columns fit the 200-column region,
  // 3 x 2 H1 barriers fit the 7 there are)
  pl.interleave = (L == 3) && pl.hpad[0] <= 128 && getenv("OPZ_TC_NO_INTERLEAVE") == nullptr;   // measured: small 50/20 mish net 2.63e8 -> 2.91e8 col-steps/s
  for (int net = 0; net < 3; ++net) {
    const float* wn = w.data() + (size_t)net * nd.P;
    for (int i = 0; i < nd.sizes[L]; ++i) pl.b3[net * 32 + i] = wn[nd.b_off[L - 1] + i];
  }
  for (int idx = 0; idx < 3 * (L - 1); ++idx) {
    const int net = pl.interleave ? idx % 3 : idx / (L - 1);
    const float* wn = w.data() + (size_t)net * nd.P;
    {
      const int j = pl.interleave ? idx / 3 : idx % (L - 1);
      const bool fused = (j == L - 2);
      const int kin = (j == 0) ? kS : pl.hpad[j - 1];
      const int hp = pl.hpad[j];
      const TapeQuant tq{fp16, tape_idx, n_tapes, net, j};
      const int nc1 = (pl.hpad[0] + kCW - 1) / kCW;
      const int h1_base = pl.interleave ? net * nc1 : 0;
      const uint32_t hf_base = kColHF + (pl.interleave ? (uint32_t)(net * (pl.hpad[0] / 2)) : 0u);
      for (int c = 0; c * kCW < hp; ++c) {
        const int rows = std::min(kCW, hp - c * kCW);
        const bool split = (rows == kCW) && (kin / 16 == 25);   // two stages: k-steps [0, kSplitK) | [kSplitK, 25) + bias + W_out
        if (pl.n_chunks >= kMaxChunks) { overflow = true; break; }
        ChunkRec& cr = pl.ch[pl.n_chunks++];
        cr = ChunkRec{};
        cr.a_col = (uint16_t)(j == 0 ? kColX : hf_base);
        cr.dst_col = (uint16_t)(fused ? kColH2 : hf_base + (uint32_t)c * (kCW / 2));
        cr.rows = (uint8_t)rows;
        cr.kin = (uint8_t)(kin / 16);
        cr.stage = (uint8_t)pl.n_stages;
        cr.flags = (uint8_t)((split ? CH_SPLIT : 0) | (fused ? CH_FUSED : 0) | ((j > 0 && (c == 0 || pl.interleave)) ? CH_WAIT_H1 : 0));
        cr.h1_bar = (uint8_t)(h1_base + c);
        cr.h1_base = (uint8_t)h1_base;
        cr.h1_seq = (uint8_t)(pl.interleave ? 0 : net);
        cr.out_ks = (uint8_t)(q1.valid ? q1.cols / 16 : 0);
        cr.out_net = (uint8_t)q1.net;
        if (split) {
          for (int hf = 0; hf < 2; ++hf)
            append_half(half[hf], tq, wn + nd.w_off[j], nd.sizes[j], nd.sizes[j + 1], c * kCW + hf * (rows / 2), rows / 2, 0, kSplitK);
          close_stage();
          for (int hf = 0; hf < 2; ++hf)
            append_half(half[hf], tq, wn + nd.w_off[j], nd.sizes[j], nd.sizes[j + 1], c * kCW + hf * (rows / 2), rows / 2, kSplitK * 16,
                        25 - kSplitK);
        } else {
          for (int hf = 0; hf < 2; ++hf)
            append_half(half[hf], tq, wn + nd.w_off[j], nd.sizes[j], nd.sizes[j + 1], c * kCW + hf * (rows / 2), rows / 2, 0, kin / 16);
        }
        for (int hf = 0; hf < 2; ++hf)   // bias block: one k-group, row r = [b_r, 0, 0, 0, 0, 0, 0, 0]
          for (int r = 0; r < rows / 2; ++r) {
            const int n = c * kCW + hf * (rows / 2) + r;
            const float bv = n < nd.sizes[j + 1] ? wn[nd.b_off[j] + n] : 0.0f;
            half[hf].push_back(quant16(bv, fp16, tape_idx, n_tapes, dither_key(net, j, 0xFFFF, n)));
            for (int z = 0; z < 7; ++z) half[hf].push_back(0);
          }
        if (q1.valid) out_halves(q1);
        else for (int hf = 0; hf < 2; ++hf) half[hf].resize(half[hf].size() + (size_t)(rows / 2) * 8, 0);  // the bias MMA's second k-group stays inside the stage
        close_stage();
        q1 = q0;
        q0 = Pending{fused, net, c, rows};
      }
    }
  }
  if (q1.valid) out_halves(q1);
  out_halves(q0);   // the last chunk of a net is always a fused one
  close_stage();
  pl.tail_ks[0] = (uint8_t)(q1.valid ? q1.cols / 16 : 0);
  pl.tail_net[0] = (uint8_t)q1.net;
  pl.tail_ks[1] = (uint8_t)(q0.cols / 16);
  pl.tail_net[1] = (uint8_t)q0.net;
  if (overflow) { set_error("tensor-core path: schedule longer than the stage table"); return OPZ_EUNSUP; }
  pl.tape_bytes = tape.size() * 2;
  // ---- ring placement (byte FIFO, every right-hand side starts at offset 0) and in-flight allowance ----
  const int n = pl.n_stages;
  uint32_t cur = 0;
  for (int s = 0; s < n; ++s) {
    const uint32_t sz = (pl.st_bytes[s] + 127u) / 128u * 128u;
    if (sz > (uint32_t)kRingBytes) { set_error("tensor-core path: weight stage larger than the ring"); return OPZ_EUNSUP; }
    if (cur + sz > (uint32_t)kRingBytes) cur = 0;
    pl.st_off[s] = cur;
    cur += sz;
  }
  for (int s = 0; s < n; ++s) {
    // walk back over the cyclic stage sequence: the stage can be filled while the `keep` previous stages are in flight
    int keep = 0;
    const uint32_t lo = pl.st_off[s], hi = lo + pl.st_bytes[s];
    while (keep < kNB - 1) {
      const int t = ((s - 1 - keep) % n + n) % n;
      const uint32_t tlo = pl.st_off[t], thi = tlo + pl.st_bytes[t];
      if (tlo < hi && lo < thi) break;   // overlaps: stage t must have been consumed
      if (keep + 1 >= n) break;          // never count the stage itself
      ++keep;
    }
    pl.st_keep[s] = (uint8_t)keep;
  }
  for (int i = 0; i < pl.n_chunks; ++i) {   // ring offsets of the chunks' stages
    ChunkRec& cr = pl.ch[i];
    const int sA = cr.stage;
    cr.offA = pl.st_off[sA];
    if (cr.flags & CH_SPLIT) {
      cr.offB = pl.st_off[sA + 1];
      cr.out_off = cr.offB + (kCW / 2) * ((uint32_t)(25 - kSplitK) * 32u + 16u);
    } else {
      cr.offB = cr.offA;
      cr.out_off = cr.offA + (uint32_t)(cr.rows / 2) * ((uint32_t)cr.kin * 32u + 16u);
    }
  }
  pl.tail_off = pl.st_off[n - 1];
  return OPZ_OK;
}

}  // namespace tc2

bool tc2_supports(const opz_model* m) {
  const int L = m->nd.n_layers;
  if (m->nz != tc2::kNz || (L != 2 && L != 3)) return false;
  for (int j = 0; j < L - 1; ++j)
    if ((m->nd.sizes[j + 1] + 15) / 16 * 16 > tc2::kMaxHidden) return false;
  return true;
}

int tc2_model_prepare(opz_model* m, int precision) {
  using namespace tc2;
  const bool fp16 = precision == OPZ_PREC_FP16;
  OPZ_CUDA(cudaSetDevice(m->ctx->device));
  std::vector<float> w((size_t)3 * m->nd.P);
  OPZ_CUDA(cudaMemcpyAsync(w.data(), m->w_dev, sizeof(float) * w.size(), cudaMemcpyDeviceToHost, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
  Plan* pl = static_cast<Plan*>(m->tc2_plan);
  if (!pl) { pl = new Plan(); m->tc2_plan = pl; }
  // OPZ_TC_TAPES=1 gives the plainly rounded single tape (the 8-day error study runs both), OPZ_TC_NO_XDITHER=1 the
  // round-to-nearest state image
  int n_tapes = kTapes;
  if (const char* e = getenv("OPZ_TC_TAPES")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) n_tapes = v; }
  std::vector<uint16_t> tape;
  int rc = build_plan(m, w, fp16, 0, n_tapes, *pl, tape);
  if (rc) return rc;
  const size_t tape_al = (pl->tape_bytes + 255) / 256 * 256;
  pl->tape_stride = tape_al;
  pl->n_tapes = n_tapes;
  const size_t need = tape_al * n_tapes + 256;
  if (m->tc2_pack_bytes < need) {
    if (m->tc2_pack) { OPZ_CUDA(cudaFree(m->tc2_pack)); m->tc2_pack = nullptr; m->tc2_pack_bytes = 0; }
    if (cudaMalloc(&m->tc2_pack, need) != cudaSuccess) { (void)cudaGetLastError(); set_error("device allocation failed"); return OPZ_ENOMEM; }
    m->tc2_pack_bytes = need;
  }
  pl->tape_dev = static_cast<uint8_t*>(m->tc2_pack);
  for (int t = 0; t < n_tapes; ++t) {
    if (t > 0) {
      Plan scratch = *pl;   // same schedule; only the rounding of the entries differs
      if ((rc = build_plan(m, w, fp16, t, n_tapes, scratch, tape))) return rc;
      if (scratch.tape_bytes != pl->tape_bytes) { set_error("tensor-core path: tape versions differ in size"); return OPZ_EINVAL; }
    }
    OPZ_CUDA(cudaMemcpyAsync(pl->tape_dev + (size_t)t * tape_al, tape.data(), pl->tape_bytes, cudaMemcpyHostToDevice, m->ctx->stream));
    OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));  // the host vector is rebuilt / goes out of scope
  }
  return OPZ_OK;
}

void tc2_model_release(opz_model* m) {
  if (m->tc2_pack) cudaFree(m->tc2_pack);
  m->tc2_pack = nullptr;
  m->tc2_pack_bytes = 0;
  delete static_cast<tc2::Plan*>(m->tc2_plan);
  m->tc2_plan = nullptr;
}

int tc2_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c, int precision) {
  using namespace tc2;
  const bool fp16 = precision == OPZ_PREC_FP16;
  const Plan* pl = static_cast<const Plan*>(m->tc2_plan);
  if (!pl || pl->fp16 != fp16) { set_error("tensor-core path: model not prepared for this precision"); return OPZ_EINVAL; }
  Args a{};
  a.P = c.P;
  {
    const DevParams& P = c.P;
    const double nzf = P.nzf, c2 = (double)P.two_inv_dRi * 1.4426950408889634;
    const double pk[18] = {P.cBz * nzf, (double)P.cBz * P.eps, P.s_u * nzf, (double)P.s_u * P.eps, P.s_v * nzf, (double)P.s_v * P.eps,
                           c2, -(double)P.Ric * c2, P.c_u * nzf, P.c_v * nzf, P.c_T * nzf, P.A_u * nzf, P.A_v * nzf, P.A_T * nzf,
                           (double)P.cor_u * P.s_v, (double)P.cor_u * P.mu_v, -(double)P.cor_v * P.s_u, -(double)P.cor_v * P.mu_u};
    for (int i = 0; i < 18; ++i) a.pk[i] = (float)pk[i];
  }
  a.tape = pl->tape_dev;
  a.x0 = c.x0; a.bc = c.bc; a.out = c.out; a.B = c.B; a.scheme = c.scheme; a.dt = c.dt; a.t0 = c.t0;
  a.n_steps = c.n_steps; a.save_every = c.save_every; a.n_save = c.n_save;
  a.L = pl->L; a.n_stages = pl->n_stages;
  a.hpad[0] = pl->hpad[0]; a.hpad[1] = pl->hpad[1];
  std::memcpy(a.b3, pl->b3, sizeof(a.b3));
  a.n_mma_warps = 2;
  a.tape_mask = pl->n_tapes - 1;
  a.tape_stride = (uint32_t)pl->tape_stride;
  a.n0 = c.dt > 0.0f ? (int)lrintf(c.t0 / c.dt) : 0;   // the dither phases continue where the previous window of the rollout stopped
  a.dither_x = getenv("OPZ_TC_NO_XDITHER") == nullptr;
  a.interleave = pl->interleave ? 1 : 0;
  a.nc1 = (pl->hpad[0] + kCW - 1) / kCW;
  a.n_chunks = pl->n_chunks;
  a.h1_mult = pl->interleave ? 1 : 3;
  a.tail_off = pl->tail_off;
  std::memcpy(a.tail_ks, pl->tail_ks, sizeof(a.tail_ks));
  std::memcpy(a.tail_net, pl->tail_net, sizeof(a.tail_net));
  std::memcpy(a.ch, pl->ch, sizeof(a.ch));
  if (const char* e = getenv("OPZ_TC_MMA_WARPS")) a.n_mma_warps = atoi(e) == 1 ? 1 : 2;
  std::memcpy(a.st_off, pl->st_off, sizeof(a.st_off));
  std::memcpy(a.st_bytes, pl->st_bytes, sizeof(a.st_bytes));
  std::memcpy(a.st_src, pl->st_src, sizeof(a.st_src));
  std::memcpy(a.st_keep, pl->st_keep, sizeof(a.st_keep));
  const size_t smem = SmemLayout::bytes;
  if ((int)smem > ctx->smem_optin) { set_error("tensor-core path: shared memory budget exceeded"); return OPZ_EUNSUP; }
  const bool profile = getenv("OPZ_TC_PROFILE") != nullptr && (pl->act == OPZ_ACT_RELU || (pl->act == OPZ_ACT_MISH && fp16));
  using KernelT = void (*)(const Args);
  // epilogue warps per CTA: 8 (two threads per ocean column).  16 measured slower on wide nets (1.4e8 vs 1.6e8: the drain
  // is tensor-memory traffic, more warps only add barrier arrivals) and, since the activations are branch-free, no faster on
  // narrow ones (4.34e8 vs 4.45e8 on the 50/20 mish net); OPZ_TC_EPI_WARPS=16 keeps it selectable
  int E = 8;
  if (const char* e = getenv("OPZ_TC_EPI_WARPS")) E = atoi(e) == 16 ? 16 : 8;
#define OPZ_TC2_ROW(F, EW) {rollout_tc2_kernel<F, 0, EW, false>, rollout_tc2_kernel<F, 1, EW, false>, rollout_tc2_kernel<F, 2, EW, false>, \
                            rollout_tc2_kernel<F, 3, EW, false>, rollout_tc2_kernel<F, 4, EW, false>, rollout_tc2_kernel<F, 5, EW, false>}
  static const KernelT table[2][2][6] = {{OPZ_TC2_ROW(false, 8), OPZ_TC2_ROW(true, 8)}, {OPZ_TC2_ROW(false, 16), OPZ_TC2_ROW(true, 16)}};
#undef OPZ_TC2_ROW
  KernelT kern = table[E == 16 ? 1 : 0][fp16 ? 1 : 0][pl->act];
  if (c.P.flags & (OPZ_FLAG_SMOOTH_NN | OPZ_FLAG_SMOOTH_RI)) {   // box-filter instantiations: 8 epilogue warps only
#define OPZ_TC2_SROW(F) {rollout_tc2_kernel<F, 0, 8, false, true>, rollout_tc2_kernel<F, 1, 8, false, true>, rollout_tc2_kernel<F, 2, 8, false, true>, \
                         rollout_tc2_kernel<F, 3, 8, false, true>, rollout_tc2_kernel<F, 4, 8, false, true>, rollout_tc2_kernel<F, 5, 8, false, true>}
    static const KernelT stable[2][6] = {OPZ_TC2_SROW(false), OPZ_TC2_SROW(true)};
#undef OPZ_TC2_SROW
    kern = stable[fp16 ? 1 : 0][pl->act];
    E = 8;
  } else if (profile) {
    if (pl->act == OPZ_ACT_MISH) kern = E == 16 ? rollout_tc2_kernel<true, OPZ_ACT_MISH, 16, true> : rollout_tc2_kernel<true, OPZ_ACT_MISH, 8, true>;
    else if (E == 16) kern = fp16 ? rollout_tc2_kernel<true, OPZ_ACT_RELU, 16, true> : rollout_tc2_kernel<false, OPZ_ACT_RELU, 16, true>;
    else kern = fp16 ? rollout_tc2_kernel<true, OPZ_ACT_RELU, 8, true> : rollout_tc2_kernel<false, OPZ_ACT_RELU, 8, true>;
  }
  OPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (c.B + kM - 1) / kM;
  const int64_t n_items = (n_tiles + 1) / 2;
  int64_t grid = std::min<int64_t>(n_items, ctx->sm_count / 2) * 2;
  if (grid < 2) grid = 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((3 + E) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OPZ_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  OPZ_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (profile) {
    OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long w[16];
    OPZ_CUDA(cudaMemcpyFromSymbol(w, g_tc2_wait, sizeof(w)));
    const int64_t my_items = (n_items + grid / 2 - 1) / (grid / 2);
    const int stages = c.scheme == OPZ_RK4 ? 4 : (c.scheme == OPZ_RK2 ? 2 : 1);
    const double per = 1.0 / ((double)my_items * c.n_steps * stages);
    fprintf(stderr, "opz tc2 profile (CTA 0, cycles per tile right-hand side): MMA warp total %.0f | wait X %.0f, D-empty %.0f, H1 %.0f, "
            "weights %.0f, H2-full %.0f; decode %.0f, issue %.0f || epilogue warp total %.0f | wait D-full %.0f, H2-empty %.0f, Y-full %.0f, physics A %.0f, physics B + publish %.0f, drain ld %.0f, drain st+arrive %.0f\n",
            w[W_TOTAL] * per, w[W_X] * per, w[W_DEMPTY] * per, w[W_H1] * per, w[W_FULL] * per, w[W_H2FULL] * per, w[W_DECODE] * per, w[W_ISSUE] * per,
            w[E_TOTAL] * per, w[E_DFULL] * per, w[E_H2EMPTY] * per, w[E_YFULL] * per, w[E_PHYS_A] * per, w[E_PHYS_B] * per, w[E_LD] * per, w[E_ST] * per);
  }
  return OPZ_OK;
}

}  // namespace opz
