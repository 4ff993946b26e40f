// opz_tc.cu — tcgen05 / TMEM tensor-core rollout kernel (OPZ_PREC_BF16).
//
// One CTA owns a tile of 128 ocean columns for the whole rollout; tile row r == TMEM lane r ==
// the thread that runs that column's physics.  Per right-hand side the three Flux Chains
// (wind_mixing/construct_NN.jl:29-34) run as a statically scheduled stream of tcgen05.mma
// instructions (BF16 operands, FP32 accumulation in tensor memory):
//
//   TMEM columns  [  0,128)  D0, D1   two 64-wide FP32 accumulator chunks (double buffered)
//                 [128,224)  Y        3 x 32 FP32 outputs of the last Dense layers (uw | vw | wT)
//                 [224,272)  X        the scaled state [u;v;T] as 96 BF16 (A operand of layer 1)
//                 [272,304)  H2c      one 64-wide BF16 chunk of the last hidden layer
//                 [304,504)  HF       a full hidden activation, up to 400 BF16 (A of layer 2)
//
//   * every A operand comes from TENSOR MEMORY (tcgen05.mma with A in TMEM), so shared-memory
//     bandwidth carries only the weights;
//   * the weights (1.27 MB BF16 for construct_NN, too large for shared memory) are pre-packed once
//     into a "tape" of no-swizzle K-major blocks in exact consumption order and streamed from L2
//     through a 6-slot ring of bulk async copies (cp.async.bulk + mbarrier complete_tx);
//   * the last hidden layer is fused chunk-by-chunk into the output layer (its 64-wide chunks are
//     the K-blocks of the 31-wide output GEMM), so it is never stored in full;
//   * four epilogue warps drain the accumulators (tcgen05.ld -> bias -> activation -> BF16 ->
//     tcgen05.st as the next A operand) while the next chunk's MMAs run, then evaluate the physics
//     (opz_device.cuh::column_rhs, FP32) and the Runge-Kutta update with the FP32 state in shared
//     memory, and hand the new BF16 stage state back to the MMA warp.
//
// Warp roles: warp 0 = weight-tape producer, warp 1 = TMEM allocator + MMA issuer (one thread),
// warps 2..5 = epilogue / physics (warp w owns TMEM lanes 32*(w%4) .. +31).
//
// Replaces the batched form of NDE! / predict_flux! (wind_mixing/src/training_postprocessing.jl:105-153)
// inside a fixed-step solve (solve_NDE_mutating, :55-159).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "opz_internal.h"
#include "opz_ptx.cuh"

namespace opz {
namespace tc {

constexpr int kNz = 32;
constexpr int kS = 3 * kNz;          // 96 state entries per column
constexpr int kM = 128;              // columns per tile
constexpr int kLd = kM + 1;          // padded leading dimension of the FP32 state arrays
constexpr int kCW = 64;              // accumulator chunk width (TMEM columns)
constexpr int kSlotBytes = 12288;    // 64 rows x 96 k x 2 B
constexpr int kSlots = 6;
constexpr int kMaxOps = 160;
constexpr int kMaxKsteps = 6;         // k-steps per weight stage and CTA of the group (the issue sequence is unrolled this far)
constexpr int kMaxHidden = 400;      // HF region: 200 TMEM columns
constexpr int kThreads = 192;

constexpr uint32_t kColD0 = 0, kColY = 128, kColX = 224, kColH2 = 272, kColHF = 304;

// mbarrier indices
enum : int {
  B_FULL = 0,                   // [kSlots] tape slot filled            (producer expect_tx)
  B_EMPTY = B_FULL + kSlots,    // [kSlots] tape slot consumed          (tcgen05.commit)
  B_DFULL = B_EMPTY + kSlots,   // [2] accumulator chunk complete       (tcgen05.commit)
  B_DEMPTY = B_DFULL + 2,       // [2] accumulator chunk drained        (4 epilogue warps)
  B_H1 = B_DEMPTY + 2,          // [7] chunk c of HF written            (4 epilogue warps)
  B_H2FULL = B_H1 + 7,          // H2c written                          (4 epilogue warps)
  B_H2EMPTY = B_H2FULL + 1,     // H2c consumed by the output-layer MMA (tcgen05.commit)
  B_YFULL = B_H2EMPTY + 1,      // Y complete                           (tcgen05.commit)
  B_XFULL = B_YFULL + 1,        // X written, Y consumed                (4 epilogue warps)
  B_COUNT = B_XFULL + 1
};
static_assert(B_COUNT <= 32, "phase bits are tracked in one 32-bit mask");

// op flags
enum : uint16_t {
  F_WAIT_X = 1 << 0,      // first op of a right-hand side
  F_NEW_CHUNK = 1 << 1,   // starts a new accumulator chunk: take the next D buffer, wait until drained
  F_H1_RESET = 1 << 2,    // first op that reads HF for this net
  F_WAIT_H2 = 1 << 3,     // output-layer op: A = H2c
  F_ACC = 1 << 4,         // accumulate on the first k-step
  F_D_IS_Y = 1 << 5,      // D = Y + d_col instead of the current chunk buffer
  F_COMMIT_D = 1 << 6,    // last op of a chunk
  F_COMMIT_H2E = 1 << 7,
  F_COMMIT_Y = 1 << 8
};

struct Op {            // one weight-tape stage and the MMAs that consume it (16 bytes)
  uint16_t a_col;      // TMEM column of the A operand's first k-step
  uint16_t d_col;      // with F_D_IS_Y: TMEM column of D
  uint8_t n;           // MMA N (rows of the weight block), multiple of 16
  uint8_t ksteps;      // k-steps of 16
  uint8_t h1_need;     // HF chunks [0, h1_need) must have been written
  uint8_t pad;
  uint16_t flags;
  uint16_t pad2;
  uint32_t bytes;      // n * ksteps * 32
};
static_assert(sizeof(Op) == 16, "Op must stay 16 bytes");

struct Plan {                     // host-side, cached on the model
  int n_ops = 0;
  Op ops[kMaxOps];
  int L = 0;                      // Dense layers (2 or 3)
  int hpad[2] = {0, 0};           // padded hidden widths
  int ks_full[2] = {1, 1}, ks_tail[2] = {1, 1};  // k-steps per stage (64-row chunks / tail chunk)
  int act = 0;
  bool fp16 = false;              // operand format of the tape (IEEE half instead of bfloat16)
  int ncta = 1;                   // 2: every weight block is stored as two row halves, one per CTA of a pair
  int bias_per_net = 0;           // floats
  size_t tape_bytes = 0;          // per right-hand side
  size_t tape_stride = 0;         // bytes between replicas
  int n_copies = 1;               // replicas of the tape in global memory (spreads the L2 slices the CTAs hit)
  uint8_t* tape_dev = nullptr;    // [tape_bytes]
  float* bias_dev = nullptr;      // [3][bias_per_net]
};

struct Args {
  DevParams P;
  const uint8_t* tape;
  const float* bias;
  const float* x0;
  const float* bc;
  float* out;
  int64_t B;
  int scheme;
  float dt, t0;
  int n_steps, save_every, n_save;
  int n_ops, L, act, bias_per_net;
  int hpad[2];
  int ks_full[2], ks_tail[2];   // k-steps per weight stage for 64-row / tail chunks of hidden layer j
  uint32_t tape_bytes;
  uint32_t tape_stride;
  int n_copies;
  int profile;
  Op ops[kMaxOps];
};

struct SmemLayout {
  // s_xn is padded (leading dimension kLd) so that the transposing snapshot copy is conflict-free;
  // s_acc and s_xs are only ever touched by their own column's thread (leading dimension kM)
  static constexpr size_t off_xn = 0;
  static constexpr size_t off_acc = off_xn + (size_t)kS * kLd * 4;
  static constexpr size_t off_xs = off_acc + (size_t)kS * kM * 4;
  static constexpr size_t off_ring = (off_xs + (size_t)kS * kM * 4 + 127) / 128 * 128;
  static constexpr size_t off_bars = off_ring + (size_t)kSlots * kSlotBytes;
  static constexpr size_t off_tmem = off_bars + 8 * B_COUNT;
  static constexpr size_t off_bias = off_tmem + 16;
  static size_t bytes(int bias_per_net) { return off_bias + sizeof(float) * 3 * (size_t)bias_per_net; }
};

__device__ __forceinline__ void bar_wait(uint64_t* bars, int id, uint32_t& mask) {
  ptx::mbar_wait(bars + id, (mask >> id) & 1u);
  mask ^= 1u << id;
}
// Optional wait-time profile (OPZ_TC_PROFILE=1): cycles CTA 0's MMA warp and first epilogue warp spend
// waiting on each class of barrier.  [0..7] MMA warp, [8..15] epilogue warp.
__device__ unsigned long long g_tc_wait[16];
enum { W_X = 0, W_DEMPTY = 1, W_H1 = 2, W_FULL = 3, W_H2FULL = 4, W_TOTAL = 5,
       E_DFULL = 8, E_H2EMPTY = 9, E_YFULL = 10, E_PHYS = 11, E_TOTAL = 12, E_PUBLISH = 13 };
__device__ __forceinline__ void bar_wait_t(uint64_t* bars, int id, uint32_t& mask, bool prof, long long* acc, int slot) {
  if (!prof) { bar_wait(bars, id, mask); return; }
  const long long t0 = clock64();
  bar_wait(bars, id, mask);
  acc[slot] += clock64() - t0;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// two FP32 values -> one 32-bit TMEM column of the A operand (low half = even k), activation applied
template <bool FP16, int ACT>
__device__ __forceinline__ uint32_t pack_act(float a, float b) {
  if constexpr (ACT == OPZ_ACT_RELU) return FP16 ? ptx::pack_f16x2_relu(a, b) : ptx::pack_bf16x2_relu(a, b);
  a = act_fast<ACT>(a);
  b = act_fast<ACT>(b);
  return FP16 ? ptx::pack_f16x2(a, b) : ptx::pack_bf16x2(a, b);
}

// NCTA = 2: CTA pair (cta_group::2).  The two CTAs of a cluster own two adjacent column tiles; every tcgen05.mma is
// issued once, by the leader, with M = 256: rows 0..127 are the leader's tile (A and D in its tensor memory), rows
// 128..255 the peer's, and each CTA streams only HALF of every weight block into its own ring.  That halves the
// MMA-issue work and the L2 -> SM weight traffic per SM.  Barriers the issuer waits on live in the leader CTA and are
// arrived on remotely by the peer's warps; completions are multicast to both CTAs by tcgen05.commit.
template <bool FP16, int ACT, int NCTA>
__global__ void __launch_bounds__(kThreads, 1) rollout_tc_kernel(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_xn = reinterpret_cast<float*>(smem + SmemLayout::off_xn);
  float* s_acc = reinterpret_cast<float*>(smem + SmemLayout::off_acc);
  float* s_xs = reinterpret_cast<float*>(smem + SmemLayout::off_xs);
  uint8_t* s_ring = smem + SmemLayout::off_ring;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::off_bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SmemLayout::off_tmem);
  float* s_bias = reinterpret_cast<float*>(smem + SmemLayout::off_bias);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t n_tiles = (a.B + kM - 1) / kM;
  const int stages = a.scheme == OPZ_RK4 ? 4 : (a.scheme == OPZ_RK2 ? 2 : 1);
  const int n_rhs = a.n_steps * stages;
  const uint32_t rank = NCTA == 2 ? ptx::cluster_ctarank() : 0u;
  // work items: one tile (NCTA = 1) or one pair of adjacent tiles (NCTA = 2) per cluster and iteration
  const int64_t n_items = (n_tiles + NCTA - 1) / NCTA;
  const int64_t item0 = blockIdx.x / NCTA, item_stride = gridDim.x / NCTA;
  // arrive on a barrier the MMA issuer waits on: it lives in the leader CTA
  const uint32_t leader_bars = NCTA == 2 ? ptx::mapa(ptx::smem_u32(bars), 0u) : 0u;
  auto arrive_issuer = [&](int id) {
    if (NCTA == 2 && rank != 0) ptx::mbar_arrive_remote(leader_bars + 8u * (uint32_t)id);
    else ptx::mbar_arrive(bars + id);
  };

  for (int i = threadIdx.x; i < 3 * a.bias_per_net; i += kThreads) s_bias[i] = a.bias[i];
  if (warp == 1) {
    if (lane == 0) {
      // the leader's "full" barriers also take the peer's forwarded arrival; the peer's own count only its producer
      for (int s = 0; s < kSlots; ++s) { ptx::mbar_init(bars + B_FULL + s, rank == 0 ? NCTA : 1); ptx::mbar_init(bars + B_EMPTY + s, 1); }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(bars + B_DFULL + b, 1); ptx::mbar_init(bars + B_DEMPTY + b, 4 * NCTA); }
      for (int c = 0; c < 7; ++c) ptx::mbar_init(bars + B_H1 + c, 4 * NCTA);
      ptx::mbar_init(bars + B_H2FULL, 4 * NCTA);
      ptx::mbar_init(bars + B_H2EMPTY, 1);
      ptx::mbar_init(bars + B_YFULL, 1);
      ptx::mbar_init(bars + B_XFULL, 4 * NCTA);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    if (NCTA == 2) ptx::tmem_alloc2<512>(tmem_slot);
    else ptx::tmem_alloc<512>(tmem_slot);
  }
  ptx::tc_fence_before();
  if (NCTA == 2) ptx::cluster_sync();  // both CTAs' barriers exist before any remote arrive / multicast commit
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= weight-tape producer =================
    if (lane == 0 && n_rhs > 0) {
      uint32_t mask = 0xFFFFFFFFu;  // "empty" barriers start in the released state
      int slot = 0;
      const uint8_t* tape0 = a.tape + (size_t)((blockIdx.x / NCTA) % a.n_copies) * a.tape_stride;
      for (int64_t item = item0; item < n_items; item += item_stride) {
        for (int r = 0; r < n_rhs; ++r) {
          const uint8_t* src = tape0;
          for (int i = 0; i < a.n_ops; ++i) {
            const uint32_t bytes = a.ops[i].bytes / NCTA;   // this CTA's half of the weight block
            bar_wait(bars, B_EMPTY + slot, mask);
            ptx::mbar_arrive_expect_tx(bars + B_FULL + slot, bytes);
            ptx::bulk_g2s(s_ring + (size_t)slot * kSlotBytes, src + rank * bytes, bytes, bars + B_FULL + slot);
            src += a.ops[i].bytes;
            slot = slot + 1 == kSlots ? 0 : slot + 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && NCTA == 2 && rank != 0) {
    // ================= peer CTA: forward "my half of the weight block has landed" to the leader =================
    if (lane == 0 && n_rhs > 0) {
      uint32_t mask = 0;
      int slot = 0;
      for (int64_t item = item0; item < n_items; item += item_stride)
        for (int r = 0; r < n_rhs; ++r)
          for (int i = 0; i < a.n_ops; ++i) {
            bar_wait(bars, B_FULL + slot, mask);
            ptx::mbar_arrive_remote(leader_bars + 8u * (uint32_t)(B_FULL + slot));
            slot = slot + 1 == kSlots ? 0 : slot + 1;
          }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp walks the schedule with warp-uniform loop counters (the same nested loops that
    // build_plan() used to lay out the tape), so descriptors and addresses live in uniform registers;
    // one elected lane issues the tcgen05.mma / tcgen05.commit instructions.
    if (n_rhs > 0) {
      // released state for the barriers this warp waits on as "empty": D-drained
      uint32_t mask = (1u << (B_DEMPTY + 0)) | (1u << (B_DEMPTY + 1));
      uint32_t slot = 0, g = 0;
      const bool prof = a.profile && blockIdx.x == 0;
      long long wacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_begin = clock64();
      const uint32_t ring_u32 = ptx::smem_u32(s_ring);
      const uint32_t idesc0 = ptx::idesc_f16kind(kM * NCTA, 0, FP16);
      const int L = a.L;
      // one weight stage: wait for its bytes, issue ks MMAs D (+)= A[k] * B[k]^T, release the slot
      // The MMAs are issued from a fully unrolled, predicated sequence: descriptor and address offsets are
      // immediates, so consecutive tcgen05.mma leave back to back (a rolled loop with 64-bit descriptor
      // arithmetic between the MMAs measured 1.6x slower end to end).
      auto stage = [&](uint32_t d, uint32_t a_col, uint32_t rows, int ks, bool acc_first) {
        if (prof) {
          const long long tw0 = clock64();
          bar_wait(bars, B_FULL + slot, mask);
          wacc[W_FULL] += clock64() - tw0;
        } else {
          bar_wait(bars, B_FULL + slot, mask);
        }
        ptx::tc_fence_after();  // orders the MMAs after the epilogue warps' tcgen05.st seen through the barriers waited on above
        // each CTA holds rows / NCTA rows of the block: [k/8][rows / NCTA][8]
        const uint64_t bd0 = ptx::smem_desc(ring_u32 + slot * kSlotBytes, (rows / NCTA) * 16u, 128);
        const uint32_t idesc = idesc0 | ((rows >> 3) << 17);
        const uint32_t kstride = 2u * rows / NCTA;  // descriptor units (16 B) per k-step
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < kMaxKsteps * NCTA; ++k)
            if (k < ks) {
              if (NCTA == 2) ptx::mma_ts2(tmem + d, tmem + a_col + (uint32_t)k * 8u, bd0 + (uint64_t)((uint32_t)k * kstride), idesc, acc_first || k > 0);
              else ptx::mma_ts(tmem + d, tmem + a_col + (uint32_t)k * 8u, bd0 + (uint64_t)((uint32_t)k * kstride), idesc, acc_first || k > 0);
            }
          if (NCTA == 2) ptx::tc_commit2(bars + B_EMPTY + slot);
          else ptx::tc_commit(bars + B_EMPTY + slot);
        }
        __syncwarp();
        slot = slot + 1 == kSlots ? 0 : slot + 1;
      };
      for (int64_t item = item0; item < n_items; item += item_stride) {
#pragma unroll 1
        for (int r = 0; r < n_rhs; ++r) {
          bool have_pend = false;
          uint32_t pend_net = 0, pend_c = 0, pend_rows = 0;
          // output-layer MMAs for one H2c chunk: Y_net (+)= H2c * W_out[:, chunk]^T
          auto out_op = [&](bool final_op) {
            bar_wait_t(bars, B_H2FULL, mask, prof, wacc, W_H2FULL);
            stage(kColY + 32u * pend_net, kColH2, 32u, (int)(pend_rows >> 4), pend_c > 0);
            if (ptx::elect_one()) {
              if (NCTA == 2) { ptx::tc_commit2(bars + B_H2EMPTY); if (final_op) ptx::tc_commit2(bars + B_YFULL); }
              else { ptx::tc_commit(bars + B_H2EMPTY); if (final_op) ptx::tc_commit(bars + B_YFULL); }
            }
            __syncwarp();
          };
          bar_wait_t(bars, B_XFULL, mask, prof, wacc, W_X);
#pragma unroll 1
          for (int net = 0; net < 3; ++net) {
#pragma unroll 1
            for (int j = 0; j < L - 1; ++j) {
              const bool fused = (j == L - 2);
              const int kin = (j == 0) ? kS / 16 : a.hpad[j - 1] / 16;
              const int hp = a.hpad[j];
              const uint32_t a_base = (j == 0) ? kColX : kColHF;
              int h1_waited = 0;
#pragma unroll 1
              for (int c0 = 0; c0 < hp; c0 += kCW) {
                const uint32_t rows = (uint32_t)min(kCW, hp - c0);
                const int ks_per = (rows == (uint32_t)kCW) ? a.ks_full[j] : a.ks_tail[j];
                const uint32_t buf = g & 1u;
                ++g;
                bar_wait_t(bars, B_DEMPTY + buf, mask, prof, wacc, W_DEMPTY);
#pragma unroll 1
                for (int k0 = 0; k0 < kin; k0 += ks_per) {
                  const int ks = min(ks_per, kin - k0);
                  if (j > 0) {  // HF chunks covering inputs [0, 16 (k0 + ks)) must have been written
                    const int need = ((k0 + ks) * 16 + kCW - 1) / kCW;
                    while (h1_waited < need) { bar_wait_t(bars, B_H1 + h1_waited, mask, prof, wacc, W_H1); ++h1_waited; }
                  }
                  stage(kColD0 + buf * kCW, a_base + (uint32_t)k0 * 8u, rows, ks, k0 > 0);
                }
                if (ptx::elect_one()) {
                  if (NCTA == 2) ptx::tc_commit2(bars + B_DFULL + buf);
                  else ptx::tc_commit(bars + B_DFULL + buf);
                }
                __syncwarp();
                // the output-layer MMAs of the PREVIOUS fused chunk go after this chunk's MMAs
                if (have_pend) { out_op(false); have_pend = false; }
                if (fused) { have_pend = true; pend_net = (uint32_t)net; pend_c = (uint32_t)(c0 / kCW); pend_rows = rows; }
              }
            }
          }
          out_op(true);
        }
      }
      if (prof && lane == 0) {
        wacc[W_TOTAL] = clock64() - t_begin;
        for (int i = 0; i < 8; ++i) g_tc_wait[i] = (unsigned long long)wacc[i];
      }
    }
  } else {
    // ================= epilogue + physics: thread == column == TMEM lane =================
    const int q = warp % 4;
    const int m = q * 32 + lane;
    const int etid = (warp - 2) * 32 + lane;  // 0..127 for cooperative copies
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t mask = 1u << B_H2EMPTY;  // released state
    uint32_t g = 0;
    const float h = a.dt;
    const bool prof = a.profile && blockIdx.x == 0 && warp == 2;
    long long eacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();

    for (int64_t item = item0; item < n_items; item += item_stride) {
      const int64_t tile = item * NCTA + rank;   // may be one past the last tile for the peer: a padding tile, never stored
      const int64_t col0 = tile * kM;
      const int64_t col = col0 + m;
      // ---- load the initial profiles (coalesced through the padded state array) ----
      epi_bar_sync();  // previous tile's snapshot readers are done with s_xn
      for (int e = etid; e < kM * kS; e += 128) {
        const int mm = e / kS, i = e - mm * kS;
        const int64_t c = col0 + mm;
        s_xn[i * kLd + mm] = c < a.B ? a.x0[c * kS + i] : 0.0f;
      }
      float bc[OPZ_BC_STRIDE];
#pragma unroll
      for (int qq = 0; qq < OPZ_BC_STRIDE; ++qq) bc[qq] = col < a.B ? a.bc[col * OPZ_BC_STRIDE + qq] : 0.0f;
      epi_bar_sync();
      auto save = [&](int slot_idx) {
        epi_bar_sync();  // every column's s_xn is final
        for (int e = etid; e < kM * kS; e += 128) {
          const int mm = e / kS, i = e - mm * kS;
          const int64_t c = col0 + mm;
          if (c < a.B) a.out[(c * a.n_save + slot_idx) * kS + i] = s_xn[i * kLd + mm];
        }
      };
      if (a.save_every > 0) save(0);
      // stage state and its BF16 image in tensor memory
      auto publish_x = [&]() {
#pragma unroll
        for (int i0 = 0; i0 < kS; i0 += 16) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pk[j] = pack_act<FP16, OPZ_ACT_IDENTITY>(s_xs[(i0 + 2 * j) * kM + m], s_xs[(i0 + 2 * j + 1) * kM + m]);
          ptx::tmem_st8(tmem + lane_addr + kColX + i0 / 2, pk);
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_issuer(B_XFULL);
      };
#pragma unroll 4
      for (int i = 0; i < kS; ++i) s_xs[i * kM + m] = s_xn[i * kLd + m];
      if (n_rhs > 0) publish_x();

      int rhs_idx = 0;
#pragma unroll 1
      for (int n = 0; n < a.n_steps; ++n) {
        const float tn = a.t0 + (float)n * h;
#pragma unroll 1
        for (int st = 0; st < stages; ++st, ++rhs_idx) {
          // ---- drain the accumulator chunks of the hidden layers ----
          for (int net = 0; net < 3; ++net) {
            const float* bnet = s_bias + net * a.bias_per_net;
            int boff = 0;
            for (int j = 0; j < a.L - 1; ++j) {
              const bool fused = (j == a.L - 2);
              const int hp = a.hpad[j];
              for (int c = 0; c * kCW < hp; ++c) {
                const uint32_t buf = g & 1u;
                ++g;
                const int ncols = min(kCW, hp - c * kCW);
                bar_wait_t(bars, B_DFULL + buf, mask, prof, eacc, E_DFULL - 8);
                ptx::tc_fence_after();
                const uint32_t src = tmem + lane_addr + kColD0 + buf * kCW;
                const uint32_t dst = tmem + lane_addr + (fused ? kColH2 : kColHF + (uint32_t)c * (kCW / 2));
                const float* bb = bnet + boff + c * kCW;
                uint32_t rr[4][16];
#pragma unroll
                for (int grp = 0; grp < 4; ++grp)
                  if (grp * 16 < ncols) ptx::tmem_ld16(src + grp * 16, rr[grp]);
                ptx::tmem_wait_ld();
                uint32_t pk[4][8];
#pragma unroll
                for (int grp = 0; grp < 4; ++grp) {
                  if (grp * 16 < ncols) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                      const float2 b2 = *reinterpret_cast<const float2*>(bb + grp * 16 + 2 * jj);
                      pk[grp][jj] = pack_act<FP16, ACT>(__uint_as_float(rr[grp][2 * jj]) + b2.x, __uint_as_float(rr[grp][2 * jj + 1]) + b2.y);
                    }
                  }
                }
                if (fused) {  // the output-layer MMAs of the previous chunk must have consumed H2c
                  bar_wait_t(bars, B_H2EMPTY, mask, prof, eacc, E_H2EMPTY - 8);
                  ptx::tc_fence_after();
                }
#pragma unroll
                for (int grp = 0; grp < 4; ++grp)
                  if (grp * 16 < ncols) ptx::tmem_st8(dst + grp * 8, pk[grp]);
                ptx::tmem_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  arrive_issuer(B_DEMPTY + buf);
                  arrive_issuer(fused ? B_H2FULL : B_H1 + c);
                }
              }
              boff += hp;
            }
          }
          // ---- NN outputs -> physics -> Runge-Kutta update ----
          bar_wait_t(bars, B_YFULL, mask, prof, eacc, E_YFULL - 8);
          ptx::tc_fence_after();
          const long long t_phys = prof ? clock64() : 0;
          float y[3][32];
          {
            int boff = 0;
            for (int j = 0; j < a.L - 1; ++j) boff += a.hpad[j];
#pragma unroll
            for (int net = 0; net < 3; ++net) {
              uint32_t rr[32];
              ptx::tmem_ld32(tmem + lane_addr + kColY + net * 32, rr);
              ptx::tmem_wait_ld();
              const float* b3 = s_bias + net * a.bias_per_net + boff;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) y[net][jj] = __uint_as_float(rr[jj]) + b3[jj];
            }
          }
          float ts = tn;
          if (a.scheme == OPZ_RK4) ts = (st == 0) ? tn : (st == 3 ? tn + h : tn + h * 0.5f);
          else if (a.scheme == OPZ_RK2) ts = (st == 0) ? tn : tn + h;
          // Runge-Kutta bookkeeping in the oracle's operation order: acc = k1 (+ 2 k2 + 2 k3);
          // stage input xs = xn + cs k; last stage xn += cf (acc + k)
          // (selects, not branches: the unrolled physics must exist once in the instruction stream)
          const bool first = (st == 0), last = (st == stages - 1);
          const float wk = (a.scheme == OPZ_RK4 && !first && !last) ? 2.0f : 1.0f;
          const float cs = (a.scheme == OPZ_RK4 && st < 2) ? h * 0.5f : h;
          const float cf = a.scheme == OPZ_RK4 ? h / 6.0f : (a.scheme == OPZ_RK2 ? h * 0.5f : h);
          const float cx = last ? cf : cs;
          auto X = [&](int i) { return s_xs[i * kM + m]; };
          auto Y = [&](int net, int j) { return y[net][j]; };
          auto K = [&](int i, float k) {
            const int o = i * kM + m, on = i * kLd + m;
            const float xo = s_xn[on];
            const float acc_in = first ? 0.0f : s_acc[o];
            const float sum = acc_in + wk * k;       // first stage: 0 + 1*k == k exactly
            const float xv = xo + cx * (last ? sum : k);
            s_acc[o] = sum;
            s_xs[o] = xv;
            if (last) s_xn[on] = xv;
          };
          column_rhs<kNz, false, true>(a.P, bc, ts, X, Y, K, NoFlux());
          const long long t_pub = prof ? clock64() : 0;
          if (rhs_idx + 1 < n_rhs) publish_x();  // also tells the MMA warp that Y has been consumed
          if (prof) { eacc[E_PHYS - 8] += t_pub - t_phys; eacc[E_PUBLISH - 8] += clock64() - t_pub; }
        }
        if (a.save_every > 0 && (n + 1) % a.save_every == 0) save((n + 1) / a.save_every);
      }
      if (a.save_every <= 0) save(0);
    }
    if (prof && lane == 0) {
      eacc[E_TOTAL - 8] = clock64() - t_begin;
      for (int i = 0; i < 8; ++i) g_tc_wait[8 + i] = (unsigned long long)eacc[i];
    }
  }
  ptx::tc_fence_before();
  if (NCTA == 2) ptx::cluster_sync();  // the peer's tensor / shared memory stay valid until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    if (NCTA == 2) ptx::tmem_dealloc2<512>(tmem);
    else ptx::tmem_dealloc<512>(tmem);
  }
}

// ---- host: weight tape + op program ---------------------------------------------------------------
static uint16_t f16_rne(float f) {  // IEEE binary16, round to nearest even, saturating to +-65504
  uint32_t u;
  std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t absu = u & 0x7FFFFFFFu;
  if (absu > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);         // nan
  if (absu >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);        // >= 65520 (or inf): saturate
  if (absu < 0x33000001u) return (uint16_t)sign;                     // < 2^-25: rounds to zero
  const int e = (int)(absu >> 23) - 127;
  uint32_t mant = (absu & 0x7FFFFFu) | 0x800000u;
  int shift = e >= -14 ? 13 : 13 + (-14 - e);                        // subnormal halves lose more bits
  uint32_t half_m = mant >> shift;
  const uint32_t rem = mant & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_m & 1u))) ++half_m;
  uint32_t out;
  if (e >= -14) out = ((uint32_t)(e + 15) << 10) + (half_m - 0x400u);  // carry into the exponent is correct
  else out = half_m;                                                   // subnormal (may round up to 0x400)
  return (uint16_t)(sign | out);
}

static uint16_t bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);  // inf / nan
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// Appends the no-swizzle K-major block of W (Flux layout: element (k_in, n_out) at w[k*nout + n]) for
// output rows [r0, r0+rows) and inputs [k0, k0 + 16*ksteps):  [k/8][row][k%8] BF16, zero padded.
static void append_block(std::vector<uint16_t>& tape, bool fp16, const float* w, int nin, int nout, int r0, int rows,
                         int k0, int ksteps, int ncta) {
  // ncta = 2: two consecutive images of rows / 2 rows each (the leader's half first)
  const int hrows = rows / ncta;
  for (int half = 0; half < ncta; ++half) {
    const size_t base = tape.size();
    tape.resize(base + (size_t)hrows * ksteps * 16, 0);
    for (int kk = 0; kk < ksteps * 16; ++kk) {
      const int k = k0 + kk;
      if (k >= nin) continue;
      for (int r = 0; r < hrows; ++r) {
        const int n = r0 + half * hrows + r;
        if (n >= nout) continue;
        const float v = w[(size_t)k * nout + n];
        tape[base + ((size_t)(kk / 8) * hrows + r) * 8 + (kk % 8)] = fp16 ? f16_rne(v) : bf16_rne(v);
      }
    }
  }
}

static int build_plan(const opz_model* m, const std::vector<float>& w, bool fp16, int ncta, Plan& pl, std::vector<uint16_t>& tape,
                      std::vector<float>& bias) {
  const NetDesc& nd = m->nd;
  const int L = nd.n_layers;
  if (m->nz != kNz) { set_error("tensor-core path: nz must be 32"); return OPZ_EUNSUP; }
  if (L != 2 && L != 3) { set_error("tensor-core path: 2 or 3 Dense layers per net"); return OPZ_EUNSUP; }
  pl.L = L;
  pl.act = nd.act;
  pl.fp16 = fp16;
  pl.ncta = ncta;
  int bias_per_net = 0;
  for (int j = 0; j < L - 1; ++j) {
    pl.hpad[j] = (nd.sizes[j + 1] + 15) / 16 * 16;
    if (pl.hpad[j] > kMaxHidden) { set_error("tensor-core path: hidden width above 400"); return OPZ_EUNSUP; }
    bias_per_net += pl.hpad[j];
  }
  bias_per_net += 32;
  pl.bias_per_net = bias_per_net;
  bias.assign((size_t)3 * bias_per_net, 0.0f);
  tape.clear();
  pl.n_ops = 0;
  bool overflow = false;
  auto push = [&](Op op) {
    if (pl.n_ops >= kMaxOps) { overflow = true; return; }
    op.bytes = (uint32_t)op.n * op.ksteps * 32u;
    pl.ops[pl.n_ops++] = op;
  };
  struct Pending { bool valid; int net, c, cols; } pend{false, 0, 0, 0};
  const int hp_last = pl.hpad[L - 2];
  const int n_last_chunks = (hp_last + kCW - 1) / kCW;
  auto emit_out = [&](const Pending& p, bool final_op) {  // output-layer MMAs for one H2c chunk
    const float* wn = w.data() + (size_t)p.net * nd.P;
    Op op{};
    op.a_col = (uint16_t)kColH2;
    op.d_col = (uint16_t)(kColY + 32 * p.net);
    op.n = 32;
    op.ksteps = (uint8_t)(p.cols / 16);  // <= 4
    op.flags = F_WAIT_H2 | F_D_IS_Y | F_COMMIT_H2E | (p.c > 0 ? F_ACC : 0) | (final_op ? F_COMMIT_Y : 0);
    append_block(tape, fp16, wn + nd.w_off[L - 1], nd.sizes[L - 1], nd.sizes[L], 0, 32, p.c * kCW, op.ksteps, ncta);
    push(op);
  };
  for (int net = 0; net < 3; ++net) {
    const float* wn = w.data() + (size_t)net * nd.P;
    int boff = 0;
    for (int j = 0; j < L - 1; ++j) {
      for (int i = 0; i < nd.sizes[j + 1]; ++i) bias[(size_t)net * bias_per_net + boff + i] = wn[nd.b_off[j] + i];
      boff += pl.hpad[j];
    }
    for (int i = 0; i < nd.sizes[L]; ++i) bias[(size_t)net * bias_per_net + boff + i] = wn[nd.b_off[L - 1] + i];
    for (int j = 0; j < L - 1; ++j) {
      const bool fused = (j == L - 2);
      const int kin = (j == 0) ? kS : pl.hpad[j - 1];
      const int ksteps_total = kin / 16;
      const int hp = pl.hpad[j];
      for (int c = 0; c * kCW < hp; ++c) {
        const int rows = std::min(kCW, hp - c * kCW);
        const int max_ks = std::min(kMaxKsteps * ncta, kSlotBytes / ((rows / ncta) * 32));
        const int n_stage = (ksteps_total + max_ks - 1) / max_ks;
        const int ks_per = (ksteps_total + n_stage - 1) / n_stage;
        if (net == 0) (rows == kCW ? pl.ks_full[j] : pl.ks_tail[j]) = ks_per;
        for (int s = 0, k = 0; k < ksteps_total; ++s, k += ks_per) {
          const int ks = std::min(ks_per, ksteps_total - k);
          Op op{};
          op.a_col = (uint16_t)((j == 0 ? kColX : kColHF) + k * 8);
          op.n = (uint8_t)rows;
          op.ksteps = (uint8_t)ks;
          op.flags = (k > 0 ? F_ACC : 0);
          if (k == 0) op.flags |= F_NEW_CHUNK;
          if (k + ks >= ksteps_total) op.flags |= F_COMMIT_D;
          if (net == 0 && j == 0 && c == 0 && k == 0) op.flags |= F_WAIT_X;
          if (j > 0) {
            if (c == 0 && k == 0) op.flags |= F_H1_RESET;
            op.h1_need = (uint8_t)(((k + ks) * 16 + kCW - 1) / kCW);  // HF chunks covering inputs [0, 16(k+ks))
          }
          append_block(tape, fp16, wn + nd.w_off[j], nd.sizes[j], nd.sizes[j + 1], c * kCW, rows, k * 16, ks, ncta);
          push(op);
        }
        // the output-layer MMAs of the PREVIOUS fused chunk go after this chunk's MMAs, so the tensor
        // pipe has work while the epilogue warps convert that chunk
        if (pend.valid) { emit_out(pend, false); pend.valid = false; }
        if (fused) pend = Pending{true, net, c, rows};
      }
    }
  }
  if (pend.valid) emit_out(pend, true);
  (void)n_last_chunks;
  if (overflow) { set_error("tensor-core path: schedule longer than the op table"); return OPZ_EUNSUP; }
  pl.tape_bytes = tape.size() * 2;
  return OPZ_OK;
}

}  // namespace tc

// OPZ_TC_V1=1 keeps the first-generation kernel of this file; the default is opz_tc2.cu
static bool use_v2() {
  const char* e = getenv("OPZ_TC_V1");
  return !(e && atoi(e) != 0);
}

int tc_model_prepare(opz_model* m, int precision) {
  using namespace tc;
  if (use_v2()) {
    const int rc2 = tc2_model_prepare(m, precision);
    if (rc2) return rc2;
    m->tc_pack_valid = true;
    return OPZ_OK;
  }
  const bool fp16 = precision == OPZ_PREC_FP16;
  OPZ_CUDA(cudaSetDevice(m->ctx->device));
  std::vector<float> w((size_t)3 * m->nd.P);
  OPZ_CUDA(cudaMemcpyAsync(w.data(), m->w_dev, sizeof(float) * w.size(), cudaMemcpyDeviceToHost, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
  Plan* pl = static_cast<Plan*>(m->tc_plan);
  if (!pl) { pl = new Plan(); m->tc_plan = pl; }
  std::vector<uint16_t> tape;
  std::vector<float> bias;
  int ncta = 1;  // OPZ_TC_PAIR=1 selects the CTA-pair (cta_group::2) variant
  if (const char* e = getenv("OPZ_TC_PAIR")) ncta = atoi(e) ? 2 : 1;
  int rc = build_plan(m, w, fp16, ncta, *pl, tape, bias);
  if (rc) return rc;
  int copies = 1;
  if (const char* e = getenv("OPZ_TC_TAPE_COPIES")) copies = atoi(e);
  if (copies < 1) copies = 1;
  if (copies > 148) copies = 148;
  pl->n_copies = copies;
  pl->tape_stride = (pl->tape_bytes + 4095) / 4096 * 4096 + 4096 * 3;  // odd multiple of 4 KiB apart
  const size_t need = pl->tape_stride * copies + bias.size() * 4 + 256;
  if (m->tc_pack_bytes < need) {
    if (m->tc_pack) { OPZ_CUDA(cudaFree(m->tc_pack)); m->tc_pack = nullptr; m->tc_pack_bytes = 0; }
    if (cudaMalloc(&m->tc_pack, need) != cudaSuccess) { (void)cudaGetLastError(); set_error("device allocation failed"); return OPZ_ENOMEM; }
    m->tc_pack_bytes = need;
  }
  pl->tape_dev = static_cast<uint8_t*>(m->tc_pack);
  pl->bias_dev = reinterpret_cast<float*>(pl->tape_dev + pl->tape_stride * copies);
  for (int c = 0; c < copies; ++c)
    OPZ_CUDA(cudaMemcpyAsync(pl->tape_dev + pl->tape_stride * c, tape.data(), pl->tape_bytes, cudaMemcpyHostToDevice, m->ctx->stream));
  OPZ_CUDA(cudaMemcpyAsync(pl->bias_dev, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));  // the host vectors go out of scope
  m->tc_pack_valid = true;
  return OPZ_OK;
}

void tc_model_release(opz_model* m) {
  tc2_model_release(m);
  if (m->tc_pack) cudaFree(m->tc_pack);
  m->tc_pack = nullptr;
  m->tc_pack_bytes = 0;
  delete static_cast<tc::Plan*>(m->tc_plan);
  m->tc_plan = nullptr;
  m->tc_pack_valid = false;
}

int tc_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c, int precision) {
  using namespace tc;
  if (precision != OPZ_PREC_BF16 && precision != OPZ_PREC_FP16) {
    set_error("tensor-core path: OPZ_PREC_BF16 and OPZ_PREC_FP16 are built; the split-operand modes are not");
    return OPZ_EUNSUP;
  }
  const bool fp16 = precision == OPZ_PREC_FP16;
  if (c.P.flags & (OPZ_FLAG_SMOOTH_NN | OPZ_FLAG_SMOOTH_RI)) {
    set_error("tensor-core path: smooth_NN / smooth_Ri are only available with OPZ_PREC_FP32");
    return OPZ_EUNSUP;
  }
  if (use_v2()) {
    if (!m->tc_pack_valid) { set_error("tensor-core path: model not prepared"); return OPZ_EINVAL; }
    return tc2_rollout(ctx, m, c, precision);
  }
  const Plan* pl = static_cast<const Plan*>(m->tc_plan);
  if (!pl || !m->tc_pack_valid || pl->fp16 != fp16) { set_error("tensor-core path: model not prepared for this precision"); return OPZ_EINVAL; }
  Args a{};
  a.P = c.P;
  a.tape = pl->tape_dev;
  a.bias = pl->bias_dev;
  a.x0 = c.x0; a.bc = c.bc; a.out = c.out; a.B = c.B; a.scheme = c.scheme; a.dt = c.dt; a.t0 = c.t0;
  a.n_steps = c.n_steps; a.save_every = c.save_every; a.n_save = c.n_save;
  a.n_ops = pl->n_ops; a.L = pl->L; a.act = pl->act; a.bias_per_net = pl->bias_per_net;
  a.hpad[0] = pl->hpad[0]; a.hpad[1] = pl->hpad[1];
  for (int j = 0; j < 2; ++j) { a.ks_full[j] = pl->ks_full[j]; a.ks_tail[j] = pl->ks_tail[j]; }
  a.tape_bytes = (uint32_t)pl->tape_bytes;
  a.tape_stride = (uint32_t)pl->tape_stride;
  a.n_copies = pl->n_copies;
  a.profile = getenv("OPZ_TC_PROFILE") != nullptr;
  std::memcpy(a.ops, pl->ops, sizeof(Op) * pl->n_ops);
  const size_t smem = SmemLayout::bytes(pl->bias_per_net);
  if ((int)smem > ctx->smem_optin) { set_error("tensor-core path: shared memory budget exceeded"); return OPZ_EUNSUP; }
  using KernelT = void (*)(const Args);
#define OPZ_TC_ROW(F, N) {rollout_tc_kernel<F, 0, N>, rollout_tc_kernel<F, 1, N>, rollout_tc_kernel<F, 2, N>, \
                          rollout_tc_kernel<F, 3, N>, rollout_tc_kernel<F, 4, N>, rollout_tc_kernel<F, 5, N>}
  static const KernelT table[2][2][6] = {{OPZ_TC_ROW(false, 1), OPZ_TC_ROW(true, 1)}, {OPZ_TC_ROW(false, 2), OPZ_TC_ROW(true, 2)}};
#undef OPZ_TC_ROW
  const int ncta = pl->ncta;
  const KernelT kern = table[ncta - 1][fp16 ? 1 : 0][pl->act];
  OPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (c.B + kM - 1) / kM;
  const int64_t n_items = (n_tiles + ncta - 1) / ncta;
  int64_t grid = std::min<int64_t>(n_items, ctx->sm_count / ncta) * ncta;
  if (grid < ncta) grid = ncta;
  if (ncta == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
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
  } else {
    kern<<<(unsigned)grid, kThreads, smem, ctx->stream>>>(a);
  }
  OPZ_CUDA(cudaGetLastError());
  ctx->launches += 1;
  if (a.profile) {
    OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long w[16];
    OPZ_CUDA(cudaMemcpyFromSymbol(w, g_tc_wait, sizeof(w)));
    const int64_t my_tiles = (n_items + grid / ncta - 1) / (grid / ncta);
    const int stages = c.scheme == OPZ_RK4 ? 4 : (c.scheme == OPZ_RK2 ? 2 : 1);
    const double per = 1.0 / ((double)my_tiles * c.n_steps * stages);
    fprintf(stderr, "opz tc profile (CTA 0, cycles per tile right-hand side): MMA warp total %.0f | wait X %.0f, D-empty %.0f, H1 %.0f, "
            "weights %.0f, H2-full %.0f || epilogue warp total %.0f | wait D-full %.0f, H2-empty %.0f, Y-full %.0f, physics %.0f, publish %.0f\n",
            w[W_TOTAL] * per, w[W_X] * per, w[W_DEMPTY] * per, w[W_H1] * per, w[W_FULL] * per, w[W_H2FULL] * per,
            w[E_TOTAL] * per, w[E_DFULL] * per, w[E_H2EMPTY] * per, w[E_YFULL] * per, w[E_PHYS] * per, w[E_PUBLISH] * per);
  }
  return OPZ_OK;
}

}  // namespace opz
