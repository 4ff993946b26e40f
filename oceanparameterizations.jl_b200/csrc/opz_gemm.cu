// opz_gemm.cu — the LAYERED tensor-core path: every Dense layer of the three nets as one batched tcgen05 GEMM over the
// column ensemble, activations exchanged through HBM / L2 as 16-bit operand images, physics + Runge-Kutta bookkeeping in an
// FP32 kernel (opz_device.cuh::column_rhs, the same code as the FP32 parity path).
//
// Why a second tensor path beside opz_tc2.cu: the fused persistent kernel keeps a whole column tile's activations in tensor
// memory, which fixes its shapes (Nz = 32, hidden <= 400, one 16-bit image per operand).  This path has no such limits:
//   * any Nz <= 128 and any hidden width / depth (BASELINE configs[4]: Nz = 128, 384-1024-1024-127 nets: here the MLP is
//     9.4 MFLOP per column and right-hand side and a layer's activations for one tile do not fit an SM anyway);
//   * split-operand precisions (include/opz.h): OPZ_PREC_BF16X2 = weights as hi + lo BF16 images (2 MMAs per k-step),
//     OPZ_PREC_BF16X3 = activations split as well (hi*hi + hi*lo + lo*hi, 3 MMAs): ~16 mantissa bits per operand with FP32
//     accumulation in tensor memory, the near-FP32 tensor-core mode;
//   * the batched MLP of predict() (src/predict.jl:12-34) at tensor-core precision.
//
// Replaces, for those configurations, the batched form of NDE! / predict_flux! (training_postprocessing.jl:105-153) inside a
// fixed-step solve (solve_NDE_mutating, :55-159).
//
// Operand image (activations AND weights), chosen so that one 1-D bulk copy fetches an MMA operand block and the epilogue's
// stores coalesce: features in blocks of kKB = 32 (two k-steps); one block of R rows is the K-major no-swizzle core-matrix layout
//   half index = ((k / 8) * R + row) * 8 + k % 8          (LBO = 16 R bytes, SBO = 128 bytes)      [k inside the block]
// activations: R = 128 ocean columns of one tile, blocks ordered [tile][k-block]; weights: R = the outputs of one pass
// (<= 256 = one MMA's N = the tensor-memory columns a CTA allocates), blocks ordered [net][pass][k-block].
// One CTA = one 128-column tile x one pass of outputs x one net: warp 0 streams operand blocks (cp.async.bulk -> mbarrier),
// warp 1 issues tcgen05.mma (A and B from shared memory, D in tensor memory), warps 2..5 drain the accumulator
// (tcgen05.ld -> bias -> activation -> 16-bit hi [, lo] -> 16-byte stores into the next layer's image).
#include <cuda_fp16.h>
#include <algorithm>
#include <cstring>
#include <vector>
#include "opz_internal.h"
#include "opz_ptx.cuh"

namespace opz {
namespace gm {

constexpr int kM = 128;                    // ocean columns per tile
constexpr int kKB = 32;                    // features per operand block (two k-steps): 24 KB (48 KB split) per stage, so that two
                                           // CTAs share an SM and one's epilogue overlaps the other's MMAs; widths pad to 32
constexpr int kNT = 256;                   // outputs per pass
constexpr int kABytes = kM * kKB * 2;      // 8 KB
constexpr int kMaxStages = 8;
// Two shapes of the persistent pair kernel (template parameters kAcc, kEpiWarps):
//   <1, 4>  one 256-column accumulator, four epilogue warps, TWO CTAs per SM: one's epilogue overlaps the other's MMAs, and two
//           producers per SM pull operands — the plain 16-bit modes are bound by operand bytes (configs[4]: 577 against 504
//           algorithmic TFLOP/s for the other shape);
//   <2, 8>  two accumulators, eight epilogue warps, ONE CTA per SM: the epilogue of item i overlaps the k-loop of item i + 1 —
//           the three-MMA split modes do three times the tensor work per byte and gain from the deeper overlap (369 against 337).

enum { MODE_PLAIN = 0, MODE_X2 = 1, MODE_X3 = 2 };

struct LayerArgs {
  const uint16_t* a_hi;       // input image [net?][tile][k-block][kM*kKB]
  const uint16_t* a_lo;       // (MODE_X3)
  size_t a_net;               // halves between the nets' input images (0: shared state image)
  int kb_in;                  // k-blocks of the input
  int k_in;                   // true number of input features
  const uint16_t* w_hi;       // weight image [net][pass][k-block][rows*kKB]
  const uint16_t* w_lo;       // (MODE_X2, MODE_X3)
  size_t w_net;               // halves per net
  const float* bias;          // [net][n_pad]
  int n_out, n_pad, n_pass;   // true outputs, padded to a multiple of kKB, passes of <= 256
  uint16_t* o_hi;             // output image (hidden layers) [net][tile][k-block][kM*kKB]
  uint16_t* o_lo;             // (MODE_X3)
  size_t o_net;
  float* y;                   // last layer: y[(net * n_out + j) * ldy + column]
  int64_t ldy;
  int act;                    // activation of this layer (OPZ_ACT_IDENTITY for the last)
  int mode;
  int fp16;
  int stages;                 // pipeline depth
  int n_nets;                 // nets in this launch (3, or 1 for predict())
  int64_t tiles;              // 128-column tiles
  uint32_t stage_stride;      // bytes per ring slot (a stage of a full 256-row pass)
};

__device__ __forceinline__ uint32_t pack2(float a, float b, bool fp16) { return fp16 ? ptx::pack_f16x2(a, b) : ptx::pack_bf16x2(a, b); }
__device__ __forceinline__ float2 unpack2(uint32_t p, bool fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&p));
  return make_float2(__uint_as_float(p << 16), __uint_as_float(p & 0xFFFF0000u));
}

// Persistent CTA PAIRS (tcgen05 cta_group::2, M = 256): one pair per two SMs walks the work items (tile pair x pass x net,
// (pass, net)-major so that the pairs running at the same time share activation tiles in L2) with the operand ring and the
// barriers carried across items, and TWO accumulators of 256 tensor-memory columns per CTA: the eight epilogue warps drain item i
// while the issuing warp is already in the k-loop of item i + 1.  Each CTA stages its own activation tile and HALF of every
// weight block — the GEMM is bound by L2 -> SM bytes (24 KB per 256 tensor cycles and CTA with whole weight blocks: 27 B/clk/SM
// measured, a third of what the tensor pipe would take), and the pair moves 16 KB for the same MMAs.
template <int kAcc, int kEpiWarps>
__global__ void __launch_bounds__((2 + kEpiWarps) * 32, kAcc == 1 ? 2 : 1) layer_gemm_kernel(const __grid_constant__ LayerArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n_parts_a = a.mode == MODE_X3 ? 2 : 1, n_parts_w = a.mode == MODE_PLAIN ? 1 : 2;
  enum { B_FULL = 0, B_EMPTY = kMaxStages, B_DFULL = 2 * kMaxStages, B_DEMPTY = 2 * kMaxStages + 2, B_N = 2 * kMaxStages + 4 };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * B_N);
  uint8_t* ring = smem + 256;
  const int S = a.stages;
  const int per_tile = a.n_pass * a.n_nets;
  const int64_t n_items = (a.tiles / 2) * per_tile;           // a.tiles is even (the last tile may be padding: zero image, never read back)
  const uint32_t rank = ptx::cluster_ctarank();
  const int64_t item0 = blockIdx.x / 2, item_stride = gridDim.x / 2;
  const uint32_t leader_bars = ptx::mapa(ptx::smem_u32(bars), 0u);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { ptx::mbar_init(bars + B_FULL + s, rank == 0 ? 2 : 1); ptx::mbar_init(bars + B_EMPTY + s, 1); }
      for (int b = 0; b < 2; ++b) { ptx::mbar_init(bars + B_DFULL + b, 1); ptx::mbar_init(bars + B_DEMPTY + b, 2 * kEpiWarps); }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc2<kAcc * kNT>(tmem_slot);
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();   // both CTAs' barriers exist before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // item -> (tile of THIS CTA, pass, net, rows)
  auto decode = [&](int64_t it, int64_t& tile, int& pass, int& net, int& rows) {
    const int64_t pair = it / per_tile;
    const int sub = (int)(it - pair * per_tile);
    tile = 2 * pair + rank;
    pass = sub % a.n_pass;
    net = sub / a.n_pass;
    rows = min(kNT, a.n_pad - pass * kNT);   // outputs of this pass (multiple of kKB)
  };

  if (warp == 0) {
    // ================= operand producer (each CTA: its tile, its half of the weight block) =================
    if (lane == 0) {
      uint32_t gs = 0;   // stages issued so far (all items)
      for (int64_t it = item0; it < n_items; it += item_stride) {
        int64_t tile; int pass, net, rows;
        decode(it, tile, pass, net, rows);
        const uint32_t w_bytes = (uint32_t)(rows / 2) * kKB * 2;
        const uint32_t stage_bytes = n_parts_a * kABytes + n_parts_w * w_bytes;
        const size_t w_off = (size_t)net * a.w_net + (size_t)pass * kNT * kKB * a.kb_in;   // the passes before this one hold 256 rows each
        for (int kb = 0; kb < a.kb_in; ++kb, ++gs) {
          const uint32_t s = gs % (uint32_t)S;
          if (gs >= (uint32_t)S) ptx::mbar_wait(bars + B_EMPTY + s, ((gs / (uint32_t)S) - 1u) & 1u);
          uint8_t* dst = ring + (size_t)s * a.stage_stride;
          ptx::mbar_arrive_expect_tx(bars + B_FULL + s, stage_bytes);
          const size_t a_off = (size_t)net * a.a_net + ((size_t)tile * a.kb_in + kb) * (kM * kKB);
          ptx::bulk_g2s(dst, a.a_hi + a_off, kABytes, bars + B_FULL + s);
          dst += kABytes;
          if (n_parts_a == 2) { ptx::bulk_g2s(dst, a.a_lo + a_off, kABytes, bars + B_FULL + s); dst += kABytes; }
          const size_t wo = w_off + (size_t)kb * rows * kKB + (size_t)rank * (rows / 2) * kKB;   // [k-block][half][(k/8)(rows/2) + row][8]
          ptx::bulk_g2s(dst, a.w_hi + wo, w_bytes, bars + B_FULL + s);
          dst += w_bytes;
          if (n_parts_w == 2) ptx::bulk_g2s(dst, a.w_lo + wo, w_bytes, bars + B_FULL + s);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank != 0) {
    // ================= peer CTA: forward "my half of the stage has landed" to the leader =================
    if (lane == 0) {
      uint32_t gs = 0;
      for (int64_t it = item0; it < n_items; it += item_stride)
        for (int kb = 0; kb < a.kb_in; ++kb, ++gs) {
          const uint32_t s = gs % (uint32_t)S;
          ptx::mbar_wait(bars + B_FULL + s, (gs / (uint32_t)S) & 1u);
          ptx::mbar_arrive_remote_relaxed(leader_bars + 8u * (uint32_t)(B_FULL + s));
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA) =================
    const uint32_t ring_u32 = ptx::smem_u32(ring);
    uint32_t gs = 0, li = 0;   // stages / items so far
    for (int64_t it = item0; it < n_items; it += item_stride, ++li) {
      int64_t tile; int pass, net, rows;
      decode(it, tile, pass, net, rows);
      const uint32_t w_bytes = (uint32_t)(rows / 2) * kKB * 2;
      const uint32_t idesc = ptx::idesc_f16kind(2 * kM, rows, a.fp16 != 0);
      const uint32_t buf = li % (uint32_t)kAcc;
      const uint32_t d = tmem + buf * (uint32_t)kNT;
      if (li >= (uint32_t)kAcc) ptx::mbar_wait(bars + B_DEMPTY + buf, ((li / (uint32_t)kAcc) - 1u) & 1u);   // both epilogues have drained item li - kAcc
      for (int kb = 0; kb < a.kb_in; ++kb, ++gs) {
        const uint32_t s = gs % (uint32_t)S;
        ptx::mbar_wait(bars + B_FULL + s, (gs / (uint32_t)S) & 1u);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t base = ring_u32 + s * a.stage_stride;
          const uint32_t a_hi = base, a_lo = base + kABytes;
          const uint32_t w_hi = base + n_parts_a * kABytes, w_lo = w_hi + w_bytes;
          const int nks = min(kKB / 16, (a.k_in - kb * kKB + 15) / 16);   // k-steps that hold real features
          for (int ks = 0; ks < nks; ++ks) {
            const uint64_t ad = ptx::smem_desc(a_hi + ks * (2 * kM * 16), kM * 16, 128);
            const uint64_t bd = ptx::smem_desc(w_hi + ks * (rows * 16), (rows / 2) * 16, 128);   // this CTA's rows / 2 rows
            ptx::mma_ss2(d, ad, bd, idesc, kb > 0 || ks > 0);
            if (a.mode != MODE_PLAIN) {
              const uint64_t bl = ptx::smem_desc(w_lo + ks * (rows * 16), (rows / 2) * 16, 128);
              ptx::mma_ss2(d, ad, bl, idesc, true);
              if (a.mode == MODE_X3) {
                const uint64_t al = ptx::smem_desc(a_lo + ks * (2 * kM * 16), kM * 16, 128);
                ptx::mma_ss2(d, al, bd, idesc, true);
              }
            }
          }
          ptx::tc_commit2(bars + B_EMPTY + s);
          if (kb == a.kb_in - 1) ptx::tc_commit2(bars + B_DFULL + buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue: this thread owns ocean column r of the tile (tensor-memory lane r); the two warps of a lane
    //                   quadrant take alternate groups of 32 accumulator columns =================
    const int q = warp % 4, r = q * 32 + lane, half = (warp - 2) / 4;
    constexpr int kColStep = 32 * (kEpiWarps / 4);
    const bool fp16 = a.fp16 != 0;
    uint32_t li = 0;
    for (int64_t it = item0; it < n_items; it += item_stride, ++li) {
      int64_t tile; int pass, net, rows;
      decode(it, tile, pass, net, rows);
      const uint32_t buf = li % (uint32_t)kAcc;
      ptx::mbar_wait(bars + B_DFULL + buf, (li / (uint32_t)kAcc) & 1u);
      ptx::tc_fence_after();
      const float* bias = a.bias + (size_t)net * a.n_pad + pass * kNT;
      for (int c0 = half * 32; c0 < rows; c0 += kColStep) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)kNT + (uint32_t)c0, v);
        ptx::tmem_wait_ld();
        if (a.y) {   // last layer: FP32 outputs, feature-major (coalesced over the ocean columns)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = pass * kNT + c0 + j;
            if (n < a.n_out) a.y[((size_t)net * a.n_out + n) * a.ldy + tile * kM + r] = __uint_as_float(v[j]) + bias[c0 + j];
          }
        } else {
          float z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = __uint_as_float(v[j]) + bias[c0 + j];
          switch (a.act) {   // one dispatch per 32 values, not per value
#define OPZ_ACT_CASE(A) case A: _Pragma("unroll") for (int j = 0; j < 32; ++j) z[j] = act_fn<A>(z[j]); break;
            OPZ_ACT_CASE(OPZ_ACT_RELU) OPZ_ACT_CASE(OPZ_ACT_TANH) OPZ_ACT_CASE(OPZ_ACT_MISH) OPZ_ACT_CASE(OPZ_ACT_SWISH) OPZ_ACT_CASE(OPZ_ACT_LEAKYRELU)
#undef OPZ_ACT_CASE
            default: break;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {   // groups of 8 features = one 16-byte chunk of the image
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = g * 8 + 2 * e;
              const float z0 = z[j], z1 = z[j + 1];
              hi[e] = pack2(z0, z1, fp16);
              if (a.mode == MODE_X3) {
                const float2 h = unpack2(hi[e], fp16);
                lo[e] = pack2(z0 - h.x, z1 - h.y, fp16);
              }
            }
            const int n = pass * kNT + c0 + g * 8;            // first feature of the chunk
            const size_t o = (size_t)net * a.o_net + ((size_t)tile * (a.n_pad / kKB) + n / kKB) * (kM * kKB) + ((size_t)((n % kKB) / 8) * kM + r) * 8;
            *reinterpret_cast<uint4*>(a.o_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (a.mode == MODE_X3) *reinterpret_cast<uint4*>(a.o_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      ptx::tc_fence_before();   // this warp's tcgen05.ld of the accumulator are complete (wait::ld above)
      __syncwarp();
      if (lane == 0) {   // the issuer's barriers live in the leader CTA
        if (rank != 0) ptx::mbar_arrive_remote_relaxed(leader_bars + 8u * (uint32_t)(B_DEMPTY + buf));
        else ptx::mbar_arrive_relaxed(bars + B_DEMPTY + buf);
      }
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();   // the peer's tensor / shared memory stay valid until both CTAs are done
  if (warp == 1) ptx::tmem_dealloc2<kAcc * kNT>(tmem);
}

// ---- state image: FP32 stage state xs[i][column] -> 16-bit operand image [tile][k-block] (hi [, lo]) ----
__global__ void pack_state_kernel(const float* __restrict__ xs, int64_t ld, int S3, int kb_in, int64_t B, uint16_t* __restrict__ hi,
                                  uint16_t* __restrict__ lo, int fp16) {
  const int64_t tile = blockIdx.x;
  const int kb = blockIdx.y;
  const int r = threadIdx.x % kM, g = threadIdx.x / kM;   // 128 columns x kKB / 8 feature groups
  const int64_t col = tile * kM + r;
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i0 = kb * kKB + g * 8 + 2 * e;
    const float x0 = (col < B && i0 < S3) ? xs[(size_t)i0 * ld + col] : 0.0f;
    const float x1 = (col < B && i0 + 1 < S3) ? xs[(size_t)(i0 + 1) * ld + col] : 0.0f;
    ph[e] = pack2(x0, x1, fp16 != 0);
    const float2 h = unpack2(ph[e], fp16 != 0);
    pl[e] = pack2(x0 - h.x, x1 - h.y, fp16 != 0);
  }
  const size_t o = ((size_t)tile * kb_in + kb) * (kM * kKB) + ((size_t)g * kM + r) * 8;
  *reinterpret_cast<uint4*>(hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// ---- physics + Runge-Kutta bookkeeping: one thread per (ocean column, chunk of kPhysCells cells), state feature-major [i][column]
//      in HBM.  A whole column per thread left an ensemble of 65 536 columns of Nz = 128 with 14 warps per SM, each walking 128
//      faces of dependent loads; split by cells the faces are independent (opz_device.cuh::column_rhs_range: the same expressions
//      per face, bit-identical tendencies). ----
constexpr int kPhysCells = 16;
struct PhysArgs {
  DevParams P;
  const float* bc;    // [B][8]
  const float* y;     // [(net * nout + j)][ld]
  float* xn;          // [S3][ld] state at the start of the step
  float* kacc;        // [S3][ld]
  const float* xin;   // [S3][ld] stage input
  float* xout;        // [S3][ld] next stage input (may alias xin's partner buffer)
  int64_t B, ld, col0;
  int scheme, stage;
  float h, ts;
};
__global__ void __launch_bounds__(128) physics_rk_kernel(const __grid_constant__ PhysArgs a) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.B) return;
  const DevParams& P = a.P;
  const int nz = P.nz, nout = nz - 1;
  const int cell0 = (int)blockIdx.y * kPhysCells, cell1 = min(nz, cell0 + kPhysCells);
  float bc[OPZ_BC_STRIDE];
#pragma unroll
  for (int q = 0; q < OPZ_BC_STRIDE; ++q) bc[q] = a.bc[(a.col0 + c) * OPZ_BC_STRIDE + q];
  const int64_t ld = a.ld;
  const int scheme = a.scheme, st = a.stage;
  const float h = a.h;
  auto X = [&](int i) { return a.xin[(size_t)i * ld + c]; };
  auto Y = [&](int net, int j) { return a.y[((size_t)net * nout + j) * ld + c]; };
  auto K = [&](int i, float k) {   // the oracle's operation order (as in rollout_fp32_kernel)
    const size_t o = (size_t)i * ld + c;
    const float xo = a.xn[o];
    if (scheme == OPZ_RK4) {
      if (st == 0) { a.kacc[o] = k; a.xout[o] = xo + (h * 0.5f) * k; }
      else if (st == 1) { a.kacc[o] = a.kacc[o] + 2.0f * k; a.xout[o] = xo + (h * 0.5f) * k; }
      else if (st == 2) { a.kacc[o] = a.kacc[o] + 2.0f * k; a.xout[o] = xo + h * k; }
      else { const float xv = xo + (h / 6.0f) * (a.kacc[o] + k); a.xn[o] = xv; a.xout[o] = xv; }
    } else if (scheme == OPZ_RK2) {
      if (st == 0) { a.kacc[o] = k; a.xout[o] = xo + h * k; }
      else { const float xv = xo + (h * 0.5f) * (a.kacc[o] + k); a.xn[o] = xv; a.xout[o] = xv; }
    } else {
      const float xv = xo + h * k; a.xn[o] = xv; a.xout[o] = xv;
    }
  };
  column_rhs_range(P, bc, a.ts, X, Y, K, cell0, cell1);
}

// x0[B][S3] (row per column) <-> feature-major [S3][ld]
__global__ void load_state_kernel(const float* __restrict__ x0, int S3, int64_t B, int64_t ld, float* __restrict__ xn, float* __restrict__ xs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * S3) return;
  const int64_t c = e / S3;
  const int i = (int)(e - c * S3);
  const float v = x0[e];
  xn[(size_t)i * ld + c] = v;
  xs[(size_t)i * ld + c] = v;
}
__global__ void save_state_kernel(const float* __restrict__ xn, int S3, int64_t B, int64_t ld, float* __restrict__ out, int n_save, int slot) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * S3) return;
  const int64_t c = e / S3;
  const int i = (int)(e - c * S3);
  out[(c * n_save + slot) * S3 + i] = xn[(size_t)i * ld + c];
}
// y[(j)][ld] (one net) -> out[Nt][nout]: the batched MLP of predict()
__global__ void store_mlp_kernel(const float* __restrict__ y, int nout, int64_t Nt, int64_t ld, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Nt * nout) return;
  const int64_t c = e / nout;
  const int j = (int)(e - c * nout);
  out[e] = y[(size_t)j * ld + c];
}
// x[Nt][S3] -> feature-major [S3][ld]
__global__ void load_rows_kernel(const float* __restrict__ x, int S3, int64_t Nt, int64_t ld, float* __restrict__ xs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Nt * S3) return;
  const int64_t c = e / S3;
  xs[(size_t)(e - c * S3) * ld + c] = x[e];
}

// ---- host: weight images ---------------------------------------------------------------------------------------------
struct Plan {
  int L = 0, mode = 0;
  bool fp16 = false;
  int kb[kMaxLayers + 1];        // k-blocks of the input of layer l (kb[L] unused)
  int n_pad[kMaxLayers];         // outputs of layer l padded to 64
  size_t w_off[kMaxLayers];      // halves, inside one net's image
  size_t w_net = 0;              // halves per net (one of hi / lo)
  size_t b_off[kMaxLayers];      // floats inside the bias block of one net... (bias block is [layer][net][n_pad])
  size_t bias_floats = 0;
  uint16_t* w_hi = nullptr;
  uint16_t* w_lo = nullptr;
  float* bias = nullptr;
  void* block = nullptr;
};

static uint16_t f16_bits(float f) { const __half h = __float2half_rn(fminf(fmaxf(f, -65504.0f), 65504.0f)); uint16_t b; std::memcpy(&b, &h, 2); return b; }
static float f16_val(uint16_t b) { __half h; std::memcpy(&h, &b, 2); return __half2float(h); }
static uint16_t bf16_bits(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_val(uint16_t b) { const uint32_t u = (uint32_t)b << 16; float f; std::memcpy(&f, &u, 4); return f; }

}  // namespace gm

void gemm_model_release(opz_model* m) {
  gm::Plan* pl = static_cast<gm::Plan*>(m->gm_plan);
  if (pl) {
    if (pl->block) cudaFree(pl->block);
    delete pl;
  }
  m->gm_plan = nullptr;
  m->gm_valid = false;
}

static int precision_mode(int precision, int* mode, bool* fp16) {
  switch (precision) {
    case OPZ_PREC_BF16: *mode = gm::MODE_PLAIN; *fp16 = false; return OPZ_OK;
    case OPZ_PREC_FP16: *mode = gm::MODE_PLAIN; *fp16 = true; return OPZ_OK;
    case OPZ_PREC_BF16X2: *mode = gm::MODE_X2; *fp16 = false; return OPZ_OK;
    case OPZ_PREC_BF16X3: *mode = gm::MODE_X3; *fp16 = false; return OPZ_OK;
    case OPZ_PREC_FP16X2: *mode = gm::MODE_X2; *fp16 = true; return OPZ_OK;
    case OPZ_PREC_FP16X3: *mode = gm::MODE_X3; *fp16 = true; return OPZ_OK;
    default: set_error("layered tensor path: unknown precision"); return OPZ_EINVAL;
  }
}

int gemm_model_prepare(opz_model* m, int precision) {
  using namespace gm;
  int mode; bool fp16;
  int rc = precision_mode(precision, &mode, &fp16);
  if (rc) return rc;
  const NetDesc& nd = m->nd;
  const int L = nd.n_layers;
  OPZ_CUDA(cudaSetDevice(m->ctx->device));
  gemm_model_release(m);
  Plan* pl = new Plan();
  m->gm_plan = pl;
  pl->L = L; pl->mode = mode; pl->fp16 = fp16;
  size_t off = 0, boff = 0;
  for (int l = 0; l < L; ++l) {
    pl->kb[l] = (nd.sizes[l] + kKB - 1) / kKB;
    pl->n_pad[l] = (nd.sizes[l + 1] + kKB - 1) / kKB * kKB;
    pl->w_off[l] = off;
    off += (size_t)pl->n_pad[l] * pl->kb[l] * kKB;
    pl->b_off[l] = boff;
    boff += 3 * (size_t)pl->n_pad[l];
    if (l > 0 && pl->kb[l] * kKB != pl->n_pad[l - 1]) { set_error("layered tensor path: internal padding mismatch"); return OPZ_EINVAL; }
  }
  pl->w_net = off;
  pl->bias_floats = boff;
  std::vector<float> w((size_t)3 * nd.P);
  OPZ_CUDA(cudaMemcpyAsync(w.data(), m->w_dev, sizeof(float) * w.size(), cudaMemcpyDeviceToHost, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
  const bool split_w = mode != MODE_PLAIN;
  std::vector<uint16_t> hi(3 * off, 0), lo(split_w ? 3 * off : 0, 0);
  std::vector<float> bias(boff, 0.0f);
  for (int net = 0; net < 3; ++net) {
    const float* wn = w.data() + (size_t)net * nd.P;
    for (int l = 0; l < L; ++l) {
      const int K = nd.sizes[l], N = nd.sizes[l + 1], n_pad = pl->n_pad[l], kbn = pl->kb[l];
      for (int n = 0; n < N; ++n) bias[pl->b_off[l] + (size_t)net * n_pad + n] = wn[nd.b_off[l] + n];
      for (int pass = 0; pass * kNT < n_pad; ++pass) {
        const int rows = std::min(kNT, n_pad - pass * kNT);
        const size_t pbase = (size_t)net * off + pl->w_off[l] + (size_t)pass * kNT * kKB * kbn;
        for (int kb = 0; kb < kbn; ++kb)
          for (int kk = 0; kk < kKB; ++kk) {
            const int k = kb * kKB + kk;
            if (k >= K) continue;
            for (int r = 0; r < rows; ++r) {
              const int n = pass * kNT + r;
              if (n >= N) continue;
              const float v = wn[nd.w_off[l] + (size_t)k * N + n];   // Flux layout: element (k_in, n_out)
              // [k-block][half of the rows (one per CTA of the pair)][(k / 8) (rows / 2) + row][k % 8]
              const int hr = rows / 2, half = r / hr, rl = r - half * hr;
              const size_t o = pbase + (size_t)kb * rows * kKB + (size_t)half * hr * kKB + ((size_t)(kk / 8) * hr + rl) * 8 + kk % 8;
              const uint16_t h = fp16 ? f16_bits(v) : bf16_bits(v);
              hi[o] = h;
              if (split_w) { const float res = v - (fp16 ? f16_val(h) : bf16_val(h)); lo[o] = fp16 ? f16_bits(res) : bf16_bits(res); }
            }
          }
      }
    }
  }
  const size_t hb = (hi.size() * 2 + 255) / 256 * 256, lb = (lo.size() * 2 + 255) / 256 * 256, bb = bias.size() * 4;
  if (cudaMalloc(&pl->block, hb + lb + bb + 256) != cudaSuccess) { (void)cudaGetLastError(); set_error("device allocation failed"); return OPZ_ENOMEM; }
  pl->w_hi = static_cast<uint16_t*>(pl->block);
  pl->w_lo = split_w ? reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(pl->block) + hb) : nullptr;
  pl->bias = reinterpret_cast<float*>(static_cast<uint8_t*>(pl->block) + hb + lb);
  OPZ_CUDA(cudaMemcpyAsync(pl->w_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice, m->ctx->stream));
  if (split_w) OPZ_CUDA(cudaMemcpyAsync(pl->w_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice, m->ctx->stream));
  OPZ_CUDA(cudaMemcpyAsync(pl->bias, bias.data(), bb, cudaMemcpyHostToDevice, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
  m->gm_valid = true;
  m->gm_precision = precision;
  return OPZ_OK;
}

namespace gm {

// device buffers of one chunk of C ocean columns (C a multiple of 128), carved from one scratch block
struct Work {
  float *xn, *kacc, *xs[2], *y;
  uint16_t *x_hi, *x_lo;
  uint16_t *h_hi[kMaxLayers], *h_lo[kMaxLayers];   // output image of hidden layer l (three nets)
  size_t h_net[kMaxLayers];
};
static size_t carve(const Plan& pl, const NetDesc& nd, int nz, int64_t C, uint8_t* base, Work* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
  const int S3 = 3 * nz, nout = nz - 1;
  const int64_t tiles = C / kM;
  Work tmp;
  Work& ww = w ? *w : tmp;
  ww.xn = (float*)take(sizeof(float) * S3 * C);
  ww.kacc = (float*)take(sizeof(float) * S3 * C);
  ww.xs[0] = (float*)take(sizeof(float) * S3 * C);
  ww.xs[1] = (float*)take(sizeof(float) * S3 * C);
  ww.y = (float*)take(sizeof(float) * 3 * nout * C);
  const size_t xi = (size_t)tiles * pl.kb[0] * kM * kKB * 2;
  ww.x_hi = (uint16_t*)take(xi);
  ww.x_lo = pl.mode == MODE_X3 ? (uint16_t*)take(xi) : nullptr;
  for (int l = 0; l < pl.L - 1; ++l) {
    ww.h_net[l] = (size_t)tiles * (pl.n_pad[l] / kKB) * kM * kKB;
    ww.h_hi[l] = (uint16_t*)take(3 * ww.h_net[l] * 2);
    ww.h_lo[l] = pl.mode == MODE_X3 ? (uint16_t*)take(3 * ww.h_net[l] * 2) : nullptr;
  }
  (void)nd;
  return off;
}

static size_t layer_stage(const Plan& pl, int l) {   // per CTA of the pair: its activation tile + half of the weight block
  const int rows = std::min(kNT, pl.n_pad[l]);
  return (pl.mode == MODE_X3 ? 2 : 1) * (size_t)kABytes + (pl.mode == MODE_PLAIN ? 1 : 2) * (size_t)(rows / 2) * kKB * 2;
}
static size_t layer_smem(const Plan& pl, int l, int stages) { return 256 + layer_stage(pl, l) * stages; }

// one MLP evaluation of the three nets (or of net `only`, if >= 0) on the state image of `tiles` tiles
static int run_layers(opz_ctx* ctx, const opz_model* m, const Plan& pl, const Work& w, int64_t tiles, int64_t ld, int only) {
  const NetDesc& nd = m->nd;
  for (int l = 0; l < pl.L; ++l) {
    LayerArgs a{};
    const bool first = l == 0, last = l == pl.L - 1;
    a.a_hi = first ? w.x_hi : w.h_hi[l - 1];
    a.a_lo = first ? w.x_lo : w.h_lo[l - 1];
    a.a_net = first ? 0 : w.h_net[l - 1];
    a.kb_in = pl.kb[l];
    a.k_in = nd.sizes[l];
    a.w_hi = pl.w_hi + pl.w_off[l];
    a.w_lo = pl.w_lo ? pl.w_lo + pl.w_off[l] : nullptr;
    a.w_net = pl.w_net;
    a.bias = pl.bias + pl.b_off[l];
    a.n_out = nd.sizes[l + 1];
    a.n_pad = pl.n_pad[l];
    a.n_pass = (pl.n_pad[l] + kNT - 1) / kNT;
    a.o_hi = last ? nullptr : w.h_hi[l];
    a.o_lo = last ? nullptr : w.h_lo[l];
    a.o_net = last ? 0 : w.h_net[l];
    a.y = last ? w.y : nullptr;
    a.ldy = ld;
    a.act = last ? OPZ_ACT_IDENTITY : nd.act;
    a.mode = pl.mode;
    a.fp16 = pl.fp16 ? 1 : 0;
    int stages = kMaxStages;   // one persistent CTA per SM: as deep a ring as shared memory holds
    // as deep as fits under the 164 KB shared-memory carve-out: the next one (228 KB) leaves 28 KB of L1 for the epilogue's bias
    // reads and measured erratic (configs[4] FP16X3: 216 - 283 TFLOP/s with 4 stages of 48 KB against 350 with 3)
    const bool deep = pl.mode == MODE_X3;   // kernel shape, see the template's comment
    while (stages > 2 && layer_smem(pl, l, stages) > (size_t)(deep ? 160 : 100) * 1024) --stages;   // under the carve-out that leaves L1 / room for two CTAs
    while (stages > 1 && layer_smem(pl, l, stages) > (size_t)ctx->smem_optin) --stages;
    if (layer_smem(pl, l, stages) > (size_t)ctx->smem_optin) { set_error("layered tensor path: operand stage exceeds shared memory"); return OPZ_EUNSUP; }
    a.stages = stages;
    a.stage_stride = (uint32_t)layer_stage(pl, l);
    a.tiles = tiles;
    const size_t smem = layer_smem(pl, l, a.stages);
    auto kern = deep ? layer_gemm_kernel<2, 8> : layer_gemm_kernel<1, 4>;
    OPZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (only >= 0) {   // a single net: shift the per-net bases
      a.w_hi += (size_t)only * pl.w_net;
      if (a.w_lo) a.w_lo += (size_t)only * pl.w_net;
      a.bias += (size_t)only * a.n_pad;
      if (!first) { a.a_hi += (size_t)only * a.a_net; if (a.a_lo) a.a_lo += (size_t)only * a.a_net; }
      if (!last) { a.o_hi += (size_t)only * a.o_net; if (a.o_lo) a.o_lo += (size_t)only * a.o_net; }
      else a.y += (size_t)only * a.n_out * ld;
    }
    a.n_nets = only >= 0 ? 1 : 3;
    if (tiles % 2) { set_error("layered tensor path: internal error (odd tile count)"); return OPZ_EINVAL; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * std::min<int64_t>((tiles / 2) * a.n_pass * a.n_nets, deep ? ctx->sm_count / 2 : ctx->sm_count)));
    cfg.blockDim = dim3(deep ? 320 : 192);
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
  }
  return OPZ_OK;
}

}  // namespace gm

int gemm_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c, int precision) {
  using namespace gm;
  const Plan* pl = static_cast<const Plan*>(m->gm_plan);
  if (!pl || !m->gm_valid || m->gm_precision != precision) { set_error("layered tensor path: model not prepared for this precision"); return OPZ_EINVAL; }
  if (c.scheme == OPZ_IMEX_EULER) { set_error("OPZ_IMEX_EULER is built for OPZ_PREC_FP32 only"); return OPZ_EUNSUP; }
  const NetDesc& nd = m->nd;
  const int nz = m->nz, S3 = 3 * nz;
  cudaStream_t st = ctx->stream;
  const int stages = c.scheme == OPZ_RK4 ? 4 : (c.scheme == OPZ_RK2 ? 2 : 1);
  // chunk of columns whose buffers fit a quarter of the free memory (the whole ensemble when it can)
  size_t free_b = 0, total_b = 0;
  OPZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->scratch_bytes[7];
  int64_t C = (c.B + 2 * kM - 1) / (2 * kM) * (2 * kM);   // whole tile PAIRS (the GEMM kernel runs as CTA pairs)
  const size_t per128 = carve(*pl, nd, nz, kM, nullptr, nullptr);
  const int64_t cmax = std::max<int64_t>(1, (int64_t)((free_b / 4) / per128) / 2) * (2 * kM);
  if (C > cmax) C = cmax;
  if (const char* e = getenv("OPZ_GEMM_CHUNK")) { const int64_t v = atoll(e) / (2 * kM) * (2 * kM); if (v >= 2 * kM && v < C) C = v; }
  uint8_t* base;
  int rc;
  if ((rc = ensure_scratch(ctx, 7, carve(*pl, nd, nz, C, nullptr, nullptr), (void**)&base))) return rc;
  Work w;
  carve(*pl, nd, nz, C, base, &w);
  for (int64_t c0 = 0; c0 < c.B; c0 += C) {
    const int64_t nb = std::min<int64_t>(C, c.B - c0);
    const int64_t tiles = (nb + 2 * kM - 1) / (2 * kM) * 2;   // an even number: the last tile may be padding
    const unsigned eg = (unsigned)((nb * S3 + 255) / 256);
    load_state_kernel<<<eg, 256, 0, st>>>(c.x0 + c0 * S3, S3, nb, C, w.xn, w.xs[0]);
    if (c.save_every > 0) save_state_kernel<<<eg, 256, 0, st>>>(w.xn, S3, nb, C, c.out + c0 * c.n_save * S3, c.n_save, 0);
    ctx->launches += 2;
    int cur = 0;
    for (int n = 0; n < c.n_steps; ++n) {
      const float tn = c.t0 + (float)n * c.dt;
      for (int sg = 0; sg < stages; ++sg) {
        float ts = tn;
        if (c.scheme == OPZ_RK4) ts = (sg == 0) ? tn : (sg == 3 ? tn + c.dt : tn + c.dt * 0.5f);
        else if (c.scheme == OPZ_RK2) ts = (sg == 0) ? tn : tn + c.dt;
        pack_state_kernel<<<dim3((unsigned)tiles, (unsigned)pl->kb[0]), kM * (kKB / 8), 0, st>>>(w.xs[cur], C, S3, pl->kb[0], nb, w.x_hi, w.x_lo, pl->fp16 ? 1 : 0);
        ctx->launches += 1;
        if ((rc = run_layers(ctx, m, *pl, w, tiles, C, -1))) return rc;
        PhysArgs pa{};
        pa.P = c.P; pa.bc = c.bc; pa.y = w.y; pa.xn = w.xn; pa.kacc = w.kacc; pa.xin = w.xs[cur]; pa.xout = w.xs[cur ^ 1];
        pa.B = nb; pa.ld = C; pa.col0 = c0; pa.scheme = c.scheme; pa.stage = sg; pa.h = c.dt; pa.ts = ts;
        physics_rk_kernel<<<dim3((unsigned)((nb + 127) / 128), (unsigned)((nz + kPhysCells - 1) / kPhysCells)), 128, 0, st>>>(pa);
        ctx->launches += 1;
        cur ^= 1;
      }
      if (c.save_every > 0 && (n + 1) % c.save_every == 0) {
        save_state_kernel<<<eg, 256, 0, st>>>(w.xn, S3, nb, C, c.out + c0 * c.n_save * S3, c.n_save, (n + 1) / c.save_every);
        ctx->launches += 1;
      }
    }
    if (c.save_every <= 0) { save_state_kernel<<<eg, 256, 0, st>>>(w.xn, S3, nb, C, c.out + c0 * S3, 1, 0); ctx->launches += 1; }
    OPZ_CUDA(cudaGetLastError());
  }
  return OPZ_OK;
}

// batched MLP of one net at tensor-core precision (predict(), src/predict.jl:12-34): x[Nt][3nz] -> y[Nt][nz-1], device pointers
int gemm_mlp(opz_ctx* ctx, const opz_model* m, int which_net, const float* x, int64_t Nt, int precision, float* y) {
  using namespace gm;
  const Plan* pl = static_cast<const Plan*>(m->gm_plan);
  if (!pl || !m->gm_valid || m->gm_precision != precision) { set_error("layered tensor path: model not prepared for this precision"); return OPZ_EINVAL; }
  const NetDesc& nd = m->nd;
  const int nz = m->nz, S3 = 3 * nz, nout = nz - 1;
  cudaStream_t st = ctx->stream;
  size_t free_b = 0, total_b = 0;
  OPZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->scratch_bytes[7];
  int64_t C = (Nt + 2 * kM - 1) / (2 * kM) * (2 * kM);
  const size_t per128 = carve(*pl, nd, nz, kM, nullptr, nullptr);
  const int64_t cmax = std::max<int64_t>(1, (int64_t)((free_b / 4) / per128) / 2) * (2 * kM);
  if (C > cmax) C = cmax;
  uint8_t* base;
  int rc;
  if ((rc = ensure_scratch(ctx, 7, carve(*pl, nd, nz, C, nullptr, nullptr), (void**)&base))) return rc;
  Work w;
  carve(*pl, nd, nz, C, base, &w);
  for (int64_t c0 = 0; c0 < Nt; c0 += C) {
    const int64_t nb = std::min<int64_t>(C, Nt - c0);
    const int64_t tiles = (nb + 2 * kM - 1) / (2 * kM) * 2;   // an even number: the last tile may be padding
    load_rows_kernel<<<(unsigned)((nb * S3 + 255) / 256), 256, 0, st>>>(x + c0 * S3, S3, nb, C, w.xs[0]);
    pack_state_kernel<<<dim3((unsigned)tiles, (unsigned)pl->kb[0]), kM * (kKB / 8), 0, st>>>(w.xs[0], C, S3, pl->kb[0], nb, w.x_hi, w.x_lo, pl->fp16 ? 1 : 0);
    ctx->launches += 2;
    if ((rc = run_layers(ctx, m, *pl, w, tiles, C, which_net))) return rc;
    store_mlp_kernel<<<(unsigned)((nb * nout + 255) / 256), 256, 0, st>>>(w.y + (size_t)which_net * nout * C, nout, nb, C, y + c0 * nout);
    ctx->launches += 1;
    OPZ_CUDA(cudaGetLastError());
  }
  return OPZ_OK;
}

}  // namespace opz
