// opz_internal.h — host-side objects behind the opaque handles of include/opz.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "opz_device.cuh"

struct opz_ctx {
  int device = 0;
  int sm_count = 0;
  int smem_optin = 0;        // max dynamic shared memory per block
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int64_t launches = 0;
  // grow-only device scratch (staging of host buffers; BPTT checkpoints)
  void* scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // host-buffer rollouts: copy-in / copy-out streams and events for the slab pipeline (created on first use)
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t slab_ev[34] = {};   // [0] entry fence, [1 + 2i] slab i copied in, [2 + 2i] slab i computed
  // BPTT path (opz_bptt.cu): private stream for graph capture (the adopted stream may be the legacy
  // default stream, which cannot be captured) and the cached graph executables
  cudaStream_t capture_stream = nullptr;
  void* bptt_cache = nullptr;
  // lifetime: models keep their context alive (opz_ctx_destroy is deferred until the last model is gone)
  int live_models = 0;
  bool destroy_requested = false;
  // the training step's collective (opz_comm.cu): ncclComm_t of this rank, or null
  void* nccl_comm = nullptr;
  bool owns_comm = false;
  int comm_ranks = 1, comm_rank = 0;
  int64_t collectives = 0;
};

struct opz_model {
  opz_ctx* ctx = nullptr;
  int nz = 0;
  opz::NetDesc nd{};
  int max_hidden = 0;
  float* w_dev = nullptr;        // [3P] FP32, Flux.destructure order, nets uw|vw|wT
  // tensor-core path (opz_tc2.cu): the weights re-packed as the stage-ordered 16-bit tape; a cache of w_dev
  bool tc_pack_valid = false;
  int tc_pack_precision = -1;    // OPZ_PREC_* the tape was packed for
  void* tc2_pack = nullptr;
  size_t tc2_pack_bytes = 0;
  void* tc2_plan = nullptr;
  // layered tensor path (opz_gemm.cu): per-layer weight images (hi [, lo]) + padded biases; a cache of w_dev
  void* gm_plan = nullptr;
  bool gm_valid = false;
  int gm_precision = -1;
};

namespace opz {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);  // records the message, returns OPZ_ECUDA

#define OPZ_CUDA(expr)                                          \
  do {                                                          \
    cudaError_t _e = (expr);                                    \
    if (_e != cudaSuccess) return ::opz::cuda_fail(_e, #expr);  \
  } while (0)

int ensure_scratch(opz_ctx* ctx, int slot, size_t bytes, void** out);

// opz_comm.cu: in-place sum over the ranks of up to n device buffers, one NCCL group on ctx->stream (no-op for 1 rank)
int comm_allreduce_sum(opz_ctx* ctx, float* const* bufs, const size_t* counts, int n);
void comm_release(opz_ctx* ctx);

struct RolloutCall {
  DevParams P;
  const float* x0;   // device [B][3nz]
  const float* bc;   // device [B][8]
  float* out;        // device [B][n_save][3nz]
  int64_t B;
  int scheme;
  float dt, t0;
  int n_steps, save_every, n_save;
};

// FP32 CUDA-core path (opz_fp32.cu)
int fp32_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c);
int fp32_rhs(opz_ctx* ctx, const opz_model* m, const DevParams& P, const float* x, const float* bc, float t,
             int64_t B, float* dxdt, float* fluxes);
int fp32_diagnostics(opz_ctx* ctx, const opz_model* m, const DevParams& P, const float* sol, const float* bc, int64_t B, int n_save,
                     float t0, float dt_save, const float* target, float* fluxes, float* fluxes_mpp, float* ri, float* loss_t);
int fp32_mlp(opz_ctx* ctx, const opz_model* m, int which_net, const float* x, int64_t Nt, float* y);

struct LossGradCall {
  RolloutCall r;          // r.out unused
  const float* target;    // device [B][n_save][3nz]
  float loss_w[6];
  int64_t B_total;
  float* loss_out;        // device [7]
  float* grad_out;        // device [3P]
  float* gphys_out;       // device [5] or null: d loss / d (nu0, nu_m, dRi, Ric, Pr)
};
int fp32_loss_grad(opz_ctx* ctx, const opz_model* m, const LossGradCall& c);
void bptt_release(opz_ctx* ctx);
// supervised flux pre-training of one net (opz_bptt.cu): device x [N][3nz], bc [N][8], flux_target [N][nz+1]; grad_out device [P]
int fp32_flux_loss_grad(opz_ctx* ctx, const opz_model* m, const DevParams& P, int which_net, const float* x, const float* bc,
                        const float* flux_target, int64_t N, float gradient_scaling, float* loss_out_host, float* grad_out);

// tcgen05 tensor-core path (opz_tc2.cu): CTA pairs, chunk-sized weight stages, level-parallel physics
int tc2_model_prepare(opz_model* m, int precision);          // (re)pack weights after set_weights
void tc2_model_release(opz_model* m);
int tc2_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c, int precision);
bool tc2_supports(const opz_model* m);   // Nz = 32, two or three Dense layers, hidden widths <= 400
// layered tensor path (opz_gemm.cu): one batched tcgen05 GEMM per Dense layer; any shape, split-operand precisions
int gemm_model_prepare(opz_model* m, int precision);
void gemm_model_release(opz_model* m);
int gemm_rollout(opz_ctx* ctx, const opz_model* m, const RolloutCall& c, int precision);
int gemm_mlp(opz_ctx* ctx, const opz_model* m, int which_net, const float* x_dev, int64_t Nt, int precision, float* y_dev);

}  // namespace opz
