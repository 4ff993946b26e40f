// opz_capi.cu — the extern "C" surface declared in include/opz.h.
// Host-pointer entry points stage through grow-only device scratch owned by the context and
// are blocking; `_dev` entry points enqueue on the context stream and return.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include "opz_internal.h"

namespace opz {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return OPZ_ECUDA;
}

int ensure_scratch(opz_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) {
      OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
      OPZ_CUDA(cudaFree(ctx->scratch[slot]));
      ctx->scratch[slot] = nullptr;
      ctx->scratch_bytes[slot] = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], bytes);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("device allocation of " + std::to_string(bytes) + " bytes failed");
      return OPZ_ENOMEM;
    }
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->scratch[slot];
  return OPZ_OK;
}

static int check_common(const opz_ctx* ctx, const opz_model* m, const opz_params* p) {
  if (!ctx || !m || !p) { set_error("null ctx/model/params"); return OPZ_EINVAL; }
  if (m->ctx != ctx) { set_error("model belongs to another context"); return OPZ_EINVAL; }
  if (p->nz != m->nz) { set_error("params.nz does not match the model"); return OPZ_EINVAL; }
  return OPZ_OK;
}

static int check_rollout(int64_t B, int scheme, int n_steps, int save_every) {
  if (B < 0 || n_steps < 0) { set_error("negative B or n_steps"); return OPZ_EINVAL; }
  if (scheme != OPZ_EULER && scheme != OPZ_RK2 && scheme != OPZ_RK4 && scheme != OPZ_IMEX_EULER) { set_error("unknown scheme"); return OPZ_EINVAL; }
  if (save_every > 0 && n_steps % save_every != 0) {
    set_error("n_steps must be a multiple of save_every (ts must be multiples of dt)");
    return OPZ_EINVAL;
  }
  return OPZ_OK;
}

static int n_save_of(int n_steps, int save_every) { return save_every > 0 ? n_steps / save_every + 1 : 1; }

}  // namespace opz

using namespace opz;

extern "C" {

int opz_version(void) { return OPZ_VERSION; }
const char* opz_last_error(void) { return g_err.c_str(); }

int opz_ctx_create(int device, opz_ctx** out) {
  if (!out) { set_error("null out"); return OPZ_EINVAL; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device: libopz has no CPU fallback");
    return OPZ_ENODEV;
  }
  if (device < 0 || device >= n) { set_error("device index out of range"); return OPZ_EINVAL; }
  cudaDeviceProp prop{};
  OPZ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("device '") + prop.name + "' is not sm_100: libopz is built for sm_100a only");
    return OPZ_ENODEV;
  }
  OPZ_CUDA(cudaSetDevice(device));
  opz_ctx* c = new (std::nothrow) opz_ctx();
  if (!c) { set_error("host allocation failed"); return OPZ_ENOMEM; }
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
  c->owns_stream = true;
  *out = c;
  return OPZ_OK;
}

static void ctx_free(opz_ctx* ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  bptt_release(ctx);
  for (int i = 0; i < 8; ++i)
    if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  for (cudaEvent_t e : ctx->slab_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int opz_ctx_destroy(opz_ctx* ctx) {
  if (!ctx) return OPZ_OK;
  if (ctx->live_models > 0) {   // models still point at this context: the last opz_model_destroy frees it
    ctx->destroy_requested = true;
    return OPZ_OK;
  }
  ctx_free(ctx);
  return OPZ_OK;
}

int opz_ctx_set_stream(opz_ctx* ctx, void* cuda_stream) {
  if (!ctx) { set_error("null ctx"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->owns_stream && ctx->stream) OPZ_CUDA(cudaStreamDestroy(ctx->stream));
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->owns_stream = false;
  return OPZ_OK;
}

int opz_ctx_synchronize(opz_ctx* ctx) {
  if (!ctx) { set_error("null ctx"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPZ_OK;
}

int64_t opz_ctx_launch_count(const opz_ctx* ctx) { return ctx ? ctx->launches : 0; }

int opz_model_create(opz_ctx* ctx, int nz, int n_sizes, const int* sizes, int act, const float* w_uw,
                     const float* w_vw, const float* w_wT, opz_model** out) {
  if (!ctx || !sizes || !out || !w_uw || !w_vw || !w_wT) { set_error("null argument"); return OPZ_EINVAL; }
  *out = nullptr;
  const int L = n_sizes - 1;
  if (L < 1 || L > kMaxLayers) { set_error("1..6 Dense layers supported"); return OPZ_EINVAL; }
  if (nz < 2 || nz > kMaxNz) { set_error("nz must be in [2, 128]"); return OPZ_EINVAL; }
  if (sizes[0] != 3 * nz || sizes[L] != nz - 1) {
    set_error("the nets must map 3*nz inputs to nz-1 interior-face outputs");
    return OPZ_EINVAL;
  }
  if (act < OPZ_ACT_IDENTITY || act > OPZ_ACT_LEAKYRELU) { set_error("unknown activation"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  opz_model* m = new (std::nothrow) opz_model();
  if (!m) { set_error("host allocation failed"); return OPZ_ENOMEM; }
  m->ctx = ctx;
  m->nz = nz;
  m->nd.n_layers = L;
  m->nd.act = act;
  int64_t off = 0;
  m->max_hidden = 0;
  for (int l = 0; l <= L; ++l) {
    if (sizes[l] < 1) { delete m; set_error("layer sizes must be positive"); return OPZ_EINVAL; }
    m->nd.sizes[l] = sizes[l];
    if (l > 0 && l < L && sizes[l] > m->max_hidden) m->max_hidden = sizes[l];
  }
  for (int l = 0; l < L; ++l) {
    m->nd.w_off[l] = off; off += (int64_t)sizes[l] * sizes[l + 1];
    m->nd.b_off[l] = off; off += sizes[l + 1];
  }
  m->nd.P = off;
  cudaError_t e = cudaMalloc(&m->w_dev, sizeof(float) * 3 * (size_t)off);
  if (e != cudaSuccess) { delete m; (void)cudaGetLastError(); set_error("device allocation failed"); return OPZ_ENOMEM; }
  const float* src[3] = {w_uw, w_vw, w_wT};
  for (int i = 0; i < 3; ++i) {
    e = cudaMemcpyAsync(m->w_dev + (size_t)i * off, src[i], sizeof(float) * (size_t)off, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { cudaFree(m->w_dev); delete m; return cuda_fail(e, "weights H2D"); }
  }
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { cudaFree(m->w_dev); delete m; return cuda_fail(e, "weights H2D sync"); }
  m->tc_pack_valid = false;
  m->gm_valid = false;
  ctx->live_models += 1;
  *out = m;
  return OPZ_OK;
}

int opz_model_set_weights(opz_model* m, const float* w) {
  if (!m || !w) { set_error("null argument"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(m->ctx->device));
  OPZ_CUDA(cudaMemcpyAsync(m->w_dev, w, sizeof(float) * 3 * (size_t)m->nd.P, cudaMemcpyHostToDevice, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
  m->tc_pack_valid = false;
  m->gm_valid = false;
  return OPZ_OK;
}

int opz_model_set_weights_dev(opz_model* m, const float* w) {
  if (!m || !w) { set_error("null argument"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(m->ctx->device));
  OPZ_CUDA(cudaMemcpyAsync(m->w_dev, w, sizeof(float) * 3 * (size_t)m->nd.P, cudaMemcpyDeviceToDevice, m->ctx->stream));
  m->tc_pack_valid = false;
  m->gm_valid = false;
  return OPZ_OK;
}

int opz_model_get_weights(const opz_model* m, float* w) {
  if (!m || !w) { set_error("null argument"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(m->ctx->device));
  OPZ_CUDA(cudaMemcpyAsync(w, m->w_dev, sizeof(float) * 3 * (size_t)m->nd.P, cudaMemcpyDeviceToHost, m->ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return OPZ_OK;
}

int64_t opz_model_n_params(const opz_model* m) { return m ? m->nd.P : 0; }

int opz_model_destroy(opz_model* m) {
  if (!m) return OPZ_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  tc2_model_release(m);
  gemm_model_release(m);
  if (m->w_dev) cudaFree(m->w_dev);
  opz_ctx* ctx = m->ctx;
  delete m;
  if (--ctx->live_models == 0 && ctx->destroy_requested) ctx_free(ctx);
  return OPZ_OK;
}

// ---- single RHS -------------------------------------------------------------------------------
int opz_rhs(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x, const float* bc, float t,
            int64_t B, int precision, float* dxdt, float* fluxes) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if (!x || !bc || !dxdt || B < 0) { set_error("null buffer or negative B"); return OPZ_EINVAL; }
  if (precision != OPZ_PREC_FP32) { set_error("opz_rhs: only OPZ_PREC_FP32"); return OPZ_EUNSUP; }
  if (B == 0) return OPZ_OK;
  OPZ_CUDA(cudaSetDevice(ctx->device));
  const size_t S = 3 * (size_t)m->nz, nf = 3 * (size_t)(m->nz + 1);
  float *dx, *dbc, *dk, *dfl = nullptr;
  if ((rc = ensure_scratch(ctx, 0, sizeof(float) * B * S, (void**)&dx))) return rc;
  if ((rc = ensure_scratch(ctx, 1, sizeof(float) * B * OPZ_BC_STRIDE, (void**)&dbc))) return rc;
  if ((rc = ensure_scratch(ctx, 2, sizeof(float) * B * S, (void**)&dk))) return rc;
  if (fluxes && (rc = ensure_scratch(ctx, 3, sizeof(float) * B * nf, (void**)&dfl))) return rc;
  OPZ_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * B * S, cudaMemcpyHostToDevice, ctx->stream));
  OPZ_CUDA(cudaMemcpyAsync(dbc, bc, sizeof(float) * B * OPZ_BC_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = fp32_rhs(ctx, m, make_dev_params(*p), dx, dbc, t, B, dk, dfl))) return rc;
  OPZ_CUDA(cudaMemcpyAsync(dxdt, dk, sizeof(float) * B * S, cudaMemcpyDeviceToHost, ctx->stream));
  if (fluxes) OPZ_CUDA(cudaMemcpyAsync(fluxes, dfl, sizeof(float) * B * nf, cudaMemcpyDeviceToHost, ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPZ_OK;
}

// ---- forward rollout --------------------------------------------------------------------------
int opz_rollout_forward_dev(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0,
                            const float* bc, int64_t B, int scheme, float dt_nd, float t0, int n_steps,
                            int save_every, int precision, float* out) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if ((rc = check_rollout(B, scheme, n_steps, save_every))) return rc;
  if (!x0 || !bc || !out) { set_error("null buffer"); return OPZ_EINVAL; }
  if (B == 0) return OPZ_OK;
  OPZ_CUDA(cudaSetDevice(ctx->device));
  RolloutCall c{};
  c.P = make_dev_params(*p);
  c.x0 = x0; c.bc = bc; c.out = out; c.B = B; c.scheme = scheme; c.dt = dt_nd; c.t0 = t0;
  c.n_steps = n_steps; c.save_every = save_every; c.n_save = n_save_of(n_steps, save_every);
  if (precision == OPZ_PREC_FP32) return fp32_rollout(ctx, m, c);
  if (scheme == OPZ_IMEX_EULER) { set_error("OPZ_IMEX_EULER is built for OPZ_PREC_FP32 only"); return OPZ_EUNSUP; }
  if (m->nd.n_layers < 2) { set_error("tensor-core paths: a single Dense layer has no hidden GEMM; use OPZ_PREC_FP32"); return OPZ_EUNSUP; }
  opz_model* mm = const_cast<opz_model*>(m);  // the packed weight images are caches of w_dev
  // FP16 / BF16 operands: the fused persistent kernel (opz_tc2.cu) for the shapes it is built for — Nz = 32, hidden <= 400 —
  // the layered path (opz_gemm.cu) for everything else (Nz = 128, wider / deeper nets); OPZ_FORCE_LAYERED=1 selects the
  // layered path regardless (cross-checks of the two in the tests)
  if ((precision == OPZ_PREC_BF16 || precision == OPZ_PREC_FP16) && tc2_supports(m) && !getenv("OPZ_FORCE_LAYERED")) {
    if (!m->tc_pack_valid || m->tc_pack_precision != precision) {
      if ((rc = tc2_model_prepare(mm, precision))) return rc;
      mm->tc_pack_valid = true;
      mm->tc_pack_precision = precision;
    }
    return tc2_rollout(ctx, m, c, precision);
  }
  // split-operand precisions (near-FP32 on the tensor cores) and all other shapes: one batched GEMM per Dense layer
  if (precision == OPZ_PREC_BF16 || precision == OPZ_PREC_FP16 || precision == OPZ_PREC_BF16X2 || precision == OPZ_PREC_BF16X3 ||
      precision == OPZ_PREC_FP16X2 || precision == OPZ_PREC_FP16X3) {
    if ((!m->gm_valid || m->gm_precision != precision) && (rc = gemm_model_prepare(mm, precision))) return rc;
    return gemm_rollout(ctx, m, c, precision);
  }
  set_error("unknown precision");
  return OPZ_EINVAL;
}

int opz_rollout_forward(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0,
                        const float* bc, int64_t B, int scheme, float dt_nd, float t0, int n_steps,
                        int save_every, int precision, float* out) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if ((rc = check_rollout(B, scheme, n_steps, save_every))) return rc;
  if (!x0 || !bc || !out) { set_error("null buffer"); return OPZ_EINVAL; }
  if (B == 0) return OPZ_OK;
  OPZ_CUDA(cudaSetDevice(ctx->device));
  const size_t S = 3 * (size_t)m->nz;
  const size_t n_save = n_save_of(n_steps, save_every);
  float *dx, *dbc, *dout;
  if ((rc = ensure_scratch(ctx, 0, sizeof(float) * B * S, (void**)&dx))) return rc;
  if ((rc = ensure_scratch(ctx, 1, sizeof(float) * B * OPZ_BC_STRIDE, (void**)&dbc))) return rc;
  if ((rc = ensure_scratch(ctx, 2, sizeof(float) * B * n_save * S, (void**)&dout))) return rc;
  // Large ensembles are pipelined in slabs of whole waves of column tiles: slab i + 1 is copied in and slab i - 1 copied out
  // while slab i is stepped (columns are independent, so the slab boundaries do not change a single bit of the result).
  // With pageable host memory the copies degrade to staged synchronous ones; the result is the same.
  const int64_t wave = (int64_t)ctx->sm_count * 128;          // one 128-column tile per SM
  int64_t slab = wave * 8;
  int n_slabs = (int)((B + slab - 1) / slab);
  if (n_slabs > 16) { n_slabs = 16; slab = ((B + n_slabs - 1) / n_slabs + wave - 1) / wave * wave; n_slabs = (int)((B + slab - 1) / slab); }
  if (n_slabs < 3) {
    OPZ_CUDA(cudaMemcpyAsync(dx, x0, sizeof(float) * B * S, cudaMemcpyHostToDevice, ctx->stream));
    OPZ_CUDA(cudaMemcpyAsync(dbc, bc, sizeof(float) * B * OPZ_BC_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = opz_rollout_forward_dev(ctx, m, p, dx, dbc, B, scheme, dt_nd, t0, n_steps, save_every, precision, dout)))
      return rc;
    OPZ_CUDA(cudaMemcpyAsync(out, dout, sizeof(float) * B * n_save * S, cudaMemcpyDeviceToHost, ctx->stream));
    OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return OPZ_OK;
  }
  if (!ctx->h2d_stream) OPZ_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) OPZ_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 1 + 2 * n_slabs; ++i)
    if (!ctx->slab_ev[i]) OPZ_CUDA(cudaEventCreateWithFlags(&ctx->slab_ev[i], cudaEventDisableTiming));
  // work queued on the context's stream before this call comes first
  OPZ_CUDA(cudaEventRecord(ctx->slab_ev[0], ctx->stream));
  OPZ_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, ctx->slab_ev[0], 0));
  OPZ_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, ctx->slab_ev[0], 0));
  int rc_slab = OPZ_OK;
  cudaError_t ce = cudaSuccess;
  const char* ce_what = "";
  // no early return inside the loop: whatever has been enqueued must drain before the caller gets its buffers back
#define OPZ_SLAB(expr) do { if (ce == cudaSuccess && rc_slab == OPZ_OK) { ce = (expr); if (ce != cudaSuccess) ce_what = #expr; } } while (0)
  for (int i = 0; i < n_slabs && ce == cudaSuccess && rc_slab == OPZ_OK; ++i) {
    const int64_t c0 = (int64_t)i * slab, nb = std::min<int64_t>(slab, B - c0);
    cudaEvent_t ev_in = ctx->slab_ev[1 + 2 * i], ev_done = ctx->slab_ev[2 + 2 * i];
    OPZ_SLAB(cudaMemcpyAsync(dx + c0 * S, x0 + c0 * S, sizeof(float) * nb * S, cudaMemcpyHostToDevice, ctx->h2d_stream));
    OPZ_SLAB(cudaMemcpyAsync(dbc + c0 * OPZ_BC_STRIDE, bc + c0 * OPZ_BC_STRIDE, sizeof(float) * nb * OPZ_BC_STRIDE,
                             cudaMemcpyHostToDevice, ctx->h2d_stream));
    OPZ_SLAB(cudaEventRecord(ev_in, ctx->h2d_stream));
    OPZ_SLAB(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
    if (ce == cudaSuccess)
      rc_slab = opz_rollout_forward_dev(ctx, m, p, dx + c0 * S, dbc + c0 * OPZ_BC_STRIDE, nb, scheme, dt_nd, t0, n_steps, save_every,
                                        precision, dout + c0 * n_save * S);
    OPZ_SLAB(cudaEventRecord(ev_done, ctx->stream));
    OPZ_SLAB(cudaStreamWaitEvent(ctx->d2h_stream, ev_done, 0));
    OPZ_SLAB(cudaMemcpyAsync(out + c0 * n_save * S, dout + c0 * n_save * S, sizeof(float) * nb * n_save * S, cudaMemcpyDeviceToHost,
                             ctx->d2h_stream));
  }
#undef OPZ_SLAB
  // blocking call: nothing of it is in flight when it returns (also after an error in the middle)
  cudaError_t e1 = cudaStreamSynchronize(ctx->h2d_stream), e2 = cudaStreamSynchronize(ctx->stream), e3 = cudaStreamSynchronize(ctx->d2h_stream);
  if (rc_slab) return rc_slab;
  if (ce != cudaSuccess) return cuda_fail(ce, ce_what);
  OPZ_CUDA(e1);
  OPZ_CUDA(e2);
  OPZ_CUDA(e3);
  return OPZ_OK;
}

// ---- loss + gradient ----------------------------------------------------------------------------
static int loss_grad_dev_impl(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0, const float* bc,
                              int64_t B, int64_t B_total, int scheme, float dt_nd, float t0, int n_steps,
                              int save_every, const float* target, const float* loss_w, int precision,
                              float* loss_out, float* grad_out, float* gphys_out) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if ((rc = check_rollout(B, scheme, n_steps, save_every))) return rc;
  // a rank that owns no columns (B == 0) still calls in and gets zeros for the all-reduce
  if ((B > 0 && (!x0 || !bc || !target)) || !loss_w || !loss_out || !grad_out) { set_error("null buffer"); return OPZ_EINVAL; }
  if (save_every <= 0) { set_error("opz_loss_grad needs save_every > 0"); return OPZ_EINVAL; }
  if (B_total < B || B_total < 1) { set_error("B_total must be >= B and >= 1"); return OPZ_EINVAL; }
  if (precision != OPZ_PREC_FP32) { set_error("opz_loss_grad: only OPZ_PREC_FP32 so far"); return OPZ_EUNSUP; }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  LossGradCall c{};
  c.r.P = make_dev_params(*p);
  c.r.x0 = x0; c.r.bc = bc; c.r.out = nullptr; c.r.B = B; c.r.scheme = scheme; c.r.dt = dt_nd; c.r.t0 = t0;
  c.r.n_steps = n_steps; c.r.save_every = save_every; c.r.n_save = n_save_of(n_steps, save_every);
  c.target = target;
  for (int i = 0; i < 6; ++i) c.loss_w[i] = loss_w[i];
  c.B_total = B_total;
  c.loss_out = loss_out;
  c.grad_out = grad_out;
  c.gphys_out = gphys_out;
  if ((rc = fp32_loss_grad(ctx, m, c))) return rc;
  // the one exchange of the training step: sum of the shard gradients and loss components over the ranks
  float* bufs[3] = {grad_out, loss_out, gphys_out};
  const size_t counts[3] = {3 * (size_t)m->nd.P, 7, gphys_out ? 5u : 0u};
  return comm_allreduce_sum(ctx, bufs, counts, 3);
}

int opz_loss_grad_dev(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0, const float* bc,
                      int64_t B, int64_t B_total, int scheme, float dt_nd, float t0, int n_steps,
                      int save_every, const float* target, const float* loss_w, int precision,
                      float* loss_out, float* grad_out) {
  return loss_grad_dev_impl(ctx, m, p, x0, bc, B, B_total, scheme, dt_nd, t0, n_steps, save_every, target, loss_w, precision,
                            loss_out, grad_out, nullptr);
}

int opz_loss_grad_mpp_dev(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0, const float* bc,
                          int64_t B, int64_t B_total, int scheme, float dt_nd, float t0, int n_steps,
                          int save_every, const float* target, const float* loss_w, int precision,
                          float* loss_out, float* grad_out, float* grad_mpp_out) {
  if (!grad_mpp_out) { set_error("null buffer"); return OPZ_EINVAL; }
  if (p && !(p->flags & OPZ_FLAG_MPP)) { set_error("opz_loss_grad_mpp: OPZ_FLAG_MPP is off, the parameters have no effect"); return OPZ_EINVAL; }
  return loss_grad_dev_impl(ctx, m, p, x0, bc, B, B_total, scheme, dt_nd, t0, n_steps, save_every, target, loss_w, precision,
                            loss_out, grad_out, grad_mpp_out);
}

static int loss_grad_host_impl(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0, const float* bc,
                               int64_t B, int64_t B_total, int scheme, float dt_nd, float t0, int n_steps, int save_every,
                               const float* target, const float* loss_w, int precision, float* loss_out, float* grad_out,
                               float* grad_mpp_out) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if ((rc = check_rollout(B, scheme, n_steps, save_every))) return rc;
  // a rank that owns no case (B == 0) may pass null data pointers, exactly as in the _dev variant
  if ((B > 0 && (!x0 || !bc || !target)) || !loss_w || !loss_out || !grad_out) { set_error("null buffer"); return OPZ_EINVAL; }
  if (save_every <= 0) { set_error("opz_loss_grad needs save_every > 0"); return OPZ_EINVAL; }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  const size_t S = 3 * (size_t)m->nz;
  const size_t n_save = n_save_of(n_steps, save_every);
  const size_t P3 = 3 * (size_t)m->nd.P;
  float *dx, *dbc, *dtgt, *dloss, *dgrad;
  if ((rc = ensure_scratch(ctx, 0, sizeof(float) * (B ? B : 1) * S, (void**)&dx))) return rc;
  if ((rc = ensure_scratch(ctx, 1, sizeof(float) * (B ? B : 1) * OPZ_BC_STRIDE, (void**)&dbc))) return rc;
  if ((rc = ensure_scratch(ctx, 2, sizeof(float) * (B ? B : 1) * n_save * S, (void**)&dtgt))) return rc;
  if ((rc = ensure_scratch(ctx, 3, sizeof(float) * (P3 + 16), (void**)&dgrad))) return rc;
  dloss = dgrad + P3;
  float* dmpp = grad_mpp_out ? dgrad + P3 + 8 : nullptr;
  if (B > 0) {
    OPZ_CUDA(cudaMemcpyAsync(dx, x0, sizeof(float) * B * S, cudaMemcpyHostToDevice, ctx->stream));
    OPZ_CUDA(cudaMemcpyAsync(dbc, bc, sizeof(float) * B * OPZ_BC_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
    OPZ_CUDA(cudaMemcpyAsync(dtgt, target, sizeof(float) * B * n_save * S, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (grad_mpp_out && !(p->flags & OPZ_FLAG_MPP)) { set_error("opz_loss_grad_mpp: OPZ_FLAG_MPP is off, the parameters have no effect"); return OPZ_EINVAL; }
  if ((rc = loss_grad_dev_impl(ctx, m, p, dx, dbc, B, B_total, scheme, dt_nd, t0, n_steps, save_every, dtgt, loss_w,
                               precision, dloss, dgrad, dmpp)))
    return rc;
  OPZ_CUDA(cudaMemcpyAsync(grad_out, dgrad, sizeof(float) * P3, cudaMemcpyDeviceToHost, ctx->stream));
  OPZ_CUDA(cudaMemcpyAsync(loss_out, dloss, sizeof(float) * 7, cudaMemcpyDeviceToHost, ctx->stream));
  if (grad_mpp_out) OPZ_CUDA(cudaMemcpyAsync(grad_mpp_out, dmpp, sizeof(float) * 5, cudaMemcpyDeviceToHost, ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPZ_OK;
}

int opz_loss_grad(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0, const float* bc,
                  int64_t B, int64_t B_total, int scheme, float dt_nd, float t0, int n_steps, int save_every,
                  const float* target, const float* loss_w, int precision, float* loss_out, float* grad_out) {
  return loss_grad_host_impl(ctx, m, p, x0, bc, B, B_total, scheme, dt_nd, t0, n_steps, save_every, target, loss_w, precision,
                             loss_out, grad_out, nullptr);
}

int opz_loss_grad_mpp(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0, const float* bc,
                      int64_t B, int64_t B_total, int scheme, float dt_nd, float t0, int n_steps, int save_every,
                      const float* target, const float* loss_w, int precision, float* loss_out, float* grad_out,
                      float* grad_mpp_out) {
  if (!grad_mpp_out) { set_error("null buffer"); return OPZ_EINVAL; }
  return loss_grad_host_impl(ctx, m, p, x0, bc, B, B_total, scheme, dt_nd, t0, n_steps, save_every, target, loss_w, precision,
                             loss_out, grad_out, grad_mpp_out);
}

// ---- supervised flux pre-training of one net --------------------------------------------------------
int opz_flux_loss_grad(opz_ctx* ctx, const opz_model* m, const opz_params* p, int which_net, const float* x, const float* bc,
                       const float* flux_target, int64_t N, float gradient_scaling, float* loss_out, float* grad_out) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if (which_net < 0 || which_net > 2) { set_error("which_net must be 0 (uw), 1 (vw) or 2 (wT)"); return OPZ_EINVAL; }
  if (N < 0 || (N > 0 && (!x || !bc || !flux_target)) || !loss_out || !grad_out) { set_error("null buffer or negative N"); return OPZ_EINVAL; }
  if (m->nd.n_layers < 2) { set_error("opz_flux_loss_grad: the nets need at least one hidden layer"); return OPZ_EUNSUP; }
  OPZ_CUDA(cudaSetDevice(ctx->device));
  const size_t S = 3 * (size_t)m->nz, nf = (size_t)m->nz + 1, P1 = (size_t)m->nd.P;
  float *dx, *dbc, *dt, *dg;
  if ((rc = ensure_scratch(ctx, 0, sizeof(float) * (N ? N : 1) * S, (void**)&dx))) return rc;
  if ((rc = ensure_scratch(ctx, 1, sizeof(float) * (N ? N : 1) * OPZ_BC_STRIDE, (void**)&dbc))) return rc;
  if ((rc = ensure_scratch(ctx, 2, sizeof(float) * (N ? N : 1) * nf, (void**)&dt))) return rc;
  if ((rc = ensure_scratch(ctx, 3, sizeof(float) * P1, (void**)&dg))) return rc;
  if (N > 0) {
    OPZ_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * N * S, cudaMemcpyHostToDevice, ctx->stream));
    OPZ_CUDA(cudaMemcpyAsync(dbc, bc, sizeof(float) * N * OPZ_BC_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
    OPZ_CUDA(cudaMemcpyAsync(dt, flux_target, sizeof(float) * N * nf, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = fp32_flux_loss_grad(ctx, m, make_dev_params(*p), which_net, dx, dbc, dt, N, gradient_scaling, loss_out, dg))) return rc;
  OPZ_CUDA(cudaMemcpyAsync(grad_out, dg, sizeof(float) * P1, cudaMemcpyDeviceToHost, ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPZ_OK;
}

// ---- post-rollout diagnostics -------------------------------------------------------------------
int opz_profile_diagnostics(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* sol, const float* bc,
                            int64_t B, int n_save, float t0, float dt_save, const float* target, float* fluxes,
                            float* fluxes_mpp, float* ri, float* loss_t) {
  int rc = check_common(ctx, m, p);
  if (rc) return rc;
  if (!sol || !bc || B < 0 || n_save < 1) { set_error("null buffer, negative B or n_save < 1"); return OPZ_EINVAL; }
  if (loss_t && !target) { set_error("loss_t needs a target"); return OPZ_EINVAL; }
  if (B == 0) return OPZ_OK;
  OPZ_CUDA(cudaSetDevice(ctx->device));
  const size_t nz = (size_t)m->nz, S = 3 * nz, nf = 3 * (nz + 1), rows = (size_t)B * n_save;
  float *dsol, *dbc, *dtgt = nullptr, *dout;
  if ((rc = ensure_scratch(ctx, 0, sizeof(float) * rows * S, (void**)&dsol))) return rc;
  if ((rc = ensure_scratch(ctx, 1, sizeof(float) * B * OPZ_BC_STRIDE, (void**)&dbc))) return rc;
  if (target && (rc = ensure_scratch(ctx, 2, sizeof(float) * rows * S, (void**)&dtgt))) return rc;
  if ((rc = ensure_scratch(ctx, 3, sizeof(float) * rows * (2 * nf + (nz + 1) + 3), (void**)&dout))) return rc;
  float* dfl = dout;
  float* dfm = dfl + rows * nf;
  float* dri = dfm + rows * nf;
  float* dlt = dri + rows * (nz + 1);
  OPZ_CUDA(cudaMemcpyAsync(dsol, sol, sizeof(float) * rows * S, cudaMemcpyHostToDevice, ctx->stream));
  OPZ_CUDA(cudaMemcpyAsync(dbc, bc, sizeof(float) * B * OPZ_BC_STRIDE, cudaMemcpyHostToDevice, ctx->stream));
  if (target) OPZ_CUDA(cudaMemcpyAsync(dtgt, target, sizeof(float) * rows * S, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = fp32_diagnostics(ctx, m, make_dev_params(*p), dsol, dbc, B, n_save, t0, dt_save, dtgt, fluxes ? dfl : nullptr,
                             fluxes_mpp ? dfm : nullptr, ri ? dri : nullptr, loss_t ? dlt : nullptr)))
    return rc;
  if (fluxes) OPZ_CUDA(cudaMemcpyAsync(fluxes, dfl, sizeof(float) * rows * nf, cudaMemcpyDeviceToHost, ctx->stream));
  if (fluxes_mpp) OPZ_CUDA(cudaMemcpyAsync(fluxes_mpp, dfm, sizeof(float) * rows * nf, cudaMemcpyDeviceToHost, ctx->stream));
  if (ri) OPZ_CUDA(cudaMemcpyAsync(ri, dri, sizeof(float) * rows * (nz + 1), cudaMemcpyDeviceToHost, ctx->stream));
  if (loss_t) OPZ_CUDA(cudaMemcpyAsync(loss_t, dlt, sizeof(float) * rows * 3, cudaMemcpyDeviceToHost, ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPZ_OK;
}

// ---- batched MLP --------------------------------------------------------------------------------
int opz_mlp_forward(opz_ctx* ctx, const opz_model* m, int which_net, const float* x, int64_t Nt, int precision,
                    float* y) {
  if (!ctx || !m || !x || !y) { set_error("null argument"); return OPZ_EINVAL; }
  if (m->ctx != ctx) { set_error("model belongs to another context"); return OPZ_EINVAL; }
  if (which_net < 0 || which_net > 2 || Nt < 0) { set_error("which_net must be 0..2, Nt >= 0"); return OPZ_EINVAL; }
  const bool tensor = precision == OPZ_PREC_BF16 || precision == OPZ_PREC_FP16 || precision == OPZ_PREC_BF16X2 || precision == OPZ_PREC_BF16X3 ||
                      precision == OPZ_PREC_FP16X2 || precision == OPZ_PREC_FP16X3;
  if (precision != OPZ_PREC_FP32 && !tensor) { set_error("opz_mlp_forward: unknown precision"); return OPZ_EINVAL; }
  if (tensor && m->nd.n_layers < 2) { set_error("opz_mlp_forward: a single Dense layer has no hidden GEMM; use OPZ_PREC_FP32"); return OPZ_EUNSUP; }
  if (Nt == 0) return OPZ_OK;
  OPZ_CUDA(cudaSetDevice(ctx->device));
  const size_t S = 3 * (size_t)m->nz, NO = (size_t)m->nz - 1;
  float *dx, *dy;
  int rc;
  if ((rc = ensure_scratch(ctx, 0, sizeof(float) * Nt * S, (void**)&dx))) return rc;
  if ((rc = ensure_scratch(ctx, 2, sizeof(float) * Nt * NO, (void**)&dy))) return rc;
  OPZ_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * Nt * S, cudaMemcpyHostToDevice, ctx->stream));
  if (tensor) {   // the layered tensor path (opz_gemm.cu)
    opz_model* mm = const_cast<opz_model*>(m);
    if ((!m->gm_valid || m->gm_precision != precision) && (rc = gemm_model_prepare(mm, precision))) return rc;
    if ((rc = gemm_mlp(ctx, m, which_net, dx, Nt, precision, dy))) return rc;
  } else if ((rc = fp32_mlp(ctx, m, which_net, dx, Nt, dy))) return rc;
  OPZ_CUDA(cudaMemcpyAsync(y, dy, sizeof(float) * Nt * NO, cudaMemcpyDeviceToHost, ctx->stream));
  OPZ_CUDA(cudaStreamSynchronize(ctx->stream));
  return OPZ_OK;
}

}  // extern "C"
