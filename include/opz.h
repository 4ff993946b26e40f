/* opz.h — C ABI of libopz.so: the B200-native NDE column-model hot path.
 *
 * Drop-in boundary for ONE path of CliMA/OceanParameterizations.jl: the neural differential
 * equation (NDE) rollout of the wind_mixing column model and the gradient of its profile loss.
 * The reference has no FFI layer (it is 100 % Julia); each entry point below replaces the Julia
 * function named in its comment and is what a `ccall` shim binds (see INTEGRATION.md and
 * julia/OPZ.jl).  Conventions:
 *   - every call returns 0 on success, a negative OPZ_E* code on failure; the message is
 *     available from opz_last_error() (thread-local);  nothing throws across the ABI;
 *   - the caller owns all host buffers; Julia column-major `3Nz x Nt x B` is C `[B][Nt][3Nz]`;
 *   - host-pointer calls are blocking (stream-synchronised before return);  `_dev` calls take
 *     DEVICE pointers, enqueue on the context's stream and return without synchronising;
 *   - one opz_ctx drives one GPU (one process or host thread per GPU); there is NO CPU fallback:
 *     creating a context on a machine without an sm_100 device fails with OPZ_ENODEV;
 *   - forward ensembles shard by column with no communication; the training step's one exchange (the
 *     sum of the NN-parameter gradient over the case shards) happens INSIDE opz_loss_grad* once the
 *     context carries a communicator (opz_ctx_comm_init / opz_ctx_comm_adopt below).
 */
#ifndef OPZ_H
#define OPZ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPZ_VERSION 100 /* 0.1.0 */

/* error codes */
#define OPZ_OK 0
#define OPZ_EINVAL (-1)  /* bad argument */
#define OPZ_ENODEV (-2)  /* no usable sm_100 device */
#define OPZ_ECUDA (-3)   /* CUDA runtime error (message has the detail) */
#define OPZ_ENOMEM (-4)  /* device allocation failed */
#define OPZ_EUNSUP (-5)  /* configuration not supported by the requested precision path */
#define OPZ_ENCCL (-6)   /* NCCL could not be loaded, or a NCCL call failed (message has the detail) */

/* conditions NamedTuple -> flag bits (wind_mixing/src/NDE_training.jl:205-207) */
#define OPZ_FLAG_MPP (1u << 0)                  /* modified_pacanowski_philander */
#define OPZ_FLAG_CA (1u << 1)                   /* convective_adjustment (training_postprocessing.jl:118-121) */
#define OPZ_FLAG_CA_SELECT_DUDZ (1u << 2)       /* reference quirk: predicate on du/dz (tp.jl:120); default dT/dz */
#define OPZ_FLAG_BC_MINUS_SCALE0 (1u << 3)      /* boundary flux = bc - scale(0) (tp.jl:82-93; zero_weights) */
#define OPZ_FLAG_DIURNAL (1u << 4)              /* wT_top(t) = scale_wT(Q sin(2 pi t tau/period)/(alpha g)) */
#define OPZ_FLAG_DIURNAL_MINUS_SCALE0 (1u << 5) /* consistent offset; the reference omits it (tp.jl:143) */
#define OPZ_FLAG_RI_EPS (1u << 6)               /* add eps to each gradient inside Ri (NDE_training.jl:115-119) */
#define OPZ_FLAG_SMOOTH_NN (1u << 7)            /* 3-point box filter on NN output (NDE_training.jl:98-102) */
#define OPZ_FLAG_SMOOTH_RI (1u << 8)            /* 3-point box filter on Ri (NDE_training.jl:121-123) */

/* timestepper (replaces the OrdinaryDiffEq `timestepper` argument; fixed step) */
#define OPZ_EULER 0
#define OPZ_RK2 1 /* Heun */
#define OPZ_RK4 2
/* Split step of the reference's production path (NDE_oceananigans.jl:61-101): explicit Euler for the NN fluxes, boundary fluxes
 * and Coriolis, then backward-Euler vertical diffusion (tridiagonal solve per variable) with the diffusivities of the updated
 * state; lifts the explicit limit dt <= ~0.7 dz^2 / kappa of convective adjustment.  opz_rollout_forward[_dev], OPZ_PREC_FP32. */
#define OPZ_IMEX_EULER 3

/* activation of the hidden Dense layers (NNlib names); the last layer is linear */
#define OPZ_ACT_IDENTITY 0
#define OPZ_ACT_RELU 1
#define OPZ_ACT_TANH 2
#define OPZ_ACT_MISH 3
#define OPZ_ACT_SWISH 4
#define OPZ_ACT_LEAKYRELU 5

/* arithmetic of the batched MLP (physics and time stepping are always FP32) */
#define OPZ_PREC_FP32 0   /* FP32 FMA on CUDA cores: the parity path (<= 1e-4) */
#define OPZ_PREC_BF16 1   /* tcgen05 kind::f16, BF16 operands, FP32 accumulate in TMEM */
#define OPZ_PREC_BF16X2 2 /* as BF16 with the weights split hi + lo (2 MMAs per k-step) */
#define OPZ_PREC_BF16X3 3 /* activations and weights split hi + lo (hi*hi + hi*lo + lo*hi, 3 MMAs): 16 mantissa bits per operand */
#define OPZ_PREC_FP16 4   /* tcgen05 kind::f16, IEEE-half operands (saturating), FP32 accumulate in TMEM:
                             same speed as BF16, 8x finer operand rounding; the default reduced-precision path */
#define OPZ_PREC_FP16X2 5 /* IEEE-half operands with the weights split hi + lo */
#define OPZ_PREC_FP16X3 6 /* IEEE-half operands, both split (3 MMAs): ~21 mantissa bits per operand, FP32 accumulation —
                             the near-FP32 tensor-core mode (inside the 1e-4 contract of the FP32 path) */
/* OPZ_PREC_FP16 / OPZ_PREC_BF16 run the fused persistent kernel for Nz = 32 and hidden widths <= 400, and the layered
 * path (one batched tcgen05 GEMM per Dense layer) for every other shape; the split precisions always run the layered path. */

#define OPZ_BC_STRIDE 8 /* floats per column in `bc` */

/* constants + scalings of prepare_parameters_NDE_training (NDE_training.jl:1-44), flattened.
 * mu/sigma order: u, v, T, uw, vw, wT  (six ZeroMeanUnitVarianceScaling objects,
 * src/DataWrangling/feature_scaling.jl:7-23). 37 x 4 bytes, no padding. */
typedef struct opz_params {
  int32_t nz;    /* cells per column (32) */
  float H;       /* abs(z_F[end]-z_F[1]) [m] */
  float tau;     /* abs(t[end]-t[1]) [s]; model time is t/tau */
  float f;       /* Coriolis [1/s] */
  float g;       /* gravity */
  float alpha;   /* thermal expansion */
  float nu0;     /* background diffusivity */
  float nu_m;    /* nu_- : shear-instability increment */
  float Ric;     /* critical Richardson number */
  float dRi;     /* width of the tanh switch */
  float Pr;      /* Prandtl number */
  float kappa;   /* convective-adjustment diffusivity */
  float eps;     /* used when OPZ_FLAG_RI_EPS */
  float mu[6];
  float sigma[6];
  uint32_t flags;
  float diurnal_period; /* seconds (86400) */
  float reserved[10];
} opz_params;

typedef struct opz_ctx opz_ctx;
typedef struct opz_model opz_model;

/* ---- lifecycle / errors ------------------------------------------------------------------ */
int opz_version(void);
const char* opz_last_error(void);
int opz_ctx_create(int device, opz_ctx** out);
int opz_ctx_destroy(opz_ctx* ctx);
/* adopt an externally owned cudaStream_t (e.g. torch's current stream) for all later calls */
int opz_ctx_set_stream(opz_ctx* ctx, void* cuda_stream);
int opz_ctx_synchronize(opz_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t opz_ctx_launch_count(const opz_ctx* ctx);
/* Destroying a context that still has live models is deferred until the last of them is destroyed (host garbage
 * collectors and finalizers give no destruction order): the handle must not be used after opz_ctx_destroy either way. */

/* ---- multi-GPU: the gradient all-reduce of the training step, inside the boundary --------------------------------
 * Replaces the mean over simulations of loss_NDE / loss_gradient_NDE (NDE_training.jl:290-301) when the simulations
 * are sharded over GPUs: rank r passes ITS cases (opz_shard_range) with B_total = the global count, and
 * opz_loss_grad[_mpp][_dev] returns the global loss_out / grad_out / grad_mpp_out on every rank (NCCL all-reduce, sum,
 * FP32, enqueued on the context's stream behind the BPTT kernels; NVLink / NVSwitch between the GPUs of a node).
 * NCCL is bound at run time (libnccl.so.2, or the path in OPZ_NCCL_LIB): single-GPU use needs no NCCL.
 *   rank 0:      opz_comm_get_unique_id(id)  -> ship the 128 bytes to the other ranks by any host means
 *   every rank:  opz_ctx_comm_init(ctx, id, n_ranks, rank)        (collective: blocks until all ranks have called)
 * or hand over a communicator the host already owns (ncclComm_t):  opz_ctx_comm_adopt(ctx, comm, n_ranks, rank).
 * A rank that owns no case (B = 0) still calls opz_loss_grad* and contributes zeros. */
#define OPZ_COMM_ID_BYTES 128
int opz_comm_get_unique_id(void* id_out /* OPZ_COMM_ID_BYTES */);
int opz_ctx_comm_init(opz_ctx* ctx, const void* id, int n_ranks, int rank);
int opz_ctx_comm_adopt(opz_ctx* ctx, void* nccl_comm, int n_ranks, int rank);
int opz_ctx_comm_destroy(opz_ctx* ctx);
/* any output may be NULL; collectives = all-reduce groups enqueued so far */
int opz_ctx_comm_info(const opz_ctx* ctx, int* n_ranks, int* rank, int64_t* collectives);
int opz_comm_nccl_version(int* version);
/* contiguous shard [lo, hi) of B cases / columns owned by `rank`: the first B % n_ranks ranks own one more */
int opz_shard_range(int64_t B, int rank, int n_ranks, int64_t* lo, int64_t* hi);

/* ---- model: three Flux Chains uw_NN, vw_NN, wT_NN of one architecture ----------------------
 * Replaces `Flux.destructure(NN)` / `re(weights)` (NDE_training.jl:11-13,62-64;
 * wind_mixing/construct_NN.jl:29-34).  sizes = {in, hidden..., out}, n_sizes = #layers + 1;
 * in must be 3*nz and out nz-1.  Each w_* is one net's flat vector in Flux.destructure order
 * [vec(W1) (out x in, column-major); b1; vec(W2); b2; ...] (HOST pointers). */
int opz_model_create(opz_ctx* ctx, int nz, int n_sizes, const int* sizes, int act,
                     const float* w_uw, const float* w_vw, const float* w_wT, opz_model** out);
/* weights = [uw; vw; wT] flat, 3P floats (the `weights` vector train_NDE optimises) */
int opz_model_set_weights(opz_model* m, const float* weights_host);
int opz_model_get_weights(const opz_model* m, float* weights_host);
int opz_model_set_weights_dev(opz_model* m, const float* weights_dev);
int64_t opz_model_n_params(const opz_model* m); /* P per net */
int opz_model_destroy(opz_model* m);

/* ---- single right-hand-side evaluation ---------------------------------------------------
 * Replaces NDE(x,p,t,...) / predict_NDE / predict_flux (NDE_training.jl:56-165) and NDE! /
 * predict_flux! (training_postprocessing.jl:105-153) for B columns at once.
 * x[B][3nz] scaled state; bc[B][8] = uw_bottom, uw_top, vw_bottom, vw_top, wT_bottom, wT_top
 * (scaled, NDE_training.jl:238-243), Q_diurnal [m^2 s^-3], unused; t = time / tau.
 * dxdt[B][3nz]; fluxes (optional, may be NULL) [B][3][nz+1] = total scaled uw, vw, wT. */
int opz_rhs(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x, const float* bc,
            float t, int64_t B, int precision, float* dxdt, float* fluxes);

/* ---- forward rollout ------------------------------------------------------------------------
 * Replaces solve_NDE_mutating(uw_NN, vw_NN, wT_NN, scalings, constants, BCs, derivatives, uvT0,
 * ts, timestepper, conditions) (training_postprocessing.jl:55-159) for an ensemble of B
 * columns: `timestepper` -> (scheme, dt_nd); ts -> t0 + i*save_every*dt_nd, i = 0..n_steps/save_every.
 * out[B][n_save][3nz], n_save = n_steps/save_every + 1, snapshot 0 = x0 (as `saveat=ts`).
 * save_every <= 0 saves the final state only: out[B][1][3nz]. */
int opz_rollout_forward(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0,
                        const float* bc, int64_t B, int scheme, float dt_nd, float t0,
                        int n_steps, int save_every, int precision, float* out);
int opz_rollout_forward_dev(opz_ctx* ctx, const opz_model* m, const opz_params* p,
                            const float* x0_dev, const float* bc_dev, int64_t B, int scheme,
                            float dt_nd, float t0, int n_steps, int save_every, int precision,
                            float* out_dev);

/* ---- loss and its gradient w.r.t. the NN weights ----------------------------------------------
 * Replaces loss_NDE / loss_gradient_NDE + Zygote.gradient through
 * solve(...; sensealg=InterpolatingAdjoint(autojacvec=ZygoteVJP())) (NDE_training.jl:290-333)
 * by the exact discrete adjoint (BPTT) of the fixed-step scheme.
 * target[B][n_save][3nz]; loss_w[6] = weights of (u, v, T, du/dz, dv/dz, dT/dz) MSE terms
 * (loss.jl:1-42); loss_out[7] = total then the six weighted components; grad_out[3P].
 * The sums are over THIS context's B columns with the mean taken over `B_total` simulations
 * (pass B_total = B on one GPU).  With a communicator on the context (opz_ctx_comm_init) the outputs are
 * all-reduced (sum) inside the call and every rank receives the global loss and gradient; without
 * one, multi-GPU callers all-reduce grad_out and loss_out themselves. */
int opz_loss_grad(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0,
                  const float* bc, int64_t B, int64_t B_total, int scheme, float dt_nd, float t0,
                  int n_steps, int save_every, const float* target, const float* loss_w,
                  int precision, float* loss_out, float* grad_out);
int opz_loss_grad_dev(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0_dev,
                      const float* bc_dev, int64_t B, int64_t B_total, int scheme, float dt_nd,
                      float t0, int n_steps, int save_every, const float* target_dev,
                      const float* loss_w_host, int precision, float* loss_out_dev,
                      float* grad_out_dev);

/* ---- the same, plus the gradient w.r.t. the five modified Pacanowski-Philander parameters -------
 * Replaces the Zygote gradient inside optimise_modified_pacanowski_philander / DE
 * (wind_mixing/src/diffusivity_parameter_optimisation.jl:1-33 the right-hand side, :112-200 the
 * loss and its optimisation): grad_mpp_out[5] = d loss / d (nu0, nu_m, dRi, Ric, Pr), the
 * reference's parameter order (:2), UNSCALED (the reference optimises p / p_initial, :44-48,70;
 * that factor stays on the host).  The reference's DE is this library's right-hand side with a
 * zero-weight model and flags OPZ_FLAG_MPP | OPZ_FLAG_RI_EPS | OPZ_FLAG_BC_MINUS_SCALE0; with a
 * non-zero model the NN fluxes are simply part of the differentiated solve.  Needs OPZ_FLAG_MPP.
 * Obtained in the same backward sweep as grad_out and all-reduced with it. */
int opz_loss_grad_mpp(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* x0,
                      const float* bc, int64_t B, int64_t B_total, int scheme, float dt_nd,
                      float t0, int n_steps, int save_every, const float* target,
                      const float* loss_w, int precision, float* loss_out, float* grad_out,
                      float* grad_mpp_out);
int opz_loss_grad_mpp_dev(opz_ctx* ctx, const opz_model* m, const opz_params* p,
                          const float* x0_dev, const float* bc_dev, int64_t B, int64_t B_total,
                          int scheme, float dt_nd, float t0, int n_steps, int save_every,
                          const float* target_dev, const float* loss_w_host, int precision,
                          float* loss_out_dev, float* grad_out_dev, float* grad_mpp_out_dev);

/* ---- supervised flux pre-training of one net ---------------------------------------
 * Replaces NN_loss / total_loss + Zygote inside train_NN (wind_mixing/src/NN_training.jl:207-249) with predict_uw /
 * predict_vw / predict_wT (:25-169) as the model: for sample i (one snapshot: profile x[i][3nz], boundary fluxes bc[i][8])
 *     loss_i = mse(F_pred, F_data) + gradient_scaling * mse(D^c F_data, D^c F_pred)          (:224-228)
 * where F_pred is the total flux of net `which_net` (0 uw, 1 vw, 2 wT) on the nz+1 faces — the fluxes output of opz_rhs —
 * and flux_target[N][nz+1] the scaled data.  loss_out[3] = mean over the N samples of (total, first term, second term);
 * grad_out[P] = gradient of the total w.r.t. that net's weights (Flux.destructure order).  The reference takes one
 * optimiser step per sample (Flux.train!); pass N = 1 for that, or a batch for the mean gradient. */
int opz_flux_loss_grad(opz_ctx* ctx, const opz_model* m, const opz_params* p, int which_net,
                       const float* x, const float* bc, const float* flux_target, int64_t N,
                       float gradient_scaling, float* loss_out, float* grad_out);

/* ---- diagnostics of a finished rollout --------------------------------------------
 * Replaces the per-snapshot loops of NDE_profile (wind_mixing/src/training_postprocessing.jl:250-632):
 *   fluxes     [B][n_save][3][nz+1]  total scaled fluxes uw, vw, wT = predict_flux at every saved state (:408-419), the
 *                                    diurnal top flux re-evaluated at each snapshot's time t0 + k*dt_save;
 *   fluxes_mpp [B][n_save][3][nz+1]  the same with the NN outputs taken as zero: boundary + diffusive part (:432-450);
 *                                    fluxes - fluxes_mpp on the interior faces is the NN contribution;
 *   ri         [B][n_save][nz+1]     local_richardson of D^f applied to the state (:428-431): no epsilon, so the two
 *                                    boundary faces (zero gradients) are NaN, as in the reference;
 *   loss_t     [B][n_save][3]        loss_per_tstep (loss.jl:44-46): mean squared error of u, v, T against `target`.
 * sol / target: [B][n_save][3nz] scaled states (the output of opz_rollout_forward); every output pointer may be NULL. */
int opz_profile_diagnostics(opz_ctx* ctx, const opz_model* m, const opz_params* p, const float* sol,
                            const float* bc, int64_t B, int n_save, float t0, float dt_save,
                            const float* target, float* fluxes, float* fluxes_mpp, float* ri,
                            float* loss_t);

/* ---- batched MLP forward ------------------------------------------------------------------------
 * Replaces predict(V, model; scaled=true) (src/predict.jl:12-34): one launch over Nt input
 * columns.  which_net: 0 uw, 1 vw, 2 wT.  x[Nt][3nz] -> y[Nt][nz-1]. */
int opz_mlp_forward(opz_ctx* ctx, const opz_model* m, int which_net, const float* x, int64_t Nt,
                    int precision, float* y);

#ifdef __cplusplus
}
#endif
#endif /* OPZ_H */
