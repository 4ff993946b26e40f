# OPZ.jl — thin `ccall` shim over libopz.so (include/opz.h) that keeps the reference's Julia API for the
# NDE hot path of wind_mixing/src (NDE_training.jl:56-165,290-333; training_postprocessing.jl:35-159;
# src/predict.jl:12-34).  Julia is not installed in the build image, so this file is NOT executed by
# the test-suite; the same entry points are driven through ctypes by tests/ (see INTEGRATION.md).
module OPZ

using Flux

export Context, Model, Params, Euler, Heun, RK4, IMEXEuler, solve_NDE_mutating, predict_NDE, predict_flux,
       loss_and_gradient, predict, comm_unique_id, comm_init!, shard_range

const libopz = get(ENV, "LIBOPZ", "libopz.so")

# ---- constants mirrored from include/opz.h -----------------------------------------------------------
const FLAG_MPP, FLAG_CA, FLAG_CA_SELECT_DUDZ, FLAG_BC_MINUS_SCALE0 = UInt32(1) << 0, UInt32(1) << 1, UInt32(1) << 2, UInt32(1) << 3
const FLAG_DIURNAL, FLAG_DIURNAL_MINUS_SCALE0, FLAG_RI_EPS = UInt32(1) << 4, UInt32(1) << 5, UInt32(1) << 6
const FLAG_SMOOTH_NN, FLAG_SMOOTH_RI = UInt32(1) << 7, UInt32(1) << 8
const SCHEME_EULER, SCHEME_RK2, SCHEME_RK4, SCHEME_IMEX = Cint(0), Cint(1), Cint(2), Cint(3)
const PREC_FP32, PREC_BF16, PREC_BF16X2, PREC_BF16X3, PREC_FP16, PREC_FP16X2, PREC_FP16X3 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4), Cint(5), Cint(6)
const ACT = Dict(identity => 0, relu => 1, tanh => 2, mish => 3, swish => 4, leakyrelu => 5)

struct OpzParams
    nz::Int32
    H::Float32; tau::Float32; f::Float32; g::Float32; alpha::Float32
    nu0::Float32; nu_m::Float32; Ric::Float32; dRi::Float32; Pr::Float32; kappa::Float32; eps::Float32
    mu::NTuple{6,Float32}; sigma::NTuple{6,Float32}
    flags::UInt32
    diurnal_period::Float32
    reserved::NTuple{10,Float32}
end

check(rc) = rc == 0 || error("libopz: " * unsafe_string(ccall((:opz_last_error, libopz), Cstring, ())))

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:opz_ctx_create, libopz), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(x -> ccall((:opz_ctx_destroy, libopz), Cint, (Ptr{Cvoid},), x.h), c)
    end
end
const _ctx = Ref{Union{Nothing,Context}}(nothing)
ctx() = (_ctx[] === nothing && (_ctx[] = Context(0)); _ctx[])
# Finalizers run in no particular order; that is safe: opz_ctx_destroy on a context that still has live models is
# deferred by the library until the last opz_model_destroy (include/opz.h).

# ---- multi-GPU: one Julia process per GPU; the gradient all-reduce happens inside opz_loss_grad -------------------
# rank 0:  id = comm_unique_id(); ship the 128 bytes to the other ranks (MPI.Bcast!, Distributed, a file ...)
# all:     comm_init!(ctx(), id, n_ranks, rank)   # collective;  afterwards loss_and_gradient returns GLOBAL values
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:opz_comm_get_unique_id, libopz), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(c::Context, id::Vector{UInt8}, n_ranks::Integer, rank::Integer) =
    check(ccall((:opz_ctx_comm_init, libopz), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), c.h, id, n_ranks, rank))
function shard_range(B::Integer, rank::Integer, n_ranks::Integer)   # 0-based [lo, hi) of the cases this rank owns
    lo = Ref{Int64}(0); hi = Ref{Int64}(0)
    check(ccall((:opz_shard_range, libopz), Cint, (Int64, Cint, Cint, Ref{Int64}, Ref{Int64}), B, rank, n_ranks, lo, hi))
    return lo[], hi[]
end

# ---- model: three Flux Chains -> flat Float32 vectors (Flux.destructure order) --------------------------
mutable struct Model
    h::Ptr{Cvoid}
    P::Int
    Nz::Int
    function Model(uw_NN, vw_NN, wT_NN, Nz)
        ws = [Float32.(Flux.destructure(nn)[1]) for nn in (uw_NN, vw_NN, wT_NN)]
        sizes = Cint[size(uw_NN[1].W, 2); [size(l.W, 1) for l in uw_NN]...]
        act = ACT[uw_NN[1].σ]
        r = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve ws sizes check(ccall((:opz_model_create, libopz), Cint,
            (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ref{Ptr{Cvoid}}),
            ctx().h, Nz, length(sizes), sizes, act, ws[1], ws[2], ws[3], r))
        m = new(r[], length(ws[1]), Nz)
        finalizer(x -> ccall((:opz_model_destroy, libopz), Cint, (Ptr{Cvoid},), x.h), m)
    end
end
set_weights!(m::Model, w::Vector{Float32}) = GC.@preserve w check(ccall((:opz_model_set_weights, libopz), Cint, (Ptr{Cvoid}, Ptr{Float32}), m.h, w))

# ---- parameters ------------------------------------------------------------------------------------------
function Params(constants, scalings, conditions; inference::Bool, eps=1f-7, diurnal=false)
    g(k, d=0f0) = haskey(constants, k) ? Float32(constants[k]) : Float32(d)
    flags = UInt32(0)
    if inference          # solve_NDE_mutating: mPP unconditional, BC - scale(0)  (training_postprocessing.jl:65,82-93)
        flags |= FLAG_MPP | FLAG_BC_MINUS_SCALE0
        conditions.convective_adjustment && (flags |= FLAG_CA)
    else                  # NDE / predict_flux (NDE_training.jl:83-147)
        conditions.modified_pacanowski_philander && (flags |= FLAG_MPP | FLAG_RI_EPS)
        get(conditions, :zero_weights, false) && (flags |= FLAG_BC_MINUS_SCALE0)
        get(conditions, :smooth_NN, false) && (flags |= FLAG_SMOOTH_NN)
        get(conditions, :smooth_Ri, false) && (flags |= FLAG_SMOOTH_RI)
    end
    diurnal && (flags |= FLAG_DIURNAL)
    s = (scalings.u, scalings.v, scalings.T, scalings.uw, scalings.vw, scalings.wT)
    OpzParams(constants.Nz, g(:H), g(:τ), g(:f), g(:g), g(:α), g(:ν₀), g(:ν₋), g(:Riᶜ), g(:ΔRi, 1), g(:Pr, 1), g(:κ), eps,
              ntuple(i -> Float32(s[i].μ), 6), ntuple(i -> Float32(s[i].σ), 6), flags, 86400f0, ntuple(_ -> 0f0, 10))
end

pack_BCs(BCs; Q=0f0) = Float32[BCs.uw.bottom, BCs.uw.top, BCs.vw.bottom, BCs.vw.top, BCs.wT.bottom,
                               BCs.wT.top isa Function ? 0f0 : BCs.wT.top, Q, 0f0]

# Device model for three Chains, cached per ARCHITECTURE (layer sizes, activation, Nz) — never per object identity: NDE()
# rebuilds its Chains from `p` on every call (NDE_training.jl:62-64).  The weights are uploaded on every call (<= 2.5 MB).
const _models = Dict{Any,Model}()
function cached_model(uw_NN, vw_NN, wT_NN, Nz)
    key = (Tuple(size(l.W) for l in uw_NN), uw_NN[1].σ, Int(Nz))
    m = get(_models, key, nothing)
    if m === nothing
        m = _models[key] = Model(uw_NN, vw_NN, wT_NN, Nz)
    else
        set_weights!(m, vcat((Float32.(Flux.destructure(nn)[1]) for nn in (uw_NN, vw_NN, wT_NN))...))
    end
    return m
end

# ---- timesteppers (replace OrdinaryDiffEq algorithm objects; dt in units of τ) --------------------------
struct Stepper; dt::Float32; scheme::Cint; precision::Cint; end
Euler(dt; precision=PREC_FP32) = Stepper(dt, SCHEME_EULER, precision)
Heun(dt; precision=PREC_FP32) = Stepper(dt, SCHEME_RK2, precision)
RK4(dt; precision=PREC_FP32) = Stepper(dt, SCHEME_RK4, precision)
# explicit Euler + backward-Euler vertical diffusion: the split step of NDE_oceananigans.jl:61-101 (FP32 only; differentiable:
# opz_loss_grad carries the adjoint of the tridiagonal solve)
IMEXEuler(dt) = Stepper(dt, SCHEME_IMEX, PREC_FP32)

function steps_from_ts(ts, dt)
    length(ts) < 2 && return 0, 1, Float32(first(ts))
    save_every = round(Int, (ts[2] - ts[1]) / dt)
    @assert save_every ≥ 1 && all(isapprox.(diff(ts), ts[2] - ts[1]; rtol=1e-5)) "ts must be uniform multiples of dt"
    return save_every * (length(ts) - 1), save_every, Float32(ts[1])
end

# ---- solve_NDE_mutating (training_postprocessing.jl:55-159) ---------------------------------------------
function solve_NDE_mutating(uw_NN, vw_NN, wT_NN, scalings, constants, BCs, derivatives, uvT₀, ts, timestepper::Stepper, conditions; Q_diurnal=0f0)
    model = cached_model(uw_NN, vw_NN, wT_NN, constants.Nz)
    params = Ref(Params(constants, scalings, conditions; inference=true, diurnal=BCs.wT.top isa Function))
    n_steps, save_every, t0 = steps_from_ts(ts, timestepper.dt)
    x0 = Float32.(uvT₀); bc = pack_BCs(BCs; Q=Q_diurnal)
    out = Array{Float32}(undef, 3constants.Nz, length(ts))
    GC.@preserve x0 bc out check(ccall((:opz_rollout_forward, libopz), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{OpzParams}, Ptr{Float32}, Ptr{Float32}, Int64, Cint, Cfloat, Cfloat, Cint, Cint, Cint, Ptr{Float32}),
        ctx().h, model.h, params, x0, bc, 1, timestepper.scheme, timestepper.dt, t0, n_steps, save_every, timestepper.precision, out))
    return out
end

# ---- predict_NDE / predict_flux (NDE_training.jl:83-165) --------------------------------------------------
function _rhs(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants; t=0f0, fluxes=false)
    model = cached_model(uw_NN, vw_NN, wT_NN, constants.Nz)
    params = Ref(Params(constants, scalings, conditions; inference=false))
    x32 = Float32.(x); bc = pack_BCs(BCs)
    dxdt = similar(x32); fl = Array{Float32}(undef, constants.Nz + 1, 3)
    GC.@preserve x32 bc dxdt fl check(ccall((:opz_rhs, libopz), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{OpzParams}, Ptr{Float32}, Ptr{Float32}, Cfloat, Int64, Cint, Ptr{Float32}, Ptr{Float32}),
        ctx().h, model.h, params, x32, bc, t, 1, PREC_FP32, dxdt, fluxes ? pointer(fl) : C_NULL))
    return dxdt, fl
end
predict_NDE(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants, derivatives, filters) =
    _rhs(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants)[1]
function predict_flux(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants, derivatives, filters)
    _, fl = _rhs(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants; fluxes=true)
    return fl[:, 1], fl[:, 2], fl[:, 3]
end

# ---- loss_NDE + gradient (NDE_training.jl:290-333) -> (total, components, grad) ---------------------------
# Multi-GPU: x0 / bc / target hold THIS rank's cases (shard_range) and B_total the global number of simulations; with a
# communicator on the context (comm_init!) the returned loss and gradient are already summed over the ranks.
function loss_and_gradient(model::Model, params::OpzParams, weights::Vector{Float32}, x0::Matrix{Float32}, bc::Matrix{Float32},
                           target::Array{Float32,3}, loss_w::Vector{Float32}, timestepper::Stepper, n_steps, save_every; t0=0f0,
                           B_total=size(x0, 2))
    set_weights!(model, weights)
    B = size(x0, 2)
    loss = zeros(Float32, 7); grad = zeros(Float32, 3model.P)
    GC.@preserve x0 bc target loss_w loss grad check(ccall((:opz_loss_grad, libopz), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{OpzParams}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Cint, Cfloat, Cfloat, Cint, Cint,
         Ptr{Float32}, Ptr{Float32}, Cint, Ptr{Float32}, Ptr{Float32}),
        ctx().h, model.h, Ref(params), x0, bc, B, B_total, timestepper.scheme, timestepper.dt, t0, n_steps, save_every,
        target, loss_w, timestepper.precision, loss, grad))
    return loss[1], loss[2:7], grad
end

# ---- the same plus d loss / d (ν₀, ν₋, ΔRi, Riᶜ, Pr): the gradient optimise_modified_pacanowski_philander takes with Zygote
#      (diffusivity_parameter_optimisation.jl:1-33, 112-200); unscaled, the caller multiplies by its scalings.parameters ------
function loss_and_gradient_mpp(model::Model, params::OpzParams, x0::Matrix{Float32}, bc::Matrix{Float32},
                               target::Array{Float32,3}, loss_w::Vector{Float32}, timestepper::Stepper, n_steps, save_every; t0=0f0)
    B = size(x0, 2)
    loss = zeros(Float32, 7); grad = zeros(Float32, 3model.P); gmpp = zeros(Float32, 5)
    GC.@preserve x0 bc target loss_w loss grad gmpp check(ccall((:opz_loss_grad_mpp, libopz), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ref{OpzParams}, Ptr{Float32}, Ptr{Float32}, Int64, Int64, Cint, Cfloat, Cfloat, Cint, Cint,
         Ptr{Float32}, Ptr{Float32}, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
        ctx().h, model.h, Ref(params), x0, bc, B, B, timestepper.scheme, timestepper.dt, t0, n_steps, save_every,
        target, loss_w, timestepper.precision, loss, grad, gmpp))
    return loss[1], loss[2:7], grad, gmpp
end

# ---- predict(𝒱, model) (src/predict.jl:12-34): one batched launch --------------------------------------
function predict(𝒱, model::Model, which_net::Integer; scaled=false, precision=PREC_FP32)
    x = Float32.(𝒱.uvT_scaled)                      # 3Nz × Nt  == C [Nt][3Nz]
    y = Array{Float32}(undef, model.Nz - 1, size(x, 2))
    GC.@preserve x y check(ccall((:opz_mlp_forward, libopz), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float32}, Int64, Cint, Ptr{Float32}), ctx().h, model.h, which_net, x, size(x, 2), precision, y))
    return scaled ? (y, 𝒱.scaled) : (𝒱.unscale_fn(y), 𝒱.coarse)
end

end # module
