"""Host-side mirror of the reference's Julia interface for the NDE hot path.

Julia is not available in this image, so the host side above the C ABI is written in Python and
keeps the reference's names, argument order and meaning (INTEGRATION.md shows the equivalent
Julia `ccall` shim, julia/OPZ.jl).  Everything numerical happens in libopz.so on the GPU; this
file only packs arguments.

Mirrored reference symbols (file:line in /root/reference):
  ZeroMeanUnitVarianceScaling, scale, unscale      src/DataWrangling/feature_scaling.jl:7-23,53-54
  Dᶜ, Dᶠ                                           src/differentiation_operators.jl:6-29
  Dense / Chain / destructure                      Flux 0.11.6 as used by wind_mixing/construct_NN.jl:29-34
  prepare_parameters_NDE_training                  wind_mixing/src/NDE_training.jl:1-44
  NDE, predict_NDE, predict_flux                   wind_mixing/src/NDE_training.jl:56-165
  prepare_BCs, solve_NDE_mutating                  wind_mixing/src/training_postprocessing.jl:35-159
  loss, loss_NDE / loss_gradient_NDE, train_NDE    wind_mixing/src/loss.jl:1-42, NDE_training.jl:167-374
  predict                                          src/predict.jl:12-34
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import unicodedata
from types import SimpleNamespace
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib as L

# --------------------------------------------------------------------------------------------------
# activations (names as in NNlib) and scalers
# --------------------------------------------------------------------------------------------------
identity, relu, tanh, mish, swish, leakyrelu = "identity", "relu", "tanh", "mish", "swish", "leakyrelu"
_ACT = {identity: L.ACT_IDENTITY, relu: L.ACT_RELU, tanh: L.ACT_TANH, mish: L.ACT_MISH, swish: L.ACT_SWISH,
        leakyrelu: L.ACT_LEAKYRELU}


@dataclasses.dataclass
class ZeroMeanUnitVarianceScaling:
    """feature_scaling.jl:7-23.  Callable: s(x) == scale(x, s); inv(s)(y) == unscale(y, s)."""
    μ: float
    σ: float

    @classmethod
    def from_data(cls, data):
        data = np.asarray(data, dtype=np.float64)
        return cls(float(data.mean()), float(data.std(ddof=1)))  # Julia std() is the n-1 estimator

    def __call__(self, x):
        return (np.asarray(x) - self.μ) / self.σ


def scale(x, s: ZeroMeanUnitVarianceScaling):
    return (np.asarray(x) - s.μ) / s.σ


def unscale(y, s: ZeroMeanUnitVarianceScaling):
    return s.σ * np.asarray(y) + s.μ


def inv(s: ZeroMeanUnitVarianceScaling) -> Callable:
    return lambda y: unscale(y, s)


def Dᶜ(N: int, Δ: float) -> np.ndarray:
    """N x (N+1) face->cell derivative (differentiation_operators.jl:6-14).  Kept for API parity;
    on the device both operators are two-point differences."""
    D = np.zeros((N, N + 1))
    idx = np.arange(N)
    D[idx, idx] = -1.0
    D[idx, idx + 1] = 1.0
    return D / Δ


def Dᶠ(N: int, Δ: float) -> np.ndarray:
    """(N+1) x N cell->face derivative with zero first/last rows (:21-29)."""
    D = np.zeros((N + 1, N))
    idx = np.arange(1, N)
    D[idx, idx - 1] = -1.0
    D[idx, idx] = 1.0
    return D / Δ


D_cell, D_face = Dᶜ, Dᶠ

# --------------------------------------------------------------------------------------------------
# Flux-shaped NN containers (weights only; evaluation happens on the GPU)
# --------------------------------------------------------------------------------------------------


class Dense:
    """Dense(in, out, σ): W is out x in, b is out; default init glorot_uniform / zeros (Flux 0.11)."""

    def __init__(self, nin: int, nout: int, σ: str = identity, *, rng: Optional[np.random.Generator] = None,
                 W: Optional[np.ndarray] = None, b: Optional[np.ndarray] = None):
        self.nin, self.nout, self.σ = int(nin), int(nout), σ
        if W is None:
            rng = rng or np.random.default_rng()
            lim = math.sqrt(6.0 / (nin + nout))
            W = rng.uniform(-lim, lim, size=(nout, nin))
        self.W = np.asarray(W, dtype=np.float32).reshape(nout, nin)
        self.b = np.zeros(nout, dtype=np.float32) if b is None else np.asarray(b, dtype=np.float32).reshape(nout)


class Chain:
    def __init__(self, *layers: Dense):
        self.layers = list(layers)
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if a.nout != b.nin:
                raise ValueError("layer sizes do not chain")
        acts = {l.σ for l in self.layers[:-1]}
        if len(acts) > 1:
            raise ValueError("libopz supports one hidden activation per Chain")
        if self.layers[-1].σ != identity:
            raise ValueError("the last Dense layer must be linear")

    @property
    def sizes(self):
        return [self.layers[0].nin] + [l.nout for l in self.layers]

    @property
    def activation(self) -> str:
        return self.layers[0].σ if len(self.layers) > 1 else identity


def destructure(chain: Chain):
    """Flux.destructure: flat Float32 vector [vec(W1); b1; vec(W2); b2; ...] (column-major W) and a
    restructure closure `re`."""
    flat = np.concatenate([np.concatenate([l.W.T.reshape(-1), l.b]) for l in chain.layers]).astype(np.float32)
    spec = [(l.nin, l.nout, l.σ) for l in chain.layers]

    def re(w):
        w = np.asarray(w, dtype=np.float32)
        if w.size != flat.size:
            raise ValueError("restructure: wrong number of parameters")
        layers, off = [], 0
        for nin, nout, σ in spec:
            W = w[off:off + nin * nout].reshape(nin, nout).T
            off += nin * nout
            b = w[off:off + nout]
            off += nout
            layers.append(Dense(nin, nout, σ, W=W.copy(), b=b.copy()))
        return Chain(*layers)

    return flat, re


# --------------------------------------------------------------------------------------------------
# context / model handles
# --------------------------------------------------------------------------------------------------


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def comm_unique_id() -> bytes:
    """opz_comm_get_unique_id: the 128-byte NCCL id rank 0 creates and ships to the other ranks by any host means."""
    buf = C.create_string_buffer(L.COMM_ID_BYTES)
    L.check(L.lib.opz_comm_get_unique_id(buf))
    return buf.raw


class Context:
    """opz_ctx: one GPU, one stream; optionally one rank of the training step's communicator."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        L.check(L.lib.opz_ctx_create(device, C.byref(h)))
        self._h = h
        self.device = device

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int):
        """opz_ctx_comm_init (collective over all ranks): afterwards loss_grad* return the GLOBAL loss and gradient."""
        if len(unique_id) != L.COMM_ID_BYTES:
            raise ValueError("unique_id must be the 128 bytes of comm_unique_id()")
        L.check(L.lib.opz_ctx_comm_init(self._h, C.c_char_p(unique_id), int(n_ranks), int(rank)))

    def comm_adopt(self, nccl_comm_ptr: int, n_ranks: int, rank: int):
        L.check(L.lib.opz_ctx_comm_adopt(self._h, C.c_void_p(nccl_comm_ptr), int(n_ranks), int(rank)))

    def comm_destroy(self):
        L.check(L.lib.opz_ctx_comm_destroy(self._h))

    @property
    def comm_info(self):
        """(n_ranks, rank, all-reduce groups enqueued so far)"""
        n, r, c = C.c_int(), C.c_int(), C.c_int64()
        L.check(L.lib.opz_ctx_comm_info(self._h, C.byref(n), C.byref(r), C.byref(c)))
        return n.value, r.value, c.value

    def set_stream(self, cuda_stream_ptr: int):
        L.check(L.lib.opz_ctx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        L.check(L.lib.opz_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(L.lib.opz_ctx_launch_count(self._h))

    def close(self):
        if self._h:
            L.lib.opz_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class NDEModel:
    """opz_model: the three Chains uw_NN, vw_NN, wT_NN resident on the GPU."""

    def __init__(self, uw_NN: Chain, vw_NN: Chain, wT_NN: Chain, Nz: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        if not (uw_NN.sizes == vw_NN.sizes == wT_NN.sizes and uw_NN.activation == vw_NN.activation == wT_NN.activation):
            raise ValueError("the three nets must share one architecture")
        self.Nz = int(Nz)
        self.sizes = list(uw_NN.sizes)
        self.activation = uw_NN.activation
        ws = [destructure(n)[0] for n in (uw_NN, vw_NN, wT_NN)]
        sizes = (C.c_int * len(self.sizes))(*self.sizes)
        h = C.c_void_p()
        L.check(L.lib.opz_model_create(self.ctx._h, self.Nz, len(self.sizes), sizes, _ACT[self.activation],
                                       _fptr(ws[0]), _fptr(ws[1]), _fptr(ws[2]), C.byref(h)))
        self._h = h
        self.P = int(L.lib.opz_model_n_params(h))

    def set_weights(self, weights):
        w = _f32(weights, (3 * self.P,))
        L.check(L.lib.opz_model_set_weights(self._h, _fptr(w)))

    def get_weights(self) -> np.ndarray:
        w = np.empty(3 * self.P, dtype=np.float32)
        L.check(L.lib.opz_model_get_weights(self._h, _fptr(w)))
        return w

    def close(self):
        if getattr(self, "_h", None):
            L.lib.opz_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------------
_SCALING_ORDER = ("u", "v", "T", "uw", "vw", "wT")


def make_params(constants, scalings, flags: int, eps: float = 1e-7, diurnal_period: float = 86400.0) -> L.OpzParams:
    # Python NFKC-normalises identifiers (Riᶜ -> Ric), so attribute names are looked up normalised
    g = lambda k, d=0.0: float(getattr(constants, unicodedata.normalize("NFKC", k), d))
    p = L.OpzParams()
    p.nz = int(constants.Nz)
    p.H, p.tau, p.f, p.g, p.alpha = g("H"), g("τ"), g("f"), g("g"), g("α")
    p.nu0, p.nu_m, p.Ric, p.dRi, p.Pr = g("ν0"), g("ν_m"), g("Riᶜ"), g("ΔRi", 1.0), g("Pr", 1.0)
    p.kappa = g("κ")
    p.eps = eps
    for i, n in enumerate(_SCALING_ORDER):
        s = scalings[n] if isinstance(scalings, dict) else getattr(scalings, n)
        p.mu[i], p.sigma[i] = s.μ, s.σ
    p.flags = flags
    p.diurnal_period = diurnal_period
    return p


def _cond(conditions, key, default=False):
    if isinstance(conditions, dict):
        return bool(conditions.get(key, default))
    return bool(getattr(conditions, key, default))


def inference_flags(conditions) -> int:
    """Flag set of solve_NDE_mutating: mPP unconditional, boundary fluxes minus scale(0)
    (training_postprocessing.jl:65,82-93,114-121)."""
    f = L.FLAG_MPP | L.FLAG_BC_MINUS_SCALE0
    if _cond(conditions, "convective_adjustment"):
        f |= L.FLAG_CA
        if _cond(conditions, "ca_select_dudz"):
            f |= L.FLAG_CA_SELECT_DUDZ
    if _cond(conditions, "diurnal"):
        f |= L.FLAG_DIURNAL
        if _cond(conditions, "diurnal_minus_scale0"):
            f |= L.FLAG_DIURNAL_MINUS_SCALE0
    return f


def training_flags(conditions) -> int:
    """Flag set of NDE / predict_flux (NDE_training.jl:83-147)."""
    f = 0
    if _cond(conditions, "modified_pacanowski_philander"):
        f |= L.FLAG_MPP | L.FLAG_RI_EPS
    elif _cond(conditions, "convective_adjustment"):
        raise NotImplementedError("the reference's CA-only training branch reads an undefined κ "
                                  "(NDE_training.jl:142); it is not reproduced")
    if _cond(conditions, "zero_weights"):
        f |= L.FLAG_BC_MINUS_SCALE0
    if _cond(conditions, "smooth_NN"):
        f |= L.FLAG_SMOOTH_NN
    if _cond(conditions, "smooth_Ri"):
        f |= L.FLAG_SMOOTH_RI
    if _cond(conditions, "diurnal"):
        f |= L.FLAG_DIURNAL
    return f


def prepare_parameters_NDE_training(𝒟train, uw_NN, vw_NN, wT_NN, f, Nz, g, α, ν0, ν_m, Riᶜ, ΔRi, Pr, κ, conditions):
    """NDE_training.jl:1-44.  𝒟train needs: .uw.z (face heights), .t (times), .scalings (dict)."""
    z = np.asarray(𝒟train.uw.z)
    t = np.asarray(𝒟train.t).reshape(len(𝒟train.t), -1)[:, 0]
    H = abs(float(z[-1] - z[0]))
    τ = abs(float(t[-1] - t[0]))
    sc = 𝒟train.scalings
    uw_w, re_uw = destructure(uw_NN)
    vw_w, re_vw = destructure(vw_NN)
    wT_w, re_wT = destructure(wT_NN)
    n1, n2, n3 = len(uw_w), len(vw_w), len(wT_w)
    constants = SimpleNamespace(H=H, τ=τ, f=f, Nz=Nz, g=g, α=α)
    if _cond(conditions, "modified_pacanowski_philander"):
        constants.ν0, constants.ν_m, constants.Riᶜ, constants.ΔRi, constants.Pr = ν0, ν_m, Riᶜ, ΔRi, Pr
    if _cond(conditions, "convective_adjustment"):
        constants.κ = κ
    scalings = SimpleNamespace(**{n: sc[n] for n in _SCALING_ORDER})
    derivatives = SimpleNamespace(cell=Dᶜ(Nz, 1 / Nz).astype(np.float32), face=Dᶠ(Nz, 1 / Nz).astype(np.float32))
    NN_constructions = SimpleNamespace(uw=re_uw, vw=re_vw, wT=re_wT)
    weights = np.concatenate([uw_w, vw_w, wT_w]).astype(np.float32)
    NN_sizes = SimpleNamespace(uw=n1, vw=n2, wT=n3)
    NN_ranges = SimpleNamespace(uw=slice(0, n1), vw=slice(n1, n1 + n2), wT=slice(n1 + n2, n1 + n2 + n3))
    filters = None  # the 3-point box filters are applied on the device (flags SMOOTH_NN / SMOOTH_RI)
    return constants, scalings, derivatives, NN_constructions, weights, NN_sizes, NN_ranges, filters


def prepare_BCs(𝒟, scalings, wT_top_function=None):
    """training_postprocessing.jl:35-53: scaled boundary fluxes of the first data column."""
    sc = scalings if not isinstance(scalings, dict) else SimpleNamespace(**scalings)
    uw, vw, wT = np.asarray(𝒟.uw.coarse), np.asarray(𝒟.vw.coarse), np.asarray(𝒟.wT.coarse)
    top = (lambda t: sc.wT(wT_top_function(t))) if wT_top_function is not None else float(sc.wT(wT[-1, 0]))
    return SimpleNamespace(uw=SimpleNamespace(top=float(sc.uw(uw[-1, 0])), bottom=float(sc.uw(uw[0, 0]))),
                           vw=SimpleNamespace(top=float(sc.vw(vw[-1, 0])), bottom=float(sc.vw(vw[0, 0]))),
                           wT=SimpleNamespace(top=top, bottom=float(sc.wT(wT[0, 0]))))


def pack_BCs(BCs, Q_diurnal: float = 0.0) -> np.ndarray:
    """NamedTuple-shaped BCs -> one row of the C ABI's bc[B][8]."""
    top = BCs.wT.top
    return np.array([BCs.uw.bottom, BCs.uw.top, BCs.vw.bottom, BCs.vw.top, BCs.wT.bottom,
                     0.0 if callable(top) else top, Q_diurnal, 0.0], dtype=np.float32)


# --------------------------------------------------------------------------------------------------
# timesteppers (replace OrdinaryDiffEq algorithm objects; dt is in units of τ)
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class _Stepper:
    dt: float
    scheme: int = L.SCHEME_RK4
    precision: int = L.PREC_FP32


def Euler(dt, precision=L.PREC_FP32):
    return _Stepper(dt, L.SCHEME_EULER, precision)


def Heun(dt, precision=L.PREC_FP32):
    return _Stepper(dt, L.SCHEME_RK2, precision)


def RK4(dt, precision=L.PREC_FP32):
    return _Stepper(dt, L.SCHEME_RK4, precision)


def IMEXEuler(dt):
    """Explicit Euler for the NN / boundary fluxes and Coriolis + backward-Euler vertical diffusion: the split step of the
    reference's Oceananigans path (NDE_oceananigans.jl:61-101).  FP32 only."""
    return _Stepper(dt, L.SCHEME_IMEX, L.PREC_FP32)


def _steps_from_ts(ts, dt):
    ts = np.asarray(ts, dtype=np.float64)
    if ts.size < 2:
        return 0, 1, float(ts[0]) if ts.size else 0.0
    gap = (ts[1] - ts[0]) / dt
    save_every = int(round(gap))
    if save_every < 1 or abs(gap - save_every) > 1e-4 * max(1.0, gap) or not np.allclose(np.diff(ts), ts[1] - ts[0], rtol=1e-5):
        raise ValueError("ts must be uniformly spaced multiples of the timestepper's dt")
    return save_every * (ts.size - 1), save_every, float(ts[0])


# --------------------------------------------------------------------------------------------------
# raw ensemble entry points (thin wrappers over the C ABI)
# --------------------------------------------------------------------------------------------------


def rhs(model: NDEModel, params: L.OpzParams, x, bc, t: float, *, with_fluxes=False, precision=L.PREC_FP32):
    x = _f32(x)
    B = x.shape[0]
    bc = _f32(bc, (B, L.BC_STRIDE))
    dxdt = np.empty_like(x)
    fl = np.empty((B, 3, model.Nz + 1), dtype=np.float32) if with_fluxes else None
    L.check(L.lib.opz_rhs(model.ctx._h, model._h, C.byref(params), _fptr(x), _fptr(bc), float(t), B, precision,
                          _fptr(dxdt), _fptr(fl) if with_fluxes else None))
    return (dxdt, fl) if with_fluxes else dxdt


def rollout_forward(model: NDEModel, params: L.OpzParams, x0, bc, scheme: int, dt_nd: float, n_steps: int,
                    save_every: int, *, t0: float = 0.0, precision: int = L.PREC_FP32) -> np.ndarray:
    """Ensemble solve: x0[B,3Nz], bc[B,8] -> out[B, n_save, 3Nz]."""
    x0 = _f32(x0)
    B = x0.shape[0]
    bc = _f32(bc, (B, L.BC_STRIDE))
    n_save = n_steps // save_every + 1 if save_every > 0 else 1
    out = np.empty((B, n_save, x0.shape[1]), dtype=np.float32)
    L.check(L.lib.opz_rollout_forward(model.ctx._h, model._h, C.byref(params), _fptr(x0), _fptr(bc), B, scheme,
                                      float(dt_nd), float(t0), int(n_steps), int(save_every), precision, _fptr(out)))
    return out


def rollout_forward_dev(model: NDEModel, params: L.OpzParams, x0_ptr: int, bc_ptr: int, B: int, scheme: int,
                        dt_nd: float, n_steps: int, save_every: int, out_ptr: int, *, t0: float = 0.0,
                        precision: int = L.PREC_FP32) -> None:
    """Device-pointer variant: enqueues on the context's stream and returns."""
    L.check(L.lib.opz_rollout_forward_dev(model.ctx._h, model._h, C.byref(params), C.c_void_p(x0_ptr),
                                          C.c_void_p(bc_ptr), B, scheme, float(dt_nd), float(t0), int(n_steps),
                                          int(save_every), precision, C.c_void_p(out_ptr)))


def loss_grad(model: NDEModel, params: L.OpzParams, x0, bc, target, loss_w, scheme: int, dt_nd: float, n_steps: int,
              save_every: int, *, t0: float = 0.0, B_total: Optional[int] = None, precision: int = L.PREC_FP32):
    """Returns (total, components[6], grad[3P]) of the weighted profile loss (NDE_training.jl:290-323)."""
    x0 = _f32(x0)
    B = x0.shape[0]
    bc = _f32(bc, (B, L.BC_STRIDE))
    n_save = n_steps // save_every + 1
    target = _f32(target, (B, n_save, x0.shape[1]))
    lw = _f32(loss_w, (6,))
    loss_out = np.zeros(7, dtype=np.float32)
    grad = np.zeros(3 * model.P, dtype=np.float32)
    L.check(L.lib.opz_loss_grad(model.ctx._h, model._h, C.byref(params), _fptr(x0), _fptr(bc), B,
                                B if B_total is None else int(B_total), scheme, float(dt_nd), float(t0), int(n_steps),
                                int(save_every), _fptr(target), _fptr(lw), precision, _fptr(loss_out), _fptr(grad)))
    return float(loss_out[0]), loss_out[1:].copy(), grad


def loss_grad_mpp(model: NDEModel, params: L.OpzParams, x0, bc, target, loss_w, scheme: int, dt_nd: float, n_steps: int,
                  save_every: int, *, t0: float = 0.0, B_total: Optional[int] = None, precision: int = L.PREC_FP32):
    """loss_grad plus d loss / d (nu0, nu_m, dRi, Ric, Pr) (diffusivity_parameter_optimisation.jl:2, unscaled):
    returns (total, components[6], grad[3P], grad_mpp[5])."""
    x0 = _f32(x0)
    B = x0.shape[0]
    bc = _f32(bc, (B, L.BC_STRIDE))
    n_save = n_steps // save_every + 1
    target = _f32(target, (B, n_save, x0.shape[1]))
    lw = _f32(loss_w, (6,))
    loss_out = np.zeros(7, dtype=np.float32)
    grad = np.zeros(3 * model.P, dtype=np.float32)
    gmpp = np.zeros(5, dtype=np.float32)
    L.check(L.lib.opz_loss_grad_mpp(model.ctx._h, model._h, C.byref(params), _fptr(x0), _fptr(bc), B,
                                    B if B_total is None else int(B_total), scheme, float(dt_nd), float(t0), int(n_steps),
                                    int(save_every), _fptr(target), _fptr(lw), precision, _fptr(loss_out), _fptr(grad),
                                    _fptr(gmpp)))
    return float(loss_out[0]), loss_out[1:].copy(), grad, gmpp


def loss_grad_dev(model: NDEModel, params: L.OpzParams, x0_ptr: int, bc_ptr: int, B: int, B_total: int, target_ptr: int,
                  loss_w, scheme: int, dt_nd: float, n_steps: int, save_every: int, loss_out_ptr: int, grad_out_ptr: int, *,
                  t0: float = 0.0, precision: int = L.PREC_FP32) -> None:
    """Device-pointer variant of loss_grad: loss_out[7] and grad_out[3P] are DEVICE buffers (e.g. one torch
    tensor that the caller then all-reduces over NCCL); enqueues on the context's stream and returns."""
    lw = _f32(loss_w, (6,))
    L.check(L.lib.opz_loss_grad_dev(model.ctx._h, model._h, C.byref(params), C.c_void_p(x0_ptr), C.c_void_p(bc_ptr), int(B),
                                    int(B_total), scheme, float(dt_nd), float(t0), int(n_steps), int(save_every),
                                    C.c_void_p(target_ptr), _fptr(lw), precision, C.c_void_p(loss_out_ptr),
                                    C.c_void_p(grad_out_ptr)))


def flux_loss_grad(model: NDEModel, params: L.OpzParams, which_net: int, x, bc, flux_target, gradient_scaling: float = 1e-4):
    """Supervised pre-training loss of ONE net and its gradient (train_NN / NN_loss, NN_training.jl:207-249): x [N, 3Nz],
    bc [N, 8], flux_target [N, Nz+1] (scaled) -> (total, flux mse, flux-gradient mse, grad[P]); means over the N samples."""
    x = _f32(x)
    N = x.shape[0]
    bc = _f32(bc, (N, L.BC_STRIDE))
    ft = _f32(flux_target, (N, x.shape[1] // 3 + 1))
    loss_out = np.zeros(3, dtype=np.float32)
    grad = np.zeros(model.P, dtype=np.float32)
    L.check(L.lib.opz_flux_loss_grad(model.ctx._h, model._h, C.byref(params), int(which_net), _fptr(x), _fptr(bc), _fptr(ft), N,
                                     float(gradient_scaling), _fptr(loss_out), _fptr(grad)))
    return float(loss_out[0]), float(loss_out[1]), float(loss_out[2]), grad


def profile_diagnostics(model: NDEModel, params: L.OpzParams, sol, bc, *, t0: float = 0.0, dt_save: float = 0.0, target=None,
                        with_mpp_part: bool = True):
    """Diagnostics of a finished rollout (NDE_profile, training_postprocessing.jl:396-450): returns a dict with
    `fluxes` [B, n_save, 3, Nz+1] (predict_flux at every snapshot), `fluxes_mpp` (NN outputs zeroed), `Ri`
    [B, n_save, Nz+1] and, with a target, `loss_per_tstep` [B, n_save, 3] (loss.jl:44-46)."""
    sol = _f32(sol)
    B, n_save, S = sol.shape
    nz = S // 3
    bc = _f32(bc, (B, L.BC_STRIDE))
    fl = np.empty((B, n_save, 3, nz + 1), dtype=np.float32)
    fm = np.empty_like(fl) if with_mpp_part else None
    ri = np.empty((B, n_save, nz + 1), dtype=np.float32)
    tg = _f32(target, sol.shape) if target is not None else None
    lt = np.empty((B, n_save, 3), dtype=np.float32) if tg is not None else None
    null = C.POINTER(C.c_float)()
    L.check(L.lib.opz_profile_diagnostics(model.ctx._h, model._h, C.byref(params), _fptr(sol), _fptr(bc), B, n_save, float(t0),
                                          float(dt_save), _fptr(tg) if tg is not None else null, _fptr(fl),
                                          _fptr(fm) if fm is not None else null, _fptr(ri), _fptr(lt) if lt is not None else null))
    out = {"fluxes": fl, "Ri": ri}
    if fm is not None:
        out["fluxes_mpp"] = fm
    if lt is not None:
        out["loss_per_tstep"] = lt
    return out


def mlp_forward(model: NDEModel, which_net: int, x, precision: int = L.PREC_FP32) -> np.ndarray:
    x = _f32(x)
    y = np.empty((x.shape[0], model.Nz - 1), dtype=np.float32)
    L.check(L.lib.opz_mlp_forward(model.ctx._h, model._h, which_net, _fptr(x), x.shape[0], precision, _fptr(y)))
    return y


# --------------------------------------------------------------------------------------------------
# reference-shaped entry points
# --------------------------------------------------------------------------------------------------
_model_cache: dict = {}


def _model_for(uw_NN, vw_NN, wT_NN, Nz) -> NDEModel:
    """Device model for three host Chains.  Cached per ARCHITECTURE (layer sizes, activation, Nz, context) — never per
    object identity: `NDE()` rebuilds its Chains from `p` on every call (NDE_training.jl:62-64) and CPython reuses freed
    addresses, so an id()-keyed cache would silently evaluate with the weights of an earlier call.  The weights are
    uploaded on every call (<= 2.5 MB), hit or miss."""
    ctx = default_context()
    if not (uw_NN.sizes == vw_NN.sizes == wT_NN.sizes and uw_NN.activation == vw_NN.activation == wT_NN.activation):
        raise ValueError("the three nets must share one architecture")
    key = (tuple(uw_NN.sizes), uw_NN.activation, int(Nz), id(ctx))
    m = _model_cache.get(key)
    if m is not None and (m._h is None or m.ctx is not ctx or m.ctx._h is None):
        m = None
    if m is None:
        if len(_model_cache) > 8:
            for old in _model_cache.values():
                old.close()
            _model_cache.clear()
        m = _model_cache[key] = NDEModel(uw_NN, vw_NN, wT_NN, Nz, ctx=ctx)
    else:
        m.set_weights(np.concatenate([destructure(n)[0] for n in (uw_NN, vw_NN, wT_NN)]))
    return m


def solve_NDE_mutating(uw_NN, vw_NN, wT_NN, scalings, constants, BCs, derivatives, uvT0, ts, timestepper, conditions,
                       Q_diurnal: float = 0.0):
    """training_postprocessing.jl:55-159.  `timestepper` is Euler/Heun/RK4(dt) and `ts` uniformly spaced
    multiples of dt.  Returns the 3Nz x length(ts) solution array like `Array(solve(...; saveat=ts))`.
    A time-dependent `BCs.wT.top` is expressed through `Q_diurnal` (data_containers.jl:131-156)."""
    cond = dict(conditions) if isinstance(conditions, dict) else dict(vars(conditions))
    if callable(BCs.wT.top):
        cond["diurnal"] = True
    params = make_params(constants, scalings, inference_flags(cond))
    model = _model_for(uw_NN, vw_NN, wT_NN, int(constants.Nz))
    n_steps, save_every, t0 = _steps_from_ts(ts, timestepper.dt)
    out = rollout_forward(model, params, np.asarray(uvT0, dtype=np.float32)[None, :], pack_BCs(BCs, Q_diurnal)[None, :],
                          timestepper.scheme, timestepper.dt, n_steps, save_every, t0=t0,
                          precision=timestepper.precision)
    return out[0].T.copy()


def predict_NDE(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants, derivatives, filters, t: float = 0.0):
    """NDE_training.jl:149-165: the right-hand side for one column."""
    params = make_params(constants, scalings, training_flags(conditions))
    model = _model_for(uw_NN, vw_NN, wT_NN, int(constants.Nz))
    return rhs(model, params, np.asarray(x, dtype=np.float32)[None, :], pack_BCs(BCs)[None, :], t)[0]


def predict_flux(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants, derivatives, filters, t: float = 0.0):
    """NDE_training.jl:83-147: total scaled fluxes (uw, vw, wT) on the Nz+1 faces."""
    params = make_params(constants, scalings, training_flags(conditions))
    model = _model_for(uw_NN, vw_NN, wT_NN, int(constants.Nz))
    _, fl = rhs(model, params, np.asarray(x, dtype=np.float32)[None, :], pack_BCs(BCs)[None, :], t, with_fluxes=True)
    return fl[0, 0], fl[0, 1], fl[0, 2]


def NDE(x, p, t, NN_ranges, NN_constructions, conditions, scalings, constants, derivatives, filters):
    """NDE_training.jl:56-66: p = [weights; uw_b, uw_t, vw_b, vw_t, wT_b, wT_t]."""
    p = np.asarray(p, dtype=np.float32)
    nW = NN_ranges.wT.stop
    uw_NN = NN_constructions.uw(p[NN_ranges.uw])
    vw_NN = NN_constructions.vw(p[NN_ranges.vw])
    wT_NN = NN_constructions.wT(p[NN_ranges.wT])
    b = p[nW:nW + 6]
    BCs = SimpleNamespace(uw=SimpleNamespace(bottom=b[0], top=b[1]), vw=SimpleNamespace(bottom=b[2], top=b[3]),
                          wT=SimpleNamespace(bottom=b[4], top=b[5]))
    return predict_NDE(uw_NN, vw_NN, wT_NN, x, BCs, conditions, scalings, constants, derivatives, filters, t)


def loss(a, b):
    """Flux.mse (loss.jl:1-3)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.mean((a - b) ** 2))


def calculate_loss_scalings(losses, fractions, train_gradient):
    """loss.jl:11-31."""
    vs = (1 - fractions.T) / fractions.T * losses.T / (losses.u + losses.v)
    profile_loss = vs * (losses.u + losses.v) + losses.T
    if train_gradient:
        vgs = (1 - fractions.dTdz) / fractions.dTdz * losses.dTdz / (losses.dudz + losses.dvdz)
        gradient_loss = vgs * (losses.dudz + losses.dvdz) + losses.dTdz
        tgs = (1 - fractions.profile) / fractions.profile * profile_loss / gradient_loss
    else:
        vgs = 0
        tgs = 0
    return SimpleNamespace(u=vs, v=vs, T=1, dudz=tgs * vgs, dvdz=tgs * vgs, dTdz=tgs)


class ADAM:
    """Flux.ADAM(η, β): m <- β1 m + (1-β1) g; v <- β2 v + (1-β2) g²; Δ = η m̂ / (sqrt(v̂) + 1e-8)."""

    def __init__(self, η=1e-3, β=(0.9, 0.999)):
        self.eta, self.beta = η, β
        self.state = None

    def apply(self, w: np.ndarray, g: np.ndarray) -> np.ndarray:
        if self.state is None:
            self.state = [np.zeros_like(w), np.zeros_like(w), np.array(self.beta, dtype=np.float64)]
        m, v, bp = self.state
        b1, b2 = self.beta
        m *= b1
        m += (1 - b1) * g
        v *= b2
        v += (1 - b2) * g * g
        step = m / (1 - bp[0]) / (np.sqrt(v / (1 - bp[1])) + 1e-8) * self.eta
        bp *= self.beta
        return (w - step).astype(np.float32)


def train_NDE(uw_NN, vw_NN, wT_NN, 𝒟train, tsteps, timestepper, optimizers, epochs, *, maxiters=500, ν0=1e-4,
              ν_m=1e-1, ΔRi=1.0, Riᶜ=0.25, Pr=1.0, κ=10.0, f=1e-4, α=2e-4, g=9.80665,
              modified_pacanowski_philander=False, convective_adjustment=False, smooth_NN=False, smooth_Ri=False,
              train_gradient=False, zero_weights=False, gradient_scaling=5e-3, training_fractions=None,
              diurnal=False, Q_diurnal=None, cb=None, comm=None, ctx: Optional[Context] = None, _loss_grad_fn=None):
    """NDE_training.jl:167-374 with the data already in memory (the JLD2 loader is out of scope):
    𝒟train carries .uvT_scaled [3Nz, n_steps*n_sim], .t, .uw/.vw/.wT (.scaled, .z), .scalings, .n_simulations.
    Loss + exact discrete-adjoint gradient come from libopz; ADAM and the callback stay on the host.

    Multi-GPU (one process per GPU, every rank calls train_NDE with the SAME data): the n_sim simulations are sharded
    over the ranks (shard_range), every rank passes B_total = n_sim, and the shard losses / gradients are summed
      * inside libopz when `ctx` carries a communicator (Context.comm_init: NCCL all-reduce behind the BPTT kernels), or
      * by `comm` (parallel.make_allreduce_comm: an object with .rank, .world and comm(buf) summing a float32 numpy
        buffer in place), the host-side route used by the gloo tests.
    Every rank then takes the identical ADAM step.  `_loss_grad_fn` (tests only) stands in for the libopz call."""
    from .parallel import shard_range
    assert not (modified_pacanowski_philander and convective_adjustment)
    if zero_weights:
        assert modified_pacanowski_philander
    conditions = dict(modified_pacanowski_philander=modified_pacanowski_philander,
                      convective_adjustment=convective_adjustment, smooth_NN=smooth_NN, smooth_Ri=smooth_Ri,
                      train_gradient=train_gradient, zero_weights=zero_weights, diurnal=diurnal)
    Nz = len(𝒟train.u.z)
    n_sim = int(𝒟train.n_simulations)
    constants, scalings, derivatives, NN_constructions, weights, NN_sizes, NN_ranges, _ = \
        prepare_parameters_NDE_training(𝒟train, uw_NN, vw_NN, wT_NN, f, Nz, g, α, ν0, ν_m, Riᶜ, ΔRi, Pr, κ, conditions)
    t_all = np.asarray(𝒟train.t).reshape(len(𝒟train.t), -1)[:, 0]
    n_steps_data = len(t_all) // n_sim
    tsteps = np.asarray(tsteps)
    uvT = np.asarray(𝒟train.uvT_scaled, dtype=np.float32)
    x0 = np.stack([uvT[:, n_steps_data * i + tsteps[0]] for i in range(n_sim)])
    target = np.stack([uvT[:, n_steps_data * i:n_steps_data * (i + 1)][:, tsteps].T for i in range(n_sim)])
    t_train = t_all[tsteps] / constants.τ
    bc = np.zeros((n_sim, L.BC_STRIDE), dtype=np.float32)
    for i in range(n_sim):
        c = n_steps_data * i + tsteps[0]
        bc[i, :6] = [𝒟train.uw.scaled[0, c], 𝒟train.uw.scaled[-1, c], 𝒟train.vw.scaled[0, c],
                     𝒟train.vw.scaled[-1, c], 𝒟train.wT.scaled[0, c], 𝒟train.wT.scaled[-1, c]]
        if diurnal:
            bc[i, 6] = Q_diurnal[i]
    params = make_params(constants, scalings, training_flags(conditions))
    n_steps, save_every, t0 = _steps_from_ts(t_train, timestepper.dt)
    model = None
    lib_world, lib_rank = 1, 0
    if _loss_grad_fn is None:
        model = NDEModel(uw_NN, vw_NN, wT_NN, Nz, ctx=ctx)
        lib_world, lib_rank, _ = model.ctx.comm_info
    if lib_world > 1 and comm is not None:
        raise ValueError("train_NDE: the context already carries a communicator; do not pass `comm` as well")
    rank, world = (lib_rank, lib_world) if lib_world > 1 else ((comm.rank, comm.world) if comm is not None else (0, 1))
    lo, hi = shard_range(n_sim, rank, world)
    x0_s, bc_s, target_s = x0[lo:hi], bc[lo:hi], target[lo:hi]

    def evaluate(w, lw):
        if _loss_grad_fn is not None:
            total, comps, grad = _loss_grad_fn(w, x0_s, bc_s, target_s, lw, n_sim)
        else:
            model.set_weights(w)
            total, comps, grad = loss_grad(model, params, x0_s, bc_s, target_s, lw, timestepper.scheme, timestepper.dt,
                                           n_steps, save_every, t0=t0, B_total=n_sim, precision=timestepper.precision)
        if comm is not None and world > 1:
            buf = np.concatenate([[total], comps, grad]).astype(np.float32)
            comm(buf)
            total, comps, grad = float(buf[0]), buf[1:7], buf[7:]
        return total, comps, grad

    if training_fractions is None:
        gs = gradient_scaling if train_gradient else 0.0
        lw = np.array([1, 1, 1, gs, gs, gs], dtype=np.float32)
    else:
        ones = np.array([1, 1, 1, 1, 1, 1] if train_gradient else [1, 1, 1, 0, 0, 0], dtype=np.float32)
        _, c, _ = evaluate(weights, ones)
        ls = calculate_loss_scalings(SimpleNamespace(u=c[0], v=c[1], T=c[2], dudz=c[3], dvdz=c[4], dTdz=c[5]),
                                     training_fractions, train_gradient)
        lw = np.array([ls.u, ls.v, ls.T, ls.dudz, ls.dvdz, ls.dTdz], dtype=np.float32)
    history = []
    for opt in optimizers:
        for _epoch in range(epochs):
            for it in range(maxiters):
                total, comps, grad = evaluate(weights, lw)
                history.append((total, comps.copy()))
                if cb is not None:
                    cb(weights, total, comps, lw)
                weights = opt.apply(weights, grad)
    if model is not None:
        model.close()
    out = tuple(NN_constructions.__dict__[k](weights[NN_ranges.__dict__[k]]) for k in ("uw", "vw", "wT"))
    return out + (history,)


def optimise_modified_pacanowski_philander(model: NDEModel, params: L.OpzParams, x0, bc, target, loss_w, timestepper,
                                           n_steps: int, save_every: int, optimizers, maxiters: int = 100, *,
                                           callback=None, comm=None, B_total: Optional[int] = None):
    """Host loop of optimise_modified_pacanowski_philander (diffusivity_parameter_optimisation.jl:36-231): the five
    parameters p = (nu0, nu_m, dRi, Ric, Pr) are optimised in the reference's scaled form p / p_initial (:44-48), clipped to
    its box [0, 10] (:197), with the gradient of the profile loss taken by libopz (opz_loss_grad_mpp) instead of Zygote.
    `model` carries the NN weights that stay fixed (an all-zero model reproduces the reference's NN-free DE, :1-33);
    `params` supplies everything else and is updated in place.  `comm(buf)` sums a float32 buffer over ranks (multi-GPU).
    Returns (p[5], history of total losses)."""
    if comm is not None and model.ctx.comm_info[0] > 1:
        raise ValueError("the context already carries a communicator (results are reduced inside libopz); do not pass `comm`")
    p0 = np.array([params.nu0, params.nu_m, params.dRi, params.Ric, params.Pr], dtype=np.float64)
    if np.any(p0 <= 0):
        raise ValueError("the initial parameters must be positive (they are the scalings)")
    scaled = np.ones(5, dtype=np.float64)
    history = []
    for opt in optimizers:
        for it in range(maxiters):
            p = scaled * p0
            params.nu0, params.nu_m, params.dRi, params.Ric, params.Pr = (float(v) for v in p)
            total, comps, _, gm = loss_grad_mpp(model, params, x0, bc, target, loss_w, timestepper.scheme, timestepper.dt,
                                                n_steps, save_every, B_total=B_total, precision=timestepper.precision)
            buf = np.concatenate([[total], comps, gm]).astype(np.float32)
            if comm is not None:
                comm(buf)
            total, comps, gm = float(buf[0]), buf[1:7], buf[7:].astype(np.float64)
            history.append(total)
            if callback is not None:
                callback(p.copy(), total, comps, it)
            g_scaled = gm * p0                      # d loss / d (p / p0)
            scaled = np.clip(np.asarray(opt.apply(scaled, g_scaled), dtype=np.float64), 0.0, 10.0)
    p = scaled * p0
    params.nu0, params.nu_m, params.dRi, params.Ric, params.Pr = (float(v) for v in p)
    return p, history


def predict(𝒱, model: NDEModel, which_net: int, *, scaled: bool = False):
    """src/predict.jl:12-34: apply one net to every column of 𝒱.uvT_scaled (3Nz x Nt) in one launch.
    Returns (predictions (Nz-1) x Nt, truth)."""
    x = np.asarray(𝒱.uvT_scaled, dtype=np.float32).T
    y = mlp_forward(model, which_net, x).T
    if scaled:
        return y, 𝒱.scaled
    return 𝒱.unscale_fn(y), 𝒱.coarse
