"""CPU oracle for the NDE column-model hot path (TEST INFRASTRUCTURE, not product).

This is a NumPy restatement of the reference's algorithm, written from SURVEY.md §9 and
the reference sources cited per function.  It is the checker that the CUDA path is
compared against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product path never does.

PARITY PIN STATUS: **parity unpinned**.  The reference (Julia; Flux 0.11.6, OrdinaryDiffEq
5.55.1, DiffEqSensitivity 6.45.0 un-vendored) cannot run in this container, and its own
tests hold no golden vectors for the NDE path.  What IS pinned here:
  * the helper identities the reference's tests do hold (feature-scaling moments / round
    trip: test/test_feature_scaling.jl:7-12; loss-scaling fraction identities:
    wind_mixing/test/test_training_scaling.jl:17-19; Dᶜ/Dᶠ docstring contracts),
  * a float64 twin (noise floor), analytic limits, dt-convergence order,
  * a PyTorch-autograd twin + finite differences for gradients (oracle/nde_oracle_torch.py).
The time integrator is fixed-step RK (north_star replaces ROCK4 by fixed-dt RK), so
bit-level agreement with OrdinaryDiffEq is not a goal.

All arrays are batched over columns: state x has shape [B, 3*Nz] (C order), which is the
same memory as Julia's column-major 3Nz x B.  Index 0 of each Nz block is the BOTTOM cell.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------------------
# flags (mirrored bit-for-bit by include/opz.h)
# ----------------------------------------------------------------------------------------
FLAG_MPP = 1 << 0               # modified Pacanowski-Philander diffusivity on
FLAG_CA = 1 << 1                # convective adjustment on (inference path only)
FLAG_CA_SELECT_DUDZ = 1 << 2    # reference quirk: CA predicate on du/dz (tp.jl:120) instead of dT/dz
FLAG_BC_MINUS_SCALE0 = 1 << 3   # boundary flux = bc - scale(0)   (tp.jl:82-93; zero_weights training)
FLAG_DIURNAL = 1 << 4           # time-dependent top heat flux    (data_containers.jl:131-156)
FLAG_DIURNAL_MINUS_SCALE0 = 1 << 5  # physically consistent diurnal offset (the reference omits it: tp.jl:143)
FLAG_RI_EPS = 1 << 6            # add eps to each gradient inside Ri (NDE_training.jl:115-119)
FLAG_SMOOTH_NN = 1 << 7         # 3-point box filter on NN output (NDE_training.jl:98-102)
FLAG_SMOOTH_RI = 1 << 8         # 3-point box filter on Ri (NDE_training.jl:121-123)

SCHEME_EULER, SCHEME_RK2, SCHEME_RK4, SCHEME_IMEX = 0, 1, 2, 3
STAGES = {SCHEME_EULER: 1, SCHEME_RK2: 2, SCHEME_RK4: 4, SCHEME_IMEX: 1}

ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_MISH, ACT_SWISH, ACT_LEAKYRELU = 0, 1, 2, 3, 4, 5


# ----------------------------------------------------------------------------------------
# operators  (src/differentiation_operators.jl:6-29)
# ----------------------------------------------------------------------------------------
def D_cell(N: int, delta: float) -> np.ndarray:
    """Dᶜ(N, Δ): N x (N+1) face->cell difference matrix (differentiation_operators.jl:6-14)."""
    D = np.zeros((N, N + 1))
    for k in range(N):
        D[k, k] = -1.0
        D[k, k + 1] = 1.0
    return (1.0 / delta) * D


def D_face(N: int, delta: float) -> np.ndarray:
    """Dᶠ(N, Δ): (N+1) x N cell->face difference matrix, first/last rows zero (:21-29)."""
    D = np.zeros((N + 1, N))
    for k in range(1, N):
        D[k, k - 1] = -1.0
        D[k, k] = 1.0
    return (1.0 / delta) * D


def face_gradient(c: np.ndarray, Nz: int) -> np.ndarray:
    """Stencil form of Dᶠ(Nz, 1/Nz) @ c for c[..., Nz] -> [..., Nz+1]."""
    g = np.zeros(c.shape[:-1] + (Nz + 1,), dtype=c.dtype)
    g[..., 1:Nz] = (c[..., 1:] - c[..., :-1]) * c.dtype.type(Nz)
    return g


def cell_divergence(F: np.ndarray, Nz: int) -> np.ndarray:
    """Stencil form of Dᶜ(Nz, 1/Nz) @ F for F[..., Nz+1] -> [..., Nz]."""
    return (F[..., 1:] - F[..., :-1]) * F.dtype.type(Nz)


def smoothing_filter(N: int, filter_width: int) -> np.ndarray:
    """Dense edge-shrinking box filter (wind_mixing/src/filtering_operators.jl:1-15)."""
    assert N >= filter_width and filter_width % 2 == 1
    filt = np.zeros((N, N), dtype=np.float32)
    half = (filter_width - 1) // 2
    for i in range(1, half + 1):
        filt[i - 1, 0:half + i] = 1.0 / (half + i)
        filt[N - i, N - (half + i):N] = 1.0 / (half + i)
    for i in range(half + 1, N - half + 1):
        filt[i - 1, i - 1 - half:i + half] = 1.0 / filter_width
    return filt


# ----------------------------------------------------------------------------------------
# scalers  (src/DataWrangling/feature_scaling.jl:7-23,53-54)
# ----------------------------------------------------------------------------------------
@dataclasses.dataclass
class ZeroMeanUnitVarianceScaling:
    mu: float
    sigma: float

    @classmethod
    def from_data(cls, data) -> "ZeroMeanUnitVarianceScaling":
        data = np.asarray(data)
        # Julia's std() is the corrected (n-1) sample standard deviation.
        return cls(float(np.mean(data)), float(np.std(data, ddof=1)))

    def __call__(self, x):
        return (x - self.mu) / self.sigma

    def scale(self, x):
        return (x - self.mu) / self.sigma

    def unscale(self, y):
        return self.sigma * y + self.mu

    def inv(self) -> Callable:
        return self.unscale


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
@dataclasses.dataclass
class Params:
    """constants + scalings + conditions of prepare_parameters_NDE_training (NDE_training.jl:1-44)."""
    Nz: int = 32
    H: float = 256.0
    tau: float = 172800.0
    f: float = 1e-4
    g: float = 9.80665
    alpha: float = 2e-4
    nu0: float = 1e-4
    nu_m: float = 1e-1
    Ric: float = 0.25
    dRi: float = 1e-1
    Pr: float = 1.0
    kappa: float = 1.0
    eps: float = 1e-7
    # order: u, v, T, uw, vw, wT  (synthetic LES-like magnitudes)
    mu: Sequence[float] = (0.0, 0.0, 18.7, -5e-5, 0.0, 7e-6)
    sigma: Sequence[float] = (0.05, 0.05, 0.8, 1e-4, 1e-4, 5e-6)
    flags: int = FLAG_MPP | FLAG_CA | FLAG_BC_MINUS_SCALE0
    diurnal_period: float = 86400.0

    def scalings(self):
        names = ("u", "v", "T", "uw", "vw", "wT")
        return {n: ZeroMeanUnitVarianceScaling(self.mu[i], self.sigma[i]) for i, n in enumerate(names)}


# ----------------------------------------------------------------------------------------
# MLP  (Flux 0.11.6 Chain/Dense + Flux.destructure; NNlib 0.7.20 activations, restated)
# ----------------------------------------------------------------------------------------
def n_params(sizes: Sequence[int]) -> int:
    return sum(sizes[i] * sizes[i + 1] + sizes[i + 1] for i in range(len(sizes) - 1))


def unpack_weights(flat: np.ndarray, sizes: Sequence[int]):
    """Flux.destructure layout per net: [vec(W1) column-major (out x in); b1; vec(W2); b2; ...]
    (NDE_training.jl:11-13; construct_NN.jl:32-34).  Returns [(W[out,in], b[out]), ...]."""
    out, off = [], 0
    for i in range(len(sizes) - 1):
        nin, nout = sizes[i], sizes[i + 1]
        W = flat[off:off + nin * nout].reshape(nin, nout).T  # column-major out x in
        off += nin * nout
        b = flat[off:off + nout]
        off += nout
        out.append((W, b))
    assert off == flat.size
    return out


def glorot_uniform_flat(sizes: Sequence[int], rng: np.random.Generator, dtype=np.float32) -> np.ndarray:
    """Flux default init restated: W ~ U(+-sqrt(6/(in+out))), b = 0, in destructure layout."""
    parts = []
    for i in range(len(sizes) - 1):
        nin, nout = sizes[i], sizes[i + 1]
        lim = math.sqrt(6.0 / (nin + nout))
        W = rng.uniform(-lim, lim, size=(nout, nin)).astype(dtype)
        parts.append(W.T.reshape(-1))  # column-major vec of out x in
        parts.append(np.zeros(nout, dtype=dtype))
    return np.concatenate(parts)


def activation(z: np.ndarray, act: int) -> np.ndarray:
    if act == ACT_IDENTITY:
        return z
    if act == ACT_RELU:
        return np.maximum(z, z.dtype.type(0))
    if act == ACT_TANH:
        return np.tanh(z)
    if act == ACT_MISH:  # x * tanh(softplus(x))
        sp = np.logaddexp(z, z.dtype.type(0))
        return z * np.tanh(sp)
    if act == ACT_SWISH:  # x * sigmoid(x)
        return z / (z.dtype.type(1) + np.exp(-z))
    if act == ACT_LEAKYRELU:
        return np.maximum(z.dtype.type(0.01) * z, z)
    raise ValueError(act)


def mlp_forward(x: np.ndarray, flat: np.ndarray, sizes: Sequence[int], act: int,
                quant: Optional[Callable] = None, net: int = 0) -> np.ndarray:
    """y = W3 act(W2 act(W1 x + b1) + b2) + b3, batched: x[B, in] -> [B, out].
    `quant` (optional) rounds GEMM operands, emulating a reduced-precision tensor path: a plain callable rounds every
    operand; a TensorPathEmulation restates the product kernel's temporally dithered roundings."""
    layers = unpack_weights(flat, sizes)
    h = x
    emu = isinstance(quant, TensorPathEmulation)
    for li, (W, b) in enumerate(layers):
        last = li == len(layers) - 1
        if isinstance(quant, SplitOperandEmulation):
            z = quant.matmul(h, W) + b
        elif emu:
            z = quant.activations(h, li) @ quant.weights(W, net, li).T + (b if last else quant.bias(b, net, li))
        elif quant is not None:
            z = quant(h) @ quant(W).T + b
        else:
            z = h @ W.T + b
        h = activation(z, act) if not last else z
    return h


def quant_bf16(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 -> float32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def quant_fp16(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> IEEE half (saturating at +-65504) -> float32."""
    a = np.clip(np.ascontiguousarray(a, dtype=np.float32), -65504.0, 65504.0)
    return a.astype(np.float16).astype(np.float32)


def quant_tf32(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u >> 13) & 1) + np.uint32(0x0FFF)
    return ((u + r) & np.uint32(0xFFFFE000)).view(np.float32)


class SplitOperandEmulation:
    """CPU restatement of the operand handling of the layered tensor path (csrc/opz_gemm.cu): every operand rounded to nearest
    in the 16-bit format (mode "plain"), the weights additionally split as hi + lo ("x2": h_hi W_hi + h_hi W_lo), or both
    operands split ("x3": h_hi W_hi + h_hi W_lo + h_lo W_hi), products accumulated in FP32."""

    def __init__(self, fmt: str = "bf16", mode: str = "x3"):
        assert fmt in ("fp16", "bf16") and mode in ("plain", "x2", "x3")
        self.q = quant_fp16 if fmt == "fp16" else quant_bf16
        self.mode = mode

    def matmul(self, h, W):
        h = np.ascontiguousarray(h, dtype=np.float32)
        W = np.ascontiguousarray(W, dtype=np.float32)
        hh, wh = self.q(h), self.q(W)
        z = hh @ wh.T
        if self.mode != "plain":
            z = z + hh @ self.q(W - wh).T
            if self.mode == "x3":
                z = z + self.q(h - hh) @ wh.T
        return z


class TensorPathEmulation:
    """CPU restatement of the operand roundings of the tcgen05 rollout kernel (csrc/opz_tc2.cu), for the parity tests that
    separate "the kernel computes what it is meant to" from "16-bit operands differ from FP32":
      * weights and hidden biases: version (n + s) mod n_tapes of the temporally dithered 16-bit rounding (quant16 /
        dither_key of opz_tc2.cu) for the right-hand side of step n, stage s;
      * the state fed to the first layers: magnitude bits + Weyl threshold of the right-hand side's global index, truncated;
      * hidden activations: round to nearest (IEEE half saturating).
    `stages` = right-hand sides per time step of the scheme; `n0` = global index of the first step (t0 / dt).
    predict_flux() advances the right-hand-side counter after the three nets."""

    def __init__(self, fmt: str = "fp16", stages: int = 4, n_tapes: int = 16, n0: int = 0, dither_x: bool = True):
        assert fmt in ("fp16", "bf16")
        self.fp16 = fmt == "fp16"
        self.stages, self.n_tapes, self.n0, self.dither_x = stages, n_tapes, n0, dither_x
        self.r = 0
        self._cache = {}

    # 16-bit grid helpers on float32 arrays
    def _rn(self, a):
        return quant_fp16(a) if self.fp16 else quant_bf16(a)

    def _neighbours(self, W):
        W = np.ascontiguousarray(W, dtype=np.float32)
        if self.fp16:
            q = np.clip(W, -65504.0, 65504.0).astype(np.float16)
            lo = np.where(q.astype(np.float32) > W, np.nextafter(q, np.float16(-np.inf)), q).astype(np.float16)
            hi = np.nextafter(lo, np.float16(np.inf)).astype(np.float16)
            return q.astype(np.float32), lo.astype(np.float32), hi.astype(np.float32)
        q = quant_bf16(W)
        u = q.view(np.uint32)
        below = np.where(u == 0, np.uint32(0x80010000), np.where(u & np.uint32(0x80000000), u + np.uint32(0x10000), u - np.uint32(0x10000)))
        lo_u = np.where(q > W, below, u).astype(np.uint32)
        above = np.where(lo_u == np.uint32(0x80000000), np.uint32(0x00010000),
                         np.where(lo_u & np.uint32(0x80000000), lo_u - np.uint32(0x10000), lo_u + np.uint32(0x10000))).astype(np.uint32)
        return q, lo_u.view(np.float32), above.view(np.float32)

    def _dither(self, W, key, tape):
        q, lo, hi = self._neighbours(W)
        if self.n_tapes <= 1:
            return q
        with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
            frac = (W.astype(np.float64) - lo.astype(np.float64)) / (hi.astype(np.float64) - lo.astype(np.float64))
        bits = int(np.log2(self.n_tapes))
        rot = (key >> np.uint32(13)) & np.uint32(self.n_tapes - 1)
        idx = (np.uint32(tape) + rot) & np.uint32(self.n_tapes - 1)
        rev = np.zeros_like(idx)
        for b in range(bits):
            rev |= ((idx >> np.uint32(b)) & np.uint32(1)) << np.uint32(bits - 1 - b)
        thr = (rev.astype(np.float64) + 0.5) / self.n_tapes
        ok = np.isfinite(W) & (hi > lo) & np.isfinite(hi) & (np.abs(W) < (65504.0 if self.fp16 else np.inf))
        return np.where(ok, np.where(frac > thr, hi, lo), q).astype(np.float32)

    def _tape(self):
        n, s = self.n0 + self.r // self.stages, self.r % self.stages
        return (n + s) % self.n_tapes

    @staticmethod
    def _key(net, layer, k, n):
        with np.errstate(over="ignore"):
            return (k.astype(np.uint32) * np.uint32(2654435761)) ^ (n.astype(np.uint32) * np.uint32(2246822519)) ^ \
                   np.uint32(((net * 8 + layer) * 3266489917) & 0xFFFFFFFF)

    def weights(self, W, net, layer):
        """W is out x in (Flux); returns the version of the current right-hand side."""
        tape = self._tape()
        ck = (net, layer, tape, W.shape, float(W.flat[0]), float(W.flat[-1]))
        if ck in self._cache:
            return self._cache[ck]
        nout, nin = W.shape
        n, k = np.meshgrid(np.arange(nout, dtype=np.uint32), np.arange(nin, dtype=np.uint32), indexing="ij")
        out = self._dither(W, self._key(net, layer, k, n), tape)
        self._cache[ck] = out
        return out

    def bias(self, b, net, layer):
        tape = self._tape()
        ck = ("b", net, layer, tape, b.shape, float(b.flat[0]) if b.size else 0.0)
        if ck in self._cache:
            return self._cache[ck]
        n = np.arange(b.shape[0], dtype=np.uint32)
        out = self._dither(b, self._key(net, layer, np.full_like(n, 0xFFFF), n), tape)
        self._cache[ck] = out
        return out

    def activations(self, h, layer):
        h = np.ascontiguousarray(h, dtype=np.float32)
        if layer > 0:
            return self._rn(h)
        drop = 13 if self.fp16 else 16
        t = np.uint32(((self.n0 * self.stages + self.r) * 5061) & 0xFFFFFFFF)
        idx = np.arange(h.shape[-1], dtype=np.uint32) & np.uint32(7)
        thr = ((t + idx * np.uint32(1021)) & np.uint32(8191)) << np.uint32(drop - 13)
        if not self.dither_x:
            thr = np.full_like(idx, 1 << (drop - 1))
        with np.errstate(over="ignore"):
            u = (h.view(np.uint32) + thr) & np.uint32(~((1 << drop) - 1) & 0xFFFFFFFF)
        v = u.view(np.float32)
        return np.clip(v, -65504.0, 65504.0) if self.fp16 else v

    def advance(self):
        self.r += 1


@dataclasses.dataclass
class Model:
    """Three nets uw, vw, wT of identical architecture on the shared input [u;v;T]."""
    sizes: Sequence[int]
    act: int
    w_uw: np.ndarray
    w_vw: np.ndarray
    w_wT: np.ndarray

    @property
    def P(self) -> int:
        return n_params(self.sizes)

    def flat(self) -> np.ndarray:
        return np.concatenate([self.w_uw, self.w_vw, self.w_wT])

    def astype(self, dt) -> "Model":
        return Model(self.sizes, self.act, self.w_uw.astype(dt), self.w_vw.astype(dt), self.w_wT.astype(dt))

    @classmethod
    def from_flat(cls, flat, sizes, act) -> "Model":
        P = n_params(sizes)
        return cls(sizes, act, flat[:P].copy(), flat[P:2 * P].copy(), flat[2 * P:3 * P].copy())


def make_model(sizes, act, seed=0, out_scale=1.0, dtype=np.float32, bias_scale=0.0) -> Model:
    """Seeded Glorot-uniform nets; `out_scale` multiplies the LAST layer's W (b=0), the knob
    used to keep the synthetic random-weight NDE well conditioned.  `bias_scale` > 0
    replaces Flux's zero biases by U(+-bias_scale) (hidden layers) and U(+-bias_scale*out_scale) (output
    layer), as a trained model would have them; drawn AFTER the weights so that bias_scale=0 models
    are unchanged."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nets = []
    for _ in range(3):
        flat = glorot_uniform_flat(sizes, rng, dtype)
        if out_scale != 1.0:
            nin, nout = sizes[-2], sizes[-1]
            off = n_params(sizes) - nout - nin * nout
            flat[off:off + nin * nout] *= dtype(out_scale)
        nets.append(flat)
    if bias_scale > 0.0:
        brng = np.random.Generator(np.random.PCG64(seed + 7919))
        for flat in nets:
            off = 0
            for i in range(len(sizes) - 1):
                nin, nout = sizes[i], sizes[i + 1]
                off += nin * nout
                sc = bias_scale * (out_scale if i == len(sizes) - 2 else 1.0)
                flat[off:off + nout] = brng.uniform(-sc, sc, size=nout).astype(dtype)
                off += nout
    return Model(tuple(sizes), act, *nets)


# ----------------------------------------------------------------------------------------
# right-hand side
# ----------------------------------------------------------------------------------------
def diurnal_wT_top(Q, t_seconds, p: Params, dt):
    """wT_flux(Q, t) = Q sin(2 pi t / 86400) / (alpha g)  [K m/s] (data_containers.jl:135)."""
    return (Q * np.sin(dt(2.0 * math.pi / p.diurnal_period) * t_seconds) / (dt(p.alpha) * dt(p.g))).astype(dt)


def _box3(a: np.ndarray) -> np.ndarray:
    """Edge-shrinking 3-point box filter along the last axis == smoothing_filter(N,3) @ a."""
    out = np.empty_like(a)
    third = a.dtype.type(1.0 / 3.0)
    half = a.dtype.type(0.5)
    out[..., 1:-1] = (a[..., :-2] + a[..., 1:-1] + a[..., 2:]) * third
    out[..., 0] = (a[..., 0] + a[..., 1]) * half
    out[..., -1] = (a[..., -2] + a[..., -1]) * half
    return out


def predict_flux(x: np.ndarray, bc: np.ndarray, t, model: Model, p: Params,
                 quant: Optional[Callable] = None, return_diag: bool = False):
    """Total scaled fluxes on the Nz+1 faces.

    Follows predict_flux! (training_postprocessing.jl:105-129) and predict_flux
    (NDE_training.jl:83-147); the flag set selects the flavour.
    x: [B, 3Nz] scaled; bc: [B, 8] = uw_b, uw_t, vw_b, vw_t, wT_b, wT_t (scaled), Q_diurnal, unused.
    Returns uw, vw, wT each [B, Nz+1].
    """
    dt = x.dtype.type
    Nz = p.Nz
    B = x.shape[0]
    mu_uw, mu_vw, mu_wT = p.mu[3], p.mu[4], p.mu[5]
    s_u, s_v, s_T, s_uw, s_vw, s_wT = (dt(s) for s in p.sigma)
    u, v, T = x[:, :Nz], x[:, Nz:2 * Nz], x[:, 2 * Nz:]

    nn = [mlp_forward(x, w.astype(x.dtype, copy=False), model.sizes, model.act, quant, net)
          for net, w in enumerate((model.w_uw, model.w_vw, model.w_wT))]
    if isinstance(quant, TensorPathEmulation):
        quant.advance()
    if p.flags & FLAG_SMOOTH_NN:
        nn = [_box3(a) for a in nn]

    # boundary values; scale(0) = -mu/sigma
    off = [dt(-mu_uw) / s_uw, dt(-mu_vw) / s_vw, dt(-mu_wT) / s_wT]
    sub = dt(1) if (p.flags & FLAG_BC_MINUS_SCALE0) else dt(0)
    F = []
    for i in range(3):
        Fi = np.empty((B, Nz + 1), dtype=x.dtype)
        Fi[:, 0] = bc[:, 2 * i] - sub * off[i]
        Fi[:, Nz] = bc[:, 2 * i + 1] - sub * off[i]
        Fi[:, 1:Nz] = nn[i]
        F.append(Fi)
    if p.flags & FLAG_DIURNAL:
        # wT_top = scalings.wT(wT_top_function(t*tau))  (NDE_training.jl:73; tp.jl:142-144)
        phys = diurnal_wT_top(bc[:, 6].astype(x.dtype), dt(t) * dt(p.tau), p, dt)
        top = (phys - dt(mu_wT)) / s_wT
        if p.flags & FLAG_DIURNAL_MINUS_SCALE0:
            top = top - off[2]
        F[2][:, Nz] = top

    diag = {}
    if p.flags & FLAG_MPP:
        gu, gv, gT = face_gradient(u, Nz), face_gradient(v, Nz), face_gradient(T, Nz)
        e = dt(p.eps) if (p.flags & FLAG_RI_EPS) else dt(0)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            # local_richardson (NDE_training.jl:46-52): Bz = H*g*alpha*sigma_T*dTdz ; S2 = (su*du)^2+(sv*dv)^2
            Bz = dt(p.H) * dt(p.g) * dt(p.alpha) * s_T * (gT + e)
            S2 = (s_u * (gu + e)) ** 2 + (s_v * (gv + e)) ** 2
            Ri = Bz / S2
            if p.flags & FLAG_SMOOTH_RI:
                Ri = _box3(Ri)
            # nu = nu0 + nu_m * tanh_step((Ri - Ric)/dRi)   (tp.jl:116 ; tanh_step :54)
            nu = dt(p.nu0) + dt(p.nu_m) * ((dt(1) - np.tanh((Ri - dt(p.Ric)) / dt(p.dRi))) / dt(2))
        if p.flags & FLAG_CA:
            sel = gu if (p.flags & FLAG_CA_SELECT_DUDZ) else gT
            nuT = np.where(sel > 0, nu / dt(p.Pr), dt(p.kappa))  # tp.jl:118-121
        else:
            nuT = nu / dt(p.Pr)
        Hd = dt(p.H)
        sl = slice(1, Nz)
        F[0][:, sl] -= s_u / s_uw / Hd * nu[:, sl] * gu[:, sl]
        F[1][:, sl] -= s_v / s_vw / Hd * nu[:, sl] * gv[:, sl]
        F[2][:, sl] -= s_T / s_wT / Hd * nuT[:, sl] * gT[:, sl]
        diag = dict(Ri=Ri, nu=nu, nuT=nuT, gu=gu, gv=gv, gT=gT)
    if return_diag:
        return F[0], F[1], F[2], diag
    return F[0], F[1], F[2]


def rhs(x: np.ndarray, bc: np.ndarray, t, model: Model, p: Params, quant=None) -> np.ndarray:
    """d(x)/d(t/tau) (predict_NDE NDE_training.jl:149-165; NDE! tp.jl:131-153)."""
    dt = x.dtype.type
    Nz = p.Nz
    mu_u, mu_v = dt(p.mu[0]), dt(p.mu[1])
    s_u, s_v, s_T, s_uw, s_vw, s_wT = (dt(s) for s in p.sigma)
    u, v = x[:, :Nz], x[:, Nz:2 * Nz]
    uw, vw, wT = predict_flux(x, bc, t, model, p, quant)
    A = -dt(p.tau) / dt(p.H)
    ft = dt(p.f) * dt(p.tau)
    out = np.empty_like(x)
    out[:, :Nz] = A * s_uw / s_u * cell_divergence(uw, Nz) + ft / s_u * (s_v * v + mu_v)
    out[:, Nz:2 * Nz] = A * s_vw / s_v * cell_divergence(vw, Nz) - ft / s_v * (s_u * u + mu_u)
    out[:, 2 * Nz:] = A * s_wT / s_T * cell_divergence(wT, Nz)
    return out


# ----------------------------------------------------------------------------------------
# fixed-step time stepping (replaces OrdinaryDiffEq ROCK4 per north_star)
# ----------------------------------------------------------------------------------------
def step(x, bc, t, h, model, p, scheme, quant=None):
    dt = x.dtype.type
    h = dt(h)
    if scheme == SCHEME_EULER:
        return x + h * rhs(x, bc, t, model, p, quant)
    if scheme == SCHEME_RK2:  # Heun
        k1 = rhs(x, bc, t, model, p, quant)
        k2 = rhs(x + h * k1, bc, t + h, model, p, quant)
        return x + (h * dt(0.5)) * (k1 + k2)
    if scheme == SCHEME_RK4:
        k1 = rhs(x, bc, t, model, p, quant)
        k2 = rhs(x + (h * dt(0.5)) * k1, bc, t + h * dt(0.5), model, p, quant)
        k3 = rhs(x + (h * dt(0.5)) * k2, bc, t + h * dt(0.5), model, p, quant)
        k4 = rhs(x + h * k3, bc, t + h, model, p, quant)
        return x + (h / dt(6)) * (k1 + dt(2) * k2 + dt(2) * k3 + k4)
    if scheme == SCHEME_IMEX:
        return step_imex(x, bc, t, h, model, p, quant)
    raise ValueError(scheme)


def step_imex(x, bc, t, h, model, p, quant=None):
    """Split step of the reference's production (Oceananigans) path, NDE_oceananigans.jl:61-101: an explicit Euler step
    with everything but the vertical diffusion (NN fluxes, boundary fluxes, Coriolis), then a backward-Euler solve of the
    modified Pacanowski-Philander / convective-adjustment diffusion, one tridiagonal system per variable,
        (1 + D_c + D_{c+1}) phi_c - D_c phi_{c-1} - D_{c+1} phi_{c+1} = phi*_c ,   D_f = dt nu_f / dz^2  (0 on the two
    boundary faces), with the diffusivities evaluated on the state after the explicit part (:40-58 computes them from
    the model state at that point).  In the scaled variables D_f = h (tau / H^2) Nz^2 nu_f.  The reference's
    `T'[1] = T_bottom` pin (:95) is NOT applied (it overrides the no-flux bottom boundary of the column model)."""
    dt = x.dtype.type
    Nz = p.Nz
    h = dt(h)
    p_ex = dataclasses.replace(p, flags=p.flags & ~(FLAG_MPP | FLAG_CA | FLAG_SMOOTH_RI))
    xs = x + h * rhs(x, bc, t, model, p_ex, quant)
    if not (p.flags & FLAG_MPP):
        return xs
    _, _, _, diag = predict_flux(xs, bc, t, model, p, quant, return_diag=True)
    coef = h * (dt(p.tau) / dt(p.H) / dt(p.H)) * dt(Nz) * dt(Nz)
    out = np.empty_like(xs)
    for v, nu in enumerate((diag["nu"], diag["nu"], diag["nuT"])):
        D = np.zeros((x.shape[0], Nz + 1), dtype=x.dtype)
        D[:, 1:Nz] = coef * nu[:, 1:Nz].astype(x.dtype)
        phi = xs[:, v * Nz:(v + 1) * Nz]
        # Thomas algorithm, bottom (c = 0) to top
        cp = np.empty_like(phi)
        dp = np.empty_like(phi)
        b0 = dt(1) + D[:, 0] + D[:, 1]
        cp[:, 0] = -D[:, 1] / b0
        dp[:, 0] = phi[:, 0] / b0
        for c in range(1, Nz):
            a_c = -D[:, c]
            b_c = dt(1) + D[:, c] + D[:, c + 1]
            den = b_c - a_c * cp[:, c - 1]
            cp[:, c] = -D[:, c + 1] / den
            dp[:, c] = (phi[:, c] - a_c * dp[:, c - 1]) / den
        sol = np.empty_like(phi)
        sol[:, Nz - 1] = dp[:, Nz - 1]
        for c in range(Nz - 2, -1, -1):
            sol[:, c] = dp[:, c] - cp[:, c] * sol[:, c + 1]
        out[:, v * Nz:(v + 1) * Nz] = sol
    return out


def rollout(x0, bc, model, p, scheme, dt_nd, n_steps, save_every, t0=0.0, quant=None):
    """solve_NDE_mutating with a fixed-step scheme: returns [B, n_save, 3Nz] where
    n_save = n_steps // save_every + 1 and snapshot 0 is the initial state (saveat=ts)."""
    x = np.array(x0, copy=True)
    ft = x.dtype.type
    n_save = n_steps // save_every + 1
    out = np.empty((x.shape[0], n_save, x.shape[1]), dtype=x.dtype)
    out[:, 0] = x
    for n in range(n_steps):
        t = ft(t0) + ft(n) * ft(dt_nd)
        x = step(x, bc, t, dt_nd, model, p, scheme, quant)
        if (n + 1) % save_every == 0:
            out[:, (n + 1) // save_every] = x
    return out


# ----------------------------------------------------------------------------------------
# loss  (wind_mixing/src/loss.jl:1-42; NDE_training.jl:290-323)
# ----------------------------------------------------------------------------------------
def mse(a, b):
    return np.mean((a - b) ** 2)


def loss_components(sol, target, Nz, with_gradients=True):
    """sol, target: [B, n_save, 3Nz].  Returns (u, v, T, dudz, dvdz, dTdz) unweighted losses:
    mean over simulations of Flux.mse over an Nz x Nt (or (Nz+1) x Nt) block."""
    B = sol.shape[0]
    comps = []
    for i in range(3):
        a, b = sol[:, :, i * Nz:(i + 1) * Nz], target[:, :, i * Nz:(i + 1) * Nz]
        comps.append(np.mean([mse(a[s], b[s]) for s in range(B)]))
    for i in range(3):
        if with_gradients:
            a = face_gradient(sol[:, :, i * Nz:(i + 1) * Nz], Nz)
            b = face_gradient(target[:, :, i * Nz:(i + 1) * Nz], Nz)
            comps.append(np.mean([mse(a[s], b[s]) for s in range(B)]))
        else:
            comps.append(sol.dtype.type(0))
    return tuple(comps)


def calculate_loss_scalings(losses, fractions, train_gradient):
    """loss.jl:11-31.  losses: dict u,v,T,dudz,dvdz,dTdz ; fractions: dict T, dTdz, profile."""
    velocity_scaling = (1 - fractions["T"]) / fractions["T"] * losses["T"] / (losses["u"] + losses["v"])
    profile_loss = velocity_scaling * (losses["u"] + losses["v"]) + losses["T"]
    if train_gradient:
        vgs = (1 - fractions["dTdz"]) / fractions["dTdz"] * losses["dTdz"] / (losses["dudz"] + losses["dvdz"])
        gradient_loss = vgs * (losses["dudz"] + losses["dvdz"]) + losses["dTdz"]
        tgs = (1 - fractions["profile"]) / fractions["profile"] * profile_loss / gradient_loss
    else:
        vgs = 0
        tgs = 0
    return dict(u=velocity_scaling, v=velocity_scaling, T=1, dudz=tgs * vgs, dvdz=tgs * vgs, dTdz=tgs)


def total_loss(sol, target, Nz, w6):
    c = loss_components(sol, target, Nz, with_gradients=any(w != 0 for w in w6[3:]))
    return sum(float(w) * float(ci) for w, ci in zip(w6, c)), c


# ----------------------------------------------------------------------------------------
# synthetic inputs  
# ----------------------------------------------------------------------------------------
def synthetic_columns(B: int, p: Params, dtype=np.float32, seed: int = 0, diurnal: bool = False):
    """IC: u=v=0, T(z)=20+0.01 z at cell centres; forcing swept on a (Q_u, Q_b) grid:
    uw_top = -Q_u, Q_u in [1.5e-4, 1e-3]; wT_top = Q_b/(alpha g), Q_b in [-5e-8, 1.2e-7];
    bottom fluxes 0, vw_top 0.  Returns x0[B,3Nz] (scaled) and bc[B,8] (scaled fluxes)."""
    Nz = p.Nz
    sc = p.scalings()
    dz = p.H / Nz
    z = -p.H + dz * (np.arange(Nz) + 0.5)
    T0 = 20.0 + 0.01 * z
    x0 = np.empty((B, 3 * Nz), dtype=np.float64)
    x0[:, :Nz] = sc["u"](0.0)
    x0[:, Nz:2 * Nz] = sc["v"](0.0)
    x0[:, 2 * Nz:] = sc["T"](T0)[None, :]
    n1 = int(math.ceil(math.sqrt(B)))
    idx = np.arange(B)
    a = (idx % n1) / max(n1 - 1, 1)
    b = (idx // n1) / max((B - 1) // n1, 1)
    Qu = 1.5e-4 + a * (1e-3 - 1.5e-4)
    Qb = -5e-8 + b * (1.2e-7 + 5e-8)
    bc = np.zeros((B, 8), dtype=np.float64)
    bc[:, 0] = sc["uw"](0.0)
    bc[:, 1] = sc["uw"](-Qu)
    bc[:, 2] = sc["vw"](0.0)
    bc[:, 3] = sc["vw"](0.0)
    bc[:, 4] = sc["wT"](0.0)
    bc[:, 5] = sc["wT"](Qb / (p.alpha * p.g))
    bc[:, 6] = np.abs(Qb) if diurnal else 0.0
    return x0.astype(dtype), bc.astype(dtype)


def profile_error(a: np.ndarray, ref: np.ndarray, Nz: int) -> float:
    """Parity norm: per column, variable and snapshot,
    max_k|a-ref| / max_k|ref| ; max over everything.  a, ref: [B, n_save, 3Nz]."""
    worst = 0.0
    for i in range(3):
        d = np.abs(a[..., i * Nz:(i + 1) * Nz].astype(np.float64) - ref[..., i * Nz:(i + 1) * Nz].astype(np.float64))
        den = np.abs(ref[..., i * Nz:(i + 1) * Nz].astype(np.float64)).max(axis=-1)
        num = d.max(axis=-1)
        # u,v start at exactly 0 (scaled): measure those snapshots against the run's amplitude
        floor = np.maximum(den.max(), 1e-30) * 1e-3
        worst = max(worst, float((num / np.maximum(den, floor)).max()))
    return worst
