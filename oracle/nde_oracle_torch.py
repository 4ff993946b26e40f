"""PyTorch-autograd twin of oracle/nde_oracle.py (TEST INFRASTRUCTURE, not product).

Same right-hand side, same fixed-step schemes, same loss, written with differentiable torch ops on
the CPU so that `torch.autograd` yields the EXACT discrete-adjoint (BPTT) gradient of the profile
loss with respect to the flat NN weight vector — the ground truth the CUDA BPTT kernel
(opz_loss_grad) is compared with.  It replaces, as a checker, what the reference obtains from
Zygote through `solve(...; sensealg=InterpolatingAdjoint(autojacvec=ZygoteVJP()))`
(wind_mixing/src/NDE_training.jl:290-333); the reference's own adjoint is continuous and its
solver adaptive, so agreement with Julia to rounding is not a goal (parity unpinned).

Follows: NDE / predict_NDE / predict_flux (NDE_training.jl:56-165), loss (loss.jl:1-42),
loss_NDE / loss_gradient_NDE (NDE_training.jl:290-323).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import nde_oracle as O


def _act(z, act):
    if act == O.ACT_IDENTITY:
        return z
    if act == O.ACT_RELU:
        return torch.relu(z)
    if act == O.ACT_TANH:
        return torch.tanh(z)
    if act == O.ACT_MISH:
        return z * torch.tanh(torch.nn.functional.softplus(z))
    if act == O.ACT_SWISH:
        return z * torch.sigmoid(z)
    if act == O.ACT_LEAKYRELU:
        return torch.maximum(0.01 * z, z)
    raise ValueError(act)


def mlp(x, flat, sizes, act):
    """x[B,in] -> [B,out] with `flat` in Flux.destructure order (NDE_training.jl:11-13)."""
    h, off = x, 0
    L = len(sizes) - 1
    for l in range(L):
        nin, nout = sizes[l], sizes[l + 1]
        W = flat[off:off + nin * nout].reshape(nin, nout)  # element (k_in, n_out)
        off += nin * nout
        b = flat[off:off + nout]
        off += nout
        z = h @ W + b
        h = _act(z, act) if l < L - 1 else z
    return h


def _box3(a):
    third = 1.0 / 3.0
    mid = (a[..., :-2] + a[..., 1:-1] + a[..., 2:]) * third
    lo = ((a[..., 0] + a[..., 1]) * 0.5).unsqueeze(-1)
    hi = ((a[..., -2] + a[..., -1]) * 0.5).unsqueeze(-1)
    return torch.cat([lo, mid, hi], dim=-1)


def rhs(x, bc, t, w, sizes, act, p: O.Params, phys=None, return_fluxes=False):
    """d x / d(t/tau); x[B,3Nz], bc[B,8], w flat [3P] (torch tensors of one dtype).  `phys` (optional tensor [5]) overrides
    (nu0, nu_m, dRi, Ric, Pr) of `p` so that autograd can differentiate with respect to them
    (diffusivity_parameter_optimisation.jl:2)."""
    Nz = p.Nz
    P = O.n_params(sizes)
    dt = x.dtype
    s_u, s_v, s_T, s_uw, s_vw, s_wT = (float(s) for s in p.sigma)
    mu_u, mu_v = float(p.mu[0]), float(p.mu[1])
    u, v, T = x[:, :Nz], x[:, Nz:2 * Nz], x[:, 2 * Nz:]
    nn = [mlp(x, w[i * P:(i + 1) * P], sizes, act) for i in range(3)]
    if p.flags & O.FLAG_SMOOTH_NN:
        nn = [_box3(a) for a in nn]
    sub = 1.0 if (p.flags & O.FLAG_BC_MINUS_SCALE0) else 0.0
    off = [-p.mu[3 + i] / p.sigma[3 + i] for i in range(3)]
    bot = [bc[:, 2 * i] - sub * off[i] for i in range(3)]
    top = [bc[:, 2 * i + 1] - sub * off[i] for i in range(3)]
    if p.flags & O.FLAG_DIURNAL:
        wT_phys = bc[:, 6] * math.sin(2.0 * math.pi / p.diurnal_period * (float(t) * p.tau)) / (p.alpha * p.g)
        tp = (wT_phys - p.mu[5]) / p.sigma[5]
        if p.flags & O.FLAG_DIURNAL_MINUS_SCALE0:
            tp = tp - off[2]
        top[2] = tp
    F = [a for a in nn]
    if p.flags & O.FLAG_MPP:
        g = [(c[:, 1:] - c[:, :-1]) * float(Nz) for c in (u, v, T)]  # interior faces 1..Nz-1
        e = p.eps if (p.flags & O.FLAG_RI_EPS) else 0.0
        if p.flags & O.FLAG_SMOOTH_RI:
            z = torch.zeros_like(g[0][:, :1])
            gf = [torch.cat([z, a, z], dim=1) for a in g]
            Ri = (p.H * p.g * p.alpha * s_T * (gf[2] + e)) / ((s_u * (gf[0] + e)) ** 2 + (s_v * (gf[1] + e)) ** 2)
            Ri = _box3(Ri)[:, 1:-1]
        else:
            Ri = (p.H * p.g * p.alpha * s_T * (g[2] + e)) / ((s_u * (g[0] + e)) ** 2 + (s_v * (g[1] + e)) ** 2)
        nu0, nu_m, dRi, Ric, Pr = (p.nu0, p.nu_m, p.dRi, p.Ric, p.Pr) if phys is None else (phys[0], phys[1], phys[2], phys[3], phys[4])
        nu = nu0 + nu_m * ((1.0 - torch.tanh((Ri - Ric) / dRi)) / 2.0)
        if p.flags & O.FLAG_CA:
            sel = g[0] if (p.flags & O.FLAG_CA_SELECT_DUDZ) else g[2]
            nuT = torch.where(sel > 0, nu / Pr, torch.full_like(nu, p.kappa))
        else:
            nuT = nu / Pr
        F[0] = F[0] - s_u / s_uw / p.H * nu * g[0]
        F[1] = F[1] - s_v / s_vw / p.H * nu * g[1]
        F[2] = F[2] - s_T / s_wT / p.H * nuT * g[2]
    full = [torch.cat([bot[i].unsqueeze(1), F[i], top[i].unsqueeze(1)], dim=1) for i in range(3)]
    if return_fluxes:
        return full
    div = [(f[:, 1:] - f[:, :-1]) * float(Nz) for f in full]
    A = -p.tau / p.H
    ft = p.f * p.tau
    du = A * s_uw / s_u * div[0] + ft / s_u * (s_v * v + mu_v)
    dv = A * s_vw / s_v * div[1] - ft / s_v * (s_u * u + mu_u)
    dT = A * s_wT / s_T * div[2]
    return torch.cat([du, dv, dT], dim=1).to(dt)


def diffusivities(x, p: O.Params, phys=None):
    """nu, nu_T on the interior faces 1..Nz-1 of the state x (the part of rhs() that the split implicit step needs)."""
    Nz = p.Nz
    s_u, s_v, s_T = (float(s) for s in p.sigma[:3])
    u, v, T = x[:, :Nz], x[:, Nz:2 * Nz], x[:, 2 * Nz:]
    g = [(c[:, 1:] - c[:, :-1]) * float(Nz) for c in (u, v, T)]
    e = p.eps if (p.flags & O.FLAG_RI_EPS) else 0.0
    if p.flags & O.FLAG_SMOOTH_RI:
        z = torch.zeros_like(g[0][:, :1])
        gf = [torch.cat([z, a, z], dim=1) for a in g]
        Ri = (p.H * p.g * p.alpha * s_T * (gf[2] + e)) / ((s_u * (gf[0] + e)) ** 2 + (s_v * (gf[1] + e)) ** 2)
        Ri = _box3(Ri)[:, 1:-1]
    else:
        Ri = (p.H * p.g * p.alpha * s_T * (g[2] + e)) / ((s_u * (g[0] + e)) ** 2 + (s_v * (g[1] + e)) ** 2)
    nu0, nu_m, dRi, Ric, Pr = (p.nu0, p.nu_m, p.dRi, p.Ric, p.Pr) if phys is None else (phys[0], phys[1], phys[2], phys[3], phys[4])
    nu = nu0 + nu_m * ((1.0 - torch.tanh((Ri - Ric) / dRi)) / 2.0)
    if p.flags & O.FLAG_CA:
        sel = g[0] if (p.flags & O.FLAG_CA_SELECT_DUDZ) else g[2]
        nuT = torch.where(sel > 0, nu / Pr, torch.full_like(nu, p.kappa))
    else:
        nuT = nu / Pr
    return nu, nuT


def step_imex(x, bc, t, h, w, sizes, act, p, phys=None):
    """Differentiable twin of nde_oracle.step_imex (NDE_oceananigans.jl:61-101): explicit Euler without the vertical
    diffusion, then backward-Euler diffusion with the diffusivities of the updated state (Thomas algorithm per variable)."""
    import dataclasses
    Nz = p.Nz
    p_ex = dataclasses.replace(p, flags=p.flags & ~(O.FLAG_MPP | O.FLAG_CA | O.FLAG_SMOOTH_RI))
    xs = x + h * rhs(x, bc, t, w, sizes, act, p_ex, phys)
    if not (p.flags & O.FLAG_MPP):
        return xs
    nu, nuT = diffusivities(xs, p, phys)
    coef = h * (p.tau / p.H / p.H) * Nz * Nz
    outs = []
    zero = torch.zeros_like(xs[:, :1])
    for v, nn in enumerate((nu, nu, nuT)):
        D = torch.cat([zero, coef * nn, zero], dim=1)           # faces 0..Nz
        phi = xs[:, v * Nz:(v + 1) * Nz]
        cp, dp = [], []
        b0 = 1.0 + D[:, 0] + D[:, 1]
        cp.append(-D[:, 1] / b0)
        dp.append(phi[:, 0] / b0)
        for c in range(1, Nz):
            a_c = -D[:, c]
            den = (1.0 + D[:, c] + D[:, c + 1]) - a_c * cp[c - 1]
            cp.append(-D[:, c + 1] / den)
            dp.append((phi[:, c] - a_c * dp[c - 1]) / den)
        sol = [None] * Nz
        sol[Nz - 1] = dp[Nz - 1]
        for c in range(Nz - 2, -1, -1):
            sol[c] = dp[c] - cp[c] * sol[c + 1]
        outs.append(torch.stack(sol, dim=1))
    return torch.cat(outs, dim=1)


def step(x, bc, t, h, w, sizes, act, p, scheme, phys=None):
    if scheme == O.SCHEME_IMEX:
        return step_imex(x, bc, t, h, w, sizes, act, p, phys)
    if scheme == O.SCHEME_EULER:
        return x + h * rhs(x, bc, t, w, sizes, act, p, phys)
    if scheme == O.SCHEME_RK2:
        k1 = rhs(x, bc, t, w, sizes, act, p, phys)
        k2 = rhs(x + h * k1, bc, t + h, w, sizes, act, p, phys)
        return x + (h * 0.5) * (k1 + k2)
    k1 = rhs(x, bc, t, w, sizes, act, p, phys)
    k2 = rhs(x + (h * 0.5) * k1, bc, t + h * 0.5, w, sizes, act, p, phys)
    k3 = rhs(x + (h * 0.5) * k2, bc, t + h * 0.5, w, sizes, act, p, phys)
    k4 = rhs(x + h * k3, bc, t + h, w, sizes, act, p, phys)
    return x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def rollout(x0, bc, w, sizes, act, p, scheme, dt_nd, n_steps, save_every, t0=0.0, phys=None):
    x = x0
    out = [x]
    for n in range(n_steps):
        x = step(x, bc, t0 + n * dt_nd, dt_nd, w, sizes, act, p, scheme, phys)
        if (n + 1) % save_every == 0:
            out.append(x)
    return torch.stack(out, dim=1)  # [B, n_save, 3Nz]


def loss_components(sol, target, Nz, B_total=None):
    """Six unweighted components (u, v, T, du/dz, dv/dz, dT/dz): mean over simulations of Flux.mse over
    an Nz x Nt (resp. (Nz+1) x Nt with two zero rows) block (loss.jl:1-9; NDE_training.jl:290-317)."""
    B, Nt = sol.shape[0], sol.shape[1]
    Bt = B if B_total is None else B_total
    comps = []
    for i in range(3):
        d = sol[:, :, i * Nz:(i + 1) * Nz] - target[:, :, i * Nz:(i + 1) * Nz]
        comps.append((d ** 2).sum() / (Nz * Nt * Bt))
    for i in range(3):
        a = sol[:, :, i * Nz:(i + 1) * Nz]
        b = target[:, :, i * Nz:(i + 1) * Nz]
        d = ((a[:, :, 1:] - a[:, :, :-1]) - (b[:, :, 1:] - b[:, :, :-1])) * float(Nz)
        comps.append((d ** 2).sum() / ((Nz + 1) * Nt * Bt))
    return comps


def loss_and_grad(x0, bc, model: O.Model, p: O.Params, scheme, dt_nd, n_steps, save_every, target, loss_w, t0=0.0,
                  dtype=torch.float64, B_total=None, with_mpp=False):
    """Returns (total, components[6] (weighted), grad[3P]) as numpy float64; with_mpp=True appends
    d total / d (nu0, nu_m, dRi, Ric, Pr) [5]."""
    w = torch.tensor(np.asarray(model.flat(), dtype=np.float64), dtype=dtype, requires_grad=True)
    x0t = torch.tensor(np.asarray(x0, dtype=np.float64), dtype=dtype)
    bct = torch.tensor(np.asarray(bc, dtype=np.float64), dtype=dtype)
    tg = torch.tensor(np.asarray(target, dtype=np.float64), dtype=dtype)
    phys = torch.tensor([p.nu0, p.nu_m, p.dRi, p.Ric, p.Pr], dtype=dtype, requires_grad=True) if with_mpp else None
    sol = rollout(x0t, bct, w, tuple(model.sizes), model.act, p, scheme, dt_nd, n_steps, save_every, t0, phys)
    comps = loss_components(sol, tg, p.Nz, B_total)
    weighted = [float(loss_w[i]) * comps[i] for i in range(6)]
    total = sum(weighted)
    total.backward()
    out = (float(total.detach()), np.array([float(c.detach()) for c in weighted]), w.grad.detach().numpy().astype(np.float64))
    if with_mpp:
        out = out + (phys.grad.detach().numpy().astype(np.float64),)
    return out


def flux_loss_and_grad(x, bc, model: O.Model, p: O.Params, which_net, flux_target, gradient_scaling, dtype=torch.float64):
    """Supervised pre-training loss of one net (train_NN / NN_loss, NN_training.jl:207-249; predict_uw/vw/wT :25-169 = the
    total flux of that net) and its gradient w.r.t. that net's weights: returns (total, mse_flux, mse_flux_gradient, grad[P]),
    means over the samples."""
    w = torch.tensor(np.asarray(model.flat(), dtype=np.float64), dtype=dtype, requires_grad=True)
    xt = torch.tensor(np.asarray(x, dtype=np.float64), dtype=dtype)
    bct = torch.tensor(np.asarray(bc, dtype=np.float64), dtype=dtype)
    tg = torch.tensor(np.asarray(flux_target, dtype=np.float64), dtype=dtype)
    F = rhs(xt, bct, 0.0, w, tuple(model.sizes), model.act, p, return_fluxes=True)[which_net]
    Nz = p.Nz
    l1 = ((F - tg) ** 2).mean(dim=1)
    dF = (F[:, 1:] - F[:, :-1]) * float(Nz)
    dT = (tg[:, 1:] - tg[:, :-1]) * float(Nz)
    l2 = ((dT - dF) ** 2).mean(dim=1)
    total = (l1 + gradient_scaling * l2).mean()
    total.backward()
    P = model.P
    g = w.grad.detach().numpy().astype(np.float64)
    return float(total.detach()), float(l1.mean().detach()), float(l2.mean().detach()), g[which_net * P:(which_net + 1) * P], g
