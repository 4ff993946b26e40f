"""Hand-derived discrete adjoint (BPTT) of the fixed-step NDE solve (TEST INFRASTRUCTURE, not product).

NumPy restatement, stage by stage, of exactly the recurrences the CUDA BPTT path
(oceanparameterizations.jl_b200/csrc/opz_bptt.cu) implements; it is checked on the CPU against the
autograd twin (oracle/nde_oracle_torch.py), so a formula error is caught before a GPU run.
Replaces, as a checker, Zygote through `InterpolatingAdjoint` (wind_mixing/src/NDE_training.jl:290-333).

Forward per right-hand side, per column:
    h_0 = x;  z_l = h_{l-1} W_l + b_l;  h_l = act(z_l) (l < L);  y = z_L           (three nets)
    g = D^f x;  Ri(g);  nu(Ri);  nu_T;  F = y - c nu g;  k = A D^c F + Coriolis
Backward, given kbar = dL/dk:
    Fbar[f] = A nz (kbar[f-1] - kbar[f])            interior faces f = 1..Nz-1
    ybar = Fbar;  nubar = -sum_v c_v g_v Fbar_v;  Ribar = nubar dnu/dRi;  gbar = -c nu Fbar + Ribar dRi/dg
    xbar[f] += nz gbar[f];  xbar[f-1] -= nz gbar[f];  Coriolis transpose;  xbar += W_1^T dz_1 (MLP backward)
    dW_l += h_{l-1}^T dz_l;  db_l += sum dz_l
Runge-Kutta adjoint (b = final weights, a = stage offsets):
    kbar_S = b_S lam;  xbar_s = J_s^T kbar_s;  kbar_{s-1} = b_{s-1} lam + a_{s-1} xbar_s;  lam_n = lam + sum_s xbar_s
"""
from __future__ import annotations

import numpy as np

from . import nde_oracle as O


def act_grad(z, act):
    one = z.dtype.type(1)
    if act == O.ACT_IDENTITY:
        return np.ones_like(z)
    if act == O.ACT_RELU:
        return (z > 0).astype(z.dtype)
    if act == O.ACT_TANH:
        return one - np.tanh(z) ** 2
    if act == O.ACT_MISH:
        sp = np.logaddexp(z, z.dtype.type(0))
        t = np.tanh(sp)
        sg = one / (one + np.exp(-z))
        return t + z * (one - t * t) * sg
    if act == O.ACT_SWISH:
        sg = one / (one + np.exp(-z))
        return sg + z * sg * (one - sg)
    if act == O.ACT_LEAKYRELU:
        return np.where(z > 0, one, z.dtype.type(0.01))
    raise ValueError(act)


def rk_tables(scheme, h):
    """(b, a, c): x+ = x + sum_s b_s k_s ; stage input s+1 = x + a_s k_s ; stage time = t + c_s h."""
    if scheme == O.SCHEME_EULER:
        return [h], [], [0.0]
    if scheme == O.SCHEME_RK2:
        return [h * 0.5, h * 0.5], [h], [0.0, 1.0]
    return [h / 6.0, h / 3.0, h / 3.0, h / 6.0], [h * 0.5, h * 0.5, h], [0.0, 0.5, 0.5, 1.0]


def rhs_forward(x, bc, t, nets, sizes, act, p: O.Params):
    """Returns k[B,3Nz] and the record needed by rhs_backward."""
    Nz = p.Nz
    dt = x.dtype.type
    rec = {"x": x, "h": [], "gp": []}
    ys = []
    for layers in nets:
        h, hs, gps = x, [], []
        for li, (W, b) in enumerate(layers):
            z = h @ W.T + b
            if li < len(layers) - 1:
                h = O.activation(z, act)
                hs.append(h)
                gps.append(act_grad(z, act))
            else:
                ys.append(z)
        rec["h"].append(hs)
        rec["gp"].append(gps)
    s_u, s_v, s_T, s_uw, s_vw, s_wT = (dt(s) for s in p.sigma)
    sub = dt(1) if (p.flags & O.FLAG_BC_MINUS_SCALE0) else dt(0)
    off = [dt(-p.mu[3 + i]) / dt(p.sigma[3 + i]) for i in range(3)]
    B = x.shape[0]
    F = np.empty((3, B, Nz + 1), dtype=x.dtype)
    for i in range(3):
        F[i, :, 0] = bc[:, 2 * i] - sub * off[i]
        F[i, :, Nz] = bc[:, 2 * i + 1] - sub * off[i]
        F[i, :, 1:Nz] = ys[i]
    if p.flags & O.FLAG_DIURNAL:
        phys = O.diurnal_wT_top(bc[:, 6].astype(x.dtype), dt(t) * dt(p.tau), p, dt)
        top = (phys - dt(p.mu[5])) / s_wT
        if p.flags & O.FLAG_DIURNAL_MINUS_SCALE0:
            top = top - off[2]
        F[2, :, Nz] = top
    u, v, T = x[:, :Nz], x[:, Nz:2 * Nz], x[:, 2 * Nz:]
    if p.flags & O.FLAG_MPP:
        assert not (p.flags & (O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI)), "adjoint restatement: no box filters"
        g = np.stack([(c[:, 1:] - c[:, :-1]) * dt(Nz) for c in (u, v, T)])  # interior faces
        e = dt(p.eps) if (p.flags & O.FLAG_RI_EPS) else dt(0)
        cBz = dt(p.H) * dt(p.g) * dt(p.alpha) * s_T
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            a_, b_ = s_u * (g[0] + e), s_v * (g[1] + e)
            S2 = a_ * a_ + b_ * b_
            Ri = cBz * (g[2] + e) / S2
            th = np.tanh((Ri - dt(p.Ric)) / dt(p.dRi))
        nu = dt(p.nu0) + dt(p.nu_m) * ((dt(1) - th) / dt(2))
        if p.flags & O.FLAG_CA:
            sel = g[0] if (p.flags & O.FLAG_CA_SELECT_DUDZ) else g[2]
            keep = sel > 0
            nuT = np.where(keep, nu / dt(p.Pr), dt(p.kappa))
        else:
            keep = np.ones_like(nu, dtype=bool)
            nuT = nu / dt(p.Pr)
        c = [s_u / s_uw / dt(p.H), s_v / s_vw / dt(p.H), s_T / s_wT / dt(p.H)]
        F[0, :, 1:Nz] -= c[0] * nu * g[0]
        F[1, :, 1:Nz] -= c[1] * nu * g[1]
        F[2, :, 1:Nz] -= c[2] * nuT * g[2]
        rec.update(g=g, e=e, S2=S2, Ri=Ri, th=th, nu=nu, nuT=nuT, keep=keep, c=c, cBz=cBz)
    A = -dt(p.tau) / dt(p.H)
    ft = dt(p.f) * dt(p.tau)
    Ac = [A * s_uw / s_u, A * s_vw / s_v, A * s_wT / s_T]
    k = np.empty_like(x)
    k[:, :Nz] = Ac[0] * ((F[0, :, 1:] - F[0, :, :-1]) * dt(Nz)) + ft / s_u * (s_v * v + dt(p.mu[1]))
    k[:, Nz:2 * Nz] = Ac[1] * ((F[1, :, 1:] - F[1, :, :-1]) * dt(Nz)) - ft / s_v * (s_u * u + dt(p.mu[0]))
    k[:, 2 * Nz:] = Ac[2] * ((F[2, :, 1:] - F[2, :, :-1]) * dt(Nz))
    rec.update(Ac=Ac, ft=ft)
    return k, rec


def rhs_backward(kbar, rec, nets, sizes, act, p: O.Params, grads):
    """xbar = J^T kbar; accumulates dW/db into grads[net][layer] = [dW (out x in), db]."""
    Nz = p.Nz
    x = rec["x"]
    dt = x.dtype.type
    s_u, s_v = dt(p.sigma[0]), dt(p.sigma[1])
    nz = dt(Nz)
    xbar = np.zeros_like(x)
    kb = [kbar[:, :Nz], kbar[:, Nz:2 * Nz], kbar[:, 2 * Nz:]]
    # Coriolis transpose
    xbar[:, Nz:2 * Nz] += (rec["ft"] / s_u * s_v) * kb[0]
    xbar[:, :Nz] -= (rec["ft"] / s_v * s_u) * kb[1]
    # flux cotangents on interior faces f = 1..Nz-1 : top of cell f-1 (+), bottom of cell f (-)
    Fbar = [rec["Ac"][i] * nz * (kb[i][:, :-1] - kb[i][:, 1:]) for i in range(3)]
    if p.flags & O.FLAG_MPP:
        g, nu, nuT, c = rec["g"], rec["nu"], rec["nuT"], rec["c"]
        nubar = -c[0] * g[0] * Fbar[0] - c[1] * g[1] * Fbar[1] - np.where(rec["keep"], c[2] * g[2] * Fbar[2] / dt(p.Pr), dt(0))
        dnu_dRi = -dt(p.nu_m) / (dt(2) * dt(p.dRi)) * (dt(1) - rec["th"] ** 2)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            Ribar = nubar * dnu_dRi
            dRi_dgT = rec["cBz"] / rec["S2"]
            dRi_dgu = -rec["Ri"] * dt(2) * s_u * s_u * (g[0] + rec["e"]) / rec["S2"]
            dRi_dgv = -rec["Ri"] * dt(2) * s_v * s_v * (g[1] + rec["e"]) / rec["S2"]
            ok = np.isfinite(rec["Ri"]) & np.isfinite(dRi_dgT)
            gbar = [-c[0] * nu * Fbar[0] + np.where(ok, Ribar * dRi_dgu, dt(0)),
                    -c[1] * nu * Fbar[1] + np.where(ok, Ribar * dRi_dgv, dt(0)),
                    -c[2] * nuT * Fbar[2] + np.where(ok, Ribar * dRi_dgT, dt(0))]
        for i in range(3):
            xbar[:, i * Nz + 1:(i + 1) * Nz] += nz * gbar[i]
            xbar[:, i * Nz:(i + 1) * Nz - 1] -= nz * gbar[i]
    # MLP backward
    for ni, layers in enumerate(nets):
        L = len(layers)
        d = Fbar[ni]  # dy
        for li in range(L - 1, -1, -1):
            W, _ = layers[li]
            hin = x if li == 0 else rec["h"][ni][li - 1]
            grads[ni][li][0] += d.T @ hin
            grads[ni][li][1] += d.sum(axis=0)
            dh = d @ W
            if li > 0:
                d = dh * rec["gp"][ni][li - 1]
            else:
                xbar += dh
    return xbar


def loss_and_grad(x0, bc, model: O.Model, p: O.Params, scheme, dt_nd, n_steps, save_every, target, loss_w, t0=0.0,
                  dtype=np.float64, B_total=None):
    """Returns (total, weighted components[6], grad[3P] in Flux.destructure order) like the torch twin."""
    Nz = p.Nz
    ft = np.dtype(dtype).type
    x = np.asarray(x0, dtype=dtype)
    bc = np.asarray(bc, dtype=dtype)
    tg = np.asarray(target, dtype=dtype)
    sizes = tuple(model.sizes)
    nets = [O.unpack_weights(np.asarray(w, dtype=dtype), sizes) for w in (model.w_uw, model.w_vw, model.w_wT)]
    B = x.shape[0]
    Bt = B if B_total is None else B_total
    n_save = n_steps // save_every + 1
    h = ft(dt_nd)
    b, a, c = rk_tables(scheme, h)
    S = len(b)
    # ---- forward sweep with records ----
    xs = [x]
    recs = []
    for n in range(n_steps):
        tn = ft(t0) + ft(n) * h
        xn = xs[-1]
        stage_recs, ks = [], []
        xin = xn
        for s in range(S):
            k, rec = rhs_forward(xin, bc, tn + ft(c[s]) * h, nets, sizes, model.act, p)
            ks.append(k)
            stage_recs.append(rec)
            if s + 1 < S:
                xin = xn + ft(a[s]) * k
        xs.append(xn + sum(ft(b[s]) * ks[s] for s in range(S)))
        recs.append(stage_recs)
    # ---- loss and its state cotangents at the save points ----
    comps = np.zeros(6)
    inj = {}
    for i in range(n_save):
        n = i * save_every
        lam_i = np.zeros_like(x)
        for v in range(3):
            sl = slice(v * Nz, (v + 1) * Nz)
            d = xs[n][:, sl] - tg[:, i, sl]
            comps[v] += float(loss_w[v]) * float((d ** 2).sum()) / (Nz * n_save * Bt)
            lam_i[:, sl] += ft(loss_w[v]) * 2 * d / (Nz * n_save * Bt)
            dg = (d[:, 1:] - d[:, :-1]) * ft(Nz)  # face-gradient difference on the interior faces
            comps[3 + v] += float(loss_w[3 + v]) * float((dg ** 2).sum()) / ((Nz + 1) * n_save * Bt)
            sc = ft(loss_w[3 + v]) * 2 * ft(Nz) / ((Nz + 1) * n_save * Bt)
            lam_i[:, v * Nz + 1:(v + 1) * Nz] += sc * dg
            lam_i[:, v * Nz:(v + 1) * Nz - 1] -= sc * dg
        inj[n] = lam_i
    # ---- backward sweep ----
    grads = [[[np.zeros_like(W), np.zeros_like(bb)] for (W, bb) in layers] for layers in nets]
    lam = inj[n_steps].copy()
    for n in range(n_steps - 1, -1, -1):
        xsum = np.zeros_like(x)
        kbar = ft(b[S - 1]) * lam
        for s in range(S - 1, -1, -1):
            xbar = rhs_backward(kbar, recs[n][s], nets, sizes, model.act, p, grads)
            xsum += xbar
            if s > 0:
                kbar = ft(b[s - 1]) * lam + ft(a[s - 1]) * xbar
        lam = lam + xsum
        if n in inj:
            lam = lam + inj[n]
    flat = np.concatenate([np.concatenate([np.concatenate([gW.T.reshape(-1), gb]) for gW, gb in net]) for net in grads])
    return float(comps.sum()), comps, flat
