"""Shared helpers for the tests: build product objects from oracle objects."""
import numpy as np

ACT_NAMES = {0: "identity", 1: "relu", 2: "tanh", 3: "mish", 4: "swish", 5: "leakyrelu"}


def chains_from_oracle_model(opz, m):
    """oracle Model (flat Flux vectors) -> three product Chains."""
    from oracle import nde_oracle as O
    chains = []
    for w in (m.w_uw, m.w_vw, m.w_wT):
        layers = O.unpack_weights(np.asarray(w, dtype=np.float32), m.sizes)
        dense = []
        for li, (W, b) in enumerate(layers):
            act = ACT_NAMES[m.act] if li < len(layers) - 1 else "identity"
            dense.append(opz.Dense(W.shape[1], W.shape[0], act, W=W, b=b))
        chains.append(opz.Chain(*dense))
    return chains


def product_model(opz, m, Nz=32, ctx=None):
    uw, vw, wT = chains_from_oracle_model(opz, m)
    return opz.NDEModel(uw, vw, wT, Nz, ctx=ctx)


def product_params(opz, p):
    """oracle Params -> struct opz_params."""
    q = opz.OpzParams()
    q.nz = p.Nz
    q.H, q.tau, q.f, q.g, q.alpha = p.H, p.tau, p.f, p.g, p.alpha
    q.nu0, q.nu_m, q.Ric, q.dRi, q.Pr, q.kappa, q.eps = p.nu0, p.nu_m, p.Ric, p.dRi, p.Pr, p.kappa, p.eps
    for i in range(6):
        q.mu[i] = p.mu[i]
        q.sigma[i] = p.sigma[i]
    q.flags = p.flags
    q.diurnal_period = p.diurnal_period
    return q


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
