"""Generates the committed gradient fixture tests/golden/bptt_config4_small.npz from the float64
autograd twin (oracle/nde_oracle_torch.py): a shortened BASELINE configs[3] training step
(18 forcing cases, construct_NN, training-flavour RHS, RK4, profile + gradient loss).
Run from the repo root:  python tests/golden/make_golden_bptt.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nde_oracle as O  # noqa: E402
from oracle import nde_oracle_torch as OT  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def config4_like(n_steps=48, save_every=12):
    """18 cases = 3x3 wind x cooling + 3x3 wind x heating (train_NDE_args.jl:39-60 mirrored by the
    synthetic sweep), tau = 8 days, mPP with eps + zero_weights boundary form, dt = 60 s."""
    p = O.Params(tau=691200.0, flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    x0, bc = O.synthetic_columns(18, p)
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=0, out_scale=0.1)
    m_true = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=5, out_scale=0.1)
    dt = float(np.float32(60.0 / p.tau))  # the value a Float32 caller passes through the ABI
    target = O.rollout(x0, bc, m_true, p, O.SCHEME_RK4, dt, n_steps, save_every)
    return p, x0, bc, m, target, dt


def main():
    n_steps, save_every = 48, 12
    p, x0, bc, m, target, dt = config4_like(n_steps, save_every)
    lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    total, comps, grad = OT.loss_and_grad(x0, bc, m, p, O.SCHEME_RK4, dt, n_steps, save_every, target, lw)
    np.savez_compressed(os.path.join(HERE, "bptt_config4_small.npz"), x0=x0, bc=bc, target=target, loss_w=lw,
                        dt_nd=np.float32(dt), n_steps=n_steps, save_every=save_every, total=total, comps=comps,
                        grad=grad.astype(np.float32), seed=0, out_scale=0.1, flags=p.flags, tau=p.tau)
    print("bptt fixture: loss", total, "|grad|", np.linalg.norm(grad))


if __name__ == "__main__":
    main()
