"""Generates tests/golden/bptt_long_1152.npz: a LONG-horizon gradient fixture from the float64 autograd twin
(oracle/nde_oracle_torch.py): BASELINE configs[3]'s training step shape (construct_NN, training-flavour RHS: mPP with eps +
zero_weights boundary form, RK4 dt = 60 s, tau = 8 days, profile + gradient loss) over 1 152 time steps (4 608 right-hand
sides, 13 snapshots) for 6 of the forcing cases.  The 634 893-entry gradient is stored as every 7th entry plus the L2 norm
of every (net, layer, W | b) block and of the whole vector.
Run from the repo root:  python tests/golden/make_golden_bptt_long.py      (~2 min, ~2 GB of autograd graph)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nde_oracle as O  # noqa: E402
from oracle import nde_oracle_torch as OT  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SIZES = (96, 400, 400, 31)
N_STEPS, SAVE_EVERY, B, STRIDE = 1152, 96, 6, 7


def blocks(P, sizes):
    """(start, stop) of every W / b block of the three nets inside the flat [uw; vw; wT] vector."""
    out = []
    for net in range(3):
        off = net * P
        for i in range(len(sizes) - 1):
            for n_el in (sizes[i] * sizes[i + 1], sizes[i + 1]):
                out.append((off, off + n_el))
                off += n_el
    return out


def case():
    p = O.Params(tau=691200.0, flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    x0, bc = O.synthetic_columns(18, p)
    pick = np.array([0, 4, 8, 9, 13, 17])   # weak / medium / strong wind, cooling and heating
    x0, bc = x0[pick], bc[pick]
    m = O.make_model(SIZES, O.ACT_RELU, seed=0, out_scale=0.1)
    m_true = O.make_model(SIZES, O.ACT_RELU, seed=5, out_scale=0.1)
    dt = float(np.float32(60.0 / p.tau))
    return p, x0, bc, m, m_true, dt


def main():
    from oracle import c_oracle as CO
    p, x0, bc, m, m_true, dt = case()
    target = CO.rollout(x0, bc, m_true, p, O.SCHEME_RK4, dt, N_STEPS, SAVE_EVERY)
    lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    t = time.time()
    total, comps, grad = OT.loss_and_grad(x0, bc, m, p, O.SCHEME_RK4, dt, N_STEPS, SAVE_EVERY, target, lw)
    print(f"autograd twin: {time.time() - t:.1f} s, loss {total:.6e}, |grad| {np.linalg.norm(grad):.6e}")
    bl = blocks(m.P, SIZES)
    np.savez_compressed(os.path.join(HERE, "bptt_long_1152.npz"), x0=x0, bc=bc, target=target, loss_w=lw, dt_nd=np.float32(dt),
                        n_steps=N_STEPS, save_every=SAVE_EVERY, total=total, comps=comps, grad_sub=grad[::STRIDE].astype(np.float32),
                        stride=STRIDE, block_norms=np.array([np.linalg.norm(grad[a:b]) for a, b in bl]), grad_norm=np.linalg.norm(grad),
                        seed=0, out_scale=0.1, flags=p.flags, tau=p.tau)


if __name__ == "__main__":
    main()
