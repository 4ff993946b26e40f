"""Generates tests/golden/config3_rk4_8day.npz: the BENCHMARKED configuration (BASELINE configs[2]: construct_NN
96-400-400-31 relu x3, mPP + convective adjustment kappa = 1, RK4 dt = 30 s, 8 days = 23 040 steps, tau = 691 200 s,
exactly the model / parameters / forcing sweep bench.py's workload() builds) solved by the C oracle
(oracle/nde_oracle.c) for 64 columns spread over the 64 x 64 forcing sweep, in Float32 (the reference's arithmetic)
and in double (the noise floor), one snapshot every 8 hours.
Run from the repo root:  python tests/golden/make_golden_8day.py      (~2 min on 8 cores)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as CO  # noqa: E402
from oracle import nde_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_STEPS, SAVE_EVERY, SEED, OUT_SCALE, TAU, DT_S = 23040, 960, 0, 0.1, 691200.0, 30.0


def columns():
    """Every 65th column of the 64 x 64 (Q_u, Q_b) sweep: the diagonal of the grid plus wrap-around, 64 columns that
    cover weak/strong wind and heating/cooling."""
    return (np.arange(64) * 65) % 4096


def main():
    p = O.Params(tau=TAU)
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=SEED, out_scale=OUT_SCALE)
    x0s, bcs = O.synthetic_columns(4096, p)
    pick = columns()
    x0, bc = x0s[pick], bcs[pick]
    dt = float(np.float32(DT_S / p.tau))
    t = time.time()
    sol32 = CO.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, N_STEPS, SAVE_EVERY)
    t32 = time.time() - t
    t = time.time()
    sol64 = CO.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, N_STEPS, SAVE_EVERY, f64=True)
    t64 = time.time() - t
    floor = [O.profile_error(sol32[:, k:k + 1], sol64[:, k:k + 1], 32) for k in range(sol32.shape[1])]
    print(f"f32 {t32:.1f} s, f64 {t64:.1f} s; fp32-vs-fp64 floor per snapshot:", " ".join(f"{e:.2e}" for e in floor))
    assert np.isfinite(sol32).all() and np.isfinite(sol64).all()
    np.savez_compressed(os.path.join(HERE, "config3_rk4_8day.npz"), x0=x0, bc=bc, pick=pick, sol32=sol32,
                        sol64=sol64.astype(np.float64), dt_nd=np.float32(dt), n_steps=N_STEPS, save_every=SAVE_EVERY,
                        seed=SEED, out_scale=OUT_SCALE, tau=TAU, flags=p.flags, floor=np.array(floor))


if __name__ == "__main__":
    main()
