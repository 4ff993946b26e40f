"""Generates the committed golden vectors from the CPU oracle (oracle/nde_oracle.py).

The reference (Julia) cannot run in this image and holds no fixtures for the NDE path, so these
vectors pin the ORACLE (regression) and give the GPU tests full-length rollouts to compare with
without paying the oracle's run time on the GPU box.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nde_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def config1_like(B, seed=0):
    """BASELINE.json configs[0]: Nz=32, construct_NN (96-400-400-31 relu), mPP + CA(kappa=1),
    2-day rollout, RK4 dt = 30 s (5760 steps), 11 snapshots; B columns of the swept forcing."""
    p = O.Params()
    x0, bc = O.synthetic_columns(B, p)
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=seed, out_scale=0.1)
    return p, x0, bc, m


def main():
    # (1) full-length config-1 rollout, fp32 oracle + fp64 twin (noise floor)
    p, x0, bc, m = config1_like(8)
    dt = 30.0 / p.tau
    s32 = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, 5760, 576)
    s64 = O.rollout(x0.astype(np.float64), bc.astype(np.float64), m.astype(np.float64), p, O.SCHEME_RK4, dt, 5760, 576)
    np.savez_compressed(os.path.join(HERE, "config1_rk4_2day.npz"), x0=x0, bc=bc, sol32=s32,
                        sol64=s64.astype(np.float32), dt_nd=np.float32(dt), n_steps=5760, save_every=576, seed=0,
                        out_scale=0.1, floor=O.profile_error(s32, s64, 32))
    print("config1: fp32-vs-fp64 noise floor", O.profile_error(s32, s64, 32))

    # (2) small production net (96-50-20-31 mish), training flavour, diurnal, Heun, 8-day tau
    p2 = O.Params(tau=691200.0, flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL)
    x02, bc2 = O.synthetic_columns(6, p2, diurnal=True)
    m2 = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=1, out_scale=0.1)
    dt2 = 60.0 / p2.tau
    s2 = O.rollout(x02, bc2, m2, p2, O.SCHEME_RK2, dt2, 2880, 288)
    np.savez_compressed(os.path.join(HERE, "small_mish_diurnal_heun.npz"), x0=x02, bc=bc2, sol32=s2,
                        dt_nd=np.float32(dt2), n_steps=2880, save_every=288, seed=1, out_scale=0.1)

    # (3) single RHS evaluations with raw Glorot weights for every flag flavour
    rng = np.random.default_rng(7)
    out = {}
    flavours = {
        "inference_ca": O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0,
        "inference_ca_quirk": O.FLAG_MPP | O.FLAG_CA | O.FLAG_CA_SELECT_DUDZ | O.FLAG_BC_MINUS_SCALE0,
        "training_plain": O.FLAG_MPP | O.FLAG_RI_EPS,
        "training_zero_weights": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0,
        "training_smooth": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI,
        "nn_only": 0,
        "diurnal": O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL,
        "diurnal_offset": O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0,
    }
    pr = O.Params()
    xr = rng.normal(size=(5, 96)).astype(np.float32)
    _, bcr = O.synthetic_columns(5, pr, diurnal=True)
    mr = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=3)
    out["x"], out["bc"] = xr, bcr
    for name, fl in flavours.items():
        pf = O.Params(flags=fl)
        out["k_" + name] = O.rhs(xr, bcr, 0.37, mr, pf)
        uw, vw, wT = O.predict_flux(xr, bcr, 0.37, mr, pf)
        out["F_" + name] = np.stack([uw, vw, wT], axis=1)
        out["flags_" + name] = fl
    np.savez_compressed(os.path.join(HERE, "rhs_flavours.npz"), **out)
    print("written")


if __name__ == "__main__":
    main()
