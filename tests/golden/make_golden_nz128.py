"""Generates tests/golden/config5_nz128_1day.npz: BASELINE configs[4] — Nz = 128 columns (dz = 2 m), widened nets
384-1024-1024-127 relu x3 (a choice: the reference has no Nz = 128 net), modified Pacanowski-Philander diffusivity,
DIURNAL top heat flux (data_containers.jl:131-156), RK4 dt = 10 s over one day (8 640 steps; at dz = 2 m the explicit limit of the
mPP diffusion is ~28 s and dt = 20 s already amplifies rounding noise to 5e-3; random-init output layer x0.03: with x0.1 the
Float32-vs-double gap of this configuration grows to 4e-2 within a day, leaving nothing to study) — solved by the C oracle
(oracle/nde_oracle.c) in Float32 and in double for 8 columns of the forcing sweep, one snapshot every hour.  It is the
reference of the TF32/BF16-vs-FP32 tolerance study (tests/test_gpu_layered.py::test_config5_precision_study).
Run from the repo root:  python tests/golden/make_golden_nz128.py      (~5 min on 8 cores)
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
SIZES = (384, 1024, 1024, 127)
N_STEPS, SAVE_EVERY, DT_S, TAU, OUT_SCALE = 8640, 360, 10.0, 172800.0, 0.03


def case():
    p = O.Params(Nz=128, tau=TAU, flags=O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0)
    x0, bc = O.synthetic_columns(64, p, diurnal=True)
    pick = np.arange(8) * 9     # 8 columns across the (wind, diurnal amplitude) sweep
    m = O.make_model(SIZES, O.ACT_RELU, seed=0, out_scale=OUT_SCALE, bias_scale=0.05)
    return p, x0[pick], bc[pick], m, float(np.float32(DT_S / p.tau))


def main():
    p, x0, bc, m, dt = case()
    t = time.time()
    sol32 = CO.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, N_STEPS, SAVE_EVERY)
    t32 = time.time() - t
    t = time.time()
    sol64 = CO.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, N_STEPS, SAVE_EVERY, f64=True)
    t64 = time.time() - t
    floor = [O.profile_error(sol32[:, k:k + 1], sol64[:, k:k + 1], 128) for k in range(sol32.shape[1])]
    print(f"f32 {t32:.1f} s, f64 {t64:.1f} s; finite {np.isfinite(sol32).all()}; |x|max {np.abs(sol32).max():.2f}; floor:", " ".join(f"{e:.1e}" for e in floor))
    assert np.isfinite(sol32).all() and np.isfinite(sol64).all()
    np.savez_compressed(os.path.join(HERE, "config5_nz128_1day.npz"), x0=x0, bc=bc, sol32=sol32, sol64=sol64, dt_nd=np.float32(dt),
                        n_steps=N_STEPS, save_every=SAVE_EVERY, seed=0, out_scale=OUT_SCALE, bias_scale=0.05, tau=TAU, flags=p.flags,
                        floor=np.array(floor))


if __name__ == "__main__":
    main()
