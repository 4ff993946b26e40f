"""GPU parity of the LAYERED tensor-core path (csrc/opz_gemm.cu: one batched tcgen05 GEMM per Dense layer, FP32 physics
kernel): the shapes the fused kernel declines — BASELINE configs[4]: Nz = 128 columns, widened 384-1024-1024-127 nets,
diurnal forcing — and the split-operand precisions OPZ_PREC_BF16X2 / OPZ_PREC_BF16X3 (near-FP32 on the tensor cores).
Everything goes through the C ABI.  The error-growth study of configs[4] is tests/test_gpu_layered.py::test_config5_*."""
import json
import os

import numpy as np
import pytest

from helpers import product_model, product_params, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _prec(opz, name):
    return {"fp16": opz.PREC_FP16, "bf16": opz.PREC_BF16, "bf16x2": opz.PREC_BF16X2, "bf16x3": opz.PREC_BF16X3,
            "fp16x2": opz.PREC_FP16X2, "fp16x3": opz.PREC_FP16X3}[name]


def _emu(oracle, name):
    return {"fp16": oracle.SplitOperandEmulation("fp16", "plain"), "bf16": oracle.SplitOperandEmulation("bf16", "plain"),
            "bf16x2": oracle.SplitOperandEmulation("bf16", "x2"), "bf16x3": oracle.SplitOperandEmulation("bf16", "x3"),
            "fp16x2": oracle.SplitOperandEmulation("fp16", "x2"), "fp16x3": oracle.SplitOperandEmulation("fp16", "x3")}[name]


# error vs the FP32 oracle a mode may show after 24 RK4 steps (its operand rounding), and vs its own CPU emulation
TOL = {"fp16": (2e-3, 5e-4), "bf16": (1e-2, 5e-3), "bf16x2": (1e-2, 5e-3), "bf16x3": (5e-5, 2e-5), "fp16x2": (2e-3, 5e-4),
       "fp16x3": (2e-5, 2e-5)}


@pytest.mark.parametrize("prec", ["fp16", "bf16", "bf16x2", "bf16x3", "fp16x2", "fp16x3"])
@pytest.mark.parametrize("sizes,act,B", [((96, 400, 400, 31), 1, 300), ((96, 50, 20, 31), 3, 129), ((96, 400, 31), 4, 64),
                                          ((96, 64, 48, 40, 31), 2, 7), ((96, 640, 31), 5, 130)])
def test_layered_path_nz32_architectures(opz, oracle, monkeypatch, sizes, act, B, prec):
    """Nz = 32 nets of every depth / width (also ones the fused kernel runs: OPZ_FORCE_LAYERED=1) through the layered path."""
    monkeypatch.setenv("OPZ_FORCE_LAYERED", "1")
    O = oracle
    p = O.Params()
    x0, bc = O.synthetic_columns(B, p)
    m = O.make_model(sizes, act, seed=4, out_scale=0.1, bias_scale=0.2)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    n0 = model.ctx.launch_count
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8, precision=_prec(opz, prec))
    assert model.ctx.launch_count - n0 >= 24 * 4 * (len(sizes) + 1)      # pack + one GEMM launch per layer + physics, per right-hand side
    assert out.shape == (B, 4, 96) and np.isfinite(out).all() and np.array_equal(out[:, 0], x0)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, 24, 8)
    refq = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, 24, 8, quant=_emu(O, prec))
    err, errq = O.profile_error(out, ref, 32), O.profile_error(out, refq, 32)
    print(f"layered {sizes} act {act}: {prec} vs fp32 oracle {err:.2e}; vs its CPU emulation {errq:.2e}")
    assert err < TOL[prec][0] and errq < TOL[prec][1]
    # final-state-only mode and the other schemes
    fin = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 0, precision=_prec(opz, prec))
    assert np.array_equal(fin[:, 0], out[:, -1])
    for scheme in (0, 1):
        o2 = opz.rollout_forward(model, pp, x0[:9], bc[:9], scheme, dt, 6, 3, precision=_prec(opz, prec))
        r2 = O.rollout(x0[:9], bc[:9], m, p, scheme, dt, 6, 3)
        assert O.profile_error(o2, r2, 32) < TOL[prec][0]
    model.close()


def test_fused_and_layered_paths_agree(opz, oracle, monkeypatch):
    """The two tensor paths on the same FP16 problem: same operand format, different rounding policy (the fused kernel dithers
    in time) — both inside the reduced-precision tolerance of the FP32 oracle and of each other."""
    O = oracle
    p = O.Params()
    x0, bc = O.synthetic_columns(200, p)
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=0, out_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    fused = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8, precision=opz.PREC_FP16)
    monkeypatch.setenv("OPZ_FORCE_LAYERED", "1")
    layered = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8, precision=opz.PREC_FP16)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, 24, 8)
    assert O.profile_error(fused, ref, 32) < 1e-3 and O.profile_error(layered, ref, 32) < 1e-3
    assert O.profile_error(fused, layered, 32) < 1e-3 and not np.array_equal(fused, layered)
    model.close()


@pytest.mark.parametrize("prec", ["fp16", "bf16", "bf16x2", "bf16x3", "fp16x3"])
def test_config5_nz128_widened_diurnal_on_the_tensor_path(opz, oracle, prec):
    """BASELINE configs[4]: Nz = 128 (state 384, 127 interior faces), 384-1024-1024-127 relu nets x3 (9.4 MFLOP per column and
    right-hand side), diurnal top heat flux; ragged ensemble (2 full tiles + 44 columns)."""
    O = oracle
    p = O.Params(Nz=128, flags=O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0)
    B = 300
    x0, bc = O.synthetic_columns(B, p, diurnal=True)
    x0 = (x0 + 0.01 * np.random.default_rng(6).normal(size=x0.shape)).astype(np.float32)
    sizes = (384, 1024, 1024, 127)
    m = O.make_model(sizes, O.ACT_RELU, seed=0, out_scale=0.1, bias_scale=0.05)
    model = product_model(opz, m, Nz=128)
    pp = product_params(opz, p)
    dt = (30.0 / 16.0) / p.tau        # dz is 4x finer: the explicit diffusion limit shrinks 16x
    n_steps, save_every = 24, 8
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, n_steps, save_every, t0=0.1, precision=_prec(opz, prec))
    assert out.shape == (B, 4, 384) and np.isfinite(out).all()
    pick = np.array([0, 1, 127, 128, 255, 256, 299])            # the oracle on a sample: first / last rows of every tile
    ref = O.rollout(x0[pick], bc[pick], m, p, O.SCHEME_RK4, dt, n_steps, save_every, t0=0.1)
    refq = O.rollout(x0[pick], bc[pick], m, p, O.SCHEME_RK4, dt, n_steps, save_every, t0=0.1, quant=_emu(O, prec))
    err, errq = O.profile_error(out[pick], ref, 128), O.profile_error(out[pick], refq, 128)
    print(f"configs[4] Nz=128 {sizes}: {prec} vs fp32 oracle {err:.2e}; vs its CPU emulation {errq:.2e}")
    assert err < TOL[prec][0] and errq < TOL[prec][1]
    # column independence across tiles: a sub-ensemble reproduces its columns bit for bit
    sub = opz.rollout_forward(model, pp, x0[pick], bc[pick], opz.SCHEME_RK4, dt, n_steps, save_every, t0=0.1, precision=_prec(opz, prec))
    assert np.array_equal(sub, out[pick])
    # the batched MLP of predict() at the same precision
    y = opz.mlp_forward(model, 1, x0, precision=_prec(opz, prec))
    yr = O.mlp_forward(x0, m.w_vw, sizes, m.act)
    assert y.shape == (B, 127) and rel_err(y, yr) < 20 * TOL[prec][0]
    model.close()


def test_near_fp32_tensor_mode_on_config1(opz, oracle, golden_dir, monkeypatch):
    """BASELINE configs[0] (construct_NN, 2 days = 5 760 RK4 steps) with OPZ_PREC_FP16X3 (both operands split into two IEEE
    halves, ~21 mantissa bits): inside the FP32 contract (<= 1e-4 of the Float32 CPU solve at every snapshot) on the tensor
    cores.  OPZ_PREC_BF16X3 (16 bits per operand) lands at the Float32-vs-double noise floor of this rollout (2.2e-4): its
    rounded weights are a constant model perturbation of 2^-17; the weights-only splits are reported."""
    g = np.load(f"{golden_dir}/config1_rk4_2day.npz")
    O = oracle
    p = O.Params()
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    model = product_model(opz, m)
    pp = product_params(opz, p)
    res = {}
    for name in ("fp16x3", "bf16x3", "fp16x2", "bf16x2"):
        out = opz.rollout_forward(model, pp, g["x0"], g["bc"], opz.SCHEME_RK4, float(g["dt_nd"]), int(g["n_steps"]), int(g["save_every"]),
                                  precision=_prec(opz, name))
        res[name] = O.profile_error(out, g["sol32"], 32)
        assert np.isfinite(out).all()
    print("config1 2-day rollout on the layered tensor path: " + ", ".join(f"{k} err {v:.2e}" for k, v in res.items()) +
          f" (fp32-vs-fp64 floor {float(g['floor']):.2e})")
    assert res["fp16x3"] <= 1e-4
    assert res["bf16x3"] <= 2.0 * float(g["floor"])
    assert res["bf16x2"] <= 1e-2 and res["fp16x2"] <= 1e-2
    model.close()


def test_config5_precision_study(opz, oracle, golden_dir):
    """The tolerance study BASELINE configs[4] asks for ("TF32/BF16 vs FP32"): Nz = 128, 384-1024-1024-127 nets, diurnal
    forcing, RK4 dt = 10 s over one day (8 640 steps) on the tensor cores in every operand precision, against the C oracle's
    Float32 AND double solutions of the same problem (tests/golden/config5_nz128_1day.npz).  One curve per mode, written to
    gpurun_out/precision_study_nz128.json (committed under profiles/).  TF32 itself is not built: its operand format has
    the 10-bit mantissa of IEEE half (the FP16 curve is its rounding error; its range is BF16's) — documented in DESIGN.md."""
    g = np.load(os.path.join(golden_dir, "config5_nz128_1day.npz"))
    O = oracle
    sizes = (384, 1024, 1024, 127)
    p = O.Params(Nz=128, tau=float(g["tau"]), flags=int(g["flags"]))
    m = O.make_model(sizes, O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]), bias_scale=float(g["bias_scale"]))
    model = product_model(opz, m, Nz=128)
    pp = product_params(opz, p)
    n_steps, se, dt = int(g["n_steps"]), int(g["save_every"]), float(g["dt_nd"])
    sol32, sol64 = g["sol32"], g["sol64"]
    curve = lambda out, ref: [float(O.profile_error(out[:, k:k + 1], ref[:, k:k + 1], 128)) for k in range(out.shape[1])]
    report = {"config": "BASELINE configs[4]: Nz=128, 384-1024-1024-127 relu x3, mPP, diurnal, RK4 dt=10 s, 8640 steps (1 day), 8 columns",
              "hours": [k * se * 10.0 / 3600.0 for k in range(sol32.shape[1])],
              "floor_f32_oracle_vs_f64_oracle": curve(sol32, sol64)}
    for name in ("bf16", "fp16", "bf16x2", "fp16x2", "bf16x3", "fp16x3"):
        out = opz.rollout_forward(model, pp, g["x0"], g["bc"], opz.SCHEME_RK4, dt, n_steps, se, precision=_prec(opz, name))
        assert out.shape == sol32.shape and np.isfinite(out).all(), name
        report[name + "_vs_f32_oracle"] = curve(out, sol32)
        report[name + "_vs_f64_oracle"] = curve(out, sol64)
        report[name + "_max"] = max(report[name + "_vs_f32_oracle"])
    model.close()
    floor = max(report["floor_f32_oracle_vs_f64_oracle"])
    report["floor_max"] = floor
    print("\nconfigs[4] precision study, max over the one-day rollout of the error vs the Float32 oracle (floor = f32 vs f64 oracle):")
    print("  floor %.2e | " % floor + " | ".join(f"{n} {report[n + '_max']:.2e}" for n in ("bf16", "fp16", "bf16x2", "fp16x2", "bf16x3", "fp16x3")))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(report, open(os.path.join(ROOT, "gpurun_out", "precision_study_nz128.json"), "w"), indent=1)
    except OSError:
        pass
    assert report["fp16_max"] <= 1e-2                       # the reduced-precision tolerance
    assert report["fp16x3_max"] <= max(1e-4, 2.0 * floor)   # the near-FP32 mode: the FP32 contract, down to the Float32 noise floor
    assert report["bf16x3_max"] <= max(1e-3, 4.0 * floor)
    assert report["bf16_max"] > report["fp16_max"] > report["fp16x3_max"]
