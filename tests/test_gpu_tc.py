"""GPU parity of the tcgen05 tensor-core rollout (OPZ_PREC_FP16 / OPZ_PREC_BF16) against the CPU oracle.
north_star tolerance for the reduced-precision path: max relative profile error <= 1e-2 after the
full rollout.  Everything goes through the C ABI."""
import numpy as np
import pytest

from helpers import product_model, product_params

pytestmark = pytest.mark.gpu

TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def ctx(opz):
    return opz.default_context()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("sizes,act,B", [((96, 400, 400, 31), 1, 200), ((96, 50, 20, 31), 3, 129), ((96, 400, 31), 4, 64),
                                          ((96, 128, 256, 31), 2, 7), ((96, 31), 0, 5)])
def test_tc_short_rollout_architectures(opz, oracle, ctx, sizes, act, B, prec):
    PREC = opz.PREC_FP16 if prec == "fp16" else opz.PREC_BF16
    quant = oracle.TensorPathEmulation(prec, stages=4)   # the kernel's own operand roundings (temporally dithered), on the CPU
    p = oracle.Params()
    x0, bc = oracle.synthetic_columns(B, p)
    m = oracle.make_model(sizes, act, seed=4, out_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    if len(sizes) == 2:  # a single Dense layer has no hidden GEMM: the tensor-core path declines it
        with pytest.raises(opz.OpzError) as e:
            opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 8, 4, precision=PREC)
        assert e.value.code == opz.OPZ_EUNSUP
        model.close()
        return
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8, precision=PREC)
    assert out.shape == (B, 4, 96)
    assert np.array_equal(out[:, 0], x0)
    ref = oracle.rollout(x0, bc, m, p, oracle.SCHEME_RK4, dt, 24, 8)
    err = oracle.profile_error(out, ref, 32)
    # the same rollout with GEMM operands rounded to the tensor path's format on the CPU: what the tensor path should reproduce
    refq = oracle.rollout(x0, bc, m, p, oracle.SCHEME_RK4, dt, 24, 8, quant=quant)
    errq = oracle.profile_error(out, refq, 32)
    print(f"{sizes} act {act}: {prec} vs fp32 oracle {err:.2e}; vs {prec}-emulating oracle {errq:.2e}")
    assert np.isfinite(out).all()
    assert err < (TOL_BF16 if prec == "bf16" else 1e-3)
    assert errq < 5e-4
    model.close()


@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_tc_schemes_and_final_only(opz, oracle, ctx, scheme):
    p = oracle.Params()
    B = 131
    x0, bc = oracle.synthetic_columns(B, p)
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    out = opz.rollout_forward(model, pp, x0, bc, scheme, dt, 12, 3, precision=opz.PREC_FP16)
    ref = oracle.rollout(x0, bc, m, p, scheme, dt, 12, 3)
    assert oracle.profile_error(out, ref, 32) < TOL_BF16
    fin = opz.rollout_forward(model, pp, x0, bc, scheme, dt, 12, 0, precision=opz.PREC_FP16)
    assert np.array_equal(fin[:, 0], out[:, -1])
    # zero steps: the initial state comes back
    z = opz.rollout_forward(model, pp, x0, bc, scheme, dt, 0, 1, precision=opz.PREC_FP16)
    assert np.array_equal(z[:, 0], x0)
    model.close()


def test_tc_config1_full_two_day_rollout(opz, oracle, golden_dir, ctx):
    """BASELINE configs[0] length (5760 RK4 steps, construct_NN) on the reduced-precision tensor path
    (IEEE-half operands, FP32 accumulation, temporally dithered weight / state rounding): within 1e-2 of the FP32 CPU
    solve at every saved snapshot.  The BF16-operand variant (8x coarser grid) is reported and bounded too.  The full
    8-day rollout of the benchmarked configuration is tests/test_gpu_long.py."""
    g = np.load(f"{golden_dir}/config1_rk4_2day.npz")
    p = oracle.Params()
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    model = product_model(opz, m)
    out = opz.rollout_forward(model, product_params(opz, p), g["x0"], g["bc"], opz.SCHEME_RK4, float(g["dt_nd"]),
                              int(g["n_steps"]), int(g["save_every"]), precision=opz.PREC_FP16)
    err = oracle.profile_error(out, g["sol32"], 32)
    outb = opz.rollout_forward(model, product_params(opz, p), g["x0"], g["bc"], opz.SCHEME_RK4, float(g["dt_nd"]),
                               int(g["n_steps"]), int(g["save_every"]), precision=opz.PREC_BF16)
    errb = oracle.profile_error(outb, g["sol32"], 32)
    print(f"config1 tensor-core rollout: fp16 err vs fp32 oracle {err:.3e}; bf16 err {errb:.3e}")
    assert np.isfinite(out).all() and np.isfinite(outb).all()
    assert err < 1e-3       # (plainly rounded FP16 operands: 1.4e-3; dithered: ~1e-4)
    assert errb < TOL_BF16  # (plainly rounded BF16 operands: 1.06e-2)
    model.close()


def test_tc_ensemble_shard_invariance_and_multi_tile(opz, oracle, ctx):
    """More tiles than SMs (persistent loop over tiles) and bit-exact column independence."""
    p = oracle.Params()
    B = 148 * 128 + 77
    x0s, bcs = oracle.synthetic_columns(4096, p)
    reps = (B + 4095) // 4096
    x0 = np.tile(x0s, (reps, 1))[:B]
    bc = np.tile(bcs, (reps, 1))[:B]
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    full = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 4, 2, precision=opz.PREC_FP16)
    assert np.isfinite(full).all()
    pick = np.array([0, 127, 128, 4095, 4096, 148 * 128 - 1, 148 * 128, B - 1])
    sub = opz.rollout_forward(model, pp, x0[pick], bc[pick], opz.SCHEME_RK4, dt, 4, 2, precision=opz.PREC_FP16)
    assert np.array_equal(sub, full[pick])
    ref = oracle.rollout(x0[pick], bc[pick], m, p, oracle.SCHEME_RK4, dt, 4, 2)
    assert oracle.profile_error(full[pick], ref, 32) < TOL_BF16
    # updated weights invalidate the packed tape
    w = model.get_weights()
    model.set_weights(w * 0.5)
    half = opz.rollout_forward(model, pp, x0[:9], bc[:9], opz.SCHEME_RK4, dt, 4, 2, precision=opz.PREC_FP16)
    assert not np.array_equal(half, full[:9])
    model.close()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("sizes,act,B", [((96, 400, 400, 31), 1, 300), ((96, 50, 20, 31), 3, 77), ((96, 400, 31), 4, 130)])
def test_tc_nonzero_biases(opz, oracle, ctx, sizes, act, B, prec):
    """Flux initialises biases to zero, trained models do not keep them there: every Dense bias non-zero (the tensor-core
    path folds the hidden biases into the weight tape as an extra k-step against a constant-one operand)."""
    PREC = opz.PREC_FP16 if prec == "fp16" else opz.PREC_BF16
    quant = oracle.TensorPathEmulation(prec, stages=4)
    p = oracle.Params()
    x0, bc = oracle.synthetic_columns(B, p)
    m = oracle.make_model(sizes, act, seed=11, out_scale=0.1, bias_scale=0.3)
    assert all(np.abs(w).max() > 0 for w in (m.w_uw, m.w_vw, m.w_wT))
    m0 = oracle.make_model(sizes, act, seed=11, out_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8, precision=PREC)
    ref = oracle.rollout(x0, bc, m, p, oracle.SCHEME_RK4, dt, 24, 8)
    ref0 = oracle.rollout(x0, bc, m0, p, oracle.SCHEME_RK4, dt, 24, 8)
    refq = oracle.rollout(x0, bc, m, p, oracle.SCHEME_RK4, dt, 24, 8, quant=quant)
    err = oracle.profile_error(out, ref, 32)
    errq = oracle.profile_error(out, refq, 32)
    bias_effect = oracle.profile_error(ref0, ref, 32)
    print(f"{sizes} act {act} biased: {prec} vs fp32 oracle {err:.2e}; vs {prec}-emulating oracle {errq:.2e}; bias effect {bias_effect:.2e}")
    assert bias_effect > 20 * max(err, 1e-4)   # the biases matter at the size of this test
    assert np.isfinite(out).all()
    assert err < (TOL_BF16 if prec == "bf16" else 2e-3)
    assert errq < (5e-3 if prec == "bf16" else 5e-4)   # biases are rounded to the operand format on the tensor path (bf16: 2^-9 relative)
    model.close()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("sizes,act,B,which", [((96, 400, 400, 31), 1, 150, "both"), ((96, 50, 20, 31), 3, 129, "both"),
                                                ((96, 400, 31), 4, 40, "nn"), ((96, 50, 20, 31), 3, 33, "ri")])
def test_tc_box_filters(opz, oracle, ctx, sizes, act, B, which, prec):
    """smooth_NN / smooth_Ri (filtering_operators.jl:1-15 at NDE_training.jl:98-102,121-123) on the tensor-core path: the
    filtered outputs / Richardson numbers of the two threads that share an ocean column meet at cell 16, and the edge rows of
    the filter sit on faces 1 and 31."""
    O = oracle
    PREC = opz.PREC_FP16 if prec == "fp16" else opz.PREC_BF16
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    fl |= {"both": O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI, "nn": O.FLAG_SMOOTH_NN, "ri": O.FLAG_SMOOTH_RI}[which]
    p = O.Params(flags=fl)
    x0, bc = O.synthetic_columns(B, p)
    x0 = (x0 + 0.02 * np.random.default_rng(5).normal(size=x0.shape)).astype(np.float32)
    m = O.make_model(sizes, act, seed=6, out_scale=0.1, bias_scale=0.2)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8, precision=PREC)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, 24, 8)
    refq = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, 24, 8, quant=O.TensorPathEmulation(prec, stages=4))
    p0 = O.Params(flags=fl & ~(O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI))
    ref0 = O.rollout(x0, bc, m, p0, O.SCHEME_RK4, dt, 24, 8)
    err, errq, effect = O.profile_error(out, ref, 32), O.profile_error(out, refq, 32), O.profile_error(ref0, ref, 32)
    print(f"{sizes} filters={which}: {prec} vs fp32 oracle {err:.2e}; vs emulating oracle {errq:.2e}; filter effect {effect:.2e}")
    assert np.isfinite(out).all()
    assert effect > 20 * max(err, 1e-5)           # the filters matter at the size of this test
    assert err < (TOL_BF16 if prec == "bf16" else 2e-3)
    assert errq < (5e-3 if prec == "bf16" else 5e-4)
    # and the FP32 path agrees with itself through the same flags
    o32 = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 24, 8)
    assert O.profile_error(o32, ref, 32) < 2e-5
    model.close()


def test_tc_full_size_ensemble_periodicity(opz, oracle, ctx):
    """BASELINE configs[2] size (2^20 columns on one GPU) on the tensor-core path: the synthetic forcing sweep repeats
    every 4096 columns, so the result must repeat bit for bit (column independence and determinism across all 8192 tiles,
    the slab pipeline of the host-buffer call included), and a sample of columns must sit on the oracle."""
    p = oracle.Params()
    B = 1 << 20
    x0s, bcs = oracle.synthetic_columns(4096, p)
    x0 = np.ascontiguousarray(np.tile(x0s, (B // 4096, 1)))
    bc = np.ascontiguousarray(np.tile(bcs, (B // 4096, 1)))
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1, bias_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 4, 0, precision=opz.PREC_FP16)
    assert out.shape == (B, 1, 96) and np.isfinite(out).all()
    blocks = out.reshape(B // 4096, 4096, 96)
    assert np.array_equal(blocks, np.broadcast_to(blocks[0], blocks.shape))
    pick = np.array([0, 63, 64, 4095])
    ref = oracle.rollout(x0s[pick], bcs[pick], m, p, oracle.SCHEME_RK4, dt, 4, 4)[:, -1:]
    assert oracle.profile_error(blocks[0][pick][:, None, :], ref, 32) < 1e-3
    model.close()
