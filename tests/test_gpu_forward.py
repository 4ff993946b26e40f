"""GPU parity of the forward path (rhs, rollout, batched MLP) against the CPU oracle and the
committed golden vectors.  Everything goes through the C ABI (ctypes -> libopz.so)."""
import numpy as np
import pytest

from helpers import product_model, product_params, rel_err

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4  # north_star: max relative profile error of the FP32 path after the full rollout


@pytest.fixture(scope="module")
def ctx(opz):
    return opz.default_context()


def test_no_cpu_fallback_and_native_lib_loaded(opz, ctx):
    import os
    assert os.path.exists(opz.LIB_PATH)
    assert opz.lib.opz_version() == 100
    n0 = ctx.launch_count
    assert n0 >= 0


def test_rhs_flavours_match_golden(opz, oracle, golden_dir, ctx):
    g = np.load(f"{golden_dir}/rhs_flavours.npz")
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=3)
    model = product_model(opz, m)
    names = [k[2:] for k in g.files if k.startswith("k_")]
    assert len(names) == 8
    for name in names:
        p = oracle.Params(flags=int(g["flags_" + name]))
        k, F = opz.rhs(model, product_params(opz, p), g["x"], g["bc"], 0.37, with_fluxes=True)
        # raw Glorot weights: O(1) fluxes; compare per variable against its own amplitude
        for v in range(3):
            assert rel_err(F[:, v], g["F_" + name][:, v]) < 2e-5, (name, "flux", v)
            sl = slice(32 * v, 32 * (v + 1))
            assert rel_err(k[:, sl], g["k_" + name][:, sl]) < 5e-5, (name, "tendency", v)
    model.close()


@pytest.mark.parametrize("sizes,act", [((96, 50, 20, 31), "mish"), ((96, 400, 31), "swish"), ((96, 31), "identity"),
                                        ((96, 64, 48, 40, 31), "tanh"), ((96, 400, 400, 31), "leakyrelu")])
def test_rhs_architectures(opz, oracle, ctx, sizes, act):
    actc = {"identity": 0, "relu": 1, "tanh": 2, "mish": 3, "swish": 4, "leakyrelu": 5}[act]
    m = oracle.make_model(sizes, actc, seed=11)
    p = oracle.Params(flags=oracle.FLAG_MPP | oracle.FLAG_RI_EPS)
    rng = np.random.default_rng(5)
    B = 37  # ragged: not a multiple of any tile
    x = rng.normal(size=(B, 96)).astype(np.float32)
    _, bc = oracle.synthetic_columns(B, p)
    model = product_model(opz, m)
    k = opz.rhs(model, product_params(opz, p), x, bc, 0.0)
    kr = oracle.rhs(x, bc, 0.0, m, p)
    for v in range(3):
        sl = slice(32 * v, 32 * (v + 1))
        assert rel_err(k[:, sl], kr[:, sl]) < 5e-5
    model.close()


def test_mlp_forward_matches_oracle(opz, oracle, ctx):
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=2)
    model = product_model(opz, m)
    x = np.random.default_rng(0).normal(size=(130, 96)).astype(np.float32)
    for net, w in enumerate((m.w_uw, m.w_vw, m.w_wT)):
        y = opz.mlp_forward(model, net, x)
        yr = oracle.mlp_forward(x, w, m.sizes, m.act)
        assert rel_err(y, yr) < 1e-5
    model.close()


@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_short_rollout_all_schemes(opz, oracle, ctx, scheme):
    p = oracle.Params()
    B = 9
    x0, bc = oracle.synthetic_columns(B, p)
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1)
    model = product_model(opz, m)
    dt = 30.0 / p.tau
    out = opz.rollout_forward(model, product_params(opz, p), x0, bc, scheme, dt, 40, 10)
    ref = oracle.rollout(x0, bc, m, p, scheme, dt, 40, 10)
    assert out.shape == ref.shape == (B, 5, 96)
    assert np.array_equal(out[:, 0], x0)  # snapshot 0 is the initial state (saveat=ts)
    assert oracle.profile_error(out, ref, 32) < 2e-5
    # final-state-only mode
    fin = opz.rollout_forward(model, product_params(opz, p), x0, bc, scheme, dt, 40, 0)
    assert fin.shape == (B, 1, 96)
    assert np.array_equal(fin[:, 0], out[:, -1])
    model.close()


def test_config1_full_two_day_rollout_fp32(opz, oracle, golden_dir, ctx):
    """BASELINE configs[0]: construct_NN + mPP + CA, 2 days, RK4 dt=30 s (5760 steps), FP32 path
    within 1e-4 of the FP32 CPU solve at every saved snapshot."""
    g = np.load(f"{golden_dir}/config1_rk4_2day.npz")
    p = oracle.Params()
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    model = product_model(opz, m)
    out = opz.rollout_forward(model, product_params(opz, p), g["x0"], g["bc"], opz.SCHEME_RK4, float(g["dt_nd"]),
                              int(g["n_steps"]), int(g["save_every"]))
    err = oracle.profile_error(out, g["sol32"], 32)
    floor = float(g["floor"])
    print(f"config1 fp32 rollout: err vs fp32 oracle {err:.3e} (fp32-vs-fp64 oracle floor {floor:.3e})")
    assert np.isfinite(out).all()
    assert err < TOL_FP32
    model.close()


def test_small_mish_diurnal_heun_golden(opz, oracle, golden_dir, ctx):
    g = np.load(f"{golden_dir}/small_mish_diurnal_heun.npz")
    p = oracle.Params(tau=691200.0, flags=oracle.FLAG_MPP | oracle.FLAG_RI_EPS | oracle.FLAG_BC_MINUS_SCALE0 | oracle.FLAG_DIURNAL)
    m = oracle.make_model((96, 50, 20, 31), oracle.ACT_MISH, seed=1, out_scale=0.1)
    model = product_model(opz, m)
    out = opz.rollout_forward(model, product_params(opz, p), g["x0"], g["bc"], opz.SCHEME_RK2, float(g["dt_nd"]),
                              int(g["n_steps"]), int(g["save_every"]))
    err = oracle.profile_error(out, g["sol32"], 32)
    print(f"small mish diurnal heun: err {err:.3e}")
    assert err < TOL_FP32
    model.close()


def test_ensemble_4096_shard_invariance(opz, oracle, ctx):
    """BASELINE configs[1] size: 4,096 columns.  Columns are independent, so any subset solved alone
    must reproduce the ensemble BIT-exactly (this is what makes column sharding communication-free),
    and a sample of columns must match the oracle."""
    p = oracle.Params()
    B = 4096
    x0, bc = oracle.synthetic_columns(B, p)
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    full = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, 8, 4)
    assert np.isfinite(full).all()
    pick = np.array([0, 1, 31, 32, 63, 64, 1000, 2047, 2048, 4095])
    sub = opz.rollout_forward(model, pp, x0[pick], bc[pick], opz.SCHEME_RK4, dt, 8, 4)
    assert np.array_equal(sub, full[pick])
    lo, hi = 1024, 1024 + 517  # a ragged shard, as a rank of a multi-GPU run would own
    shard = opz.rollout_forward(model, pp, x0[lo:hi], bc[lo:hi], opz.SCHEME_RK4, dt, 8, 4)
    assert np.array_equal(shard, full[lo:hi])
    ref = oracle.rollout(x0[pick], bc[pick], m, p, oracle.SCHEME_RK4, dt, 8, 4)
    assert oracle.profile_error(full[pick], ref, 32) < 2e-5
    model.close()


def test_edge_cases_and_errors(opz, oracle, ctx):
    p = oracle.Params()
    m = oracle.make_model((96, 50, 20, 31), oracle.ACT_MISH, seed=1)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    x0, bc = oracle.synthetic_columns(3, p)
    # empty ensemble
    out = opz.rollout_forward(model, pp, x0[:0], bc[:0], opz.SCHEME_RK4, 1e-4, 4, 2)
    assert out.shape == (0, 3, 96)
    # zero steps returns the initial state
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, 1e-4, 0, 1)
    assert np.array_equal(out[:, 0], x0)
    # ts not multiples of dt
    with pytest.raises(opz.OpzError) as e:
        opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, 1e-4, 5, 2)
    assert e.value.code == opz.OPZ_EINVAL
    # nz mismatch
    bad = product_params(opz, p)
    bad.nz = 16
    with pytest.raises(opz.OpzError):
        opz.rollout_forward(model, bad, x0, bc, opz.SCHEME_RK4, 1e-4, 4, 2)
    # unknown scheme
    with pytest.raises(opz.OpzError):
        opz.rollout_forward(model, pp, x0, bc, 7, 1e-4, 4, 2)
    # IEEE hazard of the reference kept: S2 = 0 and dT/dz = 0 with eps = 0 gives NaN
    xz = np.zeros((1, 96), dtype=np.float32)
    k = opz.rhs(model, pp, xz, bc[:1], 0.0)
    kr = oracle.rhs(xz, bc[:1], 0.0, m, p)
    assert np.array_equal(np.isnan(k), np.isnan(kr))
    model.close()


def test_reference_shaped_api(opz, oracle, ctx):
    """solve_NDE_mutating / predict_NDE / predict_flux with the reference's argument lists."""
    from types import SimpleNamespace as NS
    p = oracle.Params()
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1)
    from helpers import chains_from_oracle_model
    uw_NN, vw_NN, wT_NN = chains_from_oracle_model(opz, m)
    sc = {n: opz.ZeroMeanUnitVarianceScaling(p.mu[i], p.sigma[i]) for i, n in enumerate(("u", "v", "T", "uw", "vw", "wT"))}
    scalings = NS(**sc)
    constants = NS(H=p.H, τ=p.tau, f=p.f, Nz=32, g=p.g, α=p.alpha, ν0=p.nu0, ν_m=p.nu_m, Riᶜ=p.Ric, ΔRi=p.dRi, Pr=p.Pr, κ=p.kappa)
    derivatives = NS(cell=opz.Dᶜ(32, 1 / 32), face=opz.Dᶠ(32, 1 / 32))
    x0, bc = oracle.synthetic_columns(4, p)
    BCs = NS(uw=NS(bottom=bc[3, 0], top=bc[3, 1]), vw=NS(bottom=bc[3, 2], top=bc[3, 3]), wT=NS(bottom=bc[3, 4], top=bc[3, 5]))
    dt = 30.0 / p.tau
    ts = np.arange(0, 5) * (10 * dt)
    conditions = NS(convective_adjustment=True)
    sol = opz.solve_NDE_mutating(uw_NN, vw_NN, wT_NN, scalings, constants, BCs, derivatives, x0[3], ts, opz.RK4(dt), conditions)
    assert sol.shape == (96, 5)
    ref = oracle.rollout(x0[3:4], bc[3:4], m, p, oracle.SCHEME_RK4, dt, 40, 10)
    assert oracle.profile_error(sol.T[None], ref, 32) < 2e-5
    # training-flavour RHS
    cond_t = NS(modified_pacanowski_philander=True, convective_adjustment=False, zero_weights=True, smooth_NN=False, smooth_Ri=False)
    k = opz.predict_NDE(uw_NN, vw_NN, wT_NN, x0[3], BCs, cond_t, scalings, constants, derivatives, None)
    pt = oracle.Params(flags=oracle.FLAG_MPP | oracle.FLAG_RI_EPS | oracle.FLAG_BC_MINUS_SCALE0)
    kr = oracle.rhs(x0[3:4], bc[3:4], 0.0, m, pt)[0]
    assert rel_err(k, kr) < 5e-5
    uw, vw, wT = opz.predict_flux(uw_NN, vw_NN, wT_NN, x0[3], BCs, cond_t, scalings, constants, derivatives, None)
    assert uw.shape == (33,)


def test_host_buffer_rollout_slab_pipeline_is_bit_exact(opz, oracle):
    """opz_rollout_forward pipelines large ensembles in slabs (copy-in / step / copy-out on three streams).  Columns are
    independent, so the result must equal the one-launch device-resident call bit for bit, and a sample of columns
    must sit on the oracle."""
    import torch
    p = oracle.Params()
    ctx = opz.default_context()
    B = 3 * 8 * 148 * 128 + 4321            # four slabs, the last one ragged
    x0s, bcs = oracle.synthetic_columns(4096, p)
    reps = (B + 4095) // 4096
    x0 = np.ascontiguousarray(np.tile(x0s, (reps, 1))[:B])
    bc = np.ascontiguousarray(np.tile(bcs, (reps, 1))[:B])
    x0[:, :32] += np.linspace(0.0, 1e-3, B, dtype=np.float32)[:, None]   # no two columns alike
    m = oracle.make_model((96, 400, 400, 31), oracle.ACT_RELU, seed=0, out_scale=0.1, bias_scale=0.2)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    for prec in (opz.PREC_FP16, opz.PREC_FP32):
        n_steps, save_every = (2, 1) if prec == opz.PREC_FP16 else (1, 0)
        host = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, n_steps, save_every, precision=prec)
        xd, bd = torch.from_numpy(x0).cuda(), torch.from_numpy(bc).cuda()
        od = torch.empty(host.shape, dtype=torch.float32, device="cuda")
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        opz.rollout_forward_dev(model, pp, xd.data_ptr(), bd.data_ptr(), B, opz.SCHEME_RK4, dt, n_steps, save_every, od.data_ptr(),
                                precision=prec)
        torch.cuda.synchronize()
        assert np.array_equal(host, od.cpu().numpy())
        pick = np.array([0, 1, 148 * 128 * 8 - 1, 148 * 128 * 8, 2 * 148 * 128 * 8 + 5, B - 1])
        ref = oracle.rollout(x0[pick], bc[pick], m, p, oracle.SCHEME_RK4, dt, n_steps, save_every if save_every else n_steps)
        if not save_every:
            ref = ref[:, -1:]   # final state only
        assert oracle.profile_error(host[pick], ref, 32) < (1e-3 if prec == opz.PREC_FP16 else 1e-4)
    model.close()


@pytest.mark.parametrize("diurnal", [False, True])
def test_profile_diagnostics_match_oracle(opz, oracle, diurnal):
    """SURVEY §8 f-2 (NDE_profile, training_postprocessing.jl:396-450): fluxes at every saved snapshot, their NN-free part,
    Ri profiles and loss_per_tstep, computed on the device from a finished rollout, against the oracle's predict_flux."""
    O = oracle
    flags = O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0 | ((O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0) if diurnal else 0)
    p = O.Params(flags=flags)
    B, n_steps, save_every = 37, 40, 8
    x0, bc = O.synthetic_columns(B, p, diurnal=diurnal)
    x0 = (x0 + 0.02 * np.random.default_rng(2).normal(size=x0.shape)).astype(np.float32)
    m = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=3, out_scale=0.1, bias_scale=0.1)
    mz = O.Model(m.sizes, m.act, np.zeros_like(m.w_uw), np.zeros_like(m.w_vw), np.zeros_like(m.w_wT))
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 30.0 / p.tau
    sol = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, n_steps, save_every)
    n_save = sol.shape[1]
    target = O.rollout(x0, bc, O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=9, out_scale=0.1), p, O.SCHEME_RK4, dt, n_steps, save_every)
    d = opz.profile_diagnostics(model, pp, sol, bc, t0=0.0, dt_save=save_every * dt, target=target)
    assert d["fluxes"].shape == (B, n_save, 3, 33) and d["Ri"].shape == (B, n_save, 33)
    for k in range(n_save):
        t = k * save_every * dt
        uw, vw, wT, diag = O.predict_flux(sol[:, k], bc, t, m, p, return_diag=True)
        ref = np.stack([uw, vw, wT], axis=1)
        assert rel_err(d["fluxes"][:, k], ref) < 2e-5, k
        uz, vz, wz = O.predict_flux(sol[:, k], bc, t, mz, p)
        assert rel_err(d["fluxes_mpp"][:, k], np.stack([uz, vz, wz], axis=1)) < 2e-5, k
        ri_ref = diag["Ri"]
        assert np.isnan(d["Ri"][:, k, 0]).all() and np.isnan(d["Ri"][:, k, 32]).all()          # boundary faces: 0/0
        fin = np.isfinite(ri_ref[:, 1:32])
        assert np.allclose(d["Ri"][:, k, 1:32][fin], ri_ref[:, 1:32][fin], rtol=2e-4, atol=1e-6)
        lt = np.stack([((sol[:, k, v * 32:(v + 1) * 32] - target[:, k, v * 32:(v + 1) * 32]) ** 2).mean(axis=1) for v in range(3)], axis=1)
        assert np.allclose(d["loss_per_tstep"][:, k], lt, rtol=1e-5, atol=1e-12)
    # the NN contribution is what the nets put out on the interior faces
    nn = d["fluxes"][:, :, :, 1:32] - d["fluxes_mpp"][:, :, :, 1:32]
    y = np.stack([O.mlp_forward(sol[:, 0], w, m.sizes, m.act) for w in (m.w_uw, m.w_vw, m.w_wT)], axis=1)
    assert rel_err(nn[:, 0], y) < 1e-4
    model.close()


def test_nz128_diurnal_fp32_matches_oracle(opz, oracle):
    """BASELINE configs[4] on the FP32 path: Nz = 128 columns (state 384, 127 interior faces) with a widened MLP
    (384-256-256-127 x3) and diurnal surface forcing; parity with the oracle at the FP32 tolerance (1e-4), the single
    right-hand side, the fluxes and the batched MLP included.  (The tensor-core side of configs[4]: tests/test_gpu_layered.py.)"""
    O = oracle
    p = O.Params(Nz=128, flags=O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL | O.FLAG_DIURNAL_MINUS_SCALE0)
    B = 37
    x0, bc = O.synthetic_columns(B, p, diurnal=True)
    x0 = (x0 + 0.01 * np.random.default_rng(6).normal(size=x0.shape)).astype(np.float32)
    sizes = (384, 256, 256, 127)
    m = O.make_model(sizes, O.ACT_RELU, seed=0, out_scale=0.1, bias_scale=0.05)
    model = product_model(opz, m, Nz=128)
    pp = product_params(opz, p)
    assert model.P == O.n_params(sizes)
    dt = (30.0 / 16.0) / p.tau        # dz is 4x finer: the explicit diffusion limit shrinks 16x
    n_steps, save_every = 48, 12
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_RK4, dt, n_steps, save_every, t0=0.1)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, n_steps, save_every, t0=0.1)
    assert out.shape == (B, 5, 384) and np.isfinite(out).all()
    err = O.profile_error(out, ref, 128)
    print(f"Nz=128 diurnal FP32 rollout: max relative profile error {err:.2e}")
    assert err < 1e-4
    k, fl = opz.rhs(model, pp, x0, bc, 0.3, with_fluxes=True)
    assert rel_err(k, O.rhs(x0, bc, 0.3, m, p)) < 2e-5
    uw, vw, wT = O.predict_flux(x0, bc, 0.3, m, p)
    assert fl.shape == (B, 3, 129) and rel_err(fl, np.stack([uw, vw, wT], axis=1)) < 2e-5
    y = opz.mlp_forward(model, 2, x0)
    assert rel_err(y, O.mlp_forward(x0, m.w_wT, sizes, m.act)) < 2e-5
    model.close()


def test_imex_euler_matches_oracle_and_is_stable_at_kappa_10(opz, oracle):
    """OPZ_IMEX_EULER (explicit Euler + backward-Euler vertical diffusion, NDE_oceananigans.jl:61-101) on the
    FP32 path against the oracle's step_imex; one day at dt = 60 s with kappa = 10, where the explicit schemes need dt <= 4 s."""
    O = oracle
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0, kappa=10.0)
    B = 70
    x0, bc = O.synthetic_columns(B, p)
    x0 = (x0 + 0.02 * np.random.default_rng(5).normal(size=x0.shape)).astype(np.float32)
    m = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=2, out_scale=0.1, bias_scale=0.05)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    dt = 60.0 / p.tau
    # short rollout: FP32 tolerance of the north star
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_IMEX, dt, 48, 12)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_IMEX, dt, 48, 12)
    err = O.profile_error(out, ref, 32)
    assert np.isfinite(out).all() and err < 1e-4, err
    # one day: the convective-adjustment switch amplifies rounding (float32 vs float64 oracle: ~5e-3), so the bound is the
    # oracle's own float32 noise floor
    n_steps, save_every = 1440, 360
    out = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_IMEX, dt, n_steps, save_every)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_IMEX, dt, n_steps, save_every)
    m64 = O.Model(m.sizes, m.act, m.w_uw.astype(np.float64), m.w_vw.astype(np.float64), m.w_wT.astype(np.float64))
    ref64 = O.rollout(x0.astype(np.float64), bc.astype(np.float64), m64, p, O.SCHEME_IMEX, dt, n_steps, save_every)
    floor = O.profile_error(ref, ref64, 32)
    err_day = O.profile_error(out, ref64, 32)
    print(f"IMEX Euler, kappa = 10, dt = 60 s: 48 steps err {err:.2e}; one day err vs float64 oracle {err_day:.2e} "
          f"(float32 oracle: {floor:.2e}); |x|max {np.abs(out).max():.2f}")
    assert np.isfinite(out).all() and np.abs(out).max() < 50
    assert err_day < max(1e-4, 2.0 * floor)
    # the host mirror's stepper object; the tensor-core precisions decline the scheme (its gradient: tests/test_gpu_bptt.py)
    st = opz.IMEXEuler(dt)
    assert st.scheme == opz.SCHEME_IMEX
    for prec in (opz.PREC_FP16, opz.PREC_BF16):
        with pytest.raises(opz.OpzError) as e:
            opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_IMEX, dt, 4, 2, precision=prec)
        assert e.value.code == opz.OPZ_EUNSUP
    model.close()


def test_reference_shaped_calls_never_reuse_stale_weights(opz, oracle, ctx):
    """NDE(x, p, t, ...) rebuilds its three Chains from `p` on every call (NDE_training.jl:62-64).  The device model behind
    the reference-shaped entry points is cached per architecture and its weights are uploaded on every call: two calls
    with different parameter vectors must each match the oracle (an id()-keyed cache returned the first call's weights)."""
    import gc
    from types import SimpleNamespace as NS
    sizes = (96, 50, 20, 31)
    pt = oracle.Params(flags=oracle.FLAG_MPP | oracle.FLAG_RI_EPS)
    sc = {n: opz.ZeroMeanUnitVarianceScaling(pt.mu[i], pt.sigma[i]) for i, n in enumerate(("u", "v", "T", "uw", "vw", "wT"))}
    scalings = NS(**sc)
    constants = NS(H=pt.H, τ=pt.tau, f=pt.f, Nz=32, g=pt.g, α=pt.alpha, ν0=pt.nu0, ν_m=pt.nu_m, Riᶜ=pt.Ric, ΔRi=pt.dRi, Pr=pt.Pr)
    cond = NS(modified_pacanowski_philander=True, convective_adjustment=False, zero_weights=False, smooth_NN=False, smooth_Ri=False)
    from helpers import chains_from_oracle_model
    x0, bc = oracle.synthetic_columns(3, pt)
    x = (x0[1] + 0.05 * np.random.default_rng(3).normal(size=96)).astype(np.float32)
    res = []
    for seed in (1, 2, 3, 1):
        m = oracle.make_model(sizes, oracle.ACT_MISH, seed=seed)
        chains = chains_from_oracle_model(opz, m)
        res_w = [opz.destructure(c) for c in chains]
        NN_constructions = NS(uw=res_w[0][1], vw=res_w[1][1], wT=res_w[2][1])
        P = m.P
        NN_ranges = NS(uw=slice(0, P), vw=slice(P, 2 * P), wT=slice(2 * P, 3 * P))
        pvec = np.concatenate([m.flat(), bc[1, :6]]).astype(np.float32)
        del chains, res_w
        gc.collect()
        k = opz.NDE(x, pvec, 0.0, NN_ranges, NN_constructions, cond, scalings, constants, None, None)
        kr = oracle.rhs(x[None], bc[1:2], 0.0, m, pt)[0]
        for v in range(3):
            sl = slice(32 * v, 32 * (v + 1))
            assert rel_err(k[sl], kr[sl]) < 5e-5, (seed, v)
        res.append(k)
    assert not np.array_equal(res[0], res[1]) and not np.array_equal(res[1], res[2])
    assert np.array_equal(res[0], res[3])
    # in-place edit of a Dense's W between two solve calls is picked up as well
    m = oracle.make_model(sizes, oracle.ACT_MISH, seed=5, out_scale=0.1)
    uw_NN, vw_NN, wT_NN = chains_from_oracle_model(opz, m)
    BCs = NS(uw=NS(bottom=bc[1, 0], top=bc[1, 1]), vw=NS(bottom=bc[1, 2], top=bc[1, 3]), wT=NS(bottom=bc[1, 4], top=bc[1, 5]))
    k1 = opz.predict_NDE(uw_NN, vw_NN, wT_NN, x, BCs, cond, scalings, constants, None, None)
    uw_NN.layers[-1].W *= 3.0
    k2 = opz.predict_NDE(uw_NN, vw_NN, wT_NN, x, BCs, cond, scalings, constants, None, None)
    assert not np.array_equal(k1[:32], k2[:32]) and np.array_equal(k1[64:], k2[64:])


def test_context_destroyed_before_its_models_is_safe(opz, oracle):
    """Host garbage collectors give no destruction order: opz_ctx_destroy with live models defers the teardown to the last
    opz_model_destroy (no use-after-free of the context's device / stream)."""
    import ctypes as C
    h = C.c_void_p()
    opz.check(opz.lib.opz_ctx_create(0, C.byref(h)))
    m = oracle.make_model((96, 16, 31), oracle.ACT_RELU, seed=0)
    sizes = (C.c_int * 3)(96, 16, 31)
    mh = [C.c_void_p(), C.c_void_p()]
    ws = [np.ascontiguousarray(w, dtype=np.float32) for w in (m.w_uw, m.w_vw, m.w_wT)]
    for k in range(2):
        opz.check(opz.lib.opz_model_create(h, 32, 3, sizes, 1, ws[0].ctypes.data_as(C.c_void_p), ws[1].ctypes.data_as(C.c_void_p),
                                           ws[2].ctypes.data_as(C.c_void_p), C.byref(mh[k])))
    assert opz.lib.opz_ctx_destroy(h) == 0        # deferred
    back = np.empty(3 * m.P, dtype=np.float32)
    opz.check(opz.lib.opz_model_get_weights(mh[0], back.ctypes.data_as(C.c_void_p)))   # the context is still alive
    assert np.array_equal(back, m.flat())
    assert opz.lib.opz_model_destroy(mh[0]) == 0
    assert opz.lib.opz_model_destroy(mh[1]) == 0  # frees the context too
