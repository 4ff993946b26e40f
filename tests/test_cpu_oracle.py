"""CPU tests of the oracle (no GPU): golden vectors, the C port against the NumPy oracle, the
identities the reference's own tests hold (test/test_feature_scaling.jl:7-12,
wind_mixing/test/test_training_scaling.jl:17-19, the D^c / D^f docstring contracts of
src/differentiation_operators.jl:1-20), analytic limits, and the autograd twin."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import nde_oracle as O


def test_difference_operators_contract():
    N = 32
    Dc, Df = O.D_cell(N, 1 / N), O.D_face(N, 1 / N)
    assert Dc.shape == (N, N + 1) and Df.shape == (N + 1, N)
    assert np.all(Df[0] == 0) and np.all(Df[-1] == 0)  # zero first / last rows (differentiation_operators.jl:21-29)
    rng = np.random.default_rng(0)
    c, F = rng.normal(size=N), rng.normal(size=N + 1)
    assert np.allclose(Df @ c, O.face_gradient(c[None], N)[0])
    assert np.allclose(Dc @ F, O.cell_divergence(F[None], N)[0])
    # derivative of a linear profile is exact in the interior
    z = (np.arange(N) + 0.5) / N
    assert np.allclose((Df @ (3 * z))[1:-1], 3.0)


@pytest.mark.parametrize("shape", [(10,), (5, 5), (3, 3, 3)])
def test_feature_scaling_identities(shape):
    """test/test_feature_scaling.jl:7-12 restated."""
    rng = np.random.default_rng(1)
    data = rng.random(shape)
    s = O.ZeroMeanUnitVarianceScaling.from_data(data)
    assert s.mu == pytest.approx(data.mean()) and s.sigma == pytest.approx(data.std(ddof=1))
    sc = s(data)
    assert abs(sc.mean()) < 1e-10
    assert sc.std(ddof=1) == pytest.approx(1.0)
    assert np.allclose(s.inv()(sc), data)


def test_loss_scaling_fraction_identities():
    """wind_mixing/test/test_training_scaling.jl:17-19 re-targeted at calculate_loss_scalings (loss.jl:11-31)."""
    losses = dict(u=0.3, v=0.7, T=2.0, dudz=5.0, dvdz=1.0, dTdz=4.0)
    fr = dict(T=0.8, dTdz=0.8, profile=0.5)
    s = O.calculate_loss_scalings(losses, fr, True)
    prof_T, prof_uv = s["T"] * losses["T"], s["u"] * losses["u"] + s["v"] * losses["v"]
    assert prof_T / prof_uv == pytest.approx(fr["T"] / (1 - fr["T"]))
    grad_T, grad_uv = s["dTdz"] * losses["dTdz"], s["dudz"] * losses["dudz"] + s["dvdz"] * losses["dvdz"]
    assert grad_T / grad_uv == pytest.approx(fr["dTdz"] / (1 - fr["dTdz"]))
    assert (prof_T + prof_uv) / (grad_T + grad_uv) == pytest.approx(fr["profile"] / (1 - fr["profile"]))
    s0 = O.calculate_loss_scalings(losses, fr, False)
    assert s0["dudz"] == 0 and s0["dTdz"] == 0


def test_smoothing_filter_matches_box3():
    a = np.random.default_rng(2).normal(size=(4, 31)).astype(np.float32)
    assert np.allclose(O._box3(a), a @ O.smoothing_filter(31, 3).T, atol=1e-6)


def test_flux_divergence_telescopes():
    """Column integral of dT/dt equals -(tau/H) (s_wT/s_T) Nz (wT_top - wT_bottom)."""
    p = O.Params()
    x0, bc = O.synthetic_columns(7, p, dtype=np.float64)
    m = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=3).astype(np.float64)
    rng = np.random.default_rng(3)
    x = x0 + 0.01 * rng.normal(size=x0.shape)
    k = O.rhs(x, bc, 0.1, m, p)
    uw, vw, wT = O.predict_flux(x, bc, 0.1, m, p)
    lhs = k[:, 64:].sum(axis=1)
    rhs = -(p.tau / p.H) * (p.sigma[5] / p.sigma[2]) * p.Nz * (wT[:, -1] - wT[:, 0])
    assert np.allclose(lhs, rhs, rtol=1e-9)


def test_diffusivity_limits_and_ieee_cases():
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0)
    m = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=3)
    x0, bc = O.synthetic_columns(1, p)
    # strongly stable (Ri -> +inf, no shear): nu -> nu0
    _, _, _, d = O.predict_flux(x0, bc, 0.0, m, p, return_diag=True)
    assert np.allclose(d["nu"][:, 1:32], p.nu0, rtol=1e-5)
    # unstable stratification: Ri -> -inf: nu -> nu0 + nu_m
    xu = x0.copy()
    xu[:, 64:] = xu[:, 64:][:, ::-1]
    _, _, _, d = O.predict_flux(xu, bc, 0.0, m, p, return_diag=True)
    assert np.allclose(d["nu"][:, 1:32], p.nu0 + p.nu_m, rtol=1e-5)
    # S2 = 0 and dT/dz = 0 with eps = 0: NaN, as in the reference
    xz = np.zeros_like(x0)
    assert np.isnan(O.rhs(xz, bc, 0.0, m, p)).any()
    # the training flavour adds eps and stays finite
    pe = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS)
    assert np.isfinite(O.rhs(xz, bc, 0.0, m, pe)).all()


def test_inertial_oscillation_without_fluxes():
    """No NN, no diffusion, no stress: u + i v rotates at f (the Coriolis terms of NDE_training.jl:160-161)."""
    p = O.Params(flags=0, mu=(0.0, 0.0, 18.7, 0.0, 0.0, 0.0))
    m = O.make_model((96, 8, 31), O.ACT_RELU, seed=0, out_scale=0.0).astype(np.float64)
    x0 = np.zeros((1, 96))
    x0[:, :32] = 1.0
    bc = np.zeros((1, 8))
    n = 2000
    dt = (2 * np.pi / p.f) / p.tau / n
    sol = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, n, n)
    assert np.allclose(sol[0, -1, :64], x0[0, :64], atol=1e-9)  # one inertial period later
    q = O.rollout(x0, bc, m, p, O.SCHEME_RK4, dt, n // 4, n // 4)
    assert np.allclose(q[0, -1, :32], 0.0, atol=1e-9) and np.allclose(q[0, -1, 32:64], -1.0, atol=1e-9)


def test_time_step_convergence_order():
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    x0, bc = O.synthetic_columns(2, p, dtype=np.float64)
    m = O.make_model((96, 50, 20, 31), O.ACT_TANH, seed=5, out_scale=0.1).astype(np.float64)
    T = 3600.0 / p.tau
    ref = O.rollout(x0, bc, m, p, O.SCHEME_RK4, T / 512, 512, 512)[:, -1]
    err = {}
    for scheme in (O.SCHEME_EULER, O.SCHEME_RK2):
        e = []
        for n in (16, 32):
            e.append(np.abs(O.rollout(x0, bc, m, p, scheme, T / n, n, n)[:, -1] - ref).max())
        err[scheme] = np.log2(e[0] / e[1])
    assert 0.8 < err[O.SCHEME_EULER] < 1.3
    assert 1.7 < err[O.SCHEME_RK2] < 2.8


def test_golden_vectors_pin_the_oracle(golden_dir):
    g = np.load(f"{golden_dir}/rhs_flavours.npz")
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=3)
    for name in [k[2:] for k in g.files if k.startswith("k_")]:
        p = O.Params(flags=int(g["flags_" + name]))
        k = O.rhs(g["x"], g["bc"], 0.37, m, p)
        assert np.allclose(k, g["k_" + name], rtol=2e-5, atol=1e-4 * np.abs(g["k_" + name]).max()), name
    g2 = np.load(f"{golden_dir}/small_mish_diurnal_heun.npz")
    p2 = O.Params(tau=691200.0, flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL)
    m2 = O.make_model((96, 50, 20, 31), O.ACT_MISH, seed=1, out_scale=0.1)
    sol = O.rollout(g2["x0"], g2["bc"], m2, p2, O.SCHEME_RK2, float(g2["dt_nd"]), 288, 288)
    assert O.profile_error(sol, g2["sol32"][:, :2], 32) < 1e-5


def test_c_port_matches_numpy_oracle():
    for sizes, act, flags in [((96, 400, 400, 31), O.ACT_RELU, O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0),
                              ((96, 50, 20, 31), O.ACT_MISH, O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI),
                              ((96, 400, 31), O.ACT_SWISH, O.FLAG_MPP | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_DIURNAL)]:
        p = O.Params(flags=flags)
        x0, bc = O.synthetic_columns(5, p, diurnal=True)
        m = O.make_model(sizes, act, seed=7, out_scale=0.1)
        for scheme in (0, 1, 2):
            a = CO.rollout(x0, bc, m, p, scheme, 30.0 / p.tau, 20, 5)
            b = O.rollout(x0, bc, m, p, scheme, 30.0 / p.tau, 20, 5)
            assert O.profile_error(a, b, 32) < 2e-5
    fin = CO.rollout(x0, bc, m, p, 2, 30.0 / p.tau, 20, 0)
    assert np.array_equal(fin[:, 0], CO.rollout(x0, bc, m, p, 2, 30.0 / p.tau, 20, 5)[:, -1])


def test_c_port_config1_golden(golden_dir):
    """The full-length BASELINE configs[0] rollout: the C port (the cpu_baseline of bench.py) reproduces
    the committed FP32 NumPy solution, and the FP32-vs-FP64 noise floor is what the file records."""
    g = np.load(f"{golden_dir}/config1_rk4_2day.npz")
    p = O.Params()
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    sol = CO.rollout(g["x0"][:2], g["bc"][:2], m, p, O.SCHEME_RK4, float(g["dt_nd"]), int(g["n_steps"]), int(g["save_every"]))
    assert O.profile_error(sol, g["sol32"][:2], 32) < 1e-4
    assert O.profile_error(g["sol32"], g["sol64"], 32) == pytest.approx(float(g["floor"]), rel=1e-3)
    assert float(g["floor"]) < 1e-3  # FP32 round-off over 23 040 RHS evaluations: 2.2e-4, the context for the 1e-4 GPU-vs-FP32-CPU bar


def test_reduced_precision_emulation_helpers():
    a = np.array([1.0, 1.0 + 2 ** -8, 1.0 + 3 * 2 ** -9, 3.14159265, 70000.0, -1e-8], dtype=np.float32)
    b = O.quant_bf16(a)
    assert b[0] == 1.0 and b[1] == 1.0 and b[2] == np.float32(1.0 + 2 ** -7)  # 7 mantissa bits, ties to even
    h = O.quant_fp16(a)
    assert h[4] == 65504.0 and abs(h[3] - a[3]) < 2 ** -10 * 4
    assert np.all(np.abs(O.quant_tf32(a) - a) <= np.abs(a) * 2 ** -11)


def test_autograd_twin_matches_oracle_and_finite_differences():
    from oracle import nde_oracle_torch as OT
    import torch
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS)
    m = O.make_model((96, 24, 16, 31), O.ACT_MISH, seed=1, out_scale=0.1)
    x0, bc = O.synthetic_columns(3, p)
    dt = 60.0 / p.tau
    x64, b64, m64 = x0.astype(np.float64), bc.astype(np.float64), m.astype(np.float64)
    ref = O.rollout(x64, b64, m64, p, O.SCHEME_RK4, dt, 8, 4)
    sol = OT.rollout(torch.tensor(x64), torch.tensor(b64), torch.tensor(m64.flat()), m.sizes, m.act, p, O.SCHEME_RK4, dt, 8, 4)
    assert np.abs(sol.numpy() - ref).max() < 1e-12
    tgt = O.rollout(x64, b64, O.make_model(m.sizes, m.act, seed=2, out_scale=0.1).astype(np.float64), p, O.SCHEME_RK4, dt, 8, 4)
    lw = [1, 1, 1, 5e-3, 5e-3, 5e-3]
    tot, comps, g = OT.loss_and_grad(x0, bc, m, p, O.SCHEME_RK4, dt, 8, 4, tgt, lw)
    c = O.loss_components(ref, tgt, 32)
    assert tot == pytest.approx(sum(a * b for a, b in zip(lw, c)), rel=1e-10)
    flat = m64.flat()

    def L(f):
        s = O.rollout(x64, b64, O.Model.from_flat(f, m.sizes, m.act), p, O.SCHEME_RK4, dt, 8, 4)
        return sum(a * b for a, b in zip(lw, O.loss_components(s, tgt, 32)))
    for idx in np.argsort(-np.abs(g))[:4]:
        fp, fm = flat.copy(), flat.copy()
        fp[idx] += 1e-6
        fm[idx] -= 1e-6
        assert (L(fp) - L(fm)) / 2e-6 == pytest.approx(g[idx], rel=1e-5)


@pytest.mark.parametrize("flags,scheme,sizes,act", [
    (O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0, O.SCHEME_RK4, (96, 50, 20, 31), O.ACT_MISH),
    (O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0, O.SCHEME_RK4, (96, 64, 31), O.ACT_RELU),
    (O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_DIURNAL, O.SCHEME_RK2, (96, 40, 30, 31), O.ACT_SWISH),
    (0, O.SCHEME_EULER, (96, 40, 30, 20, 31), O.ACT_TANH),
])
def test_hand_adjoint_matches_autograd_twin(flags, scheme, sizes, act):
    """oracle/nde_adjoint.py (the recurrences the CUDA BPTT kernels implement) against torch.autograd."""
    from oracle import nde_adjoint as A
    from oracle import nde_oracle_torch as OT
    p = O.Params(flags=flags)
    B = 5
    x0, bc = O.synthetic_columns(B, p, dtype=np.float64, diurnal=True)
    x0 = x0 + 0.05 * np.random.default_rng(1).normal(size=x0.shape)
    m = O.make_model(sizes, act, seed=1, out_scale=0.1, dtype=np.float64)
    dt, n_steps, se = 30.0 / p.tau, 12, 4
    tgt = O.rollout(x0, bc, O.make_model(sizes, act, seed=2, out_scale=0.1, dtype=np.float64), p, scheme, dt, n_steps, se)
    lw = [1, 0.7, 1.3, 5e-3, 4e-3, 6e-3]
    L1, c1, g1 = OT.loss_and_grad(x0, bc, m, p, scheme, dt, n_steps, se, tgt, lw, B_total=7)
    L2, c2, g2 = A.loss_and_grad(x0, bc, m, p, scheme, dt, n_steps, se, tgt, lw, B_total=7)
    assert L2 == pytest.approx(L1, rel=1e-12)
    assert np.allclose(c1, c2, rtol=1e-12, atol=0)
    assert np.linalg.norm(g1 - g2) <= 1e-12 * np.linalg.norm(g1)


def test_bptt_fixture_pins_the_adjoint(golden_dir):
    """The committed gradient fixture (generated by the autograd twin) against the hand adjoint in float64
    and float32: the latter is the noise floor the FP32 CUDA path is judged against."""
    from oracle import nde_adjoint as A
    g = np.load(f"{golden_dir}/bptt_config4_small.npz")
    p = O.Params(tau=float(g["tau"]), flags=int(g["flags"]))
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    args = (g["x0"], g["bc"], m, p, O.SCHEME_RK4, float(g["dt_nd"]), int(g["n_steps"]), int(g["save_every"]), g["target"], g["loss_w"])
    t64, _, g64 = A.loss_and_grad(*args, dtype=np.float64)
    assert t64 == pytest.approx(float(g["total"]), rel=1e-9)
    assert np.linalg.norm(g64 - g["grad"]) <= 1e-6 * np.linalg.norm(g["grad"])
    t32, _, g32 = A.loss_and_grad(*args, dtype=np.float32)
    assert t32 == pytest.approx(float(g["total"]), rel=1e-4)
    assert np.linalg.norm(g32 - g["grad"]) <= 1e-3 * np.linalg.norm(g["grad"])


def test_twin_mpp_parameter_gradient_matches_finite_differences():
    """The autograd twin's d loss / d (nu0, nu_m, dRi, Ric, Pr) (ground truth of opz_loss_grad_mpp; the reference's DE of
    diffusivity_parameter_optimisation.jl:1-33 is the NN-free special case) against central differences in float64."""
    from oracle import nde_oracle_torch as OT
    import dataclasses
    flags = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    p = O.Params(flags=flags)
    B = 4
    x0, bc = O.synthetic_columns(B, p, dtype=np.float64)
    x0 = x0 + 0.05 * np.random.default_rng(3).normal(size=x0.shape)
    sizes = (96, 20, 31)
    m = O.make_model(sizes, O.ACT_TANH, seed=1, out_scale=0.1, dtype=np.float64)
    dt, n_steps, se = 30.0 / p.tau, 8, 4
    p_true = dataclasses.replace(p, nu_m=0.07, Pr=1.4)
    tgt = O.rollout(x0, bc, m, p_true, O.SCHEME_RK4, dt, n_steps, se)
    lw = [1, 1, 1, 1e-3, 1e-3, 1e-3]
    L0, _, _, gm = OT.loss_and_grad(x0, bc, m, p, O.SCHEME_RK4, dt, n_steps, se, tgt, lw, with_mpp=True)
    names = ("nu0", "nu_m", "dRi", "Ric", "Pr")
    for i, nme in enumerate(names):
        v = getattr(p, nme)
        h = 1e-5 * v
        Lp = OT.loss_and_grad(x0, bc, m, dataclasses.replace(p, **{nme: v + h}), O.SCHEME_RK4, dt, n_steps, se, tgt, lw)[0]
        Lm = OT.loss_and_grad(x0, bc, m, dataclasses.replace(p, **{nme: v - h}), O.SCHEME_RK4, dt, n_steps, se, tgt, lw)[0]
        assert (Lp - Lm) / (2 * h) == pytest.approx(gm[i], rel=2e-5), nme


def test_twin_flux_pretraining_gradient_matches_finite_differences():
    """Ground truth of opz_flux_loss_grad (train_NN / NN_loss, NN_training.jl:207-249): the twin's gradient of the
    supervised flux loss of one net against central differences, and zero gradient for the other two nets."""
    from oracle import nde_oracle_torch as OT
    flags = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    p = O.Params(flags=flags)
    N = 6
    x, bc = O.synthetic_columns(N, p, dtype=np.float64)
    x = x + 0.05 * np.random.default_rng(8).normal(size=x.shape)
    sizes = (96, 12, 31)
    m = O.make_model(sizes, O.ACT_SWISH, seed=1, out_scale=0.1, dtype=np.float64, bias_scale=0.1)
    tgt = O.predict_flux(x, bc, 0.0, O.make_model(sizes, O.ACT_SWISH, seed=2, out_scale=0.1, dtype=np.float64), p)[1]
    gs = 3.2e-3
    total, l1, l2, g, g_all = OT.flux_loss_and_grad(x, bc, m, p, 1, tgt, gs)
    assert total == pytest.approx(l1 + gs * l2, rel=1e-12)
    P = m.P
    assert np.all(g_all[:P] == 0) and np.all(g_all[2 * P:] == 0)
    # numpy restatement of the loss
    def L(w_vw):
        mm = O.Model(m.sizes, m.act, m.w_uw, w_vw, m.w_wT)
        F = O.predict_flux(x, bc, 0.0, mm, p)[1]
        d = (F[:, 1:] - F[:, :-1]) * 32.0 - (tgt[:, 1:] - tgt[:, :-1]) * 32.0
        return float((((F - tgt) ** 2).mean(axis=1) + gs * (d ** 2).mean(axis=1)).mean())
    assert L(m.w_vw) == pytest.approx(total, rel=1e-10)
    rng = np.random.default_rng(0)
    for idx in rng.choice(P, size=6, replace=False):
        wp, wm = m.w_vw.copy(), m.w_vw.copy()
        h = 1e-6
        wp[idx] += h
        wm[idx] -= h
        assert (L(wp) - L(wm)) / (2 * h) == pytest.approx(g[idx], rel=1e-5, abs=1e-12)


def test_optimise_mpp_host_loop_with_a_stub_gradient(monkeypatch):
    """Host logic of api.optimise_modified_pacanowski_philander without a GPU: scaled parameters p / p_initial
    (diffusivity_parameter_optimisation.jl:44-48), box [0, 10] (:197), gradient chain rule, params updated in place."""
    import opz_import
    opz = opz_import.load()
    api = opz.api
    target = np.array([2e-4, 0.05, 0.2, 0.3, 1.5])

    def fake(model, params, x0, bc, tgt, lw, scheme, dt, n_steps, save_every, **kw):
        pv = np.array([params.nu0, params.nu_m, params.dRi, params.Ric, params.Pr], dtype=np.float64)
        r = pv / target - 1.0
        return float((r ** 2).sum()), np.zeros(6, np.float32), np.zeros(1, np.float32), (2 * r / target).astype(np.float32)

    monkeypatch.setattr(api, "loss_grad_mpp", fake)
    pp = opz.OpzParams()
    pp.nu0, pp.nu_m, pp.dRi, pp.Ric, pp.Pr = 1e-4, 0.1, 0.1, 0.25, 1.0
    seen = []
    pfin, hist = api.optimise_modified_pacanowski_philander(None, pp, None, None, None, None, api.RK4(1e-4), 8, 4, [api.ADAM(5e-2)],
                                                            maxiters=300, callback=lambda p, t, c, i: seen.append(t))
    assert len(hist) == 300 and len(seen) == 300 and hist[-1] < 1e-2 * hist[0]
    assert np.allclose(pfin, target, rtol=0.1)
    assert pp.nu_m == pytest.approx(float(pfin[1]), rel=1e-6) and np.all(pfin >= 0) and np.all(pfin <= 10 * np.array([1e-4, 0.1, 0.1, 0.25, 1.0]))


def test_imex_step_conserves_converges_and_lifts_the_diffusion_limit():
    """the split explicit / implicit-diffusion step (NDE_oceananigans.jl:61-101).  (i) the implicit solve
    conserves each column integral (no-flux diffusion), (ii) it converges to the explicit RK4 solution at first order,
    (iii) with convective adjustment at kappa = 10 it stays bounded at dt = 60 s where explicit RK4 blows up."""
    flags = O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0
    p = O.Params(flags=flags, kappa=10.0)
    B = 6
    x0, bc = O.synthetic_columns(B, p, dtype=np.float64)
    m = O.make_model((96, 20, 31), O.ACT_TANH, seed=1, out_scale=0.1, dtype=np.float64)
    # (i) pure diffusion step (zero NN, zero boundary fluxes, no Coriolis)
    mz = O.Model(m.sizes, m.act, np.zeros_like(m.w_uw), np.zeros_like(m.w_vw), np.zeros_like(m.w_wT))
    import dataclasses
    pz = dataclasses.replace(p, f=0.0, flags=O.FLAG_MPP | O.FLAG_CA)
    bcz = np.zeros_like(bc)
    xr = x0 + 0.1 * np.random.default_rng(0).normal(size=x0.shape)
    x1 = O.step(xr, bcz, 0.0, 60.0 / p.tau, mz, pz, O.SCHEME_IMEX)
    for v in range(3):
        assert np.allclose(x1[:, v * 32:(v + 1) * 32].sum(axis=1), xr[:, v * 32:(v + 1) * 32].sum(axis=1), rtol=1e-12, atol=1e-12)
        assert np.ptp(x1[:, v * 32:(v + 1) * 32], axis=1).max() <= np.ptp(xr[:, v * 32:(v + 1) * 32], axis=1).max() + 1e-12   # maximum principle
    # (ii) first-order convergence towards RK4 with a small step (kappa = 1 so that the explicit reference is stable)
    p1 = dataclasses.replace(p, kappa=1.0)
    T_end = 1800.0 / p.tau
    ref = O.rollout(x0, bc, m, p1, O.SCHEME_RK4, T_end / 3600, 3600, 3600)[:, -1]
    errs = []
    for n in (60, 120, 240):
        sol = O.rollout(x0, bc, m, p1, O.SCHEME_IMEX, T_end / n, n, n)[:, -1]
        errs.append(np.abs(sol - ref).max())
    assert errs[1] < 0.65 * errs[0] and errs[2] < 0.65 * errs[1], errs
    # (iii) kappa = 10, dt = 60 s (dt kappa / dz^2 = 9.4): the implicit solve obeys the maximum principle step after step on a
    # noisy (convectively unstable in places) column, explicit RK4 at the same step amplifies the noise without bound
    xi = xr.copy()
    for _ in range(30):
        xi = O.step(xi, bcz, 0.0, 60.0 / p.tau, mz, pz, O.SCHEME_IMEX)
    assert np.isfinite(xi).all() and np.abs(xi).max() <= np.abs(xr).max() + 1e-9
    with np.errstate(all="ignore"):
        xe = xr.copy()
        for _ in range(30):
            xe = O.step(xe, bcz, 0.0, 60.0 / p.tau, mz, pz, O.SCHEME_RK4)
    assert (not np.isfinite(xe).all()) or np.abs(xe).max() > 1e3 * np.abs(xr).max()
    # and a full day of the forced NDE at dt = 60 s stays bounded
    n = 1440
    sol = O.rollout(x0.astype(np.float32), bc.astype(np.float32), O.make_model((96, 20, 31), O.ACT_TANH, seed=1, out_scale=0.1), p,
                    O.SCHEME_IMEX, 60.0 / p.tau, n, n)[:, -1]
    assert np.isfinite(sol).all() and np.abs(sol).max() < 50
