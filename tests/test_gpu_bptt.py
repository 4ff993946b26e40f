"""GPU parity of opz_loss_grad (loss + exact discrete-adjoint gradient, FP32) against the float64
autograd twin (oracle/nde_oracle_torch.py) and the committed gradient fixture.  Everything goes
through the C ABI.  Tolerance (north_star): gradients within 1e-3 relative (L2)."""
import os

import numpy as np
import pytest

from helpers import product_model, product_params

pytestmark = pytest.mark.gpu

TOL_GRAD = 1e-3
TOL_LOSS = 1e-4


@pytest.fixture(scope="module")
def ctx(opz):
    return opz.default_context()


def _rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _case(oracle, flags, sizes, act, B, n_steps, save_every, seed=1, tau=172800.0, dt_s=30.0, diurnal=False, jitter=0.02):
    p = oracle.Params(flags=flags, tau=tau)
    x0, bc = oracle.synthetic_columns(B, p, diurnal=diurnal)
    rng = np.random.default_rng(seed)
    x0 = (x0 + jitter * rng.normal(size=x0.shape)).astype(np.float32)
    m = oracle.make_model(sizes, act, seed=seed, out_scale=0.1)
    m_true = oracle.make_model(sizes, act, seed=seed + 100, out_scale=0.1)
    dt = dt_s / p.tau
    return p, x0, bc, m, m_true, dt


@pytest.mark.parametrize("name,flags,sizes,act,scheme,B", [
    ("training_small_mish_rk4", "train", (96, 50, 20, 31), 3, 2, 5),
    ("inference_ca_relu_rk4", "infer", (96, 400, 400, 31), 1, 2, 18),
    ("two_layer_swish_heun_diurnal", "diurnal", (96, 400, 31), 4, 1, 7),
    ("four_layer_tanh_euler_nn_only", "none", (96, 40, 30, 20, 31), 2, 0, 3),
    ("leakyrelu_ragged_two_column_tiles", "train", (96, 64, 48, 31), 5, 2, 37),
    # three hidden layers too wide for one SM: the 16-CTA cluster path with two all-gathers and the double-buffered
    # reduce-scatter channel (st.async exchanges of opz_bptt_cluster.cuh)
    ("four_layer_relu_wide_cluster", "train", (96, 256, 250, 240, 31), 1, 2, 6),
])
def test_loss_grad_matches_autograd_twin(opz, oracle, ctx, name, flags, sizes, act, scheme, B):
    from oracle import nde_oracle_torch as OT
    O = oracle
    fl = {"train": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0,
          "infer": O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0,
          "diurnal": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_DIURNAL,
          "none": 0}[flags]
    n_steps, save_every = 12, 4
    p, x0, bc, m, m_true, dt = _case(O, fl, sizes, act, B, n_steps, save_every, diurnal=(flags == "diurnal"))
    target = O.rollout(x0, bc, m_true, p, scheme, dt, n_steps, save_every)
    lw = np.array([1.0, 0.7, 1.3, 5e-3, 4e-3, 6e-3], dtype=np.float32)
    model = product_model(opz, m)
    n0 = ctx.launch_count
    total, comps, grad = opz.loss_grad(model, product_params(opz, p), x0, bc, target, lw, scheme, dt, n_steps, save_every,
                                       B_total=B + 2)
    assert ctx.launch_count > n0
    tr, cr, gr = OT.loss_and_grad(x0, bc, m, p, scheme, dt, n_steps, save_every, target, lw, B_total=B + 2)
    assert np.isfinite(grad).all()
    assert abs(total - tr) <= TOL_LOSS * abs(tr), (name, total, tr)
    assert np.allclose(comps, cr, rtol=2e-4, atol=1e-4 * abs(tr)), (name, comps, cr)
    assert _rel_l2(grad, gr) < TOL_GRAD, (name, _rel_l2(grad, gr))
    # per net and per layer block as well: a wrong small block must not hide behind a large one
    P = m.P
    off = 0
    for i in range(len(sizes) - 1):
        for n_el in (sizes[i] * sizes[i + 1], sizes[i + 1]):
            for net in range(3):
                sl = slice(net * P + off, net * P + off + n_el)
                if np.linalg.norm(gr[sl]) > 1e-12 * np.linalg.norm(gr):
                    assert _rel_l2(grad[sl], gr[sl]) < 5e-3, (name, "layer", i, "net", net, _rel_l2(grad[sl], gr[sl]))
            off += n_el
    model.close()


def test_config4_fixture(opz, oracle, golden_dir, ctx):
    """Shortened BASELINE configs[3] step (18 cases, construct_NN) against the committed float64 fixture."""
    g = np.load(os.path.join(golden_dir, "bptt_config4_small.npz"))
    O = oracle
    p = O.Params(tau=float(g["tau"]), flags=int(g["flags"]))
    m = O.make_model((96, 400, 400, 31), O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    model = product_model(opz, m)
    total, comps, grad = opz.loss_grad(model, product_params(opz, p), g["x0"], g["bc"], g["target"], g["loss_w"],
                                       opz.SCHEME_RK4, float(g["dt_nd"]), int(g["n_steps"]), int(g["save_every"]))
    assert abs(total - float(g["total"])) <= TOL_LOSS * float(g["total"])
    assert _rel_l2(grad, g["grad"]) < TOL_GRAD, _rel_l2(grad, g["grad"])
    model.close()


def test_segmented_recompute_equals_single_segment(opz, oracle, ctx, monkeypatch):
    O = oracle
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    n_steps, save_every, B = 21, 7, 6
    p, x0, bc, m, m_true, dt = _case(O, fl, (96, 50, 20, 31), 3, B, n_steps, save_every, seed=4)
    target = O.rollout(x0, bc, m_true, p, 2, dt, n_steps, save_every)
    lw = np.array([1, 1, 1, 0, 0, 0], dtype=np.float32)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    t1, c1, g1 = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, save_every)
    # budget for 5 time steps of records -> segments of 5, 5, 5, 5, 1 steps, forward sweep without records
    rec_floats = 96 + 3 * 31 + 6 * (50 + 20)
    monkeypatch.setenv("OPZ_BPTT_RECORD_BYTES", str(4 * rec_floats * B * 4 * 5 + 64))
    monkeypatch.setenv("OPZ_BPTT_GRAPH_STEPS", "3")
    t2, c2, g2 = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, save_every)
    assert t1 == pytest.approx(t2, rel=1e-6)
    assert _rel_l2(g2, g1) < 1e-5
    model.close()


def test_cluster_and_graph_paths_agree(opz, oracle, ctx, monkeypatch):
    """The 16-CTA cluster sweep (weights resident in shared memory) and the layer-parallel CUDA-graph path
    compute the same loss and gradient (different summation orders only)."""
    O = oracle
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    n_steps, save_every = 12, 4
    lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    for sizes, act, B in (((96, 400, 400, 31), 1, 18), ((96, 50, 20, 31), 3, 7), ((96, 64, 31), 4, 3), ((96, 40, 30, 20, 31), 2, 9)):
        p, x0, bc, m, m_true, dt = _case(O, fl, sizes, act, B, n_steps, save_every, seed=9)
        target = O.rollout(x0, bc, m_true, p, 2, dt, n_steps, save_every)
        model = product_model(opz, m)
        pp = product_params(opz, p)
        res = {}
        for name, env in (("cluster", {}), ("graph", {"OPZ_BPTT_NO_CLUSTER": "1"}), ("cluster4", {"OPZ_BPTT_CLUSTER_COLUMNS": "4"}),
                          ("cluster1", {"OPZ_BPTT_CLUSTER_COLUMNS": "1"})):
            for k in ("OPZ_BPTT_NO_CLUSTER", "OPZ_BPTT_CLUSTER_COLUMNS"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            res[name] = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, save_every)
        for name in ("graph", "cluster4", "cluster1"):
            assert res[name][0] == pytest.approx(res["cluster"][0], rel=1e-5), (sizes, name)
            assert _rel_l2(res[name][2], res["cluster"][2]) < 2e-4, (sizes, name, _rel_l2(res[name][2], res["cluster"][2]))
        model.close()


def test_gradient_is_additive_over_column_shards(opz, oracle, ctx):
    """The multi-GPU contract: each rank passes its shard with B_total = global batch; the all-reduce is a sum."""
    O = oracle
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    n_steps, save_every, B = 8, 4, 18
    p, x0, bc, m, m_true, dt = _case(O, fl, (96, 50, 20, 31), 3, B, n_steps, save_every, seed=6)
    target = O.rollout(x0, bc, m_true, p, 2, dt, n_steps, save_every)
    lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    t, c, g = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, save_every)
    ts, gs = 0.0, np.zeros_like(g)
    for lo, hi in ((0, 7), (7, 7), (7, 18)):  # one empty shard
        ti, ci, gi = opz.loss_grad(model, pp, x0[lo:hi], bc[lo:hi], target[lo:hi], lw, 2, dt, n_steps, save_every, B_total=B)
        ts += ti
        gs += gi
    assert ts == pytest.approx(t, rel=1e-5)
    assert _rel_l2(gs, g) < 1e-5
    model.close()


def test_gradient_descends_the_loss(opz, oracle, ctx):
    """A few ADAM steps through the reference-shaped host loop reduce the profile loss."""
    O = oracle
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    n_steps, save_every, B = 16, 4, 6
    p, x0, bc, m, m_true, dt = _case(O, fl, (96, 50, 20, 31), 3, B, n_steps, save_every, seed=8)
    target = O.rollout(x0, bc, m_true, p, 2, dt, n_steps, save_every)
    lw = np.array([1, 1, 1, 0, 0, 0], dtype=np.float32)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    w = model.get_weights()
    opt = opz.ADAM(1e-3)
    losses = []
    for _ in range(6):
        model.set_weights(w)
        t, _, g = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, save_every)
        losses.append(t)
        w = opt.apply(w, g)
    assert losses[-1] < losses[0]
    model.close()


def test_loss_grad_rejects_unsupported(opz, oracle, ctx):
    O = oracle
    m = O.make_model((96, 31), 0, seed=0)
    model = product_model(opz, m)
    p = O.Params()
    x0, bc = O.synthetic_columns(2, p)
    tgt = np.zeros((2, 2, 96), dtype=np.float32)
    lw = np.ones(6, dtype=np.float32)
    with pytest.raises(opz.OpzError) as e:
        opz.loss_grad(model, product_params(opz, p), x0, bc, tgt, lw, 2, 1e-4, 2, 2)
    assert e.value.code == opz.OPZ_EUNSUP
    model.close()
    m2 = O.make_model((96, 50, 31), 1, seed=0)
    model2 = product_model(opz, m2)
    with pytest.raises(opz.OpzError) as e:   # reduced-precision BPTT is not built: the gradient contract is FP32
        opz.loss_grad(model2, product_params(opz, p), x0, bc, tgt, lw, 2, 1e-4, 2, 2, precision=opz.PREC_FP16)
    assert e.value.code == opz.OPZ_EUNSUP
    with pytest.raises(opz.OpzError):  # n_steps not a multiple of save_every
        opz.loss_grad(model2, product_params(opz, p), x0, bc, np.zeros((2, 2, 96), np.float32), lw, 2, 1e-4, 3, 2)
    model2.close()


def _rel_each(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@pytest.mark.parametrize("name,flags,sizes,act,zero_nn,B", [
    ("reference_DE_nn_free", "train", (96, 50, 20, 31), 3, True, 18),     # diffusivity_parameter_optimisation.jl:1-33
    ("construct_nn_with_ca", "infer", (96, 400, 400, 31), 1, False, 18),
    ("small_mish_training_flags", "train", (96, 50, 20, 31), 3, False, 7),
])
def test_loss_grad_mpp_matches_autograd_twin(opz, oracle, ctx, monkeypatch, name, flags, sizes, act, zero_nn, B):
    """d loss / d (nu0, nu_m, dRi, Ric, Pr) from the same backward sweep, against float64 autograd through
    the twin; on the cluster path and on the layer-parallel graph path."""
    from oracle import nde_oracle_torch as OT
    O = oracle
    fl = {"train": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0, "infer": O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0}[flags]
    n_steps, save_every, scheme = 12, 4, 2
    p, x0, bc, m, m_true, dt = _case(O, fl, sizes, act, B, n_steps, save_every, seed=5)
    if zero_nn:
        m = O.Model(m.sizes, m.act, np.zeros_like(m.w_uw), np.zeros_like(m.w_vw), np.zeros_like(m.w_wT))
    # targets from a solve with DIFFERENT diffusivity parameters (what the reference's optimisation descends towards)
    p_true = O.Params(flags=fl, tau=p.tau, nu0=2e-4, nu_m=0.06, dRi=0.15, Ric=0.3, Pr=1.3)
    target = O.rollout(x0, bc, m if zero_nn else m_true, p_true, scheme, dt, n_steps, save_every)
    lw = np.array([1.0, 0.7, 1.3, 5e-3, 4e-3, 6e-3], dtype=np.float32)
    t_ref, c_ref, g_ref, gm_ref = OT.loss_and_grad(x0, bc, m, p, scheme, dt, n_steps, save_every, target, lw, B_total=B + 1,
                                                   with_mpp=True)
    assert np.all(np.abs(gm_ref) > 0)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    for path, env in (("cluster", {}), ("graph", {"OPZ_BPTT_NO_CLUSTER": "1"})):
        monkeypatch.delenv("OPZ_BPTT_NO_CLUSTER", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        total, comps, grad, gm = opz.loss_grad_mpp(model, pp, x0, bc, target, lw, scheme, dt, n_steps, save_every, B_total=B + 1)
        rel = _rel_each(gm, gm_ref)
        print(f"{name} [{path}]: d loss/d(nu0, nu_m, dRi, Ric, Pr) = {gm}, rel err {rel}")
        assert total == pytest.approx(t_ref, rel=TOL_LOSS)
        if not zero_nn:
            assert _rel_l2(grad, g_ref) < TOL_GRAD
        # the tolerance of the north star (1e-3 relative, L2) on the gradient in the reference's scaled form p / p_initial
        # (diffusivity_parameter_optimisation.jl:44-48), plus a per-component bound (FP32 cancellation in the smallest one)
        pv = np.array([p.nu0, p.nu_m, p.dRi, p.Ric, p.Pr])
        assert _rel_l2(gm * pv, gm_ref * pv) < TOL_GRAD, (path, rel)
        assert np.all(rel < 3e-3), (path, rel)
        # the plain entry point returns the same loss and NN gradient
        t2, c2, g2 = opz.loss_grad(model, pp, x0, bc, target, lw, scheme, dt, n_steps, save_every, B_total=B + 1)
        assert t2 == pytest.approx(total, rel=1e-6) and _rel_l2(g2, grad) < 1e-5 or zero_nn
    # needs the mPP flag
    p_off = O.Params(flags=O.FLAG_BC_MINUS_SCALE0, tau=p.tau)
    with pytest.raises(opz.OpzError) as e:
        opz.loss_grad_mpp(model, product_params(opz, p_off), x0, bc, target, lw, scheme, dt, n_steps, save_every)
    assert e.value.code == opz.OPZ_EINVAL
    model.close()


def test_optimise_modified_pacanowski_philander_descends(opz, oracle, ctx):
    """Host loop of optimise_modified_pacanowski_philander (diffusivity_parameter_optimisation.jl:36-231) on the NN-free
    model: ADAM on the scaled parameters with libopz gradients lowers the loss towards the generating parameters."""
    O = oracle
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    n_steps, save_every = 24, 6
    p, x0, bc, m, _, dt = _case(O, fl, (96, 50, 20, 31), 3, 18, n_steps, save_every, seed=3)
    m = O.Model(m.sizes, m.act, np.zeros_like(m.w_uw), np.zeros_like(m.w_vw), np.zeros_like(m.w_wT))
    p_true = O.Params(flags=fl, tau=p.tau, nu0=p.nu0, nu_m=p.nu_m * 1.5, dRi=p.dRi, Ric=p.Ric, Pr=p.Pr * 1.2)
    target = O.rollout(x0, bc, m, p_true, 2, dt, n_steps, save_every)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    lw = np.array([1, 1, 1, 0, 0, 0], dtype=np.float32)
    pfin, hist = opz.optimise_modified_pacanowski_philander(model, pp, x0, bc, target, lw, opz.RK4(dt), n_steps, save_every,
                                                            [opz.ADAM(2e-2)], maxiters=40)
    print("mPP optimisation: loss", hist[0], "->", hist[-1], "parameters", pfin)
    assert hist[-1] < 0.5 * hist[0]
    assert np.all(pfin >= 0) and pp.nu_m == pytest.approx(float(pfin[1]), rel=1e-6)
    model.close()


@pytest.mark.parametrize("which,sizes,act,flags,N", [
    (0, (96, 400, 400, 31), 1, "train", 300),
    (2, (96, 50, 20, 31), 3, "train_nozw", 1),        # one sample: the reference's Flux.train! step
    (1, (96, 64, 31), 4, "none", 77),
])
def test_flux_loss_grad_matches_autograd_twin(opz, oracle, ctx, which, sizes, act, flags, N):
    """supervised flux pre-training loss of one net and its gradient (NN_training.jl:207-249) against
    float64 autograd through the twin's flux."""
    from oracle import nde_oracle_torch as OT
    O = oracle
    fl = {"train": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0, "train_nozw": O.FLAG_MPP | O.FLAG_RI_EPS, "none": 0}[flags]
    p = O.Params(flags=fl)
    x0, bc = O.synthetic_columns(N, p)
    rng = np.random.default_rng(4)
    x = (x0 + 0.05 * rng.normal(size=x0.shape)).astype(np.float32)
    m = O.make_model(sizes, act, seed=2, out_scale=0.1, bias_scale=0.1)
    m_true = O.make_model(sizes, act, seed=12, out_scale=0.1)
    target = np.ascontiguousarray(O.predict_flux(x, bc, 0.0, m_true, p)[which]).astype(np.float32)   # "data": another model's fluxes
    gs = 1e-4 * 32.0    # the reference's gradient_scaling (:209) in our nondimensional z (D^c multiplies by Nz)
    t_ref, l1_ref, l2_ref, g_ref, g_all = OT.flux_loss_and_grad(x, bc, m, p, which, target, gs)
    others = np.delete(g_all.reshape(3, -1), which, axis=0)
    assert np.all(others == 0)                         # the other two nets do not enter
    model = product_model(opz, m)
    total, l1, l2, grad = opz.flux_loss_grad(model, product_params(opz, p), which, x, bc, target, gs)
    print(f"flux pre-training net {which} {sizes}: loss {total:.6e} (ref {t_ref:.6e}), grad rel err {_rel_l2(grad, g_ref):.2e}")
    assert total == pytest.approx(t_ref, rel=TOL_LOSS) and l1 == pytest.approx(l1_ref, rel=TOL_LOSS) and l2 == pytest.approx(l2_ref, rel=TOL_LOSS)
    assert np.isfinite(grad).all() and np.linalg.norm(g_ref) > 0
    assert _rel_l2(grad, g_ref) < TOL_GRAD
    # a descent step along -grad lowers the loss (backtracking on the step length)
    w0 = model.get_weights()
    P = model.P
    lr, t_end = 1.0, np.inf
    for _ in range(30):
        w = w0.copy()
        w[which * P:(which + 1) * P] -= lr * grad
        model.set_weights(w)
        t_end = opz.flux_loss_grad(model, product_params(opz, p), which, x, bc, target, gs)[0]
        if t_end < total:
            break
        lr *= 0.5
    assert t_end < total
    with pytest.raises(opz.OpzError):
        opz.flux_loss_grad(model, product_params(opz, p), 3, x, bc, target, gs)
    model.close()


def test_train_NDE_host_loop_learns_on_synthetic_les_shaped_data(opz, oracle, ctx):
    """The reference-shaped training entry point (train_NDE, NDE_training.jl:167-374) end to end: an LES-shaped dataset
    object (scaled profiles of 4 simulations at 600 s spacing, fluxes, scalings, times) generated by the oracle with a
    'true' model; ADAM steps with libopz gradients reduce the profile + gradient loss, and the first evaluation agrees
    with the autograd twin."""
    from helpers import chains_from_oracle_model, les_shaped_dataset
    from oracle import nde_oracle_torch as OT
    O = oracle
    n_sim, nt = 4, 7
    sizes = (96, 50, 20, 31)
    D, p, x0, bc, data, dt, per = les_shaped_dataset(opz, O, n_sim=n_sim, nt=nt, sizes=sizes, act=O.ACT_MISH)
    m0 = O.make_model(sizes, O.ACT_MISH, seed=22, out_scale=0.1)
    uw_NN, vw_NN, wT_NN = chains_from_oracle_model(opz, m0)
    tsteps = np.arange(0, nt, 2)
    seen = []
    out = opz.train_NDE(uw_NN, vw_NN, wT_NN, D, tsteps, opz.RK4(dt), [opz.ADAM(2e-3)], 1, maxiters=12,
                        modified_pacanowski_philander=True, zero_weights=True, train_gradient=True, gradient_scaling=5e-3,
                        f=p.f, α=p.alpha, g=p.g, ν0=p.nu0, ν_m=p.nu_m, Riᶜ=p.Ric, ΔRi=p.dRi, Pr=p.Pr,
                        cb=lambda w, total, comps, lw: seen.append(total))
    uw2, vw2, wT2, history = out
    assert len(history) == 12 and len(seen) == 12
    losses = [h[0] for h in history]
    print("train_NDE losses:", ["%.4e" % v for v in losses])
    assert losses[-1] < 0.8 * losses[0] and np.isfinite(losses).all()
    # first evaluation = the twin's loss on the same problem
    target = data[:, tsteps]
    lw = np.array([1, 1, 1, 5e-3, 5e-3, 5e-3], dtype=np.float32)
    t_ref, _, _ = OT.loss_and_grad(x0, bc, m0, p, O.SCHEME_RK4, dt, per * (nt - 1), 2 * per, target, lw)
    assert losses[0] == pytest.approx(t_ref, rel=2e-4)
    # the returned chains carry the trained weights
    w_new, _ = opz.destructure(uw2)
    w_old, _ = opz.destructure(uw_NN)
    assert w_new.shape == w_old.shape and not np.array_equal(w_new, w_old)


@pytest.mark.parametrize("path", ["cluster", "graph"])
@pytest.mark.parametrize("sizes,act,B,which", [((96, 50, 20, 31), 3, 5, "both"), ((96, 400, 400, 31), 1, 18, "both"),
                                                ((96, 64, 31), 4, 3, "nn"), ((96, 40, 30, 20, 31), 2, 9, "ri")])
def test_loss_grad_with_box_filters(opz, oracle, ctx, monkeypatch, path, sizes, act, B, which):
    """smooth_NN / smooth_Ri (3-point box filters of the training path, filtering_operators.jl:1-15 at NDE_training.jl:98-102,
    121-123) in the forward solve AND in its adjoint, on both BPTT layouts, against the float64 autograd twin; the gradient
    w.r.t. the five mPP parameters as well (the filtered Richardson number enters d nu / d Ric and d nu / d dRi)."""
    from oracle import nde_oracle_torch as OT
    O = oracle
    fl = O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0
    fl |= {"both": O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI, "nn": O.FLAG_SMOOTH_NN, "ri": O.FLAG_SMOOTH_RI}[which]
    n_steps, save_every = 12, 4
    p, x0, bc, m, m_true, dt = _case(O, fl, sizes, act, B, n_steps, save_every, seed=3)
    target = O.rollout(x0, bc, m_true, p, 2, dt, n_steps, save_every)
    lw = np.array([1.0, 0.7, 1.3, 5e-3, 4e-3, 6e-3], dtype=np.float32)
    if path == "graph":
        monkeypatch.setenv("OPZ_BPTT_NO_CLUSTER", "1")
    model = product_model(opz, m)
    total, comps, grad, gm = opz.loss_grad_mpp(model, product_params(opz, p), x0, bc, target, lw, 2, dt, n_steps, save_every)
    tr, cr, gr, gmr = OT.loss_and_grad(x0, bc, m, p, 2, dt, n_steps, save_every, target, lw, with_mpp=True)
    # the filters change the answer at the size of this test (otherwise it would prove nothing)
    p0 = O.Params(flags=fl & ~(O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI), tau=p.tau)
    t0, _, g0 = OT.loss_and_grad(x0, bc, m, p0, 2, dt, n_steps, save_every, target, lw)
    print(f"{path} {sizes} filters={which}: loss {total:.6e} (twin {tr:.6e}, unfiltered {t0:.6e}); grad rel err {_rel_l2(grad, gr):.2e}, "
          f"mPP grad rel err {_rel_l2(gm, gmr):.2e}; filter effect on grad {_rel_l2(g0, gr):.2e}")
    assert _rel_l2(g0, gr) > 5 * max(_rel_l2(grad, gr), 1e-5)   # the error is far below what the filters change
    assert abs(total - tr) <= TOL_LOSS * abs(tr)
    assert np.isfinite(grad).all() and _rel_l2(grad, gr) < TOL_GRAD
    assert _rel_l2(gm, gmr) < 5e-3
    model.close()


def test_flux_loss_grad_with_smooth_NN(opz, oracle, ctx):
    """Supervised flux pre-training with the box-filtered NN output (NN_training.jl predict_* with smooth_NN)."""
    from oracle import nde_oracle_torch as OT
    O = oracle
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI)
    N, sizes = 37, (96, 50, 20, 31)
    x, bc = O.synthetic_columns(N, p)
    x = (x + 0.05 * np.random.default_rng(7).normal(size=x.shape)).astype(np.float32)
    m = O.make_model(sizes, O.ACT_MISH, seed=3, out_scale=0.1)
    target = np.random.default_rng(8).normal(size=(N, 33)).astype(np.float32) * 0.1
    model = product_model(opz, m)
    for which in range(3):
        t_ref, l1_ref, l2_ref, g_ref, _ = OT.flux_loss_and_grad(x, bc, m, p, which, target, 1e-2)
        total, l1, l2, grad = opz.flux_loss_grad(model, product_params(opz, p), which, x, bc, target, 1e-2)
        assert total == pytest.approx(t_ref, rel=TOL_LOSS)
        assert _rel_l2(grad, g_ref) < TOL_GRAD, (which, _rel_l2(grad, g_ref))
    model.close()


@pytest.mark.parametrize("mode", ["cluster", "segmented"])
def test_long_horizon_gradient_fixture(opz, oracle, golden_dir, ctx, monkeypatch, mode):
    """1 152 RK4 steps (4 608 right-hand sides) of the configs[3]-shaped training step against the committed float64
    autograd fixture (tests/golden/make_golden_bptt_long.py): the gradient the bench TIMES is a long-horizon one, and rounding
    in a 4 608-link adjoint chain is a different regime from the 12-step checks above.  `segmented` forces the record budget
    down so that the rollout is cut into recomputed time segments (the path taken when the records do not fit memory)."""
    g = np.load(os.path.join(golden_dir, "bptt_long_1152.npz"))
    O = oracle
    sizes = (96, 400, 400, 31)
    p = O.Params(tau=float(g["tau"]), flags=int(g["flags"]))
    m = O.make_model(sizes, O.ACT_RELU, seed=int(g["seed"]), out_scale=float(g["out_scale"]))
    n_steps, se, B = int(g["n_steps"]), int(g["save_every"]), g["x0"].shape[0]
    if mode == "segmented":
        rec_floats = 96 + 3 * 31 + 6 * (400 + 400)
        monkeypatch.setenv("OPZ_BPTT_RECORD_BYTES", str(4 * rec_floats * B * 4 * 100 + 64))   # records of 100 steps: 12 segments
    model = product_model(opz, m)
    total, comps, grad = opz.loss_grad(model, product_params(opz, p), g["x0"], g["bc"], g["target"], g["loss_w"],
                                       opz.SCHEME_RK4, float(g["dt_nd"]), n_steps, se)
    model.close()
    stride = int(g["stride"])
    e_sub = _rel_l2(grad[::stride], g["grad_sub"])
    from tests_golden_blocks import blocks
    bn = np.array([np.linalg.norm(grad[a:b]) for a, b in blocks(m.P, sizes)])
    e_blk = np.abs(bn - g["block_norms"]) / np.maximum(g["block_norms"], 1e-30)
    print(f"long-horizon BPTT ({mode}): loss {total:.6e} (fixture {float(g['total']):.6e}); gradient rel L2 err on the stored "
          f"sample {e_sub:.2e}; worst block-norm err {e_blk.max():.2e}; |grad| {np.linalg.norm(grad):.4e} (fixture {float(g['grad_norm']):.4e})")
    assert abs(total - float(g["total"])) <= TOL_LOSS * float(g["total"])
    assert np.isfinite(grad).all() and e_sub < TOL_GRAD
    assert e_blk.max() < 5e-3
    assert abs(np.linalg.norm(grad) - float(g["grad_norm"])) <= TOL_GRAD * float(g["grad_norm"])


@pytest.mark.parametrize("sizes,act,B,flags,kappa", [
    ((96, 50, 20, 31), 3, 5, "ca", 10.0), ((96, 400, 400, 31), 1, 18, "ca", 10.0), ((96, 64, 31), 4, 7, "train_smooth", 1.0),
    ((96, 40, 30, 20, 31), 2, 3, "diurnal", 1.0)])
def test_loss_grad_through_the_split_implicit_step(opz, oracle, ctx, sizes, act, B, flags, kappa):
    """Gradient through OPZ_IMEX_EULER (explicit NN / boundary / Coriolis step, then backward-Euler vertical diffusion with the
    diffusivities of the updated state: the reference's production step, NDE_oceananigans.jl:61-101) — including the adjoint
    of the tridiagonal solve and of its coefficients — against the float64 autograd twin, at a step (dt = 60 s, kappa = 10)
    where the explicit schemes are unstable.  The five mPP parameters' gradient goes through the same solve."""
    from oracle import nde_oracle_torch as OT
    O = oracle
    fl = {"ca": O.FLAG_MPP | O.FLAG_CA | O.FLAG_BC_MINUS_SCALE0,
          "train_smooth": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0 | O.FLAG_SMOOTH_NN | O.FLAG_SMOOTH_RI,
          "diurnal": O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_CA | O.FLAG_DIURNAL}[flags]
    n_steps, save_every = 12, 4
    p = O.Params(flags=fl, kappa=kappa)
    x0, bc = O.synthetic_columns(B, p, diurnal=(flags == "diurnal"))
    x0 = (x0 + 0.02 * np.random.default_rng(4).normal(size=x0.shape)).astype(np.float32)
    m = O.make_model(sizes, act, seed=2, out_scale=0.1)
    m_true = O.make_model(sizes, act, seed=102, out_scale=0.1)
    dt = 60.0 / p.tau
    target = O.rollout(x0, bc, m_true, p, O.SCHEME_IMEX, dt, n_steps, save_every)
    lw = np.array([1.0, 0.7, 1.3, 5e-3, 4e-3, 6e-3], dtype=np.float32)
    model = product_model(opz, m)
    pp = product_params(opz, p)
    # forward first: the FP32 rollout of the same scheme sits on the oracle (smooth_Ri inside the implicit part included)
    sol = opz.rollout_forward(model, pp, x0, bc, opz.SCHEME_IMEX, dt, n_steps, save_every)
    ref = O.rollout(x0, bc, m, p, O.SCHEME_IMEX, dt, n_steps, save_every)
    assert O.profile_error(sol, ref, 32) < 2e-5
    total, comps, grad, gm = opz.loss_grad_mpp(model, pp, x0, bc, target, lw, opz.SCHEME_IMEX, dt, n_steps, save_every, B_total=B + 1)
    tr, cr, gr, gmr = OT.loss_and_grad(x0, bc, m, p, O.SCHEME_IMEX, dt, n_steps, save_every, target, lw, B_total=B + 1, with_mpp=True)
    print(f"IMEX {sizes} {flags}: loss {total:.6e} (twin {tr:.6e}); grad rel err {_rel_l2(grad, gr):.2e}; mPP grad {gm} vs {gmr}, rel err {_rel_l2(gm, gmr):.2e}")
    assert abs(total - tr) <= TOL_LOSS * abs(tr)
    assert np.isfinite(grad).all() and _rel_l2(grad, gr) < TOL_GRAD
    assert _rel_l2(gm, gmr) < 5e-3
    # the implicit part matters: the gradient of the purely explicit Euler step differs
    _, _, g_ex = OT.loss_and_grad(x0, bc, m, O.Params(flags=fl & ~(O.FLAG_MPP | O.FLAG_CA | O.FLAG_SMOOTH_RI), kappa=kappa), O.SCHEME_EULER, dt, n_steps, save_every, target, lw, B_total=B + 1)
    assert _rel_l2(g_ex, gr) > 20 * TOL_GRAD
    model.close()
