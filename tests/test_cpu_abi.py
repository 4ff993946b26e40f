"""CPU tests of the boundary (no GPU, no compute calls): libopz.so loads, exports every symbol that
include/opz.h declares, mirrors the header's constants, and fails loudly without an sm_100 device.
Also the host-side mirror of the reference interface (api.py)."""
import ctypes as C
import os
import re
from types import SimpleNamespace as NS

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_text():
    return open(os.path.join(ROOT, "include", "opz.h")).read()


def test_every_declared_symbol_is_exported(opz):
    names = set(re.findall(r"\b(opz_[a-z_0-9]+)\s*\(", header_text()))
    assert len(names) >= 19
    lib = C.CDLL(opz.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} is declared in include/opz.h but not exported by libopz.so"
    assert names == set(opz.EXPORTED_SYMBOLS)
    assert opz.lib.opz_version() == int(re.search(r"#define OPZ_VERSION (\d+)", header_text()).group(1))


def test_header_constants_are_mirrored(opz):
    h = header_text()
    defs = dict(re.findall(r"#define (OPZ_[A-Z0-9_]+) \(?(-?\d+)u?\)?(?:\s|$)", h))
    shifts = dict(re.findall(r"#define (OPZ_FLAG_[A-Z0-9_]+) \(1u << (\d+)\)", h))
    for k, v in shifts.items():
        assert getattr(opz, k.replace("OPZ_", "")) == 1 << int(v), k
    for k in ("OPZ_EINVAL", "OPZ_ENODEV", "OPZ_ECUDA", "OPZ_ENOMEM", "OPZ_EUNSUP"):
        assert getattr(opz, k) == int(defs[k])
    for k, name in (("OPZ_EULER", "SCHEME_EULER"), ("OPZ_RK2", "SCHEME_RK2"), ("OPZ_RK4", "SCHEME_RK4"),
                    ("OPZ_PREC_FP32", "PREC_FP32"), ("OPZ_PREC_BF16", "PREC_BF16"), ("OPZ_PREC_FP16", "PREC_FP16"),
                    ("OPZ_ACT_RELU", "ACT_RELU"), ("OPZ_ACT_MISH", "ACT_MISH"), ("OPZ_ACT_LEAKYRELU", "ACT_LEAKYRELU"),
                    ("OPZ_BC_STRIDE", "BC_STRIDE")):
        assert getattr(opz, name) == int(defs[k]), k
    assert C.sizeof(opz.OpzParams) == 37 * 4


def test_no_device_no_fallback(opz):
    """On a machine without an sm_100 GPU context creation fails with OPZ_ENODEV: there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(opz.OpzError) as e:
        opz.Context(0)
    assert e.value.code == opz.OPZ_ENODEV
    assert "no CPU fallback" in str(e.value)
    # null handles are rejected before any CUDA call
    assert opz.lib.opz_ctx_synchronize(None) == opz.OPZ_EINVAL
    assert opz.lib.opz_model_n_params(None) == 0


def test_product_does_not_import_the_oracle():
    """No product source imports, includes, links or executes anything under oracle/ (comments may say the word)."""
    import re
    pkg = os.path.join(ROOT, "oceanparameterizations.jl_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle)|#\s*include\s*[\"<][^\">]*oracle|liboracle|oracle/_build|oracle\.(nde_|c_)", re.M)
    seen = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                seen += 1
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f
    assert seen > 8


def test_flux_containers_and_destructure(opz):
    rng = np.random.default_rng(0)
    c = opz.Chain(opz.Dense(96, 400, opz.relu, rng=rng), opz.Dense(400, 400, opz.relu, rng=rng), opz.Dense(400, 31, rng=rng))
    flat, re_ = opz.destructure(c)
    assert flat.dtype == np.float32 and flat.size == 96 * 400 + 400 + 400 * 400 + 400 + 400 * 31 + 31 == 211631
    # column-major vec(W1) first: element (row o, col i) at i*nout + o  (construct_NN.jl:29-34; NDE_training.jl:11-13)
    W1 = c.layers[0].W
    assert flat[5 * 400 + 7] == W1[7, 5]
    assert np.all(flat[96 * 400:96 * 400 + 400] == 0)  # Flux initialises biases to zero
    lim = np.sqrt(6 / (96 + 400))
    assert np.abs(W1).max() <= lim  # glorot_uniform
    c2 = re_(flat * 2)
    assert np.array_equal(c2.layers[1].W, c.layers[1].W * 2) and c2.sizes == [96, 400, 400, 31] and c2.activation == "relu"
    with pytest.raises(ValueError):
        re_(flat[:-1])
    with pytest.raises(ValueError):
        opz.Chain(opz.Dense(96, 10, opz.relu), opz.Dense(11, 31))


def test_scalers_operators_and_flags(opz):
    s = opz.ZeroMeanUnitVarianceScaling.from_data(np.array([1.0, 2.0, 4.0]))
    assert s.μ == pytest.approx(7 / 3) and s.σ == pytest.approx(np.std([1, 2, 4], ddof=1))
    assert np.allclose(opz.inv(s)(s(np.array([3.0, 5.0]))), [3.0, 5.0])
    assert np.allclose(opz.unscale(opz.scale(2.5, s), s), 2.5)
    D = opz.Dᶠ(4, 0.25)
    assert D.shape == (5, 4) and np.all(D[0] == 0) and np.all(D[-1] == 0) and D[2, 1] == -4 and D[2, 2] == 4
    assert opz.Dᶜ(4, 0.25).shape == (4, 5)
    f = opz.inference_flags(NS(convective_adjustment=True))
    assert f == opz.FLAG_MPP | opz.FLAG_BC_MINUS_SCALE0 | opz.FLAG_CA
    f = opz.training_flags(NS(modified_pacanowski_philander=True, convective_adjustment=False, zero_weights=True,
                              smooth_NN=True, smooth_Ri=False))
    assert f == opz.FLAG_MPP | opz.FLAG_RI_EPS | opz.FLAG_BC_MINUS_SCALE0 | opz.FLAG_SMOOTH_NN
    with pytest.raises(NotImplementedError):  # the reference's CA-only training branch reads an undefined κ
        opz.training_flags(NS(modified_pacanowski_philander=False, convective_adjustment=True))


def test_params_packing_and_time_grid(opz):
    sc = {n: opz.ZeroMeanUnitVarianceScaling(float(i), float(i + 1)) for i, n in enumerate(("u", "v", "T", "uw", "vw", "wT"))}
    constants = NS(H=256.0, τ=172800.0, f=1e-4, Nz=32, g=9.80665, α=2e-4, ν0=1e-4, ν_m=0.1, Riᶜ=0.25, ΔRi=0.1, Pr=1.0, κ=10.0)
    p = opz.make_params(constants, sc, opz.FLAG_MPP)
    assert p.nz == 32 and p.tau == 172800.0 and p.Ric == 0.25 and p.dRi == pytest.approx(0.1) and p.kappa == 10.0
    assert list(p.mu) == [0, 1, 2, 3, 4, 5] and list(p.sigma) == [1, 2, 3, 4, 5, 6]
    dt = 30.0 / 172800.0
    assert opz.api._steps_from_ts(np.arange(5) * 10 * dt, dt) == (40, 10, 0.0)
    with pytest.raises(ValueError):
        opz.api._steps_from_ts(np.array([0, 10.5 * dt, 21 * dt]), dt)
    BCs = NS(uw=NS(top=1.0, bottom=2.0), vw=NS(top=3.0, bottom=4.0), wT=NS(top=5.0, bottom=6.0))
    assert list(opz.pack_BCs(BCs, 7.0)) == [2, 1, 4, 3, 6, 5, 7, 0]


def test_adam_and_loss_helpers(opz):
    """Flux.ADAM(η, β) restated (NDE_training.jl:370 via GalacticOptim): first step is -η sign(g)."""
    opt = opz.ADAM(1e-2)
    w = np.array([1.0, -2.0, 3.0], dtype=np.float32)
    g = np.array([0.5, -0.1, 0.0], dtype=np.float32)
    w1 = opt.apply(w, g)
    assert np.allclose(w1, w - 1e-2 * np.sign(g), atol=1e-6)
    w2 = opt.apply(w1, g)
    assert np.allclose(w2, w1 - 1e-2 * np.sign(g), atol=1e-5)
    assert opz.loss(np.ones((3, 4)), np.zeros((3, 4))) == 1.0
    ls = opz.calculate_loss_scalings(NS(u=0.3, v=0.7, T=2.0, dudz=5.0, dvdz=1.0, dTdz=4.0), NS(T=0.8, dTdz=0.8, profile=0.5), True)
    assert ls.T == 1 and (ls.T * 2.0) / (ls.u * 1.0) == pytest.approx(4.0)


def test_prepare_BCs_and_pack(opz, oracle):
    """prepare_BCs (training_postprocessing.jl:35-53): scaled boundary fluxes of the FIRST data column; a
    wT_top_function makes BCs.wT.top a function of time in seconds (t -> scalings.wT(wT_top_function(t)))."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import les_shaped_dataset
    D, p, x0, bc, data, dt, per = les_shaped_dataset(opz, oracle, n_sim=2, nt=3, sizes=(96, 8, 31), act=1)
    sc = D.scalings
    BCs = opz.prepare_BCs(D, sc)
    for k, n in enumerate(("uw", "vw", "wT")):
        f = getattr(D, n)
        assert getattr(BCs, n).bottom == pytest.approx(float(sc[n](f.coarse[0, 0])), rel=1e-6)
        assert getattr(BCs, n).top == pytest.approx(float(sc[n](f.coarse[-1, 0])), rel=1e-6)
        # the scaled boundary fluxes of simulation 0, the values bc[0] was built from
        assert getattr(BCs, n).bottom == pytest.approx(float(bc[0, 2 * k]), abs=1e-5)
        assert getattr(BCs, n).top == pytest.approx(float(bc[0, 2 * k + 1]), abs=1e-5)
    row = opz.pack_BCs(BCs)
    assert row.shape == (8,) and row.dtype == np.float32 and np.allclose(row[:6], bc[0, :6], atol=1e-5)
    # diurnal form (data_containers.jl:131-156): top flux is a closure over time in seconds
    Q = 1e-7
    wT_top = lambda t: Q * np.sin(2 * np.pi / 86400.0 * t) / (p.alpha * p.g)
    BCd = opz.prepare_BCs(D, sc, wT_top)
    assert callable(BCd.wT.top)
    assert BCd.wT.top(21600.0) == pytest.approx(float(sc["wT"](Q / (p.alpha * p.g))), rel=1e-6)
    rowd = opz.pack_BCs(BCd, Q_diurnal=Q)
    assert rowd[5] == 0.0 and rowd[6] == np.float32(Q)


def test_c_shard_range_matches_host_and_nccl_binds(opz):
    """opz_shard_range (what a ccall host uses to pick its cases) == parallel.shard_range; NCCL is bound at run time."""
    lo, hi = C.c_int64(), C.c_int64()
    for B in (0, 1, 5, 18, 4097):
        for world in (1, 2, 3, 8):
            for r in range(world):
                opz.check(opz.lib.opz_shard_range(B, r, world, C.byref(lo), C.byref(hi)))
                assert (lo.value, hi.value) == opz.shard_range(B, r, world)
    assert opz.lib.opz_shard_range(4, 2, 2, C.byref(lo), C.byref(hi)) == opz.OPZ_EINVAL
    v = C.c_int()
    rc = opz.lib.opz_comm_nccl_version(C.byref(v))
    if rc == opz.OPZ_OK:
        assert v.value >= 21800, v.value          # NCCL >= 2.18
    else:
        assert rc == opz.OPZ_ENCCL and b"nccl" in opz.lib.opz_last_error().lower()


def test_nccl_library_preference(opz, monkeypatch):
    """libopz binds NCCL with dlopen; the Python package points it at the wheel PyTorch uses unless the user chose one
    (the loader keeps one library per SONAME: loading the system copy first would bind a later `import torch` to it)."""
    import importlib.util
    import sys
    L = sys.modules[opz.__name__ + "._lib"]
    monkeypatch.setenv("OPZ_NCCL_LIB", "/somewhere/libnccl.so.2")
    L._prefer_bundled_nccl()
    assert os.environ["OPZ_NCCL_LIB"] == "/somewhere/libnccl.so.2"     # an explicit choice wins
    monkeypatch.delenv("OPZ_NCCL_LIB")
    L._prefer_bundled_nccl()
    spec = importlib.util.find_spec("nvidia.nccl")
    if spec is not None and any(os.path.exists(os.path.join(p, "lib", "libnccl.so.2")) for p in spec.submodule_search_locations):
        assert os.environ["OPZ_NCCL_LIB"].endswith(os.path.join("nvidia", "nccl", "lib", "libnccl.so.2"))
    else:
        assert "OPZ_NCCL_LIB" not in os.environ
