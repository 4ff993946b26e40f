"""world_size-2 gloo tests of the multi-GPU host logic: column sharding of the forward
ensemble and the one exchange of the training step (sum all-reduce of gradient + loss components with
every rank normalising by the global batch).  The numerics on each rank come from the oracle here (no
GPU in this container); the GPU tests check the same additivity through libopz
(tests/test_gpu_bptt.py::test_gradient_is_additive_over_column_shards)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import opz_import
    from oracle import nde_adjoint as A
    from oracle import nde_oracle as O
    opz = opz_import.load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    B, n_steps, se = 5, 6, 3
    x0, bc = O.synthetic_columns(B, p)
    m = O.make_model((96, 24, 16, 31), O.ACT_MISH, seed=1, out_scale=0.1)
    dt = 30.0 / p.tau
    target = O.rollout(x0, bc, O.make_model((96, 24, 16, 31), O.ACT_MISH, seed=2, out_scale=0.1), p, 2, dt, n_steps, se)
    lw = [1, 1, 1, 5e-3, 5e-3, 5e-3]

    def fn(xs, bs, ts, B_total):
        if xs.shape[0] == 0:
            return 0.0, np.zeros(6), np.zeros(3 * m.P)
        return A.loss_and_grad(xs, bs, m, p, 2, dt, n_steps, se, ts, lw, B_total=B_total)

    total, comps, grad = opz.sharded_loss_grad(fn, dist, rank, world, x0, bc, target)
    # forward ensemble: every rank solves its own contiguous shard, no communication
    lo, hi = opz.shard_range(B, rank, world)
    sol = O.rollout(x0[lo:hi], bc[lo:hi], m, p, 2, dt, n_steps, se)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), total=total, comps=comps, grad=grad, sol=sol, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


def test_lib_shard_range(opz):
    for B in (0, 1, 5, 18, 4096):
        for world in (1, 2, 4, 8):
            r = [opz.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    assert [opz.shard_range(18, k, 8) for k in range(8)][:3] == [(0, 3), (3, 6), (6, 8)]
    with pytest.raises(ValueError):
        opz.shard_range(4, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gloo_training_exchange_and_ensemble_sharding(tmp_path):
    import torch.multiprocessing as mp
    from oracle import nde_adjoint as A
    from oracle import nde_oracle as O
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # both ranks hold the same reduced result ...
    assert r[0]["total"] == r[1]["total"] and np.array_equal(r[0]["grad"], r[1]["grad"])
    # ... equal to the single-process full-batch result
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    B, n_steps, se = 5, 6, 3
    x0, bc = O.synthetic_columns(B, p)
    m = O.make_model((96, 24, 16, 31), O.ACT_MISH, seed=1, out_scale=0.1)
    dt = 30.0 / p.tau
    target = O.rollout(x0, bc, O.make_model((96, 24, 16, 31), O.ACT_MISH, seed=2, out_scale=0.1), p, 2, dt, n_steps, se)
    t, c, g = A.loss_and_grad(x0, bc, m, p, 2, dt, n_steps, se, target, [1, 1, 1, 5e-3, 5e-3, 5e-3])
    assert float(r[0]["total"]) == pytest.approx(t, rel=1e-5)
    assert np.allclose(r[0]["comps"], c, rtol=1e-5, atol=1e-12)
    assert np.linalg.norm(r[0]["grad"] - g) <= 1e-5 * np.linalg.norm(g)
    # forward shards tile the ensemble exactly (bit-identical to the unsharded solve)
    full = O.rollout(x0, bc, m, p, 2, dt, n_steps, se)
    assert (int(r[0]["lo"]), int(r[0]["hi"]), int(r[1]["lo"]), int(r[1]["hi"])) == (0, 3, 3, 5)
    assert np.array_equal(np.concatenate([r[0]["sol"], r[1]["sol"]]), full)


# ---- train_NDE itself on two ranks (host logic: sharding by shard_range, B_total = n_sim, sum all-reduce, identical
#      ADAM steps on every rank); the libopz call is replaced by the oracle's hand-derived adjoint (no GPU here) ----
def _train_case(opz, O, A):
    from helpers import chains_from_oracle_model, les_shaped_dataset
    sizes = (96, 24, 16, 31)
    D, p, x0, bc, data, dt, per = les_shaped_dataset(opz, O, n_sim=3, nt=5, sizes=sizes, act=O.ACT_MISH)
    m0 = O.make_model(sizes, O.ACT_MISH, seed=22, out_scale=0.1)
    chains = chains_from_oracle_model(opz, m0)
    tsteps = np.arange(0, 5, 2)
    n_steps, se = per * 4, per * 2

    def fn(w, xs, bs, ts, lw, B_total):
        if xs.shape[0] == 0:
            return 0.0, np.zeros(6, dtype=np.float32), np.zeros(3 * m0.P, dtype=np.float32)
        m = O.Model.from_flat(np.asarray(w, dtype=np.float32), sizes, O.ACT_MISH)
        t, c, g = A.loss_and_grad(xs, bs, m, p, 2, dt, n_steps, se, ts, lw, B_total=B_total)
        return float(t), np.asarray(c, dtype=np.float32), np.asarray(g, dtype=np.float32)

    kw = dict(maxiters=3, modified_pacanowski_philander=True, zero_weights=True, train_gradient=True, gradient_scaling=5e-3,
              f=p.f, α=p.alpha, g=p.g, ν0=p.nu0, ν_m=p.nu_m, Riᶜ=p.Ric, ΔRi=p.dRi, Pr=p.Pr, _loss_grad_fn=fn)
    return chains, D, tsteps, opz.RK4(dt), kw


def _train_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import opz_import
    from oracle import nde_adjoint as A
    from oracle import nde_oracle as O
    opz = opz_import.load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chains, D, tsteps, stepper, kw = _train_case(opz, O, A)
    comm = opz.make_allreduce_comm(dist)
    assert (comm.rank, comm.world) == (rank, world)
    uw, vw, wT, hist = opz.train_NDE(*chains, D, tsteps, stepper, [opz.ADAM(2e-3)], 1, comm=comm, **kw)
    np.savez(os.path.join(out_dir, f"train{rank}.npz"), losses=np.array([h[0] for h in hist]),
             comps=np.stack([h[1] for h in hist]), w=opz.destructure(uw)[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_train_NDE_matches_single_process(tmp_path, opz):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import nde_adjoint as A
    from oracle import nde_oracle as O
    world, port = 2, _free_port()
    mp.spawn(_train_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"train{k}.npz") for k in range(world)]
    chains, D, tsteps, stepper, kw = _train_case(opz, O, A)
    uw, vw, wT, hist = opz.train_NDE(*chains, D, tsteps, stepper, [opz.ADAM(2e-3)], 1, **kw)
    losses = np.array([h[0] for h in hist])
    # both ranks saw the same global losses, equal to the single-process history (NOT world x larger), and took the same steps
    assert np.array_equal(r[0]["losses"], r[1]["losses"]) and np.array_equal(r[0]["w"], r[1]["w"])
    assert np.allclose(r[0]["losses"], losses, rtol=1e-5)
    assert np.allclose(r[0]["comps"], np.stack([h[1] for h in hist]), rtol=1e-4, atol=1e-9)
    assert np.allclose(r[0]["w"], opz.destructure(uw)[0], rtol=0, atol=2e-6)
    assert losses[-1] < losses[0]
