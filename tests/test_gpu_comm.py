"""The training step's collective INSIDE the boundary (include/opz.h: opz_ctx_comm_init / opz_loss_grad*):
every rank passes its shard of the simulations with B_total = the global count and receives the global loss and
gradient — what replaces the mean over simulations of loss_NDE (wind_mixing/src/NDE_training.jl:290-301) when the
simulations are spread over GPUs.  One process per GPU; NCCL is bound by libopz at run time.
  * world = 1 on one GPU: the whole plumbing (dlopen, id, ncclCommInitRank, group all-reduce on the context's stream);
  * world = 2 on two GPUs (skipped with fewer): sharded == unsharded, for opz_loss_grad, opz_loss_grad_mpp and train_NDE.
"""
import os
import socket
import sys

import numpy as np
import pytest

from helpers import product_model, product_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _case(O, B=7, sizes=(96, 50, 20, 31), act=3, n_steps=12, se=4):
    p = O.Params(flags=O.FLAG_MPP | O.FLAG_RI_EPS | O.FLAG_BC_MINUS_SCALE0)
    x0, bc = O.synthetic_columns(B, p)
    x0 = (x0 + 0.02 * np.random.default_rng(2).normal(size=x0.shape)).astype(np.float32)
    m = O.make_model(sizes, act, seed=1, out_scale=0.1)
    dt = 30.0 / p.tau
    target = O.rollout(x0, bc, O.make_model(sizes, act, seed=101, out_scale=0.1), p, 2, dt, n_steps, se)
    lw = np.array([1.0, 0.7, 1.3, 5e-3, 4e-3, 6e-3], dtype=np.float32)
    return p, x0, bc, m, dt, target, lw, n_steps, se


def test_single_rank_communicator_is_identity(opz, oracle):
    ctx = opz.Context(0)
    p, x0, bc, m, dt, target, lw, n_steps, se = _case(oracle)
    model = product_model(opz, m, ctx=ctx)
    pp = product_params(opz, p)
    ref = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, se)
    ctx.comm_init(opz.comm_unique_id(), 1, 0)
    assert ctx.comm_info[:2] == (1, 0)
    got = opz.loss_grad(model, pp, x0, bc, target, lw, 2, dt, n_steps, se)
    assert got[0] == ref[0] and np.array_equal(got[2], ref[2])
    ctx.comm_destroy()
    model.close()
    ctx.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import opz_import
    from helpers import chains_from_oracle_model, les_shaped_dataset
    from oracle import nde_oracle as O
    opz = opz_import.load()
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)    # host plumbing only: ships the NCCL id
    ctx = opz.Context(rank)
    opz.init_library_comm(ctx, dist)
    assert ctx.comm_info[:2] == (world, rank)
    p, x0, bc, m, dt, target, lw, n_steps, se = _case(O)
    B = x0.shape[0]
    lo, hi = opz.shard_range(B, rank, world)
    model = product_model(opz, m, ctx=ctx)
    pp = product_params(opz, p)
    total, comps, grad = opz.loss_grad(model, pp, x0[lo:hi], bc[lo:hi], target[lo:hi], lw, 2, dt, n_steps, se, B_total=B)
    t2, c2, g2, gm = opz.loss_grad_mpp(model, pp, x0[lo:hi], bc[lo:hi], target[lo:hi], lw, 2, dt, n_steps, se, B_total=B)
    # a rank that owns nothing still takes part: shard everything onto rank 0
    lo0, hi0 = (0, B) if rank == 0 else (B, B)
    t3, c3, g3 = opz.loss_grad(model, pp, x0[lo0:hi0], bc[lo0:hi0], target[lo0:hi0], lw, 2, dt, n_steps, se, B_total=B)
    n_coll = ctx.comm_info[2]
    model.close()
    # train_NDE on a context with a communicator: shards internally, reduced inside libopz
    sizes = (96, 24, 16, 31)
    D, pt, _, _, _, dtt, per = les_shaped_dataset(opz, O, n_sim=3, nt=5, sizes=sizes, act=O.ACT_MISH)
    chains = chains_from_oracle_model(opz, O.make_model(sizes, O.ACT_MISH, seed=22, out_scale=0.1))
    uw, vw, wT, hist = opz.train_NDE(*chains, D, np.arange(0, 5, 2), opz.RK4(dtt), [opz.ADAM(2e-3)], 1, maxiters=3,
                                     modified_pacanowski_philander=True, zero_weights=True, train_gradient=True,
                                     f=pt.f, α=pt.alpha, g=pt.g, ν0=pt.nu0, ν_m=pt.nu_m, Riᶜ=pt.Ric, ΔRi=pt.dRi, Pr=pt.Pr, ctx=ctx)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), total=total, comps=comps, grad=grad, gm=gm, g2=g2, t3=t3, g3=g3,
             n_coll=n_coll, losses=np.array([h[0] for h in hist]), w=opz.destructure(uw)[0])
    dist.barrier()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_loss_grad_is_reduced_inside_the_library(opz, oracle, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from helpers import chains_from_oracle_model, les_shaped_dataset
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    O = oracle
    p, x0, bc, m, dt, target, lw, n_steps, se = _case(O)
    ctx = opz.Context(0)
    model = product_model(opz, m, ctx=ctx)
    pp = product_params(opz, p)
    t, c, g, gm = opz.loss_grad_mpp(model, pp, x0, bc, target, lw, 2, dt, n_steps, se)
    model.close()
    for k in range(world):
        assert float(r[k]["total"]) == pytest.approx(t, rel=1e-5)
        assert np.allclose(r[k]["comps"], c, rtol=1e-4, atol=1e-9)
        assert _rel_l2(r[k]["grad"], g) < 1e-5 and _rel_l2(r[k]["g2"], g) < 1e-5 and _rel_l2(r[k]["g3"], g) < 1e-5
        assert _rel_l2(r[k]["gm"], gm) < 1e-4
        assert float(r[k]["t3"]) == pytest.approx(t, rel=1e-5)
        assert int(r[k]["n_coll"]) == 3
    assert np.array_equal(r[0]["grad"], r[1]["grad"]) and np.array_equal(r[0]["w"], r[1]["w"])
    # train_NDE: the two-rank history is the single-GPU history
    sizes = (96, 24, 16, 31)
    D, pt, _, _, _, dtt, per = les_shaped_dataset(opz, O, n_sim=3, nt=5, sizes=sizes, act=O.ACT_MISH)
    chains = chains_from_oracle_model(opz, O.make_model(sizes, O.ACT_MISH, seed=22, out_scale=0.1))
    uw, vw, wT, hist = opz.train_NDE(*chains, D, np.arange(0, 5, 2), opz.RK4(dtt), [opz.ADAM(2e-3)], 1, maxiters=3,
                                     modified_pacanowski_philander=True, zero_weights=True, train_gradient=True,
                                     f=pt.f, α=pt.alpha, g=pt.g, ν0=pt.nu0, ν_m=pt.nu_m, Riᶜ=pt.Ric, ΔRi=pt.dRi, Pr=pt.Pr, ctx=ctx)
    assert np.allclose(r[0]["losses"], [h[0] for h in hist], rtol=1e-4)
    assert np.allclose(r[0]["w"], opz.destructure(uw)[0], atol=5e-6)
    ctx.close()
