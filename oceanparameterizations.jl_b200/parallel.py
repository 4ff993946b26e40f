"""Multi-GPU host logic: one process per GPU.

  * forward ensembles shard by column, contiguous ranges, no communication;
  * the training step has ONE exchange: the sum of the per-shard NN-parameter gradients and loss
    components (the loss is a mean over simulations, wind_mixing/src/NDE_training.jl:293-295, so
    every rank normalises by the GLOBAL batch `B_total` and the all-reduce is a plain sum).

On GPUs the collective lives INSIDE libopz (include/opz.h: opz_ctx_comm_init; api.Context.comm_init /
init_library_comm below): opz_loss_grad* return the already-reduced global loss and gradient, which is what a
`ccall` user gets too.  The host-side route in this file (`make_allreduce_comm`: torch.distributed.all_reduce,
gloo in the CPU tests) carries the same shard + global-B_total contract for hosts that own their collective.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of the columns owned by `rank`; the first B % world ranks get one extra.
    Ranks beyond B own an empty range (they still take part in the all-reduce with zeros)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(B), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class HostComm:
    """comm(buf): in-place sum of a float32 numpy buffer over all ranks; .rank / .world say which shard is ours
    (api.train_NDE's `comm`)."""

    def __init__(self, dist, device=None):
        self.dist, self.device = dist, device
        self.rank, self.world = int(dist.get_rank()), int(dist.get_world_size())

    def __call__(self, buf: np.ndarray) -> None:
        import torch
        t = torch.from_numpy(buf)
        if self.device is not None:
            d = t.to(self.device)
            self.dist.all_reduce(d, op=self.dist.ReduceOp.SUM)
            t.copy_(d)
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)


def make_allreduce_comm(dist, device=None) -> HostComm:
    return HostComm(dist, device)


def init_library_comm(ctx, dist) -> None:
    """Gives `ctx` its rank of a NCCL communicator owned by libopz (opz_ctx_comm_init), using an initialised
    torch.distributed process group only to ship rank 0's 128-byte id.  Afterwards api.loss_grad* / train_NDE on this
    context return globally reduced results with no host-side collective."""
    from . import api
    rank, world = int(dist.get_rank()), int(dist.get_world_size())
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], world, rank)


def sharded_loss_grad(loss_grad_fn: Callable, dist, rank: int, world: int, x0, bc, target, *, device=None):
    """Runs `loss_grad_fn(x0_shard, bc_shard, target_shard, B_total)` -> (total, comps[6], grad) on this
    rank's shard and all-reduces (sum).  `loss_grad_fn` is the libopz call (api.loss_grad) on the GPU
    path; the CPU tests pass the oracle in its place to check exactly this sharding/normalisation logic."""
    B = x0.shape[0]
    lo, hi = shard_range(B, rank, world)
    total, comps, grad = loss_grad_fn(x0[lo:hi], bc[lo:hi], target[lo:hi], B)
    buf = np.concatenate([[total], np.asarray(comps, dtype=np.float32), np.asarray(grad, dtype=np.float32)]).astype(np.float32)
    if world > 1:
        make_allreduce_comm(dist, device)(buf)
    return float(buf[0]), buf[1:7].copy(), buf[7:].copy()
