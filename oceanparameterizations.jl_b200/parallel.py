"""Multi-GPU host logic: one process per GPU.

  * forward ensembles shard by column, contiguous ranges, no communication;
  * the training step has ONE exchange: the sum of the per-shard NN-parameter gradients and loss
    components (the loss is a mean over simulations, wind_mixing/src/NDE_training.jl:293-295, so
    every rank normalises by the GLOBAL batch `B_total` and the all-reduce is a plain sum).

The collective itself is `torch.distributed.all_reduce` (NCCL over NVLink on GPUs, gloo in the CPU
tests); nothing here touches the numerics.
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


def make_allreduce_comm(dist, device=None) -> Callable[[np.ndarray], None]:
    """Returns comm(buf): in-place sum of a float32 numpy buffer over all ranks (api.train_NDE's `comm`)."""
    import torch

    def comm(buf: np.ndarray) -> None:
        t = torch.from_numpy(buf)
        if device is not None:
            d = t.to(device)
            dist.all_reduce(d, op=dist.ReduceOp.SUM)
            t.copy_(d)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    return comm


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
