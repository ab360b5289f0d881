"""Data-parallel glue: samples are independent, theta is replicated, and the only
exchange per update step is the TD-gradient accumulator ``[sum td*dQ/dtheta (n), sum td, n_valid]``
(examples/linear_system_mpc_qlearning.py:193,203).  One process per GPU; ``torch.distributed`` with the
NCCL backend on GPUs (gloo in the CPU tests of this host-side logic).
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples for this rank (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_accumulator(acc: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank accumulators (in place).  No-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def td_parameter_step(theta: torch.Tensor, acc: torch.Tensor, lr: float, n_total: int | None = None) -> torch.Tensor:
    """theta + mean_i(lr * td_i * dQ_i/dtheta) from the (all-reduced) accumulator.  The reference divides
    by the number of samples of the episode (np.mean over the stacked rows, example line 203); pass
    ``n_total`` for that, default: the number of valid samples counted in the accumulator."""
    n = acc.shape[0] - 2
    denom = float(n_total) if n_total is not None else float(acc[n + 1].item())
    if denom <= 0:
        return theta
    out = theta.clone()
    out[:n] += lr * acc[:n].to(theta.dtype) / denom
    return out
