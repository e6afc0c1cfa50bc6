"""Multi-GPU host logic (one process per GPU, torch.distributed).

The reference has no multi-device code at all.  What shards naturally:
  * expansion + drift: embarrassingly parallel over parents -> contiguous row shards;
  * partitioned recombination: each rank partitions and recombines the children of its own
    shard ("device-local partitioning", a labelled variant of the reference's single global
    partition) -- no data-path collective; every rank draws the SAME local index vector, which
    is what quirk Q4 prescribes for every partition anyway;
  * moments / W2: per-rank fp64 partial sums + ONE all-reduce of d + 1 (+ d*d) doubles.
These helpers are device agnostic (they only touch torch.distributed and small tensors), so the
N > 1 logic is covered by world_size-2 gloo tests on CPU (tests/test_distributed_cpu.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous row range [lo, hi) of `rank` when n rows are split over `world` ranks
    (first n % world ranks get one extra row)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_mean(local_sums: torch.Tensor, group=None):
    """local_sums = [sum_r w_r x_r (d entries), sum_r w_r] of this rank's shard (fp64) ->
    (global mean[d], global weight total) after one all-reduce."""
    tot = all_reduce_sum_(local_sums.clone(), group)
    return tot[:-1] / tot[-1], tot[-1]


def global_cov(local_m2: torch.Tensor, local_sums: torch.Tensor, centre: torch.Tensor, mean: torch.Tensor,
               wtot: torch.Tensor, ddof: int = 1, group=None):
    """local_m2 = sum_r w_r (x_r - centre)(x_r - centre)^T over this rank's shard with a centre
    COMMON to all ranks (fp64).  One all-reduce, then the exact shift to the global mean:
    sum w (x-m)(x-m)^T = sum w (x-c)(x-c)^T - W (m-c)(m-c)^T."""
    del local_sums
    m2 = all_reduce_sum_(local_m2.clone(), group)
    delta = mean - centre.to(mean.dtype)
    m2 = m2 - wtot * torch.outer(delta, delta)
    return m2 / (wtot - ddof)


def distributed_mean_and_cov(particles: torch.Tensor, weights=None, ddof: int = 1, group=None):
    """jnp.mean / jnp.cov of the GLOBAL particle set from per-rank shards: two GPU reductions
    (mccb200_moments_sum, mccb200_moments_m2) and two tiny all-reduces."""
    from . import _lib

    d = particles.shape[1]
    mean_local, wtot_local = _lib.column_means(particles.contiguous(), d, weights)
    sums = torch.cat([mean_local * wtot_local, wtot_local.reshape(1)])
    mean, wtot = global_mean(sums, group)
    centre = mean.to(torch.float32).contiguous()
    m2 = _lib.second_moment(particles.contiguous(), d, centre, weights)
    cov = global_cov(m2, sums, centre, mean, wtot, ddof, group)
    return mean, cov
