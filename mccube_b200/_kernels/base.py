"""Kernel protocols and PartitioningRecombinationKernel
(mirror of /root/reference/mccube/_kernels/base.py:18-177)."""
from __future__ import annotations

import abc

import torch

from .._tree import tree_map


class AbstractKernel(abc.ABC):
    """`kernel(t, particles, args, weighted=False) -> particles` (base.py:18-51)."""

    @abc.abstractmethod
    def __call__(self, t, particles, args, weighted: bool = False):
        ...


class AbstractPartitioningKernel(AbstractKernel):
    """Reshapes every leaf `[n, d] -> [m, n/m, d]`, equal-size partitions (base.py:54-75)."""

    def __init__(self, partition_count=None):
        self.partition_count = partition_count


class AbstractRecombinationKernel(AbstractKernel):
    """Strictly reduces the leading dimension of every leaf (base.py:78-100)."""

    def __init__(self, recombination_count=None):
        self.recombination_count = recombination_count


class PartitioningRecombinationKernel(AbstractRecombinationKernel):
    """Partition, then apply one recombination kernel to every partition (base.py:103-177).

    The reference vmaps the recombination kernel with `in_axes=(None, 0, None, None)`; the
    kernel module -- and its PRNG key -- is closed over, not batched, so every partition is
    resampled with the SAME local index vector (quirk Q4).  `MonteCarloKernel` exposes that
    vector (`local_indices`), which lets this class run all partitions with one gather."""

    def __init__(self, partitioning_kernel: AbstractPartitioningKernel,
                 recombination_kernel: AbstractRecombinationKernel):
        self.partitioning_kernel = partitioning_kernel
        self.recombination_kernel = recombination_kernel
        self.recombination_count = tree_map(
            lambda r, p: r * p, recombination_kernel.recombination_count, partitioning_kernel.partition_count)

    def __call__(self, t, particles, args, weighted: bool = False):
        partitioned = self.partitioning_kernel(t, particles, args, weighted)

        def _recombine(part, count):
            batched = getattr(self.recombination_kernel, "call_batched", None)
            if batched is not None:  # one gather for all partitions (shared indices, Q4)
                return batched(t, part, args, weighted, count)
            return torch.stack([_call_with_count(self.recombination_kernel, count, t, q, args, weighted)
                                for q in part])

        recombined = tree_map(_recombine, partitioned, self.recombination_kernel.recombination_count)
        return tree_map(lambda r, p: r.reshape(-1, p.shape[-1]), recombined, particles)


def _call_with_count(kernel, count, t, q, args, weighted):
    """eqx.tree_at(lambda k: k.recombination_count, kernel, c) of base.py:162-167."""
    old = kernel.recombination_count
    kernel.recombination_count = count
    try:
        return kernel(t, q, args, weighted)
    finally:
        kernel.recombination_count = old
