"""Norm-stratified partitioning (mirror of /root/reference/mccube/_kernels/stratified.py:13-47)."""
from __future__ import annotations

import torch

from .. import _lib
from .._tree import tree_map
from .base import AbstractPartitioningKernel

L2_NORM = "l2"


class StratifiedPartitioningKernel(AbstractPartitioningKernel):
    """Centre on the (weighted) centre of mass, sort by norm, cut into `m` equal strata
    (stratified.py:36-45).  `norm=None` is a plain reshape (stratified.py:37-38).  Only the
    reference's default norm (Euclidean, `jnp.linalg.norm(axis=-1)`) has a CUDA key kernel;
    any other callable is applied to the centred rows as torch code and its output sorted."""

    def __init__(self, partition_count, norm=L2_NORM):
        super().__init__(partition_count)
        self.norm = norm

    def indices(self, p: torch.Tensor, count: int, weighted: bool = False) -> torch.Tensor:
        d = p.shape[1] - (1 if weighted else 0)
        w = p[:, -1] if weighted else None
        mean64, _ = _lib.column_means(p, d, w)
        com = mean64.to(torch.float32).contiguous()
        if self.norm == L2_NORM:
            keys = _lib.stratified_keys(p, d, com)
        else:
            vals = self.norm(p[:, :d] - com).to(torch.float32).contiguous()
            keys = _float_sort_keys(vals)
        return _lib.sort_pairs(keys, None, 32).reshape(count, -1)

    def __call__(self, t, particles, args, weighted: bool = False):
        def _one(p, count):
            if self.norm is None:
                return p.reshape(count, -1, p.shape[-1])
            p = p.contiguous()
            idx = self.indices(p, count, weighted)
            return _lib.gather_rows(p, idx.reshape(-1)).reshape(count, -1, p.shape[-1])

        return tree_map(_one, particles, self.partition_count)


def _float_sort_keys(v: torch.Tensor) -> torch.Tensor:
    """Order-preserving float32 -> uint32 map (sign flip trick), as int32 bit patterns."""
    bits = v.view(torch.int32)
    mask = (bits >> 31) | torch.tensor(-0x80000000, dtype=torch.int32, device=v.device)
    return bits ^ mask
