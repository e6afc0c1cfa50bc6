"""BoxPartitionKernel -- box binning as an integer key-sort.

NEW kernel: the brief (BASELINE.json north_star, config 3) names it, the reference does
not contain it (nearest analogues are `StratifiedPartitioningKernel`,
/root/reference/mccube/_kernels/stratified.py:36-45 -- scalar key, argsort, equal chunks --
and `BinaryTreePartitioningKernel`, tree.py -- axis-aligned boxes).  It satisfies the
`AbstractPartitioningKernel` contract (/root/reference/mccube/_kernels/base.py:54-75):
`[n, d] -> [m, n/m, d]`, equal sizes.  Definition (oracle/mccube_ref.py::box_keys):

    cell_a = clamp(floor((x_a - lo_a) * inv_h_a), 0, 2^bits - 1)   for each key dim a
    key    = lexicographic id of (cell_a)_a                         (n_dims * bits bits)
    order  = stable sort by key (ties keep row order);  partition b = order[b*n/m : (b+1)*n/m]

`bounds=None` uses the bounding box of the rows (lo = min, inv_h = 2^bits / (max - min)).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._tree import tree_map
from .base import AbstractPartitioningKernel


class BoxPartitionKernel(AbstractPartitioningKernel):
    def __init__(self, partition_count, dims=None, bits: int = 3, bounds=None):
        super().__init__(partition_count)
        self.dims = None if dims is None else tuple(int(a) for a in dims)
        self.bits = int(bits)
        self.bounds = bounds

    def resolve_dims(self, d: int):
        dims = self.dims if self.dims is not None else tuple(range(min(4, d)))
        if len(dims) > _lib.BOX_MAX_DIMS or len(dims) * self.bits > 32:
            raise ValueError(f"BoxPartitionKernel: {len(dims)} dims x {self.bits} bits unsupported")
        return dims

    def fixed_bounds(self, dims, device):
        """[lo..., inv_h...] on the device from user-given (lo, hi), or None."""
        if self.bounds is None:
            return None
        lo = np.asarray(self.bounds[0], np.float32)
        hi = np.asarray(self.bounds[1], np.float32)
        inv_h = (np.float32(1 << self.bits) / (hi - lo)).astype(np.float32)
        return torch.tensor(np.concatenate([lo, inv_h]), dtype=torch.float32, device=device)

    def indices(self, p: torch.Tensor, count: int, weighted: bool = False) -> torch.Tensor:
        d = p.shape[1] - (1 if weighted else 0)
        dims = self.resolve_dims(d)
        bounds = self.fixed_bounds(dims, p.device)
        if bounds is None:
            bounds = _lib.box_bounds(p, dims, self.bits)
        keys = _lib.box_keys(p, dims, self.bits, bounds)
        return _lib.sort_pairs(keys, None, len(dims) * self.bits).reshape(count, -1)

    def __call__(self, t, particles, args, weighted: bool = False):
        def _one(p, count):
            p = p.contiguous()
            idx = self.indices(p, count, weighted)
            return _lib.gather_rows(p, idx.reshape(-1)).reshape(count, -1, p.shape[-1])

        return tree_map(_one, particles, self.partition_count)
