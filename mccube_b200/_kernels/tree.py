"""Binary-tree partitioning (mirror of /root/reference/mccube/_kernels/tree.py:18-60).

The reference leaves the XLA program through `jax.pure_callback` and builds an sklearn
KDTree / BallTree on the HOST (tree.py:44-58); its output is sklearn's `idx_array` cut
into `leaf_size` chunks, i.e. it inherits `std::nth_element`'s platform-dependent order
(SURVEY.md Q7).  This mirror keeps exactly that behaviour -- a host callback into the same
sklearn call -- because no device algorithm can reproduce those indices; only the row
gather runs on the GPU.  It is the one place the hot path crosses to the host, as in the
reference.  `BoxPartitionKernel` is the device-resident partitioner.

`device=True` builds the same median-split tree ON THE DEVICE (csrc/kd_split.cu): sklearn's split
rule (dimension of largest spread, median at (end - start) // 2) level by level, so every node holds
the same point SET as sklearn's node, carried one level deeper so that a partition is a leaf (a KD
box) rather than half of an `nth_element`-ordered leaf.  The order inside a partition differs from
sklearn's (there it is platform dependent, Q7); no host round trip remains on the path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._tree import tree_map
from .base import AbstractPartitioningKernel


class BinaryTreePartitioningKernel(AbstractPartitioningKernel):
    def __init__(self, partition_count, mode: str = "BallTree", metric="minkowski", metric_kwargs=None,
                 device: bool = False, levels: int | None = None):
        super().__init__(partition_count)
        self.device = device
        self.levels = levels        # device=True: split levels (None = down to leaves of one partition)
        if mode not in ("KDTree", "BallTree"):
            raise ValueError(f"mode must be 'KDTree' or 'BallTree'; got {mode}")
        self.mode = mode
        self.metric = metric
        self.metric_kwargs = metric_kwargs or {}

    def indices(self, p: torch.Tensor, count: int, weighted: bool = False) -> torch.Tensor:
        if self.device:
            # both sklearn trees partition the points identically (_recursive_build is shared; the metric only
            # enters the node radii), so one device routine serves both modes
            if p.shape[0] % count != 0:
                raise ValueError(f"{p.shape[0]} particles do not divide into {count} partitions")
            leaf_size = p.shape[0] // count
            d = p.shape[1] - (1 if weighted else 0)
            levels = _lib.kd_partition_levels(p.shape[0], leaf_size) if self.levels is None else self.levels
            return _lib.kd_partition(p, d, levels).reshape(-1, leaf_size)
        from sklearn.neighbors import BallTree, KDTree

        trees = {"KDTree": KDTree, "BallTree": BallTree}
        leaf_size = p.shape[0] // count
        pu = p[:, :-1] if weighted else p
        host = pu.detach().cpu().numpy()
        t = trees[self.mode](host, leaf_size, self.metric, **self.metric_kwargs)
        idx = np.asarray(t.get_arrays()[1]).reshape(-1, leaf_size).astype(np.int32)
        return torch.from_numpy(idx).to(p.device)

    def __call__(self, t, particles, args, weighted: bool = False):
        def _one(p, count):
            p = p.contiguous()
            idx = self.indices(p, count, weighted)
            return _lib.gather_rows(p, idx.reshape(-1)).reshape(count, -1, p.shape[-1])

        return tree_map(_one, particles, self.partition_count)
