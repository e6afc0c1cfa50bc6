"""Monte Carlo recombination / partitioning
(mirror of /root/reference/mccube/_kernels/random.py:22-115).

The index vector `jr.choice` would gather with is produced on the GPU by
`mccb200_mc_indices` (in-kernel threefry2x32 + stable radix-sort shuffle), bit-exact
with jax.random on the integer branches (p = None).  Weighted sampling with a real
`weighting_function` (Gumbel top-k / inverse-CDF, random.py:62-65) uses the same random
bits but float arithmetic (log, cumsum) that cannot be promised bit-exact across XLA-CPU and
CUDA libm: `mccb200_gumbel_keys` / `mccb200_weighted_choice_replace`.
"""
from __future__ import annotations

import math

import torch

from .. import _lib
from .._tree import tree_map_with_index
from .._utils import nop, unpack_particles
from .base import AbstractPartitioningKernel, AbstractRecombinationKernel


def key_words(key):
    """Accepts `(k0, k1)`, an int seed (`jr.key(seed)`), or a 2-element array/tensor."""
    if isinstance(key, int):
        return ((key >> 32) & 0xFFFFFFFF, key & 0xFFFFFFFF)
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().tolist()
    k0, k1 = key
    return (int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF)


class MonteCarloKernel(AbstractRecombinationKernel):
    """`MonteCarloKernel(recombination_count, with_replacement=False, weighting_function=nop, *, key)`
    (random.py:22-68).  `threefry_partitionable` mirrors the `jax_threefry_partitionable`
    flag of the JAX the reference would run under (False = the reference's era)."""

    def __init__(self, recombination_count, with_replacement: bool = False, weighting_function=nop, *, key,
                 threefry_partitionable: bool = False, x64: bool = False):
        super().__init__(recombination_count)
        # mirrors jax_enable_x64 for the index draw of the sharded pull path (int64 randint: 64-bit draws,
        # more than 2^32 children); the single-device paths use 32-bit child ids
        self.x64 = x64
        self.with_replacement = with_replacement
        self.weighting_function = weighting_function
        self.key = key_words(key)
        self.threefry_partitionable = threefry_partitionable

    # -- index vector (what jr.choice gathers with; the reference never exposes it)
    def indices(self, t, n_inputs: int, count, device, leaf: int = 0, n_leaves: int = 1) -> torch.Tensor:
        shape = tuple(count) if isinstance(count, (tuple, list)) else (count,)
        rng = _lib.make_rng(self.key, t, leaf=leaf, n_leaves=n_leaves, partitionable=self.threefry_partitionable)
        if self.x64 and self.with_replacement:
            # jax_enable_x64: randint draws int64 (different bits from the int32 recipe even for small spans);
            # the single-device kernels take 32-bit child ids
            if n_inputs > 2 ** 32:
                raise ValueError(f"{n_inputs} inputs exceed the 32-bit child ids of the single-device path (use step_pull)")
            idx = _lib.mc_indices_slice(rng, n_inputs, math.prod(shape), 0, math.prod(shape), device, x64=True)
            return idx.to(torch.int32)          # wraps to the uint32 bit pattern the kernels read
        return _lib.mc_indices(rng, n_inputs, math.prod(shape), self.with_replacement, device)

    def indices_slice(self, t, n_inputs: int, count: int, first: int, size: int, device) -> torch.Tensor:
        """Entries [first, first + size) of `indices(t, n_inputs, count)` for with_replacement=True (counter
        based draws: any slice can be produced on its own)."""
        if not self.with_replacement:
            raise ValueError("indices_slice needs with_replacement=True (the shuffle is not sliceable)")
        rng = _lib.make_rng(self.key, t, partitionable=self.threefry_partitionable)
        x64 = getattr(self, "x64", False)
        if n_inputs > 2 ** 32 - 1 and not x64:
            raise ValueError(f"{n_inputs} inputs need 64-bit indices: set kernel.x64 = True (jax_enable_x64 semantics)")
        return _lib.mc_indices_slice(rng, n_inputs, count, first, size, device, x64=x64)

    def _probabilities(self, p, weighted):
        """`weights = self.weighting_function(weights)` (random.py:60-63); None = uniform."""
        if not weighted:
            return None
        _, w = unpack_particles(p, True)
        w = self.weighting_function(w)
        if w is None:
            return None
        return torch.as_tensor(w, dtype=torch.float32, device=p.device)

    def weighted_indices(self, t, probs, count, leaf: int = 0, n_leaves: int = 1) -> torch.Tensor:
        """Gumbel top-k (without replacement) / inverse-CDF (with) indices (random.py:65 with p)."""
        shape = tuple(count) if isinstance(count, (tuple, list)) else (count,)
        rng = _lib.make_rng(self.key, t, leaf=leaf, n_leaves=n_leaves, partitionable=self.threefry_partitionable)
        return _lib.weighted_indices(rng, probs, math.prod(shape), self.with_replacement)

    def __call__(self, t, particles, args, weighted: bool = False):
        def _choice(i, n_leaves, p, count):
            probs = self._probabilities(p, weighted)
            shape = tuple(count) if isinstance(count, (tuple, list)) else (count,)
            if probs is None:
                idx = self.indices(t, p.shape[0], shape, p.device, i, n_leaves)
            else:
                idx = self.weighted_indices(t, probs, shape, i, n_leaves)
            out = _lib.gather_rows(p.contiguous(), idx)
            return out.reshape(*shape, p.shape[-1])

        return tree_map_with_index(_choice, particles, self.recombination_count)

    # -- all partitions at once with the shared local index vector (quirk Q4)
    def call_batched(self, t, part, args, weighted, count):
        m, per, ld = part.shape
        if self._probabilities(part[0], weighted) is not None:
            # probabilities differ per partition (same key, same Gumbel noise): one draw per partition
            flat = part.reshape(m * per, ld).contiguous()
            rows = []
            for b in range(m):
                probs = self._probabilities(part[b], weighted)
                rows.append(self.weighted_indices(t, probs, count).to(torch.int64) + b * per)
            idx = torch.cat(rows).to(torch.int32)
            return _lib.gather_rows(flat, idx).reshape(m, -1, ld)
        idx_local = self.indices(t, per, count, part.device)
        idx = (torch.arange(m, device=part.device, dtype=torch.int64)[:, None] * per
               + idx_local.to(torch.int64)[None, :]).reshape(-1).to(torch.int32)
        return _lib.gather_rows(part.reshape(m * per, ld).contiguous(), idx).reshape(m, -1, ld)


class MonteCarloPartitioningKernel(AbstractPartitioningKernel):
    """Random assignment to `m` equal partitions (random.py:71-115): the wrapped kernel is
    called with `recombination_count = (m, n // m)`."""

    def __init__(self, partition_count, monte_carlo_kernel: MonteCarloKernel):
        super().__init__(partition_count)
        self.monte_carlo_kernel = monte_carlo_kernel

    def __call__(self, t, particles, args, weighted: bool = False):
        from .._tree import tree_map
        from .base import _call_with_count

        def _one(p, c):
            return _call_with_count(self.monte_carlo_kernel, (c, p.shape[0] // c), t, p, args, weighted)

        return tree_map(_one, particles, self.partition_count)
