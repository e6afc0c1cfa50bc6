"""Helpers (mirror of /root/reference/mccube/_utils.py:8-107) on torch device tensors."""
from __future__ import annotations

import torch

from ._tree import tree_map


def nop(*args, **kwargs) -> None:
    """Accepts anything, does nothing (_utils.py:8-17)."""
    ...


def pack_particles(particles, weights):
    """`jnp.c_[p, squeeze(w) / sum(w)]`, or `p` unchanged when `w is None` (_utils.py:47-51)."""

    def _pack(p, w):
        if w is None:
            return p
        w = torch.as_tensor(w, dtype=p.dtype, device=p.device)
        w = w.reshape(-1) / w.sum()
        return torch.cat([p, w[:, None]], dim=1)

    if weights is None:
        return particles
    return tree_map(_pack, particles, weights)


def unpack_particles(particles, weighted: bool = False):
    """`(p[..., :-1], p[..., -1])` if weighted else `(p, None)` (_utils.py:82-84)."""
    _p = tree_map(lambda p: p[..., :-1] if weighted else p, particles)
    _w = tree_map(lambda p: p[..., -1] if weighted else None, particles)
    return _p, _w


def all_subclasses(cls: type) -> set:
    """_utils.py:87-101."""
    out = set()
    for sub in cls.__subclasses__():
        out.add(sub)
        out.update(all_subclasses(sub))
    return out


def requires_weighing(is_weighted: bool) -> None:
    """_utils.py:104-107."""
    if not is_weighted:
        raise ValueError("Kernel requires `weighted=True`; got {`weighted=False`}.")
