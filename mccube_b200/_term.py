"""MCC term (mirror of /root/reference/mccube/_term.py:19-77)."""
from __future__ import annotations

import torch

from ._tree import tree_map
from .diffrax_compat import AbstractTerm, ODETerm, WeaklyDiagonalControlTerm


def _tree_flatten(tree):
    """_term.py:19-22: every leaf reshaped to [-1, d]."""
    return tree_map(lambda x: x.reshape(-1, x.shape[-1]), tree)


class MCCTerm(AbstractTerm):
    """Convenience term for CDEs driven by a cubature path (_term.py:25-77).

    `vf` evaluates the drift on the PARENTS only; `prod` is the `[n, 1, d] + [1, k, d]`
    broadcast add that defines the `[n, k, d]` expansion layout.  The fused CUDA path in
    `MCCSolver.step` reproduces this layout without building it."""

    def __init__(self, ode: ODETerm, cde: WeaklyDiagonalControlTerm):
        self.ode = ode
        self.cde = cde

    def vf(self, t, y, args):
        return self.ode.vf(t, _tree_flatten(y), args), self.cde.vf(t, _tree_flatten(y), args)

    def contr(self, t0, t1):
        return self.ode.contr(t0, t1), self.cde.contr(t0, t1)

    def prod(self, vf, control):
        ode_prod = self.ode.prod(vf[0], control[0])
        cde_prod = self.cde.prod(vf[1], control[1])

        def _bcast(o, c):
            c = torch.as_tensor(c, dtype=o.dtype, device=o.device)
            return o[:, None, ...] + c[None, ...]

        return tree_map(_bcast, ode_prod, cde_prod)
