"""Minimal PyTree helpers (the reference uses jax.tree_util; particles may be a dict /
list / tuple of arrays, e.g. /root/reference/tests/test_kernels.py:44-53,145-154)."""
from __future__ import annotations

import torch


def is_leaf(x) -> bool:
    return not isinstance(x, (dict, list, tuple))


def tree_leaves(tree):
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out += tree_leaves(tree[k])
        return out
    if isinstance(tree, (list, tuple)):
        out = []
        for v in tree:
            out += tree_leaves(v)
        return out
    return [tree]


def tree_map(fn, tree, *rest):
    """Maps over the leaves of `tree`; `rest` trees may be prefixes (a bare leaf broadcasts)."""
    if isinstance(tree, dict):
        return {k: tree_map(fn, tree[k], *[r[k] if isinstance(r, dict) else r for r in rest]) for k in sorted(tree)}
    if isinstance(tree, (list, tuple)):
        out = [tree_map(fn, v, *[r[i] if isinstance(r, (list, tuple)) else r for r in rest])
               for i, v in enumerate(tree)]
        return type(tree)(out) if not hasattr(tree, "_fields") else type(tree)(*out)
    return fn(tree, *rest)


def tree_map_with_index(fn, tree, *rest):
    """Like tree_map but `fn(leaf_index, n_leaves, leaf, *rest_leaves)`; leaf order is
    jax.tree_util's (dict keys sorted)."""
    n = len(tree_leaves(tree))
    counter = [0]

    def wrapped(leaf, *r):
        i = counter[0]
        counter[0] += 1
        return fn(i, n, leaf, *r)

    return tree_map(wrapped, tree, *rest)


def tree_equal(a, b, rtol=0.0, atol=0.0) -> bool:
    la, lb = tree_leaves(a), tree_leaves(b)
    if len(la) != len(lb):
        return False
    for x, y in zip(la, lb):
        if x is None or y is None:
            if x is not y:
                return False
            continue
        x = torch.as_tensor(x)
        y = torch.as_tensor(y)
        if x.shape != y.shape:
            return False
        if not torch.allclose(x.double().cpu(), y.double().cpu(), rtol=rtol, atol=atol):
            return False
    return True
