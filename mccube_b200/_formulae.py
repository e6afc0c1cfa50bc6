"""Cubature formulae (mirror of /root/reference/mccube/_formulae.py).

`Hadamard` is the formula on the north-star path: the CUDA kernels generate its signs in
registers.  The Stroud-Secrest formulae (_formulae.py:230-487) and the registry search
(_formulae.py:560-625) are host-side NumPy constants; their `[k, d]` point tables reach the
kernels through `mccb200_cubature.points` / `.lam` (the generic table path of `expand_kernel`).
"""
from __future__ import annotations

import abc
from functools import cached_property
from itertools import combinations, permutations

import numpy as np

from ._regions import AbstractRegion, GaussianRegion
from ._utils import all_subclasses


class AbstractCubature(abc.ABC):
    """_formulae.py:33-120: `points`, `weights`, `point_count`, stacked views, `__call__`."""

    degree: int
    sparse: bool = False

    def __init__(self, region: AbstractRegion):
        self.region = region

    @property
    @abc.abstractmethod
    def weights(self):
        ...

    @property
    @abc.abstractmethod
    def points(self):
        ...

    @property
    @abc.abstractmethod
    def point_count(self) -> int:
        ...

    @cached_property
    def stacked_weights(self) -> np.ndarray:
        """_formulae.py:94-98."""
        pts = self.points if isinstance(self.points, (list, tuple)) else [self.points]
        wts = self.weights if isinstance(self.weights, (list, tuple)) else [self.weights] * len(pts)
        return np.hstack([w * np.ones(p.shape[0]) for w, p in zip(wts, pts)])

    @cached_property
    def stacked_points(self) -> np.ndarray:
        """_formulae.py:100-103."""
        pts = self.points if isinstance(self.points, (list, tuple)) else [self.points]
        return np.vstack(pts)

    def __call__(self, integrand):
        """Q[f] = sum_j lambda_j f(x_j) (_formulae.py:105-142)."""
        vals = np.array([w * integrand(x) for w, x in zip(self.stacked_weights, self.stacked_points)])
        return vals.sum(), vals

    def __repr__(self):
        return f"{type(self).__name__}(region={self.region!r})"


class AbstractGaussianCubature(AbstractCubature):
    """Formulae valid for GaussianRegion (_formulae.py:145-188)."""

    def __init__(self, region: GaussianRegion):
        super().__init__(region)


class Hadamard(AbstractGaussianCubature):
    """Degree-3 Gaussian cubature of Victoir (2004) (_formulae.py:191-227).

    points = vstack(H[:, :d], -H[:, :d]) with H the Sylvester Hadamard matrix of order
    k/2, generated here as H[j, c] = (-1)^popcount(j & c) -- the same closed form the CUDA
    kernels evaluate in registers, so the [k, d] table is never shipped to the device.
    """

    degree = 3

    @cached_property
    def point_count(self) -> int:
        max_power = np.ceil(np.log2(self.region.dimension)) + 1
        return int(2**max_power)

    @cached_property
    def weights(self) -> float:
        return self.region.volume / self.point_count

    @cached_property
    def points(self) -> np.ndarray:
        half = self.point_count // 2
        j = np.arange(half, dtype=np.int64)[:, None]
        c = np.arange(self.region.dimension, dtype=np.int64)[None, :]
        v = j & c
        par = np.zeros_like(v)
        while np.any(v):
            par ^= v & 1
            v >>= 1
        h = 1 - 2 * par
        return np.vstack((h, -h))


# ------------------------------------------------------------------ Stroud-Secrest (1963) formulae
def _generate_point_permutations(point, mode: str = "FS") -> np.ndarray:
    """Orbit of a point under the reflection group ("R": every sign pattern of its non-zero
    coordinates, first coordinate slowest, -1 before +1), the symmetric group ("S": its
    non-zero coordinates placed in every ordered choice of positions) or both ("FS"), as in
    _formulae.py:490-558.  "S" and "FS" return the distinct points in lexicographic row order."""
    if mode not in {"R", "S", "FS"}:
        raise ValueError(f"Mode must be one of {{'R', 'S', 'FS'}}, got {mode}.")
    pt = np.atleast_1d(np.asarray(point, dtype=float))
    if pt.ndim > 1:
        pt = pt.reshape(-1, pt.shape[-1])[0]
    d = pt.shape[-1]
    nz = pt[pt != 0]
    m = nz.shape[0]
    if mode in ("R", "FS"):
        rows = np.arange(1 << m)[:, None]
        bits = (rows >> (m - 1 - np.arange(m))[None, :]) & 1          # bit set -> +1
        reflected = (2.0 * bits - 1.0) * nz[None, :]
        if mode == "R":
            return reflected
    else:
        reflected = nz[None, :]
    out = []
    for pos in permutations(range(d), m):
        block = np.zeros((reflected.shape[0], d))
        block[:, list(pos)] = reflected
        out.append(block)
    return np.unique(np.vstack(out), axis=0)


class StroudSecrest63_31(AbstractGaussianCubature):
    """Degree 3, k = 2d: +-sqrt(d) e_i with equal weights (E_n^{r^2} 3-1; _formulae.py:230-280)."""

    degree = 3
    sparse = True

    @cached_property
    def point_count(self) -> int:
        return int(2 * self.region.dimension)

    @cached_property
    def weights(self) -> float:
        return self.region.volume / self.point_count

    @cached_property
    def points(self) -> np.ndarray:
        d = self.region.dimension
        axes = np.sqrt(d) * np.eye(d)
        return np.vstack([axes, -axes])


class StroudSecrest63_32(AbstractGaussianCubature):
    """Degree 3, k = 2^d: (+-1, ..., +-1) with equal weights (E_n^{r^2} 3-2; _formulae.py:283-337)."""

    degree = 3

    @cached_property
    def point_count(self) -> int:
        return int(2 ** self.region.dimension)

    @cached_property
    def weights(self) -> float:
        return self.region.volume / 2 ** self.region.dimension

    @cached_property
    def points(self) -> np.ndarray:
        return _generate_point_permutations(np.ones(self.region.dimension), "R")


class StroudSecrest63_52(AbstractGaussianCubature):
    """Degree 5, k = 2d^2 + 1: the origin, (+-sqrt(d+2), 0, ...)_FS and
    (+-sqrt((d+2)/2), +-sqrt((d+2)/2), 0, ...)_FS (E_n^{r^2} 5-2; _formulae.py:340-406)."""

    degree = 5
    sparse = True

    @cached_property
    def point_count(self) -> int:
        return int(2 * self.region.dimension ** 2 + 1)

    @cached_property
    def weights(self):
        d, vol = self.region.dimension, self.region.volume
        return (2 / (d + 2) * vol, (4 - d) / (2 * (d + 2) ** 2) * vol, 1 / (d + 2) ** 2 * vol)

    @cached_property
    def points(self):
        d = self.region.dimension
        origin = np.zeros((1, d))
        one = np.zeros(d)
        one[0] = np.sqrt(d + 2)
        two = np.zeros(d)
        two[:2] = np.sqrt((d + 2) / 2)
        return (origin, _generate_point_permutations(one, "FS"), _generate_point_permutations(two, "FS"))


class StroudSecrest63_53(AbstractGaussianCubature):
    """Degree 5, k = 2^d + 2d, d > 2: (+-sqrt((d+2)/2), 0, ...)_FS and
    (+-sqrt((d+2)/(d-2)), ...)_R (E_n^{r^2} 5-3; _formulae.py:409-487)."""

    degree = 5

    def __init__(self, region: GaussianRegion):
        super().__init__(region)
        d = self.region.dimension
        if d <= 2:
            raise ValueError(f"StroudSecrest63_53 is only valid for regions with d > 2; got d={d}")

    @cached_property
    def point_count(self) -> int:
        return 2 * self.region.dimension + 2 ** self.region.dimension

    @cached_property
    def weights(self):
        d = self.region.dimension
        return (8 * d / (d + 2) ** 2 / self.points[0].shape[0], ((d - 2) / (d + 2)) ** 2 / self.points[1].shape[0])

    @cached_property
    def points(self):
        d = self.region.dimension
        axis = np.zeros(d)
        axis[0] = np.sqrt((d + 2) / 2)
        corner = np.sqrt((d + 2) / (d - 2)) * np.ones(d)
        return (_generate_point_permutations(axis, "FS"), _generate_point_permutations(corner, "R"))


builtin_cubature_registry = all_subclasses(AbstractCubature)
"""Every cubature formula defined above (_formulae.py:561-566)."""


def search_cubature_registry(region, degree=None, sparse_only=False, minimal_only=False,
                             searchable_formulae=None):
    """Formulae of the given degree (any if None) that can be built for `region`, sorted by point
    count; `minimal_only` keeps those with the fewest points (_formulae.py:569-625)."""
    pool = builtin_cubature_registry if searchable_formulae is None else searchable_formulae
    found = []
    for cls in pool:
        if getattr(cls, "__abstractmethods__", None):
            continue
        try:
            formula = cls(region)
        except (ValueError, TypeError):
            continue
        if (degree is not None and cls.degree != degree) or (sparse_only and not cls.sparse):
            continue
        found.append(formula)
    if minimal_only and found:
        fewest = min(f.point_count for f in found)
        found = [f for f in found if f.point_count == fewest]
    return sorted(found, key=lambda f: f.point_count)
