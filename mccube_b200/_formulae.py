"""Cubature formulae on the hot path (mirror of /root/reference/mccube/_formulae.py:33-227).

Only `Hadamard` is on the north-star path; the Stroud-Secrest formulae of
_formulae.py:230-487 are host-side NumPy constants outside this round's scope
(any `[k, d]` point table can still be fed to the CUDA kernels through
`mccb200_cubature.points`).
"""
from __future__ import annotations

import abc
from functools import cached_property

import numpy as np

from ._regions import AbstractRegion, GaussianRegion


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
