"""Cubature control paths (mirror of /root/reference/mccube/_path.py:13-76)."""
from __future__ import annotations

import abc

import numpy as np

from ._formulae import AbstractGaussianCubature


class AbstractCubaturePath(abc.ABC):
    """_path.py:13-38."""

    @property
    @abc.abstractmethod
    def weights(self):
        ...


class LocalLinearCubaturePath(AbstractCubaturePath):
    r"""Piecewise linear cubature paths f(t0, t1) = sqrt(t1 - t0) M (_path.py:41-76)."""

    t0 = 0.0
    t1 = 1.0

    def __init__(self, gaussian_cubature: AbstractGaussianCubature):
        self.gaussian_cubature = gaussian_cubature

    def coefficient(self, t0, t1=None, dtype=np.float32):
        """`where(t0 == t1, 0, sqrt(t1 - t0))` in the time dtype (_path.py:63-69)."""
        dtype = np.dtype(dtype).type
        if t1 is None:
            t0, t1 = dtype(self.t0), t0
        t0, t1 = dtype(t0), dtype(t1)
        return dtype(0.0) if t0 == t1 else dtype(np.sqrt(dtype(t1 - t0)))

    def evaluate(self, t0, t1=None, left: bool = True, dtype=np.float32) -> np.ndarray:
        del left
        coeff = self.coefficient(t0, t1, dtype)
        return (coeff * self.gaussian_cubature.stacked_points.astype(dtype)).astype(dtype)

    @property
    def weights(self) -> np.ndarray:
        return self.gaussian_cubature.stacked_weights
