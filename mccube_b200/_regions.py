"""Integration regions (mirror of /root/reference/mccube/_regions.py:9-50)."""
from __future__ import annotations

import abc


class AbstractRegion(abc.ABC):
    """Measure space over which cubature formulae are defined (_regions.py:9-25)."""

    dimension: int

    def __init__(self, dimension: int):
        self.dimension = int(dimension)

    @property
    @abc.abstractmethod
    def volume(self) -> float:
        ...

    def __repr__(self):
        return f"{type(self).__name__}(dimension={self.dimension})"

    def __eq__(self, other):
        return type(self) is type(other) and self.dimension == other.dimension

    def __hash__(self):
        return hash((type(self).__name__, self.dimension))


class GaussianRegion(AbstractRegion):
    """R^d with the standard Gaussian measure; volume 1 (_regions.py:28-50)."""

    @property
    def volume(self) -> float:
        return 1.0
