"""mccube_b200 -- B200-native (sm_100a) implementation of MCCube's data-parallel hot path.

The export list mirrors /root/reference/mccube/__init__.py:4-51 for everything on the
path (same names, constructor argument order and error behaviour), plus:
  * `BoxPartitionKernel`     -- named by the brief, absent from the reference;
  * `GaussianDiagDrift`, `GaussianFullDrift` -- analytic drifts the fused kernels recognise;
  * `diffrax_compat`         -- stand-ins for the diffrax protocol (diffrax is not installed here);
  * `mean_and_cov`           -- the jnp.mean / jnp.cov pair of README.md:85-86 as GPU reductions.
All compute goes through the C ABI in `include/mccube_b200.h` (`mccube_b200._lib`).
"""
from . import diffrax_compat as diffrax_compat
from ._drift import GaussianDiagDrift as GaussianDiagDrift, GaussianFullDrift as GaussianFullDrift
from ._formulae import (
    AbstractCubature as AbstractCubature,
    AbstractGaussianCubature as AbstractGaussianCubature,
    Hadamard as Hadamard,
    StroudSecrest63_31 as StroudSecrest63_31,
    StroudSecrest63_32 as StroudSecrest63_32,
    StroudSecrest63_52 as StroudSecrest63_52,
    StroudSecrest63_53 as StroudSecrest63_53,
    search_cubature_registry as search_cubature_registry,
)
from ._kernels import (
    AbstractKernel as AbstractKernel,
    AbstractPartitioningKernel as AbstractPartitioningKernel,
    AbstractRecombinationKernel as AbstractRecombinationKernel,
    BinaryTreePartitioningKernel as BinaryTreePartitioningKernel,
    BoxPartitionKernel as BoxPartitionKernel,
    MonteCarloKernel as MonteCarloKernel,
    MonteCarloPartitioningKernel as MonteCarloPartitioningKernel,
    PartitioningRecombinationKernel as PartitioningRecombinationKernel,
    StratifiedPartitioningKernel as StratifiedPartitioningKernel,
)
from ._metrics import (
    center_of_mass as center_of_mass,
    gaussian_wasserstein_metric as gaussian_wasserstein_metric,
    mean_and_cov as mean_and_cov,
    pairwise_metric as pairwise_metric,
)
from ._path import (
    AbstractCubaturePath as AbstractCubaturePath,
    LocalLinearCubaturePath as LocalLinearCubaturePath,
)
from ._regions import AbstractRegion as AbstractRegion, GaussianRegion as GaussianRegion
from ._solvers import MCCSolver as MCCSolver
from ._term import MCCTerm as MCCTerm
from ._utils import (
    all_subclasses as all_subclasses,
    nop as nop,
    pack_particles as pack_particles,
    unpack_particles as unpack_particles,
)

__version__ = "0.1.0"
