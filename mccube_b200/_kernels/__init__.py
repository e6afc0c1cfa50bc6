"""Kernel library (mirror of /root/reference/mccube/_kernels/__init__.py:6-17) plus the
BoxPartitionKernel the brief adds."""
from .base import (
    AbstractKernel as AbstractKernel,
    AbstractPartitioningKernel as AbstractPartitioningKernel,
    AbstractRecombinationKernel as AbstractRecombinationKernel,
    PartitioningRecombinationKernel as PartitioningRecombinationKernel,
)
from .box import BoxPartitionKernel as BoxPartitionKernel
from .random import (
    MonteCarloKernel as MonteCarloKernel,
    MonteCarloPartitioningKernel as MonteCarloPartitioningKernel,
)
from .stratified import StratifiedPartitioningKernel as StratifiedPartitioningKernel
from .tree import BinaryTreePartitioningKernel as BinaryTreePartitioningKernel
