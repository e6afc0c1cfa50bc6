"""The reference's README example (/root/reference/README.md:43-90) on this library: the same names, the same
constructor arguments, the same call to `diffeqsolve` -- `diffrax` is `mccube_b200.diffrax_compat` (diffrax and
JAX are not installable in this image), arrays are CUDA tensors, and the drift is either any callable returning
an [n, d] tensor (as below, evaluated on the parents) or the analytic `GaussianDiagDrift` the fused kernel
inlines.  Needs a B200 (no CPU fallback)."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mccube_b200 import (  # noqa: E402
    GaussianDiagDrift,
    GaussianRegion,
    Hadamard,
    LocalLinearCubaturePath,
    MCCSolver,
    MCCTerm,
    MonteCarloKernel,
    diffrax_compat as diffrax,
    gaussian_wasserstein_metric,
    mean_and_cov,
)

dev = torch.device("cuda:0")
key = 42                                   # jr.key(42)
n, d = 512, 10
t0 = 0.0
epochs = 512
dt0 = 0.05
t1 = t0 + dt0 * epochs
y0 = torch.ones((n, d), device=dev)

target_mean = 2 * np.ones(d)
target_cov = 3 * np.eye(d)
mean_t = torch.tensor(target_mean, dtype=torch.float32, device=dev)

# grad log N(p; mean, cov) = -(p - mean) / 3, written out (no autodiff here)
ode = diffrax.ODETerm(lambda t, p, args: (mean_t - p) * (1.0 / 3.0))
cde = diffrax.WeaklyDiagonalControlTerm(
    lambda t, p, args: math.sqrt(2.0),
    LocalLinearCubaturePath(Hadamard(GaussianRegion(d))),
)
terms = MCCTerm(ode, cde)
solver = MCCSolver(diffrax.Euler(), MonteCarloKernel(n, key=key))

sol = diffrax.diffeqsolve(terms, solver, t0, t1, dt0, y0)
res_mean, res_cov = mean_and_cov(sol.ys[-1])          # jnp.mean(axis=0), jnp.cov(rowvar=False)
metric = gaussian_wasserstein_metric((target_mean, res_mean), (target_cov, res_cov))
print(f"Result 2-Wasserstein distance: {metric}   [{solver.last_path}]")

# the same solve with the analytic drift inlined in the fused kernel, replayed from one CUDA graph
terms_fused = MCCTerm(diffrax.ODETerm(GaussianDiagDrift(target_mean, 3.0, device=dev)), cde)
solve = diffrax.capture_solve(terms_fused, solver, t0, t1, dt0, y0)
sol_g = solve(y0)
res_mean, res_cov = mean_and_cov(sol_g.ys[-1])
print(f"captured solve:                {gaussian_wasserstein_metric((target_mean, res_mean), (target_cov, res_cov))}")
