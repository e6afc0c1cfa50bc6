"""Pins the oracle's whole chain to a number the REFERENCE itself published.

docs/quickstart.md:184-208 of the reference prints, for the MCCube ULA with
`MonteCarloKernel(n, key=jr.key(42))`, Hadamard cubature, n=512, d=10, x64, 1024 epochs:
    2-Wasserstein distance: (3.3503528099696664+0j)
evaluated on `sol.ys[-128]` of an 11-entry save grid, i.e. (JAX clamps the out-of-range
index) the first saved time t1 - 10*dt0, reached after 1014 steps.  Reproducing it needs
every piece to be right: jr.key/split/normal for y0, fold_in of the float32 bits of the
OUTER t0 (quirk Q2), split_by_tree, the 2-round stable-sort shuffle truncated to n,
the child order c = i*k + j, the Hadamard row order, the Euler update and diffrax's
time stepping.  Any index error moves W2 by O(0.1); we match to < 1e-9.
"""
import numpy as np

from oracle import jax_random as jr
from oracle import mccube_ref as ref

PUBLISHED_W2 = 3.3503528099696664  # /root/reference/docs/quickstart.md:207


def test_quickstart_w2_is_reproduced():
    n, d = 512, 10
    key0 = jr.prng_key(42)
    key, _ = jr.split(key0, 2)                                  # quickstart.md:57
    y0 = jr.normal(key, n * d, np.float64).reshape(n, d)          # multivariate_normal(0, I)
    t0, dt0, n_epochs = 0.0, 0.05, 1024
    t1 = t0 + dt0 * n_epochs
    mean, inv_var = 2 * np.ones(d), np.ones(d) / 3.0
    kern = ref.MonteCarloKernel(n, key=key0)
    _, saved = ref.diffeqsolve(y0, t0, t1, dt0, lambda y: ref.gaussian_diag_drift(y, mean, inv_var),
                               ref.hadamard_points(d), ref.hadamard_weights(d), np.sqrt(2.0), kern,
                               time_dtype=np.float64, save_steps={1013})
    y = saved[1013]
    w2 = ref.gaussian_wasserstein_metric((mean, y.mean(0)), (3 * np.eye(d), np.cov(y, ddof=0, rowvar=False)))
    assert abs(w2.real - PUBLISHED_W2) < 1e-9, w2
    assert w2.imag == 0.0
