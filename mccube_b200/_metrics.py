"""Metrics (mirror of /root/reference/mccube/_metrics.py:17-96).

The O(n d) / O(n d^2) reductions run on the GPU (`mccb200_moments_*`, fp64 accumulators);
the d x d matrix square roots of the Gaussian 2-Wasserstein formula are a host-side
eigendecomposition (d <= 512: microseconds, `jax.scipy.linalg.sqrtm` is CPU-only in JAX
as well).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._tree import tree_map


def _to_np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().double().numpy()
    return np.asarray(x, dtype=np.float64)


def _sqrtm_psd(a: np.ndarray) -> np.ndarray:
    w, v = np.linalg.eigh((a + a.T) / 2)
    return (v * np.sqrt(np.clip(w, 0.0, None))) @ v.T


def gaussian_wasserstein_metric(means, covs):
    """||m1 - m2||^2 + tr(c1 + c2 - 2 sqrtm(sqrtm(c2) c1 sqrtm(c2))) (_metrics.py:39-44).
    Returned as complex128 like the reference's Schur `sqrtm` output
    (/root/reference/tests/test_metrics.py:13); for SPD inputs the imaginary part is 0."""
    m1, m2 = (_to_np(m) for m in means)
    c1, c2 = (_to_np(c) for c in covs)
    root_c2 = _sqrtm_psd(c2)
    mean_dist = np.linalg.norm(m1 - m2) ** 2
    cov_dist = np.trace(c1 + c2 - 2 * _sqrtm_psd(root_c2 @ c1 @ root_c2))
    return np.complex128(mean_dist + cov_dist)


def center_of_mass(particles, weights=None):
    """`jnp.average(p, 0, w)` per leaf (_metrics.py:79-96), on the GPU."""

    def _com(p, w):
        if p.dim() == 1:
            p = p[:, None]
            squeeze = True
        else:
            squeeze = False
        mean, _ = _lib.column_means(p.contiguous().float(), p.shape[1], None if w is None else w.float())
        mean = mean.to(p.dtype)
        return mean[0] if squeeze else mean

    if weights is None:
        return tree_map(lambda p: _com(p, None), particles)
    return tree_map(_com, particles, weights)


def mean_and_cov(particles: torch.Tensor, weights=None, ddof: int = 1):
    """`jnp.mean(p, 0)` and `jnp.cov(p, rowvar=False, ddof=ddof, aweights=w)`
    (/root/reference/README.md:85-86; docs/quickstart.md:98-101 uses ddof=0 and aweights),
    as float64 tensors, from one sum pass and one centred second-moment pass on the GPU."""
    d = particles.shape[1]
    if weights is not None:
        weights = torch.as_tensor(weights, device=particles.device).to(torch.float32)   # the kernels read float32
    mean, wtot, m2 = _lib.moments(particles.contiguous(), d, weights)
    if weights is None:
        cov = m2 / (wtot - ddof)
    else:
        # numpy/jnp.cov with aweights: fact = sum(w) - ddof * sum(w^2) / sum(w)
        w64 = weights.double()
        fact = wtot - ddof * (w64 * w64).sum() / wtot
        cov = m2 / fact
    return mean, cov


def euclidean_metric(xs, ys):
    """_metrics.py:47-52."""
    return tree_map(lambda x, y: torch.linalg.norm(x - y, dim=-1), xs, ys)


def squared_euclidean_metric(xs, ys):
    """_metrics.py:55-60."""
    return tree_map(lambda x, y: torch.linalg.norm(x - y, dim=-1) ** 2, xs, ys)


def pairwise_metric(xs, ys, metric=euclidean_metric):
    """_metrics.py:63-76: result[j, i] = metric(xs[i], ys[j]).  O(n^2) helper, not on the
    hot path (out of scope for kernels); plain broadcast."""
    return tree_map(lambda x, y: metric(x[None, :, :].expand(y.shape[0], -1, -1),
                                        y[:, None, :].expand(-1, x.shape[0], -1)), xs, ys)
