"""Analytic drifts the fused kernels recognise.

The reference takes any callable as the ODE vector field (README.md:72-76 uses
`vmap(grad(multivariate_normal.logpdf))`).  A callable of one of these classes is
pattern-matched by `MCCSolver.step` and evaluated inside the fused kernel; any other
callable is evaluated by the caller on the parents `[n, d]` and handed to the kernel as
an array (drift kind 1) -- the generic path, never a silent CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class GaussianDiagDrift:
    """grad log N(y; mean, diag(var)) = (mean - y) * (1 / var)   [drift kind 2]."""

    def __init__(self, mean, var, device="cuda"):
        mean = np.asarray(mean, dtype=np.float64)
        var = np.broadcast_to(np.asarray(var, dtype=np.float64), mean.shape)
        self.mean = torch.tensor(mean, dtype=torch.float32, device=device)
        self.inv_var = torch.tensor((1.0 / var).astype(np.float32), dtype=torch.float32, device=device)

    def __call__(self, t, y, args=None):
        # generic-path evaluation (same values as the fused form)
        return (self.mean - y) * self.inv_var


class GaussianFullDrift:
    """grad log N(y; mean, cov) = -(y - mean) @ cov^-1 for a full covariance (config 4).
    Evaluated into an [n, d] array (drift kind 1): on the tensor cores (tcgen05 kind::tf32 with
    the 3xTF32 split, csrc/gemm_drift.cu) when d % 32 == 0, otherwise by the fp32 SIMT tiles.
    `tensor_cores=False` forces the SIMT kernel."""

    def __init__(self, mean, cov, device="cuda", tensor_cores: bool = True):
        cov = np.asarray(cov, dtype=np.float64)
        self.mean = torch.tensor(np.asarray(mean), dtype=torch.float32, device=device)
        self.prec = torch.tensor(np.linalg.inv(cov).astype(np.float32), dtype=torch.float32, device=device)
        self.tensor_cores = tensor_cores
        self._prec_split = None
        self.last_path = None

    def __call__(self, t, y, args=None):
        d = self.mean.numel()
        y = y.contiguous()
        if self.tensor_cores and _lib.gaussian_full_drift_tc_supported(y.shape[0], d, y.shape[1]) \
                and y.data_ptr() % 16 == 0:
            if self._prec_split is None:
                self._prec_split = _lib.gaussian_full_drift_tc_prepare(self.prec)
            self.last_path = "tcgen05:3xtf32"
            return _lib.gaussian_full_drift_tc(y, d, self.mean, self._prec_split)
        self.last_path = "simt:fp32"
        return _lib.gaussian_full_drift(y, d, self.mean, self.prec)
