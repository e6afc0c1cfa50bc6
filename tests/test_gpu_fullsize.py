"""Parity evidence at BASELINE.json's FULL sizes, where the oracle cannot materialise the
expansion (n*k*d = 512 GiB): size-independent properties plus bit-exact agreement between the
two independent CUDA formulations (fused counting-sort vs child keys -> stable radix sort)."""
import math

import numpy as np
import pytest
import torch

import mccube_b200 as mccube
from mccube_b200 import _lib
from mccube_b200 import diffrax_compat as dfx

pytestmark = pytest.mark.gpu


def _problem(n, d, dev, seed=0):
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    y0 = torch.randn((n, d), generator=gen, device=dev) * 1.5 + 1.0
    drift = mccube.GaussianDiagDrift(np.full(d, 2.0), 3.0, device=dev)
    terms = mccube.MCCTerm(
        dfx.ODETerm(drift),
        dfx.WeaklyDiagonalControlTerm(lambda t, y, args: math.sqrt(2.0),
                                      mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d)))))
    return y0, terms


def test_config3_fused_step_full_size(dev):
    """BASELINE config 3: n=2^24, d=64 (2^31 implicit children), Box + PRK + MonteCarloKernel."""
    n, d, per = 1 << 24, 64, 64
    dims, bits = (0, 1, 2, 3), 3
    y0, terms = _problem(n, d, dev)
    m = n // per
    kernel = mccube.PartitioningRecombinationKernel(mccube.BoxPartitionKernel(m, dims=dims, bits=bits),
                                                    mccube.MonteCarloKernel(per, key=42))
    solver = mccube.MCCSolver(dfx.Euler(), kernel)
    t0, t1 = np.float32(0.05), np.float32(0.1)
    y1, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    assert solver.last_path == "fused:box_prk_step" and y1.shape == (n, d)
    assert torch.isfinite(y1).all()
    # (1) determinism: the step is a pure function of (key, t, y0)
    y1b, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    assert torch.equal(y1, y1b)
    del y1b
    # (2) sortedness: partitions are consecutive chunks of the key-sorted children, so box keys of
    #     the output are non-decreasing from one partition to the next
    drift, cub, dt, k, dd = solver._describe(terms, y0, *solver.substep_times(t0, t1)[0], None)
    bounds = _lib.box_child_bounds(y0, d, drift, dt, cub, dims, bits)
    keys = _lib.box_keys(y1, dims, bits, bounds).view(m, per).to(torch.int64)
    kmin, kmax = keys.min(dim=1).values, keys.max(dim=1).values
    assert bool((kmax[:-1] <= kmin[1:]).all())
    del keys
    # (3) the independent formulation (child keys for all 2^31 children -> stable radix sort ->
    #     permuted gather) gives the same state bit for bit
    solver2 = mccube.MCCSolver(dfx.Euler(), kernel)
    solver2.use_box_prk_kernel = False
    y2, *_ = solver2.step(terms, t0, t1, y0, None, None, False)
    assert solver2.last_path == "fused:box_child_keys+sort+expand_select"
    assert torch.equal(y1, y2)
    del y2
    _lib.release_workspace()
    # (4) one Langevin step contracts the mean towards the target by dt/var
    mean0, _ = mccube.mean_and_cov(y0)
    mean1, _ = mccube.mean_and_cov(y1)
    expect = mean0 + float(dt) * (2.0 - mean0) / 3.0
    assert float((mean1 - expect).abs().max()) < 0.02


def test_config2_monte_carlo_full_size(dev):
    """BASELINE config 2: n=2^20, d=10, MonteCarloKernel (default without replacement): replay is
    bit-identical, selected children are distinct, and the fused step equals materialise +
    gather with the (oracle-verified) index vector."""
    n, d = 1 << 20, 10
    y0, terms = _problem(n, d, dev, seed=1)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=42))
    t0, t1 = np.float32(0.05), np.float32(0.1)
    y1, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    y1b, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    assert torch.equal(y1, y1b)
    k = 32
    idx = _lib.mc_indices(_lib.make_rng((0, 42), t0), n * k, n, False, dev)  # bit-exact vs oracle: test_gpu_parity
    assert torch.unique(idx).numel() == n
    drift, cub, dt, kk, dd = solver._describe(terms, y0, *solver.substep_times(t0, t1)[0], None)
    children = _lib.expand_all(y0, d, False, drift, dt, cub)      # 2^25 x 10 floats = 1.25 GiB, fits
    want = _lib.gather_rows(children, idx)
    assert torch.equal(y1, want)


def test_monte_carlo_config3_size_with_replacement(dev):
    """n=2^24, d=64, global MonteCarloKernel(with_replacement=True): every sampled output row is
    its parent plus drift plus a +-scale noise vector (checked through the index vector)."""
    n, d = 1 << 24, 64
    y0, terms = _problem(n, d, dev, seed=2)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, True, key=42))
    t0, t1 = np.float32(0.1), np.float32(0.15)
    y1, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    idx = _lib.mc_indices(_lib.make_rng((0, 42), t0), n * 128, n, True, dev).to(torch.int64) & 0xFFFFFFFF
    assert int(idx.max()) < n * 128
    parent = idx // 128
    drift, cub, dt, k, dd = solver._describe(terms, y0, *solver.substep_times(t0, t1)[0], None)
    sample = torch.arange(0, n, 4099, device=dev)
    yp = y0[parent[sample]]
    centre = yp + (float(dt) * ((2.0 - yp) * (1.0 / 3.0)))
    resid = (y1[sample] - centre).abs()
    assert float((resid - float(cub.scale)).abs().max()) < 1e-5


def test_config4_full_covariance_step_full_size(dev):
    """BASELINE config 4 at its FULL size (n = 2^20, d = 512, k = 1024: 2^30 implicit children): the tcgen05
    drift and the fused step, checked on sampled rows against float64.  Tolerances: drift rows within 1e-5 RELATIVE
    to each row's own norm-scale (max |F_row|), states within 1e-5 relative (the north-star bar)."""
    n, d = 1 << 20, 512
    rs = np.random.default_rng(1)
    am = rs.normal(size=(d, d))
    cov = am @ am.T / d + np.eye(d)
    mu = rs.normal(size=d) * 2.0
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    y0 = torch.randn((n, d), generator=gen, device=dev) * 1.5 + torch.tensor(mu, dtype=torch.float32, device=dev)
    drift = mccube.GaussianFullDrift(mu, cov, device=dev)
    f = drift(0.0, y0)
    assert drift.last_path == "tcgen05:3xtf32"
    rows = torch.arange(0, n, 1237, device=dev)                      # 848 sampled rows
    ys = y0[rows].cpu().numpy().astype(np.float64)
    prec = drift.prec.cpu().numpy().astype(np.float64)
    want = -(ys - drift.mean.cpu().numpy().astype(np.float64)) @ prec
    got = f[rows].cpu().numpy().astype(np.float64)
    row_scale = np.abs(want).max(axis=1, keepdims=True)
    assert (np.abs(got - want) <= 1e-5 * row_scale).all(), float((np.abs(got - want) / row_scale).max())
    # one whole step: global MonteCarloKernel with replacement, drift handed over as an array
    terms = mccube.MCCTerm(
        dfx.ODETerm(drift),
        dfx.WeaklyDiagonalControlTerm(lambda t, y, args: math.sqrt(2.0),
                                      mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d)))))
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, True, key=42))
    t0, t1 = np.float32(0.05), np.float32(0.1)
    y1, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    k = 1024
    idx = (_lib.mc_indices(_lib.make_rng((0, 42), t0), n * k, n, True, dev).to(torch.int64) & 0xFFFFFFFF)[rows]
    parent, point = (idx // k).cpu().numpy(), (idx % k).cpu().numpy()
    yp = y0[torch.as_tensor(parent, device=dev)].cpu().numpy().astype(np.float64)
    fp = -(yp - drift.mean.cpu().numpy().astype(np.float64)) @ prec
    dt = float(np.float32(t1) - np.float32(t0))     # quirk Q1 shifts the interval, not its length
    scale = float(np.float32(np.float32(math.sqrt(2.0)) * np.float32(math.sqrt(np.float32(dt)))))
    cols = np.arange(d)
    half = k // 2 - 1
    par = np.array([[bin(int(j) & half & int(c)).count("1") & 1 for c in cols] for j in point])
    sign = 1.0 - 2.0 * (par ^ (point[:, None] > half))
    want1 = yp + dt * fp + scale * sign
    got1 = y1[rows].cpu().numpy().astype(np.float64)
    assert np.abs(got1 - want1).max() <= 1e-5 * np.abs(want1).max()


def test_config5_single_gpu_in_kernel_draw_full_size(dev):
    """BASELINE config 5's per-GPU share (n = 2^24, d = 64, k = 128: 2^31 children), MonteCarloKernel(with_replacement=True):
    the gather that draws its own child ids (no index vector in HBM) equals index draw + gather bit for bit; every drawn
    id is in range and its parent row reproduces the child (Hadamard sign + Euler step recomputed in torch for a sample)."""
    n, d = 1 << 24, 64
    y0, terms = _problem(n, d, dev, seed=3)
    kern = mccube.MonteCarloKernel(n, True, key=7)
    solver = mccube.MCCSolver(dfx.Euler(), kern)
    t0, t1 = np.float32(0.2), np.float32(0.25)
    y1, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    assert solver.last_path == "fused:expand_select_draw"
    solver.fuse_draw = False
    y2, *_ = solver.step(terms, t0, t1, y0, None, None, False)
    assert solver.last_path == "fused:mc_indices+expand_select"
    assert torch.equal(y1, y2)
    del y2
    drift, cub, dt, k, _ = solver._describe(terms, y0, *solver.substep_times(t0, t1)[0], None)
    idx = kern.indices(t0, n * k, n, dev).to(torch.int64) & 0xFFFFFFFF
    assert int(idx.min()) >= 0 and int(idx.max()) < n * k
    rows = torch.arange(0, n, 4099, device=dev)
    parent, point = idx[rows] // k, idx[rows] % k
    cols = torch.arange(d, device=dev)
    # Hadamard row j of the doubled matrix [H; -H]: sign = (-1)^(popcount(j & c)) for j < k/2, negated above
    half = k // 2
    jj = (point % half)[:, None] & cols[None, :]
    pop = torch.zeros_like(jj)
    for b in range(8):
        pop += (jj >> b) & 1
    sign = 1.0 - 2.0 * ((pop & 1) ^ (point >= half)[:, None].long()).float()
    scale = float(cub.scale)                                         # g * sqrt(dt), as the step computed it
    yp = y0[parent]
    f = (torch.full((d,), 2.0, device=dev) - yp) * np.float32(1.0 / 3.0)
    want = yp + (np.float32(dt) * f + sign * float(scale))
    assert torch.allclose(y1[rows], want, rtol=1e-5, atol=1e-5)


def test_device_median_split_partitioner_large(dev):
    """2^20 x 64 points, leaves of 64 (14 levels): the device tree is a permutation, and every leaf is separated from its
    sibling at the median of the widest dimension of their parent (the property sklearn's split rule defines; the host
    build of this size takes 11 s and is compared at small sizes in test_gpu_parity)."""
    n, d, leaf = 1 << 20, 64, 64
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    y = torch.randn((n, d), generator=gen, device=dev) * torch.linspace(0.5, 3.0, d, device=dev)
    levels = _lib.kd_partition_levels(n, leaf)
    assert levels == 14
    idx = _lib.kd_partition(y, d, levels).to(torch.int64)
    assert torch.equal(torch.sort(idx).values, torch.arange(n, device=dev))
    pts = y[idx].view(n // (2 * leaf), 2 * leaf, d)                    # parents of the leaves
    spread = pts.max(dim=1).values.double() - pts.min(dim=1).values.double()
    j = spread.argmax(dim=1)
    col = torch.gather(pts, 2, j[:, None, None].expand(-1, 2 * leaf, 1)).squeeze(2)
    assert bool((col[:, :leaf].max(dim=1).values <= col[:, leaf:].min(dim=1).values).all())
    # the root split: the first half holds the n/2 smallest values of the widest dimension
    j0 = int((y.max(0).values.double() - y.min(0).values.double()).argmax())
    med = torch.sort(y[:, j0]).values[n // 2]
    assert bool((y[idx[: n // 2], j0] <= med).all()) and bool((y[idx[n // 2:], j0] >= med).all())
