"""GPU parity tests: the CUDA path (called through the C ABI) against the oracle and the
committed golden vectors.  Integer / index work must be BIT-EXACT; fp32 particle states are
bit-exact too for the elementwise drifts (same op order, no FMA), and within 1e-5 relative
(the north-star tolerance) wherever a reduction order differs; W2 within 1e-4."""
import math

import numpy as np
import pytest
import torch

import mccube_b200 as mccube
from mccube_b200 import _lib
from mccube_b200 import diffrax_compat as dfx
from oracle import jax_random as jr
from oracle import mccube_ref as ref

from .conftest import GOLDEN

pytestmark = pytest.mark.gpu

KEY = (0, 42)
SQRT2 = np.float32(np.sqrt(2.0))


def u32(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().astype(np.int64) & 0xFFFFFFFF


def cuda(a, dev):
    return torch.tensor(np.ascontiguousarray(a), device=dev)


# ------------------------------------------------------------------ A8: indices
def test_mc_indices_golden(dev):
    z = np.load(GOLDEN / "mc_indices.npz")
    for name in z.files:
        P, n, t, r, p = name.split("_")
        P, n, t, r, p = int(P[1:]), int(n[1:]), float(t[1:]), bool(int(r[1:])), bool(int(p[1:]))
        rng = _lib.make_rng(KEY, t, partitionable=p)
        got = u32(_lib.mc_indices(rng, P, n, r, dev))
        assert np.array_equal(got, z[name]), name


@pytest.mark.parametrize("partitionable", [False, True])
@pytest.mark.parametrize("replace", [False, True])
def test_mc_indices_vs_oracle_ragged(dev, partitionable, replace):
    # up to 16384 inputs the shuffle is one single-CTA launch (per-warp chunks: sizes around the 32 / 1024
    # boundaries, the largest size and one past it), above that the multi-kernel radix sort
    for (P, n, t, leaf, n_leaves) in [(1, 1, 0.0, 0, 1), (2, 2, 0.3, 0, 1), (31, 7, 0.1, 0, 1), (33, 33, 0.1, 0, 1),
                                      (1023, 100, 0.2, 0, 1), (1025, 1025, 0.2, 0, 1), (4097, 4097, -1.5, 1, 3),
                                      (8192, 64, 0.05, 0, 1), (12345, 12345, 0.7, 0, 1), (16384, 512, 0.0, 0, 1),
                                      (16385, 512, 0.0, 0, 1), ((1 << 20) + 3, 5000, 12.75, 0, 2)]:
        key = jr.mc_step_key(jr.prng_key(42), t, leaf=leaf, n_leaves=n_leaves, partitionable=partitionable)
        want = jr.choice_indices(key, P, n, replace, partitionable=partitionable)
        rng = _lib.make_rng(KEY, t, leaf=leaf, n_leaves=n_leaves, partitionable=partitionable)
        got = u32(_lib.mc_indices(rng, P, n, replace, dev))
        assert np.array_equal(got, want), (P, n, t)
        if not replace:
            assert len(np.unique(got)) == n


def test_mc_indices_config2_size_properties(dev):
    """BASELINE config 2 (n=2^20, d=10 -> P=2^25, 3 rounds): full-size run; a permutation
    prefix has no repeats, and the head agrees with the oracle's (slow) full shuffle."""
    P, n = 1 << 25, 1 << 20
    rng = _lib.make_rng(KEY, 0.05)
    got = u32(_lib.mc_indices(rng, P, n, False, dev))
    assert got.max() < P and len(np.unique(got)) == n
    want = jr.choice_indices(jr.mc_step_key(jr.prng_key(42), 0.05), P, n, False)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("partitionable", [False, True])
def test_mc_indices_truncated_last_round(dev, partitionable):
    """Large inputs with count << n_inputs: the last shuffle round only orders the `count` smallest keys
    (one MSD pass + a sort of a short prefix); bit-exact vs the oracle's full shuffle, ragged size."""
    P, n = (1 << 21) + 123, 1 << 15
    rng = _lib.make_rng(KEY, 0.15, partitionable=partitionable)
    got = u32(_lib.mc_indices(rng, P, n, False, dev))
    want = jr.choice_indices(jr.mc_step_key(jr.prng_key(42), 0.15, partitionable=partitionable), P, n, False,
                             partitionable=partitionable)
    assert np.array_equal(got, want)


def test_sort_pairs_stable(dev):
    rs = np.random.default_rng(0)
    for n, hi, bits in [(1, 10, 32), (31, 4, 8), (4096, 1 << 32, 32), (4097, 256, 8), (100003, 1 << 12, 12),
                        (1 << 20, 1 << 32, 32), (300000, 7, 3)]:
        keys = rs.integers(0, hi, size=n, dtype=np.uint64).astype(np.uint32)
        vals = rs.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        order = np.argsort(keys, kind="stable")
        kt, vt = cuda(keys.view(np.int32), dev), cuda(vals.view(np.int32), dev)
        assert np.array_equal(u32(_lib.sort_pairs(kt, None, bits)), order)
        ko, vo = _lib.sort_pairs(kt, vt, bits, want_keys=True)
        assert np.array_equal(u32(ko), keys[order]) and np.array_equal(u32(vo), vals[order])


# ------------------------------------------------------------------ A1-A5: expansion
def _setup(n, d, seed=0, iso=False):
    rs = np.random.default_rng(seed)
    y0 = rs.normal(size=(n, d)).astype(np.float32)
    mean = (2 * np.ones(d) if iso else rs.normal(size=d)).astype(np.float32)
    inv_var = (np.ones(d) / 3.0 if iso else 1.0 / rs.uniform(0.5, 4.0, size=d)).astype(np.float32)
    return y0, mean, inv_var


@pytest.mark.parametrize("d", [1, 3, 10, 64, 130])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_expand_all_bit_exact(dev, d, kind):
    n = 37
    y0, mean, inv_var = _setup(n, d, seed=d)
    t0s, t1s = np.float32(0.05), np.float32(0.1)
    if kind == 2:
        fn = lambda y: ref.gaussian_diag_drift(y, mean, inv_var)  # noqa: E731
    elif kind == 1:
        fn = lambda y: (np.sin(y) - y).astype(np.float32)  # noqa: E731
    else:
        fn = lambda y: np.zeros_like(y)  # noqa: E731
    M, lam = ref.hadamard_points(d), ref.hadamard_weights(d)
    want = ref.euler_expand(y0, fn, M, SQRT2, t0s, t1s)
    k = M.shape[0]
    dt = np.float32(t1s - t0s)
    scale = np.float32(SQRT2 * np.float32(np.sqrt(dt)))
    y0t = cuda(y0, dev)
    if kind == 2:
        drift = _lib.make_drift(2, mean=cuda(mean, dev), inv_var=cuda(inv_var, dev))
    elif kind == 1:
        drift = _lib.make_drift(1, f=cuda(fn(y0), dev))
    else:
        drift = _lib.make_drift(0)
    got = _lib.expand_all(y0t, d, False, drift, dt, _lib.make_cubature(k, scale=scale)).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    # generic [k, d] point table (any cubature formula): same bits as the Hadamard closed form
    table = (SQRT2 * ref.path_evaluate(M, t0s, t1s, np.float32)).astype(np.float32)
    got2 = _lib.expand_all(y0t, d, False, drift, dt, _lib.make_cubature(k, points=cuda(table, dev))).cpu().numpy()
    assert np.array_equal(got2, want)
    # select form
    idx = np.random.default_rng(1).integers(0, n * k, size=55)
    got3 = _lib.expand_select(y0t, d, False, drift, dt, _lib.make_cubature(k, scale=scale),
                              cuda(idx.astype(np.int32), dev)).cpu().numpy()
    assert np.array_equal(got3, want[idx])


def _terms(d, mean, inv_var, dev, fused=True):
    if fused:
        f = mccube.GaussianDiagDrift(mean, 1.0 / inv_var.astype(np.float64), device=dev)
        f.inv_var = cuda(inv_var, dev)  # exact fp32 bits of the test vector
        f.mean = cuda(mean, dev)
    else:
        mt, it = cuda(mean, dev), cuda(inv_var, dev)
        f = lambda t, y, args: (mt - y) * it  # noqa: E731
    ode = dfx.ODETerm(f)
    cde = dfx.WeaklyDiagonalControlTerm(lambda t, y, args: math.sqrt(2.0),
                                        mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d))))
    return mccube.MCCTerm(ode, cde)


@pytest.mark.parametrize("replace", [False, True])
@pytest.mark.parametrize("partitionable", [False, True])
@pytest.mark.parametrize("fused_drift", [True, False])
def test_step_monte_carlo_kernel_bit_exact(dev, replace, partitionable, fused_drift):
    n, d = 300, 10
    y0, mean, inv_var = _setup(n, d, seed=3)
    t0, t1 = np.float32(0.35), np.float32(0.4)
    kern = ref.MonteCarloKernel(n, replace, key=jr.prng_key(42), partitionable=partitionable)
    want = ref.mcc_step(y0, t0, t1, lambda y: ref.gaussian_diag_drift(y, mean, inv_var), ref.hadamard_points(d),
                        ref.hadamard_weights(d), SQRT2, kern)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, replace, key=42,
                                                                   threefry_partitionable=partitionable))
    y1, err, dense, state, res = solver.step(_terms(d, mean, inv_var, dev, fused_drift), t0, t1, cuda(y0, dev), None,
                                             None, False)
    assert solver.last_path.startswith("fused") and err is None and res == dfx.RESULTS.successful
    assert set(dense) == {"y0", "y1"}
    assert np.array_equal(y1.cpu().numpy(), want)


def test_readme_example_trajectory(dev):
    """BASELINE config 1 / README.md:60-89 in float32, through the mirrored public API."""
    z = np.load(GOLDEN / "readme_trajectory.npz")
    n, d = 512, 10
    t0, dt0 = 0.0, 0.05
    t1 = t0 + dt0 * 512
    for replace in (False, True):
        tag = f"r{int(replace)}"
        terms = _terms(d, (2 * np.ones(d)).astype(np.float32), (np.ones(d) / 3.0).astype(np.float32), dev)
        solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, replace, key=42))
        y0 = torch.ones((n, d), device=dev)
        sol = dfx.diffeqsolve(terms, solver, t0, t1, dt0, y0, saveat=dfx.SaveAt(steps=True))
        assert sol.ys.shape[0] == int(z[f"n_steps_{tag}"])
        for s in (0, 1, 63, sol.ys.shape[0] - 1):
            assert np.array_equal(sol.ys[s].cpu().numpy(), z[f"y_step{s}_{tag}"]), (tag, s)
        mean, cov = mccube.mean_and_cov(sol.ys[-1])
        w2 = mccube.gaussian_wasserstein_metric((2 * np.ones(d), mean), (3 * np.eye(d), cov))
        assert abs(w2.real - float(z[f"w2_{tag}"])) < 1e-4


def test_table_cubature_and_generic_kernel_protocol(dev):
    """User-defined kernels go through the generic path (expansion materialised on the GPU):
    /root/reference/tests/test_kernels.py:19-60 on a PyTree, plus a solver step."""

    class PartitioningKernel(mccube.AbstractPartitioningKernel):
        def __call__(self, t, particles, args, weighted=False):
            from mccube_b200._tree import tree_map
            return tree_map(lambda p, c: p.reshape(-1, c, p.shape[-1]), particles, self.partition_count)

    class RecombinationKernel(mccube.AbstractRecombinationKernel):
        def __call__(self, t, particles, args, weighted=False):
            from mccube_b200._tree import tree_map
            return tree_map(lambda p, c: p[:c], particles, self.recombination_count)

    y0 = torch.tensor([[2.0, 4.0, 6.0, 8.0]], device=dev).T
    y1 = torch.tensor([[1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0]], device=dev).T
    tree = [y0, y1]
    kernel = mccube.PartitioningRecombinationKernel(PartitioningKernel([2, 3]), RecombinationKernel([2, 3]))
    out = kernel(0.0, tree, None)
    assert all(torch.equal(a, b) for a, b in zip(out, tree))
    treew = [mccube.pack_particles(x, torch.ones(x.shape[0], device=dev)) for x in tree]
    outw = kernel(0.0, treew, None, weighted=True)
    assert all(torch.equal(a, b) for a, b in zip(outw, treew))

    n, d = 12, 3
    yy, mean, inv_var = _setup(n, d, seed=5)
    solver = mccube.MCCSolver(dfx.Euler(), RecombinationKernel(n))
    got, *_ = solver.step(_terms(d, mean, inv_var, dev), np.float32(0), np.float32(0.05), cuda(yy, dev), None, None, False)
    want = ref.expand(yy, np.float32(0), np.float32(0.05), lambda y: ref.gaussian_diag_drift(y, mean, inv_var),
                      ref.hadamard_points(d), ref.hadamard_weights(d), SQRT2)[:n]
    assert solver.last_path.startswith("generic") and np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("name", ["StroudSecrest63_31", "StroudSecrest63_32", "StroudSecrest63_52", "StroudSecrest63_53"])
def test_stroud_secrest_formulae_steps(dev, name):
    """Every Gaussian formula of the reference (tests/test_solvers.py:67 parametrises over them) through
    the table-cubature path of the fused kernel: unweighted step bit-exact vs the oracle, weighted step
    (non-uniform lambda for 5-2 / 5-3, quirk Q3) with weights summing to one."""
    n, d = 24, 3
    yy, mean, inv_var = _setup(n, d, seed=11)
    formula = getattr(mccube, name)(mccube.GaussianRegion(d))
    k = formula.point_count
    M, lam = formula.stacked_points, formula.stacked_weights.astype(np.float64)
    drift = lambda y: ref.gaussian_diag_drift(y, mean, inv_var)  # noqa: E731
    f = mccube.GaussianDiagDrift(mean, 1.0 / inv_var.astype(np.float64), device=dev)
    f.inv_var, f.mean = cuda(inv_var, dev), cuda(mean, dev)
    terms = mccube.MCCTerm(dfx.ODETerm(f), dfx.WeaklyDiagonalControlTerm(
        lambda t, y, args: math.sqrt(2.0), mccube.LocalLinearCubaturePath(formula)))
    t0, t1 = np.float32(0.2), np.float32(0.25)
    want = ref.mcc_step(yy, t0, t1, drift, M, lam, SQRT2, ref.MonteCarloKernel(n, key=jr.prng_key(3)))
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=3))
    got, *_ = solver.step(terms, t0, t1, cuda(yy, dev), None, None, False)
    assert solver.last_path.startswith("fused") and got.shape == (n, d)
    assert np.array_equal(got.cpu().numpy(), want), (name, k)

    w0 = np.random.default_rng(4).uniform(0.5, 1.5, size=n).astype(np.float32)
    yw = ref.pack_particles(yy, w0)
    want_w = ref.mcc_step(yw, t0, t1, drift, M, lam, SQRT2, ref.MonteCarloKernel(n, key=jr.prng_key(3)), weighted=True)
    solver_w = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=3), weighted=True)
    got_w, *_ = solver_w.step(terms, t0, t1, cuda(yw, dev), None, None, False)
    assert solver_w.last_path == "fused:mc_indices+expand_select(weighted)"
    got_w = got_w.cpu().numpy()
    assert got_w.shape == (n, d + 1) and abs(got_w[:, -1].sum() - 1.0) < 1e-5
    assert np.array_equal(got_w[:, :-1], want_w[:, :-1])
    np.testing.assert_allclose(got_w[:, -1], want_w[:, -1], rtol=1e-5, atol=1e-8)
    # the generic (materialising) path gives the same weighted step
    solver_g = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=3), weighted=True, fused=False)
    got_g, *_ = solver_g.step(terms, t0, t1, cuda(yw, dev), None, None, False)
    assert solver_g.last_path.startswith("generic")
    assert np.array_equal(got_g.cpu().numpy()[:, :-1], got_w[:, :-1])
    np.testing.assert_allclose(got_g.cpu().numpy()[:, -1], got_w[:, -1], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("n,d,s,replace", [(16, 3, 2, False), (40, 3, 3, True), (9, 8, 2, False), (5, 4, 4, False),
                                            (64, 10, 2, True)])
def test_fused_substeps_step(dev, n, d, s, replace):
    """n_substeps > 1 fused into one kernel (only the selected children of the k^s expansion are
    computed, drift re-evaluated at every intermediate state) == oracle == materialising path."""
    yy, mean, inv_var = _setup(n, d, seed=20 + s)
    drift = lambda y: ref.gaussian_diag_drift(y, mean, inv_var)  # noqa: E731
    M, lam = ref.hadamard_points(d), ref.hadamard_weights(d)
    t0, t1 = np.float32(0.3), np.float32(0.35)
    want = ref.mcc_step(yy, t0, t1, drift, M, lam, SQRT2, ref.MonteCarloKernel(n, replace, key=jr.prng_key(7)),
                        n_substeps=s)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, replace, key=7), n_substeps=s)
    got, *_ = solver.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(yy, dev), None, None, False)
    assert solver.last_path == f"fused:mc_indices+expand_select(substeps={s})"
    assert np.array_equal(got.cpu().numpy(), want)
    generic = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, replace, key=7), n_substeps=s, fused=False)
    got_g, *_ = generic.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(yy, dev), None, None, False)
    assert generic.last_path.startswith("generic") and torch.equal(got, got_g)


def test_substeps_and_weighted_step(dev):
    """n_substeps=2 and the weighted state layout (quirk Q3) through the generic path
    (/root/reference/tests/test_solvers.py:93-125)."""
    n, d = 16, 3
    yy, mean, inv_var = _setup(n, d, seed=9)
    drift = lambda y: ref.gaussian_diag_drift(y, mean, inv_var)  # noqa: E731
    M, lam = ref.hadamard_points(d), ref.hadamard_weights(d)
    t0, t1 = np.float32(0.1), np.float32(0.15)
    want = ref.mcc_step(yy, t0, t1, drift, M, lam, SQRT2, ref.MonteCarloKernel(n, key=jr.prng_key(7)), n_substeps=2)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=7), n_substeps=2)
    got, *_ = solver.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(yy, dev), None, None, False)
    assert np.array_equal(got.cpu().numpy(), want)

    w0 = np.random.default_rng(2).uniform(0.5, 1.5, size=n).astype(np.float32)
    yw = ref.pack_particles(yy, w0)
    want_w = ref.mcc_step(yw, t0, t1, drift, M, lam, SQRT2, ref.MonteCarloKernel(n, key=jr.prng_key(7)), n_substeps=2,
                          weighted=True)
    solver_w = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=7), n_substeps=2, weighted=True)
    got_w, *_ = solver_w.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(yw, dev), None, None, False)
    got_w = got_w.cpu().numpy()
    assert got_w.shape == (n, d + 1) and abs(got_w[:, -1].sum() - 1.0) < 1e-5
    assert np.array_equal(got_w[:, :-1], want_w[:, :-1])
    np.testing.assert_allclose(got_w[:, -1], want_w[:, -1], rtol=1e-5)


# ------------------------------------------------------------------ A9/A10: partitioning
def test_stratified_reference_expectations(dev):
    # /root/reference/tests/test_kernels.py:95-133
    y0 = torch.tensor([[-4.0, -3.0, -2.0, -1.0, 1.0, 2.0, 3.0, 4.0]], device=dev).T
    exp = torch.tensor([[[-1.0], [1.0]], [[-2.0], [2.0]], [[-3.0], [3.0]], [[-4.0], [4.0]]], device=dev)
    k = mccube.StratifiedPartitioningKernel(4)
    v = k(0.0, y0, ...)
    assert v.shape == (4, 2, 1) and torch.equal(v, exp)
    y3 = y0.repeat(1, 3)
    assert torch.equal(k(0.0, y3, ...), exp.repeat(1, 1, 3))
    yw = y3.clone(); yw[:, -1] = yw[:, -1].abs()
    ew = exp.repeat(1, 1, 3).clone(); ew[:, :, -1] = ew[:, :, -1].abs()
    assert torch.equal(k(0.0, yw, ..., weighted=True), ew)
    assert torch.equal(mccube.StratifiedPartitioningKernel(4, norm=None)(0.0, yw, ...), yw.reshape(4, -1, 3))


def test_partitioners_vs_oracle(dev):
    rs = np.random.default_rng(11)
    p = rs.normal(size=(4096, 7)).astype(np.float32)
    pt = cuda(p, dev)
    # stratified
    want_idx, _, _ = ref.stratified_indices(p, 64)
    got_idx = u32(mccube.StratifiedPartitioningKernel(64).indices(pt, 64))
    assert np.array_equal(got_idx, want_idx)
    # box, bounding-box bounds and fixed bounds
    for kw in (dict(dims=(0, 3, 6), bits=3), dict(bits=2), dict(dims=(1,), bits=8),
               dict(dims=(0, 1), bits=4, bounds=([-2.0, -2.0], [2.0, 2.0]))):
        want = ref.BoxPartitionKernel(64, **kw).indices(p)
        got = u32(mccube.BoxPartitionKernel(64, **kw).indices(pt, 64))
        assert np.array_equal(got, want), kw
    out = mccube.BoxPartitionKernel(64, bits=2)(0.0, pt, None)
    assert out.shape == (64, 64, 7)
    assert np.array_equal(out.cpu().numpy(), ref.BoxPartitionKernel(64, bits=2)(0.0, p))
    # binary tree: same sklearn host callback as the reference (tests/test_kernels.py:136-154)
    for mode in ("KDTree", "BallTree"):
        v = mccube.BinaryTreePartitioningKernel({"y0": 3, "y1": 4}, mode)(0.0, {"y0": pt[:6, :2].contiguous(), "y1": pt[:4, :2].contiguous()}, None)
        assert v["y0"].shape == (3, 2, 2) and v["y1"].shape == (4, 1, 2)
        assert np.array_equal(v["y0"].cpu().numpy(), ref.BinaryTreePartitioningKernel(3, mode)(0.0, p[:6, :2]))
    # Monte Carlo partitioning (tests/test_kernels.py:69-81): shape + multiset preserved
    y0 = torch.tensor([[1.0, 0.01], [2.0, 1.0], [3.0, 100.0], [4.0, 10000.0]], device=dev)
    v = mccube.MonteCarloPartitioningKernel(4, mccube.MonteCarloKernel(None, key=42))(0.0, y0, ...)
    assert v.shape == (4, 1, 2)
    assert torch.equal(torch.sort(v.reshape(-1, 2), dim=0)[0], torch.sort(y0, dim=0)[0])
    want = ref.MonteCarloPartitioningKernel(4, ref.MonteCarloKernel(None, key=jr.prng_key(42)))(0.0, y0.cpu().numpy())
    assert np.array_equal(v.cpu().numpy(), want)


def test_partitioned_step_golden(dev):
    """PartitioningRecombinationKernel(Box | Stratified, MonteCarloKernel) through
    MCCSolver.step vs the committed oracle fixture (shared local indices, quirk Q4)."""
    z = np.load(GOLDEN / "partition_step.npz")
    y0, mean, inv_var = z["y0"], z["mean"], z["inv_var"]
    n, d = y0.shape
    terms = _terms(d, mean, inv_var, dev)
    y0t = cuda(y0, dev)
    k = ref.hadamard_point_count(d)
    # children and their partition indices
    solver0 = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=42))
    drift, cub, dt, kk, dd = solver0._describe(terms, y0t, *solver0.substep_times(z["t0"], z["t1"])[0], None)
    children = _lib.expand_all(y0t, d, False, drift, dt, cub)
    assert np.array_equal(children.cpu().numpy(), z["children"])
    box = mccube.BoxPartitionKernel(32, dims=(0, 2, 5), bits=2)
    assert np.array_equal(u32(box.indices(children, 32)), z["box_idx"])
    assert np.array_equal(u32(mccube.StratifiedPartitioningKernel(32).indices(children, 32)), z["strat_idx"])
    # implicit-children bounds / keys == materialised ones
    b_rows = _lib.box_bounds(children, (0, 2, 5), 2).cpu().numpy()
    b_child = _lib.box_child_bounds(y0t, d, drift, dt, cub, (0, 2, 5), 2).cpu().numpy()
    assert np.array_equal(b_rows, b_child)
    k_rows = u32(_lib.box_keys(children, (0, 2, 5), 2, cuda(b_rows, dev)))
    k_child = u32(_lib.box_child_keys(y0t, d, drift, dt, cub, (0, 2, 5), 2, cuda(b_rows, dev)))
    assert np.array_equal(k_rows, k_child)
    # full steps
    for part, name in ((box, "box_step"), (mccube.StratifiedPartitioningKernel(32), "strat_step")):
        for use_kernel in (True, False):
            solver = mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
                part, mccube.MonteCarloKernel(n // 32, key=42)))
            solver.use_box_prk_kernel = use_kernel
            got, *_ = solver.step(terms, z["t0"], z["t1"], y0t, None, None, False)
            assert np.array_equal(got.cpu().numpy(), z[name]), (name, solver.last_path)


def test_box_step_key_column_carry(dev):
    """solver_state hand-over of the packed key columns: a trajectory with the carry is bit-identical
    to one without, the carry is used from the second step on, and it is dropped when the state
    tensor handed back is not the one the previous step produced (or was modified in place)."""
    rs = np.random.default_rng(5)
    n, d = 4096, 16
    y0 = cuda((rs.normal(size=(n, d)) * 1.5 + 1.0).astype(np.float32), dev)
    terms = _terms(d, np.full(d, 2.0, np.float32), np.full(d, 1.0 / 3.0, np.float32), dev)

    def make(carry):
        solver = mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
            mccube.BoxPartitionKernel(n // 16, dims=(4, 5, 6, 7), bits=2), mccube.MonteCarloKernel(16, key=7)))
        solver.carry_key_columns = carry
        return solver

    times = dfx.step_times(0.0, 0.05 * 6, 0.05, np.float32)
    outs = {}
    for carry in (True, False):
        solver, y, state, paths = make(carry), y0, None, []
        for (a, b) in times[:5]:
            y, _, _, state, _ = solver.step(terms, a, b, y, None, state, False)
            paths.append(solver.last_path)
        outs[carry] = y
        if carry:
            assert paths[0] == "fused:box_prk_step" and all(q == "fused:box_prk_step+keycols" for q in paths[1:]), paths
            assert torch.equal(state["mccb200_keycols"], y[:, 4:8])
            # a different tensor with the same values: no carry
            solver.step(terms, *times[5], y.clone(), None, state, False)
            assert solver.last_path == "fused:box_prk_step"
            # modified in place since: no carry
            y.add_(0.0)
            solver.step(terms, *times[5], y, None, state, False)
            assert solver.last_path == "fused:box_prk_step"
        else:
            assert all(q == "fused:box_prk_step" for q in paths), paths
    assert torch.equal(outs[True], outs[False])


def test_captured_solve_matches_eager(dev):
    """The whole README solve (config 1) replayed from one CUDA graph: bit-identical to the eager loop, for
    the captured y0 and for a different y0 fed to the same graph, on both index paths."""
    z = np.load(GOLDEN / "readme_trajectory.npz")
    n, d = 512, 10
    t0, dt0 = 0.0, 0.05
    t1 = t0 + dt0 * 512
    terms = _terms(d, (2 * np.ones(d)).astype(np.float32), (np.ones(d) / 3.0).astype(np.float32), dev)
    for replace in (False, True):
        tag = f"r{int(replace)}"
        solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, replace, key=42))
        y0 = torch.ones((n, d), device=dev)
        solve = dfx.capture_solve(terms, solver, t0, t1, dt0, y0, saveat=dfx.SaveAt(ts=[dt0, 64 * dt0], t1=True))
        sol = solve(y0)
        assert sol.stats["cuda_graph"] and sol.ys.shape[0] == 3
        assert np.array_equal(sol.ys[-1].cpu().numpy(), z[f"y_step{int(z[f'n_steps_{tag}']) - 1}_{tag}"]), tag
        other = cuda(np.random.default_rng(9).normal(size=(n, d)).astype(np.float32), dev)
        eager = dfx.diffeqsolve(terms, solver, t0, t1, dt0, other, saveat=dfx.SaveAt(ts=[dt0, 64 * dt0], t1=True))
        again = solve(other)
        assert torch.equal(again.ys, eager.ys) and np.array_equal(again.ts, eager.ts)
        assert torch.equal(solve(y0).ys, sol.ys)           # and back: replays are independent


def test_captured_solve_box_partition_and_table_formula(dev):
    """Capture covers the side-stream index draw of the box + PRK step (with the key-column carry) and a
    point-table formula (whose scaled points are uploaded once, outside the capture)."""
    rs = np.random.default_rng(5)
    n, d = 4096, 16
    y0 = cuda((rs.normal(size=(n, d)) * 1.5 + 1.0).astype(np.float32), dev)
    terms = _terms(d, np.full(d, 2.0, np.float32), np.full(d, 1.0 / 3.0, np.float32), dev)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
        mccube.BoxPartitionKernel(n // 16, dims=(4, 5, 6, 7), bits=2), mccube.MonteCarloKernel(16, key=7)))
    eager = dfx.diffeqsolve(terms, solver, 0.0, 0.4, 0.05, y0, saveat=dfx.SaveAt(steps=True))
    solve = dfx.capture_solve(terms, solver, 0.0, 0.4, 0.05, y0, saveat=dfx.SaveAt(steps=True))
    assert solver.last_path == "fused:box_prk_step+keycols"
    for _ in range(2):
        assert torch.equal(solve(y0).ys, eager.ys)

    d = 3
    f = mccube.GaussianDiagDrift(np.full(d, 2.0, np.float32), np.full(d, 3.0), device=dev)
    terms = mccube.MCCTerm(dfx.ODETerm(f), dfx.WeaklyDiagonalControlTerm(
        lambda t, y, args: math.sqrt(2.0), mccube.LocalLinearCubaturePath(mccube.StroudSecrest63_31(mccube.GaussianRegion(d)))))
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(256, key=3))
    y0 = cuda(rs.normal(size=(256, d)).astype(np.float32), dev)
    eager = dfx.diffeqsolve(terms, solver, 0.0, 1.0, 0.05, y0)
    assert torch.equal(dfx.capture_solve(terms, solver, 0.0, 1.0, 0.05, y0)(y0).ys, eager.ys)
    assert len(solver._point_tables) <= 8                   # one upload per distinct float32 step length


# ------------------------------------------------------------------ A11: metrics
def test_moments_and_w2(dev):
    rs = np.random.default_rng(5)
    # ragged d (4-byte copies), whole float4 rows (16-byte copies), one tile and several (upper triangle + mirror),
    # row counts that leave partial stages
    for n, d in [(1000, 3), (50000, 10), (20000, 64), (3000, 130), (5001, 132), (40000, 256)]:
        p = (rs.normal(size=(n, d)) * rs.uniform(0.5, 2, size=d) + rs.normal(size=d) * 3).astype(np.float32)
        w = rs.uniform(0.1, 1.0, size=n).astype(np.float32)
        pt = cuda(p, dev)
        mean, cov = mccube.mean_and_cov(pt)
        p64 = p.astype(np.float64)
        np.testing.assert_allclose(mean.cpu().numpy(), p64.mean(0), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(cov.cpu().numpy(), np.cov(p64, rowvar=False), rtol=1e-5, atol=1e-6)
        mw, cw = mccube.mean_and_cov(pt, cuda(w, dev), ddof=0)
        np.testing.assert_allclose(mw.cpu().numpy(), np.average(p64, axis=0, weights=w.astype(np.float64)), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(cw.cpu().numpy(), np.cov(p64, rowvar=False, ddof=0, aweights=w.astype(np.float64)),
                                   rtol=1e-5, atol=1e-6)
        tm, tc = np.zeros(d), np.eye(d) * 2
        w2 = mccube.gaussian_wasserstein_metric((tm, mean), (tc, cov))
        w2_ref = ref.gaussian_wasserstein_metric((tm, p64.mean(0)), (tc, np.cov(p64, rowvar=False)))
        assert abs(w2 - w2_ref) < 1e-4
    # center_of_mass (/root/reference/tests/test_metrics.py:24-34)
    y0 = torch.tensor([[1.0, 2.0], [3.0, -4.0], [5.0, 6.0]], device=dev)
    assert torch.allclose(mccube.center_of_mass(y0, None), torch.tensor([3.0, 4.0 / 3.0], device=dev))
    yu, wu = mccube.unpack_particles(y0, True)
    assert torch.allclose(mccube.center_of_mass(yu.contiguous(), wu), torch.tensor([5.0], device=dev))


def test_full_covariance_drift(dev):
    rs = np.random.default_rng(1)
    n, d = 1000, 96
    a = rs.normal(size=(d, d))
    cov = a @ a.T / d + np.eye(d)
    mean = rs.normal(size=d)
    y = rs.normal(size=(n, d)).astype(np.float32)
    f = mccube.GaussianFullDrift(mean, cov, device=dev)
    got = f(0.0, cuda(y, dev)).cpu().numpy()
    want = -(y.astype(np.float64) - mean) @ np.linalg.inv(cov)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("n,d", [(1000, 96), (128, 32), (4096, 128), (777, 512), (2048, 1024)])
def test_full_covariance_drift_tensor_cores(dev, n, d):
    """tcgen05 3xTF32 drift (config 4's dense contraction) vs float64, and vs the fp32 SIMT kernel."""
    rs = np.random.default_rng(d)
    a = rs.normal(size=(d, d))
    cov = a @ a.T / d + np.eye(d)
    mean = rs.normal(size=d) * 2.0
    y = (rs.normal(size=(n, d)) * 1.5 + mean).astype(np.float32)
    f = mccube.GaussianFullDrift(mean, cov, device=dev)
    yt = cuda(y, dev)
    got = f(0.0, yt).cpu().numpy()
    assert f.last_path == "tcgen05:3xtf32"
    prec32 = f.prec.cpu().numpy().astype(np.float64)
    want = -(y.astype(np.float64) - f.mean.cpu().numpy().astype(np.float64)) @ prec32
    scale = np.abs(want).max()
    # tensor-core accumulation truncates: the bias grows with d / 8 accumulation steps (measured
    # 7.6e-6 at d = 512, 1.3e-5 at d = 1024, relative to max |F|)
    tol = 1e-5 if d <= 512 else 2e-5
    assert np.abs(got - want).max() <= tol * scale
    # the north-star bar is on the particle STATES (y + dt f + noise within 1e-5 relative)
    dt = 0.05
    y1_got, y1_want = y + np.float32(dt) * got, y.astype(np.float64) + dt * want
    assert np.abs(y1_got - y1_want).max() <= 1e-5 * np.abs(y1_want).max()
    simt = mccube.GaussianFullDrift(mean, cov, device=dev, tensor_cores=False)
    ref = simt(0.0, yt).cpu().numpy()
    assert simt.last_path == "simt:fp32"
    assert np.abs(got - ref).max() <= tol * scale
    # a second call reuses the split precision matrix and is deterministic
    assert np.array_equal(got, f(0.0, yt).cpu().numpy())


# ------------------------------------------------------------------ fused Box + PRK kernel
@pytest.mark.parametrize("n,d,dims,bits,m,replace,kind", [
    (96, 6, (0, 2, 5), 2, 32, False, 2),
    (1000, 10, (0, 1, 2, 3), 3, 125, False, 2),      # in_per_part = 256, k = 32
    (1000, 10, (9,), 8, 50, True, 1),                 # duplicates in idx_local, in_per_part = 640 (not pow2)
    (777, 3, (0, 1, 2), 4, 21, False, 0),             # k = 8 < 32 lanes, in_per_part = 296
    (4096, 64, (0, 1, 2, 3), 3, 2048, False, 2),      # config-3 shape family: k = 128, in_per_part = 256
    (3000, 64, (5, 40, 63), 4, 3000, True, 2),        # one output per partition, dims across lane slots
    (2048, 100, (0, 33, 66, 99), 3, 64, False, 1),    # CPL = 4, k = 256 (8 rounds)
    (50000, 20, (0, 1, 2, 3, 4, 5), 2, 1000, False, 2),  # 64 classes: unsupported by the fused kernel
    (50000, 20, (0, 3, 7, 12, 19), 2, 1000, False, 2),   # 32 classes per parent
    (5, 2, (0, 1), 1, 1, False, 2),                   # fewer parents than segments
])
def test_box_prk_step_matches_sort_path_and_oracle(dev, n, d, dims, bits, m, replace, kind):
    """The fused counting-sort kernel must equal child keys -> stable sort -> expand_select
    bit for bit, and (at sizes the oracle can materialise) the oracle's partitioned step."""
    y0, mean, inv_var = _setup(n, d, seed=n + d)
    y0 *= 1.7
    k = ref.hadamard_point_count(d)
    total = n * k
    assert total % m == 0
    in_per, out_per = total // m, max(1, n // m)
    t0, t1 = np.float32(0.55), np.float32(0.6)
    y0t = cuda(y0, dev)
    if kind == 2:
        fn = lambda y: ref.gaussian_diag_drift(y, mean, inv_var)  # noqa: E731
        drift = _lib.make_drift(2, mean=cuda(mean, dev), inv_var=cuda(inv_var, dev))
    elif kind == 1:
        fn = lambda y: (np.cos(y) - 0.5 * y).astype(np.float32)  # noqa: E731
        drift = _lib.make_drift(1, f=cuda(fn(y0), dev))
    else:
        fn = lambda y: np.zeros_like(y)  # noqa: E731
        drift = _lib.make_drift(0)
    (a, b), = ref.substep_times(t0, t1, 1)
    dt = np.float32(b - a)
    cub = _lib.make_cubature(k, scale=np.float32(SQRT2 * np.float32(np.sqrt(dt))))
    bounds = _lib.box_child_bounds(y0t, d, drift, dt, cub, dims, bits)
    idx_local = _lib.mc_indices(_lib.make_rng(KEY, t0), in_per, out_per, replace, dev)
    if len(dims) > 5:   # > 32 classes per parent: the solver routes this shape to the sort path
        assert not _lib.box_prk_supported(n, d, cub, dims, bits, in_per)
        return
    assert _lib.box_prk_supported(n, d, cub, dims, bits, in_per)
    got = _lib.box_prk_step(y0t, d, drift, dt, cub, dims, bits, bounds, idx_local, in_per).cpu().numpy()
    keys = _lib.box_child_keys(y0t, d, drift, dt, cub, dims, bits, bounds)
    perm = _lib.sort_pairs(keys, None, len(dims) * bits)
    want = _lib.expand_select(y0t, d, False, drift, dt, cub, idx_local, perm=perm, out_per_part=out_per,
                              in_per_part=in_per, n_out=m * out_per).cpu().numpy()
    assert got.shape == (m * out_per, d)
    assert np.array_equal(got, want)
    if total * d <= 1 << 24:
        kern = ref.PartitioningRecombinationKernel(
            ref.BoxPartitionKernel(m, dims=dims, bits=bits),
            ref.MonteCarloKernel(out_per, replace, key=jr.prng_key(42)))
        oracle = ref.mcc_step(y0, t0, t1, fn, ref.hadamard_points(d), ref.hadamard_weights(d), SQRT2, kern)
        assert np.array_equal(got, oracle)


# ------------------------------------------------------------------ sharded global resample
@pytest.mark.parametrize("replace", [False, True])
def test_sharded_monte_carlo_step_emulated_ranks(dev, replace):
    """The multi-GPU global resample, with the ranks emulated one after another on this GPU and
    the NCCL all-to-all replaced by an in-memory exchange: the concatenated result must be
    bit-identical to the single-device step on the concatenated state."""
    from mccube_b200._distributed import ShardedMonteCarloStep

    G, n_loc, d = 4, 250, 10
    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=21)
    terms = _terms(d, mean, inv_var, dev)
    t0, t1 = np.float32(0.25), np.float32(0.3)
    yt = cuda(y, dev)
    single = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, replace, key=42))
    want, *_ = single.step(terms, t0, t1, yt, None, None, False)
    sharded = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, replace, key=42)),
                                    terms, G)
    prepared = [sharded.prepare(t0, t1, yt[r * n_loc:(r + 1) * n_loc].contiguous(), r) for r in range(G)]
    outs = []
    for r in range(G):
        parts = []
        for src in range(G):
            send_buf, send_splits, _, _ = prepared[src]
            off = sum(send_splits[:r])
            parts.append(send_buf[off:off + send_splits[r]])
        recv = torch.cat(parts)
        assert [p.shape[0] for p in parts] == prepared[r][2]
        outs.append(ShardedMonteCarloStep.place(recv, prepared[r][3]))
    assert torch.equal(torch.cat(outs), want)


@pytest.mark.parametrize("replace", [False, True])
def test_sharded_p2p_scatter_emulated_ranks(dev, replace):
    """Fused expand + scatter-to-peers kernel with the ranks emulated on one GPU (every "peer" shard is a
    local tensor): the shards written by all ranks together equal the single-device step bit for bit."""
    import math as _m

    G, n_loc, d = 4, 256, 12
    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=23)
    terms = _terms(d, mean, inv_var, dev)
    t0, t1 = np.float32(0.25), np.float32(0.3)
    yt = cuda(y, dev)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, replace, key=42))
    want, *_ = solver.step(terms, t0, t1, yt, None, None, False)
    shards = [torch.full((n_loc, d), float("nan"), device=dev) for _ in range(G)]
    ptrs = torch.tensor([s.data_ptr() for s in shards], dtype=torch.int64, device=dev)
    for r in range(G):
        ys = yt[r * n_loc:(r + 1) * n_loc].contiguous()
        drift, cub, dt, k, _ = solver._describe(terms, ys, *solver.substep_times(t0, t1)[0], None)
        idx = solver.recombination_kernel.indices(t0, n_tot * k, n_tot, dev)
        keys, counts = _lib.shard_owner_keys(idx, n_loc, k, G)
        order = _lib.sort_pairs(keys, None, max(1, _m.ceil(_m.log2(G))))
        _lib.expand_scatter_peers(ys, d, drift, dt, cub, idx, order, counts, r, G, n_loc, ptrs)
    assert torch.equal(torch.cat(shards), want)


@pytest.mark.parametrize("G,n_loc,d,dims,bits,m_div,fixed", [
    (4, 1024, 16, (4, 5, 6, 7), 2, 16, False),      # vectorised key columns, k = 32
    (2, 2048, 64, (0, 1, 2, 3), 3, 64, False),      # config-3 family: k = 128, 8192 children per partition
    (3, 1024, 10, (0, 3, 9), 2, 8, True),           # odd rank count, scalar key columns, fixed bounds
])
def test_sharded_box_step_emulated_ranks(dev, G, n_loc, d, dims, bits, m_div, fixed):
    """Global Box + PRK partition over row-sharded particles (ranks emulated on one GPU, "peer" shards are local
    tensors, the two collectives done by hand): the shards written by all ranks together equal the single-device
    step on the concatenated state bit for bit, over several steps."""
    from mccube_b200._distributed import ShardedBoxStep

    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=31)
    terms = _terms(d, mean, inv_var, dev)
    kw = dict(bounds=(np.full(len(dims), -8.0), np.full(len(dims), 12.0))) if fixed else {}
    out_per = m_div
    solver = mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
        mccube.BoxPartitionKernel(n_tot // out_per, dims=dims, bits=bits, **kw), mccube.MonteCarloKernel(out_per, key=11)))
    sh = ShardedBoxStep(solver, terms, G)
    times = dfx.step_times(0.0, 0.15, 0.05, np.float32)
    want = cuda(y, dev)
    cur = [want[r * n_loc:(r + 1) * n_loc].contiguous() for r in range(G)]
    for (a, b) in times:
        want, *_ = solver.step(terms, a, b, want, None, None, False)
        assert solver.last_path.startswith("fused:box_prk_step")
        nxt = [torch.full((n_loc, d), float("nan"), device=dev) for _ in range(G)]
        ptrs = torch.tensor([s_.data_ptr() for s_ in nxt], dtype=torch.int64, device=dev)
        mms = [sh.local_minmax(a, b, cur[r]) for r in range(G)]
        if mms[0] is None:
            mm = None
        else:
            nd = len(dims)
            mm = torch.cat([torch.stack([q[:nd] for q in mms]).min(0).values, torch.stack([q[nd:] for q in mms]).max(0).values])
        bounds = sh.bounds_from(a, b, cur[0], mm)
        totals_all = torch.stack([sh.count(a, b, cur[r], bounds)[1].clone() for r in range(G)])
        assert int(totals_all.sum()) == n_tot * solver._describe(terms, cur[0], a, b, None)[3]
        for r in range(G):
            ws, _ = sh.count(a, b, cur[r], bounds)          # (one device, one scratch buffer: count again per rank)
            sh.rank_and_scatter(a, b, cur[r], r, ws, sh.key_starts_for(r, totals_all), ptrs)
        assert torch.equal(torch.cat(nxt), want), (a, b)
        cur = nxt


def test_sharded_pull_gather_emulated_ranks(dev):
    """Pull form (with_replacement=True): every rank draws only its own slice of the index vector and reads
    the parents from the owners' shards; ranks emulated on one GPU.  Also checks the index slices."""
    G, n_loc, d = 4, 192, 12
    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=29)
    terms = _terms(d, mean, inv_var, dev)
    t0, t1 = np.float32(0.25), np.float32(0.3)
    yt = cuda(y, dev)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42))
    want, *_ = solver.step(terms, t0, t1, yt, None, None, False)
    kern = solver.recombination_kernel
    shards_in = [yt[r * n_loc:(r + 1) * n_loc].contiguous() for r in range(G)]
    ptrs = torch.tensor([s.data_ptr() for s in shards_in], dtype=torch.int64, device=dev)
    drift, cub, dt, k, _ = solver._describe(terms, shards_in[0], *solver.substep_times(t0, t1)[0], None)
    full = kern.indices(t0, n_tot * k, n_tot, dev)
    outs = []
    for r in range(G):
        idx = kern.indices_slice(t0, n_tot * k, n_tot, r * n_loc, n_loc, dev)
        assert torch.equal(idx, full[r * n_loc:(r + 1) * n_loc])
        out = torch.empty((n_loc, d), device=dev)
        outs.append(_lib.expand_gather_peers(ptrs, G, n_loc, d, drift, dt, cub, idx, out))
    assert torch.equal(torch.cat(outs), want)
    with pytest.raises(ValueError):
        mccube.MonteCarloKernel(n_tot, False, key=42).indices_slice(t0, n_tot * k, n_tot, 0, n_loc, dev)


@pytest.mark.parametrize("n_inputs,count", [(1 << 16, 1 << 11), (100003, 25000), (1 << 20, 1 << 15), (3_000_017, 9), (1 << 22, 1 << 20)])
@pytest.mark.parametrize("partitionable", [False, True])
def test_shuffle_by_selection_equals_sort_shuffle(dev, monkeypatch, n_inputs, count, partitionable):
    """`permutation[:count]` by rank selection (rng_select.cu) against the full multi-round radix-sort shuffle, and
    against the oracle's restatement of jax.random._shuffle where that is quick."""
    rng = _lib.make_rng((0, 42), np.float32(0.35), partitionable=partitionable)
    monkeypatch.setenv("MCCB200_SHUFFLE_SELECT", "1")
    got = _lib.mc_indices(rng, n_inputs, count, False, dev)
    monkeypatch.setenv("MCCB200_SHUFFLE_SELECT", "0")
    want = _lib.mc_indices(rng, n_inputs, count, False, dev)
    assert torch.equal(got, want)
    assert torch.unique(got).numel() == count
    if n_inputs <= (1 << 20):
        key = jr.mc_step_key(jr.prng_key(42), np.float32(0.35), partitionable=partitionable)
        ref = jr.choice_indices(key, n_inputs, count, False, partitionable=partitionable)
        assert np.array_equal(got.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, ref)


@pytest.mark.parametrize("n,d,n_out,partitionable", [(1000, 16, 1000, False), (777, 64, 2001, False), (4096, 128, 4096, True),
                                                    (513, 32, 100, True), (65536, 64, 65536, False), (1031, 64, 1031, False)])
def test_expand_select_draw_equals_indices_then_expand(dev, n, d, n_out, partitionable):
    """with_replacement=True: the gather kernel that draws its own child ids (no index vector in HBM) against the
    two-kernel path (mc_indices, expand_select) and the oracle's randint restatement: same ids, same rows, bit for bit.
    (65536 x 64 x k=128 gives the power-of-two span whose high draw drops out; 1031 a span that needs both draws.)"""
    y, mean, inv_var = _setup(n, d, seed=5)
    terms = _terms(d, mean, inv_var, dev)
    yt = cuda(y, dev)
    t0, t1 = np.float32(0.25), np.float32(0.3)
    kern = mccube.MonteCarloKernel(n_out, True, key=42)
    kern.threefry_partitionable = partitionable
    solver = mccube.MCCSolver(dfx.Euler(), kern)
    drift, cub, dt, k, _ = solver._describe(terms, yt, *solver.substep_times(t0, t1)[0], None)
    idx = kern.indices(t0, n * k, n_out, dev)
    want = _lib.expand_select(yt, d, False, drift, dt, cub, idx)
    rng = _lib.make_rng(kern.key, t0, partitionable=partitionable)
    idx_out = torch.empty(n_out, dtype=torch.int32, device=dev)
    got = _lib.expand_select_draw(yt, d, drift, dt, cub, rng, n_out, idx_out=idx_out)
    assert torch.equal(idx_out, idx)
    assert torch.equal(got, want)
    key = jr.mc_step_key(jr.prng_key(42), t0, partitionable=partitionable)
    ref = jr.choice_indices(key, n * k, n_out, True, partitionable=partitionable)
    assert np.array_equal(idx_out.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, ref)
    # the solver takes the fused path on its own, and says so
    y1, *_ = solver.step(terms, t0, t1, yt, None, None, False)
    assert solver.last_path == "fused:expand_select_draw"
    assert torch.equal(y1, want)
    solver.fuse_draw = False
    y1b, *_ = solver.step(terms, t0, t1, yt, None, None, False)
    assert solver.last_path.startswith("fused:mc_indices+expand_select") and torch.equal(y1b, want)
    assert not _lib.expand_select_draw_supported(12, 12, False) and not _lib.expand_select_draw_supported(64, 65, True)


@pytest.mark.parametrize("n_inputs,count", [(1 << 18, 1 << 12), (1 << 22, 1 << 19)])
def test_shuffle_by_selection_rerun_when_bucket_limit_too_small(dev, monkeypatch, n_inputs, count):
    """The round processed first looks only at the leading buckets (they hold the `count` smallest keys); when that
    limit is too small a device flag re-runs the round over all buckets.  Forced here with a limit of half the need."""
    rng = _lib.make_rng((0, 7), np.float32(0.6))
    monkeypatch.setenv("MCCB200_SHUFFLE_SELECT", "0")
    want = _lib.mc_indices(rng, n_inputs, count, False, dev)
    monkeypatch.setenv("MCCB200_SHUFFLE_SELECT", "1")
    monkeypatch.setenv("MCCB200_SELECT_SHORT_LIMIT", "1")
    got = _lib.mc_indices(rng, n_inputs, count, False, dev)
    assert torch.equal(got, want)


@pytest.mark.parametrize("G,n_loc,d,x64", [(4, 250, 12, False), (3, 333, 8, True), (8, 64, 64, True), (1, 100, 4, False)])
def test_sharded_owner_ids_emulated_ranks(dev, G, n_loc, d, x64):
    """Id exchange + owner-computes (`step_owner_ids`): every rank draws only its own index slice, the ranks
    exchange (slot, child id) pairs and each materialises the children of its OWN parents.  Ranks emulated on one
    GPU, the two all-to-alls done by hand: row q of rank r's shard equals row slot_map[q] of the single-device
    step on the concatenated state, every output slot appears exactly once, and every shard has n_loc rows."""
    from mccube_b200._distributed import ShardedMonteCarloStep

    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=41)
    terms = _terms(d, mean, inv_var, dev)
    t0, t1 = np.float32(0.25), np.float32(0.3)
    yt = cuda(y, dev)

    def make():
        kern = mccube.MonteCarloKernel(n_tot, True, key=42)
        kern.x64 = x64
        return mccube.MCCSolver(dfx.Euler(), kern)

    want, *_ = make().step(terms, t0, t1, yt, None, None, False) if not x64 else (None,)
    if x64:   # the single-device step draws 32-bit ids; build the x64 reference from the x64 index vector
        solver = make()
        drift, cub, dt, k, _ = solver._describe(terms, yt, *solver.substep_times(t0, t1)[0], None)
        idx = solver.recombination_kernel.indices_slice(t0, n_tot * k, n_tot, 0, n_tot, dev)
        ptrs = torch.tensor([yt.data_ptr()], dtype=torch.int64, device=dev)
        want = _lib.expand_gather_peers(ptrs, 1, n_tot, d, drift, dt, cub, idx, torch.empty_like(yt))
    sh = ShardedMonteCarloStep(make(), terms, G)
    shards = [yt[r * n_loc:(r + 1) * n_loc].contiguous() for r in range(G)]
    bucketed = [sh.owner_ids_bucket(t0, n_loc, r, dev) for r in range(G)]
    packed = [sh.owner_ids_bucket(t0, n_loc, r, dev, packed=True) for r in range(G)]
    for (s_, i_, c_), (p_, c2_) in zip(bucketed, packed):     # the packed form holds the same pairs
        assert torch.equal(p_[:, 0], s_) and torch.equal(p_[:, 1], i_) and torch.equal(c_, c2_)
        s2_, i2_ = _lib.unzip_pairs(p_)
        assert torch.equal(s2_, s_) and torch.equal(i2_, i_)
    counts = torch.stack([c for _, _, c in bucketed]).cpu().numpy()        # counts[src][owner]
    assert counts.sum() == n_tot and (counts.sum(1) == n_loc).all()
    mats = []
    for r in range(G):
        pp = []
        for src in range(G):
            off = int(counts[src, :r].sum())
            pp.append(packed[src][0][off:off + int(counts[src, r])])
        mats.append(sh.owner_ids_materialise(t0, t1, shards[r], r, torch.cat(pp), counts))
    outs, maps = [], []
    for r in range(G):
        y1, slots, _, _, _, recv_splits, keep = mats[r]
        slot_map = torch.empty(n_loc, dtype=torch.int32, device=dev)
        slot_map[:keep] = slots
        at = keep
        for src in range(G):
            c = recv_splits[src]
            if c:
                _, _, send_rows, send_slots, send_splits, _, _ = mats[src]
                off = sum(send_splits[:r])
                y1[at:at + c] = send_rows[off:off + c]
                slot_map[at:at + c] = send_slots[off:off + c]
                at += c
        assert at == n_loc
        outs.append(y1)
        maps.append(slot_map)
    all_slots = torch.cat(maps).to(torch.int64)
    assert torch.equal(torch.sort(all_slots).values, torch.arange(n_tot, device=dev))
    assert torch.equal(torch.cat(outs), want[all_slots])
    # grouping is stable: ascending slots inside every owner group of the send buffer, ids local to the owner
    k = terms.cde.control.gaussian_cubature.point_count
    for r in range(G):
        slots_r, ids_r = bucketed[r][0].to(torch.int64), bucketed[r][1].to(torch.int64) & 0xFFFFFFFF
        start = 0
        for o in range(G):
            c = int(counts[r, o])
            assert bool((slots_r[start + 1:start + c] > slots_r[start:start + c - 1]).all())
            assert c == 0 or int(ids_r[start:start + c].max()) < n_loc * k
            start += c


@pytest.mark.parametrize("partitionable", [False, True])
def test_sharded_pull_x64_indices(dev, partitionable):
    """jax_enable_x64 semantics of the index draw (int64 randint, 64-bit draws) for the pull path: slices vs the
    oracle's restatement, also for a span beyond 2^32 (BASELINE config 5 has 2^34 children), and an emulated
    sharded step vs the oracle's step with the x64 kernel.  Unpinned (no JAX in the image)."""
    t0, t1 = np.float32(0.25), np.float32(0.3)
    for span, count in ((1 << 34, 1000), (12345, 777), ((1 << 32) + 17, 64)):
        want = jr.choice_indices(jr.mc_step_key(jr.prng_key(42), t0, partitionable=partitionable), span, count, True,
                                 partitionable=partitionable, x64=True)
        kern = mccube.MonteCarloKernel(count, True, key=42, threefry_partitionable=partitionable, x64=True)
        got = torch.cat([kern.indices_slice(t0, span, count, a, b - a, dev) for a, b in ((0, 10), (10, count))])
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)
    G, n_loc, d = 2, 96, 8
    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=31)
    terms = _terms(d, mean, inv_var, dev)
    yt = cuda(y, dev)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42, threefry_partitionable=partitionable,
                                                                    x64=True))
    want = ref.mcc_step(y, t0, t1, lambda q: ref.gaussian_diag_drift(q, mean, inv_var), ref.hadamard_points(d),
                        ref.hadamard_weights(d), SQRT2,
                        ref.MonteCarloKernel(n_tot, True, key=jr.prng_key(42), partitionable=partitionable, x64=True))
    shards_in = [yt[r * n_loc:(r + 1) * n_loc].contiguous() for r in range(G)]
    ptrs = torch.tensor([s.data_ptr() for s in shards_in], dtype=torch.int64, device=dev)
    drift, cub, dt, k, _ = solver._describe(terms, shards_in[0], *solver.substep_times(t0, t1)[0], None)
    outs = []
    for r in range(G):
        idx = solver.recombination_kernel.indices_slice(t0, n_tot * k, n_tot, r * n_loc, n_loc, dev)
        outs.append(_lib.expand_gather_peers(ptrs, G, n_loc, d, drift, dt, cub, idx, torch.empty((n_loc, d), device=dev)))
    assert np.array_equal(torch.cat(outs).cpu().numpy(), want)


def test_x64_monte_carlo_step_single_device(dev):
    """MonteCarloKernel(x64=True, with_replacement=True) on one device: the fused step equals the oracle's step
    with the int64 index draw (and differs from the int32 draw)."""
    n, d = 300, 10
    y, mean, inv_var = _setup(n, d, seed=33)
    t0, t1 = np.float32(0.1), np.float32(0.15)
    want = ref.mcc_step(y, t0, t1, lambda q: ref.gaussian_diag_drift(q, mean, inv_var), ref.hadamard_points(d),
                        ref.hadamard_weights(d), SQRT2, ref.MonteCarloKernel(n, True, key=jr.prng_key(5), x64=True))
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, True, key=5, x64=True))
    got, *_ = solver.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(y, dev), None, None, False)
    assert solver.last_path.startswith("fused") and np.array_equal(got.cpu().numpy(), want)
    solver32 = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, True, key=5))
    got32, *_ = solver32.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(y, dev), None, None, False)
    assert not torch.equal(got, got32)


def test_sharded_owner_computes_emulated_ranks(dev):
    """Owner-computes variant: only the imbalance is exchanged; rows equal the single-device
    result after applying the exported slot map (a permutation of all output slots)."""
    from mccube_b200._distributed import ShardedMonteCarloStep

    G, n_loc, d = 4, 500, 10
    n_tot = G * n_loc
    y, mean, inv_var = _setup(n_tot, d, seed=33)
    terms = _terms(d, mean, inv_var, dev)
    t0, t1 = np.float32(0.4), np.float32(0.45)
    yt = cuda(y, dev)
    want, *_ = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42)).step(
        terms, t0, t1, yt, None, None, False)
    sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42)), terms, G)
    prep = [sh.prepare_owner(t0, t1, yt[r * n_loc:(r + 1) * n_loc].contiguous(), r) for r in range(G)]
    all_slots = []
    moved = 0
    for r in range(G):
        y1, _, _, recv_splits, keep, slot_map = prep[r]
        parts = []
        for src in range(G):
            _, send_buf, send_splits, _, _, _ = prep[src]
            off = sum(send_splits[:r])
            parts.append(send_buf[off:off + send_splits[r]])
        recv = torch.cat(parts)
        assert [p.shape[0] for p in parts] == recv_splits and keep + recv.shape[0] == n_loc
        y1[keep:] = recv
        moved += recv.shape[0]
        sm = slot_map.to(torch.int64) & 0xFFFFFFFF
        assert torch.equal(y1, want[sm])
        all_slots.append(sm)
    assert torch.equal(torch.sort(torch.cat(all_slots)).values, torch.arange(n_tot, device=dev))
    assert moved < n_tot // 4          # only the imbalance crosses ranks


# ------------------------------------------------------------------ weighted sampling (float branches)
def test_weighted_monte_carlo_reference_expectation(dev):
    # /root/reference/tests/test_kernels.py:83-92: identity weighting, key 42, t=0 -> rows by descending weight
    y0 = torch.tensor([[1.0, 0.01], [2.0, 1.0], [3.0, 100.0], [4.0, 10000.0]], device=dev)
    for part in (False, True):
        mc = mccube.MonteCarloKernel(None, weighting_function=lambda w: w, key=42, threefry_partitionable=part)
        v = mccube.MonteCarloPartitioningKernel(4, mc)(0.0, y0, ..., weighted=True)
        assert v.shape == (4, 1, 2)
        order = torch.argsort(y0[:, -1], descending=True)
        assert torch.equal(v[:, 0, :], y0[order])


@pytest.mark.parametrize("replace", [False, True])
@pytest.mark.parametrize("partitionable", [False, True])
def test_weighted_indices_vs_oracle(dev, replace, partitionable):
    """Same random bits as jax.random; float scores may round differently from NumPy's libm, so
    allow a tiny fraction of index differences (ties within one ulp)."""
    rs = np.random.default_rng(4)
    P, n, t = 20000, 2500, 0.35
    p = rs.uniform(0.05, 1.0, size=P).astype(np.float32)
    key = jr.mc_step_key(jr.prng_key(42), t, partitionable=partitionable)
    if replace:
        c = np.cumsum(p.astype(np.float64)).astype(np.float32)     # the CUDA path accumulates in fp64
        u = jr.uniform(key, n, np.float32, partitionable=partitionable)
        want = np.searchsorted(c, c[-1] * (np.float32(1) - u))
    else:
        want = jr.choice_indices(key, P, n, False, p, partitionable=partitionable)
    rng = _lib.make_rng(KEY, t, partitionable=partitionable)
    got = u32(_lib.weighted_indices(rng, cuda(p, dev), n, replace))
    assert got.shape == want.shape
    mismatch = float((got != want).mean())
    assert mismatch < 2e-3, mismatch
    if not replace:
        assert len(np.unique(got)) == n


def test_weighted_solver_step_with_weighting_function(dev):
    """MCCSolver(weighted=True) + MonteCarloKernel(weighting_function=identity): weights are
    used as sampling probabilities and renormalised (generic path)."""
    n, d = 64, 3
    yy, mean, inv_var = _setup(n, d, seed=12)
    w0 = np.random.default_rng(3).uniform(0.5, 1.5, size=n).astype(np.float32)
    yw = ref.pack_particles(yy, w0)
    t0, t1 = np.float32(0.1), np.float32(0.15)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, weighting_function=lambda w: w, key=7), weighted=True)
    got, *_ = solver.step(_terms(d, mean, inv_var, dev), t0, t1, cuda(yw, dev), None, None, False)
    got = got.cpu().numpy()
    want = ref.mcc_step(yw, t0, t1, lambda y: ref.gaussian_diag_drift(y, mean, inv_var), ref.hadamard_points(d),
                        ref.hadamard_weights(d), SQRT2,
                        ref.MonteCarloKernel(n, weighting_function=lambda w: w, key=jr.prng_key(7)), weighted=True)
    assert got.shape == want.shape and abs(got[:, -1].sum() - 1.0) < 1e-5
    same_rows = (got[:, :-1] == want[:, :-1]).all(axis=1).mean()
    assert same_rows > 0.95


# ------------------------------------------------------------------ edge cases of the newer entry points
def test_edge_cases_new_entry_points(dev):
    """Tiny / ragged / empty inputs through the tensor-core drift, the fused sub-steps, the peer gather and
    the split Box API."""
    rs = np.random.default_rng(77)
    # tensor-core drift: a single particle, and one row more than a tile
    d = 64
    a = rs.normal(size=(d, d)); cov = a @ a.T / d + np.eye(d); mu = rs.normal(size=d)
    f = mccube.GaussianFullDrift(mu, cov, device=dev)
    for n in (1, 129, 257):
        y = (rs.normal(size=(n, d)) + mu).astype(np.float32)
        got = f(0.0, cuda(y, dev)).cpu().numpy()
        want = -(y.astype(np.float64) - f.mean.cpu().numpy()) @ f.prec.cpu().numpy().astype(np.float64)
        assert f.last_path == "tcgen05:3xtf32" and np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # d not a multiple of 32: the SIMT tiles take over (no silent precision change: same tolerance)
    d2 = 40
    a = rs.normal(size=(d2, d2)); cov2 = a @ a.T / d2 + np.eye(d2)
    f2 = mccube.GaussianFullDrift(np.zeros(d2), cov2, device=dev)
    y = rs.normal(size=(50, d2)).astype(np.float32)
    got = f2(0.0, cuda(y, dev)).cpu().numpy()
    assert f2.last_path == "simt:fp32"
    assert np.abs(got + y.astype(np.float64) @ f2.prec.cpu().numpy().astype(np.float64)).max() <= 1e-5 * np.abs(got).max()
    # fused sub-steps with an odd dimension (scalar vector path) and a single particle
    for n, dd in ((1, 3), (7, 5)):
        yy, mean, inv_var = _setup(n, dd, seed=40 + n)
        want = ref.mcc_step(yy, np.float32(0.0), np.float32(0.05), lambda q: ref.gaussian_diag_drift(q, mean, inv_var),
                            ref.hadamard_points(dd), ref.hadamard_weights(dd), SQRT2,
                            ref.MonteCarloKernel(n, key=jr.prng_key(9)), n_substeps=2)
        solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=9), n_substeps=2)
        got, *_ = solver.step(_terms(dd, mean, inv_var, dev), np.float32(0.0), np.float32(0.05), cuda(yy, dev), None, None, False)
        assert "substeps=2" in solver.last_path and np.array_equal(got.cpu().numpy(), want)
    # peer gather with one "rank" and an empty slice
    n, dd = 33, 6
    yy, mean, inv_var = _setup(n, dd, seed=50)
    terms = _terms(dd, mean, inv_var, dev)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, True, key=11))
    t0, t1 = np.float32(0.0), np.float32(0.05)
    yt = cuda(yy, dev)
    want, *_ = solver.step(terms, t0, t1, yt, None, None, False)
    drift, cub, dt, k, _ = solver._describe(terms, yt, *solver.substep_times(t0, t1)[0], None)
    ptrs = torch.tensor([yt.data_ptr()], dtype=torch.int64, device=dev)
    idx = solver.recombination_kernel.indices_slice(t0, n * k, n, 0, n, dev)
    out = _lib.expand_gather_peers(ptrs, 1, n, dd, drift, dt, cub, idx, torch.empty((n, dd), device=dev))
    assert torch.equal(out, want)
    assert solver.recombination_kernel.indices_slice(t0, n * k, n, 5, 0, dev).numel() == 0
    # split Box API on a state smaller than one batch of 32 parents, non-power-of-two partition size
    n, dd = 24, 8
    yy, mean, inv_var = _setup(n, dd, seed=51)
    kern = mccube.PartitioningRecombinationKernel(mccube.BoxPartitionKernel(16, dims=(0, 1, 2, 3), bits=2),
                                                  mccube.MonteCarloKernel(3, key=13))
    solver = mccube.MCCSolver(dfx.Euler(), kern)
    got, _, _, state, _ = solver.step(_terms(dd, mean, inv_var, dev), t0, t1, cuda(yy, dev), None, None, False)
    okern = ref.PartitioningRecombinationKernel(ref.BoxPartitionKernel(16, dims=(0, 1, 2, 3), bits=2),
                                                ref.MonteCarloKernel(3, key=jr.prng_key(13)))
    want = ref.mcc_step(yy, t0, t1, lambda q: ref.gaussian_diag_drift(q, mean, inv_var), ref.hadamard_points(dd),
                        ref.hadamard_weights(dd), SQRT2, okern)
    assert solver.last_path == "fused:box_prk_step" and np.array_equal(got.cpu().numpy(), want)
    # and a second step that uses the carried key columns (48 rows now: the state grew)
    got2, *_ = solver.step(_terms(dd, mean, inv_var, dev), t1, np.float32(0.1), got, None, state, False)
    want2 = ref.mcc_step(want, t1, np.float32(0.1), lambda q: ref.gaussian_diag_drift(q, mean, inv_var),
                         ref.hadamard_points(dd), ref.hadamard_weights(dd), SQRT2, okern)
    assert solver.last_path == "fused:box_prk_step+keycols" and np.array_equal(got2.cpu().numpy(), want2)


def test_captured_solve_generic_path_and_refusal(dev):
    """The generic (materialising) path with a device-side partitioner is capturable as well; a solve whose
    step synchronises with the host (BinaryTree: sklearn callback) is refused with a ValueError and leaves
    the device usable."""
    n, d = 256, 4
    terms = _terms(d, np.full(d, 2.0, np.float32), np.full(d, 1.0 / 3.0, np.float32), dev)
    y0 = cuda(np.random.default_rng(2).normal(size=(n, d)).astype(np.float32), dev)
    strat = mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
        mccube.StratifiedPartitioningKernel(16), mccube.MonteCarloKernel(16, key=1)))
    eager = dfx.diffeqsolve(terms, strat, 0.0, 0.2, 0.05, y0).ys
    assert strat.last_path == "generic:expand_all+kernel"
    assert torch.equal(dfx.capture_solve(terms, strat, 0.0, 0.2, 0.05, y0)(y0).ys, eager)
    tree = mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
        mccube.BinaryTreePartitioningKernel(16), mccube.MonteCarloKernel(16, key=1)))
    with pytest.raises(ValueError, match="cannot be recorded as a CUDA graph"):
        dfx.capture_solve(terms, tree, 0.0, 0.2, 0.05, y0)
    plain = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=1))
    assert torch.equal(dfx.capture_solve(terms, plain, 0.0, 0.2, 0.05, y0)(y0).ys,
                       dfx.diffeqsolve(terms, plain, 0.0, 0.2, 0.05, y0).ys)


@pytest.mark.parametrize("n,d,leaf,weighted,mode", [(4096, 8, 64, False, "KDTree"), (1000, 5, 10, False, "BallTree"),
                                                   (3000, 64, 100, True, "KDTree"), (1 << 16, 16, 64, False, "KDTree"),
                                                   (777, 3, 7, False, "KDTree"), (512, 40, 512, False, "BallTree")])
def test_device_median_split_partitioner_matches_sklearn_sets(dev, n, d, leaf, weighted, mode):
    """`BinaryTreePartitioningKernel(device=True)`: the median-split tree built on the GPU holds, in every sklearn leaf's
    range of idx_array, exactly the point set sklearn put there (reference: mccube/_kernels/tree.py:44-58 -> sklearn
    _binary_tree.pxi _recursive_build); it is a permutation; and each of its own (one level deeper) leaves is split
    from its sibling at the median of the widest dimension of their parent."""
    from sklearn.neighbors import BallTree, KDTree

    rng = np.random.default_rng(3)
    y = rng.standard_normal((n, d + (1 if weighted else 0))).astype(np.float32) * rng.uniform(0.5, 3.0, d + (1 if weighted else 0)).astype(np.float32)
    yt = cuda(y, dev)
    count = n // leaf
    kern = mccube.BinaryTreePartitioningKernel(count, mode=mode, device=True)
    idx = kern.indices(yt, count, weighted).cpu().numpy().reshape(-1)
    assert np.array_equal(np.sort(idx), np.arange(n))
    tree = {"KDTree": KDTree, "BallTree": BallTree}[mode](y[:, :d].astype(np.float64), leaf)
    _, sk_idx, node_data, _ = tree.get_arrays()
    leaves = [(int(nd["idx_start"]), int(nd["idx_end"])) for nd in node_data if nd["is_leaf"]]
    assert sum(e - s for s, e in leaves) == n
    for s, e in leaves:
        assert set(idx[s:e].tolist()) == set(np.asarray(sk_idx[s:e]).tolist()), (s, e)
    # the extra level: inside every sklearn leaf with >= 2 * leaf points, the halves are separated along the widest
    # dimension of that leaf
    levels = _lib.kd_partition_levels(n, leaf)
    sk_levels = int(np.log2(max(1, (n - 1) // leaf)))
    assert levels in (sk_levels, sk_levels + 1)
    if levels == sk_levels + 1:
        for s, e in leaves:
            pts = y[idx[s:e], :d].astype(np.float64)
            j = int(np.argmax(pts.max(0) - pts.min(0)))
            mid = (e - s) // 2
            assert pts[:mid, j].max() <= pts[mid:, j].min()
    # the kernel call gathers the rows partition by partition
    out = kern(np.float32(0.0), yt, None, weighted)
    assert out.shape == (count, leaf, y.shape[1])
    assert torch.equal(out.reshape(n, -1), yt[torch.from_numpy(idx.astype(np.int64)).to(dev)])


@pytest.mark.parametrize("n,ld", [(1000, 65), (4097, 5), (333, 2), (2048, 4), (10, 3), (70000, 17)])
def test_normalised_copy_equals_clone_then_normalise(dev, n, ld):
    """MCCSolver.step's renormalised copy (mccube/_solvers.py:146-155) in one pass: same bits as clone + normalise_weights_,
    with the sum computed here or handed over by the producer."""
    rng = np.random.default_rng(11)
    y = rng.standard_normal((n, ld)).astype(np.float32)
    y[:, -1] = rng.uniform(0.1, 2.0, n).astype(np.float32)
    yt = cuda(y, dev)
    want = _lib.normalise_weights_(yt.clone())
    got = _lib.normalised_copy(yt)
    assert torch.equal(yt, cuda(y, dev))                       # the input keeps its raw weights
    total = yt[:, -1].double().sum().reshape(1)
    got2 = _lib.normalised_copy(yt, total)
    s32 = np.float32(total.item())
    assert torch.equal(got[:, :-1], yt[:, :-1]) and torch.equal(got2[:, :-1], yt[:, :-1])
    assert torch.equal(got2[:, -1], yt[:, -1] / torch.tensor(s32, device=dev))
    # the in-kernel sum is a double accumulation in another order: equal after the cast to float32 except at a rounding tie
    assert torch.allclose(got[:, -1], want[:, -1], rtol=2e-7, atol=0) and abs(float(got[:, -1].double().sum()) - 1.0) < 1e-5


def test_weighted_step_sums_child_weights_in_the_gather(dev):
    """Weighted fused MC step: the gather kernel hands the sum of the child weights to the renormalising copy; the
    result equals normalising dense_info's raw y1 separately."""
    n, d = 3000, 16
    y, mean, inv_var = _setup(n, d, seed=9)
    w = np.random.default_rng(2).uniform(0.5, 1.5, (n, 1)).astype(np.float32)
    yw = cuda(np.concatenate([y, w / w.sum()], axis=1), dev)
    terms = _terms(d, mean, inv_var, dev)
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, key=3), weighted=True)
    y1, _, dense, _, _ = solver.step(terms, np.float32(0.0), np.float32(0.05), yw, None, None, False)
    assert solver.last_path == "fused:mc_indices+expand_select(weighted)"
    raw = dense["y1"]
    ref = _lib.normalise_weights_(raw.clone())
    assert torch.equal(y1[:, :-1], raw[:, :-1])
    assert torch.allclose(y1[:, -1], ref[:, -1], rtol=2e-7, atol=0)
    assert abs(float(y1[:, -1].double().sum()) - 1.0) < 1e-5


@pytest.mark.parametrize("n,d", [(4096, 64), (1000, 32), (129, 96), (70001, 128), (5, 64)])
def test_stratified_keys_async_tiles_equal_oracle_and_staged_kernel(dev, monkeypatch, n, d):
    """Rows made of whole 16-byte chunks take the cp.async / swizzled-tile kernel: the same keys, bit for bit, as the
    oracle's sequential fp32 sum (mccube/_kernels/stratified.py:36-45) and as the register-staged kernel."""
    p = np.random.default_rng(n + d).normal(size=(n, d)).astype(np.float32) * 3
    pt = cuda(p, dev)
    _, com, key = ref.stratified_indices(p, 1)
    comt = cuda(com.astype(np.float32), dev)
    got = _lib.stratified_keys(pt, d, comt)
    monkeypatch.setenv("MCCB200_STRATIFIED_ASYNC", "0")
    staged = _lib.stratified_keys(pt, d, comt)
    assert torch.equal(got, staged)
    assert np.array_equal(u32(got), key.view(np.uint32))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 16385, 300001])
def test_second_moment_tcgen05_d64(dev, monkeypatch, n):
    """d = 64 unweighted second moments run on tcgen05 (MN-major TF32 operands, one M128 x N64 x K8 instruction per 8
    rows for the three 3xTF32 products, csrc/moments_tc.cu): against fp64 numpy within 1e-5 of the largest entry (the
    bar of jnp.cov in fp32, README.md:85-86), symmetric, and on the tensor-core path (phase mark)."""
    rs = np.random.default_rng(n)
    p = (rs.normal(size=(n, 64)) * rs.uniform(0.5, 2, size=64) + rs.normal(size=64) * 3).astype(np.float32)
    c = p.astype(np.float64).mean(0).astype(np.float32)
    pt, ct = cuda(p, dev), cuda(c, dev)
    _lib.profile_enable(True)
    got = _lib.second_moment(pt, 64, ct)
    torch.cuda.synchronize()
    names = [name for name, _ in _lib.profile_read()]
    _lib.profile_enable(False)
    assert "moments_m2.tcgen05" in names
    x = p.astype(np.float64) - c.astype(np.float64)
    want = x.T @ x
    g = got.cpu().numpy()
    assert np.abs(g - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)
    assert np.abs(g - g.T).max() <= 1e-6 * max(np.abs(want).max(), 1e-30)
    monkeypatch.setenv("MCCB200_M2_TC", "2")                        # second form: A = [hi ; lo] in tensor memory, B = hi in smem
    ts = _lib.second_moment(pt, 64, ct).cpu().numpy()
    assert np.abs(ts - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)
    monkeypatch.setenv("MCCB200_M2_TC", "0")                        # the mma.sync kernel stays as the general form
    old = _lib.second_moment(pt, 64, ct).cpu().numpy()
    assert np.abs(old - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)


@pytest.mark.parametrize("n,d,leaf", [(4096, 8, 64), (1000, 5, 10), (777, 3, 7), (3000, 64, 100), (1 << 15, 130, 32), (100, 2, 1)])
def test_device_median_split_equals_the_oracle_exactly(dev, n, d, leaf):
    """Index parity of csrc/kd_split.cu: the device tree equals the oracle's restatement of sklearn's split rule
    (`oracle.mccube_ref.median_split_indices`, itself checked against sklearn's node sets on the CPU) position by
    position -- same split dimensions, same medians, same stable order inside the leaves."""
    rng = np.random.default_rng(n + d)
    y = (rng.standard_normal((n, d)) * rng.uniform(0.5, 3.0, d)).astype(np.float32)
    y[rng.integers(0, n, 5), rng.integers(0, d, 5)] = 0.0           # a few exact zeros and duplicates
    y[1] = y[0]
    levels = _lib.kd_partition_levels(n, leaf)
    assert levels == ref.median_split_levels(n, leaf)
    got = _lib.kd_partition(cuda(y, dev), d, levels).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, ref.median_split_indices(y, levels))
