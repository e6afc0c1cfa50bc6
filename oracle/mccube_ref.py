"""NumPy restatement of the MCCube hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to
`/root/reference`).  The expansion is *materialised* exactly as the reference does
(`mccube/_solvers.py:136`), so this is only usable at sizes where `n*k*d` fits in RAM.

Floating-point policy: every elementwise op is a separately rounded NumPy op in the
array dtype (no FMA contraction), in the reference's op order:
    y1 = y0 + (dt * f(y0) + g * (coeff * M))         (_term.py:70-77, diffrax Euler.step)
The CUDA kernels use `__fmul_rn/__fadd_rn` in the same order, so CUDA == oracle bit
for bit on the fp32 state for the elementwise drifts.
"""
from __future__ import annotations

import numpy as np

from . import jax_random as jr


# --------------------------------------------------------------------------- formulae
def hadamard_point_count(d: int) -> int:
    """`Hadamard.point_count` (mccube/_formulae.py:223-227): k = 2^(ceil(log2 d)+1)."""
    return int(2 ** (np.ceil(np.log2(d)) + 1))


def hadamard_points(d: int) -> np.ndarray:
    """`Hadamard.points` (mccube/_formulae.py:217-221): vstack(H[:, :d], -H[:, :d]),
    H = Sylvester Hadamard matrix of order k/2: H[j, c] = (-1)^popcount(j & c)."""
    k = hadamard_point_count(d)
    half = k // 2
    j = np.arange(half, dtype=np.int64)[:, None]
    c = np.arange(d, dtype=np.int64)[None, :]
    par = np.zeros((half, d), dtype=np.int64)
    v = j & c
    while np.any(v):
        par ^= v & 1
        v >>= 1
    H = 1 - 2 * par
    return np.vstack((H, -H))


def hadamard_weights(d: int) -> np.ndarray:
    """`stacked_weights` (mccube/_formulae.py:94-98, 213-215): volume / k each."""
    k = hadamard_point_count(d)
    return np.full(k, 1.0 / k)


# --------------------------------------------------------------------------- path
def path_evaluate(points: np.ndarray, t0, t1=None, dtype=np.float32) -> np.ndarray:
    """`LocalLinearCubaturePath.evaluate` (mccube/_path.py:60-71)."""
    dtype = np.dtype(dtype)
    if t1 is None:
        t0, t1 = dtype.type(0.0), t0
    t0, t1 = dtype.type(t0), dtype.type(t1)
    dt = dtype.type(t1 - t0)
    coeff = dtype.type(0.0) if t0 == t1 else dtype.type(np.sqrt(dt))
    return (coeff * points.astype(dtype)).astype(dtype)


# --------------------------------------------------------------------------- utils
def pack_particles(p, w):
    """`pack_particles` (mccube/_utils.py:47-51)."""
    if w is None:
        return p
    w = np.asarray(w)
    return np.c_[p, (np.squeeze(w) / np.sum(w)).astype(p.dtype)]


def unpack_particles(p, weighted=False):
    """`unpack_particles` (mccube/_utils.py:82-84)."""
    if weighted:
        return p[..., :-1], p[..., -1]
    return p, None


# --------------------------------------------------------------------------- drifts
def gaussian_diag_drift(y, mean, inv_var):
    """grad log N(y; mean, diag(1/inv_var)) = (mean - y) * inv_var.

    Stands in for the user callable `vmap(grad(logpdf))` (README.md:72-76).  The
    reference evaluates it through two triangular solves against chol(cov); for
    diagonal cov that is the same value up to ~1 ulp (covered by the 1e-5 gate)."""
    dt = y.dtype
    return ((mean.astype(dt) - y) * inv_var.astype(dt)).astype(dt)


def gaussian_full_drift(y, mean, prec):
    """-(y - mean) @ prec for a full precision matrix (config 4)."""
    dt = y.dtype
    return (-((y - mean.astype(dt)).astype(np.float64) @ prec.astype(np.float64))).astype(dt)


# --------------------------------------------------------------------------- step pieces
def substep_times(t0, t1, n_substeps, dtype=np.float32):
    """(_t0, _t1) of every inner step of `MCCSolver.step` (mccube/_solvers.py:127-134).

    Reproduces quirk Q1: `_t0 = _t1` executes before the first inner step, so inner step
    i integrates over [t0 + (i+1)h, t0 + (i+2)h] in the time dtype."""
    dtype = np.dtype(dtype).type
    t0, t1 = dtype(t0), dtype(t1)
    h = dtype(dtype(t1 - t0) / dtype(n_substeps))
    _t0 = t0
    _t1 = dtype(t0 + h)
    out = []
    for _ in range(n_substeps):
        _t0 = _t1
        _t1 = dtype(_t0 + h)
        out.append((_t0, _t1))
    return out


def euler_expand(y, drift, points, g, t0s, t1s):
    """One inner `Euler.step` through `MCCTerm` (mccube/_term.py:58-77; diffrax Euler:
    y1 = y0 + vf_prod): [n, d] -> [n*k, d], child c = i*k + j (mccube/_solvers.py:136)."""
    dt_ = y.dtype.type
    n, d = y.shape
    dt = dt_(t1s - t0s)
    ode_prod = (dt * drift(y)).astype(y.dtype)                                   # ODETerm.prod
    cde_prod = (dt_(g) * path_evaluate(points, t0s, t1s, y.dtype)).astype(y.dtype)  # WDCT.prod
    vf_prod = (ode_prod[:, None, :] + cde_prod[None, :, :]).astype(y.dtype)      # _term.py:77
    y1 = (y[:, None, :] + vf_prod).astype(y.dtype)                               # Euler.step
    return y1.reshape(-1, d)


def expand(y0, t0, t1, drift, points, lam, g, n_substeps=1, weighted=False, time_dtype=None):
    """Expansion half of `MCCSolver.step` (mccube/_solvers.py:125-143): returns the packed
    `[n*k^s, d(+1)]` children.  Reproduces quirk Q3 (weights laid out k-major)."""
    y, w = unpack_particles(y0, weighted)
    td = time_dtype or y.dtype
    for (a, b) in substep_times(t0, t1, n_substeps, td):
        y = euler_expand(y, drift, points, g, a, b)
        if w is not None:
            w = (w[None, :] * lam.astype(w.dtype)[:, None]).reshape(-1)         # _solvers.py:138-141
    return pack_particles(y, w)


# --------------------------------------------------------------------------- kernels
class MonteCarloKernel:
    """`MonteCarloKernel` (mccube/_kernels/random.py:22-68)."""

    def __init__(self, recombination_count, with_replacement=False, weighting_function=None, *,
                 key, partitionable=False, x64=False):
        self.x64 = x64
        self.recombination_count = recombination_count
        self.with_replacement = with_replacement
        self.weighting_function = weighting_function
        self.key = key
        self.partitionable = partitionable

    def indices(self, t, n_inputs, count=None, p=None):
        count = self.recombination_count if count is None else count
        shape = tuple(count) if isinstance(count, tuple) else (count,)
        key = jr.mc_step_key(self.key, t, partitionable=self.partitionable)
        idx = jr.choice_indices(key, n_inputs, int(np.prod(shape)), self.with_replacement, p,
                                partitionable=self.partitionable, x64=self.x64)
        return idx.reshape(shape)

    def __call__(self, t, particles, args=None, weighted=False, count=None):
        p = None
        if weighted and self.weighting_function is not None:
            p = self.weighting_function(particles[..., -1])
        idx = self.indices(t, particles.shape[0], count, p)
        return particles[idx]


class MonteCarloPartitioningKernel:
    """`MonteCarloPartitioningKernel` (mccube/_kernels/random.py:71-115)."""

    def __init__(self, partition_count, monte_carlo_kernel):
        self.partition_count = partition_count
        self.monte_carlo_kernel = monte_carlo_kernel

    def __call__(self, t, particles, args=None, weighted=False):
        m = self.partition_count
        return self.monte_carlo_kernel(t, particles, args, weighted, count=(m, particles.shape[0] // m))


def center_of_mass(p, w=None):
    """`center_of_mass` (mccube/_metrics.py:79-96) = jnp.average(p, 0, w)."""
    return np.average(p, axis=0, weights=w)


def stratified_indices(p, m, weighted=False, norm="l2"):
    """Index form of `StratifiedPartitioningKernel` (mccube/_kernels/stratified.py:36-45).

    fp32 policy for bit-exact comparison with CUDA: centre on the COM (given, fp32),
    accumulate squares sequentially over columns in fp32, sqrt, stable argsort."""
    pu, w = unpack_particles(p, weighted)
    com = center_of_mass(pu.astype(np.float64), None if w is None else w.astype(np.float64)).astype(p.dtype)
    c = (pu - com).astype(p.dtype)
    acc = np.zeros(c.shape[0], p.dtype)
    for a in range(c.shape[1]):
        acc = (acc + (c[:, a] * c[:, a]).astype(p.dtype)).astype(p.dtype)
    key = np.sqrt(acc).astype(p.dtype)
    return np.argsort(key, kind="stable").reshape(m, -1), com, key


class StratifiedPartitioningKernel:
    def __init__(self, partition_count, norm="l2"):
        self.partition_count = partition_count
        self.norm = norm

    def __call__(self, t, particles, args=None, weighted=False):
        m = self.partition_count
        if self.norm is None:
            return particles.reshape(m, -1, particles.shape[-1])           # stratified.py:37-38
        idx, _, _ = stratified_indices(particles, m, weighted)
        return particles[idx]


class BinaryTreePartitioningKernel:
    """`BinaryTreePartitioningKernel` (mccube/_kernels/tree.py:37-60): sklearn is installed
    here, so the oracle calls it exactly as the reference's host callback does."""

    def __init__(self, partition_count, mode="BallTree", metric="minkowski", metric_kwargs=None):
        self.partition_count = partition_count
        self.mode = mode
        self.metric = metric
        self.metric_kwargs = metric_kwargs or {}

    def indices(self, particles, weighted=False):
        from sklearn.neighbors import BallTree, KDTree

        trees = {"KDTree": KDTree, "BallTree": BallTree}
        pu, _ = unpack_particles(particles, weighted)
        leaf_size = particles.shape[0] // self.partition_count
        t = trees[self.mode](np.asarray(pu), leaf_size, self.metric, **self.metric_kwargs)
        return np.asarray(t.get_arrays()[1]).reshape(-1, leaf_size).astype(np.int32)

    def __call__(self, t, particles, args=None, weighted=False):
        return particles[self.indices(particles, weighted)]


def median_split_levels(n: int, leaf_size: int) -> int:
    """Split levels of the device tree: floor(log2(n / leaf_size)); sklearn's own tree has
    floor(log2((n - 1) // leaf_size)) of them (sklearn/neighbors/_binary_tree.pxi, `n_levels`)."""
    levels = 0
    while n >= leaf_size and (n >> (levels + 1)) >= leaf_size:
        levels += 1
    return levels


def median_split_indices(points, levels: int) -> np.ndarray:
    """Restatement of sklearn's `_recursive_build` partition rule (the tree the reference builds through
    `jax.pure_callback`, mccube/_kernels/tree.py:44-58), level by level as `csrc/kd_split.cu` does it: a node is the
    positions [start, end) of the index array; its split dimension is the one of largest max - min over its points
    (first on ties, compared in float64 as sklearn's float64 trees do); its points are ordered along that dimension
    by a STABLE sort (sklearn: `nth_element`, whose order inside the halves is platform dependent) and it is cut at
    (end - start) // 2.  The point SET of every node equals sklearn's; the order inside a leaf is this function's
    definition (sorted along the last split dimension, ties in the previous order)."""
    pts = np.asarray(points)
    n = pts.shape[0]
    idx = np.arange(n, dtype=np.int64)
    bounds = [(0, n)]
    for _ in range(levels):
        nxt = []
        for s, e in bounds:
            if e > s:
                node = pts[idx[s:e]]
                spread = node.max(axis=0).astype(np.float64) - node.min(axis=0).astype(np.float64)
                j = int(np.argmax(spread))
                # -0.0 orders before +0.0 in the device's integer image of a float; numpy treats them as equal
                key = node[:, j].astype(np.float64)
                key = np.where((node[:, j] == 0) & np.signbit(node[:, j]), -np.finfo(np.float64).tiny, key)
                idx[s:e] = idx[s:e][np.argsort(key, kind="stable")]
            mid = s + (e - s) // 2
            nxt += [(s, mid), (mid, e)]
        bounds = nxt
    return idx


def box_keys(p, lo, inv_h, dims, bits):
    """Integer box key of `BoxPartitionKernel` (NEW kernel -- absent from the reference,
    SURVEY.md F2; this function is its definition).

    cell_a = clamp(int(floor((x_a - lo_a) * inv_h_a)), 0, 2^bits - 1)  in fp32, per key dim a;
    key    = sum_a cell_a << (bits * (len(dims) - 1 - i_a))             (lexicographic grid id).
    """
    p = np.asarray(p)
    key = np.zeros(p.shape[0], dtype=np.uint32)
    hi = (1 << bits) - 1
    for i, a in enumerate(dims):
        v = ((p[:, a] - p.dtype.type(lo[i])).astype(p.dtype) * p.dtype.type(inv_h[i])).astype(p.dtype)
        cell = np.clip(np.floor(v), 0, hi).astype(np.uint32)
        key |= cell << np.uint32(bits * (len(dims) - 1 - i))
    return key


def box_bounds(p, dims, bits):
    """Bounding box used by BoxPartitionKernel: lo_a = min_a, inv_h_a = 2^bits / (max_a - min_a)
    (fp32; a zero range maps everything to cell 0)."""
    lo = np.array([p[:, a].min() for a in dims], dtype=p.dtype)
    hi = np.array([p[:, a].max() for a in dims], dtype=p.dtype)
    rng = (hi - lo).astype(p.dtype)
    with np.errstate(divide="ignore"):
        inv_h = np.where(rng > 0, p.dtype.type(1 << bits) / rng, p.dtype.type(0)).astype(p.dtype)
    return lo, inv_h


class BoxPartitionKernel:
    """Box-binning partitioner: stable sort by integer box key, then `m` equal chunks
    (the `[n,d] -> [m, n/m, d]` contract of mccube/_kernels/base.py:54-75)."""

    def __init__(self, partition_count, dims=None, bits=3, bounds=None):
        self.partition_count = partition_count
        self.dims = dims
        self.bits = bits
        self.bounds = bounds

    def resolve(self, particles, weighted=False):
        pu, _ = unpack_particles(particles, weighted)
        dims = tuple(self.dims) if self.dims is not None else tuple(range(min(4, pu.shape[1])))
        if self.bounds is None:
            lo, inv_h = box_bounds(pu, dims, self.bits)
        else:
            lo = np.asarray(self.bounds[0], pu.dtype)
            hi = np.asarray(self.bounds[1], pu.dtype)
            inv_h = (pu.dtype.type(1 << self.bits) / (hi - lo)).astype(pu.dtype)
        return pu, dims, lo, inv_h

    def indices(self, particles, weighted=False):
        pu, dims, lo, inv_h = self.resolve(particles, weighted)
        key = box_keys(pu, lo, inv_h, dims, self.bits)
        return np.argsort(key, kind="stable").reshape(self.partition_count, -1)

    def __call__(self, t, particles, args=None, weighted=False):
        return particles[self.indices(particles, weighted)]


class PartitioningRecombinationKernel:
    """`PartitioningRecombinationKernel` (mccube/_kernels/base.py:139-177).  The inner kernel
    is closed over by `filter_vmap`, not batched, so every partition uses the SAME key and
    therefore the same local index vector (quirk Q4)."""

    def __init__(self, partitioning_kernel, recombination_kernel):
        self.partitioning_kernel = partitioning_kernel
        self.recombination_kernel = recombination_kernel
        self.recombination_count = recombination_kernel.recombination_count * partitioning_kernel.partition_count

    def __call__(self, t, particles, args=None, weighted=False):
        parts = self.partitioning_kernel(t, particles, args, weighted)        # [m, P/m, d]
        rk = self.recombination_kernel
        uniform = not (weighted and getattr(rk, "weighting_function", None) is not None)
        if isinstance(rk, MonteCarloKernel) and uniform:
            # quirk Q4: the vmapped kernel closes over ONE key, so the index vector is the same for every
            # partition -- draw it once (what XLA's vmap does), not m times
            idx = rk.indices(t, parts.shape[1])
            rec = parts[:, idx]
        else:
            rec = np.stack([rk(t, q, args, weighted) for q in parts])
        return rec.reshape(-1, particles.shape[-1])


# --------------------------------------------------------------------------- solver
def mcc_step(y0, t0, t1, drift, points, lam, g, kernel, n_substeps=1, weighted=False, time_dtype=None):
    """`MCCSolver.step` (mccube/_solvers.py:115-148).  Kernel is called with the OUTER t0
    (quirk Q2, `_solvers.py:144`), then weights are renormalised (`:146`)."""
    y1_packed = expand(y0, t0, t1, drift, points, lam, g, n_substeps, weighted, time_dtype)
    y1_hat = kernel(t0, y1_packed, None, weighted)
    return pack_particles(*unpack_particles(y1_hat, weighted))


def _clip_to_end(tprev, tnext, t1, dtype):
    """diffrax/_integrate.py `_clip_to_end` (keep_step is always True for ConstantStepSize)."""
    tol = 1e-10 if np.dtype(dtype) == np.float64 else 1e-6
    return dtype(t1) if tnext > t1 - dtype(tol) else tnext


def step_times(t0, t1, dt0, dtype=np.float32, max_steps=4096):
    """(tprev, tnext) of every `diffeqsolve` step under `ConstantStepSize`
    (diffrax: init -> t0 + dt0; adapt_step_size -> (t1, t1 + dt0); loop clips to the end)."""
    dtype = np.dtype(dtype).type
    t0, t1, dt0 = dtype(t0), dtype(t1), dtype(dt0)
    tprev, tnext = t0, _clip_to_end(t0, dtype(t0 + dt0), t1, dtype)
    out = []
    while tprev < t1 and len(out) < max_steps:
        out.append((tprev, tnext))
        tprev, tnext = tnext, dtype(tnext + dt0)
        tprev = min(tprev, t1)
        tnext = _clip_to_end(tprev, tnext, t1, dtype)
    return out


def diffeqsolve(y0, t0, t1, dt0, drift, points, lam, g, kernel, n_substeps=1, weighted=False,
                time_dtype=None, save_steps=None):
    """Constant-step `diffrax.diffeqsolve` loop around `mcc_step` (README.md:81-84).
    Returns the final state, plus {step_index: state-after-that-step} for `save_steps`."""
    td = time_dtype or y0.dtype
    y = y0
    saved = {}
    for i, (a, b) in enumerate(step_times(t0, t1, dt0, td)):
        y = mcc_step(y, a, b, drift, points, lam, g, kernel, n_substeps, weighted, td)
        if save_steps is not None and i in save_steps:
            saved[i] = y
    return y, saved


# --------------------------------------------------------------------------- metrics
def _sqrtm_psd(a):
    w, v = np.linalg.eigh((a + a.T) / 2)
    return (v * np.sqrt(np.clip(w, 0, None))) @ v.T


def gaussian_wasserstein_metric(means, covs):
    """`gaussian_wasserstein_metric` (mccube/_metrics.py:39-44); complex dtype like the
    reference's Schur sqrtm output (`tests/test_metrics.py:13`)."""
    m1, m2 = (np.asarray(m, np.float64) for m in means)
    c1, c2 = (np.asarray(c, np.float64) for c in covs)
    root_c2 = _sqrtm_psd(c2)
    mean_dist = np.linalg.norm(m1 - m2) ** 2
    cov_dist = np.trace(c1 + c2 - 2 * _sqrtm_psd(root_c2 @ c1 @ root_c2))
    return np.complex128(mean_dist + cov_dist)
