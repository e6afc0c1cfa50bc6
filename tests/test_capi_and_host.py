"""CPU-only checks: the C-ABI library builds, loads and exports every symbol that
include/mccube_b200.h declares (no compute calls), and the host-side mirror of the reference
API behaves like the reference's own tests expect."""
import math
import re
import warnings
from pathlib import Path

import numpy as np
import pytest
import torch

import mccube_b200 as mccube
from mccube_b200 import _lib
from mccube_b200 import diffrax_compat as dfx
from oracle import mccube_ref as ref

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "mccube_b200.h").read_text()
    declared = set(re.findall(r"MCCB200_API[^;(]*?\b(mccb200_\w+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.mccb200_version() == 100
    # host-only entry points are callable without a GPU
    assert lib.mccb200_sort_pairs_workspace_bytes(1 << 20) > 3 * 4 * (1 << 20)
    assert lib.mccb200_mc_indices_workspace_bytes(1 << 14, 512, 1) < 4096


def test_no_cpu_fallback():
    p = torch.ones(4, 2)
    with pytest.raises(_lib.MccubeB200Error, match="CUDA"):
        _lib.gather_rows(p, torch.zeros(1, dtype=torch.int32))
    k = mccube.MonteCarloKernel(2, key=42)
    solver = mccube.MCCSolver(dfx.Euler(), k)
    with pytest.raises(_lib.MccubeB200Error, match="CUDA"):
        solver.step(None, 0.0, 0.05, p, None, None, False)


def test_product_does_not_import_oracle():
    for f in (ROOT / "mccube_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_solver_init_errors_and_warnings():
    # /root/reference/tests/test_solvers.py:55-64,78-79
    with pytest.raises(ValueError) as e, pytest.warns(UserWarning) as w:
        mccube.MCCSolver(dfx.EulerHeun(), mccube.MonteCarloKernel(10, key=42), 0)
    assert "n_substeps must be at least one;" in e.value.args[0]
    assert len(w) == 1 and "diffrax.Euler solver" in w[0].message.args[0]
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        s = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(10, key=42))
    assert s.term_structure is mccube.MCCTerm and s.interpolation_cls is dfx.LocalLinearInterpolation


def test_step_times_match_oracle():
    for (t0, t1, dt0, td) in [(0.0, 25.6, 0.05, np.float32), (0.0, 51.2, 0.05, np.float64), (1.0, 1.37, 0.1, np.float32)]:
        assert dfx.step_times(t0, t1, dt0, td) == ref.step_times(t0, t1, dt0, td)
    s = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(10, key=42), n_substeps=3)
    assert s.substep_times(np.float32(0.3), np.float32(0.35)) == ref.substep_times(np.float32(0.3), np.float32(0.35), 3)


@pytest.mark.parametrize("d", [3, 5, 10])
def test_local_linear_cubature_path(d):
    # /root/reference/tests/test_path.py:10-23
    f = mccube.Hadamard(mccube.GaussianRegion(d))
    path = mccube.LocalLinearCubaturePath(f)
    assert np.array_equal(np.abs(path.evaluate(0.0)), np.zeros(f.stacked_points.shape))
    assert np.array_equal(path.evaluate(1.0), path.evaluate(0.0, 1.0))
    assert np.array_equal(path.evaluate(0.0, 1.0), f.stacked_points.astype(np.float32))
    assert np.array_equal(path.evaluate(0.0, 0.5), (np.float32(np.sqrt(np.float32(0.5))) * path.evaluate(0.0, 1.0)))
    assert np.array_equal(path.weights, f.stacked_weights)
    assert np.array_equal(f.points, ref.hadamard_points(d)) and f.point_count == ref.hadamard_point_count(d)


def test_mcc_term_shapes():
    # /root/reference/tests/test_term.py:8-48
    class Control:
        t0, t1 = 0, 1

        def evaluate(self, t0, t1=None, left=True):
            return {"y": torch.ones((8, 2))}

    ode = dfx.ODETerm(lambda t, y, args: {"y": -y["y"]})
    cde = dfx.WeaklyDiagonalControlTerm(lambda t, y, args: {"y": 1.0}, Control())
    term = mccube.MCCTerm(ode, cde)
    dx = term.contr(0, 1)
    y = {"y": torch.tensor([[1.0, 2.0], [2.0, 4.0], [5.0, 6.0]])[:, None, :]}
    vf = term.vf(0, y, None)
    vf_prod = term.vf_prod(0, y, None, dx)
    assert np.shape(dx[0]) == () and tuple(dx[1]["y"].shape) == (8, 2)
    assert tuple(vf[0]["y"].shape) == (3, 2) and vf[1]["y"] == 1.0
    assert torch.equal(vf[0]["y"], -y["y"].squeeze())
    assert tuple(vf_prod["y"].shape) == (3, 8, 2)
    y2 = vf_prod
    vf2 = term.vf(0, y2, None)
    assert tuple(vf2[0]["y"].shape) == (24, 2)
    assert tuple(term.vf_prod(0, y2, None, dx)["y"].shape) == (24, 8, 2)


def test_pack_unpack():
    # /root/reference/tests/test_utils.py:6-24
    p = torch.ones((4, 4)); w = torch.full((4, 1), 0.25)
    exp = torch.cat([p, w], 1)
    assert torch.equal(exp, mccube.pack_particles(p, w))
    assert torch.equal(exp, mccube.pack_particles(p, w * 4))
    assert mccube.pack_particles(p, None) is p
    x = torch.cat([torch.ones((3, 4)), torch.zeros((3, 1))], 1)
    assert mccube.unpack_particles(x, False)[1] is None
    a, b = mccube.unpack_particles(x, True)
    assert torch.equal(a, x[:, :-1]) and torch.equal(b, x[:, -1])


def test_gaussian_wasserstein_metric():
    # /root/reference/tests/test_metrics.py:8-21
    d = 3
    dist = mccube.gaussian_wasserstein_metric((np.ones(d), np.ones(d)), (np.eye(d), np.eye(d)))
    assert dist == np.complex128(0.0) and dist.dtype == np.complex128
    dist = mccube.gaussian_wasserstein_metric((np.ones(d), 2 * np.ones(d)), (np.eye(d), 2 * np.eye(d)))
    assert abs(dist - (3 + 3 * (1 + 2 - 2 * np.sqrt(2)))) < 1e-12
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(2, 5, 5))
    c1, c2 = a @ a.T + np.eye(5), b @ b.T + np.eye(5)
    m1, m2 = rng.normal(size=(2, 5))
    assert abs(mccube.gaussian_wasserstein_metric((m1, m2), (c1, c2)) - ref.gaussian_wasserstein_metric((m1, m2), (c1, c2))) < 1e-12


def test_prk_recombination_count():
    mc = mccube.MonteCarloKernel(4, key=42)
    prk = mccube.PartitioningRecombinationKernel(mccube.BoxPartitionKernel(8), mc)
    assert prk.recombination_count == 32
    assert mccube.BoxPartitionKernel(8).resolve_dims(64) == (0, 1, 2, 3)
    with pytest.raises(ValueError):
        mccube.BoxPartitionKernel(8, dims=range(8), bits=5).resolve_dims(64)


# ------------------------------------------------------------------ cubature formulae (host side)
def _hermite_monomial(mono):
    """E[prod x_i^m_i] under N(0, I): 0 if any exponent is odd, else prod (m_i - 1)!!
    (the closed form the reference test computes with gamma functions, tests/test_formulae.py:181-193)."""
    from scipy.special import factorial2

    return float(np.prod([0.0 if m % 2 else (1.0 if m == 0 else factorial2(m - 1)) for m in mono]))


@pytest.mark.parametrize("name,degree,dims", [
    ("Hadamard", 3, [1, 2, 3, 4]), ("StroudSecrest63_31", 3, [1, 2, 3, 4]), ("StroudSecrest63_32", 3, [1, 2, 3, 4]),
    ("StroudSecrest63_52", 5, [2, 3, 4]), ("StroudSecrest63_53", 5, [3, 4]),
])
def test_gaussian_cubature_exactness(name, degree, dims):
    """/root/reference/tests/test_formulae.py:132-161: every monomial of degree <= m integrates
    exactly; points / weights / point_count are consistent (:20-31)."""
    from itertools import product

    import mccube_b200 as mccube

    for d in dims:
        f = getattr(mccube, name)(mccube.GaussianRegion(d))
        pts = f.points if isinstance(f.points, tuple) else (f.points,)
        assert f.point_count == sum(q.shape[0] for q in pts) == f.stacked_points.shape[0] == f.stacked_weights.shape[0]
        assert f.stacked_points.shape[1] == d and f.degree == degree
        for mono in product(range(degree + 1), repeat=d):
            if sum(mono) > degree:
                continue
            got, vals = f(lambda x: float(np.prod(x ** np.array(mono))))
            assert abs(got - _hermite_monomial(mono)) <= 1e-5 * abs(_hermite_monomial(mono)) + 1e-8, (name, d, mono)
            assert vals.shape == (f.point_count,)


def test_cubature_registry_search_and_permutations():
    """/root/reference/tests/test_formulae.py:34-129."""
    import mccube_b200 as mccube
    from mccube_b200._formulae import _generate_point_permutations, builtin_cubature_registry

    assert builtin_cubature_registry == mccube.all_subclasses(mccube.AbstractCubature)
    pool = {mccube.Hadamard, mccube.StroudSecrest63_31, mccube.StroudSecrest63_32, mccube.StroudSecrest63_52}
    region = mccube.GaussianRegion(5)
    cases = [
        ((None, False, False), ["StroudSecrest63_31", "Hadamard", "StroudSecrest63_32", "StroudSecrest63_52"]),
        ((3, False, False), ["StroudSecrest63_31", "Hadamard", "StroudSecrest63_32"]),
        ((3, False, True), ["StroudSecrest63_31"]),
        ((3, True, False), ["StroudSecrest63_31"]),
        ((5, True, True), ["StroudSecrest63_52"]),
    ]
    for (deg, sparse, minimal), want in cases:
        got = mccube.search_cubature_registry(region, deg, sparse, minimal, pool)
        assert [type(f).__name__ for f in got] == want and all(f.region is region for f in got)
    assert [type(f).__name__ for f in mccube.search_cubature_registry(region, 3, False, True)] == ["StroudSecrest63_31"]
    with pytest.raises(ValueError, match="only valid for regions with d > 2"):
        mccube.StroudSecrest63_53(mccube.GaussianRegion(2))
    fs = _generate_point_permutations(np.array([3, 0, 0]), "FS")
    assert np.array_equal(fs, [[-3, 0, 0], [0, -3, 0], [0, 0, -3], [0, 0, 3], [0, 3, 0], [3, 0, 0]])
    assert np.array_equal(_generate_point_permutations(np.array([3, 0, 0]), "S"), fs[3:])
    rr = _generate_point_permutations(np.array([3, 3, 3]), "R")
    assert np.array_equal(rr, [[-3, -3, -3], [-3, -3, 3], [-3, 3, -3], [-3, 3, 3], [3, -3, -3], [3, -3, 3], [3, 3, -3], [3, 3, 3]])
    with pytest.raises(ValueError):
        _generate_point_permutations(np.ones(2), "X")


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the oracle on the host cores) needs no GPU; stdout is exactly one JSON
    line with the contract's keys, whatever else the process prints (that goes to stderr)."""
    import json
    import subprocess
    import sys

    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "particle_steps_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    for key in ("unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "config"):
        assert key in line, key


def test_diffeqsolve_protocol_on_cpu():
    """The constant-step driver (stand-in for diffrax.diffeqsolve, diffrax is not installable here) with the
    plain Euler solver on CPU tensors: step count and float32 time accumulation, SaveAt(t0 / t1 / steps / ts /
    dense), linear interpolation inside a step, max_steps."""
    f = lambda t, y, args: -y                                               # noqa: E731
    term = dfx.ODETerm(f)
    y0 = torch.tensor([[1.0, 2.0]])
    dt = 0.1
    sol = dfx.diffeqsolve(term, dfx.Euler(), 0.0, 1.0, dt, y0, saveat=dfx.SaveAt(t0=True, steps=True))
    times = dfx.step_times(0.0, 1.0, dt, np.float32)
    assert sol.stats["num_steps"] == len(times) == 10 and sol.result == dfx.RESULTS.successful
    assert sol.ys.shape == (11, 1, 2) and np.isclose(sol.ts[-1], 1.0)
    want = y0.clone()
    for (a, b) in times:
        want = want + (-want) * float(np.float32(b) - np.float32(a))
    assert torch.allclose(sol.ys[-1], want, rtol=1e-6)
    # ts inside steps are linear interpolations of the step's end points; ts at step boundaries are exact
    sol_ts = dfx.diffeqsolve(term, dfx.Euler(), 0.0, 1.0, dt, y0, saveat=dfx.SaveAt(ts=[0.0, 0.25, 0.5, 1.0]))
    assert sol_ts.ys.shape[0] == 4 and torch.equal(sol_ts.ys[0], y0)
    i25 = [i for i, (a, b) in enumerate(times) if a <= np.float32(0.25) <= b][0]
    a, b = times[i25]
    lo, hi = sol.ys[i25], sol.ys[i25 + 1]
    w = (float(np.float32(0.25)) - float(a)) / (float(b) - float(a))
    assert torch.allclose(sol_ts.ys[1], lo + w * (hi - lo), rtol=1e-6)
    assert torch.allclose(sol_ts.ys[-1], sol.ys[-1])
    # dense output evaluates anywhere; without it evaluate() raises
    dense = dfx.diffeqsolve(term, dfx.Euler(), 0.0, 1.0, dt, y0, saveat=dfx.SaveAt(dense=True))
    assert torch.allclose(dense.evaluate(0.25), sol_ts.ys[1], rtol=1e-6)
    with pytest.raises(ValueError):
        sol.evaluate(0.25)
    # max_steps
    short = dfx.diffeqsolve(term, dfx.Euler(), 0.0, 1.0, dt, y0, max_steps=4)
    assert short.result == dfx.RESULTS.max_steps_reached and short.stats["num_steps"] == 4
    # capture_solve refuses CPU states
    with pytest.raises(ValueError, match="CUDA tensor"):
        dfx.capture_solve(term, dfx.Euler(), 0.0, 1.0, dt, y0)


def test_pairwise_metric_reference_expectation():
    """/root/reference/tests/test_metrics.py:37-45 (host-side helper, plain torch)."""
    y0 = torch.tensor([[0.0, 0.0], [3.0, 4.0], [5.0, 12.0]])
    dist = mccube.pairwise_metric(y0, y0)
    want = torch.tensor([[0.0, 5.0, 13.0], [5.0, 0.0, math.sqrt(68)], [13.0, math.sqrt(68), 0.0]])
    assert torch.allclose(dist, want, rtol=1e-6)
    # result[j, i] = metric(xs[i], ys[j]) for different inputs (_metrics.py:63-76)
    xs, ys = torch.tensor([[0.0, 0.0], [1.0, 0.0]]), torch.tensor([[0.0, 2.0], [0.0, 0.0], [3.0, 4.0]])
    got = mccube.pairwise_metric(xs, ys)
    assert got.shape == (3, 2) and torch.allclose(got[2], torch.tensor([5.0, math.sqrt(20.0)]))


def test_key_words_and_tree_helpers():
    """Host-side plumbing: PRNG key forms (jr.key(seed) = (seed >> 32, seed & 0xffffffff))
    and the PyTree helpers with jax.tree_util's leaf order (dict keys sorted)."""
    from mccube_b200._kernels.random import key_words
    from mccube_b200._tree import tree_leaves, tree_map, tree_map_with_index

    assert key_words(42) == (0, 42) and key_words((1 << 35) + 7) == (8, 7)
    assert key_words((3, 4)) == (3, 4) and key_words(torch.tensor([5, 6])) == (5, 6)
    assert key_words(np.array([2 ** 32 + 1, -1])) == (1, 0xFFFFFFFF)
    tree = {"b": [torch.ones(2), (torch.zeros(1), torch.full((1,), 2.0))], "a": torch.full((3,), 7.0)}
    leaves = tree_leaves(tree)
    assert [int(x.numel()) for x in leaves] == [3, 2, 1, 1]                 # "a" before "b"
    doubled = tree_map(lambda x: 2 * x, tree)
    assert isinstance(doubled["b"][1], tuple) and float(doubled["a"][0]) == 14.0
    # a bare leaf as the second tree broadcasts (a prefix tree), like the reference's integer counts
    summed = tree_map(lambda x, c: x + c, tree, 1.0)
    assert float(summed["b"][1][1][0]) == 3.0
    seen = []
    tree_map_with_index(lambda i, n, leaf: seen.append((i, n, int(leaf.numel()))), tree)
    assert seen == [(0, 4, 3), (1, 4, 2), (2, 4, 1), (3, 4, 1)]


def test_c_abi_argument_errors_without_a_gpu(lib):
    """every entry point validates before it touches the device and reports through a
    status code + thread-local message (what the XLA-FFI shim turns into ffi::Error); size queries are pure host
    functions.  None of this needs a GPU."""
    import ctypes as C

    lib.mccb200_last_error.restype = C.c_char_p
    rng = _lib.make_rng((0, 42), 0.0)
    assert lib.mccb200_mc_indices(C.byref(rng), 10, 20, 0, None, None, 0, None) != 0
    assert b"NULL argument" in lib.mccb200_last_error()
    fake = C.c_void_p(64)       # non-NULL, never dereferenced: validation fails first
    assert lib.mccb200_mc_indices(C.byref(rng), 10, 20, 0, fake, fake, 1 << 30, None) != 0
    assert b"without replacement" in lib.mccb200_last_error()
    assert lib.mccb200_mc_indices(C.byref(rng), 1 << 20, 1 << 10, 0, fake, fake, 16, None) != 0
    assert b"workspace" in lib.mccb200_last_error()
    assert lib.mccb200_compact_slots(None, 10, None, 10, None, None) != 0
    assert b"compact_slots" in lib.mccb200_last_error()
    # workspace sizes: pure functions, monotone in the problem size, zero-replace draw needs no sort scratch
    small = lib.mccb200_mc_indices_workspace_bytes(16384, 512, 0)
    big = lib.mccb200_mc_indices_workspace_bytes(1 << 25, 1 << 20, 0)
    assert 0 < small < big and lib.mccb200_mc_indices_workspace_bytes(1 << 25, 1 << 20, 1) < small
    assert lib.mccb200_sort_pairs_workspace_bytes(1 << 20) >= 3 * 4 * (1 << 20)
    n_keys = C.c_uint32(0)
    off = lib.mccb200_box_prk_key_start_offset(1 << 20, 128, 4, 3, 8192, 64, C.byref(n_keys))
    assert n_keys.value == 1 << 12 and off % 4 == 0
    assert off < lib.mccb200_box_prk_step_workspace_bytes(1 << 20, 128, 4, 3, 8192)
    assert lib.mccb200_box_prk_key_start_offset(1 << 20, 128, 6, 6, 8192, 64, C.byref(n_keys)) == C.c_size_t(-1).value
    # tensor-core drift: multiples of 32 columns, 16-byte aligned rows
    assert lib.mccb200_gaussian_full_drift_tc_supported(1 << 20, 512, 512) == 1
    assert lib.mccb200_gaussian_full_drift_tc_supported(100, 48, 48) == 0
    assert lib.mccb200_gaussian_full_drift_tc_supported(100, 64, 66) == 0


def test_round2_entry_points_validate_and_size_without_a_gpu(lib):
    """The entry points added in round 2: shape predicates, level counts and workspace sizes are pure host functions;
    argument validation reports through the status code + message before anything touches the device."""
    import ctypes as C

    lib.mccb200_last_error.restype = C.c_char_p
    # in-kernel index draw: unweighted rows of 4 .. 32 float4
    for d, ld, w, ok in ((64, 64, 0, 1), (16, 16, 0, 1), (128, 128, 0, 1), (12, 12, 0, 0), (64, 65, 1, 0), (256, 256, 0, 0),
                         (8, 8, 0, 0)):
        assert lib.mccb200_expand_select_draw_supported(d, ld, w) == ok, (d, ld, w)
    assert lib.mccb200_expand_select_draw_workspace_bytes() >= 64
    fake = C.c_void_p(256)
    drift = _lib.Drift()
    cub = _lib.Cubature()
    cub.k = 128
    rng = _lib.make_rng((0, 42), 0.0)
    assert lib.mccb200_expand_select_draw(fake, 1 << 26, 64, C.byref(drift), 0.05, C.byref(cub), C.byref(rng), 1 << 10, fake,
                                          None, fake, 256, None) != 0       # n * k = 2^33 child ids
    assert b"child ids" in lib.mccb200_last_error()
    # median-split tree: levels = floor(log2(n / leaf)); sklearn's tree has floor(log2((n - 1) // leaf)) split levels
    for n, leaf in ((1 << 20, 64), (1000, 10), (777, 7), (512, 512), (4096, 64), (3000, 100), (65, 64), (63, 64)):
        want = 0
        while n >= leaf and (n >> (want + 1)) >= leaf:
            want += 1
        got = lib.mccb200_kd_partition_levels(n, leaf)
        assert got == want, (n, leaf)
        if n > leaf:
            sk = int(np.log2(max(1, (n - 1) // leaf)))
            assert got in (sk, sk + 1)
    lib.mccb200_kd_partition_workspace_bytes.restype = C.c_size_t
    small = lib.mccb200_kd_partition_workspace_bytes(1 << 16, 8, 10)
    big = lib.mccb200_kd_partition_workspace_bytes(1 << 20, 64, 14)
    assert 0 < small < big
    assert lib.mccb200_kd_partition(None, 100, 8, 8, 3, None, None, 0, None) != 0
    assert b"kd_partition" in lib.mccb200_last_error()
    assert lib.mccb200_kd_partition(fake, 100, 300, 300, 3, fake, fake, 1 << 30, None) != 0     # d > 256
    # selection shuffle: a quarter of the scratch of the sort shuffle at config 2's size, and only up to 2^27 inputs
    sel = lib.mccb200_mc_indices_workspace_bytes(1 << 25, 1 << 20, 0)
    assert sel < 3 * 4 * (1 << 25) + lib.mccb200_sort_pairs_workspace_bytes(1 << 25)
    assert lib.mccb200_mc_indices_workspace_bytes(1 << 28, 1 << 20, 0) >= 3 * 4 * (1 << 28)
    # renormalised copy / id bucketing
    assert lib.mccb200_normalised_copy(None, 10, 5, None, None, None, None) != 0
    assert b"normalised_copy" in lib.mccb200_last_error()
    assert lib.mccb200_shard_bucket_ids_workspace_bytes(8) > 0
