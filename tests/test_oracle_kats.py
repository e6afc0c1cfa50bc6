"""Oracle vs known answers: Random123 KATs, published JAX values, SURVEY Appendix-A goldens,
and the committed fixtures (oracle regression)."""
import numpy as np

from oracle import jax_random as jr
from oracle import mccube_ref as ref

from .conftest import GOLDEN


def test_threefry_random123_kats():
    # Random123 kat_vectors, threefry2x32 20 rounds
    assert tuple(map(int, jr.threefry2x32((0, 0), 0, 0))) == (0x6B200159, 0x99BA4EFE)
    assert tuple(map(int, jr.threefry2x32((0xFFFFFFFF, 0xFFFFFFFF), 0xFFFFFFFF, 0xFFFFFFFF))) == (0x1CB996FC, 0xBB002BE7)
    assert tuple(map(int, jr.threefry2x32((0x13198A2E, 0x03707344), 0x243F6A88, 0x85A308D3))) == (0xC4923A9C, 0x483DF7A0)


def test_published_jax_values_non_partitionable():
    # jax docs / test-suite values for the default (pre-0.5) threefry mode
    assert [tuple(map(int, k)) for k in jr.split(jr.prng_key(0))] == [(4146024105, 967050713), (2718843009, 1272950319)]
    assert [tuple(map(int, k)) for k in jr.split(jr.prng_key(42))] == [(2465931498, 3679230171), (255383827, 267815257)]
    assert jr.random_bits(jr.prng_key(1701), 32, 3).tolist() == [56197195, 4200222568, 961309823]
    assert abs(float(jr.uniform(jr.prng_key(0), 1)[0]) - 0.41845703) < 1e-7
    assert abs(float(jr.normal(jr.prng_key(0), 1)[0]) - (-0.20584226)) < 1e-6
    assert abs(float(jr.normal(jr.prng_key(42), 1)[0]) - (-0.18471177)) < 1e-6
    # the vector printed in JAX's own README / quickstart for `random.normal(random.PRNGKey(0), (10,))`: ten values pin
    # the layout of random_bits (word i pairs with word i + ceil(n/2) of one threefry block) and the bits -> uniform ->
    # normal chain; the last digit depends on the erfinv implementation
    want = [-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377, -0.1521442, -0.67135346, -0.5908641, 0.73168886,
            0.5673026]
    assert np.allclose(np.asarray(jr.normal(jr.prng_key(0), 10)), want, rtol=0, atol=1e-6)


def test_partitionable_structure():
    key = jr.prng_key(42)
    ks = jr.split(key, 5, partitionable=True)
    for i, k in enumerate(ks):
        assert tuple(map(int, k)) == tuple(map(int, jr.fold_in(key, i)))
    bits = jr.random_bits(key, 32, 7, partitionable=True)
    for i in range(7):
        o = jr.threefry2x32(key, 0, i)
        assert int(bits[i]) == int(o[0]) ^ int(o[1])


def test_survey_appendix_a_goldens():
    key = jr.prng_key(42)
    assert tuple(map(int, jr.fold_in(key, 0))) == (1832780943, 270669613)
    assert tuple(map(int, jr.fold_in(key, 0x3D4CCCCD))) == (2091507901, 2947249837)
    assert jr.float_time_bits(0.05) == 0x3D4CCCCD
    assert tuple(map(int, jr.mc_step_key(key, 0.0))) == (1705926158, 899080142)
    assert tuple(map(int, jr.mc_step_key(key, 0.05))) == (3902356039, 3131044844)
    assert tuple(map(int, jr.mc_step_key(key, 0.0, partitionable=True))) == (1012194634, 3152801799)
    k = jr.mc_step_key(key, 0.0)
    assert jr.shuffle_perm(k, 16).tolist() == [1, 9, 2, 14, 13, 6, 10, 0, 7, 3, 4, 15, 11, 12, 5, 8]
    assert jr.choice_indices(k, 16, 4, True).tolist() == [8, 8, 10, 1]
    kp = jr.mc_step_key(key, 0.0, partitionable=True)
    assert jr.choice_indices(kp, 16, 4, False, partitionable=True).tolist() == [3, 2, 10, 14]
    assert jr.choice_indices(kp, 16, 4, True, partitionable=True).tolist() == [15, 1, 5, 15]
    idx = jr.choice_indices(k, 16384, 512, False)
    assert idx[:8].tolist() == [4188, 1174, 292, 6731, 1285, 7462, 2385, 6037] and int(idx.sum()) == 4059257
    assert jr.shuffle_rounds(2**14) == 2 and jr.shuffle_rounds(2**25) == 3 and jr.shuffle_rounds(2**31) == 3


def test_golden_indices_regression():
    z = np.load(GOLDEN / "mc_indices.npz")
    key = jr.prng_key(42)
    for name in z.files:
        P, n, t, r, p = name.split("_")
        P, n, t, r, p = int(P[1:]), int(n[1:]), float(t[1:]), bool(int(r[1:])), bool(int(p[1:]))
        k = jr.mc_step_key(key, t, partitionable=p)
        assert np.array_equal(jr.choice_indices(k, P, n, r, partitionable=p), z[name]), name


def test_weighted_gumbel_orders_by_weight():
    # /root/reference/tests/test_kernels.py:83-92: key 42, t=0.0, identity weighting ->
    # rows come out ordered by descending weight (weak KAT; holds in both threefry modes)
    y0 = np.array([[1.0, 0.01], [2.0, 1.0], [3.0, 100.0], [4.0, 10000.0]], np.float32)
    for part in (False, True):
        mc = ref.MonteCarloKernel(None, weighting_function=lambda w: w, key=jr.prng_key(42), partitionable=part)
        out = ref.MonteCarloPartitioningKernel(4, mc)(0.0, y0, None, True)
        assert out.shape == (4, 1, 2)
        assert np.array_equal(out[:, 0, :], y0[np.argsort(y0[:, -1])[::-1]])


def test_hadamard_matches_scipy_and_is_degree3_exact():
    from itertools import product

    from scipy.linalg import hadamard

    for d in (1, 2, 3, 4, 5, 10, 64):
        k = ref.hadamard_point_count(d)
        H = hadamard(k // 2)[:, :d]
        M = ref.hadamard_points(d)
        assert np.array_equal(M, np.vstack((H, -H)))
        assert np.allclose(M.T @ M / k, np.eye(d))
    # /root/reference/tests/test_formulae.py:146-161: all monomials of degree <= 3 integrate exactly
    from scipy.special import factorial2

    for d in (1, 2, 3, 4):
        M, lam = ref.hadamard_points(d).astype(float), ref.hadamard_weights(d)
        for mono in product(range(4), repeat=d):
            if sum(mono) > 3:
                continue
            exact = np.prod([0.0 if m % 2 else (1.0 if m == 0 else factorial2(m - 1)) for m in mono])
            assert abs((lam * np.prod(M ** np.array(mono), axis=1)).sum() - exact) < 1e-12


def test_reference_stratified_expectations():
    # /root/reference/tests/test_kernels.py:95-133
    y0 = np.array([[-4.0, -3.0, -2.0, -1.0, 1.0, 2.0, 3.0, 4.0]], np.float32).T
    exp = np.array([[[-1.0], [1.0]], [[-2.0], [2.0]], [[-3.0], [3.0]], [[-4.0], [4.0]]], np.float32)
    k = ref.StratifiedPartitioningKernel(4)
    assert np.array_equal(k(0.0, y0), exp)
    y3 = np.tile(y0, 3)
    assert np.array_equal(k(0.0, y3), np.tile(exp, 3))
    yw = y3.copy(); yw[:, -1] = np.abs(yw[:, -1])
    ew = np.tile(exp, 3); ew[:, :, -1] = np.abs(ew[:, :, -1])
    assert np.array_equal(k(0.0, yw, None, True), ew)
    assert np.array_equal(ref.StratifiedPartitioningKernel(4, norm=None)(0.0, yw), yw.reshape(4, -1, 3))


def test_reference_tree_expectations():
    # /root/reference/tests/test_kernels.py:136-154
    y0 = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0], [7.0, 8.0], [9.0, 10.0], [11.0, 12.0]])
    for mode in ("KDTree", "BallTree"):
        v0 = ref.BinaryTreePartitioningKernel(3, mode)(0.0, y0)
        v1 = ref.BinaryTreePartitioningKernel(4, mode)(0.0, y0[:-2])
        assert v0.shape == (3, 2, 2) and v1.shape == (4, 1, 2)
        assert np.array_equal(np.unique(v0.reshape(-1, 2), axis=0), y0)


def test_solver_quirks_and_times():
    # Q1: inner interval is shifted one sub-step forward (_solvers.py:129-135)
    (a, b), = ref.substep_times(np.float32(0.0), np.float32(0.05), 1)
    assert a == np.float32(0.05) and b == np.float32(np.float32(0.05) + np.float32(0.05))
    times = ref.step_times(0.0, 0.05 * 512, 0.05, np.float32)
    assert times[0] == (np.float32(0.0), np.float32(0.05))
    assert times[1][1] == np.float32(np.float32(0.05) + np.float32(0.05))
    assert times[-1][1] == np.float32(0.05 * 512)
    t64 = ref.step_times(0.0, 0.05 * 1024, 0.05, np.float64)
    assert len(t64) == 1024
    # weights stay normalised through a weighted step (tests/test_solvers.py:106-125)
    d, n = 3, 16
    rng = np.random.default_rng(0)
    y0 = ref.pack_particles(rng.normal(size=(n, d)).astype(np.float32), np.ones(n, np.float32))
    mc = ref.MonteCarloKernel(n, key=jr.prng_key(1))
    y1 = ref.mcc_step(y0, 0.0, 0.05, lambda y: -y, ref.hadamard_points(d), ref.hadamard_weights(d),
                      np.float32(np.sqrt(2.0)), mc, n_substeps=2, weighted=True)
    assert y1.shape == (n, d + 1) and abs(float(y1[:, -1].sum()) - 1.0) < 1e-5


def test_randint64_structure():
    """x64 randint restatement (unpinned): with span = 2^34 the multiplier (2^32 mod span)^2 wraps to 0 in
    uint64, so the draw is the low 64-bit word mod span; small spans stay in range and are deterministic."""
    key = jr.mc_step_key(jr.prng_key(42), 0.25)
    for part in (False, True):
        k1, k2 = jr.split(key, 2, partitionable=part)
        lo = jr.random_bits(k2, 64, 100, partitionable=part)
        got = jr.randint64(key, 100, 0, 1 << 34, partitionable=part)
        assert np.array_equal(got, (lo % np.uint64(1 << 34)).astype(np.int64))
        small = jr.randint64(key, 1000, 0, 12345, partitionable=part)
        assert small.min() >= 0 and small.max() < 12345
        assert np.array_equal(small, jr.randint64(key, 1000, 0, 12345, partitionable=part))
        assert not np.array_equal(small, jr.randint(key, 1000, 0, 12345, partitionable=part))
        # 64-bit words are (word i << 32) | word n + i of one hash in the original mode
        if not part:
            bits = jr.random_bits(k2, 32, 200)
            assert np.array_equal(lo, (bits[:100].astype(np.uint64) << np.uint64(32)) | bits[100:].astype(np.uint64))


# ------------------------------------------------------------------ property tests of the restated jax.random chain
def test_oracle_random_properties():
    """Size-independent properties of the restatement, over random keys / sizes (hypothesis): a shuffle is a
    permutation, draws stay in range, a slice of the counter-based draws equals the slice of the full draw in
    partitionable mode, `split` and `fold_in` agree where JAX defines them to, and everything is a pure function
    of its arguments."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 2 ** 63 - 1), size=st.integers(1, 300), part=st.booleans(), data=st.integers(0, 2 ** 32 - 1))
    def run(seed, size, part, data):
        key = jr.fold_in(jr.prng_key(seed), data)
        perm = jr.shuffle_perm(key, size, partitionable=part)
        assert sorted(perm.tolist()) == list(range(size))
        assert np.array_equal(perm, jr.shuffle_perm(key, size, partitionable=part))           # deterministic
        n_draws = max(1, size // 3)
        idx = jr.choice_indices(key, size, n_draws, False, partitionable=part)
        assert np.array_equal(idx, perm[:n_draws])                                             # choice = permutation prefix
        rep = jr.choice_indices(key, size, 2 * size, True, partitionable=part)
        assert rep.min() >= 0 and rep.max() < size and rep.shape == (2 * size,)
        bits = jr.random_bits(key, 32, size, partitionable=part)
        assert bits.dtype == np.uint32 and bits.shape == (size,)
        if part:
            # counter based: element i depends on (key, i) only, so a longer draw extends a shorter one
            longer = jr.random_bits(key, 32, size + 7, partitionable=True)
            assert np.array_equal(longer[:size], bits)
            ks = jr.split(key, 3, partitionable=True)
            for i in range(3):
                assert tuple(int(v) for v in ks[i]) == tuple(int(v) for v in jr.fold_in(key, i))
        else:
            # original mode hashes iota(size) in two halves: the first ceil(size/2) words of a draw of `size`
            # are the x0 outputs of blocks (i, i + h)
            h = (size + 1) // 2
            x1 = np.arange(h, 2 * h, dtype=np.uint32)
            x1[x1 >= size] = 0
            o0, o1 = jr.threefry2x32(key, np.arange(h, dtype=np.uint32), x1)
            assert np.array_equal(bits[:h], o0) and np.array_equal(bits[h:], o1[:size - h])
        u = jr.uniform(key, size, partitionable=part)
        assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()

    run()


def test_median_split_oracle_matches_sklearn_node_sets():
    """The oracle's level-by-level restatement of sklearn's `_recursive_build` (the tree the reference builds in its host
    callback, mccube/_kernels/tree.py:44-58) puts into every sklearn leaf's range of `idx_array` exactly the points sklearn
    put there -- KDTree and BallTree alike, ragged counts included -- and levels = floor(log2(n / leaf)) is sklearn's depth
    or one more."""
    from sklearn.neighbors import BallTree, KDTree

    rng = np.random.default_rng(0)
    for n, d, leaf, tree in ((4096, 8, 64, KDTree), (1000, 5, 10, BallTree), (777, 3, 7, KDTree), (3000, 64, 100, BallTree),
                             (512, 40, 512, KDTree)):
        y = (rng.standard_normal((n, d)) * rng.uniform(0.5, 3.0, d)).astype(np.float32)
        levels = ref.median_split_levels(n, leaf)
        sk_levels = int(np.log2(max(1, (n - 1) // leaf)))
        assert levels in (sk_levels, sk_levels + 1)
        idx = ref.median_split_indices(y, levels)
        assert np.array_equal(np.sort(idx), np.arange(n))
        _, sk_idx, node_data, _ = tree(y.astype(np.float64), leaf).get_arrays()
        leaves = [(int(nd["idx_start"]), int(nd["idx_end"])) for nd in node_data if nd["is_leaf"]]
        assert sum(e - s for s, e in leaves) == n
        for s, e in leaves:
            assert set(idx[s:e].tolist()) == set(np.asarray(sk_idx[s:e]).tolist())
