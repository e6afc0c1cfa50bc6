"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges and the moment
all-reduce algebra.  The per-shard partial moments are produced with NumPy here -- on a GPU box
they come from mccb200_moments_sum / _m2 (tests/test_gpu_parity.py::test_moments_and_w2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mccube_b200._distributed import global_cov, global_mean, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.default_rng(0)
        x = rs.normal(size=(n, d)) * rs.uniform(0.5, 2.0, size=d) + 3.0
        w = rs.uniform(0.1, 1.0, size=n)
        lo, hi = shard_range(n, rank, world)
        xs, ws = x[lo:hi], w[lo:hi]
        sums = torch.tensor(np.concatenate([(ws[:, None] * xs).sum(0), [ws.sum()]]))
        mean, wtot = global_mean(sums)
        centre = mean.to(torch.float32)
        c64 = centre.double().numpy()
        m2 = torch.tensor(((xs - c64) * ws[:, None]).T @ (xs - c64))
        w2 = torch.tensor((ws * ws).sum())
        cov = global_cov(m2, sums, centre, mean, wtot, ddof=0, local_w2=w2)
        cov1 = global_cov(m2, sums, centre, mean, wtot, ddof=1, local_w2=w2)     # weighted, ddof = 1 (jnp.cov default)
        # normalised weights (what pack_particles produces): sum(w) = 1, so `wtot - ddof` would divide by zero
        wn = w / w.sum()
        wsn = wn[lo:hi]
        sums_n = torch.tensor(np.concatenate([(wsn[:, None] * xs).sum(0), [wsn.sum()]]))
        mean_n, wtot_n = global_mean(sums_n)
        m2n = torch.tensor(((xs - c64) * wsn[:, None]).T @ (xs - c64))
        cov1n = global_cov(m2n, sums_n, centre, mean_n, wtot_n, ddof=1, local_w2=torch.tensor((wsn * wsn).sum()))
        if rank == 0:
            out["mean"] = mean.numpy()
            out["cov"] = cov.numpy()
            out["cov1"] = cov1.numpy()
            out["cov1n"] = cov1n.numpy()
            out["want_mean"] = np.average(x, axis=0, weights=w)
            out["want_cov"] = np.cov(x, rowvar=False, ddof=0, aweights=w)
            out["want_cov1"] = np.cov(x, rowvar=False, ddof=1, aweights=w)
            out["want_cov1n"] = np.cov(x, rowvar=False, ddof=1, aweights=wn)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n, world in [(10, 2), (7, 3), (1 << 20, 8), (5, 8), (0, 2)]:
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_global_moments_two_ranks_gloo():
    world, n, d = 2, 1001, 5
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, d, out), nprocs=world, join=True)
        np.testing.assert_allclose(out["mean"], out["want_mean"], rtol=1e-12)
        np.testing.assert_allclose(out["cov"], out["want_cov"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out["cov1"], out["want_cov1"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out["cov1n"], out["want_cov1n"], rtol=1e-9, atol=1e-12)


def test_plan_blocks_matches_bruteforce():
    from mccube_b200._distributed import plan_blocks

    rs = np.random.default_rng(3)
    G, n_loc, k = 4, 50, 8
    n_tot = G * n_loc
    idx = rs.integers(0, n_tot * k, size=n_tot)
    owner = idx // (n_loc * k)
    dest = np.arange(n_tot) // n_loc
    counts = np.zeros((G, G), np.int64)
    np.add.at(counts, (owner, dest), 1)
    order = np.argsort(owner, kind="stable")
    for r in range(G):
        send, recv = plan_blocks(counts, r)
        mine = order[send["start"]:send["start"] + send["count"]]
        assert np.array_equal(mine, np.nonzero(owner == r)[0])
        assert send["splits"] == [int(((owner == r) & (dest == q)).sum()) for q in range(G)]
        got = np.concatenate([order[s:s + c] for s, c in recv["blocks"]])
        want = np.concatenate([np.nonzero((owner == o) & (dest == r))[0] for o in range(G)])
        assert np.array_equal(got, want) and sorted(got) == list(range(r * n_loc, (r + 1) * n_loc))


def _exchange_worker(rank, world, port, out):
    from mccube_b200._distributed import exchange_rows, plan_blocks

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.default_rng(5)
        n_loc, k, d = 64, 4, 3
        n_tot = world * n_loc
        idx = rs.integers(0, n_tot * k, size=n_tot)
        children = rs.normal(size=(n_tot * k, d)).astype(np.float32)     # stand-in for the expansion
        owner, dest = idx // (n_loc * k), np.arange(n_tot) // n_loc
        counts = np.zeros((world, world), np.int64)
        np.add.at(counts, (owner, dest), 1)
        order = np.argsort(owner, kind="stable")
        send, recv = plan_blocks(counts, rank)
        mine = order[send["start"]:send["start"] + send["count"]]
        send_buf = torch.tensor(children[idx[mine]])                       # what expand_select produces locally
        recv_buf = exchange_rows(send_buf, send["splits"], recv["splits"])
        slots = np.concatenate([order[s:s + c] for s, c in recv["blocks"]])
        y1 = np.empty((n_loc, d), np.float32)
        y1[slots - rank * n_loc] = recv_buf.numpy()
        out[rank] = (y1, children[idx[rank * n_loc:(rank + 1) * n_loc]])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_resample_exchange_two_ranks_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_exchange_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        for r in range(world):
            got, want = out[r]
            assert np.array_equal(got, want)


def test_plan_rebalance_is_consistent():
    from mccube_b200._distributed import plan_rebalance

    rs = np.random.default_rng(9)
    for G, n_loc in [(2, 100), (8, 1000), (5, 7)]:
        for _ in range(20):
            per = rs.multinomial(G * n_loc, np.ones(G) / G)
            tr = plan_rebalance(per, n_loc)
            assert np.array_equal(per - tr.sum(1) + tr.sum(0), np.full(G, n_loc))   # everybody ends with n_loc
            assert np.array_equal(tr.sum(1), np.maximum(per - n_loc, 0))            # only the surplus moves
            assert (np.diag(tr) == 0).all()


def test_global_key_starts_match_a_global_stable_sort():
    """Host logic of the sharded Box step: from the per-rank per-key child counts, rank r's first child of key q
    sits at (children of smaller keys on all ranks) + (children of key q on lower ranks) in the global stable
    order -- checked against an explicit stable sort of (key, rank-major child id), including 32-bit wrap."""
    import torch

    from mccube_b200._distributed import ShardedBoxStep

    rs = np.random.default_rng(0)
    G, n_keys, per_rank = 3, 17, 500
    keys = [rs.integers(0, n_keys, size=per_rank) for _ in range(G)]
    totals = torch.tensor(np.stack([np.bincount(k_, minlength=n_keys) for k_ in keys]), dtype=torch.int64)
    allk = np.concatenate(keys)
    order = np.argsort(allk, kind="stable")                       # global sorted position -> global child id
    pos = np.empty_like(order); pos[order] = np.arange(order.size)
    for r in range(G):
        got = ShardedBoxStep.key_starts_for(r, totals).numpy().astype(np.int64) & 0xFFFFFFFF
        for q in range(n_keys):
            mine = np.nonzero(keys[r] == q)[0]
            if mine.size:
                assert got[q] == pos[r * per_rank + mine[0]], (r, q)
    # positions are taken modulo 2^32 (uint32 bit patterns in an int32 tensor)
    big = torch.tensor([[2 ** 31, 2 ** 31 - 5, 7], [1, 2, 3]], dtype=torch.int64)
    got = ShardedBoxStep.key_starts_for(1, big)
    assert got.dtype == torch.int32
    assert (got.to(torch.int64) & 0xFFFFFFFF).tolist() == [2 ** 31, (2 ** 31 + 1 + 2 ** 31 - 5) % 2 ** 32, (2 ** 32 - 2 + 7) % 2 ** 32]


def _box_worker(rank, world, port, out):
    """The two collectives of ShardedBoxStep on CPU tensors over gloo: global min / max from one all-reduce and
    the per-rank key-count matrix, then every rank's global key starts."""
    from mccube_b200._distributed import ShardedBoxStep

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.default_rng(100 + rank)
        lo, hi = rs.normal(size=3).astype(np.float32) - 2.0, rs.normal(size=3).astype(np.float32) + 2.0
        mm = ShardedBoxStep.reduce_minmax(torch.tensor(np.concatenate([lo, hi])))
        totals = torch.tensor(rs.integers(0, 50, size=11), dtype=torch.int64)
        totals_all = ShardedBoxStep.gather_totals(totals, world)
        starts = ShardedBoxStep.key_starts_for(rank, totals_all)
        out[rank] = (mm.numpy(), totals.numpy(), starts.numpy(), lo, hi)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_box_collectives_two_ranks_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_box_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        got = {r: out[r] for r in range(world)}
    lo = np.minimum(got[0][3], got[1][3])
    hi = np.maximum(got[0][4], got[1][4])
    for r in range(world):
        assert np.array_equal(got[r][0], np.concatenate([lo, hi]))              # exact: min / max only
    t0, t1 = got[0][1], got[1][1]
    per_key = t0 + t1
    base = np.cumsum(per_key) - per_key
    assert np.array_equal(got[0][2].astype(np.int64), base)                     # rank 0 starts each key's run
    assert np.array_equal(got[1][2].astype(np.int64), base + t0)                # rank 1 follows rank 0's children
