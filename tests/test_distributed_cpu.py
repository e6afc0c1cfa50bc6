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
        cov = global_cov(m2, sums, centre, mean, wtot, ddof=0)
        if rank == 0:
            out["mean"] = mean.numpy()
            out["cov"] = cov.numpy()
            out["want_mean"] = np.average(x, axis=0, weights=w)
            out["want_cov"] = np.cov(x, rowvar=False, ddof=0, aweights=w)
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
