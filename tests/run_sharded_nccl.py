"""Multi-GPU check of the sharded global MonteCarloKernel resample (run under torchrun, one rank
per GPU, NCCL):  python -m torch.distributed.run --nproc-per-node 2 tests/run_sharded_nccl.py
Asserts that the concatenated shards equal the single-device step bit for bit (at a size rank 0
can hold), then times the sharded step at --n-log2 particles per rank."""
import argparse
import json
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mccube_b200 as mccube  # noqa: E402
from mccube_b200 import diffrax_compat as dfx  # noqa: E402
from mccube_b200._distributed import PeerShards, ShardedBoxStep, ShardedMonteCarloStep  # noqa: E402


def terms_for(d, dev):
    return mccube.MCCTerm(
        dfx.ODETerm(mccube.GaussianDiagDrift(np.full(d, 2.0), 3.0, device=dev)),
        dfx.WeaklyDiagonalControlTerm(lambda t, y, args: math.sqrt(2.0),
                                      mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-log2", type=int, default=20)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--replace", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    replace = bool(a.replace)

    # ---- parity at a small size: every rank builds the same global state
    n_loc, d = 4096, 10
    n_tot = n_loc * world
    gen = torch.Generator(device=dev); gen.manual_seed(7)
    y = torch.randn((n_tot, d), generator=gen, device=dev)
    terms = terms_for(d, dev)
    t0, t1 = np.float32(0.25), np.float32(0.3)
    for rep in (False, True):
        want, *_ = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, rep, key=42)).step(
            terms, t0, t1, y, None, None, False)
        sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, rep, key=42)), terms, world)
        got = sh.step(t0, t1, y[rank * n_loc:(rank + 1) * n_loc].contiguous(), rank)
        assert torch.equal(got, want[rank * n_loc:(rank + 1) * n_loc]), f"rank {rank}: sharded != single (replace={rep})"
    # fused expand + exchange over peer memory (no collective in the data path): same shards, two steps
    # so that both ping-pong buffers are exercised
    for rep in (False, True):
        single = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, rep, key=42))
        want1, *_ = single.step(terms, t0, t1, y, None, None, False)
        want2, *_ = single.step(terms, t1, np.float32(0.35), want1, None, None, False)
        sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, rep, key=42)), terms, world)
        shards = PeerShards(n_loc, d, rank, world, dev)
        got1 = sh.step_p2p(t0, t1, y[rank * n_loc:(rank + 1) * n_loc].contiguous(), rank, shards)
        assert torch.equal(got1, want1[rank * n_loc:(rank + 1) * n_loc]), f"rank {rank}: p2p step 1 != single (replace={rep})"
        got2 = sh.step_p2p(t1, np.float32(0.35), got1, rank, shards)
        assert torch.equal(got2, want2[rank * n_loc:(rank + 1) * n_loc]), f"rank {rank}: p2p step 2 != single (replace={rep})"
        del got1, got2
        shards.close()
    # pull form (with_replacement=True): indices of the own slots only, parents read from the owners' shards
    single = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42))
    want1, *_ = single.step(terms, t0, t1, y, None, None, False)
    want2, *_ = single.step(terms, t1, np.float32(0.35), want1, None, None, False)
    sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42)), terms, world)
    shards = PeerShards(n_loc, d, rank, world, dev)
    shards.load(y[rank * n_loc:(rank + 1) * n_loc])
    got1 = sh.step_pull(t0, t1, rank, shards)
    assert torch.equal(got1, want1[rank * n_loc:(rank + 1) * n_loc]), f"rank {rank}: pull step 1 != single"
    got2 = sh.step_pull(t1, np.float32(0.35), rank, shards)
    assert torch.equal(got2, want2[rank * n_loc:(rank + 1) * n_loc]), f"rank {rank}: pull step 2 != single"
    del got1, got2
    shards.close()
    # owner-computes variant: rows equal the single-device result through the exported slot map
    want, *_ = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42)).step(
        terms, t0, t1, y, None, None, False)
    sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42)), terms, world)
    got, slot_map = sh.step_owner(t0, t1, y[rank * n_loc:(rank + 1) * n_loc].contiguous(), rank)
    assert torch.equal(got, want[slot_map.to(torch.int64) & 0xFFFFFFFF]), f"rank {rank}: owner-computes mismatch"
    dist.barrier()
    # global Box + PRK partition over the shards (two all-reduces of 8 floats, one all-gather of the key counts,
    # fused expand + scatter over peer memory): three steps, shards equal the single-device trajectory
    n_box, d_box = 4096, 16

    def box_solver(n_all):
        return mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
            mccube.BoxPartitionKernel(n_all // 16, dims=(4, 5, 6, 7), bits=2), mccube.MonteCarloKernel(16, key=7)))

    gen.manual_seed(13)
    yb = torch.randn((n_box * world, d_box), generator=gen, device=dev) * 1.5 + 1.0
    terms_b = terms_for(d_box, dev)
    shb = ShardedBoxStep(box_solver(n_box * world), terms_b, world)
    shards = PeerShards(n_box, d_box, rank, world, dev)
    shards.load(yb[rank * n_box:(rank + 1) * n_box])
    single, want = box_solver(n_box * world), yb
    for (ta, tb) in dfx.step_times(0.0, 0.15, 0.05, np.float32):
        want, *_ = single.step(terms_b, ta, tb, want, None, None, False)
        got = shb.step(ta, tb, rank, shards)
        assert torch.equal(got, want[rank * n_box:(rank + 1) * n_box]), f"rank {rank}: sharded Box step != single at t={ta}"
    del got
    shards.close()
    dist.barrier()

    # ---- timing at n_loc = 2^n_log2 per rank
    n_loc, d = 1 << a.n_log2, a.d
    n_tot = n_loc * world
    terms = terms_for(d, dev)
    gen.manual_seed(11 + rank)
    ys = torch.randn((n_loc, d), generator=gen, device=dev)
    sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, replace, key=42)), terms, world)
    times = dfx.step_times(0.0, 0.05 * (a.steps + 4), 0.05)
    for i in range(3):
        ys = sh.step(times[i][0], times[i][1], ys, rank)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + a.steps):
        ys = sh.step(times[i][0], times[i][1], ys, rank)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # owner-computes timing
    gen.manual_seed(11 + rank)
    yo = torch.randn((n_loc, d), generator=gen, device=dev)
    for i in range(3):
        yo, _ = sh.step_owner(times[i][0], times[i][1], yo, rank)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for i in range(3, 3 + a.steps):
        yo, _ = sh.step_owner(times[i][0], times[i][1], yo, rank)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms_o = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms_o, op=dist.ReduceOp.MAX)
    # fused p2p timing
    gen.manual_seed(11 + rank)
    yp = torch.randn((n_loc, d), generator=gen, device=dev)
    shards = PeerShards(n_loc, d, rank, world, dev)
    for i in range(3):
        yp = sh.step_p2p(times[i][0], times[i][1], yp, rank, shards)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for i in range(3, 3 + a.steps):
        yp = sh.step_p2p(times[i][0], times[i][1], yp, rank, shards)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms_p = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms_p, op=dist.ReduceOp.MAX)
    # pull timing
    shp = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42)), terms, world)
    gen.manual_seed(11 + rank)
    shards.load(torch.randn((n_loc, d), generator=gen, device=dev))
    for i in range(3):
        shp.step_pull(times[i][0], times[i][1], rank, shards)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for i in range(3, 3 + a.steps):
        shp.step_pull(times[i][0], times[i][1], rank, shards)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms_q = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms_q, op=dist.ReduceOp.MAX)
    # global Box + PRK partition timing (2^6 outputs per partition of 2^13 children, as in the headline config)
    shards.close()
    shb = ShardedBoxStep(mccube.MCCSolver(dfx.Euler(), mccube.PartitioningRecombinationKernel(
        mccube.BoxPartitionKernel(n_tot // 64, dims=(0, 1, 2, 3), bits=3), mccube.MonteCarloKernel(64, key=42))), terms, world)
    shards = PeerShards(n_loc, d, rank, world, dev)
    gen.manual_seed(11 + rank)
    shards.load(torch.randn((n_loc, d), generator=gen, device=dev) * 1.5 + 2.0)
    for i in range(3):
        shb.step(times[i][0], times[i][1], rank, shards)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for i in range(3, 3 + a.steps):
        shb.step(times[i][0], times[i][1], rank, shards)
    e1.record()
    torch.cuda.synchronize(); dist.barrier()
    ms_b = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms_b, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"variant": "global Box + PRK partition over the shards (fused expand + peer scatter), bit-exact shards",
                          "n_gpus": world, "n_per_gpu": n_loc, "d": d, "ms_per_step": float(ms_b) / a.steps,
                          "particle_steps_per_sec": n_tot * a.steps / (float(ms_b) * 1e-3)}))
    if rank == 0:
        print(json.dumps({"variant": "pull: own index slice + gather from the owners' shards over peer memory, bit-exact shards",
                          "n_gpus": world, "n_per_gpu": n_loc, "d": d, "ms_per_step": float(ms_q) / a.steps,
                          "particle_steps_per_sec": n_tot * a.steps / (float(ms_q) * 1e-3)}))
    if rank == 0:
        print(json.dumps({"variant": "fused expand + peer-memory scatter (no collective in the data path), bit-exact shards",
                          "n_gpus": world, "n_per_gpu": n_loc, "d": d, "ms_per_step": float(ms_p) / a.steps,
                          "particle_steps_per_sec": n_tot * a.steps / (float(ms_p) * 1e-3)}))
        print(json.dumps({"variant": "owner-computes (imbalance exchange only)", "n_gpus": world, "n_per_gpu": n_loc, "d": d,
                          "ms_per_step": float(ms_o) / a.steps,
                          "particle_steps_per_sec": n_tot * a.steps / (float(ms_o) * 1e-3)}))
        print(json.dumps({"check": "sharded == single-device, bit-exact (replace False and True)", "n_gpus": world,
                          "n_per_gpu": n_loc, "d": d, "replace": replace, "ms_per_step": float(ms) / a.steps,
                          "particle_steps_per_sec": n_tot * a.steps / (float(ms) * 1e-3),
                          "exchange_bytes_per_rank_per_step": int(n_loc * d * 4 * (world - 1) / world)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
