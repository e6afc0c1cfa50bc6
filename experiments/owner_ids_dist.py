#!/usr/bin/env python
"""step_owner_ids under torchrun: per-phase wall/GPU times at n_loc = 2^24 rows per rank (config-5 shard at 8 GPUs)."""
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mccube_b200 as mccube  # noqa: E402
from mccube_b200 import diffrax_compat as dfx  # noqa: E402
from mccube_b200._distributed import ShardedMonteCarloStep, exchange_rows  # noqa: E402
from mccube_b200 import _lib  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    d, n_loc = 64, 1 << int(os.environ.get("N_LOC_LOG2", "24"))
    n_tot = n_loc * world
    terms = mccube.MCCTerm(
        dfx.ODETerm(mccube.GaussianDiagDrift(np.full(d, 2.0), 3.0, device=dev)),
        dfx.WeaklyDiagonalControlTerm(lambda t, y, args: math.sqrt(2.0),
                                      mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d)))))
    sh = ShardedMonteCarloStep(mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n_tot, True, key=42, x64=True)), terms, world)
    times = dfx.step_times(0.0, 0.05 * 40, 0.05, np.float32)
    y = torch.randn((n_loc, d), device=dev)
    res = {}
    for prefetch, dma in ((False, False), (True, False), (True, True)):
        yy = y
        for i in range(3):
            yy, _ = sh.step_owner_ids(times[i][0], times[i][1], yy, rank, prefetch=prefetch, dma=dma)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tic = time.perf_counter()
        e0.record()
        for i in range(3, 13):
            yy, _ = sh.step_owner_ids(times[i][0], times[i][1], yy, rank, prefetch=prefetch, dma=dma)
        e1.record()
        torch.cuda.synchronize()
        res[f"prefetch={prefetch},dma={dma}"] = {"gpu_ms_per_step": e0.elapsed_time(e1) / 10, "wall_ms_per_step": (time.perf_counter() - tic) * 100}
        sh._ids_ready = None
        torch.cuda.synchronize()
    # timeline of three steady-state steps in the copy-engine mode (ms since the first event)
    yy = y
    for i in range(3):
        yy, _ = sh.step_owner_ids(times[i][0], times[i][1], yy, rank, prefetch=True, dma=True)
    sh._trace = []
    for i in range(3, 6):
        yy, _ = sh.step_owner_ids(times[i][0], times[i][1], yy, rank, prefetch=True, dma=True)
    torch.cuda.synchronize()
    base = sh._trace[0][1]
    res["timeline_ms"] = [(lab, round(base.elapsed_time(ev), 3)) for lab, ev in sh._trace]
    sh._trace = None
    sh._ids_ready = None
    sh.close()
    # phases, no prefetch
    def ev():
        return torch.cuda.Event(enable_timing=True)
    marks = []
    t0, t1 = times[20]
    torch.cuda.synchronize(); dist.barrier()
    def mark(name):
        e = ev(); e.record(); marks.append((name, e, time.perf_counter()))
    mark("start")
    pairs, counts_dev = sh.owner_ids_bucket(t0, n_loc, rank, dev, packed=True); mark("draw+bucket")
    allc = [torch.empty_like(counts_dev) for _ in range(world)]
    dist.all_gather(allc, counts_dev); counts = torch.stack(allc).cpu().numpy(); mark("all_gather+host sync")
    recv = exchange_rows(pairs, [int(c) for c in counts[rank]], [int(c) for c in counts[:, rank]]); mark("all_to_all ids")
    out = sh.owner_ids_materialise(t0, t1, y, rank, recv, counts); mark("materialise (gather)")
    torch.cuda.synchronize()
    res["phases_gpu_ms"] = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
    res["phases_host_ms"] = {marks[i][0]: (marks[i][2] - marks[i - 1][2]) * 1e3 for i in range(1, len(marks))}
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
