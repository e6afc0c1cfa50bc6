#!/usr/bin/env python
"""Can the bookkeeping passes of step t+1 run UNDER the HBM-bound gather of step t?

Config 3 (n = 2^24, d = 64, Box + PRK), steady state.  Times, with CUDA events,
  rank alone / expand alone / both on two streams
for several residency splits (MCCB200_RANK_CTAS, MCCB200_EXPAND_CTAS, MCCB200_EXPAND_SMEM_PAD).
Output: one JSON line per configuration.
"""
import json
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mccube_b200 as mccube  # noqa: E402
from mccube_b200 import _lib, diffrax_compat as dfx  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    n_log2 = int(os.environ.get("N_LOG2", "24"))
    n, d, per, bits = 1 << n_log2, 64, 64, 3
    dims = (0, 1, 2, 3)
    drift_fn = mccube.GaussianDiagDrift(np.full(d, 2.0), 3.0, device=dev)
    terms = mccube.MCCTerm(
        dfx.ODETerm(drift_fn),
        dfx.WeaklyDiagonalControlTerm(lambda t, y, args: math.sqrt(2.0),
                                      mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d)))))
    mc = mccube.MonteCarloKernel(per, key=42)
    kernel = mccube.PartitioningRecombinationKernel(mccube.BoxPartitionKernel(n // per, dims=dims, bits=bits), mc)
    solver = mccube.MCCSolver(dfx.Euler(), kernel)
    times = dfx.step_times(0.0, 0.05 * 16, 0.05, np.float32)
    y = torch.randn((n, d), device=dev)
    state = None
    for i in range(4):
        y, _, _, state, _ = solver.step(terms, times[i][0], times[i][1], y, None, state, False)
    torch.cuda.synchronize()
    # pieces of one more step, by hand
    t0, t1 = times[4]
    drift, cub, dt, k, _ = solver._describe(terms, y, *solver.substep_times(t0, t1)[0], None)
    in_per_part = per * k
    keycols = state["mccb200_keycols"]
    tables = _lib.box_prk_tables(k, dims, bits, dev)
    idx_local = mc.indices(t0, in_per_part, per, dev)
    sel = _lib.box_prk_selection(idx_local, in_per_part)
    bounds = _lib.box_child_bounds(y, d, drift, dt, cub, dims, bits, keycols=keycols)
    y1 = torch.empty_like(y)
    idx_gather = torch.randint(0, n * k, (n,), device=dev, dtype=torch.int64).to(torch.int32)  # any valid child ids
    idx_slots = torch.empty(n, dtype=torch.int32, device=dev)
    s_rank, s_exp = torch.cuda.Stream(), torch.cuda.Stream()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def run(which, reps=5):
        out = []
        for _ in range(reps):
            # the count pass rebuilds the counter table the rank pass consumes (atomics mutate it)
            ws = _lib.box_prk_count(y, d, drift, dt, cub, dims, bits, bounds, in_per_part, tables, keycols=keycols)
            ks = _lib.box_prk_local_key_starts(ws, n, k, len(dims), bits, in_per_part, per)
            torch.cuda.synchronize()
            a, b, c, e = ev(), ev(), ev(), ev()
            s_rank.wait_stream(torch.cuda.current_stream())
            s_exp.wait_stream(torch.cuda.current_stream())
            if which in ("rank", "both"):
                with torch.cuda.stream(s_rank):
                    a.record()
                    _lib.box_prk_rank_global(y, d, drift, dt, cub, dims, bits, per, in_per_part, tables, sel, ws, ks,
                                             idx_slots)
                    b.record()
            if which in ("expand", "both"):
                with torch.cuda.stream(s_exp):
                    c.record()
                    _lib.expand_select(y, d, False, drift, dt, cub, idx_gather, out=y1)
                    e.record()
            torch.cuda.synchronize()
            r = {}
            if which in ("rank", "both"):
                r["rank_ms"] = a.elapsed_time(b)
            if which in ("expand", "both"):
                r["expand_ms"] = c.elapsed_time(e)
            if which == "both":
                r["span_ms"] = max(a.elapsed_time(b), a.elapsed_time(e), c.elapsed_time(e), c.elapsed_time(b))
            out.append(r)
        return {k2: float(np.median([o[k2] for o in out])) for k2 in out[0]}

    combos = [(4, 8, 0), (3, 8, 0), (2, 8, 0), (2, 2, 0), (2, 2, 100), (3, 1, 0), (3, 2, 60), (2, 3, 60), (1, 3, 60)]
    for rank_ctas, exp_ctas, pad_kb in combos:
        os.environ["MCCB200_RANK_CTAS"] = str(rank_ctas)
        os.environ["MCCB200_EXPAND_CTAS"] = str(exp_ctas)
        os.environ["MCCB200_EXPAND_SMEM_PAD"] = str(pad_kb * 1024)
        _lib.release_workspace()
        res = {"rank_ctas": rank_ctas, "expand_ctas": exp_ctas, "expand_pad_kb": pad_kb}
        for which in ("rank", "expand", "both"):
            res[which] = run(which)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
