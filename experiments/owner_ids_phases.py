#!/usr/bin/env python
"""Local phases of step_owner_ids at config-5 size (2^24 rows per rank of 8), timed on ONE GPU:
index slice draw (x64), stable bucketing by owner, and the local gather."""
import json
import math
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import mccube_b200 as mccube  # noqa: E402
from mccube_b200 import _lib, diffrax_compat as dfx  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    dev = torch.device("cuda:0")
    G, n_loc, d, k = 8, 1 << 24, 64, 128
    n_tot = G * n_loc
    kern = mccube.MonteCarloKernel(n_tot, True, key=42, x64=True)
    t0 = np.float32(0.25)
    res = {}
    res["draw_x64_ms"], idx = timed(lambda: kern.indices_slice(t0, n_tot * k, n_tot, 3 * n_loc, n_loc, dev))
    res["bucket_ms"], (slots, ids, counts) = timed(lambda: _lib.shard_bucket_ids(idx, 3 * n_loc, n_loc, k, G))
    res["counts"] = counts.cpu().tolist()
    y = torch.randn((n_loc, d), device=dev)
    drift = _lib.make_drift(2, mean=torch.full((d,), 2.0, device=dev), inv_var=torch.full((d,), 1 / 3.0, device=dev))
    cub = _lib.make_cubature(k, scale=np.float32(math.sqrt(2.0) * math.sqrt(0.05)))

    out = torch.empty_like(y)
    res["gather_ms"], _ = timed(lambda: _lib.expand_select(y, d, False, drift, np.float32(0.05), cub, ids, n_out=n_loc, out=out))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
