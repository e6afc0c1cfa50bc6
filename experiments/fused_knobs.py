#!/usr/bin/env python
"""Config-3 step time under the experiment knobs of the fused rank+gather kernel.

usage: fused_knobs.py "K1=v K2=v" "K1=v ..." ...   (one timing per argument; "" = defaults)
"""
import json
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("MCCB200_BENCH_NO_REDIRECT", "1")
import bench  # noqa: E402
from mccube_b200 import _lib, diffrax_compat as dfx  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    a = SimpleNamespace(n_log2=int(os.environ.get("N_LOG2", "24")), d=int(os.environ.get("D", "64")), workload="box_prk",
                        part_size_log2=6, box_dims=4, box_bits=3)
    steps = int(os.environ.get("STEPS", "10"))
    times = dfx.step_times(0.0, bench.DT0 * (steps + 8), bench.DT0, np.float32)
    knobs_seen = set()
    for spec in (sys.argv[1:] or [""]):
        kv = dict(item.split("=") for item in spec.split())
        for k in knobs_seen:
            os.environ.pop(k, None)
        for k, v in kv.items():
            os.environ[k] = v
            knobs_seen.add(k)
        _lib.release_workspace()
        terms, solver, y0 = bench.build_problem(a, dev)
        ms, path = bench.time_steps(solver, terms, y0, times, steps, warmup=4)
        _lib.profile_enable(True)
        state, y = None, y0
        for i in range(6):
            y, _, _, state, _ = solver.step(terms, times[i][0], times[i][1], y, None, state, False)
        torch.cuda.synchronize()
        ph = {}
        for name, t in _lib.profile_read():
            ph.setdefault(name, []).append(t)
        _lib.profile_enable(False)
        print(json.dumps({"knobs": spec, "ms_per_step": ms, "path": path,
                          "phase_ms": {k: float(np.median(v[2:])) for k, v in ph.items()}}), flush=True)
        del y, y0, state


if __name__ == "__main__":
    main()
