#!/usr/bin/env python
"""permutation[:2^20] of 2^25 inputs (config 2's index draw): selection shuffle vs full radix-sort shuffle, ms per draw.

With ONCE=1 runs three draws only (for an ncu launch list)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
rng = _lib.make_rng((0, 42), np.float32(0.35))
P, count = 1 << int(os.environ.get("P_LOG2", "25")), 1 << int(os.environ.get("COUNT_LOG2", "20"))


def draw(reps):
    for _ in range(3):
        idx = _lib.mc_indices(rng, P, count, False, dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        idx = _lib.mc_indices(rng, P, count, False, dev)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, idx


if os.environ.get("ONCE"):
    print(int(draw(1)[1][:4].sum()))
else:
    out = {}
    ref = None
    for name, flag in (("select", "1"), ("radix", "0")):
        os.environ["MCCB200_SHUFFLE_SELECT"] = flag
        ms, idx = draw(20)
        out[name + "_ms"] = ms
        if ref is None:
            ref = idx.clone()
        else:
            out["identical"] = bool(torch.equal(ref, idx))
    print(json.dumps(out))
