#!/usr/bin/env python
"""tcgen05 second moment (moments_tc.cu) against fp64 numpy and against the mma.sync kernel, plus timing at n = 2^24."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
out = {}
rs = np.random.default_rng(0)
for n in (64, 1000, 20000, 300001):
    p = (rs.normal(size=(n, 64)) * rs.uniform(0.5, 2, size=64) + rs.normal(size=64) * 3).astype(np.float32)
    c = p.astype(np.float64).mean(0).astype(np.float32)
    pt, ct = torch.from_numpy(p).to(dev), torch.from_numpy(c).to(dev)
    os.environ["MCCB200_M2_TC"] = "2"
    got2 = _lib.second_moment(pt, 64, ct).cpu().numpy()
    os.environ["MCCB200_M2_TC"] = "1"
    got = _lib.second_moment(pt, 64, ct).cpu().numpy()
    os.environ["MCCB200_M2_TC"] = "0"
    old = _lib.second_moment(pt, 64, ct).cpu().numpy()
    x = p.astype(np.float64) - c.astype(np.float64)
    want = x.T @ x
    scale = np.abs(want).max()
    out[f"n={n}"] = {"tc_err": float(np.abs(got - want).max() / scale), "tc_a_in_tmem_err": float(np.abs(got2 - want).max() / scale), "mma_sync_err": float(np.abs(old - want).max() / scale)}
n = 1 << 24
y = torch.randn((n, 64), device=dev) * 1.7 + 2.0
c = y[:4096].mean(0).contiguous()
for flag in ("2", "1", "0"):
    os.environ["MCCB200_M2_TC"] = flag
    _lib.second_moment(y, 64, c); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        _lib.second_moment(y, 64, c)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out[{"2": "tcgen05_a_in_tmem", "1": "tcgen05", "0": "mma.sync"}[flag]] = {"ms": ms, "read_gbs": n * 64 * 4 / ms / 1e6}
print(json.dumps(out))
