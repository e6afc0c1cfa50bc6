#!/usr/bin/env python
"""Device median-split partitioner (csrc/kd_split.cu): ms per call at config-3 size, and sklearn's host build beside it
at a size it finishes in seconds."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for n_log2, d, leaf in ((20, 64, 64), (24, 64, 64)):
    n = 1 << n_log2
    y = torch.randn((n, d), device=dev)
    levels = _lib.kd_partition_levels(n, leaf)
    _lib.kd_partition(y, d, levels)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        idx = _lib.kd_partition(y, d, levels)
    b.record()
    torch.cuda.synchronize()
    rec = {"levels": levels, "device_ms": a.elapsed_time(b) / reps}
    _lib.profile_enable(True)
    _lib.kd_partition(y, d, levels)
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    if n_log2 <= 20:
        from sklearn.neighbors import KDTree

        host = y.cpu().numpy()
        t0 = time.perf_counter()
        KDTree(host, leaf)
        rec["sklearn_host_ms"] = (time.perf_counter() - t0) * 1e3
    out[f"n=2^{n_log2} d={d} leaf={leaf}"] = rec
    del y, idx
print(json.dumps(out))
