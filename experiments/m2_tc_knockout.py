#!/usr/bin/env python
"""Knock-out timings of the tcgen05 second-moment kernel (results are wrong under debug flags): which role bounds it."""
import json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib
dev = torch.device("cuda:0")
n = 1 << 24
y = torch.randn((n, 64), device=dev) * 1.7 + 2.0
c = y[:4096].mean(0).contiguous()
out = {}
for name, flag in (("full", "0"), ("no_mma", "2"), ("no_global_loads", "3"), ("no_fp64_accumulation", "4")):
    os.environ["MCCB200_M2_TC_DEBUG"] = flag
    _lib.second_moment(y, 64, c); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        _lib.second_moment(y, 64, c)
    b.record(); torch.cuda.synchronize()
    out[name] = a.elapsed_time(b) / 5
print(json.dumps(out))
