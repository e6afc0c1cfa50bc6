#!/usr/bin/env python
"""Two calls of the device median-split partitioner at 2^22 x 64 (for an ncu launch list)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib
dev = torch.device("cuda:0")
n, d, leaf = 1 << 22, 64, 64
y = torch.randn((n, d), device=dev)
lv = _lib.kd_partition_levels(n, leaf)
for _ in range(2):
    idx = _lib.kd_partition(y, d, lv)
torch.cuda.synchronize()
print(lv, int(idx[:3].sum()))
