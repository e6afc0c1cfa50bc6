#!/usr/bin/env python
"""MonteCarloKernel(with_replacement=False) at config-3 size: permutation[:2^24] of 2^31 child ids, by rank selection
(2^28 buckets) and by three sort rounds; the two index vectors must be identical."""
import json, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib
dev = torch.device("cuda:0")
rng = _lib.make_rng((0, 42), np.float32(0.05))
P, count = 1 << 31, 1 << 24
out = {}
res = {}
for name, flag in (("select", "1"), ("radix", "0")):
    os.environ["MCCB200_SHUFFLE_SELECT"] = flag
    torch.cuda.reset_peak_memory_stats()
    idx = _lib.mc_indices(rng, P, count, False, dev); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    idx = _lib.mc_indices(rng, P, count, False, dev)
    b.record(); torch.cuda.synchronize()
    res[name] = idx.clone()
    out[name] = {"ms": a.elapsed_time(b), "peak_GiB": torch.cuda.max_memory_allocated() / 2**30}
    del idx
    _lib.release_workspace(); torch.cuda.empty_cache()
out["identical"] = bool(torch.equal(res["select"], res["radix"]))
out["distinct"] = int(torch.unique(res["select"].to(torch.int64) & 0xFFFFFFFF).numel())
print(json.dumps(out))
