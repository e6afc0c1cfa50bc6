"""Single-CTA shuffle kernel (mc_indices with <= 16384 inputs, without replacement) timed alone."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mccube_b200 import _lib

dev = torch.device("cuda:0")
out = {}
for P, n in [(1024, 64), (8192, 64), (16384, 512), (16385, 512), (65536, 512)]:
    rng = _lib.make_rng((0, 42), 0.05)
    _lib.mc_indices(rng, P, n, False, dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        _lib.mc_indices(rng, P, n, False, dev)
    b.record(); torch.cuda.synchronize()
    out[f"P={P}"] = a.elapsed_time(b) / 50 * 1e3
print(json.dumps(out))
