"""mean_and_cov passes timed alone: column sums (HBM-bound) and centred second moments (FP32-FMA-bound)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mccube_b200 import _lib

dev = torch.device("cuda:0")
out = {}
for n, d in [(1 << 24, 64), (1 << 20, 512), (1 << 20, 10)]:
    y = torch.randn((n, d), device=dev) * 1.7 + 2.0
    c = y[:4096].mean(0).contiguous()
    res = {}
    for name, fn in [("sum", lambda: _lib.column_means(y, d)), ("m2", lambda: _lib.second_moment(y, d, c))]:
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        res[name] = dict(ms=ms, read_gbs=n * d * 4 / ms / 1e6, tflops=(2 * n * d * d / ms / 1e9) if name == "m2" else None)
    out[f"n=2^{n.bit_length() - 1},d={d}"] = res
print(json.dumps(out))
