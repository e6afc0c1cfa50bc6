"""Times the full-covariance drift (config 4: n = 2^20, d = 512) on the tensor-core and SIMT kernels."""
import sys, json
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mccube_b200 as mccube
from mccube_b200 import _lib

n_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
rs = np.random.default_rng(1)
a = rs.normal(size=(d, d)); cov = a @ a.T / d + np.eye(d)
mean = np.full(d, 2.0)
y = torch.randn((1 << n_log2, d), device=dev) * 1.5 + 2.0
out = {}
for name, tc in (("tcgen05_3xtf32", True), ("simt_fp32", False)):
    f = mccube.GaussianFullDrift(mean, cov, device=dev, tensor_cores=tc)
    for _ in range(2): r = f(0.0, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if tc else 2
    e0.record()
    for _ in range(reps): r = f(0.0, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * (1 << n_log2) * d * d
    out[name] = {"ms": ms, "useful_tflops": flop / ms / 1e9, "tensor_pipe_tflops": (3 * flop / ms / 1e9) if tc else None, "path": f.last_path}
    if tc: ref = r
    else: out["max_abs_diff_vs_simt"] = float((ref - r).abs().max()); out["max_abs"] = float(r.abs().max())
print(json.dumps(out))
