"""Stage-ring timeline of the tensor-core drift kernel (CTA 0): clock64 stamps per stage."""
import sys, ctypes
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mccube_b200 as mccube
from mccube_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
d = 512
rs = np.random.default_rng(1)
a = rs.normal(size=(d, d)); cov = a @ a.T / d + np.eye(d)
y = torch.randn((1 << 20, d), device=dev) * 1.5 + 2.0
f = mccube.GaussianFullDrift(np.full(d, 2.0), cov, device=dev)
f(0.0, y); torch.cuda.synchronize()
buf = torch.zeros(256 * 5, dtype=torch.int64, device=dev)
lib.mccb200_debug_gemm_trace.argtypes = [ctypes.c_void_p]
lib.mccb200_debug_gemm_trace(buf.data_ptr())
f(0.0, y); torch.cuda.synchronize()
lib.mccb200_debug_gemm_trace(None)
t = buf.cpu().numpy().reshape(256, 5)
t0 = t[0, 0]
print("stage: producer_issue full_seen converted mma_start mma_issued   | load conv wait_mma issue  period")
for g in range(60, 100):
    r = t[g] - t0
    print(g, r.tolist(), "|", t[g,1]-t[g,0], t[g,2]-t[g,1], t[g,3]-t[g,2], t[g,4]-t[g,3], t[g,0]-t[g-1,0])
