import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mccube_b200 import _lib
dev = torch.device("cuda:0")
np.set_printoptions(linewidth=200, precision=4, suppress=True)
n = int(os.environ.get("N", "64"))
# structured input: x[r][j] = (r+1) * 0.01 + j  -> products easy to recognise
p = np.zeros((n, 64), np.float32)
mode = os.environ.get("MODE", "e")
if mode == "e":      # single non-zero entry per row: row r has 1 at column r % 64
    for r in range(n): p[r, r % 64] = 1.0 + r // 64
else:
    p = np.random.default_rng(0).normal(size=(n, 64)).astype(np.float32)
c = np.zeros(64, np.float32)
got = _lib.second_moment(torch.from_numpy(p).to(dev), 64, torch.from_numpy(c).to(dev)).cpu().numpy()
want = p.astype(np.float64).T @ p.astype(np.float64)
print("got diag", np.diag(got)[:16]); print("want diag", np.diag(want)[:16])
print("nonzero got", np.count_nonzero(got), "nonzero want", np.count_nonzero(want))
print("got[:8,:8]\n", got[:8, :8]); print("want[:8,:8]\n", want[:8, :8])
print("max err", np.abs(got - want).max())
