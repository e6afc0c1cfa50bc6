"""README configuration (n=512, d=10, 512 steps; BASELINE config 1): eager step loop vs the whole solve
replayed from one CUDA graph (`diffrax_compat.capture_solve`).  Wall-clock per solve, result on device."""
import math, time, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mccube_b200 as mccube
from mccube_b200 import diffrax_compat as dfx

dev = torch.device("cuda:0")
n, d = 512, 10
f = mccube.GaussianDiagDrift(np.full(d, 2.0, np.float32), np.full(d, 3.0), device=dev)
terms = mccube.MCCTerm(dfx.ODETerm(f), dfx.WeaklyDiagonalControlTerm(
    lambda t, y, args: math.sqrt(2.0), mccube.LocalLinearCubaturePath(mccube.Hadamard(mccube.GaussianRegion(d)))))
out = {}
for replace in (False, True):
    solver = mccube.MCCSolver(dfx.Euler(), mccube.MonteCarloKernel(n, replace, key=42))
    y0 = torch.ones((n, d), device=dev)
    run = lambda: dfx.diffeqsolve(terms, solver, 0.0, 0.05 * 512, 0.05, y0)
    run(); torch.cuda.synchronize()
    t = time.perf_counter(); [run() for _ in range(3)]; torch.cuda.synchronize()
    eager = (time.perf_counter() - t) / 3
    solve = dfx.capture_solve(terms, solver, 0.0, 0.05 * 512, 0.05, y0)
    solve(y0); torch.cuda.synchronize()
    t = time.perf_counter(); [solve(y0) for _ in range(10)]; torch.cuda.synchronize()
    graph = (time.perf_counter() - t) / 10
    out[f"replace={replace}"] = dict(eager_ms=eager * 1e3, graph_ms=graph * 1e3, steps=512,
                                     graph_us_per_step=graph * 1e6 / 512)
print(json.dumps(out))
