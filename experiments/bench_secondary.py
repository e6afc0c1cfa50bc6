"""Sweep of the kernels that are NOT in the headline step, each timed alone with its algorithmic GB/s:
a sanity check that none of them sits far below the HBM roofline without a stated reason."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mccube_b200 import _lib

dev = torch.device("cuda:0")
out = {}

def timed(name, fn, bytes_moved, reps=5):
    fn(); fn(); torch.cuda.synchronize()   # two warm-up calls: the first ones pay for the allocator growing
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out[name] = dict(ms=round(ms, 4), algorithmic_gbs=round(bytes_moved / ms / 1e6, 1))
    del r

n, d, k = 1 << 17, 64, 128
y = torch.randn((n, d), device=dev)
mean = torch.full((d,), 2.0, device=dev); ivar = torch.full((d,), 1 / 3.0, device=dev)
drift = _lib.make_drift(2, mean=mean, inv_var=ivar)
cub = _lib.make_cubature(k, scale=0.3)
timed("expand_all n=2^17 d=64 k=128 (writes 4.3 GB)", lambda: _lib.expand_all(y, d, False, drift, 0.05, cub), n * k * d * 4)
big = torch.randn((1 << 24, d), device=dev)
idx = torch.randint(0, 1 << 24, (1 << 24,), device=dev, dtype=torch.int32)
timed("gather_rows 2^24 x 256 B", lambda: _lib.gather_rows(big, idx), 2 * (1 << 24) * d * 4)
slots = torch.randperm(1 << 22, device=dev).to(torch.int32)
outbuf = torch.empty((1 << 22, d), device=dev)
timed("scatter_rows 2^22 x 256 B", lambda: _lib.scatter_rows_(outbuf, big[: 1 << 22], slots), 2 * (1 << 22) * d * 4)
yw = torch.rand((1 << 24, d + 1), device=dev)
timed("normalise_weights n=2^24 ld=65", lambda: _lib.normalise_weights_(yw), 3 * (1 << 24) * 4)
timed("normalised_copy n=2^24 ld=65 (clone + normalise; sum computed)", lambda: _lib.normalised_copy(yw), 2 * (1 << 24) * (d + 1) * 4)
tot = yw[:, -1].double().sum().reshape(1)
timed("normalised_copy n=2^24 ld=65 (sum handed over by the gather)", lambda: _lib.normalised_copy(yw, tot), 2 * (1 << 24) * (d + 1) * 4)
timed("clone n=2^24 ld=65 (torch, for scale)", lambda: yw.clone(), 2 * (1 << 24) * (d + 1) * 4)
rng = _lib.make_rng((0, 42), 0.05)
timed("mc_indices replace=True 2^24 of 2^31", lambda: _lib.mc_indices(rng, 1 << 31, 1 << 24, True, dev), (1 << 24) * 4)
com = big.mean(0).contiguous()
timed("stratified_keys n=2^24 d=64", lambda: _lib.stratified_keys(big, d, com), (1 << 24) * (d + 1) * 4)
p = torch.rand(1 << 22, device=dev); p /= p.sum()
timed("weighted_indices Gumbel top-k P=2^22 n=2^16", lambda: _lib.weighted_indices(rng, p, 1 << 16, False), (1 << 22) * 4)
timed("weighted_indices inverse-CDF P=2^22 n=2^22", lambda: _lib.weighted_indices(rng, p, 1 << 22, True), 2 * (1 << 22) * 4)
keys = torch.randint(0, 2 ** 31 - 1, (1 << 24,), device=dev, dtype=torch.int32)
timed("mc_indices replace=False 2^20 of 2^25 (selection shuffle)", lambda: _lib.mc_indices(rng, 1 << 25, 1 << 20, False, dev), (1 << 20) * 4)
timed("kd_partition n=2^20 d=64 leaf=64 (14 levels)", lambda: _lib.kd_partition(big[: 1 << 20], d, 14), 14 * (1 << 20) * d * 4)
timed("sort_pairs 2^24 x 32 bits", lambda: _lib.sort_pairs(keys, None, 32), 4 * 20 * (1 << 24))
bounds = _lib.box_bounds(big, (0, 1, 2, 3), 3)
timed("box_bounds rows n=2^24 (4 of 64 columns)", lambda: _lib.box_bounds(big, (0, 1, 2, 3), 3), (1 << 24) * 64)
timed("box_keys rows n=2^24", lambda: _lib.box_keys(big, (0, 1, 2, 3), 3, bounds), (1 << 24) * (64 + 4))
print(json.dumps(out, indent=1))
