"""Clock stamps of the phases of small_shuffle_kernel's first round (build with MCCB200_NVCC_FLAGS=-DMCCB200_SS_TRACE)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mccube_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
names = ["threefry", "atomics", "scan", "scatter", "place", "zero"]
for P in (1024, 8192, 16384):
    for part in (False, True):
        rng = _lib.make_rng((0, 42), 0.05, partitionable=part)
        for _ in range(3): _lib.mc_indices(rng, P, 64, False, dev)
        buf = (ctypes.c_longlong * 16)()
        lib.mccb200_debug_ss_trace(buf)
        t = list(buf)
        print(P, "partitionable" if part else "original", {n: t[i + 1] - t[i] for i, n in enumerate(names)})
