#!/usr/bin/env python
"""Raw pinned-memory copy rates on this box: H2D alone, D2H alone, both at once, whole buffer or split over streams.
Tells how much of the e2e leg (4 GiB in + 4 GiB out per step) is the link."""
import json
import torch

dev = torch.device("cuda:0")
n_bytes = 1 << 32
h_in = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
d_out = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
out = {}


def run(name, h2d, d2h, chunks):
    streams_in = [torch.cuda.Stream() for _ in range(chunks)]
    streams_out = [torch.cuda.Stream() for _ in range(chunks)]
    step = n_bytes // chunks
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in streams_in + streams_out:
            s.wait_event(a)
        for c in range(chunks):
            sl = slice(c * step, (c + 1) * step)
            if h2d:
                with torch.cuda.stream(streams_in[c]):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(streams_out[c]):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
        for s in streams_in + streams_out:
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    out[name] = {"ms": best, "GBps_each_way": n_bytes / best / 1e6}


for chunks in (1, 2, 4):
    run(f"h2d_x{chunks}", True, False, chunks)
    run(f"d2h_x{chunks}", False, True, chunks)
    run(f"both_x{chunks}", True, True, chunks)
print(json.dumps(out, indent=1))
