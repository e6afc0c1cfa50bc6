#!/usr/bin/env python
"""Prints the JSON lines of fused_knobs.py compactly."""
import json
import sys

for line in open(sys.argv[1]):
    d = json.loads(line)
    print(d["knobs"] or "(defaults)", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["phase_ms"].items()})
