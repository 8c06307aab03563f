#!/usr/bin/env python
"""Workload for ncu: `reps` full solves of one matrix kind with U (and V) resident on the device, bench.py's generators.
    python tools/prof_kinds.py symarrow|symdpr1|halfarrow n reps"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arrowhead_b200 as ab
import bench

kind = sys.argv[1] if len(sys.argv) > 1 else "symarrow"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
P = bench.Problem(kind, n)
ctx = ab.get_context(0)
U = torch.empty((n, n), dtype=torch.float64, device="cuda")
V = torch.empty((n, n), dtype=torch.float64, device="cuda") if kind == "halfarrow" else None
for _ in range(reps):
    r = P.solve(ctx, 1, n, U.data_ptr(), None if V is None else V.data_ptr())
print(kind, n, r.stats)
