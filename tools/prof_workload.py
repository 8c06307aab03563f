#!/usr/bin/env python
"""Workload for ncu: a few full SymArrow solves with U resident on the device (no checks, no timing)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import arrowhead_b200 as ab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(421)
D = np.sort(rng.random(n - 1))[::-1].copy(); z = rng.random(n - 1); a = float(rng.random())
ctx = ab.get_context(0)
U = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(reps):
    r = ctx.symarrow(D, z, a, vectors=False, U_dev=U.data_ptr(), ldu=n)
print(n, r.stats)
