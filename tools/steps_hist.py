#!/usr/bin/env python
"""Distribution of the sweeps per eigenpair (round 0 only for eigenpairs without Remedy 3): python tools/steps_hist.py n world rank"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import arrowhead_b200 as ab, bench
n, world, rank = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
P = bench.Problem("symarrow", n)
k0, k1 = ab.k_partition(n, world, rank)
ctx = ab.get_context(0)
r = ctx.symarrow(P.D, P.z, P.a, k0=k0, k1=k1, vectors=False)
st = r.steps; rem3 = (r.status & 0x200) != 0
a = st[~rem3]
print("no Remedy 3:", a.size, "mean", a.mean(), "percentiles 50/90/99/99.9/max", [int(np.percentile(a, q)) for q in (50, 90, 99, 99.9)], a.max())
print("histogram (sweeps: count):", {int(v): int(c) for v, c in zip(*np.unique(a, return_counts=True))})
b = st[rem3]
print("Remedy 3:", b.size, "mean", b.mean(), "min/max", b.min(), b.max(), sorted(b.tolist())[-10:])
print("knu of the 10 longest no-rem3:", [(int(st[i]), float(r.knu[i])) for i in np.argsort(np.where(rem3, 0, st))[-10:]])
