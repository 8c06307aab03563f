#!/usr/bin/env python
"""Duration of the Remedy-3 round alone (no vectors -> everything on one stream): python tools/rem3_time.py n world"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import arrowhead_b200 as ab
from helpers import gen_symarrow
n, world = int(sys.argv[1]), int(sys.argv[2])
ctx = ab.get_context(0)
D, z, a = gen_symarrow(n)
k0, k1 = ab.k_partition(n, world, 0)
best = None
for rep in range(5):
    s = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False).stats
    if rep >= 1 and (best is None or s["rem3_ms"] < best["rem3_ms"]): best = s
print(f"n={n}/{world}: n_rem3={best['n_rem3']} rem3_ms={best['rem3_ms']:.3f} bisect={best['bisect_ms']:.3f} kernel={best['kernel_ms']:.3f}")
