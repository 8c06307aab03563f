#!/usr/bin/env python
"""bisect time per bisection step (diagnostic builds change the step counts): python tools/steps_ab.py lib.so n world"""
import os, sys
lib, n, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
os.environ["ARROWHEAD_B200_LIB"] = os.path.abspath(lib)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow
ctx = ab.get_context(0)
D, z, a = gen_symarrow(n)
k0, k1 = ab.k_partition(n, world, 0)
best = None
for rep in range(4):
    r = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False)
    s = r.stats
    if rep >= 1 and (best is None or s["bisect_ms"] < best["bisect_ms"]): best = dict(s)
print(f"{os.path.basename(lib):20s} n={n}/{world}: bisect={best['bisect_ms']:.3f} ms total_steps={best['total_steps']} n_full={best['n_full']} n_rem3={best['n_rem3']} "
      f"-> {best['bisect_ms'] * 1e6 / best['total_steps']:.3f} ns per step")
