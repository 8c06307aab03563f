#!/usr/bin/env python
"""Same-box A/B of two builds of libarrowhead_b200.so (boxes differ by a few %, so never compare across gpurun calls).
    python tools/ab_compare.py libA.so libB.so [n] [rounds]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow
n = int(sys.argv[1]); world = int(sys.argv[2])
ctx = ab.get_context(0)
D, z, a = gen_symarrow(n, seed=421)
k0, k1 = ab.k_partition(n, world, 0)
U = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda")
best = None
for rep in range(6):
    s = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n).stats
    if rep >= 2 and (best is None or s["kernel_ms"] < best["kernel_ms"]): best = s
print("kernel=%%.3f bisect=%%.3f prep=%%.3f vec=%%.3f rem3=%%.3f" %% (best["kernel_ms"], best["bisect_ms"], best["prep_ms"], best["vec_ms"], best["rem3_ms"]))
''' % (ROOT, ROOT)
libs = sys.argv[1:3]
n = sys.argv[3] if len(sys.argv) > 3 else "32768"
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 3
world = sys.argv[5] if len(sys.argv) > 5 else "1"
for r in range(rounds):
    for lib in libs:
        env = dict(os.environ, ARROWHEAD_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, "-c", code, n, world], env=env, capture_output=True, text=True)
        print(f"round {r} {os.path.basename(lib):28s} n={n}/{world}: {out.stdout.strip() or out.stderr[-300:]}", flush=True)
