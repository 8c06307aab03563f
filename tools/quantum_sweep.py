#!/usr/bin/env python
"""k_bisect time of a 1/world shard for several time-slice settings (ARROWHEAD_B200_QUANTUM is read at ah_create):
    python tools/quantum_sweep.py lib.so [world] [n]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.abspath(sys.argv[1]); world = sys.argv[2] if len(sys.argv) > 2 else "8"; n = sys.argv[3] if len(sys.argv) > 3 else "32768"
code = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow
n = int(sys.argv[1]); world = int(sys.argv[2])
ctx = ab.get_context(0)
D, z, a = gen_symarrow(n, seed=421)
out = []
for r in (0, world // 2, world - 1):
    k0, k1 = ab.k_partition(n, world, r)
    U = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda")
    best = None
    for rep in range(5):
        s = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n).stats
        if rep >= 1 and (best is None or s["bisect_ms"] < best["bisect_ms"]): best = s
    out.append("r%%d bisect=%%.3f kernel=%%.3f" %% (r, best["bisect_ms"], best["kernel_ms"]))
print(" | ".join(out))
''' % (ROOT, ROOT)
for q in ("0", "4", "8", "12", "16", "24", "32"):
    env = dict(os.environ, ARROWHEAD_B200_LIB=lib, ARROWHEAD_B200_QUANTUM=q)
    out = subprocess.run([sys.executable, "-c", code, n, world], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):24s} quantum={q:3s} n={n}/{world}: {out.stdout.strip() or out.stderr[-300:]}", flush=True)
