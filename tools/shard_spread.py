#!/usr/bin/env python
"""Kernel time and bisection work of every k-range shard, run one after the other on one GPU:
    python tools/shard_spread.py n world"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow
n, world = int(sys.argv[1]), int(sys.argv[2])
ctx = ab.get_context(0)
D, z, a = gen_symarrow(n)
for r in range(world):
    k0, k1 = ab.k_partition(n, world, r)
    U = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda")
    best = None
    for rep in range(4):
        res = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)
        s = res.stats
        if rep >= 1 and (best is None or s["kernel_ms"] < best["kernel_ms"]): best = s
    st = res.steps
    print(f"rank {r}: k={k0}..{k1} kernel={best['kernel_ms']:.3f} bisect={best['bisect_ms']:.3f} vec={best['vec_ms']:.3f} rem3={best['rem3_ms']:.3f} "
          f"n_rem3={best['n_rem3']} steps sum={int(st.sum())} mean={st.mean():.2f} min={st.min()} max={st.max()} p42={np.percentile(st,42):.0f}")
    del U
