#!/usr/bin/env python
"""Kernel-time breakdown of the three matrix kinds at their BASELINE config sizes (U/V resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow, gen_symdpr1, gen_halfarrow
ctx = ab.get_context(0)
def show(name, n, s, steps):
    N = n
    print(f"{name}: kernel={s['kernel_ms']:.2f} ms prep={s['prep_ms']:.2f} bisect={s['bisect_ms']:.2f} rem3={s['rem3_ms']:.2f} vec={s['vec_ms']:.2f} "
          f"full={s['full_ms']:.2f} n_full={s['n_full']} n_rem3={s['n_rem3']} -> {n / (s['kernel_ms'] * 1e-3):.3e} pairs/s, "
          f"bisect terms/s={steps * N / ((s['bisect_ms'] + s['rem3_ms']) * 1e-3):.3e}")
n = 16384
D, u, rho = gen_symdpr1(n, seed=421)
U = torch.empty((n, n), dtype=torch.float64, device="cuda")
for _ in range(3): G = ctx.symdpr1(D, u, rho, vectors=False, U_dev=U.data_ptr(), ldu=n)
show("SymDPR1 n=16384", n, G.stats, int(G.steps.sum()))
nd = 16383
D, z = gen_halfarrow(nd, 1, seed=421)
V = torch.empty((nd + 1, nd + 1), dtype=torch.float64, device="cuda")
for _ in range(3): G = ctx.halfarrow(D, z, vectors=False, U_dev=U.data_ptr(), ldu=nd + 1, V_dev=V.data_ptr(), ldv=nd + 1)
show("HalfArrow nd=16383 tip", nd + 1, G.stats, int(G.steps.sum()))
n = 16384
D, z, a = gen_symarrow(n, seed=421)
for _ in range(3): G = ctx.symarrow(D, z, a, vectors=False, U_dev=U.data_ptr(), ldu=n)
show("SymArrow n=16384", n, G.stats, int(G.steps.sum()))
