#!/usr/bin/env python
"""Quick GPU diagnostic (not a test, not the bench): parity error statistics + kernel timings."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import arrowhead_b200 as ab
import oracle
from helpers import EPS, gen_symarrow, gen_symdpr1, gen_halfarrow

ab.build_library()
ctx = ab.get_context(0)
g, ms = ctx.measure_fp64_peak()
print(f"FP64 pipe: {g:.1f} G lane-instr/s  ({2*g/1e3:.2f} TFLOP/s as FMA), burst {ms:.2f} ms")
print(f"store bw: {ctx.measure_store_bw(4 << 30):.0f} GB/s")
olib = oracle.get_lib()
for n in (10, 1000, 4100):
    D, z, a = gen_symarrow(n, seed=421)
    O = olib.symarrow(D, z, a, oracle.default_tau(n - 1))
    for mode in (1, 0):
        ctx.set_mode(mode)
        G = ctx.symarrow(D, z, a)
        ok = G.sind == O.sind
        le = np.abs(G.lam - O.lam) / np.abs(O.lam) / EPS
        ve = np.abs(G.U - O.U) / np.abs(O.U) / EPS
        print(f"n={n} mode={mode} shift_mismatch={int((~ok).sum())} qout_mismatch={int((G.qout != O.qout).sum())} "
              f"status={np.unique(G.status)} lam_err={le[ok].max():.2f}eps vec_err={ve[:, ok].max():.1f}eps "
              f"(p99 {np.percentile(ve[:, ok].max(axis=0), 99):.1f}) steps {G.steps.sum()} vs {O.steps.sum()} "
              f"knu_rel={np.max(np.abs(G.knu-O.knu)/O.knu):.2e} kb_rel={np.max(np.abs(G.kb-O.kb)/np.abs(O.kb)):.2e} "
              f"orth={np.linalg.norm(G.U.T@G.U-np.eye(n))/EPS/n:.2f}n*eps stats={G.stats}")
ctx.set_mode(0)
import torch
hp = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)      # 1 GiB
dv = torch.zeros(1 << 27, dtype=torch.float64, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time(); hp.copy_(dv, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t0
print(f"D2H pinned 1 GiB: {hp.numel()*8/dt/1e9:.1f} GB/s")
del hp, dv
for n in (8192, 32768):
    D, z, a = gen_symarrow(n, seed=421)
    U = torch.empty((n, n), dtype=torch.float64, device="cuda")
    for rep in range(3):
        t0 = time.time()
        G = ctx.symarrow(D, z, a, vectors=False, U_dev=U.data_ptr(), ldu=n)
        dt = time.time() - t0
        s = G.stats
        terms = (G.steps.sum() + 4.0 * n) * (n - 1)
        print(f"n={n} rep={rep} wall={dt*1e3:.1f} ms kernel={s['kernel_ms']:.2f} ms main={s['main_kernel_ms']:.2f} ms "
              f"prep={s['prep_ms']:.2f} vec={s['vec_ms']:.2f} full={s['full_ms']:.2f} n_full={s['n_full']} n_rem3={s['n_rem3']} eigenpairs/s={n/(s['kernel_ms']*1e-3):.3e} terms/s={terms/(s['kernel_ms']*1e-3):.3e} "
              f"fp64 instr/s (7/term)={7*terms/(s['kernel_ms']*1e-3)/1e9:.0f} G")
