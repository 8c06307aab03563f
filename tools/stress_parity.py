#!/usr/bin/env python
"""Exploratory parity sweep over structured matrix families (GPU vs oracle); prints what differs."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import arrowhead_b200 as ab, oracle
from helpers import compare_range, shift_margin_symarrow, shift_margin_symdpr1
ctx = ab.get_context(0); olib = oracle.get_lib(); ohi = oracle.get_lib_hi()
def fams(rng, N):
    yield "clustered pairs 1e-10", np.sort(np.repeat(rng.random(N // 2), 2) + np.tile([0, 1e-10], N // 2))[::-1].copy(), rng.random(N), rng.random()
    yield "clustered 1e-14 rel", np.sort(1.0 + 1e-14 * np.arange(N) * (1 + rng.random(N)))[::-1].copy(), rng.random(N), rng.random()
    yield "wide z range", np.sort(rng.random(N) * 2 - 1)[::-1].copy(), 10.0 ** (-8 * rng.random(N)), rng.random()
    yield "huge a", np.sort(rng.random(N))[::-1].copy(), rng.random(N), 1e12
    yield "all negative", np.sort(-rng.random(N) - 0.1)[::-1].copy(), rng.random(N) - 0.5, -3.0
    yield "tiny z 1e-100", np.sort(rng.random(N))[::-1].copy(), 1e-100 * (0.5 + rng.random(N)), rng.random()
    yield "mixed tiny z", np.sort(rng.random(N))[::-1].copy(), np.where(rng.random(N) < 0.3, 1e-30, 1.0) * (0.5 + rng.random(N)), rng.random()
    yield "geometric D", np.sort(2.0 ** (-np.arange(N) * 60.0 / N))[::-1].copy(), rng.random(N), 0.0
    yield "sym about 0", np.sort(np.concatenate([np.linspace(0.1, 1, N // 2), -np.linspace(0.1, 1, N - N // 2) * (1 + 1e-3)]))[::-1].copy(), rng.random(N), 0.0
rng = np.random.default_rng(7)
for N in (60, 700):
    for name, D, z, a in fams(rng, N):
        for kind in ("arrow", "dpr1"):
            try:
                if kind == "arrow":
                    G = ctx.symarrow(D, z, a); O = olib.symarrow(D, z, a, oracle.default_tau(N)); H = ohi.symarrow(D, z, a, oracle.default_tau(N))
                    mf = lambda k: shift_margin_symarrow(D, z, a, k)
                else:
                    rho = abs(a) + 0.1
                    G = ctx.symdpr1(D, z, rho); O = olib.symdpr1(D, z, rho, oracle.default_tau(N)); H = ohi.symdpr1(D, z, rho, oracle.default_tau(N))
                    mf = lambda k: shift_margin_symdpr1(D, z, rho, k)
                st = f"status G{np.unique(G.status)} O{np.unique(O.status)} qout G{np.unique(G.qout)} O{np.unique(O.qout)} nanG={int(np.isnan(G.lam).sum())} nanO={int(np.isnan(O.lam).sum())} full={G.stats['n_full']} rem3={G.stats['n_rem3']}"
                try:
                    out = compare_range(G, O, H, margin_fn=mf)
                    print(f"OK   N={N} {kind:5s} {name:22s} lam={out['lam_err_eps']:.1f} U={out.get('U_err_eps', 0):.1f} {st}")
                except AssertionError as e:
                    print(f"DIFF N={N} {kind:5s} {name:22s} {str(e)[:150]} | {st}")
            except Exception as e:
                print(f"ERR  N={N} {kind:5s} {name:22s} {type(e).__name__}: {str(e)[:200]}")
