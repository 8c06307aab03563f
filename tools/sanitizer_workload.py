#!/usr/bin/env python
"""Small run through every kernel of the library (for compute-sanitizer): fast pipeline incl. the Remedy-3 round and the
hand-over kernel, all three matrix kinds, deflation utilities, batched tdc primitives, Hermitian phase kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import arrowhead_b200 as ab
from helpers import gen_symarrow, gen_symdpr1, gen_halfarrow
ctx = ab.get_context(0)
D, z, a = gen_symarrow(2300, seed=421)
r = ctx.symarrow(D, z, a); print("symarrow", r.stats["n_rem3"], r.stats["n_full"], int(r.status.max()))
r = ctx.symarrow(D, z, a, k0=100, k1=777)
D, u, rho = gen_symdpr1(1500, seed=421)
r = ctx.symdpr1(D, u, rho); print("symdpr1", r.stats["n_rem3"], r.stats["n_full"])
D, z = gen_halfarrow(900, 1, seed=421)
r = ctx.halfarrow(D, z); print("halfarrow", r.stats["n_rem3"], r.stats["n_full"])
rng = np.random.default_rng(5)
N = 40
Dg = np.sort(np.exp(20 * (rng.random(N) - 0.5)) * np.where(rng.random(N) < 0.5, -1, 1))[::-1].copy()
zg = np.exp(20 * (rng.random(N) - 0.5))
r = ctx.symarrow(Dg, zg, 3.0); print("graded (double-double)", int((r.qout > 0).sum()))
E, info = ab.eigen(ab.SymArrow([1.0, 2.0, 2.0, 0.5, 3.0], [0.3, 0.0, 0.4, 0.2, 0.1], 0.7, 3), ctx=ctx); print("deflation", E.values)
E, _ = ab.eigen(ab.HermArrow(rng.random(30), rng.random(30) + 1j * rng.random(30), 0.3, 7), ctx=ctx); print("herm", E.vectors.shape)
n = 300
E = ab.tdc_device(2.0 + 1e-2 * rng.standard_normal(n), -1.0 + 1e-2 * rng.standard_normal(n - 1), ctx=ctx); print("tdc", E.stats)
print("done")
