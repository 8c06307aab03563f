import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import arrowhead_b200 as ab, oracle
from helpers import compare_range, EPS, shift_margin_halfarrow
ctx = ab.get_context(0); olib = oracle.get_lib(); ohi = oracle.get_lib_hi()
def fams(rng, nd):
    yield "uniform", np.sort(rng.random(nd) + 0.01)[::-1].copy(), rng.random(nd + 1) - 0.5
    yield "graded D 1e-8..1", np.sort(10.0 ** (-8 * rng.random(nd)))[::-1].copy(), rng.random(nd + 1) - 0.5
    yield "graded z", np.sort(rng.random(nd) + 0.01)[::-1].copy(), 10.0 ** (-8 * rng.random(nd + 1))
    yield "clustered D", np.sort(1.0 + 1e-9 * np.arange(nd) * (1 + rng.random(nd)))[::-1].copy(), rng.random(nd + 1)
    yield "both graded", np.sort(np.exp(20 * (rng.random(nd) - 0.5)))[::-1].copy(), np.exp(20 * (rng.random(nd + 1) - 0.5))
    yield "tiny tip", np.sort(rng.random(nd) + 0.01)[::-1].copy(), np.concatenate([rng.random(nd), [1e-12]])
rng = np.random.default_rng(11)
for nd in (40, 500):
    for name, D, z in fams(rng, nd):
        for tip in (1, 0):
            zz = z if tip else z[:nd]
            try:
                G = ctx.halfarrow(D, zz); O = olib.halfarrow(D, zz, oracle.default_tau(nd)); H = ohi.halfarrow(D, zz, oracle.default_tau(nd))
                st = f"status G{np.unique(G.status)} O{np.unique(O.status)} qout G{np.unique(G.qout)} O{np.unique(O.qout)} full={G.stats['n_full']} rem3={G.stats['n_rem3']}"
                A = np.zeros((max(nd, zz.size), nd + 1)); A[np.arange(nd), np.arange(nd)] = D; A[:zz.size, nd] = zz
                sv = np.linalg.svd(A, compute_uv=False)
                acc = np.max(np.abs(G.lam - sv[:G.lam.size]) / sv[:G.lam.size]) / EPS
                try:
                    out = compare_range(G, O, H, margin_fn=lambda k: shift_margin_halfarrow(D, zz, k), vec_tol=1000.0, max_excluded=G.lam.size, poles=D)
                    print(f"OK   nd={nd} tip={tip} {name:18s} lam={out['lam_err_eps']:.1f} V={out.get('V_err_eps', 0):.1f} vs-lapack={acc:.1f}eps {st}")
                except AssertionError as e:
                    print(f"DIFF nd={nd} tip={tip} {name:18s} {str(e)[:140]} | vs-lapack={acc:.1f}eps {st}")
            except Exception as e:
                print(f"ERR  nd={nd} tip={tip} {name:18s} {type(e).__name__}: {str(e)[:160]}")
