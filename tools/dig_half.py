import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, mpmath as mp
import arrowhead_b200 as ab, oracle
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
        if nd != 500 or name != "tiny tip":
            continue
        G = ctx.halfarrow(D, z); O = olib.halfarrow(D, z, oracle.default_tau(nd)); H = ohi.halfarrow(D, z, oracle.default_tau(nd))
        ctx.set_mode(1); G1 = ctx.halfarrow(D, z); ctx.set_mode(0)
        A = np.zeros((nd + 1, nd + 1)); A[np.arange(nd), np.arange(nd)] = D; A[:nd + 1, nd] = z
        # sigma_min via the tiny-tip structure in high precision: det(A) = prod(D) * ztip => product of singular values
        for tag, R in (("gpu", G), ("gpu mode1", G1), ("oracle", O), ("oracle_hi", H)):
            c = nd
            print(f"{tag:10s} k={nd+1}: sigma={R.lam[c]!r} sind={R.sind[c]} knu={R.knu[c]:.4g} krho={R.krho[c]:.4g} kb={R.kb[c]:.4g} kz={R.kz[c]:.4g} qout={R.qout[c]} status={R.status[c]} steps={R.steps[c]}")
        mp.mp.prec = 300
        logdet = sum(mp.log(mp.mpf(float(d))) for d in D) + mp.log(mp.mpf(1e-12))
        for tag, R in (("gpu", G), ("oracle", O)):
            ls = sum(mp.log(mp.mpf(float(x))) for x in R.lam[:-1])
            print(tag, "sigma_min implied by det / other sigmas:", float(mp.exp(logdet - ls)))
        print("lapack:", np.linalg.svd(A, compute_uv=False)[-1])
