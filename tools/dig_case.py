import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import arrowhead_b200 as ab, oracle
ctx = ab.get_context(0); olib = oracle.get_lib(); ohi = oracle.get_lib_hi()
np.set_printoptions(linewidth=200)
def fams(rng, N):
    yield "clustered pairs 1e-10", np.sort(np.repeat(rng.random(N // 2), 2) + np.tile([0, 1e-10], N // 2))[::-1].copy(), rng.random(N), rng.random()
    yield "clustered 1e-14 rel", np.sort(1.0 + 1e-14 * np.arange(N) * (1 + rng.random(N)))[::-1].copy(), rng.random(N), rng.random()
    yield "wide z range", np.sort(rng.random(N) * 2 - 1)[::-1].copy(), 10.0 ** (-8 * rng.random(N)), rng.random()
    yield "huge a", np.sort(rng.random(N))[::-1].copy(), rng.random(N), 1e12
    yield "all negative", np.sort(-rng.random(N) - 0.1)[::-1].copy(), rng.random(N) - 0.5, -3.0
    yield "tiny z 1e-100", np.sort(rng.random(N))[::-1].copy(), 1e-100 * (0.5 + rng.random(N)), rng.random()
    yield "mixed tiny z", np.sort(rng.random(N))[::-1].copy(), np.where(rng.random(N) < 0.3, 1e-30, 1.0) * (0.5 + rng.random(N)), rng.random()
    yield "geometric D", np.sort(2.0 ** (-np.arange(N) * 60.0 / N))[::-1].copy(), rng.random(N), 0.0
def show(tag, G, O, H, ks):
    for k in ks:
        c = k - 1
        print(f"{tag} k={k}: lam G={G.lam[c]!r} O={O.lam[c]!r} H={H.lam[c]!r} | sind {G.sind[c]} {O.sind[c]} | knu {G.knu[c]:.6g} {O.knu[c]:.6g} {H.knu[c]:.6g} | krho {G.krho[c]:.4g} {O.krho[c]:.4g} | kb {G.kb[c]:.4g} {O.kb[c]:.4g} | steps {G.steps[c]} {O.steps[c]} {H.steps[c]} | qout {G.qout[c]} {O.qout[c]} status {G.status[c]} {O.status[c]}")
rng = np.random.default_rng(7)
N = 60
for name, D, z, a in fams(rng, N):
    if name != "geometric D":
        continue
    G = ctx.symarrow(D, z, a); O = olib.symarrow(D, z, a, oracle.default_tau(N)); H = ohi.symarrow(D, z, a, oracle.default_tau(N))
    show("geo", G, O, H, [N - 1, N, N + 1])
    A = np.zeros((N + 1, N + 1)); A[np.arange(N), np.arange(N)] = D; A[:N, N] = z; A[N, :N] = z; A[N, N] = a
    import mpmath as mp
    mp.mp.prec = 300
    E = mp.eigsy(mp.matrix(A.tolist()), eigvals_only=True)
    ev = sorted([E[i] for i in range(N + 1)], reverse=True)
    print("mp last 3:", [float(x) for x in ev[-3:]])
    print("G    last 3:", G.lam[-3:]); print("O    last 3:", O.lam[-3:]); print("H    last 3:", H.lam[-3:])
    ctx.set_mode(1); G1 = ctx.symarrow(D, z, a); ctx.set_mode(0)
    print("mode1 last 3:", G1.lam[-3:], "steps", G1.steps[-3:], "knu", G1.knu[-3:], "krho", G1.krho[-3:])
    tau = oracle.default_tau(N).copy(); tau[2] = 1e300
    Gn = ctx.symarrow(D, z, a, tau=tau, k0=N + 1, k1=N + 1); On = olib.symarrow(D, z, a, tau, k0=N + 1, k1=N + 1)
    print("no-remedy3 k=61: G lam", Gn.lam, "knu", Gn.knu, "steps", Gn.steps, "| O lam", On.lam, "knu", On.knu, "steps", On.steps)
    G1 = ctx.symarrow(D, z, a, k0=N + 1, k1=N + 1)
    print("only k=61: lam", G1.lam, "knu", G1.knu, "krho", G1.krho, "steps", G1.steps, "stats", {k: G1.stats[k] for k in ("n_full", "n_rem3")})
    print("z_i", z[-1], "D last", D[-2:], "pad?", D[-1] - (D[0] - D[-1]))
