"""Shared generators and comparison helpers for the parity tests."""
import numpy as np

EPS = float(np.finfo(np.float64).eps)


def gen_symarrow(n, seed=421):
    """SURVEY.md section 8d synthetic SymArrow: ordered, irreducible, arrow top last."""
    rng = np.random.default_rng(seed)
    N = n - 1
    D = np.sort(rng.random(N))[::-1].copy()
    z = rng.random(N)
    a = float(rng.random())
    return D, z, a


def gen_symdpr1(n, seed=421):
    rng = np.random.default_rng(seed)
    D = np.sort(rng.random(n))[::-1].copy()
    u = rng.random(n)
    rho = float(rng.random())
    return D, u, rho


def gen_halfarrow(nd, tip, seed=421):
    rng = np.random.default_rng(seed)
    D = np.sort(np.abs(rng.random(nd) - 0.5))[::-1].copy()
    z = rng.random(nd + (1 if tip else 0)) - 0.5
    return D, z


def shift_margin_symarrow(D, z, a, k):
    """relative distance of the shift-selection test (arrowhead3.jl:440-441) from its threshold."""
    Dl = D.astype(np.longdouble); zl = z.astype(np.longdouble)
    Dk = Dl[k - 1]
    mid = (Dl[k - 2] - Dk) / 2
    t = zl * zl / ((Dl - Dk) - mid)
    F = (np.longdouble(a) - Dk - mid) - t.sum()
    return float(abs(F) / (np.abs(t).sum() + abs(np.longdouble(a) - Dk - mid)))


def shift_margin_symdpr1(D, u, rho, k):
    Dl = D.astype(np.longdouble); ul = u.astype(np.longdouble)
    Dk = Dl[k - 1]
    mid = (Dl[k - 2] - Dk) / 2
    t = np.longdouble(rho) * ul * ul / ((Dl - Dk) - mid)
    F = 1 + t.sum()
    return float(abs(F) / (np.abs(t).sum() + 1))


def compare_range(gpu, orc, *, lam_tol=10.0, vec_tol=100.0, k_tol=1e-9, margin_fn=None, allow_status=()):
    """GPU k-range result vs oracle k-range result (same k-range).  Returns a dict of max errors.

    Bars (BASELINE.json north_star): Sind / Qout / status exact; eigenvalues <= 10 eps relative;
    eigenvector entries <= 100 eps relative; K values to k_tol relative (parallel vs sequential sums)."""
    cnt = gpu.lam.size
    assert cnt == orc.lam.size
    same_shift = gpu.sind == orc.sind
    bad = np.flatnonzero(~same_shift)
    for c in bad:
        k = gpu.k0 + c
        assert margin_fn is not None, f"shift index differs at k={k}: {gpu.sind[c]} vs {orc.sind[c]}"
        m = margin_fn(k)
        assert m < 1e-12, f"shift index differs at k={k} with a clear margin {m:.3e}"
    ok = same_shift
    assert np.array_equal(gpu.qout[ok], orc.qout[ok]), "Qout differs"
    assert np.array_equal(gpu.status[ok], orc.status[ok]), "status differs"
    out = {"n_shift_ambiguous": int(bad.size)}
    den = np.abs(orc.lam)
    den[den == 0] = 1.0
    lam_err = np.abs(gpu.lam - orc.lam) / den / EPS
    out["lam_err_eps"] = float(lam_err[ok].max()) if ok.any() else 0.0
    assert out["lam_err_eps"] <= lam_tol, f"eigenvalue error {out['lam_err_eps']:.2f} eps at c={int(np.argmax(lam_err))}"
    for name in ("kb", "kz", "knu", "krho"):
        g, o = getattr(gpu, name)[ok], getattr(orc, name)[ok]
        fin = np.isfinite(o) & (o != 0)
        rel = np.abs(g[fin] - o[fin]) / np.abs(o[fin])
        out[name + "_rel"] = float(rel.max()) if rel.size else 0.0
        assert out[name + "_rel"] <= k_tol, f"{name} differs: {out[name + '_rel']:.3e}"
        assert np.array_equal(g[~fin], o[~fin]) or not (~fin).any()
    for vname in ("U", "V"):
        G, O = getattr(gpu, vname, None), getattr(orc, vname, None)
        if G is None or O is None:
            continue
        G = np.asarray(G)[:, ok]; O = np.asarray(O)[:, ok]
        nz = O != 0
        rel = np.zeros_like(O)
        rel[nz] = np.abs(G[nz] - O[nz]) / np.abs(O[nz]) / EPS
        assert np.array_equal(G[~nz], O[~nz]), f"{vname}: zero pattern differs"
        out[vname + "_err_eps"] = float(rel.max()) if rel.size else 0.0
        assert out[vname + "_err_eps"] <= vec_tol, \
            f"{vname} entry error {out[vname + '_err_eps']:.1f} eps (col {int(np.argmax(rel.max(axis=0)))})"
    return out
