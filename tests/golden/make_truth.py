#!/usr/bin/env python
"""High-precision TRUTH eigenpairs for the BASELINE.json configs, independent of oracle/ and of the CUDA kernels.

What it computes (mpmath, 320-bit arithmetic; nothing is imported from oracle/ or arrowhead.jl_b200/):

  SymArrow  [diag(D) z; z' a]     lambda_k = k-th largest root of  f = a - lambda - sum z_l^2/(D_l - lambda)
                                  v = [z_l/(D_l - lambda); -1] / ||.||            (arrowhead3.jl:516-522 sign convention)
  SymDPR1   diag(D) + rho u u'    root of  f = 1 + rho sum u_l^2/(D_l - lambda),  v = [u_l/(D_l - lambda)] / ||.||
                                  (arrowhead4.jl:391-396)
  HalfArrow [diag(D) z] (tip)     sigma_k^2 = eigenvalue of the arrow (D.^2, z.*D, z'z), v as SymArrow,
                                  u = [sigma v_l/D_l ; z_n v_n/sigma]            (arrowhead6.jl:410-425)

Every root is CERTIFIED: it lies in its Cauchy-interlacing interval and f changes sign between
x(1 - 2^-200) and x(1 + 2^-200), x = lambda - (nearest pole).  f is strictly monotone between poles, so that
pins the k-th eigenvalue to 200 bits relative to its distance from the pole -- far below the 1e-16 the tests need.

Stored per eigenpair (tests/golden/truth_<case>.npz):
    k, i (0-based shift pole), x = lambda - D_i as (hi, lo) doubles, lambda (hi, lo), the norm (hi, lo),
    and mp-rounded eigenvector entries on a row sample (neighbours of the shift pole, the ends, random rows).
From (i, x_hi, x_lo, norm) a test can form EVERY entry of the truth vector in long double with ~2^-62 relative
error (d_l - x never cancels by more than a factor 2 because D_i is the nearest pole); tests/test_truth.py checks
that derivation against the stored mp-rounded rows before using it.

Inputs are the seeded generators of tests/helpers.py (restated below so the file is self-contained).

    python tests/golden/make_truth.py            # all cases, ~10 minutes on 8 cores
    python tests/golden/make_truth.py symarrow_n1000
"""
from __future__ import annotations

import multiprocessing as mproc
import os
import sys
import time

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PREC = 320
LD = np.longdouble


# ---- the synthetic inputs (same streams as tests/helpers.py gen_*) ---------------------------------------------
def gen_symarrow(n, seed=421):
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


# ---- one secular problem in mp: poles P (descending), weights W2 = w_l^2, form 'arrow' | 'dpr1' -------------------
_G = {}


def _init(form, poles, w2, top, wsrc):
    mp.mp.prec = PREC
    _G["form"] = form
    _G["P"] = poles            # list of mpf, exact
    _G["W2"] = w2              # list of mpf, exact
    _G["top"] = top            # mpf: arrow top a, or rho for dpr1
    _G["w"] = wsrc             # list of mpf: the numerators of the eigenvector entries (z or u or z.*D)
    _G["Pld"] = np.array([LD(float(p)) + LD(float(p - mp.mpf(float(p)))) for p in poles])
    _G["W2ld"] = np.array([LD(float(q)) + LD(float(q - mp.mpf(float(q)))) for q in w2])
    _G["topld"] = LD(float(top)) + LD(float(top - mp.mpf(float(top))))


def _f_ld(lam):
    """long-double value of f at lam (only its sign near the middle of an interval is used)."""
    t = _G["W2ld"] / (_G["Pld"] - lam)
    if _G["form"] == "arrow":
        return (_G["topld"] - lam) - t.sum()
    return 1 + _G["topld"] * t.sum()


def _f_shifted_ld(dl, x, s):
    t = _G["W2ld"] / (dl - x)
    if _G["form"] == "arrow":
        return ((_G["topld"] - s) - x) - t.sum()
    return 1 + _G["topld"] * t.sum()


def _f_mp(d, x, s):
    """(f, f', sum w2/(d-x)^2) at lambda = s + x in mp; d_l = P_l - s exact."""
    S = mp.mpf(0)
    S2 = mp.mpf(0)
    for dl, w2 in zip(d, _G["W2"]):
        t = dl - x
        q = w2 / t
        S += q
        S2 += q / t
    if _G["form"] == "arrow":
        return (_G["top"] - s) - x - S, -1 - S2, S2
    return 1 + _G["top"] * S, _G["top"] * S2, S2


def _solve_one(args):
    k, rows = args
    mp.mp.prec = PREC
    P, form = _G["P"], _G["form"]
    N = len(P)
    Pld = _G["Pld"]
    incr = form == "dpr1"                  # f increasing (dpr1, rho > 0) or decreasing (arrow) between poles
    kmax = N + 1 if form == "arrow" else N
    assert 1 <= k <= kmax
    # ---- interlacing interval (lo, hi) and the nearest pole i
    if k == 1:
        i = 0
        if form == "arrow":
            span = abs(_G["topld"] - Pld[0]) + np.sqrt(_G["W2ld"].sum()) * 2 + abs(Pld[0]) * LD(1e-3) + LD(1e-300)
        else:
            span = _G["topld"] * _G["W2ld"].sum() * LD(1.0000001) + LD(1e-300)
        xlo, xhi = LD(0), span              # x = lambda - P_0 in (0, span]
    elif k == N + 1:
        i = N - 1
        span = abs(_G["topld"] - Pld[N - 1]) + np.sqrt(_G["W2ld"].sum()) * 2 + abs(Pld[N - 1]) * LD(1e-3) + LD(1e-300)
        xlo, xhi = -span, LD(0)
    else:
        lo, hi = Pld[k - 1], Pld[k - 2]     # 0-based poles P[k-1] < lambda < P[k-2]
        mid = lo + (hi - lo) / 2
        fm = _f_ld(mid)
        # root below mid <=> (decreasing f: f(mid) < 0) | (increasing f: f(mid) > 0)
        below = (fm > 0) if incr else (fm < 0)
        if below:
            i = k - 1
            xlo, xhi = LD(0), mid - lo
        else:
            i = k - 2
            xlo, xhi = mid - hi, LD(0)
    s_ld = Pld[i]
    dl_ld = Pld - s_ld
    # ---- long-double bisection on x (the pole end is never evaluated)
    for _ in range(80):
        xm = xlo + (xhi - xlo) / 2
        if xm == xlo or xm == xhi:
            break
        fv = _f_shifted_ld(dl_ld, xm, s_ld)
        right = (fv < 0) if incr else (fv > 0)       # root to the right of xm
        if right:
            xlo = xm
        else:
            xhi = xm
    # ---- mp: safeguarded Newton from the long-double root, then the sign-change certificate
    s = P[i]
    d = [p - s for p in P]
    a_, b_ = mp.mpf(float(xlo)), mp.mpf(float(xhi))   # may not bracket in mp terms if long double was off by an ulp: widen
    wid = abs(b_ - a_) + max(abs(a_), abs(b_)) * mp.mpf(2) ** -36 + mp.mpf(2) ** -1070   # long-double sign noise: K_nu-fold
    a_ -= 4 * wid
    b_ += 4 * wid
    if i == k - 1 or k == 1:                # x > 0 side
        a_ = max(a_, mp.mpf(2) ** -1074)
    else:
        b_ = min(b_, -mp.mpf(2) ** -1074)
    x = (mp.mpf(float(xlo)) + mp.mpf(float(xhi))) / 2
    evals = 0
    for it in range(60):
        f, fp, _ = _f_mp(d, x, s)
        evals += 1
        if f == 0:
            break
        right = (f < 0) if incr else (f > 0)
        if right:
            a_ = max(a_, x)
        else:
            b_ = min(b_, x)
        xn = x - f / fp
        if abs(xn - x) <= abs(xn) * mp.mpf(2) ** -260 and a_ <= xn <= b_:
            x = xn
            break
        if not (a_ < xn < b_):
            xn = (a_ + b_) / 2
        x = xn
    else:
        raise RuntimeError(f"k={k}: Newton did not converge")
    e = mp.mpf(2) ** -200
    x1, x2 = x * (1 - e), x * (1 + e)
    f1, _, _ = _f_mp(d, min(x1, x2), s)
    f2, _, S2 = _f_mp(d, max(x1, x2), s)
    evals += 2
    assert (f1 > 0) != (f2 > 0), f"k={k}: no sign change around the computed root"
    lam = s + x
    if 1 < k <= N:
        assert P[k - 1] < lam < P[k - 2], f"k={k}: root outside its interlacing interval"
    elif k == 1:
        assert lam > P[0]
    else:
        assert lam < P[N - 1]
    _, _, S2 = _f_mp(d, x, s)
    nrm = mp.sqrt(1 + S2) if form == "arrow" else mp.sqrt(S2)

    def two(v):
        h = float(v)
        return h, float(v - mp.mpf(h))
    vals = []
    for r in rows:
        if r == N:                           # the arrow's last entry
            vals.append(float(-1 / nrm))
        else:
            vals.append(float(_G["w"][r] / (d[r] - x) / nrm))
    return (k, i) + two(x) + two(lam) + two(nrm) + (np.array(vals),) + (evals,)


def _rows_for(k, i, N, nvec, rng, nrand=24):
    near = [i + o for o in range(-3, 4) if 0 <= i + o < N]
    ends = [0, 1, N - 2, N - 1] + ([N] if nvec == N + 1 else [])
    rand = list(rng.integers(0, nvec, nrand))
    rows = sorted(set(int(r) for r in near + ends + rand if 0 <= r < nvec))
    return rows


def solve_case(form, poles_f, w_f, top_mp, ks, nvec, label, poles_mp=None, w_mp=None, procs=None):
    """poles_f/w_f: float arrays (exact inputs) unless poles_mp/w_mp give exact mp values (HalfArrow squares)."""
    mp.mp.prec = PREC
    P = poles_mp if poles_mp is not None else [mp.mpf(float(v)) for v in poles_f]
    w = w_mp if w_mp is not None else [mp.mpf(float(v)) for v in w_f]
    W2 = [v * v for v in w]
    N = len(P)
    rng = np.random.default_rng(12345)
    # shift pole guess for the row sample: k-1 or k-2 -> take both neighbourhoods
    jobs = []
    for k in ks:
        i0 = min(max(k - 1, 0), N - 1)
        rows = sorted(set(_rows_for(k, i0, N, nvec, rng) + _rows_for(k, max(i0 - 1, 0), N, nvec, rng, 0)))
        jobs.append((int(k), rows))
    t0 = time.time()
    procs = procs or min(os.cpu_count() or 1, 16)
    with mproc.Pool(procs, initializer=_init, initargs=(form, P, W2, top_mp, w)) as pool:
        res = pool.map(_solve_one, jobs, chunksize=max(1, len(jobs) // (procs * 8)))
    print(f"  {label}: {len(ks)} eigenpairs certified in {time.time() - t0:.1f} s, "
          f"{np.mean([r[-1] for r in res]):.1f} mp evaluations each", flush=True)
    nr = max(len(j[1]) for j in jobs)
    rows = np.full((len(ks), nr), -1, np.int64)
    vals = np.zeros((len(ks), nr))
    for c, (j, r) in enumerate(zip(jobs, res)):
        rows[c, :len(j[1])] = j[1]
        vals[c, :len(j[1])] = r[8]
    out = {"k": np.array([r[0] for r in res], np.int64), "i": np.array([r[1] for r in res], np.int64)}
    for name, col in (("x_hi", 2), ("x_lo", 3), ("lam_hi", 4), ("lam_lo", 5), ("nrm_hi", 6), ("nrm_lo", 7)):
        out[name] = np.array([r[col] for r in res])
    out["rows"] = rows
    out["vals"] = vals
    return out


def sample_ks(kmax, count, seed, hard=()):
    rng = np.random.default_rng(seed)
    base = [1, 2, 3, kmax - 2, kmax - 1, kmax]
    ks = set(k for k in base if 1 <= k <= kmax) | set(int(k) for k in hard)
    while len(ks) < min(count, kmax):
        ks.add(int(rng.integers(1, kmax + 1)))
    return np.array(sorted(ks), np.int64)


def hard_ks_arrow(D, w2, top, form, count):
    """eigenpairs whose eigenvalue sits closest to a pole relative to the pole spacing (these are the ones with a
    large K_nu: Remedy 3 in the reference).  Estimated with a cheap float64 secular bisection, vectorised over k."""
    N = D.size
    step = max(1, N // 2048)                          # every step-th interior interval is enough to find hard ones
    ks = np.arange(2, N + 1, step)
    a_, b_ = D[ks - 1].copy(), D[ks - 2].copy()
    for _ in range(70):
        m = (a_ + b_) / 2
        S = np.concatenate([(w2[None, :] / (D[None, :] - m[c:c + 128, None])).sum(axis=1) for c in range(0, m.size, 128)])
        f = (top - m) - S if form == "arrow" else 1 + top * S
        right = (f > 0) if form == "arrow" else (f < 0)
        a_ = np.where(right, m, a_)
        b_ = np.where(right, b_, m)
    lam = (a_ + b_) / 2
    gap = np.minimum(lam - D[ks - 1], D[ks - 2] - lam) / (D[ks - 2] - D[ks - 1])
    order = np.argsort(gap)
    return ks[order[:count]]


# ---- the cases ------------------------------------------------------------------------------------------------
def case_symarrow(n, count):
    D, z, a = gen_symarrow(n)
    hard = hard_ks_arrow(D, z * z, a, "arrow", count // 4) if count < n else ()
    ks = sample_ks(n, count, 7, hard)
    out = solve_case("arrow", D, z, mp.mpf(a), ks, n, f"symarrow n={n}")
    out.update(kind=np.array("symarrow"), n=np.array(n), seed=np.array(421))
    return out


def case_symdpr1(n, count):
    D, u, rho = gen_symdpr1(n)
    hard = hard_ks_arrow(D, u * u, rho, "dpr1", count // 4) if count < n else ()
    ks = sample_ks(n, count, 8, hard)
    out = solve_case("dpr1", D, u, mp.mpf(rho), ks, n, f"symdpr1 n={n}")
    out.update(kind=np.array("symdpr1"), n=np.array(n), seed=np.array(421))
    return out


def case_halfarrow(nd, count):
    mp.mp.prec = PREC
    D, z = gen_halfarrow(nd, 1)
    Dm = [mp.mpf(float(v)) for v in D]
    zm = [mp.mpf(float(v)) for v in z]
    P = [v * v for v in Dm]                            # exact squares
    w = [zl * dl for zl, dl in zip(zm[:nd], Dm)]       # exact z.*D
    top = mp.fsum(v * v for v in zm)                   # exact z'z
    hard = hard_ks_arrow(D * D, (z[:nd] * D) ** 2, float(top), "arrow", count // 4)
    ks = sample_ks(nd + 1, count, 9, hard)
    out = solve_case("arrow", None, None, top, ks, nd + 1, f"halfarrow nd={nd} tip", poles_mp=P, w_mp=w)
    # singular value sigma = sqrt(lambda) as (hi, lo)
    sg_hi, sg_lo = [], []
    for lh, ll in zip(out["lam_hi"], out["lam_lo"]):
        sg = mp.sqrt(mp.mpf(float(lh)) + mp.mpf(float(ll)))
        sg_hi.append(float(sg)); sg_lo.append(float(sg - mp.mpf(float(sg))))
    out.update(sig_hi=np.array(sg_hi), sig_lo=np.array(sg_lo), kind=np.array("halfarrow"), n=np.array(nd + 1), seed=np.array(421))
    return out


CASES = {
    "symarrow_n1000": lambda: case_symarrow(1000, 1000),          # BASELINE configs[0]: every eigenpair
    "symarrow_n32768": lambda: case_symarrow(32768, 256),         # configs[1]
    "symdpr1_n16384": lambda: case_symdpr1(16384, 256),           # configs[2]
    "halfarrow_n16384": lambda: case_halfarrow(16383, 256),       # configs[3], length(z) = length(D)+1 = 16384
    "symarrow_n65536": lambda: case_symarrow(65536, 64),          # the north-star target size
}


def main(argv):
    names = argv or list(CASES)
    for name in names:
        print(name, flush=True)
        out = CASES[name]()
        np.savez_compressed(os.path.join(HERE, f"truth_{name}.npz"), prec_bits=np.array(PREC), **out)


if __name__ == "__main__":
    main(sys.argv[1:])
