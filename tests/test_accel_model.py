"""Executable model of the accelerated bisection of k_bisect (csrc/ah_pipe.cuh: accel_update /
accel_descend / accel_choose), run on the CPU in binary64.

The device code decides grid points of the reference's bisection (arrowhead3.jl:686-704) from certified bounds instead of
evaluating them.  This file restates that logic operation by operation on the inverse arrow that k_prep forms
(arrowhead3.jl:152-189, 671-680) and checks, against PLAIN bisection with the very same evaluation function:

  * the final bracket is bit-identical for every eigenpair of every family -- with the production probe policy, with a
    random probe policy (any point strictly inside the certified interval: the certificates, not the heuristic, carry the
    proof) and with the probes switched off;
  * the certificates are never wrong: every grid point the descent skips is re-evaluated and must decide the same way;
  * the number of O(N) sweeps drops to well under half.

The evaluation here is numpy's (pairwise) sum of the same terms; its error bound is smaller than the theta the device
uses, so the same certificates apply."""
import numpy as np
import pytest

EPS = 2.220446049250313e-16


class InverseArrow:
    """what k_prep hands to k_bisect for eigenpair k (1-based) of the ordered SymArrow (D, z, a)"""

    def __init__(self, D, z, a, k):
        N = D.size
        w2 = z * z
        if k == 1:
            right, i = True, 1
        elif k == N + 1:
            right, i = False, N
        else:
            Dk = D[k - 1]; mid = (D[k - 2] - Dk) / 2
            right = (a - Dk - mid - np.sum(w2 / ((D - Dk) - mid))) < 0          # arrowhead3.jl:437-441
            i = k if right else k - 1
        i0 = i - 1
        sigma = D[i0]; wz = 1 / z[i0]
        d = D - sigma
        mask = np.ones(N, bool); mask[i0] = False
        self.dm, self.w2m = d[mask], w2[mask]
        Dp = 1 / self.dm; zp = -z[mask] * Dp * wz
        P = np.sum((self.w2m * Dp)[Dp > 0]); Q = np.sum((self.w2m * Dp)[Dp <= 0])
        as_ = a - sigma
        if as_ < 0: P -= as_
        else: Q -= as_
        self.b = (P + Q) * wz * wz
        self.wz2 = wz * wz
        sabs = np.sum(np.abs(zp)) + abs(wz)
        bnd = max(np.max((1.0 if right else -1.0) * Dp + np.abs(zp)), abs(wz))
        mxD = 1 / d[i0 - 1] if i0 > 0 else 0.0
        mnD = 1 / d[i0 + 1] if i0 < N - 1 else 0.0
        self.sideR = right
        self.l0, self.r0 = (mxD, max(bnd, self.b + sabs)) if right else (min(-bnd, self.b - sabs), mnD)
        self.pole = self.l0 if right else self.r0
        self.sweeps = 0
        self.noise = 0.0

    def S(self, m):                                                                # one O(N) sweep
        self.sweeps += 1
        with np.errstate(all="ignore"):
            s = float(np.sum(self.w2m / (self.dm * (1 - m * self.dm))))
        if self.noise:     # a deterministic relative perturbation of known size (a stand-in for a sloppier summation)
            h = (hash(np.float64(m).tobytes()) % 2000001) / 1000000.0 - 1.0
            s *= 1.0 + self.noise * h
        return s

    def F(self, S, m):                                                             # arrowhead3.jl:696-701, k_bisect's order
        with np.errstate(all="ignore"):
            return (self.b - m) - (S * self.wz2 + self.wz2 / (0 - m))


def converged(l, r):
    return not ((r - l) > 2 * EPS * max(abs(l), abs(r)))


def plain_bisection(I):
    l, r = I.l0, I.r0
    while not converged(l, r):
        m = (l + r) / 2
        if I.F(I.S(m), m) > 0: l = m
        else: r = m
    return l, r


def accelerated_bisection(I, th2, policy="production", rng=None, audit=False):
    """mirror of the device code: returns the final bracket; `audit` re-evaluates every skipped grid point"""
    l, r = I.l0, I.r0
    lo, hi, Flo, Fhi = -np.inf, np.inf, np.inf, -np.inf
    hist = []                                  # up to three (x, F), newest last
    probing, nprobe, nfail = policy != "off", 0, 0
    p = I.pole
    while True:
        # ---- accel_descend
        while not converged(l, r):
            m = (l + r) / 2
            if m <= lo or m >= hi:
                if audit:
                    assert (I.F(I.S(m), m) > 0) == (m <= lo), "a certificate decided a grid point wrongly"
                    I.sweeps -= 1
                if m <= lo: l = m
                else: r = m
                continue
            break
        if converged(l, r):
            return l, r
        m = (l + r) / 2
        # ---- accel_choose
        x = m
        if probing and len(hist) >= 2 and nprobe < 24:
            a_, b_ = max(lo, l), min(hi, r)
            (x1, F1), (x2, F2) = hist[-2], hist[-1]
            with np.errstate(all="ignore"):
                slope = abs((F2 - F1) / (x2 - x1))
                L1, L2 = I.b - x1, I.b - x2
                R1, R2 = F1 - L1, F2 - L2
                delta = max(1.25 * th2 * abs(R2) / slope * 4.0 ** nfail, 4 * EPS * abs(x2))
                if not ((b_ - a_) > 6 * delta):
                    probing = False
                elif policy == "random":
                    x = a_ + (b_ - a_) * rng.uniform(0.01, 0.99); nprobe += 1
                    if not (a_ < x < b_): x = m
                else:
                    def root_with_pole(q):
                        gam = (R1 - R2) / (1 / (x1 - q) - 1 / (x2 - q)); alpha = R2 - gam / (x2 - q)
                        B, C = I.b + alpha + q, (I.b + alpha) * q - gam
                        sq = np.sqrt(B * B - 4 * C)
                        ra, rb = (B + sq) / 2, (B - sq) / 2
                        return ra if a_ < ra < b_ else rb
                    xs = np.nan
                    if len(hist) >= 3:
                        x0, F0 = hist[-3]; R0 = F0 - (I.b - x0)
                        rho = (R0 - R1) * (x2 - x1) / ((R1 - R2) * (x1 - x0))
                        qq = (x2 - rho * x0) / (1 - rho)
                        if ((qq - p) if I.sideR else (p - qq)) <= 0: xs = root_with_pole(qq)
                    if not (a_ < xs < b_): xs = root_with_pole(p)
                    if not (a_ < xs < b_) and lo >= l and hi <= r: xs = lo + (hi - lo) * Flo / (Flo - Fhi)
                    xp = xs - delta if (xs - a_) > (b_ - xs) else xs + delta
                    if a_ < xp < b_: x = xp; nprobe += 1
        # ---- the sweep, then accel_update
        S = I.S(x)
        grid = x == m
        if S != S:
            assert not grid, "NaN at a grid point: the device hands the eigenpair to k_full"
            probing = False
            continue
        F = I.F(S, x)
        if abs(S) > 1e-250 and abs(S) <= 1.7976931348623157e308:
            la, lb = I.F(S * (1 + th2), x) > 0, I.F(S * (1 - th2), x) > 0
            if la and lb:
                if x > lo: lo, Flo = x, F
            elif not la and not lb:
                if x < hi: hi, Fhi = x, F
            elif not grid:
                nfail += 1
                if nfail >= 3: probing = False
        if abs(F) <= 1.7976931348623157e308:
            if not hist or x != hist[-1][0]: hist = (hist + [(x, F)])[-3:]
        else:
            probing = False
        if grid:
            if F > 0: l = m
            else: r = m


def families(n, seed):
    rng = np.random.default_rng(seed); N = n - 1
    yield "uniform", np.sort(rng.random(N))[::-1].copy(), rng.random(N), float(rng.random())
    yield "normal", np.sort(rng.standard_normal(N))[::-1].copy(), rng.standard_normal(N), float(rng.standard_normal())
    yield "clustered", np.sort(np.concatenate([rng.random(N // 2), 0.5 + 1e-9 * rng.random(N - N // 2)]))[::-1].copy(), rng.random(N), float(rng.random())
    yield "graded", np.sort(2.0 ** (-40 * rng.random(N)))[::-1].copy(), 2.0 ** (-30 * rng.random(N)), float(rng.random())
    yield "tiny z", np.sort(rng.random(N))[::-1].copy(), 1e-8 * rng.random(N) + 1e-12, float(rng.random())
    yield "huge a", np.sort(rng.random(N))[::-1].copy(), rng.random(N), 1e12
    yield "all negative", np.sort(-rng.random(N) - 0.1)[::-1].copy(), rng.random(N) - 0.5, -3.0


def th2_of(N):
    return 2.01 * ((N + 1023) // 1024 + 4 + 14) * 2.0 ** -53          # bisect_th2 (ah_pipe.cuh): tile sums of 4, one rounding per tile


@pytest.mark.parametrize("n", [300, 2500])
def test_certified_skipping_reproduces_plain_bisection_bit_for_bit(n):
    rng = np.random.default_rng(5)
    plain_sweeps = fast_sweeps = 0
    for name, D, z, a in families(n, 421):
        if np.any(np.diff(D) >= 0) or np.any(z == 0):
            continue
        N = n - 1
        ks = sorted(set([1, 2, N, N + 1] + list(rng.integers(1, N + 2, 40 if n > 1000 else 80))))
        for k in ks:
            I = InverseArrow(D, z, a, int(k))
            if not np.isfinite([I.l0, I.r0, I.b]).all():
                continue
            ref = plain_bisection(I); plain_sweeps += I.sweeps; I.sweeps = 0
            got = accelerated_bisection(I, th2_of(N), audit=(k % 7 == 0)); fast_sweeps += I.sweeps; I.sweeps = 0
            assert got == ref, (name, k)
            assert accelerated_bisection(I, th2_of(N), policy="random", rng=rng) == ref, (name, k, "random probes")
            assert accelerated_bisection(I, th2_of(N), policy="off") == ref, (name, k, "no probes")
    assert fast_sweeps < 0.5 * plain_sweeps, (fast_sweeps, plain_sweeps)


def test_a_wrong_certificate_would_be_caught():
    """The audit is sharp.  With a sweep that carries a relative error of up to 1e-12, certificates that allow for it
    (theta >= the error bound) still reproduce plain bisection exactly, while certificates that do not go wrong."""
    D, z, a = next(families(800, 3))[1:]
    bad = 0
    for k in range(2, 400, 7):
        I = InverseArrow(D, z, a, k)
        I.noise = 1e-12
        ref = plain_bisection(I)
        assert accelerated_bisection(I, 2.01 * (1e-12 + th2_of(799)), audit=True) == ref, k
        try:
            if accelerated_bisection(I, 1e-22, audit=True) != ref:
                bad += 1
        except AssertionError:
            bad += 1
    assert bad > 0
