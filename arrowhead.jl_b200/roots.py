"""rootsah(pol, D, tau) -- roots of a real polynomial with real distinct roots (src/arrowhead7.jl:53-120), Float64 branch.

Host side (what Julia keeps): the arrowhead companion matrix in barycentric form, built in Dekker double-double
exactly as the reference does (Horner values s = p(D), the products t_j = prod_{i != j}(D_j - D_i), the arrow top
alpha and the border z_j = sqrt(-s_j/(t_j p_1)), arrowhead7.jl:66-110).  Device side: the eigenvalue-only per-k solver
with a double-double border (`ah_symarrow_eigvals_dd`), which replaces the reference's serial `for k=1:n` loop (:116-118).

The double-double helpers follow src/DoubleDouble.jl operation by operation (add2 :112-116, sub :129-133, mul12/mul2
:137-149 with the Veltkamp split :34-42, div2 :152-157, sqrt2 :160-168); numpy evaluates each binary operation with
one rounding, so the results are the ones Julia gets.
"""
from __future__ import annotations

import numpy as np

from ._lib import Context, get_context

_SPLIT = 134217729.0      # half64 = 2^27 + 1 (DoubleDouble.jl:26)


def _renorm(u, v):                          # double(u,v), DoubleDouble.jl:93-96
    w = u + v
    return w, (u - w) + v


def dd_add(x, y):
    xh, xl = x; yh, yl = y
    r = xh + yh
    s = np.where(np.abs(xh) > np.abs(yh), (((xh - r) + yh) + yl) + xl, (((yh - r) + xh) + xl) + yl)
    return _renorm(r, s)


def dd_sub(x, y):
    xh, xl = x; yh, yl = y
    r = xh - yh
    s = np.where(np.abs(xh) > np.abs(yh), (((xh - r) - yh) - yl) + xl, (((-yh - r) + xh) + xl) - yl)
    return _renorm(r, s)


def _split(x):
    p = x * _SPLIT
    h = (x - p) + p
    return h, x - h


def _mul12(x, y):
    hx, lx = _split(x); hy, ly = _split(y)
    z = x * y
    return z, (((hx * hy - z) + hx * ly) + lx * hy) + lx * ly


def dd_mul(x, y):
    xh, xl = x; yh, yl = y
    ch, cl = _mul12(xh, yh)
    cc = (xh * yl + xl * yh) + cl
    return _renorm(ch, cc)


def dd_div(x, y):
    xh, xl = x; yh, yl = y
    c = xh / yh
    uh, ul = _mul12(c, yh)
    cc = ((((xh - uh) - ul) + xl) - c * yl) / yh
    return _renorm(c, cc)


def dd_sqrt(x):
    xh, xl = x
    if np.any(np.asarray(xh) <= 0):
        raise ValueError("sqrt of a non-positive double-double: D does not interlace the roots of the polynomial")
    c = np.sqrt(xh)
    uh, ul = _mul12(c, c)
    cc = (((xh - uh) - ul) + xl) * 0.5 / c
    return _renorm(c, cc)


def dd_neg(x):
    return -x[0], -x[1]


def companion_arrow(coeffs, D):
    """(D_sorted, z_hi, z_lo, alpha_hi, alpha_lo) of the arrowhead companion matrix of the polynomial with coefficients
    `coeffs` (ascending powers, like Polynomial(...) in the reference) for the barycentric points D (arrowhead7.jl:59-110)."""
    Dm = np.sort(np.asarray(D, np.float64))[::-1].copy()
    p = np.asarray(coeffs, np.float64)[::-1].copy()             # p[0] = leading coefficient
    n = p.size - 1
    if Dm.size != n - 1:
        raise ValueError("need degree-1 barycentric points")
    zero = np.zeros(n - 1)
    DD = (Dm, zero.copy())
    p1 = (np.float64(p[0]), np.float64(0.0))
    s = dd_mul((np.full(n - 1, p[0]), zero.copy()), (np.ones(n - 1), zero.copy()))      # s = pD[1]*ones
    for i in range(1, n + 1):                                   # Horner in double-double
        r = dd_mul(s, DD)
        s = dd_add(r, (np.full(n - 1, p[i]), zero.copy()))
    th, tl = np.ones(n - 1), np.zeros(n - 1)                    # t_j = prod_{i != j} (D_j - D_i), i ascending
    for i in range(n - 1):
        g = dd_sub(DD, (np.full(n - 1, Dm[i]), zero.copy()))
        nh, nl = dd_mul((th, tl), g)
        keep = np.arange(n - 1) == i
        th = np.where(keep, th, nh); tl = np.where(keep, tl, nl)
    aD = dd_div((np.float64(p[1]), np.float64(0.0)), p1)
    for i in range(n - 1):
        aD = dd_add(aD, (np.float64(Dm[i]), np.float64(0.0)))
    aD = (-aD[0], -aD[1])
    zD = dd_sqrt(dd_div(dd_neg(s), dd_mul((th, tl), (np.full(n - 1, p[0]), zero.copy()))))
    return Dm, np.asarray(zD[0], np.float64), np.asarray(zD[1], np.float64), float(aD[0]), float(aD[1])


def rootsah(coeffs, D, tau=(1e2, 1e2, 1e2, 1e2, 1e2), *, ctx: Context | None = None, return_info=False):
    """rootsah(pol, D, tau) -> (roots, Qout).  `coeffs`: ascending powers; `D`: barycentric points interlacing the roots
    (e.g. the roots of the derivative)."""
    ctx = ctx or get_context()
    Dm, zh, zl, ah, al = companion_arrow(coeffs, D)
    z = zl + zh                                                   # z[i] = Float64(zD[i].lo + zD[i].hi)   (:108-110)
    zlo = (zh - z) + zl                                           # keep (z, zlo) equal to zD
    a = al + ah
    alo = (ah - a) + al
    res = ctx.symarrow_eigvals_dd(Dm, z, zlo, a, alo, np.asarray(tau, np.float64))
    return (res.lam, res.qout, res) if return_info else (res.lam, res.qout)
