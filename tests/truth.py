"""Loader for the mpmath truth fixtures (tests/golden/truth_*.npz, made by tests/golden/make_truth.py) and the
long-double derivation of complete truth eigenvectors from the stored (shift pole, x = lambda - pole, norm).

Nothing here touches oracle/ or the CUDA library: the truth is an independent third party that both are measured
against.  Errors are returned in units of eps = 2^-52, relative, entry by entry."""
from __future__ import annotations

import os

import numpy as np

from helpers import EPS, gen_halfarrow, gen_symarrow, gen_symdpr1

LD = np.longdouble
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def available(name: str) -> bool:
    return os.path.exists(os.path.join(GOLDEN, f"truth_{name}.npz"))


class Truth:
    """One fixture: the problem (regenerated from its seed) + certified eigenvalues + what is needed to form
    every eigenvector entry in long double (error ~ 2^-61 relative, see make_truth.py's header)."""

    def __init__(self, name: str):
        t = np.load(os.path.join(GOLDEN, f"truth_{name}.npz"))
        self.name = name
        self.kind = str(t["kind"])
        self.n = int(t["n"])
        self.k = t["k"]
        self.i = t["i"]
        self.x = t["x_hi"].astype(LD) + t["x_lo"].astype(LD)
        self.lam_arrow = t["lam_hi"].astype(LD) + t["lam_lo"].astype(LD)
        self.nrm = t["nrm_hi"].astype(LD) + t["nrm_lo"].astype(LD)
        self.rows, self.vals = t["rows"], t["vals"]
        seed = int(t["seed"])
        if self.kind == "symarrow":
            self.D, self.z, self.a = gen_symarrow(self.n, seed)
            self.lam = self.lam_arrow
        elif self.kind == "symdpr1":
            self.D, self.u, self.rho = gen_symdpr1(self.n, seed)
            self.lam = self.lam_arrow
        else:
            self.D, self.z = gen_halfarrow(self.n - 1, 1, seed)
            self.lam = t["sig_hi"].astype(LD) + t["sig_lo"].astype(LD)      # singular values
        self.index = {int(k): c for c, k in enumerate(self.k)}

    # -- pole differences d_l = P_l - P_i in long double (P = D, or D.^2 for the HalfArrow)
    def _d(self, c):
        Dl = self.D.astype(LD)
        Di = Dl[self.i[c]]
        if self.kind == "halfarrow":
            return (Dl - Di) * (Dl + Di)
        return Dl - Di

    def vector(self, c: int) -> np.ndarray:
        """truth eigenvector (right singular vector) of fixture entry c, long double, reference sign convention."""
        t = self._d(c) - self.x[c]
        if self.kind == "symarrow":
            w = self.z.astype(LD)
        elif self.kind == "symdpr1":
            w = self.u.astype(LD)
        else:
            w = self.z[: self.n - 1].astype(LD) * self.D.astype(LD)
        v = w / t / self.nrm[c]
        if self.kind != "symdpr1":
            v = np.concatenate([v, [-1 / self.nrm[c]]])
        return v

    def left_vector(self, c: int) -> np.ndarray:
        """HalfArrow: u = [sigma v_l / D_l ; z_n v_n / sigma] (arrowhead6.jl:421-425)."""
        assert self.kind == "halfarrow"
        v = self.vector(c)
        nd = self.n - 1
        sg = self.lam[c]
        return np.concatenate([sg * v[:nd] / self.D.astype(LD), [LD(self.z[nd]) * v[nd] / sg]])

    def residual(self, c: int) -> float:
        """||A v - lambda v|| / ||A||_F in long double -- ties the secular-equation truth back to the matrix itself."""
        v = self.vector(c)
        t = self._d(c) - self.x[c]                       # P_l - lambda
        if self.kind == "symarrow":
            z = self.z.astype(LD)
            top = t * v[:-1] + z * v[-1]
            a_l = (LD(self.a) - self.D.astype(LD)[self.i[c]]) - self.x[c]
            bot = (z * v[:-1]).sum() + a_l * v[-1]
            nA = np.sqrt((self.D ** 2).sum() + 2 * (self.z ** 2).sum() + self.a ** 2)
            return float(np.sqrt((top * top).sum() + bot * bot) / nA)
        if self.kind == "symdpr1":
            u = self.u.astype(LD)
            r = t * v + LD(self.rho) * u * (u * v).sum()
            return float(np.sqrt((r * r).sum()) / (np.sqrt((self.D ** 2).sum()) + self.rho * (self.u ** 2).sum()))
        nd = self.n - 1                                  # A'A v - sigma^2 v with A'A = arrow(D^2, z.*D, z'z)
        w = self.z[:nd].astype(LD) * self.D.astype(LD)
        top = t * v[:-1] + w * v[-1]
        zz = (self.z.astype(LD) ** 2).sum()
        a_l = (zz - self.D.astype(LD)[self.i[c]] ** 2) - self.x[c]
        bot = (w * v[:-1]).sum() + a_l * v[-1]
        nA = (self.D ** 2).sum() + (self.z ** 2).sum()
        return float(np.sqrt((top * top).sum() + bot * bot) / nA)


def err_eps(value, truth) -> np.ndarray:
    """|value - truth| / |truth| in eps, elementwise (truth long double; zeros of the truth must be matched exactly)."""
    value = np.asarray(value).astype(LD)
    truth = np.asarray(truth, dtype=LD)
    out = np.zeros(truth.shape, dtype=np.float64)
    nz = truth != 0
    out[nz] = (np.abs(value[nz] - truth[nz]) / np.abs(truth[nz]) / LD(EPS)).astype(np.float64)
    out[~nz] = np.where(value[~nz] == 0, 0.0, np.inf)
    return out
