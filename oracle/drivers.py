"""Host-side drivers of the reference, restated in numpy (TEST INFRASTRUCTURE ONLY).

Each function follows the reference driver it cites: stable descending sort, exact-zero deflation in z,
exact-duplicate deflation in D with Givens rotations, the per-k loop (delegated to the C restatement in
arrowhead_oracle.c), the final re-sort and the row/column back-permutation.  Indices are 0-based here.

Reference quirks kept on purpose:
  * the driver-level tolerance vector is never forwarded to the per-k solver (arrowhead3.jl:631,643;
    arrowhead4.jl:521,534; arrowhead6.jl:627,640) -> the per-k default [1e3, 10*len(D_reduced), 1e3, 1e3, 1e3]
    is always what runs;
  * `lmul!(R[l][1],U)` applies R, not R' (arrowhead3.jl:634-635).
Reference code that cannot run as written (pre-1.0 syntax or undefined names: arrowhead4.jl:481,505;
arrowhead6.jl:554,569-584,617) is restated by its evident intent and the result carries
``ref_throws=True`` so tests can tell the two classes apart.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .lib import get_lib

EPS = np.finfo(np.float64).eps


def default_tau(npoles: int) -> np.ndarray:
    """Per-k default tolerances [tolb,tolz,tolnu,tolrho,tollambda] (arrowhead3.jl:420, arrowhead4.jl:290,
    arrowhead6.jl:345): tolz = 10*length(A.D) of the problem actually handed to the per-k solver."""
    return np.array([1e3, 10.0 * npoles, 1e3, 1e3, 1e3])


def jl_sortperm_desc(x) -> np.ndarray:
    """`sortperm(x, rev=true)`: stable, descending in `isless` order (0.0 before -0.0), ties by index."""
    x = np.asarray(x, np.float64)
    return np.lexsort((np.arange(x.size), np.signbit(x).astype(np.int8), -x))


def givens(f: float, g: float, i1: int, i2: int):
    """`LinearAlgebra.givens(f,g,i1,i2)` -> (i1,i2,c,s,r) with G*[f;g]=[r;0] (classic dlartg algorithm,
    unscaled branch; the scaled branches only trigger outside [~1e-146, ~1e146])."""
    if g == 0:
        c, s, r = 1.0, 0.0, f
    elif f == 0:
        c, s, r = 0.0, 1.0, g
    else:
        r = float(np.sqrt(f * f + g * g))
        c, s = f / r, g / r
        if abs(f) > abs(g) and c < 0:
            c, s, r = -c, -s, -r
    if i1 > i2:
        s = -s
        i1, i2 = i2, i1
    return i1, i2, c, s, r


def _apply_givens_rows(U, i1, i2, c, s, rotation="reference"):
    """`lmul!(G,U)` (rotation="reference": what arrowhead3.jl:635 executes, G*U) or G'*U
    (rotation="transpose": what the comment at :634 says and what the algebra needs -- with the
    reference's choice the deflated eigenvectors do not satisfy A*U = U*Lambda)."""
    if rotation == "transpose":
        s = -s
    a1 = U[i1, :].copy()
    a2 = U[i2, :].copy()
    U[i1, :] = c * a1 + s * a2
    U[i2, :] = -s * a1 + c * a2


@dataclass
class Info:
    sind: np.ndarray
    kb: np.ndarray
    kz: np.ndarray
    knu: np.ndarray
    krho: np.ndarray
    qout: np.ndarray
    steps: np.ndarray
    status: np.ndarray

    @staticmethod
    def zeros(n):
        return Info(np.zeros(n, np.int64), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n),
                    np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int32))

    def set_from(self, dst, res, c):
        for name in ("sind", "kb", "kz", "knu", "krho", "qout", "steps", "status"):
            getattr(self, name)[dst] = getattr(res, name)[c]

    def permuted(self, p):
        return Info(*(getattr(self, name)[p].copy()
                      for name in ("sind", "kb", "kz", "knu", "krho", "qout", "steps", "status")))


@dataclass
class EigenResult:
    values: np.ndarray
    vectors: np.ndarray
    info: Info
    perm: dict = field(default_factory=dict)   # is / es / rows: the ordering + deflation permutations
    ref_throws: bool = False


@dataclass
class SVDResult:
    U: np.ndarray
    S: np.ndarray
    Vt: np.ndarray
    info: Info
    perm: dict = field(default_factory=dict)
    ref_throws: bool = False


# ------------------------------------------------------------------------------------------ SymArrow
def eigen_symarrow(D, z, a, i, lib=None, nthreads=0, rotation="reference") -> EigenResult:
    """eigen(A::SymArrow,tau), arrowhead3.jl:567-657.  `i` is the 1-based arrow-top position."""
    lib = lib or get_lib()
    D = np.asarray(D, np.float64)
    z = np.asarray(z, np.float64)
    a = float(a)
    N = D.size
    n = n0 = N + 1
    is_ = jl_sortperm_desc(D)
    Ds, zs = D[is_].copy(), z[is_].copy()
    U = np.eye(n0, order="F")
    lam = np.zeros(n0)
    info = Info.zeros(n0)
    if n0 == 1:
        return EigenResult(np.array([a]), U, info)
    z0 = np.flatnonzero(zs == 0)
    zx = np.flatnonzero(zs != 0)
    if zx.size == 0:
        lam = np.concatenate([D, [a]])
        p = jl_sortperm_desc(lam)
        return EigenResult(lam[p], U[:, p], info, {"es": p})
    if z0.size:
        lam[z0] = Ds[z0]
        Ds, zs = Ds[zx], zs[zx]
        n = zs.size + 1
    zx = np.concatenate([zx, [n0 - 1]])
    g = Ds[:-1] - Ds[1:]
    g0 = np.flatnonzero(g == 0)
    gx = np.flatnonzero(g != 0)
    if g0.size:
        rots = [None] * g0.size
        for l in range(g0.size - 1, -1, -1):
            p = g0[l]
            i1, i2, c, s, r = givens(zs[p], zs[p + 1], int(zx[p]), int(zx[p + 1]))
            rots[l] = (i1, i2, c, s)
            zs[p], zs[p + 1] = r, 0.0
            lam[zx[p + 1]] = Ds[p + 1]
        keep = np.concatenate([[0], gx + 1]).astype(np.int64)
        nn = keep.size
        zxx = zx[np.concatenate([keep, [n - 1]])]
        res = lib.symarrow(Ds[keep], zs[keep], a, default_tau(nn), nthreads=nthreads)
        for c in range(nn + 1):
            lam[zxx[c]] = res.lam[c]
            U[zxx, zxx[c]] = res.U[:, c]
            info.set_from(zxx[c], res, c)
        for (i1, i2, c, s) in rots:
            _apply_givens_rows(U, i1, i2, c, s, rotation)
    else:
        res = lib.symarrow(Ds, zs, a, default_tau(n - 1), nthreads=nthreads)
        if z0.size == 0:
            lam[:] = res.lam
            U = res.U
            for name in ("sind", "kb", "kz", "knu", "krho", "qout", "steps", "status"):
                getattr(info, name)[:] = getattr(res, name)
        else:
            for c in range(n):
                lam[zx[c]] = res.lam[c]
                U[zx, zx[c]] = res.U[:, c]
                info.set_from(zx[c], res, c)
    isi = np.argsort(is_, kind="stable")
    es = jl_sortperm_desc(lam)
    rows = np.concatenate([isi[: i - 1], [n0 - 1], isi[i - 1:]]).astype(np.int64)
    return EigenResult(lam[es], U[rows][:, es], info.permuted(es), {"is": is_, "es": es, "rows": rows})


# ------------------------------------------------------------------------------------------ SymDPR1
def eigen_symdpr1(D, u, r, lib=None, nthreads=0, rotation="reference") -> EigenResult:
    """eigen(A::SymDPR1,tau), arrowhead4.jl:438-551."""
    lib = lib or get_lib()
    D = np.asarray(D, np.float64)
    u = np.asarray(u, np.float64)
    r = float(r)
    n = n0 = D.size
    signr = 1.0 if r > 0.0 else -1.0
    Dn = signr * D
    is_ = jl_sortperm_desc(Dn)
    Ds, zs = Dn[is_].copy(), u[is_].copy()
    rho = signr * r
    U = np.eye(n0, order="F")
    lam = np.zeros(n0)
    info = Info.zeros(n0)
    ref_throws = False
    if n0 == 1:                                     # arrowhead4.jl:456-471
        if Ds[0] == 0 and (rho == 0 or zs[0] == 0):
            return EigenResult(np.array([0.0]), U, info)
        L = D[0] + r * u[0] ** 2
        KD = (abs(D[0]) + abs(r) * u[0] ** 2) / abs(L)
        if KD > 1e3:
            a_hi, a_lo = lib.dd_op("mul", (u[0], 0.0), (u[0], 0.0))
            b_hi, b_lo = lib.dd_op("mul", (r, 0.0), (a_hi, a_lo))
            c_hi, c_lo = lib.dd_op("add", (D[0], 0.0), (b_hi, b_lo))
            L = c_hi + c_lo
        info.qout[0] = 1
        return EigenResult(np.array([L]), U, info)
    z0 = np.flatnonzero(zs == 0)
    zx = np.flatnonzero(zs != 0)
    if zx.size == 0:                                # arrowhead4.jl:476-482 returns undefined names
        p = jl_sortperm_desc(D)
        return EigenResult(D[p], U[:, p], info, {"es": p}, ref_throws=True)
    if z0.size:
        lam[z0] = Ds[z0]
        Ds, zs = Ds[zx], zs[zx]
        n = zs.size
    if n == 1:
        # reduced 1x1 problem: the reference would index D[0:-1] fine but per-k needs n>=2
        lam[zx[0]] = Ds[0] + rho * zs[0] ** 2
        ref_throws = True
    else:
        g = Ds[:-1] - Ds[1:]
        g0 = np.flatnonzero(g == 0)
        gx = np.flatnonzero(g != 0)
        if g0.size:                                 # arrowhead4.jl:501-526 (`Array(Tuple...)` throws)
            ref_throws = True
            rots = [None] * g0.size
            for l in range(g0.size - 1, -1, -1):
                p = g0[l]
                i1, i2, c, s, rr = givens(zs[p], zs[p + 1], int(zx[p]), int(zx[p + 1]))
                rots[l] = (i1, i2, c, s)
                zs[p], zs[p + 1] = rr, 0.0
                lam[zx[p + 1]] = Ds[p + 1]
            keep = np.concatenate([[0], gx + 1]).astype(np.int64)
            nn = keep.size
            zxx = zx[keep]
            if nn == 1:
                lam[zxx[0]] = Ds[keep][0] + rho * zs[keep][0] ** 2
            else:
                res = lib.symdpr1(Ds[keep], zs[keep], rho, default_tau(nn), nthreads=nthreads)
                for c in range(nn):
                    lam[zxx[c]] = res.lam[c]
                    U[zxx, zxx[c]] = res.U[:, c]
                    info.set_from(zxx[c], res, c)
            for (i1, i2, c, s) in rots:
                _apply_givens_rows(U, i1, i2, c, s, rotation)
        else:
            res = lib.symdpr1(Ds, zs, rho, default_tau(n), nthreads=nthreads)
            for c in range(n):
                lam[zx[c]] = res.lam[c]
                U[zx, zx[c]] = res.U[:, c]
                info.set_from(zx[c], res, c)
    isi = np.argsort(is_, kind="stable")
    lam = signr * lam
    es = jl_sortperm_desc(lam)
    return EigenResult(lam[es], U[isi][:, es], info.permuted(es), {"is": is_, "es": es, "rows": isi},
                       ref_throws=ref_throws)


# ------------------------------------------------------------------------------------------ HalfArrow
def svd_halfarrow(D, z, lib=None, nthreads=0, rotation="transpose") -> SVDResult:
    """svd(A::HalfArrow,tau), arrowhead6.jl:526-661 (no-deflation path exact; deflation by intent)."""
    lib = lib or get_lib()
    D = np.asarray(D, np.float64)
    z = np.asarray(z, np.float64)
    nd, n0 = D.size, z.size
    tip = n0 == nd + 1
    if not tip and n0 != nd:
        raise ValueError("length(z) must be length(D) or length(D)+1")
    signD = np.sign(D)
    Da = np.abs(D)
    is_ = jl_sortperm_desc(Da)
    Ds, signD = Da[is_].copy(), signD[is_].copy()
    zs = z[is_].copy()
    U = np.eye(n0, order="F")
    V = np.eye(n0 + 1, n0, order="F") if not tip else np.eye(n0, order="F")
    S = np.zeros(n0)
    info = Info.zeros(n0)
    ref_throws = False
    z0 = np.flatnonzero(zs == 0)
    zx = np.flatnonzero(zs != 0)
    if z0.size:                                     # arrowhead6.jl:569-584 skips the solve: intent
        ref_throws = True
        S[z0] = Ds[z0]
        V[z0, z0] = signD[z0]
        Ds, zs = Ds[zx], zs[zx]
    if zx.size == 0:
        if tip:
            S[n0 - 1] = abs(z[n0 - 1])
            V[n0 - 1, n0 - 1] = np.sign(z[n0 - 1])
    else:
        sgn = signD[zx]
        if tip:
            zx = zxv = np.concatenate([zx, [n0 - 1]])
            zs = np.concatenate([zs, [z[n0 - 1]]])
        else:
            zxv = np.concatenate([zx, [n0]])
        ndr = Ds.size
        # Exact-duplicate deflation, BY INTENT.  The reference's branch (arrowhead6.jl:586-632) cannot run: its gap
        # vector `D[1:n-2]-D[2:n-1]` misses the last pair of the no-tip shape, `zxxv=[zxx,n0+1]` nests a vector in a
        # vector, the reduced HalfArrow drops the tip and the loop runs to nn+1.  What it is after -- rotate the two
        # rows of a duplicate |D| so that one border entry vanishes, solve the reduced problem, rotate U AND V back --
        # is done here for every duplicate pair.
        g = Ds[:-1] - Ds[1:]
        g0 = np.flatnonzero(g == 0)
        if g0.size:
            ref_throws = True
            gx = np.flatnonzero(g != 0)
            rots = [None] * g0.size
            for l in range(g0.size - 1, -1, -1):
                p = g0[l]
                i1, i2, c, s, rr = givens(zs[p], zs[p + 1], int(zx[p]), int(zx[p + 1]))
                rots[l] = (i1, i2, c, s)
                zs[p], zs[p + 1] = rr, 0.0
                S[zx[p + 1]] = Ds[p + 1]
                V[zx[p + 1], zx[p + 1]] = sgn[p + 1]
            keep = np.concatenate([[0], gx + 1]).astype(np.int64)
            zk = np.concatenate([zs[keep], [zs[-1]]]) if tip else zs[keep]
            rows_u = np.concatenate([zx[keep], [n0 - 1]]) if tip else zx[keep]
            rows_v = np.concatenate([zx[keep], [n0 - 1 if tip else n0]])
            res = lib.halfarrow(Ds[keep], zk, default_tau(keep.size), nthreads=nthreads)
            for c in range(rows_u.size):
                col = rows_u[c]
                S[col] = res.lam[c]
                U[rows_u, col] = res.U[:, c]
                V[rows_v, col] = res.V[:, c]
                V[rows_v[:-1], col] *= sgn[keep]
                info.set_from(col, res, c)
            # A = B*diag(signD,1) with B = HalfArrow(|D|,z): B = G'*Bnew*G~ for a duplicate |D| pair, so U = G'*Unew and
            # W = G~'*Wnew; the stored V = diag(signD,1)*W carries the signs already, hence s -> s*signD[i1]*signD[i2]
            for (i1, i2, c, s) in rots:
                _apply_givens_rows(U, i1, i2, c, s, rotation)
                _apply_givens_rows(V, i1, i2, c, s * signD[i1] * signD[i2], rotation)
        else:
            res = lib.halfarrow(Ds, zs, default_tau(ndr), nthreads=nthreads)
            for c in range(zx.size):
                col = zx[c]
                S[col] = res.lam[c]
                U[zx, col] = res.U[:, c]
                V[zxv, col] = res.V[:, c]
                V[zxv[:-1], col] *= sgn
                info.set_from(col, res, c)
    isi = np.argsort(is_, kind="stable")
    es = jl_sortperm_desc(S)
    if not tip:
        rows_u, rows_v = isi, np.concatenate([isi, [n0]])
    else:
        rows_u = rows_v = np.concatenate([isi, [n0 - 1]])
    Uo = U[rows_u][:, es]
    Vo = V[rows_v][:, es]
    return SVDResult(np.ascontiguousarray(Uo), S[es], np.ascontiguousarray(Vo.T), info.permuted(es),
                     {"is": is_, "es": es, "rows_u": rows_u, "rows_v": rows_v}, ref_throws=ref_throws)


# ------------------------------------------------------------------------------------------ tdc
def tdc(dv, ev, lib=None, nthreads=0, rotation="reference"):
    """tdc(T::SymTridiagonal), arrowhead7.jl:10-36 -> (values, vectors)."""
    dv = np.asarray(dv, np.float64)
    ev = np.asarray(ev, np.float64)
    n = dv.size
    if n == 1:
        return np.array([dv[0]]), np.ones((1, 1))
    if n == 2:
        E = eigen_symarrow([dv[0]], [ev[0]], dv[1], 2, lib=lib, nthreads=nthreads, rotation=rotation)
        return E.values, E.vectors
    k = n // 2
    L1, U1 = tdc(dv[:k], ev[: k - 1], lib, nthreads, rotation)
    L2, U2 = tdc(dv[k + 1:], ev[k + 1:], lib, nthreads, rotation)
    Dm = np.concatenate([L1, L2])
    zm = np.concatenate([U1[-1, :] * ev[k - 1], U2[0, :] * ev[k]])
    E = eigen_symarrow(Dm, zm, dv[k], k + 1, lib=lib, nthreads=nthreads, rotation=rotation)
    W = np.array(E.vectors, order="F", copy=True)
    W[:k, :] = U1 @ W[:k, :]
    W[k + 1:, :] = U2 @ W[k + 1:, :]
    return E.values, W


# ------------------------------------------------------------------------------------------ Hermitian wrappers
def _phase_split(zc):
    """hermitian.jl:108-110 / 132-135: c = z./abs.(z); real border real(conj(c).*z) (components summed left to right)."""
    zc = np.asarray(zc, np.float64)
    if zc.shape[1] == 2:
        mod = np.hypot(zc[:, 0], zc[:, 1])
    else:
        mod = np.sqrt(((zc[:, 0] * zc[:, 0] + zc[:, 1] * zc[:, 1]) + zc[:, 2] * zc[:, 2]) + zc[:, 3] * zc[:, 3])
    c = zc / mod[:, None]
    zr = c[:, 0] * zc[:, 0] + c[:, 1] * zc[:, 1]
    for q in range(2, zc.shape[1]):
        zr = zr + c[:, q] * zc[:, q]
    return c, zr


def _scale_rows(c, U):
    H = c[:, None, :] * U[:, :, None]                      # Diagonal(c) * U, component by component
    return H[:, :, 0] + 1j * H[:, :, 1] if c.shape[1] == 2 else H


def eigen_hermarrow(D, zc, a, i, **kw) -> EigenResult:
    """eigen(A::HermArrow), hermitian.jl:107-115.  zc: (n, 2|4) real components of the border."""
    c, zr = _phase_split(zc)
    E = eigen_symarrow(D, zr, a, i, **kw)
    one = np.zeros((1, c.shape[1])); one[0, 0] = 1.0
    phase = np.concatenate([c[: i - 1], one, c[i - 1:]])
    return EigenResult(E.values, _scale_rows(phase, np.asarray(E.vectors)), E.info, E.perm, E.ref_throws)


def eigen_hermdpr1(D, uc, r, **kw) -> EigenResult:
    """eigen(A::HermDPR1), hermitian.jl:131-139."""
    c, ur = _phase_split(uc)
    E = eigen_symdpr1(D, ur, r, **kw)
    return EigenResult(E.values, _scale_rows(c, np.asarray(E.vectors)), E.info, E.perm, E.ref_throws)
