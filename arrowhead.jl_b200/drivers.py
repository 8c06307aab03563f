"""Host-side mirror of the reference's all-eigenpairs drivers.

In the drop-in picture Julia keeps this logic (types, stable sort, exact deflation, final permutation) and
only the `Threads.@threads for k` loop bodies become one call into libarrowhead_b200.so.  This module is
that same host logic in Python over the same C ABI, so the parity tests read like the reference's tests:

    eigen(A::SymArrow)   src/arrowhead3.jl:567-657      eigen(A::SymDPR1)  src/arrowhead4.jl:438-551
    svd(A::HalfArrow)    src/arrowhead6.jl:526-661      tdc(T)             src/arrowhead7.jl:10-36

Eigenvectors live on the GPU: U is allocated on the device, the kernels write every column in place
(`U[zx,zx[k]] = v` becomes a row/column-mapped store), the deflation rotations (`lmul!(R,U)`) and the
final `view(U,rows,es)` permutation run as device kernels, and a host copy is made only on request.

Reference behaviour kept on purpose (documented in DESIGN.md):
  * the tolerance vector passed to the driver is NOT forwarded to the per-k solver
    (arrowhead3.jl:631,643 ...) -> `forward_tau=False` by default;
  * deflation rotations are applied as R (arrowhead3.jl:635), although the algebra needs R'
    -> `rotation="reference"` by default, `"transpose"` gives the mathematically consistent vectors.
"""
from __future__ import annotations

import numpy as np

from ._lib import Context, get_context
from .types import SVD, Eigen, HalfArrow, HermArrow, HermDPR1, Info, SymArrow, SymDPR1


# ---------------------------------------------------------------------------------------------------
class DeviceMatrix:
    """Column-major float64 matrix in GPU memory owned through the C ABI (ah_device_malloc)."""

    def __init__(self, ctx: Context, nrows: int, ncols: int):
        import ctypes as C
        self.ctx, self.nrows, self.ncols, self.ld = ctx, int(nrows), int(ncols), int(nrows)
        p = C.c_void_p()
        ctx._check(ctx.lib.ah_device_malloc(ctx.handle, C.byref(p), self.ld * self.ncols * 8))
        self.ptr = p.value

    @property
    def shape(self):
        return (self.nrows, self.ncols)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.nrows, self.ncols), "typestr": "<f8", "data": (self.ptr, False), "version": 3,
                "strides": (8, 8 * self.ld)}

    def to_host(self) -> np.ndarray:
        out = np.empty((self.nrows, self.ncols), order="F")
        self.ctx.memcpy_d2h(out, self.ptr, self.ld * self.ncols * 8)
        return out

    def free(self):
        import ctypes as C
        if getattr(self, "ptr", None):
            self.ctx.lib.ah_device_free(self.ctx.handle, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def jl_sortperm_desc(x) -> np.ndarray:
    """sortperm(x, rev=true): stable; 0.0 sorts before -0.0 (isless order)."""
    x = np.asarray(x, np.float64)
    return np.lexsort((np.arange(x.size), np.signbit(x).astype(np.int8), -x))


def _is_identity(p) -> bool:
    return bool(np.array_equal(p, np.arange(p.size)))


def givens(f: float, g: float, i1: int, i2: int):
    """LinearAlgebra.givens(f,g,i1,i2) -> (i1,i2,c,s,r), G*[f;g] = [r;0]."""
    if g == 0:
        c, s, r = 1.0, 0.0, f
    elif f == 0:
        c, s, r = 0.0, 1.0, g
    else:
        r = float(np.hypot(f, g)) if not (1e-146 < max(abs(f), abs(g)) < 1e146) else float(np.sqrt(f * f + g * g))
        c, s = f / r, g / r
        if abs(f) > abs(g) and c < 0:
            c, s, r = -c, -s, -r
    if i1 > i2:
        s, i1, i2 = -s, i2, i1
    return i1, i2, c, s, r


def _finish(ctx, Udev, rows, cols, vectors, phase=None):
    """Apply the lazy `view(U,rows,cols)` of the reference and hand the vectors out.  phase (nrows x 2|4): the Hermitian
    wrappers' `Diagonal(c)*U` (hermitian.jl:107-115) fused into the same pass -- complex / quaternion entries are written once."""
    if Udev is None:
        return None
    if phase is not None:
        ncomp = phase.shape[1]
        out = DeviceMatrix(ctx, rows.size * ncomp, cols.size)
        ctx.gather_permuted_phase(out.ptr, rows.size, Udev.ptr, Udev.ld, rows, cols, phase)
        Udev.free()
        if vectors == "device":
            return out
        host = _hyper_to_host(out.to_host(), rows.size, ncomp)
        out.free()
        return host
    if not (_is_identity(rows) and _is_identity(cols) and rows.size == Udev.nrows and cols.size == Udev.ncols):
        out = DeviceMatrix(ctx, rows.size, cols.size)
        ctx.gather_permuted(out.ptr, out.ld, Udev.ptr, Udev.ld, rows, cols)
        Udev.free()
        Udev = out
    if vectors == "device":
        return Udev
    host = Udev.to_host()
    Udev.free()
    return host


def _deflate_duplicates(Ds, zs, zx, lam, rotation):
    """Exact-duplicate deflation in D (arrowhead3.jl:605-626): returns kept positions + rotations."""
    g = Ds[:-1] - Ds[1:]
    g0 = np.flatnonzero(g == 0)
    gx = np.flatnonzero(g != 0)
    rots = []
    if g0.size:
        rots = [None] * g0.size
        for l in range(g0.size - 1, -1, -1):
            p = g0[l]
            i1, i2, c, s, r = givens(zs[p], zs[p + 1], int(zx[p]), int(zx[p + 1]))
            rots[l] = (i1, i2, c, -s if rotation == "transpose" else s)
            zs[p], zs[p + 1] = r, 0.0
            lam[zx[p + 1]] = Ds[p + 1]
    keep = np.concatenate([[0], gx + 1]).astype(np.int64)
    return keep, rots


# ---------------------------------------------------------------------------------------------------
def eigen_symarrow(A: SymArrow, tau=None, *, ctx: Context | None = None, vectors="host", forward_tau=False,
                   rotation="reference", out: np.ndarray | None = None, phase: np.ndarray | None = None):
    """eigen(A::SymArrow, tau) -> (Eigen(values, vectors), Info).  vectors: "host" | "device" | None.
    out: optional caller-owned Fortran-ordered n x n float64 array for the eigenvectors (pinned memory makes the
    copy PCIe-fast: `Context.pinned_alloc`, torch `pin_memory`); when the input needs no deflation and no
    permutation the column blocks are streamed into it while the next block is being computed."""
    ctx = ctx or get_context()
    D, z, a = A.D, A.z, A.a
    N = D.size
    n = n0 = N + 1
    tau_k = np.asarray(tau, np.float64) if (forward_tau and tau is not None) else None
    is_ = jl_sortperm_desc(D)
    Ds, zs = D[is_].copy(), z[is_].copy()
    lam = np.zeros(n0)
    info = Info.zeros(n0)
    want = vectors is not None
    if n0 == 1:
        return Eigen(np.array([a]), np.ones((1, 1)) if want else None), info
    z0 = np.flatnonzero(zs == 0)
    zx = np.flatnonzero(zs != 0)
    if zx.size == 0:                       # diagonal matrix (arrowhead3.jl:585-591)
        lam = np.concatenate([D, [a]])
        p = jl_sortperm_desc(lam)
        return Eigen(lam[p], np.eye(n0)[:, p] if want else None), info
    if z0.size:
        lam[z0] = Ds[z0]
        Ds, zs = Ds[zx], zs[zx]
        n = zs.size + 1
    zx = np.concatenate([zx, [n0 - 1]])
    keep, rots = _deflate_duplicates(Ds, zs, zx, lam, rotation)
    deflated = bool(z0.size or rots)
    direct = (not deflated and out is not None and vectors == "host" and A.i == n0 and _is_identity(is_)
              and out.shape == (n0, n0) and out.dtype == np.float64 and out.flags.f_contiguous)
    Udev = DeviceMatrix(ctx, n0, n0) if (want and not direct) else None
    if deflated:
        rowmap = zx[np.concatenate([keep, [n - 1]])]
        if want:
            ctx.set_identity(Udev.ptr, n0, n0, Udev.ld)
        res = ctx.symarrow(Ds[keep], zs[keep], a, tau_k, vectors=False, U_dev=Udev.ptr if want else None,
                           ldu=n0, rowmap=rowmap if want else None)
        lam[rowmap] = res.lam
        info.scatter(rowmap, res)
        for (i1, i2, c, s) in rots:
            if want:
                ctx.apply_givens_rows(Udev.ptr, Udev.ld, n0, i1, i2, c, s)
    else:
        if direct:                         # ordered input: eigenvector blocks go straight into the caller's buffer
            res = ctx.symarrow(Ds, zs, a, tau_k, vectors=False, U_host=out.ctypes.data, ldu=n0)
            if _is_identity(jl_sortperm_desc(res.lam)):
                info.scatter(np.arange(n0), res)
                info.stats = res.stats
                return Eigen(res.lam, out), info
            Udev = DeviceMatrix(ctx, n0, n0)   # eigenvalues came out unsorted (never seen): take the general path
        res = ctx.symarrow(Ds, zs, a, tau_k, vectors=False, U_dev=Udev.ptr if want else None, ldu=n0)
        lam[:] = res.lam
        info.scatter(np.arange(n0), res)
    isi = np.argsort(is_, kind="stable")
    es = jl_sortperm_desc(lam)
    rows = np.concatenate([isi[: A.i - 1], [n0 - 1], isi[A.i - 1:]]).astype(np.int64)
    Uout = _finish(ctx, Udev, rows, es, vectors, phase)
    if out is not None and isinstance(Uout, np.ndarray):
        out[...] = Uout
        Uout = out
    E = Eigen(lam[es], Uout)
    out_info = info.permuted(es)
    out_info.stats = res.stats
    return E, out_info


def eigen_symdpr1(A: SymDPR1, tau=None, *, ctx: Context | None = None, vectors="host", forward_tau=False,
                  rotation="reference", phase: np.ndarray | None = None):
    """eigen(A::SymDPR1, tau) -> (Eigen, Info)."""
    ctx = ctx or get_context()
    D, u, r = A.D, A.u, A.r
    n = n0 = D.size
    tau_k = np.asarray(tau, np.float64) if (forward_tau and tau is not None) else None
    signr = 1.0 if r > 0.0 else -1.0
    Dn = signr * D
    is_ = jl_sortperm_desc(Dn)
    Ds, zs = Dn[is_].copy(), u[is_].copy()
    rho = signr * r
    lam = np.zeros(n0)
    info = Info.zeros(n0)
    want = vectors is not None
    if n0 == 1:                            # arrowhead4.jl:456-471 (closed form; double-double when cancelling)
        if Ds[0] == 0 and (rho == 0 or zs[0] == 0):
            return Eigen(np.array([0.0]), np.ones((1, 1)) if want else None), info
        L = D[0] + r * u[0] ** 2
        if (abs(D[0]) + abs(r) * u[0] ** 2) / abs(L) > 1e3:
            from fractions import Fraction
            L = float(Fraction(D[0]) + Fraction(r) * Fraction(u[0]) ** 2)
        info.qout[0] = 1
        return Eigen(np.array([L]), np.ones((1, 1)) if want else None), info
    z0 = np.flatnonzero(zs == 0)
    zx = np.flatnonzero(zs != 0)
    if zx.size == 0:
        p = jl_sortperm_desc(D)
        return Eigen(D[p], np.eye(n0)[:, p] if want else None), info
    if z0.size:
        lam[z0] = Ds[z0]
        Ds, zs = Ds[zx], zs[zx]
        n = zs.size
    keep, rots = (np.arange(n), []) if n == 1 else _deflate_duplicates(Ds, zs, zx, lam, rotation)
    deflated = bool(z0.size or rots)
    Udev = DeviceMatrix(ctx, n0, n0) if want else None
    res = None
    if deflated:
        rowmap = zx[keep]
        if want:
            ctx.set_identity(Udev.ptr, n0, n0, Udev.ld)
        if keep.size == 1:                 # reduced 1x1 problem
            lam[rowmap[0]] = Ds[keep][0] + rho * zs[keep][0] ** 2
        else:
            res = ctx.symdpr1(Ds[keep], zs[keep], rho, tau_k, vectors=False, U_dev=Udev.ptr if want else None,
                              ldu=n0, rowmap=rowmap if want else None)
            lam[rowmap] = res.lam
            info.scatter(rowmap, res)
        for (i1, i2, c, s) in rots:
            if want:
                ctx.apply_givens_rows(Udev.ptr, Udev.ld, n0, i1, i2, c, s)
    else:
        res = ctx.symdpr1(Ds, zs, rho, tau_k, vectors=False, U_dev=Udev.ptr if want else None, ldu=n0)
        lam[:] = res.lam
        info.scatter(np.arange(n0), res)
    isi = np.argsort(is_, kind="stable")
    lam = signr * lam
    es = jl_sortperm_desc(lam)
    E = Eigen(lam[es], _finish(ctx, Udev, isi.astype(np.int64), es, vectors, phase))
    out_info = info.permuted(es)
    out_info.stats = res.stats if res is not None else None
    return E, out_info


def svd_halfarrow(A: HalfArrow, tau=None, *, ctx: Context | None = None, vectors="host", forward_tau=False):
    """svd(A::HalfArrow, tau) -> (SVD(U,S,Vt), Info)  (arrowhead6.jl:526-661).

    Ordering by |D|, the sign row-scaling of V (:641) and the final permutation are the reference's.  Deflation:
      * zeros in z (:569-584): sigma = |D_j|, u = e_j, v = sign(D_j) e_j -- the reference records these and then SKIPS
        the solve of what remains (the `else` at :585 swallows it); here the remaining irreducible problem is solved;
      * exact duplicates of |D| (:586-632): the reference's branch cannot run (its gap vector misses the last pair of
        the no-tip shape, `zxxv=[zxx,n0+1]` nests a vector, the reduced HalfArrow loses the tip, the loop runs to
        nn+1).  What it is after is done here the way eigen(::SymArrow) does it: one Givens rotation per duplicate
        pair annihilates a border entry, the reduced problem is solved with row/column-mapped stores into the
        identity-initialised U and V on the device, and the rotations are applied to the rows of U and V
        (G' -- the algebraically consistent direction; V carries the signs of D already, so its rotation uses
        s*sign(D_i1)*sign(D_i2)).
    Not handled (the reference returns NaN there -- u = sigma*v./D divides by D, arrowhead6.jl:421): a zero in D next
    to a nonzero z entry, or a zero tip entry; these raise ValueError."""
    ctx = ctx or get_context()
    D, z = A.D, A.z
    nd, n0 = D.size, z.size
    tip = n0 == nd + 1
    tau_k = np.asarray(tau, np.float64) if (forward_tau and tau is not None) else None
    signD = np.sign(D)
    Da = np.abs(D)
    is_ = jl_sortperm_desc(Da)
    Ds, sg = Da[is_].copy(), signD[is_].copy()
    zs = z[:nd][is_].copy()
    want = vectors is not None
    nu_rows, nv_rows = n0, (n0 if tip else n0 + 1)
    if (tip and z[nd] == 0) or np.any((Ds == 0) & (zs != 0)):
        raise ValueError("svd(HalfArrow): a zero tip entry or a zero in D beside a nonzero z entry makes the reference "
                         "divide by zero (arrowhead6.jl:421-424); deflate it before calling")
    S = np.zeros(n0)
    info = Info.zeros(n0)
    z0 = np.flatnonzero(zs == 0)
    zx = np.flatnonzero(zs != 0)
    res = None
    if z0.size == 0 and not np.any(Ds[:-1] == Ds[1:]):
        # ---- no deflation: the loop of arrowhead6.jl:635-642 in one call
        zfull = np.concatenate([zs, [z[nd]]]) if tip else zs
        Udev = DeviceMatrix(ctx, nu_rows, n0) if want else None
        Vdev = DeviceMatrix(ctx, nv_rows, n0) if want else None
        res = ctx.halfarrow(Ds, zfull, tau_k, vectors=False, U_dev=Udev.ptr if want else None, ldu=nu_rows,
                            V_dev=Vdev.ptr if want else None, ldv=nv_rows, rowscale_v=sg if want else None)
        S[:] = res.lam
        info.scatter(np.arange(n0), res)
    else:
        Udev = DeviceMatrix(ctx, nu_rows, n0) if want else None
        Vdev = DeviceMatrix(ctx, nv_rows, n0) if want else None
        if want:
            ctx.set_identity(Udev.ptr, nu_rows, n0, nu_rows)
            ctx.set_identity(Vdev.ptr, nv_rows, n0, nv_rows)
        neg = []                                          # deflated columns whose v = -e_j
        S[z0] = Ds[z0]
        neg += [int(j) for j in z0 if sg[j] < 0]
        Dr, zr = Ds[zx].copy(), zs[zx].copy()
        if zx.size == 0:
            if tip:                                       # only the tip row is left (arrowhead6.jl:577-580)
                S[n0 - 1] = abs(z[nd])
                if z[nd] < 0:
                    neg.append(n0 - 1)
            rots = []
        else:
            keep, rots = _deflate_duplicates(Dr, zr, zx, S, "transpose")
            dup = np.setdiff1d(np.arange(zx.size), keep)
            neg += [int(zx[p]) for p in dup if sg[zx[p]] < 0]
            rows = zx[keep]
            zk = np.concatenate([zr[keep], [z[nd]]]) if tip else zr[keep]
            rowmap_u = np.concatenate([rows, [n0 - 1]]) if tip else rows
            rowmap_v = np.concatenate([rows, [nv_rows - 1]])
            res = ctx.halfarrow(Dr[keep], zk, tau_k, vectors=False, U_dev=Udev.ptr if want else None, ldu=nu_rows,
                                V_dev=Vdev.ptr if want else None, ldv=nv_rows, rowmap_u=rowmap_u if want else None,
                                rowmap_v=rowmap_v if want else None, rowscale_v=sg[rows] if want else None)
            S[rowmap_u] = res.lam
            info.scatter(rowmap_u, res)
        if want:
            minus = np.array([-1.0])
            for j in neg:                                 # V[j,j] = sign(D_j) (arrowhead6.jl:572-574)
                ctx.memcpy_h2d(Vdev.ptr + 8 * (j * nv_rows + j), minus, 8)
            for (i1, i2, c, s) in rots:
                ctx.apply_givens_rows(Udev.ptr, Udev.ld, n0, i1, i2, c, s)
                ctx.apply_givens_rows(Vdev.ptr, Vdev.ld, n0, i1, i2, c, s * sg[i1] * sg[i2])
    isi = np.argsort(is_, kind="stable").astype(np.int64)
    es = jl_sortperm_desc(S)
    rows_u = isi if not tip else np.concatenate([isi, [n0 - 1]])
    rows_v = np.concatenate([isi, [n0]]) if not tip else rows_u
    Uo = _finish(ctx, Udev, rows_u, es, vectors)
    Vo = _finish(ctx, Vdev, rows_v, es, vectors)
    Vt = None if Vo is None else (Vo.T if isinstance(Vo, np.ndarray) else Vo)   # device: V itself (Vt is its transpose)
    out_info = info.permuted(es)
    out_info.stats = res.stats if res is not None else None
    return SVD(Uo, S[es], Vt), out_info


# ---------------------------------------------------------------------------------------------------
# Hermitian wrappers (hermitian.jl:107-139): phase vector c = z./abs.(z), real solve, `Diagonal(c)*U`.
def phase_split(zc: np.ndarray):
    """c = z ./ abs.(z) and the real border real(conj(c).*z), component-wise as Julia evaluates them
    (complex: abs = hypot; quaternion: abs = sqrt(s^2+v1^2+v2^2+v3^2), products summed left to right)."""
    if zc.shape[1] == 2:
        mod = np.hypot(zc[:, 0], zc[:, 1])
    else:
        mod = np.sqrt(((zc[:, 0] * zc[:, 0] + zc[:, 1] * zc[:, 1]) + zc[:, 2] * zc[:, 2]) + zc[:, 3] * zc[:, 3])
    c = zc / mod[:, None]
    zr = c[:, 0] * zc[:, 0] + c[:, 1] * zc[:, 1]
    for q in range(2, zc.shape[1]):
        zr = zr + c[:, q] * zc[:, q]
    return c, zr


def _hyper_to_host(M: np.ndarray, nrows: int, ncomp: int) -> np.ndarray:
    """(nrows*ncomp, ncols) column-major real -> complex (nrows, ncols) or quaternion (nrows, ncols, 4)."""
    ncols = M.shape[1]
    T = np.ascontiguousarray(M.T).reshape(ncols, nrows, ncomp)
    if ncomp == 2:
        return np.ascontiguousarray((T[:, :, 0] + 1j * T[:, :, 1]).T)
    return np.ascontiguousarray(np.transpose(T, (1, 0, 2)))


def _apply_phase(ctx, Ereal, phase, vectors):
    """Diagonal(c) * U.vectors as a device kernel (ah_scale_rows_phase): complex/quaternion entries interleaved."""
    if Ereal is None:
        return None
    ncomp = phase.shape[1]
    if isinstance(Ereal, np.ndarray):          # trivial cases the driver answers on the host (1x1, diagonal)
        H = phase[:, None, :] * Ereal[:, :, None]
        return H[:, :, 0] + 1j * H[:, :, 1] if ncomp == 2 else H
    n, m = Ereal.nrows, Ereal.ncols
    out = DeviceMatrix(ctx, n * ncomp, m)
    ctx.scale_rows_phase(out.ptr, n, Ereal.ptr, Ereal.ld, phase, n, m)
    Ereal.free()
    if vectors == "device":
        return out
    host = _hyper_to_host(out.to_host(), n, ncomp)
    out.free()
    return host


def eigen_hermarrow(A: HermArrow, tau=None, *, ctx: Context | None = None, vectors="host", **kw):
    """eigen(A::HermArrow) (hermitian.jl:107-115) -> (Eigen(values, complex|quaternion vectors), Info)."""
    ctx = ctx or get_context()
    c, zr = phase_split(A.zc)
    one = np.zeros((1, A.ncomp)); one[0, 0] = 1.0
    phase = np.concatenate([c[: A.i - 1], one, c[A.i - 1:]])          # insert!(c, A.i, one(T))
    # the phase multiplication rides on the driver's final permutation pass (ah_gather_permuted_phase): no extra pass over U
    E, info = eigen_symarrow(SymArrow(A.D, zr, A.a, A.i), tau, ctx=ctx, vectors=vectors, phase=phase, **kw)
    V = E.vectors
    if isinstance(V, np.ndarray) and not np.iscomplexobj(V) and V.ndim == 2:      # trivial cases answered on the host (1x1, diagonal)
        V = _apply_phase(ctx, V, phase, vectors)
    return Eigen(E.values, V), info


def eigen_hermdpr1(A: HermDPR1, tau=None, *, ctx: Context | None = None, vectors="host", **kw):
    """eigen(A::HermDPR1) (hermitian.jl:131-139)."""
    ctx = ctx or get_context()
    c, ur = phase_split(A.uc)
    E, info = eigen_symdpr1(SymDPR1(A.D, ur, A.r), tau, ctx=ctx, vectors=vectors, phase=c, **kw)
    V = E.vectors
    if isinstance(V, np.ndarray) and not np.iscomplexobj(V) and V.ndim == 2:
        V = _apply_phase(ctx, V, c, vectors)
    return Eigen(E.values, V), info


def eigen(A, tau=None, **kw):
    """LinearAlgebra.eigen dispatch of the reference (arrowhead1.jl:2)."""
    if isinstance(A, SymArrow):
        return eigen_symarrow(A, tau, **kw)
    if isinstance(A, SymDPR1):
        return eigen_symdpr1(A, tau, **kw)
    if isinstance(A, HermArrow):
        return eigen_hermarrow(A, tau, **kw)
    if isinstance(A, HermDPR1):
        return eigen_hermdpr1(A, tau, **kw)
    raise TypeError(f"eigen: unsupported type {type(A).__name__}")


def svd(A, tau=None, **kw):
    if isinstance(A, HalfArrow):
        return svd_halfarrow(A, tau, **kw)
    raise TypeError(f"svd: unsupported type {type(A).__name__}")


def tdc_recursive(dv, ev, *, ctx: Context | None = None, rotation="reference"):
    """tdc(T::SymTridiagonal) (arrowhead7.jl:10-36) exactly as written: recursion, one arrow merge per call on the
    GPU, the back-transform products on the host.  Kept for small matrices and as the cross-check of tdc_device."""
    dv = np.asarray(dv, np.float64)
    ev = np.asarray(ev, np.float64)
    n = dv.size
    if n == 1:
        return Eigen(np.array([dv[0]]), np.ones((1, 1)))
    if n == 2:
        E, _ = eigen_symarrow(SymArrow([dv[0]], [ev[0]], dv[1], 2), ctx=ctx, rotation=rotation)
        return E
    k = n // 2
    E1 = tdc_recursive(dv[:k], ev[: k - 1], ctx=ctx, rotation=rotation)
    E2 = tdc_recursive(dv[k + 1:], ev[k + 1:], ctx=ctx, rotation=rotation)
    Dm = np.concatenate([E1.values, E2.values])
    zm = np.concatenate([E1.vectors[-1, :] * ev[k - 1], E2.vectors[0, :] * ev[k]])
    E, _ = eigen_symarrow(SymArrow(Dm, zm, dv[k], k + 1), ctx=ctx, rotation=rotation)
    W = np.array(E.vectors, order="F", copy=True)
    W[:k, :] = E1.vectors @ W[:k, :]
    W[k + 1:, :] = E2.vectors @ W[k + 1:, :]
    return Eigen(E.values, W)


_TDC_BIG = 1000      # merges with at least this many poles go through the fast pipeline one by one


def tdc_device(dv, ev, *, ctx: Context | None = None, vectors="host", rotation="reference"):
    """tdc(T::SymTridiagonal) (arrowhead7.jl:10-36) with everything of size n^2 resident on the GPU.

    The recursion tree (k = div(n,2); T1 = T[1:k], T2 = T[k+2:n], arrowhead7.jl:21-23) is walked bottom-up, one
    height at a time: all merges of one height are independent, so their arrow problems are solved in ONE batched
    launch (ah_symarrow_eigen_batched; large ones through the fast pipeline), their merge vectors z (:28) are read
    from the resident eigenvector blocks, the final `view(U,rows,es)` of the arrow driver is one batched gather, and
    the back-transforms `E1.vectors*E.vectors[1:k,:]`, `E2.vectors*E.vectors[k+2:n,:]` (:32-33) are two launches
    per size class of the library's own FP64 tensor-core GEMM (ah_dgemm_batched: DMMA.8x8x4, csrc/ah_gemm.cuh).
    The gathers, copies and GEMMs of a height are enqueued without host synchronisation (ah_set_async); the host
    only waits where it needs data back (eigenvalues of the merges, the merge vectors z).  Host work is the
    reference's host logic only: stable sorts, the (exact) deflation tests, index bookkeeping.  Node [lo,hi) keeps its eigenvectors in the diagonal block
    [lo:hi, lo:hi] of one of three n x n buffers (height mod 3), so children are never overwritten while read.
    Merges that need deflation take the general per-problem driver (eigen_symarrow)."""
    ctx = ctx or get_context()
    dv = np.ascontiguousarray(dv, np.float64)
    ev = np.ascontiguousarray(ev, np.float64)
    n = dv.size
    if n <= 2:
        return tdc_recursive(dv, ev, ctx=ctx, rotation=rotation)
    height = lambda m: int(m).bit_length() - 1                        # floor(log2(m)); 0 for a leaf
    by_height: dict[int, list] = {}
    stack = [(0, n)]
    while stack:
        lo, hi = stack.pop()
        m = hi - lo
        if m < 2:
            continue
        by_height.setdefault(height(m), []).append((lo, hi))
        k = m // 2
        stack.append((lo, lo + k))
        stack.append((lo + k + 1, hi))
    # one resident workspace per context, reused across calls (cudaMalloc/cudaFree of GBs cost more than the solve):
    # three n x n eigenvector buffers side by side, then the solver-output and permuted-output scratch (n^2 each)
    ws = getattr(ctx, "_tdc_workspace", None)
    if ws is None or ws.nrows != n:
        if ws is not None:
            ws.free()
        ws = ctx._tdc_workspace = DeviceMatrix(ctx, n, 5 * n)
    Q = ws
    qbase = lambda t: Q.ptr + 8 * t * n * n
    s_base, wp_base = qbase(3), qbase(4)

    class _View:                                                      # a slice of the workspace with DeviceMatrix's face
        def __init__(self, ptr):
            self.ptr = ptr

        def free(self):
            pass
    ctx.set_async(True)
    try:
        return _tdc_levels(ctx, dv, ev, n, by_height, height, Q, qbase, s_base, wp_base, _View, vectors, rotation)
    finally:
        ctx.set_async(False)
        ctx.wait()


def _tdc_levels(ctx, dv, ev, n, by_height, height, Q, qbase, s_base, wp_base, _View, vectors, rotation):
    ctx.set_identity(Q.ptr, n, n, n)                                  # leaves: U = [1] in buffer 0
    # eigenvalues of every finished node [lo,hi) live in lam[lo:hi] (node-local order): children are read and the merged
    # node is written back with index arithmetic over the whole size group -- no per-node Python work (8191 nodes at n = 8192)
    lam = dv.copy()
    stats = {"levels": 0, "batched_problems": 0, "pipeline_problems": 0, "driver_problems": 0, "gemm_calls": 0, "gemm_flops": 0.0,
             "gemm_shapes": []}
    for h in sorted(by_height):
        stats["levels"] += 1
        groups: dict[int, list] = {}
        for node in by_height[h]:
            groups.setdefault(node[1] - node[0], []).append(node)
        for m, nodes in sorted(groups.items()):
            nodes.sort()
            nb, k = len(nodes), m // 2
            r, N = m - k - 1, m - 1
            los = np.array([lo for lo, _ in nodes], np.int64)
            p = los + k                                               # pivot rows
            tl, tr, tq = height(k) % 3, (height(r) % 3 if r else 0), h % 3
            # ---- merge problems: D = [L1; L2], z = [U1[end,:]*ev[p-1]; U2[1,:]*ev[p]], a = dv[p], arrow top at k+1
            D = np.empty((nb, N)); z = np.empty((nb, N))
            cl = los[:, None] + np.arange(k)[None, :]
            idx = [((tl * n + cl) * n + (p - 1)[:, None]).ravel()]
            if r:
                cr = (p + 1)[:, None] + np.arange(r)[None, :]
                idx.append(((tr * n + cr) * n + (p + 1)[:, None]).ravel())
            g = ctx.gather_elements(Q.ptr, np.concatenate(idx))
            z[:, :k] = g[: nb * k].reshape(nb, k) * ev[p - 1][:, None]
            if r:
                z[:, k:] = g[nb * k:].reshape(nb, r) * ev[p][:, None]
            D[:, :k] = lam[cl]
            if r:
                D[:, k:] = lam[cr]
            a = dv[p]
            # ---- the arrow driver's host logic (arrowhead3.jl:571-573, 583-623), vectorised over the group
            is_ = np.argsort(-D, axis=1, kind="stable")
            Ds = np.take_along_axis(D, is_, axis=1); zs = np.take_along_axis(z, is_, axis=1)
            needs_driver = (zs == 0).any(axis=1) | (Ds == 0).any(axis=1)
            if N > 1:
                needs_driver |= (Ds[:, :-1] == Ds[:, 1:]).any(axis=1)
            fast = np.flatnonzero(~needs_driver)
            S = _View(s_base)                                         # solver output, one m x m block per node (nb*m*m <= n*n)
            Wp = _View(wp_base)                                       # after the driver's row/column permutation
            L = np.empty((nb, m))
            if fast.size:
                if N >= _TDC_BIG:
                    for b in fast:
                        res = ctx.symarrow(Ds[b], zs[b], a[b], vectors=False, U_dev=S.ptr + 8 * b * m * m, ldu=m)
                        L[b] = res.lam
                    stats["pipeline_problems"] += int(fast.size)
                else:
                    # contiguous sub-batch of the solvable nodes
                    Sf = S if fast.size == nb else DeviceMatrix(ctx, m * m, fast.size)
                    res = ctx.symarrow_batched(Ds[fast], zs[fast], a[fast], U_dev=Sf.ptr, ldu=m, ustride=m * m)
                    L[fast] = res.lam.reshape(fast.size, m)
                    if Sf is not S:
                        ctx.copy2d_batched(m * m, 1, S.ptr + 8 * m * m * fast.astype(np.uint64), m * m,
                                           Sf.ptr + 8 * m * m * np.arange(fast.size, dtype=np.uint64), m * m)
                        Sf.free()
                    stats["batched_problems"] += int(fast.size)
                es = np.argsort(-L[fast], axis=1, kind="stable")
                isi = np.argsort(is_[fast], axis=1, kind="stable")
                rows = np.concatenate([isi[:, :k], np.full((fast.size, 1), N, np.int64), isi[:, k:]], axis=1)
                if fast.size == nb:
                    ctx.gather_permuted_batched(Wp.ptr, m * m, m, S.ptr, m * m, m, rows, es)
                else:
                    for j, b in enumerate(fast):
                        ctx.gather_permuted_batched(Wp.ptr + 8 * b * m * m, m * m, m, S.ptr + 8 * b * m * m, m * m, m,
                                                    rows[j:j + 1], es[j:j + 1])
                L[fast] = np.take_along_axis(L[fast], es, axis=1)
            for b in np.flatnonzero(needs_driver):                    # deflation: the general driver, one problem at a time
                E, _ = eigen_symarrow(SymArrow(D[b], z[b], a[b], k + 1), ctx=ctx, vectors="device", rotation=rotation)
                L[b] = E.values
                ctx.copy2d_batched(m, m, [Wp.ptr + 8 * b * m * m], m, [E.vectors.ptr], E.vectors.ld)
                E.vectors.free()
                stats["driver_problems"] += 1
            S.free()
            # ---- back-transform into buffer h mod 3 (arrowhead7.jl:32-33) + the pivot row
            wp = Wp.ptr + 8 * m * m * np.arange(nb, dtype=np.uint64)
            qout = np.uint64(qbase(tq)) + 8 * (los * n + los).astype(np.uint64)
            ql = np.uint64(qbase(tl)) + 8 * (los * n + los).astype(np.uint64)
            ctx.dgemm_batched(k, m, k, ql, n, wp, m, qout, n)
            ctx.copy2d_batched(1, m, qout + np.uint64(8 * k), n, wp + np.uint64(8 * k), m)
            stats["gemm_calls"] += 1
            stats["gemm_flops"] += 2.0 * nb * k * m * k
            stats["gemm_shapes"].append((nb, k, m, k))
            if r:
                qr = np.uint64(qbase(tr)) + 8 * ((p + 1) * n + (p + 1)).astype(np.uint64)
                ctx.dgemm_batched(r, m, r, qr, n, wp + np.uint64(8 * (k + 1)), m, qout + np.uint64(8 * (k + 1)), n)
                stats["gemm_calls"] += 1
                stats["gemm_flops"] += 2.0 * nb * r * m * r
                stats["gemm_shapes"].append((nb, r, m, r))
            Wp.free()
            lam[los[:, None] + np.arange(m)[None, :]] = L
    values = lam
    troot = height(n) % 3
    E = Eigen(values, None)
    E.stats = stats
    if vectors is None:
        return E
    out = DeviceMatrix(ctx, n, n)
    ctx.copy2d_batched(n, n, [out.ptr], n, [qbase(troot)], n)
    if vectors == "device":
        E.vectors = out
    else:
        E.vectors = out.to_host()
        out.free()
    return E


def tdc(dv, ev, *, ctx: Context | None = None, rotation="reference", method="auto", vectors="host"):
    """tdc(T::SymTridiagonal) (arrowhead7.jl:10-36).  method: "recursive" (host back-transform, the literal
    recursion), "device" (level-batched, resident), "auto" (device from n = 64 on)."""
    n = np.asarray(dv).size
    if method == "recursive" or (method == "auto" and n < 64):
        return tdc_recursive(dv, ev, ctx=ctx, rotation=rotation)
    return tdc_device(dv, ev, ctx=ctx, vectors=vectors, rotation=rotation)
