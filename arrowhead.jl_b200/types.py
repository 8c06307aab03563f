"""Matrix types and generators of the reference API (host side, numpy).

SymArrow / SymDPR1   src/arrowhead1.jl:20-47, 65-83
HalfArrow            src/arrowhead5.jl:25-46
GenSymArrow/GenSymDPR1  src/arrowhead3.jl:23-44,  GenHalfArrow  src/arrowhead6.jl:37-43
Info                 src/arrowhead3.jl:395-402 (kept as structure-of-arrays, see include/arrowhead_b200.h)

Indices follow the reference: the arrow-top position `i` is 1-based.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class SymArrow:
    """[diag(D) z; z' a] with the arrow top at position [i,i] (1-based)."""

    def __init__(self, D, z, a, i):
        self.D = np.array(D, dtype=np.float64).ravel()
        self.z = np.array(z, dtype=np.float64).ravel()
        if self.D.size != self.z.size:
            raise ValueError("SymArrow: D and z must have the same length")
        self.a = float(a)
        self.i = int(i)
        if not 1 <= self.i <= self.D.size + 1:
            raise ValueError("SymArrow: arrow top position out of range")

    @property
    def shape(self):
        n = self.D.size + 1
        return (n, n)

    def __getitem__(self, ij):
        i, j = ij
        i, j, t = i + 1, j + 1, self.i          # reference getindex is 1-based (arrowhead1.jl:33-44)
        if i == j:
            return self.D[i - 1] if i < t else (self.a if i == t else self.D[i - 2])
        if i == t:
            return self.z[j - 1] if j < t else self.z[j - 2]
        if j == t:
            return self.z[i - 1] if i < t else self.z[i - 2]
        return 0.0

    def dense(self) -> np.ndarray:
        n = self.D.size + 1
        t = self.i - 1
        idx = np.array([r for r in range(n) if r != t], dtype=np.int64)
        A = np.zeros((n, n))
        A[idx, idx] = self.D
        A[idx, t] = self.z
        A[t, idx] = self.z
        A[t, t] = self.a
        return A


class SymDPR1:
    """diag(D) + r*u*u'."""

    def __init__(self, D, u, r):
        self.D = np.array(D, dtype=np.float64).ravel()
        self.u = np.array(u, dtype=np.float64).ravel()
        if self.D.size != self.u.size:
            raise ValueError("SymDPR1: D and u must have the same length")
        self.r = float(r)

    @property
    def shape(self):
        return (self.D.size, self.D.size)

    def __getitem__(self, ij):
        i, j = ij
        v = self.r * self.u[i] * self.u[j]
        return self.D[i] + v if i == j else v

    def dense(self) -> np.ndarray:
        return np.diag(self.D) + self.r * np.outer(self.u, self.u)


class HalfArrow:
    """[diag(D) z]; len(z)==len(D): nd x (nd+1); len(z)==len(D)+1: square with a tip."""

    def __init__(self, D, z):
        self.D = np.array(D, dtype=np.float64).ravel()
        self.z = np.array(z, dtype=np.float64).ravel()
        if self.z.size not in (self.D.size, self.D.size + 1):
            raise ValueError("HalfArrow: length(z) must be length(D) or length(D)+1")

    @property
    def shape(self):
        return (max(self.D.size, self.z.size), self.D.size + 1)

    def dense(self) -> np.ndarray:
        nd = self.D.size
        A = np.zeros(self.shape)
        A[np.arange(nd), np.arange(nd)] = self.D
        A[: self.z.size, nd] = self.z
        return A


def _as_hyper(z, ncomp=None):
    """complex vector -> (n,2) float array; quaternion vectors are given as (n,4) [s, v1, v2, v3] arrays."""
    z = np.asarray(z)
    if np.iscomplexobj(z):
        z = z.ravel()
        return np.stack([z.real, z.imag], axis=1).astype(np.float64)
    z = np.asarray(z, np.float64)
    if z.ndim == 1:
        return np.stack([z, np.zeros_like(z)], axis=1)
    if z.ndim != 2 or z.shape[1] not in (2, 4):
        raise ValueError("complex (n,) / (n,2) or quaternion (n,4) entries expected")
    return z.copy()


class HermArrow:
    """Hermitian arrow [diag(D) z; z' a] with complex or quaternion border (hermitian.jl:9-14).
    `z` is a complex vector or an (n,4) array of quaternion components; stored as (n, ncomp) reals."""

    def __init__(self, D, z, a, i):
        self.D = np.array(D, dtype=np.float64).ravel()
        self.zc = _as_hyper(z)
        if self.D.size != self.zc.shape[0]:
            raise ValueError("HermArrow: D and z must have the same length")
        self.a = float(a)
        self.i = int(i)

    @property
    def ncomp(self):
        return self.zc.shape[1]

    @property
    def z(self):
        return self.zc[:, 0] + 1j * self.zc[:, 1] if self.ncomp == 2 else self.zc

    def dense(self) -> np.ndarray:
        """complex dense matrix (complex border only)"""
        if self.ncomp != 2:
            raise NotImplementedError("dense() of a quaternion matrix")
        n = self.D.size + 1
        t = self.i - 1
        idx = np.array([r for r in range(n) if r != t], dtype=np.int64)
        A = np.zeros((n, n), dtype=np.complex128)
        A[idx, idx] = self.D
        A[idx, t] = self.z
        A[t, idx] = np.conj(self.z)
        A[t, t] = self.a
        return A


class HermDPR1:
    """diag(D) + r*u*u' with complex or quaternion u (hermitian.jl:47-51)."""

    def __init__(self, D, u, r):
        self.D = np.array(D, dtype=np.float64).ravel()
        self.uc = _as_hyper(u)
        if self.D.size != self.uc.shape[0]:
            raise ValueError("HermDPR1: D and u must have the same length")
        self.r = float(r)

    @property
    def ncomp(self):
        return self.uc.shape[1]

    @property
    def u(self):
        return self.uc[:, 0] + 1j * self.uc[:, 1] if self.ncomp == 2 else self.uc

    def dense(self) -> np.ndarray:
        if self.ncomp != 2:
            raise NotImplementedError("dense() of a quaternion matrix")
        return np.diag(self.D).astype(np.complex128) + self.r * np.outer(self.u, np.conj(self.u))


def GenHermArrow(n: int, i: int, ncomp: int = 2, rng=None) -> HermArrow:
    """GenHermArrow(n,i,T) (hermitian.jl:76-79): rand(T) has every component uniform in [0,1)."""
    rng = np.random.default_rng() if rng is None else rng
    return HermArrow(rng.random(n - 1), rng.random((n - 1, ncomp)), rng.random(), i)


def GenHermDPR1(n: int, ncomp: int = 2, rng=None) -> HermDPR1:
    rng = np.random.default_rng() if rng is None else rng
    return HermDPR1(rng.random(n), rng.random((n, ncomp)), rng.random())


def GenSymArrow(n: int, i: int, rng=None) -> SymArrow:
    rng = np.random.default_rng() if rng is None else rng
    return SymArrow(rng.random(n - 1), rng.random(n - 1), rng.random(), i)


def GenSymDPR1(n: int, rng=None) -> SymDPR1:
    rng = np.random.default_rng() if rng is None else rng
    return SymDPR1(rng.random(n), rng.random(n), rng.random())


def GenHalfArrow(n: int, p: int, rng=None) -> HalfArrow:
    rng = np.random.default_rng() if rng is None else rng
    if p == 0:
        return HalfArrow(rng.random(n - 1) - 0.5, rng.random(n - 1) - 0.5)
    return HalfArrow(rng.random(n - 1) - 0.5, rng.random(n) - 0.5)


_INFO_FIELDS = ("sind", "kb", "kz", "knu", "krho", "qout", "steps", "status")


@dataclass
class Info:
    """kappa: Info(Sind,Kb,Kz,Knu,Krho,Qout) per eigenpair, as arrays (+ steps/status, see the C header)."""

    sind: np.ndarray
    kb: np.ndarray
    kz: np.ndarray
    knu: np.ndarray
    krho: np.ndarray
    qout: np.ndarray
    steps: np.ndarray
    status: np.ndarray

    @staticmethod
    def zeros(n: int) -> "Info":
        return Info(np.zeros(n, np.int64), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n),
                    np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int32))

    def scatter(self, dst, res):
        for f in _INFO_FIELDS:
            getattr(self, f)[dst] = getattr(res, f)

    def permuted(self, p) -> "Info":
        return Info(*(getattr(self, f)[p].copy() for f in _INFO_FIELDS))

    def __len__(self):
        return self.sind.size

    def __getitem__(self, k):
        return tuple(getattr(self, f)[k] for f in _INFO_FIELDS[:6])


@dataclass
class Eigen:
    values: np.ndarray
    vectors: object          # numpy array (host) or DeviceMatrix, or None


@dataclass
class SVD:
    U: object
    S: np.ndarray
    Vt: object
