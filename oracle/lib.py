"""ctypes binding of oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "arrowhead_oracle.c")
_OUT = os.path.join(_HERE, "_build", "liboracle.so")
_OUT_HI = os.path.join(_HERE, "_build", "liboracle_hi.so")

# -ffp-contract=off: Julia never contracts a*b+c into an FMA; the restatement must not either.
_CFLAGS = ["-O2", "-std=c11", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math"]


def build(force: bool = False, hi: bool = False) -> str:
    """Compile the C restatement with the PATH gcc (the image's $CC has no libgomp.spec).
    hi=True builds the -DAH_ORACLE_HI variant (long-double accumulators, same algorithm)."""
    out = _OUT_HI if hi else _OUT
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(_SRC):
        return out
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("oracle: gcc not found on PATH")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = out + f".tmp{os.getpid()}"
    extra = ["-DAH_ORACLE_HI"] if hi else []
    subprocess.run([gcc, *_CFLAGS, *extra, "-shared", "-o", tmp, _SRC, "-lm"], check=True)
    os.replace(tmp, out)
    return out


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


@dataclass
class RangeResult:
    """Outputs of a per-k range solve; column c belongs to eigenpair k0+c (1-based k)."""

    k0: int
    k1: int
    lam: np.ndarray
    U: np.ndarray | None      # column-major (Fortran) n x cnt
    V: np.ndarray | None      # HalfArrow only
    sind: np.ndarray
    kb: np.ndarray
    kz: np.ndarray
    knu: np.ndarray
    krho: np.ndarray
    qout: np.ndarray
    steps: np.ndarray
    status: np.ndarray


class OracleLib:
    def __init__(self, path: str | None = None):
        self.path = path or build()
        self.lib = C.CDLL(self.path)
        common_tail = [_ip, _dp, _dp, _dp, _dp, _ip, _i32p, _i32p]
        self.lib.ah_oracle_symarrow_eigen.restype = C.c_int
        self.lib.ah_oracle_symarrow_eigen.argtypes = [
            C.c_int64, _dp, _dp, C.c_double, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int64,
            *common_tail]
        self.lib.ah_oracle_symdpr1_eigen.restype = C.c_int
        self.lib.ah_oracle_symdpr1_eigen.argtypes = [
            C.c_int64, _dp, _dp, C.c_double, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int64,
            *common_tail]
        self.lib.ah_oracle_halfarrow_svd.restype = C.c_int
        self.lib.ah_oracle_halfarrow_svd.argtypes = [
            C.c_int64, C.c_int64, _dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int64,
            _dp, C.c_int64, *common_tail]
        self.lib.ah_oracle_rootsah_eigvals.restype = C.c_int
        self.lib.ah_oracle_rootsah_eigvals.argtypes = [
            C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, _dp, C.c_int64, C.c_int64, C.c_int, _dp, *common_tail]
        self.lib.ah_oracle_dd_op.restype = None
        self.lib.ah_oracle_dd_op.argtypes = [C.c_int, _dp, _dp, _dp]
        self.lib.ah_oracle_max_threads.restype = C.c_int

    def max_threads(self) -> int:
        return int(self.lib.ah_oracle_max_threads())

    @staticmethod
    def _info(cnt):
        return (np.zeros(cnt, np.int64), np.zeros(cnt), np.zeros(cnt), np.zeros(cnt), np.zeros(cnt),
                np.zeros(cnt, np.int64), np.zeros(cnt, np.int32), np.zeros(cnt, np.int32))

    @staticmethod
    def _info_ptrs(info):
        sind, kb, kz, knu, krho, qout, steps, status = info
        return (_p(sind, _ip), _p(kb, _dp), _p(kz, _dp), _p(knu, _dp), _p(krho, _dp), _p(qout, _ip),
                _p(steps, _i32p), _p(status, _i32p))

    def symarrow(self, D, z, a, tau, k0=1, k1=None, vectors=True, nthreads=0) -> RangeResult:
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64)
        tau = np.ascontiguousarray(tau, np.float64)
        N = D.size; n = N + 1
        k1 = n if k1 is None else k1
        cnt = k1 - k0 + 1
        lam = np.zeros(cnt)
        U = np.zeros((n, cnt), order="F") if vectors else None
        info = self._info(cnt)
        rc = self.lib.ah_oracle_symarrow_eigen(N, _p(D, _dp), _p(z, _dp), float(a), _p(tau, _dp), k0, k1,
                                               nthreads, _p(lam, _dp), _p(U, _dp), n, *self._info_ptrs(info))
        if rc != 0:
            raise RuntimeError(f"ah_oracle_symarrow_eigen rc={rc}")
        return RangeResult(k0, k1, lam, U, None, *info)

    def rootsah_eigvals(self, D, z, zlo, a, alo, tau, k0=1, k1=None, nthreads=0) -> RangeResult:
        """eigen(A,zD,alphaD,k,tau) for k0..k1 (arrowhead7.jl:351-495): eigenvalues + Qout of the companion arrow"""
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64); zlo = np.ascontiguousarray(zlo, np.float64)
        tau = np.ascontiguousarray(tau, np.float64)
        N = D.size
        k1 = N + 1 if k1 is None else k1
        cnt = k1 - k0 + 1
        lam = np.zeros(cnt)
        info = self._info(cnt)
        rc = self.lib.ah_oracle_rootsah_eigvals(N, _p(D, _dp), _p(z, _dp), _p(zlo, _dp), float(a), float(alo), _p(tau, _dp),
                                                k0, k1, nthreads, _p(lam, _dp), *self._info_ptrs(info))
        if rc != 0:
            raise RuntimeError(f"ah_oracle_rootsah_eigvals rc={rc}")
        return RangeResult(k0, k1, lam, None, None, *info)

    def symdpr1(self, D, u, rho, tau, k0=1, k1=None, vectors=True, nthreads=0) -> RangeResult:
        D = np.ascontiguousarray(D, np.float64); u = np.ascontiguousarray(u, np.float64)
        tau = np.ascontiguousarray(tau, np.float64)
        n = D.size
        k1 = n if k1 is None else k1
        cnt = k1 - k0 + 1
        lam = np.zeros(cnt)
        U = np.zeros((n, cnt), order="F") if vectors else None
        info = self._info(cnt)
        rc = self.lib.ah_oracle_symdpr1_eigen(n, _p(D, _dp), _p(u, _dp), float(rho), _p(tau, _dp), k0, k1,
                                              nthreads, _p(lam, _dp), _p(U, _dp), n, *self._info_ptrs(info))
        if rc != 0:
            raise RuntimeError(f"ah_oracle_symdpr1_eigen rc={rc}")
        return RangeResult(k0, k1, lam, U, None, *info)

    def halfarrow(self, D, z, tau, k0=1, k1=None, vectors=True, nthreads=0) -> RangeResult:
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64)
        tau = np.ascontiguousarray(tau, np.float64)
        nd, nz = D.size, z.size
        k1 = nz if k1 is None else k1
        cnt = k1 - k0 + 1
        sig = np.zeros(cnt)
        U = np.zeros((nz, cnt), order="F") if vectors else None
        V = np.zeros((nd + 1, cnt), order="F") if vectors else None
        info = self._info(cnt)
        rc = self.lib.ah_oracle_halfarrow_svd(nd, nz, _p(D, _dp), _p(z, _dp), _p(tau, _dp), k0, k1, nthreads,
                                              _p(sig, _dp), _p(U, _dp), nz, _p(V, _dp), nd + 1,
                                              *self._info_ptrs(info))
        if rc != 0:
            raise RuntimeError(f"ah_oracle_halfarrow_svd rc={rc}")
        return RangeResult(k0, k1, sig, U, V, *info)

    def dd_op(self, op: str, x, y):
        code = {"add": 0, "sub": 1, "mul": 2, "div": 3}[op]
        xa = np.array(x, np.float64); ya = np.array(y, np.float64); out = np.zeros(2)
        self.lib.ah_oracle_dd_op(code, _p(xa, _dp), _p(ya, _dp), _p(out, _dp))
        return float(out[0]), float(out[1])


_LIB: OracleLib | None = None
_LIB_HI: OracleLib | None = None


def get_lib() -> OracleLib:
    global _LIB
    if _LIB is None:
        _LIB = OracleLib()
    return _LIB


def get_lib_hi() -> OracleLib:
    """Same algorithm with long-double accumulators: the accuracy yardstick, not the reference's rounding."""
    global _LIB_HI
    if _LIB_HI is None:
        _LIB_HI = OracleLib(build(hi=True))
    return _LIB_HI
