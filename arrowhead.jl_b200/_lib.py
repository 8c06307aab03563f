"""ctypes binding of libarrowhead_b200.so (the C ABI in include/arrowhead_b200.h).

This is the only compute backend of the package: if the shared library is missing or no sm_100 GPU is
usable, every solve raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ARROWHEAD_B200_LIB") or os.path.join(_PKG, "libarrowhead_b200.so")   # env: A/B builds only

AH_MEM_HOST, AH_MEM_DEVICE = 0, 1
ST_NU_INF, ST_NEEDS_BIGFLOAT, ST_POLE_RETURN, ST_REF_THROWS, ST_ITER_LIMIT, ST_INTERLACE = 1, 2, 4, 8, 16, 32

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


class AhResult(C.Structure):
    _fields_ = [
        ("lambda_", _dp), ("sind", _i64p), ("kb", _dp), ("kz", _dp), ("knu", _dp), ("krho", _dp),
        ("qout", _i64p), ("steps", _i32p), ("status", _i32p),
        ("U", C.c_void_p), ("ldu", C.c_int64), ("U_mem", C.c_int32),
        ("V", C.c_void_p), ("ldv", C.c_int64), ("V_mem", C.c_int32),
        ("rowmap_u", _i64p), ("rowmap_v", _i64p), ("rowscale_v", _dp),
    ]


class AhStats(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double), ("main_kernel_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("launches", C.c_int64), ("n_fast", C.c_int64), ("n_full", C.c_int64), ("total_steps", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("prep_ms", C.c_double), ("bisect_ms", C.c_double), ("vec_ms", C.c_double), ("full_ms", C.c_double),
        ("n_rem3", C.c_int64), ("rem3_ms", C.c_double), ("host_pre_ms", C.c_double), ("host_post_ms", C.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "ah_version", "ah_create", "ah_destroy", "ah_last_error", "ah_set_stream", "ah_get_stats", "ah_set_mode",
    "ah_symarrow_eigen", "ah_symdpr1_eigen", "ah_halfarrow_svd", "ah_resolve",
    "ah_device_malloc", "ah_device_free", "ah_host_alloc_pinned", "ah_host_free_pinned",
    "ah_memcpy_d2h", "ah_memcpy_h2d", "ah_set_identity", "ah_apply_givens_rows", "ah_gather_permuted",
    "ah_scale_rows_phase", "ah_gather_permuted_phase", "ah_symarrow_eigvals_dd", "ah_symarrow_eigen_batched", "ah_gather_elements", "ah_gather_permuted_batched",
    "ah_copy2d_batched", "ah_dgemm_batched",
    "ah_measure_fp64_peak", "ah_measure_term_peak", "ah_measure_rcp_seed_error", "ah_measure_store_bw",
    "ah_set_async", "ah_wait", "ah_last_gemm_ms", "ah_measure_dmma_peak",
    "ah_multi_create", "ah_multi_destroy", "ah_multi_ndev", "ah_multi_ctx", "ah_multi_last_error", "ah_multi_elapsed_ms",
    "ah_k_partition", "ah_multi_symarrow_eigen", "ah_multi_symdpr1_eigen", "ah_multi_halfarrow_svd",
    "ah_multi_allgather_columns", "ah_nccl_unique_id", "ah_nccl_init", "ah_allgather_columns",
]


class ArrowheadError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libarrowhead_b200 error {code}: {msg}")
        self.code = code


def load_library(path: str | None = None) -> C.CDLL:
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  arrowhead_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.ah_version.restype = C.c_int
    lib.ah_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.ah_destroy.argtypes = [vp]
    lib.ah_destroy.restype = None
    lib.ah_last_error.argtypes = [vp]
    lib.ah_last_error.restype = C.c_char_p
    lib.ah_set_stream.argtypes = [vp, vp]
    lib.ah_get_stats.argtypes = [vp, C.POINTER(AhStats)]
    lib.ah_set_mode.argtypes = [vp, C.c_int]
    lib.ah_symarrow_eigen.argtypes = [vp, C.c_int64, _dp, _dp, C.c_double, _dp, C.c_int64, C.c_int64, C.POINTER(AhResult)]
    lib.ah_symdpr1_eigen.argtypes = [vp, C.c_int64, _dp, _dp, C.c_double, _dp, C.c_int64, C.c_int64, C.POINTER(AhResult)]
    lib.ah_halfarrow_svd.argtypes = [vp, C.c_int64, C.c_int64, _dp, _dp, _dp, C.c_int64, C.c_int64, C.POINTER(AhResult)]
    lib.ah_resolve.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(AhResult)]
    lib.ah_device_malloc.argtypes = [vp, C.POINTER(vp), C.c_int64]
    lib.ah_device_free.argtypes = [vp, vp]
    lib.ah_host_alloc_pinned.argtypes = [vp, C.POINTER(vp), C.c_int64]
    lib.ah_host_free_pinned.argtypes = [vp, vp]
    lib.ah_memcpy_d2h.argtypes = [vp, vp, vp, C.c_int64]
    lib.ah_memcpy_h2d.argtypes = [vp, vp, vp, C.c_int64]
    lib.ah_set_identity.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64]
    lib.ah_apply_givens_rows.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double]
    lib.ah_gather_permuted.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, _i64p, C.c_int64, _i64p, C.c_int64]
    lib.ah_gather_permuted_phase.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, _i64p, C.c_int64, _i64p, C.c_int64, _dp, C.c_int]
    u64p = C.POINTER(C.c_uint64)
    lib.ah_symarrow_eigvals_dd.argtypes = [vp, C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, _dp, C.c_int64, C.c_int64,
                                           C.POINTER(AhResult)]
    lib.ah_symarrow_eigen_batched.argtypes = [vp, C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, C.POINTER(AhResult), C.c_int64]
    lib.ah_gather_elements.argtypes = [vp, vp, _i64p, C.c_int64, _dp]
    lib.ah_gather_permuted_batched.argtypes = [vp, C.c_int64, vp, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int64, _i64p, C.c_int64,
                                               _i64p, C.c_int64]
    lib.ah_copy2d_batched.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, u64p, C.c_int64, u64p, C.c_int64]
    lib.ah_dgemm_batched.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, u64p, C.c_int64, u64p, C.c_int64, u64p, C.c_int64]
    lib.ah_scale_rows_phase.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, _dp, C.c_int, C.c_int64, C.c_int64]
    lib.ah_set_async.argtypes = [vp, C.c_int]
    lib.ah_wait.argtypes = [vp]
    lib.ah_last_gemm_ms.argtypes = [vp, _dp]
    lib.ah_measure_dmma_peak.argtypes = [vp, _dp, _dp]
    lib.ah_measure_fp64_peak.argtypes = [vp, _dp, _dp]
    lib.ah_measure_term_peak.argtypes = [vp, _dp, _dp]
    lib.ah_measure_rcp_seed_error.argtypes = [vp, _dp, _dp]
    lib.ah_measure_store_bw.argtypes = [vp, C.c_int64, _dp]
    pp = C.POINTER(vp)
    lib.ah_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    lib.ah_multi_destroy.argtypes = [vp]
    lib.ah_multi_ndev.argtypes = [vp]
    lib.ah_multi_ctx.argtypes = [vp, C.c_int]
    lib.ah_multi_last_error.argtypes = [vp]
    lib.ah_multi_elapsed_ms.argtypes = [vp]
    lib.ah_k_partition.argtypes = [C.c_int64, C.c_int, C.c_int, _i64p, _i64p]
    lib.ah_multi_symarrow_eigen.argtypes = [vp, C.c_int64, _dp, _dp, C.c_double, _dp, C.POINTER(AhResult), pp]
    lib.ah_multi_symdpr1_eigen.argtypes = [vp, C.c_int64, _dp, _dp, C.c_double, _dp, C.POINTER(AhResult), pp]
    lib.ah_multi_halfarrow_svd.argtypes = [vp, C.c_int64, C.c_int64, _dp, _dp, _dp, C.POINTER(AhResult), pp, pp]
    lib.ah_multi_allgather_columns.argtypes = [vp, pp, C.c_int64, C.c_int64, _dp]
    lib.ah_nccl_unique_id.argtypes = [vp]
    lib.ah_nccl_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.ah_allgather_columns.argtypes = [vp, vp, C.c_int64, C.c_int64, _dp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("ah_destroy", "ah_last_error", "ah_multi_destroy", "ah_multi_ctx", "ah_multi_last_error", "ah_multi_elapsed_ms",
                        "ah_k_partition"):
            fn.restype = C.c_int
    lib.ah_multi_destroy.restype = None
    lib.ah_k_partition.restype = None
    lib.ah_multi_ctx.restype = vp
    lib.ah_multi_last_error.restype = C.c_char_p
    lib.ah_multi_elapsed_ms.restype = C.c_double
    return lib


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class RangeResult:
    """Per-k outputs of a k-range solve (entry c belongs to eigenpair k0+c)."""

    def __init__(self, k0, k1):
        cnt = k1 - k0 + 1
        self.k0, self.k1 = k0, k1
        self.lam = np.zeros(cnt)
        self.sind = np.zeros(cnt, np.int64)
        self.kb = np.zeros(cnt)
        self.kz = np.zeros(cnt)
        self.knu = np.zeros(cnt)
        self.krho = np.zeros(cnt)
        self.qout = np.zeros(cnt, np.int64)
        self.steps = np.zeros(cnt, np.int32)
        self.status = np.zeros(cnt, np.int32)
        self.U = None     # host ndarray (Fortran order) or device pointer owner
        self.V = None
        self.stats = None

    def fill(self, r: AhResult):
        r.lambda_ = _ptr(self.lam, _dp)
        r.sind = _ptr(self.sind, _i64p)
        r.kb = _ptr(self.kb, _dp); r.kz = _ptr(self.kz, _dp); r.knu = _ptr(self.knu, _dp); r.krho = _ptr(self.krho, _dp)
        r.qout = _ptr(self.qout, _i64p)
        r.steps = _ptr(self.steps, _i32p)
        r.status = _ptr(self.status, _i32p)


class Context:
    """One ah_ctx: one GPU."""

    def __init__(self, device: int = 0, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.handle = C.c_void_p()
        rc = self.lib.ah_create(C.byref(self.handle), int(device))
        if rc != 0:
            raise ArrowheadError(rc, (self.lib.ah_last_error(None) or b"").decode())
        self.device = device

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.ah_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ArrowheadError(rc, (self.lib.ah_last_error(self.handle) or b"").decode())

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.ah_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def set_mode(self, mode: int):
        self._check(self.lib.ah_set_mode(self.handle, int(mode)))

    def stats(self) -> dict:
        s = AhStats()
        self._check(self.lib.ah_get_stats(self.handle, C.byref(s)))
        return s.asdict()

    # ---- the three hot-path entry points ---------------------------------------------------------
    def _run(self, fn, args, k0, k1, nrows_u, nrows_v, vectors, U_dev, ldu, V_dev, ldv, rowmap_u, rowmap_v,
             rowscale_v, U_host=None, V_host=None, into=None):
        # `into`: a RangeResult of the same k-range to fill again (callers that solve repeatedly avoid fresh pages for
        # the nine per-k arrays: at n = 32768 first-touch page faults cost 0.2 ms per call)
        res = into if (into is not None and into.k0 == k0 and into.k1 == k1) else RangeResult(k0, k1)
        r = getattr(res, "_ah", None)            # a reused RangeResult keeps its filled-in struct (the nine numpy -> ctypes pointer
        if r is None:                            # conversions cost ~30 us of Python per call, 1 % of a 1/8-shard step)
            r = AhResult()
            res.fill(r)
            if res is into:
                res._ah = r
        else:                                    # everything but the nine per-k pointers is set afresh below
            r.U = None; r.ldu = 0; r.U_mem = 0; r.V = None; r.ldv = 0; r.V_mem = 0
            r.rowmap_u = None; r.rowmap_v = None; r.rowscale_v = None
        cnt = k1 - k0 + 1
        keep = []
        if U_dev is not None:
            r.U, r.ldu, r.U_mem = C.c_void_p(int(U_dev)), int(ldu), AH_MEM_DEVICE
        elif U_host is not None:
            r.U, r.ldu, r.U_mem = C.c_void_p(int(U_host)), int(ldu), AH_MEM_HOST
        elif vectors:
            res.U = np.zeros((nrows_u, cnt), order="F")
            r.U, r.ldu, r.U_mem = res.U.ctypes.data_as(C.c_void_p), nrows_u, AH_MEM_HOST
        if nrows_v:
            if V_dev is not None:
                r.V, r.ldv, r.V_mem = C.c_void_p(int(V_dev)), int(ldv), AH_MEM_DEVICE
            elif V_host is not None:
                r.V, r.ldv, r.V_mem = C.c_void_p(int(V_host)), int(ldv), AH_MEM_HOST
            elif vectors:
                res.V = np.zeros((nrows_v, cnt), order="F")
                r.V, r.ldv, r.V_mem = res.V.ctypes.data_as(C.c_void_p), nrows_v, AH_MEM_HOST
        for name, arr, t in (("rowmap_u", rowmap_u, np.int64), ("rowmap_v", rowmap_v, np.int64), ("rowscale_v", rowscale_v, np.float64)):
            if arr is not None:
                a = np.ascontiguousarray(arr, t)
                keep.append(a)
                setattr(r, name, _ptr(a, _i64p if t is np.int64 else _dp))
        self._check(fn(self.handle, *args, int(k0), int(k1), C.byref(r)))
        if fn is not self.lib.ah_resolve:
            self._resident = (nrows_u, nrows_v)      # ah_resolve solves this problem again from the device-resident inputs
        res.stats = self.stats()
        return res

    def symarrow(self, D, z, a, tau=None, k0=1, k1=None, vectors=True, U_dev=None, ldu=0, rowmap=None, U_host=None, into=None):
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64)
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        N = D.size
        k1 = N + 1 if k1 is None else k1
        args = (C.c_int64(N), _ptr(D, _dp), _ptr(z, _dp), C.c_double(float(a)), _ptr(tau_a, _dp))
        return self._run(self.lib.ah_symarrow_eigen, args, k0, k1, N + 1, 0, vectors, U_dev, ldu, None, 0, rowmap, None,
                         None, U_host=U_host, into=into)

    def resolve(self, k0, k1, vectors=True, U_dev=None, ldu=0, V_dev=None, ldv=0, rowmap_u=None, rowmap_v=None,
                rowscale_v=None, U_host=None, into=None):
        """ah_resolve: another k-range of the problem of the last symarrow / symdpr1 / halfarrow call on this context
        (inputs still resident on the device: no validation, padding or host->device copy)."""
        if getattr(self, "_resident", None) is None:
            raise ArrowheadError(-1, "resolve(): no solve on this context yet")
        nrows_u, nrows_v = self._resident
        return self._run(self.lib.ah_resolve, (), k0, k1, nrows_u, nrows_v, vectors, U_dev, ldu, V_dev, ldv, rowmap_u,
                         rowmap_v, rowscale_v, U_host=U_host, into=into)

    def symdpr1(self, D, u, rho, tau=None, k0=1, k1=None, vectors=True, U_dev=None, ldu=0, rowmap=None, U_host=None, into=None):
        D = np.ascontiguousarray(D, np.float64); u = np.ascontiguousarray(u, np.float64)
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        n = D.size
        k1 = n if k1 is None else k1
        args = (C.c_int64(n), _ptr(D, _dp), _ptr(u, _dp), C.c_double(float(rho)), _ptr(tau_a, _dp))
        return self._run(self.lib.ah_symdpr1_eigen, args, k0, k1, n, 0, vectors, U_dev, ldu, None, 0, rowmap, None, None,
                         U_host=U_host, into=into)

    def halfarrow(self, D, z, tau=None, k0=1, k1=None, vectors=True, U_dev=None, ldu=0, V_dev=None, ldv=0,
                  rowmap_u=None, rowmap_v=None, rowscale_v=None, into=None):
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64)
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        nd, nz = D.size, z.size
        k1 = nz if k1 is None else k1
        args = (C.c_int64(nd), C.c_int64(nz), _ptr(D, _dp), _ptr(z, _dp), _ptr(tau_a, _dp))
        return self._run(self.lib.ah_halfarrow_svd, args, k0, k1, nz, nd + 1, vectors, U_dev, ldu, V_dev, ldv, rowmap_u,
                         rowmap_v, rowscale_v, into=into)

    def symarrow_eigvals_dd(self, D, z, zlo, a, alo, tau, k0=1, k1=None):
        """eigenvalues only, double-double border (z+zlo, a+alo): rootsah's per-k solver (arrowhead7.jl:351-495)"""
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64); zlo = np.ascontiguousarray(zlo, np.float64)
        tau_a = np.ascontiguousarray(tau, np.float64)
        N = D.size
        k1 = N + 1 if k1 is None else k1
        res = RangeResult(k0, k1)
        r = AhResult()
        res.fill(r)
        self._check(self.lib.ah_symarrow_eigvals_dd(self.handle, N, _ptr(D, _dp), _ptr(z, _dp), _ptr(zlo, _dp), float(a), float(alo),
                                                    _ptr(tau_a, _dp), int(k0), int(k1), C.byref(r)))
        res.stats = self.stats()
        return res

    # ---- batched primitives (tdc) -----------------------------------------------------------------
    def symarrow_batched(self, D, z, a, tau=None, U_dev=None, ldu=0, ustride=0):
        """nb ordered SymArrow problems with N poles each: D, z (nb, N), a (nb,) -> RangeResult with (nb*(N+1),) arrays."""
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64); a = np.ascontiguousarray(a, np.float64)
        nb, N = D.shape
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        res = RangeResult(1, nb * (N + 1))
        r = AhResult()
        res.fill(r)
        if U_dev is not None:
            r.U, r.ldu, r.U_mem = C.c_void_p(int(U_dev)), int(ldu), AH_MEM_DEVICE
        self._check(self.lib.ah_symarrow_eigen_batched(self.handle, nb, N, _ptr(D, _dp), _ptr(z, _dp), _ptr(a, _dp), _ptr(tau_a, _dp),
                                                       C.byref(r), int(ustride)))
        res.stats = self.stats()
        return res

    def gather_elements(self, src_dev, idx):
        idx = np.ascontiguousarray(idx, np.int64)
        out = np.empty(idx.size)
        self._check(self.lib.ah_gather_elements(self.handle, C.c_void_p(int(src_dev)), _ptr(idx, _i64p), idx.size, _ptr(out, _dp)))
        return out

    def gather_permuted_batched(self, dst_dev, dst_stride, ld_dst, src_dev, src_stride, ld_src, rows, cols):
        rows = np.ascontiguousarray(rows, np.int64); cols = np.ascontiguousarray(cols, np.int64)
        self._check(self.lib.ah_gather_permuted_batched(self.handle, rows.shape[0], C.c_void_p(int(dst_dev)), dst_stride, ld_dst,
                                                        C.c_void_p(int(src_dev)), src_stride, ld_src, _ptr(rows, _i64p),
                                                        rows.shape[1], _ptr(cols, _i64p), cols.shape[1]))

    def copy2d_batched(self, nrows, ncols, dst_ptrs, ld_dst, src_ptrs, ld_src):
        d = np.ascontiguousarray(dst_ptrs, np.uint64); s_ = np.ascontiguousarray(src_ptrs, np.uint64)
        u64p = C.POINTER(C.c_uint64)
        self._check(self.lib.ah_copy2d_batched(self.handle, d.size, nrows, ncols, d.ctypes.data_as(u64p), ld_dst,
                                               s_.ctypes.data_as(u64p), ld_src))

    def dgemm_batched(self, m, n, k, A_ptrs, lda, B_ptrs, ldb, C_ptrs, ldc):
        A = np.ascontiguousarray(A_ptrs, np.uint64); B = np.ascontiguousarray(B_ptrs, np.uint64); Cc = np.ascontiguousarray(C_ptrs, np.uint64)
        u64p = C.POINTER(C.c_uint64)
        self._check(self.lib.ah_dgemm_batched(self.handle, A.size, m, n, k, A.ctypes.data_as(u64p), lda, B.ctypes.data_as(u64p), ldb,
                                              Cc.ctypes.data_as(u64p), ldc))

    # ---- device utilities -------------------------------------------------------------------------
    def set_identity(self, U_dev, nrows, ncols, ld):
        self._check(self.lib.ah_set_identity(self.handle, C.c_void_p(int(U_dev)), nrows, ncols, ld))

    def apply_givens_rows(self, U_dev, ld, ncols, i1, i2, c, s):
        self._check(self.lib.ah_apply_givens_rows(self.handle, C.c_void_p(int(U_dev)), ld, ncols, i1, i2, c, s))

    def gather_permuted(self, dst_dev, ld_dst, src_dev, ld_src, rows, cols):
        rows = np.ascontiguousarray(rows, np.int64); cols = np.ascontiguousarray(cols, np.int64)
        self._check(self.lib.ah_gather_permuted(self.handle, C.c_void_p(int(dst_dev)), ld_dst, C.c_void_p(int(src_dev)),
                                                ld_src, _ptr(rows, _i64p), rows.size, _ptr(cols, _i64p), cols.size))

    def gather_permuted_phase(self, dst_dev, ld_dst, src_dev, ld_src, rows, cols, phase):
        rows = np.ascontiguousarray(rows, np.int64); cols = np.ascontiguousarray(cols, np.int64)
        phase = np.ascontiguousarray(phase, np.float64)
        self._check(self.lib.ah_gather_permuted_phase(self.handle, C.c_void_p(int(dst_dev)), ld_dst, C.c_void_p(int(src_dev)), ld_src,
                                                      _ptr(rows, _i64p), rows.size, _ptr(cols, _i64p), cols.size, _ptr(phase, _dp),
                                                      int(phase.shape[1])))

    def scale_rows_phase(self, dst_dev, ld_dst, src_dev, ld_src, phase, nrows, ncols):
        phase = np.ascontiguousarray(phase, np.float64)
        self._check(self.lib.ah_scale_rows_phase(self.handle, C.c_void_p(int(dst_dev)), ld_dst, C.c_void_p(int(src_dev)),
                                                 ld_src, _ptr(phase, _dp), int(phase.shape[1]), nrows, ncols))

    def memcpy_d2h(self, dst: np.ndarray, src_dev, nbytes):
        self._check(self.lib.ah_memcpy_d2h(self.handle, dst.ctypes.data_as(C.c_void_p), C.c_void_p(int(src_dev)), nbytes))

    def memcpy_h2d(self, dst_dev, src: np.ndarray, nbytes):
        src = np.ascontiguousarray(src)
        self._check(self.lib.ah_memcpy_h2d(self.handle, C.c_void_p(int(dst_dev)), src.ctypes.data_as(C.c_void_p), nbytes))

    def pinned_alloc(self, nbytes) -> int:
        p = C.c_void_p()
        self._check(self.lib.ah_host_alloc_pinned(self.handle, C.byref(p), nbytes))
        return p.value

    def pinned_free(self, ptr):
        self._check(self.lib.ah_host_free_pinned(self.handle, C.c_void_p(ptr)))

    # ---- NCCL all-gather of U, one process per GPU (ah_nccl_init once, then ah_allgather_columns) ----------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.ah_nccl_unique_id(buf)
        if rc != 0:
            raise ArrowheadError(rc, (self.lib.ah_last_error(None) or b"").decode())
        return buf.raw

    def nccl_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(bytes(uid), 128)
        self._check(self.lib.ah_nccl_init(self.handle, int(nranks), int(rank), buf))

    def allgather_columns(self, U_full_dev, ld, kmax) -> float:
        ms = C.c_double()
        self._check(self.lib.ah_allgather_columns(self.handle, C.c_void_p(int(U_full_dev)), int(ld), int(kmax), C.byref(ms)))
        return ms.value

    def set_async(self, on: bool):
        """device utilities enqueue and return (True) or synchronise before returning (False, the default)"""
        self._check(self.lib.ah_set_async(self.handle, 1 if on else 0))

    def wait(self):
        self._check(self.lib.ah_wait(self.handle))

    def last_gemm_ms(self) -> float:
        ms = C.c_double()
        self._check(self.lib.ah_last_gemm_ms(self.handle, C.byref(ms)))
        return ms.value

    def measure_dmma_peak(self):
        g, ms = C.c_double(), C.c_double()
        self._check(self.lib.ah_measure_dmma_peak(self.handle, C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def measure_fp64_peak(self):
        g, ms = C.c_double(), C.c_double()
        self._check(self.lib.ah_measure_fp64_peak(self.handle, C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def measure_term_peak(self):
        g, ms = C.c_double(), C.c_double()
        self._check(self.lib.ah_measure_term_peak(self.handle, C.byref(g), C.byref(ms)))
        return g.value, ms.value

    def measure_rcp_seed_error(self):
        """exhaustive max |1 - x*r| of the MUFU.RCP64H seed and of the corrected reciprocal of the secular term"""
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.ah_measure_rcp_seed_error(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_store_bw(self, nbytes):
        g = C.c_double()
        self._check(self.lib.ah_measure_store_bw(self.handle, nbytes, C.byref(g)))
        return g.value


class MultiContext:
    """ah_multi: several GPUs driven from this one thread (worker threads live inside the library)."""

    def __init__(self, devices=None, ndev: int = 0, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.handle = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.ah_multi_create(C.byref(self.handle), arr, len(devices))
        else:
            rc = self.lib.ah_multi_create(C.byref(self.handle), None, int(ndev))
        if rc != 0:
            raise ArrowheadError(rc, (self.lib.ah_last_error(None) or b"").decode())
        self.ndev = int(self.lib.ah_multi_ndev(self.handle))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.ah_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ArrowheadError(rc, (self.lib.ah_multi_last_error(self.handle) or b"").decode())

    def partition(self, kmax: int, r: int):
        k0, k1 = C.c_int64(), C.c_int64()
        self.lib.ah_k_partition(int(kmax), self.ndev, int(r), C.byref(k0), C.byref(k1))
        return k0.value, k1.value

    def device_stats(self, r: int) -> dict:
        s = AhStats()
        rc = self.lib.ah_get_stats(C.c_void_p(self.lib.ah_multi_ctx(self.handle, int(r))), C.byref(s))
        if rc != 0:
            raise ArrowheadError(rc, "ah_get_stats")
        return s.asdict()

    def _ptrs(self, bufs):
        if bufs is None:
            return None, None
        arr = (C.c_void_p * self.ndev)(*[C.c_void_p(int(b)) for b in bufs])
        return arr, C.cast(arr, C.POINTER(C.c_void_p))

    def _solve(self, fn, args, kmax, nrows_u, nrows_v, vectors, U_full, V_full, U_host, ldu, V_host=None, ldv=0):
        res = RangeResult(1, kmax)
        r = AhResult()
        res.fill(r)
        if U_full is None:
            if U_host is not None:
                r.U, r.ldu, r.U_mem = C.c_void_p(int(U_host)), int(ldu), AH_MEM_HOST
            elif vectors:
                res.U = np.zeros((nrows_u, kmax), order="F")
                r.U, r.ldu, r.U_mem = res.U.ctypes.data_as(C.c_void_p), nrows_u, AH_MEM_HOST
        else:
            r.ldu = int(ldu or nrows_u)
        if nrows_v:
            if V_full is None:
                if V_host is not None:
                    r.V, r.ldv, r.V_mem = C.c_void_p(int(V_host)), int(ldv), AH_MEM_HOST
                elif vectors:
                    res.V = np.zeros((nrows_v, kmax), order="F")
                    r.V, r.ldv, r.V_mem = res.V.ctypes.data_as(C.c_void_p), nrows_v, AH_MEM_HOST
            else:
                r.ldv = int(ldv or nrows_v)
        ku, pu = self._ptrs(U_full)
        extra = (pu,)
        if nrows_v:
            kv, pv = self._ptrs(V_full)
            extra = (pu, pv)
        self._check(fn(self.handle, *args, C.byref(r), *extra))
        res.stats = {"elapsed_ms": float(self.lib.ah_multi_elapsed_ms(self.handle)),
                     "devices": [self.device_stats(i) for i in range(self.ndev)]}
        return res

    def symarrow(self, D, z, a, tau=None, vectors=True, U_full=None, U_host=None, ldu=0):
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64)
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        N = D.size
        args = (C.c_int64(N), _ptr(D, _dp), _ptr(z, _dp), C.c_double(float(a)), _ptr(tau_a, _dp))
        return self._solve(self.lib.ah_multi_symarrow_eigen, args, N + 1, N + 1, 0, vectors, U_full, None, U_host, ldu)

    def symdpr1(self, D, u, rho, tau=None, vectors=True, U_full=None, U_host=None, ldu=0):
        D = np.ascontiguousarray(D, np.float64); u = np.ascontiguousarray(u, np.float64)
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        n = D.size
        args = (C.c_int64(n), _ptr(D, _dp), _ptr(u, _dp), C.c_double(float(rho)), _ptr(tau_a, _dp))
        return self._solve(self.lib.ah_multi_symdpr1_eigen, args, n, n, 0, vectors, U_full, None, U_host, ldu)

    def halfarrow(self, D, z, tau=None, vectors=True, U_full=None, V_full=None, ldu=0, ldv=0):
        D = np.ascontiguousarray(D, np.float64); z = np.ascontiguousarray(z, np.float64)
        tau_a = None if tau is None else np.ascontiguousarray(tau, np.float64)
        nd, nz = D.size, z.size
        args = (C.c_int64(nd), C.c_int64(nz), _ptr(D, _dp), _ptr(z, _dp), _ptr(tau_a, _dp))
        return self._solve(self.lib.ah_multi_halfarrow_svd, args, nz, nz, nd + 1, vectors, U_full, V_full, None, ldu, None, ldv)

    def allgather_columns(self, U_full, ld, kmax) -> float:
        """in-place NCCL all-gather of the column blocks of the per-device full-size buffers; returns device ms"""
        keep, pu = self._ptrs(U_full)
        ms = C.c_double()
        self._check(self.lib.ah_multi_allgather_columns(self.handle, pu, int(ld), int(kmax), C.byref(ms)))
        return ms.value


_CTX: dict[int, Context] = {}


def get_context(device: int = 0) -> Context:
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]
