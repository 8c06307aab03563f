/*
 * arrowhead_b200.h -- C ABI of libarrowhead_b200.so
 *
 * B200 (sm_100a) replacement for the per-eigenpair loops of ivanslapnicar/Arrowhead.jl.
 * The reference's public API is Julia dispatch (`eigen(::SymArrow)`, `eigen(::SymDPR1)`,
 * `svd(::HalfArrow)`, src/arrowhead1.jl:2, src/Arrowhead.jl:9).  Julia keeps the types, the sorting,
 * the deflation and the final permutation; what this library replaces is exactly the body of the
 * `Threads.@threads for k` loops, i.e. "solve eigenpairs k0..k1 of an ordered irreducible problem":
 *
 *   ah_symarrow_eigen   <->  src/arrowhead3.jl:627-632 and :639-644  (calls eigen(A,Ainv,k), :419-550)
 *   ah_symdpr1_eigen    <->  src/arrowhead4.jl:516-522 and :529-535  (calls eigen(A,Ainv,k), :289-422)
 *   ah_halfarrow_svd    <->  src/arrowhead6.jl:622-628 and :635-642  (calls svd(A,Ainv,k),   :344-508)
 *   ah_apply_givens_rows<->  src/arrowhead3.jl:633-636, arrowhead4.jl:523-526, arrowhead6.jl:629-632 (lmul!(R,U))
 *   ah_set_identity     <->  src/arrowhead3.jl:574 (U=Array{T}(I,n,n)) for a device-resident U
 *   ah_gather_permuted  <->  src/arrowhead3.jl:652 (view(U,rows,es)) materialised on the device
 *
 * Plain C: no C++ or torch types cross this boundary.  All entry points return 0 on success, a negative
 * AH_ERR_* code for invalid arguments and a positive value (1000 + cudaError_t) for CUDA failures;
 * ah_last_error() gives the message.  Nothing throws or aborts.  Calls block until their results are
 * complete (they synchronise the context's stream).  One ah_ctx is bound to one GPU and is not
 * re-entrant; distinct contexts are independent (one process per GPU, or one context per GPU).
 *
 * Index conventions follow the reference: eigenpair numbers k are 1-based, descending eigenvalue order
 * of the ordered problem; `sind` is the 1-based shift index; matrices are column-major.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with AH_ERR_NO_DEVICE.
 */
#ifndef ARROWHEAD_B200_H
#define ARROWHEAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AH_VERSION 100          /* 0.1.0 */

/* return codes */
#define AH_OK 0
#define AH_ERR_INVALID_ARG (-1)   /* null pointer, bad sizes, bad k-range, bad leading dimension */
#define AH_ERR_NOT_ORDERED (-2)   /* D not strictly decreasing (arrowhead3.jl:407 "ordered") */
#define AH_ERR_REDUCIBLE (-3)     /* a zero (or non-finite) entry in z/u (arrowhead3.jl:407 "irreducible") */
#define AH_ERR_NO_DEVICE (-4)     /* no usable CUDA device / wrong architecture */
#define AH_ERR_NOMEM (-5)
#define AH_ERR_CUDA_BASE 1000     /* 1000 + cudaError_t */

/* per-eigenpair status bits (same meaning in the CPU oracle) */
#define AH_ST_OK 0
#define AH_ST_NU_INF 1            /* |nu|==Inf deflation branch (arrowhead3.jl:451-455): v=e_i, lambda=sigma */
#define AH_ST_NEEDS_BIGFLOAT 2    /* reference switches to 256-bit BigFloat here (Qout 50/100): redo k on the host */
#define AH_ST_POLE_RETURN 4       /* Remedy 3 "came back with a pole" (arrowhead3.jl:474-496; raises in the reference) */
#define AH_ST_REF_THROWS 8        /* the reference raises on this input (e.g. arrowhead4.jl:324) */
#define AH_ST_ITER_LIMIT 16       /* Remedy-3 loop guard hit (the reference would keep looping) */
#define AH_ST_INTERLACE 32        /* the computed value violates D_k <= lambda_k <= D_{k-1}: the reference's scheme lost all
                                     accuracy here (first-stage K_nu ~ 1/eps, Remedy 3 re-shifted past the eigenvalue);
                                     the reference returns such values silently -- redo this k in higher precision */

/* Branch trace, bits 8..14 of the same status word (informational: `status & 0xff` are the conditions above).
 * Which non-standard branches of the reference's per-k solver the eigenpair went through; the CPU oracle sets the
 * same bits, so the parity tests compare the branch taken for every eigenpair, not only the results. */
#define AH_BR_DD_B 0x100          /* b recomputed in Dekker double-double (inv!: Kb/Kz test failed, arrowhead3.jl:191-230) */
#define AH_BR_REM3 0x200          /* Remedy 3: re-shift between the poles (K_nu > tolnu, arrowhead3.jl:466-513) */
#define AH_BR_REM3_REPEAT 0x400   /* ... more than once (the `while Knu > tau[3]` loop went round again) */
#define AH_BR_RL_RETRY 0x800      /* 'R' -> 'L' retry of the DPR1 bisection (arrowhead3.jl:499-501) */
#define AH_BR_DD_RHO 0x1000       /* rho of an inverse at a non-pole shift in double-double (K_rho >= tolrho, :292-321) */
#define AH_BR_REM1 0x2000         /* Remedy 1: Rayleigh quotient with the inverse of the unshifted matrix (:526-544) */
#define AH_BR_REM1_ZERO 0x4000    /* ... whose rho is infinite: lambda = 0 (:538-539) */
#define AH_ST_MASK 0xff           /* status & AH_ST_MASK == 0: the eigenpair is complete and as the reference returns it */

/* where an eigenvector buffer lives */
#define AH_MEM_HOST 0
#define AH_MEM_DEVICE 1

typedef struct ah_ctx ah_ctx;

/* Outputs of one k-range solve; cnt = k1-k0+1.  Entry c (0-based) belongs to eigenpair k0+c.
 * Every pointer may be NULL when that output is not wanted, except `lambda`.
 * The info arrays are the reference's Info(Sind,Kb,Kz,Knu,Krho,Qout) (arrowhead3.jl:395-402) as SoA
 * (a Julia Vector{Info} is a vector of boxed mutable structs and cannot cross a ccall). */
typedef struct ah_result {
    double  *lambda;     /* [cnt] eigenvalues (singular values for svd), host */
    int64_t *sind;       /* [cnt] shift index i (1-based), host */
    double  *kb;         /* [cnt] host */
    double  *kz;         /* [cnt] host */
    double  *knu;        /* [cnt] host */
    double  *krho;       /* [cnt] host */
    int64_t *qout;       /* [cnt] host */
    int32_t *steps;      /* [cnt] O(N) sweeps (evaluations of the secular function) spent on this eigenpair -- not in the reference;
                          * for roofline accounting.  Plain bisection takes ~54; the library decides most of those steps from
                          * certified bounds and evaluates ~15, with the same final bracket bit for bit (DESIGN.md section 3;
                          * environment ARROWHEAD_B200_ACCEL=0 at ah_create selects plain bisection) */
    int32_t *status;     /* [cnt] AH_ST_* bits, host */
    /* Eigenvectors.  Column c of U starts at U + c*ldu.  Entry r of the vector goes to row
     * rowmap[r] (0-based) when rowmap != NULL (the deflation case `U[zx,zx[k]] = v`,
     * arrowhead3.jl:631,643), else to row r.  U_mem says whether U is a host or a device pointer
     * (device: must belong to the context's GPU; it stays resident, nothing is copied back). */
    double  *U;          /* SymArrow: (N+1) rows; SymDPR1: n rows; HalfArrow: left vectors, nz rows */
    int64_t  ldu;
    int32_t  U_mem;
    double  *V;          /* HalfArrow only: right vectors, nd+1 rows */
    int64_t  ldv;
    int32_t  V_mem;
    const int64_t *rowmap_u;   /* host, length = vector length, or NULL */
    const int64_t *rowmap_v;   /* host, HalfArrow only */
    const double  *rowscale_v; /* host, [nd] or NULL: V rows 0..nd-1 multiplied by this (signD, arrowhead6.jl:641) */
} ah_result;

/* timing + accounting of the last compute call on a context */
typedef struct ah_stats {
    double  kernel_ms;        /* device time of all kernels of the call (CUDA events on the context stream) */
    double  main_kernel_ms;   /* device time of the dominant kernel: k_bisect round 0 (bisect_ms) + round 1 (rem3_ms) */
    double  h2d_ms, d2h_ms;   /* host<->device copies issued by the call */
    int64_t launches;         /* number of kernels launched by the call */
    int64_t n_fast;           /* eigenpairs finished by the fast-path kernels (k_prep -> k_bisect -> k_vec) */
    int64_t n_full;           /* eigenpairs routed to the full (Remedy / double-double) kernel */
    int64_t total_steps;      /* sum of `steps` (sweeps evaluated) */
    int64_t h2d_bytes, d2h_bytes;
    double  prep_ms, bisect_ms, vec_ms, full_ms;   /* per-kernel device time: k_prep, k_bisect (+ Remedy-3 round), k_vec, k_full */
    int64_t n_rem3;           /* eigenpairs that took the Remedy-3 round of the fast path (counted in n_fast) */
    double  rem3_ms;          /* k_rem3_prep + k_bisect round 1 (device time on the side stream; overlaps k_vec) */
    double  host_pre_ms;      /* host wall time from entry to the first kernel launch (padding, H2D enqueue, allocation) */
    double  host_post_ms;     /* host wall time after the device finished (unpacking the per-k outputs) */
} ah_stats;

/* ---- context ---------------------------------------------------------------------------------- */
int  ah_version(void);
int  ah_create(ah_ctx **ctx, int device);       /* binds to CUDA device `device`; needs compute capability 10.x */
void ah_destroy(ah_ctx *ctx);
const char *ah_last_error(const ah_ctx *ctx);   /* ctx may be NULL: last error of a failed ah_create */
int  ah_set_stream(ah_ctx *ctx, void *cuda_stream);   /* run on a caller-owned cudaStream_t (NULL: own stream) */
int  ah_get_stats(const ah_ctx *ctx, ah_stats *out);
/* 0: fast path + full kernel for the eigenpairs that need it (default); 1: full kernel for every k */
int  ah_set_mode(ah_ctx *ctx, int mode);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* SymArrow [diag(D) z; z' a], N = length(D) poles, arrow top last; D strictly decreasing, z nonzero.
 * Solves eigenpairs k0..k1 (1 <= k0 <= k1 <= N+1).  tau = [tolb,tolz,tolnu,tolrho,tollambda]
 * (NULL: the per-k default [1e3, 10N, 1e3, 1e3, 1e3], arrowhead3.jl:420). */
int ah_symarrow_eigen(ah_ctx *ctx, int64_t N, const double *D, const double *z, double a,
                      const double *tau, int64_t k0, int64_t k1, ah_result *out);

/* SymDPR1 diag(D)+rho*u*u', n = length(D) >= 2, rho > 0, D strictly decreasing, u nonzero.
 * 1 <= k0 <= k1 <= n. */
int ah_symdpr1_eigen(ah_ctx *ctx, int64_t n, const double *D, const double *u, double rho,
                     const double *tau, int64_t k0, int64_t k1, ah_result *out);

/* HalfArrow [diag(D) z], D > 0 strictly decreasing (nd entries), z nonzero with nz = nd (no tip) or
 * nd+1 (tip) entries.  Singular triples k0..k1, 1 <= k0 <= k1 <= nz. */
int ah_halfarrow_svd(ah_ctx *ctx, int64_t nd, int64_t nz, const double *D, const double *z,
                     const double *tau, int64_t k0, int64_t k1, ah_result *out);

/* Solve again, for another k-range (or into another U), the problem of the last ah_symarrow_eigen /
 * ah_symdpr1_eigen / ah_halfarrow_svd call on this context: its padded inputs are still resident on the device, so
 * validation, padding and the host->device copy are skipped.  Used by callers that sweep k-ranges over one matrix
 * (host-staged blocks, repeated partial solves) and by bench.py's device-resident `value`.  Any other compute call on
 * the context (batched solves, ah_symarrow_eigvals_dd) invalidates the resident inputs. */
int ah_resolve(ah_ctx *ctx, int64_t k0, int64_t k1, ah_result *out);

/* ---- several GPUs ------------------------------------------------------------------------------
 * The reference parallelises the eigenpair loop inside one process (`Threads.@threads for k`, arrowhead3.jl:641-644,
 * arrowhead4.jl:531-535).  ah_multi does the same over GPUs for a single-threaded caller (Julia's ccall): one context
 * per device, one worker thread per context inside the library, the k-range 1..kmax cut into contiguous blocks
 * (ah_k_partition).  Every device receives D and z and owns its eigenpairs and its columns of U; nothing is exchanged.
 *   out->lambda / sind / ... : host arrays of FULL length kmax; device r fills its block.
 *   eigenvectors, host   : out->U (out->U_mem = AH_MEM_HOST) is the full host matrix; every device streams its columns
 *                          into it (ld = out->ldu).  Pass U_full = NULL.
 *   eigenvectors, device : U_full[r] is a full-size buffer (ldu x kmax doubles) on device r; device r writes its block
 *                          at column offset k0-1, so that ah_multi_allgather_columns can complete every copy in place.
 * Mapped stores (rowmap_u/rowmap_v: deflated problems) are not sharded.  Calls block until every device has finished;
 * ah_multi_last_error gives the first failure; per-device timing: ah_get_stats(ah_multi_ctx(m, r), ...). */
typedef struct ah_multi ah_multi;
int  ah_multi_create(ah_multi **m, const int *devices, int ndev);   /* devices NULL: 0..ndev-1; ndev <= 0: all visible GPUs */
void ah_multi_destroy(ah_multi *m);
int  ah_multi_ndev(const ah_multi *m);
ah_ctx *ah_multi_ctx(ah_multi *m, int r);                           /* borrowed; r = 0..ndev-1 */
const char *ah_multi_last_error(const ah_multi *m);
double ah_multi_elapsed_ms(const ah_multi *m);                      /* host wall time of the last ah_multi_* solve */
/* block r of nparts: the 1-based inclusive range [*k0,*k1] (sizes differ by at most one; empty: k0=1, k1=0) */
void ah_k_partition(int64_t nk, int nparts, int r, int64_t *k0, int64_t *k1);
int ah_multi_symarrow_eigen(ah_multi *m, int64_t N, const double *D, const double *z, double a, const double *tau,
                            ah_result *out, double *const *U_full);
int ah_multi_symdpr1_eigen(ah_multi *m, int64_t n, const double *D, const double *u, double rho, const double *tau,
                           ah_result *out, double *const *U_full);
int ah_multi_halfarrow_svd(ah_multi *m, int64_t nd, int64_t nz, const double *D, const double *z, const double *tau,
                           ah_result *out, double *const *U_full, double *const *V_full);
/* In-place all-gather over NVLink (ncclAllGather; ncclBroadcast per block when kmax is not a multiple of ndev): after
 * the call every U_full[r] (ld x kmax, column-major, on device r) holds all columns.  *elapsed_ms: device time, max
 * over the devices.  NCCL is dlopen()ed on first use (libnccl.so.2; env ARROWHEAD_B200_NCCL overrides the path). */
int ah_multi_allgather_columns(ah_multi *m, double *const *U_full, int64_t ld, int64_t kmax, double *elapsed_ms);
/* The same for one process per GPU (torchrun / MPI style): rank 0 calls ah_nccl_unique_id, the caller carries the
 * 128 bytes to the other ranks, everyone calls ah_nccl_init once, then ah_allgather_columns on its own full-size
 * buffer whose block [ah_k_partition(kmax, nranks, rank)] it has filled. */
int ah_nccl_unique_id(void *id128);
int ah_nccl_init(ah_ctx *ctx, int nranks, int rank, const void *id128);
int ah_allgather_columns(ah_ctx *ctx, double *U_full, int64_t ld, int64_t kmax, double *elapsed_ms);

/* ---- device utilities for a resident U -------------------------------------------------------- */
int ah_device_malloc(ah_ctx *ctx, void **dptr, int64_t bytes);
int ah_device_free(ah_ctx *ctx, void *dptr);
int ah_host_alloc_pinned(ah_ctx *ctx, void **hptr, int64_t bytes);
int ah_host_free_pinned(ah_ctx *ctx, void *hptr);
int ah_memcpy_d2h(ah_ctx *ctx, void *dst_host, const void *src_dev, int64_t bytes);
int ah_memcpy_h2d(ah_ctx *ctx, void *dst_dev, const void *src_host, int64_t bytes);
/* U (device, n x ncols, leading dimension ld) := identity pattern (ones on the main diagonal). */
int ah_set_identity(ah_ctx *ctx, double *U_dev, int64_t nrows, int64_t ncols, int64_t ld);
/* rows i1,i2 (0-based) of U := [c s; -s c] * rows, for all ncols columns (lmul!(Givens(i1,i2,c,s),U)). */
int ah_apply_givens_rows(ah_ctx *ctx, double *U_dev, int64_t ld, int64_t ncols, int64_t i1, int64_t i2,
                         double c, double s);
/* dst[r,c] = src[rows[r], cols[c]] (0-based host index vectors), both device, column-major. */
int ah_gather_permuted(ah_ctx *ctx, double *dst_dev, int64_t ld_dst, const double *src_dev, int64_t ld_src,
                       const int64_t *rows, int64_t nrows, const int64_t *cols, int64_t ncols);

/* rootsah (src/arrowhead7.jl:53-120): eigenvalues only of the arrowhead companion matrix whose border z and arrow top
 * are known in double-double (z_hi+z_lo, a_hi+a_lo; built on the host, arrowhead7.jl:66-110).  Replaces the serial loop
 * `for k=1:n; E[k],Qout[k] = eigen(A,zD,alphaD,k,tau); end` (:116-118) and the per-k solver :351-495 with one launch.
 * Same input contract as ah_symarrow_eigen; tau is required (rootsah's default is [1e2,1e2,1e2,1e2,1e2]);
 * out->U and out->V must be NULL. */
int ah_symarrow_eigvals_dd(ah_ctx *ctx, int64_t N, const double *D, const double *z_hi, const double *z_lo, double a_hi,
                           double a_lo, const double *tau, int64_t k0, int64_t k1, ah_result *out);

/* ---- batched primitives for tdc() (src/arrowhead7.jl:10-36): one divide-and-conquer level = many independent
 * arrow merges of (nearly) equal size, then two back-transform GEMMs per merge ---------------------------------- */
/* nb ordered irreducible SymArrow problems with the same number of poles N, solved in ONE launch by the complete
 * per-k solver (every branch of arrowhead3.jl:419-550).  D, z: host [nb*N] (problem b at offset b*N), a: host [nb].
 * out->lambda/sind/...: host arrays of nb*(N+1) entries (problem b at offset b*(N+1)).  out->U: NULL or a DEVICE
 * pointer; problem b's (N+1)x(N+1) eigenvector matrix goes to U + b*ustride with leading dimension out->ldu. */
int ah_symarrow_eigen_batched(ah_ctx *ctx, int64_t nb, int64_t N, const double *D, const double *z, const double *a,
                              const double *tau, ah_result *out, int64_t ustride);
/* out[t] = src_dev[idx[t]] (element offsets, host index vector, host result): the merge vectors
 * z = [U1[end,:]; U2[1,:]] of arrowhead7.jl:28 are read straight out of the resident eigenvector blocks. */
int ah_gather_elements(ah_ctx *ctx, const double *src_dev, const int64_t *idx, int64_t count, double *out);
/* nb gathers of equal shape: dst_b[r,c] = src_b[rows[b*nrows+r], cols[b*ncols+c]], dst_b = dst + b*dst_stride,
 * src_b = src + b*src_stride (all device, column-major); rows/cols are host index vectors. */
int ah_gather_permuted_batched(ah_ctx *ctx, int64_t nb, double *dst_dev, int64_t dst_stride, int64_t ld_dst,
                               const double *src_dev, int64_t src_stride, int64_t ld_src, const int64_t *rows,
                               int64_t nrows, const int64_t *cols, int64_t ncols);
/* nb strided 2-D copies of equal shape between device matrices given by their (device) addresses:
 * dst_ptrs[b][c*ld_dst + r] = src_ptrs[b][c*ld_src + r]. */
int ah_copy2d_batched(ah_ctx *ctx, int64_t nb, int64_t nrows, int64_t ncols, const uint64_t *dst_ptrs, int64_t ld_dst,
                      const uint64_t *src_ptrs, int64_t ld_src);
/* C_b = A_b * B_b (m x k times k x n), all device, column-major, addresses in host arrays: the back-transform
 * `E.vectors[1:k,:] = E1.vectors*E.vectors[1:k,:]` (arrowhead7.jl:32-33) of a whole tree height in one launch of the
 * library's own FP64 tensor-core kernel (csrc/ah_gemm.cuh: DMMA.8x8x4 = mma.sync.m8n8k4.f64, cp.async ring, 128x128 or
 * 64x64 CTA tiles).  No cuBLAS anywhere in the library.  nb <= 65535. */
int ah_dgemm_batched(ah_ctx *ctx, int64_t nb, int64_t m, int64_t n, int64_t k, const uint64_t *A_ptrs, int64_t lda,
                     const uint64_t *B_ptrs, int64_t ldb, const uint64_t *C_ptrs, int64_t ldc);

/* Hermitian wrappers (src/hermitian.jl:107-115, 131-139: `Diagonal(c)*U.vectors`):
 * dst[(c*ld_dst + r)*ncomp + q] = phase[r*ncomp + q] * src[c*ld_src + r], q < ncomp (2: complex, 4: quaternion).
 * src is the real eigenvector matrix on the device, dst a device buffer of nrows*ncols*ncomp doubles
 * (interleaved components = ComplexF64 / QuaternionF64 layout), phase a host array. */
int ah_scale_rows_phase(ah_ctx *ctx, double *dst_dev, int64_t ld_dst, const double *src_dev, int64_t ld_src,
                        const double *phase, int ncomp, int64_t nrows, int64_t ncols);
/* The same fused with the driver's final permutation `view(U,rows,es)` (arrowhead3.jl:648-654), so that the Hermitian
 * wrappers cost no extra pass over the eigenvectors:
 * dst[(c*ld_dst + r)*ncomp + q] = phase[r*ncomp + q] * src[cols[c]*ld_src + rows[r]]  (rows, cols: 0-based host index vectors). */
int ah_gather_permuted_phase(ah_ctx *ctx, double *dst_dev, int64_t ld_dst, const double *src_dev, int64_t ld_src, const int64_t *rows,
                             int64_t nrows, const int64_t *cols, int64_t ncols, const double *phase, int ncomp);

/* The device utilities above (set_identity, apply_givens_rows, gather_permuted[_batched], copy2d_batched,
 * dgemm_batched, scale_rows_phase) synchronise the stream before they return.  ah_set_async(ctx, 1) turns them into
 * enqueue-only calls (their host index / pointer arrays are staged through pinned memory, so they may be freed on
 * return); ah_wait synchronises and reports any deferred CUDA error.  tdc() uses this to keep a whole tree height on
 * the stream.  Calls that return data to the host (the solves, ah_gather_elements, ah_memcpy_d2h) always synchronise. */
int ah_set_async(ah_ctx *ctx, int on);
int ah_wait(ah_ctx *ctx);
/* device time of the last ah_dgemm_batched launch (CUDA events around the kernel) */
int ah_last_gemm_ms(ah_ctx *ctx, double *ms);

/* ---- measurement helpers (used by bench.py for the roofline denominators) ---------------------- */
/* DMMA.8x8x4 burst from registers on every SM: FP64 tensor-core rate in 1e9 FMA/s (x2 = GFLOP/s): the denominator of
 * the back-transform GEMM's roofline. */
int ah_measure_dmma_peak(ah_ctx *ctx, double *gfma_per_s, double *elapsed_ms);
/* Dependent-free DFMA burst on every SM: returns FP64-pipe instruction issue rate in
 * 1e9 warp-lane-instructions/s (x2 = GFLOP/s) and the elapsed ms. */
int ah_measure_fp64_peak(ah_ctx *ctx, double *ginstr_per_s, double *elapsed_ms);
/* The same for the instruction mix of the bisection term (7 FP64-pipe instructions + 1 MUFU.RCP64H per secular term,
 * 32 independent terms per thread and iteration, registers only): 1e9 terms/s.  The ceiling of k_bisect's inner block. */
int ah_measure_term_peak(ah_ctx *ctx, double *gterms_per_s, double *elapsed_ms);
/* Exhaustive accuracy of the reciprocal seed the secular term starts from (MUFU.RCP64H: all 2^20 high-word mantissa
 * patterns x low word 0 / all ones x both signs x several exponents): max |1 - x*r0|, and the same for the corrected
 * reciprocal of the term.  The certificates of the accelerated bisection assume seed_err <= 2^-18. */
int ah_measure_rcp_seed_error(ah_ctx *ctx, double *seed_err, double *corrected_err);
/* streaming store of `bytes` to device memory: GB/s */
int ah_measure_store_bw(ah_ctx *ctx, int64_t bytes, double *gb_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ARROWHEAD_B200_H */
