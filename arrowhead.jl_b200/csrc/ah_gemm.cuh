// ah_gemm.cuh -- the FP64 back-transform of tdc() as a hand-written sm_100a kernel.
//
//   E.vectors[1:k,:]   = E1.vectors * E.vectors[1:k,:]         (arrowhead7.jl:32)
//   E.vectors[k+2:n,:] = E2.vectors * E.vectors[k+2:n,:]       (arrowhead7.jl:33)
//
// i.e. for every merge of one tree height C_b (m x n) = A_b (m x k) * B_b (k x n), column-major, addresses in
// pointer arrays (all merges of a height have the same shape).  This is the one real contraction of the package;
// FP64 has no tcgen05/TMEM path, the tensor-core instruction for binary64 on sm_100a is DMMA.8x8x4
// (`mma.sync.aligned.m8n8k4.row.col.f64`), which is what this kernel issues.
//
// Tiling: one CTA = TM x TN tile of C, K in slices of 16 through a 3-stage cp.async ring in shared memory;
// WGM x WGN warps, each a (TM/WGM) x (TN/WGN) block of 8x8 DMMA tiles with the accumulators in registers
// (128 x 64 with 8 warps of 64 x 16, two CTAs per SM; 64 x 64 with 8 warps of 32 x 16 for small merges).
// Shared-memory layouts are padded so that every fragment load (one double per lane) is conflict-free:
//   A tile  [kk][row]  column stride TM+4 doubles: the 4 k-columns of a fragment fall 8 banks apart
//   B tile  [col][kk]  column stride 20 doubles  : the 4 columns of a half-warp fall 8 banks apart
// Global loads are 8-byte cp.async (sub-blocks of the resident eigenvector matrices start at odd element offsets,
// so 16-byte alignment cannot be assumed); rows/columns/k beyond the matrix edge are zero-filled by cp.async's
// src-size operand, the epilogue is predicated.  Rounding: each C entry is a chain of fused multiply-adds over k in
// ascending order (DMMA accumulates 4 products per instruction in order) -- one rounding per product-sum, as cuBLAS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ah {

#ifndef AH_GEMM_GK
#define AH_GEMM_GK 16
#endif
#ifndef AH_GEMM_STAGES
#define AH_GEMM_STAGES 3
#endif
constexpr int GK = AH_GEMM_GK;            // k-slice held by one stage of the ring
constexpr int GSTAGES = AH_GEMM_STAGES;

template <int TM, int TN>
struct GemmSmem {
    static constexpr int LDA = TM + 4;     // doubles per k-column of the A tile
    static constexpr int LDB = GK + 4;     // doubles per column of the B tile
    double A[GSTAGES][GK * LDA];
    double B[GSTAGES][TN * LDB];
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 8 : 0;       // src-size 0: the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// grid: (ceil(n/TN), ceil(m/TM), nb); WGM x WGN warps, each a (TM/WGM) x (TN/WGN) block of C
template <int TM, int TN, int WGM, int WGN>
__global__ void __launch_bounds__(32 * WGM * WGN) k_dgemm_dmma(const uint64_t *__restrict__ Aptrs, const uint64_t *__restrict__ Bptrs,
                                                                const uint64_t *__restrict__ Cptrs, int m, int n, int k, int64_t lda,
                                                                int64_t ldb, int64_t ldc) {
    extern __shared__ __align__(16) unsigned char gsm_raw[];
    using S = GemmSmem<TM, TN>;
    S &sm = *reinterpret_cast<S *>(gsm_raw);
    constexpr int NTHR = 32 * WGM * WGN;
    constexpr int WM = TM / WGM, WN = TN / WGN;      // warp tile
    constexpr int MT = WM / 8, NT = WN / 8;          // 8x8 DMMA tiles per warp
    constexpr int AL = (GK * TM) / NTHR, BL = (GK * TN) / NTHR;   // cp.async per thread and slice (A, B)
    constexpr int AKS = NTHR / TM;                   // k-columns of the A tile covered by one pass of the CTA
    constexpr int BCS = NTHR / GK;                   // columns of the B tile covered by one pass
    static_assert(AL >= 1 && BL >= 1 && AL * NTHR == GK * TM && BL * NTHR == GK * TN, "tile / thread-count mismatch");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % WGM) * WM, wn = (warp / WGM) * WN;
    const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    const double *__restrict__ A = reinterpret_cast<const double *>(Aptrs[blockIdx.z]);
    const double *__restrict__ B = reinterpret_cast<const double *>(Bptrs[blockIdx.z]);
    double *__restrict__ C = reinterpret_cast<double *>(Cptrs[blockIdx.z]);

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // ---- per-thread load plan, fixed for the whole K loop: the address arithmetic of a slice is one add per copy.
    // A tile: thread -> row ra, k-columns ka0 + it*AKS.   B tile: thread -> k-row kb, columns cb0 + it*BCS.
    const int ra = tid % TM, ka0 = tid / TM;
    const int kb = tid % GK, cb0 = tid / GK;
    const bool a_row_ok = row0 + ra < m;
    const double *pa = A + (int64_t)ka0 * lda + (a_row_ok ? row0 + ra : 0);
    const double *pb = B + (int64_t)(col0 + cb0) * ldb + kb;
    unsigned b_col_ok = 0;
#pragma unroll
    for (int it = 0; it < BL; ++it) b_col_ok |= (col0 + cb0 + it * BCS < n ? 1u : 0u) << it;
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&sm.A[0][ka0 * S::LDA + ra]);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&sm.B[0][cb0 * S::LDB + kb]);
    constexpr uint32_t A_STAGE = GK * S::LDA * 8, B_STAGE = TN * S::LDB * 8;

    const int nslices = (k + GK - 1) / GK;
    auto load_slice = [&](int s, int stage) {
        const int k0 = s * GK;
        const double *ga = pa + (int64_t)k0 * lda;
        const double *gb = pb + k0;
#pragma unroll
        for (int it = 0; it < AL; ++it) {
            const bool ok = a_row_ok && (k0 + ka0 + it * AKS < k);
            const int bytes = ok ? 8 : 0;                        // src-size 0: the destination is zero-filled
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa + stage * A_STAGE + it * AKS * S::LDA * 8),
                         "l"(ok ? ga + (int64_t)it * AKS * lda : A), "r"(bytes) : "memory");
        }
#pragma unroll
        for (int it = 0; it < BL; ++it) {
            const bool ok = ((b_col_ok >> it) & 1u) && (k0 + kb < k);
            const int bytes = ok ? 8 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sb + stage * B_STAGE + it * BCS * S::LDB * 8),
                         "l"(ok ? gb + (int64_t)it * BCS * ldb : B), "r"(bytes) : "memory");
        }
    };

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < nslices) load_slice(s, s);
        cp_async_commit();
    }
    const int ar = lane >> 2, ak = lane & 3;         // fragment coordinates: A[row ar][k ak], B[k ak][col ar]
    int stage = 0, nstage = GSTAGES - 1;             // stage of slice s, stage that slice s + GSTAGES - 1 goes into
    for (int s = 0; s < nslices; ++s) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();                             // slice s has landed for everyone; the stage of slice s-1 is free again
        if (s + GSTAGES - 1 < nslices) load_slice(s + GSTAGES - 1, nstage);
        cp_async_commit();
        const double *As = sm.A[stage], *Bs = sm.B[stage];
#pragma unroll
        for (int k4 = 0; k4 < GK; k4 += 4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) af[i] = As[(k4 + ak) * S::LDA + wm + i * 8 + ar];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = Bs[(wn + j * 8 + ar) * S::LDB + k4 + ak];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        stage = (stage + 1 == GSTAGES) ? 0 : stage + 1;
        nstage = (nstage + 1 == GSTAGES) ? 0 : nstage + 1;
    }
    cp_async_wait<0>();
    // epilogue: C fragment of a DMMA tile: row ar, columns 2*ak and 2*ak+1
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int r = row0 + wm + i * 8 + ar;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c = col0 + wn + j * 8 + 2 * ak;
            if (r < m) {
                if (c < n) C[(int64_t)c * ldc + r] = acc[i][j][0];
                if (c + 1 < n) C[(int64_t)(c + 1) * ldc + r] = acc[i][j][1];
            }
        }
    }
}

template <int TM, int TN, int WGM, int WGN>
inline int launch_dgemm_tile(const uint64_t *A, const uint64_t *B, const uint64_t *C, int nb, int m, int n, int k, int64_t lda,
                             int64_t ldb, int64_t ldc, cudaStream_t st) {
    const int smem = (int)sizeof(GemmSmem<TM, TN>);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_dgemm_dmma<TM, TN, WGM, WGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    dim3 grid((n + TN - 1) / TN, (m + TM - 1) / TM, nb);
    k_dgemm_dmma<TM, TN, WGM, WGN><<<grid, 32 * WGM * WGN, smem, st>>>(A, B, C, m, n, k, lda, ldb, ldc);
    return (int)cudaGetLastError();
}

#ifndef AH_GEMM_VARIANT
#define AH_GEMM_VARIANT 2
#endif
// Tile choice (measured on B200 on the back-transform shapes of tdc n = 8192, tools/gemm_bench.py): 128 x 64 tiles with
// 8 warps of 64 x 16 and TWO CTAs per SM (one computes through the other's slice barrier) reach 30.4 TFLOP/s at
// 4096 x 8192 x 4096; 128 x 128 tiles with one CTA per SM stay at 28 TFLOP/s whether 8 warps of 64 x 32 or 16 of 32 x 32.
// Small merges of the low tree levels use 64 x 64 tiles.
inline int launch_dgemm(const uint64_t *A, const uint64_t *B, const uint64_t *C, int nb, int m, int n, int k, int64_t lda,
                        int64_t ldb, int64_t ldc, int sm_count, cudaStream_t st) {
    const int64_t big = (int64_t)((n + 63) / 64) * ((m + 127) / 128) * nb;
    if (m >= 96 && n >= 48 && big >= sm_count) {
#if AH_GEMM_VARIANT == 1
        return launch_dgemm_tile<128, 128, 2, 4>(A, B, C, nb, m, n, k, lda, ldb, ldc, st);     // 8 warps of 64 x 32
#elif AH_GEMM_VARIANT == 0
        return launch_dgemm_tile<128, 128, 4, 4>(A, B, C, nb, m, n, k, lda, ldb, ldc, st);     // 16 warps of 32 x 32
#else
        return launch_dgemm_tile<128, 64, 2, 4>(A, B, C, nb, m, n, k, lda, ldb, ldc, st);      // 8 warps of 64 x 16, two CTAs per SM
#endif
    }
    return launch_dgemm_tile<64, 64, 2, 4>(A, B, C, nb, m, n, k, lda, ldb, ldc, st);
}

// DMMA.8x8x4 burst: independent accumulator chains from registers -> tensor-FP64 FMA rate (the GEMM's denominator)
__global__ void __launch_bounds__(256) k_dmma_burst(double *out, int iters, double a0, double b0) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = 1.0 - i * 1e-3; }
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += c[i][0] + c[i][1];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}

}  // namespace ah
