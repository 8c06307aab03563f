// ah_api.cu -- host side of libarrowhead_b200.so: the C ABI declared in include/arrowhead_b200.h.
//
// One ah_ctx = one GPU: a stream, reusable device buffers for the (small) inputs and the per-k outputs,
// and a double-buffered staging area for eigenvector blocks when the caller wants them on the host.
// There is no CPU fallback anywhere in this file: every compute entry point launches the kernels in
// ah_fast.cuh / ah_full.cuh or fails.
#include "../../include/arrowhead_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ah_device.cuh"
#include "ah_full.cuh"
#include "ah_pipe.cuh"
#include "ah_roots.cuh"
#include "ah_gemm.cuh"

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct ah_ctx {
    int device = 0;
    int sm_count = 0;
    int quantum = -1;           // k_bisect time slice (steps): < 0 automatic, 0 off (tuning knob ARROWHEAD_B200_QUANTUM)
    int window = 2;             // time slicing starts when fewer than window x slots eigenpairs wait (ARROWHEAD_B200_WINDOW)
    int spec = 1;               // speculative tail of k_bisect: idle slots evaluate the next midpoints (ARROWHEAD_B200_SPEC=0: off)
    int spinwait = 0;           // ARROWHEAD_B200_SPINWAIT=1: poll the stream at the end of a solve instead of blocking in the driver
    int accel = 1;              // certified skipping of bisection steps in k_bisect round 0 (ARROWHEAD_B200_ACCEL=0: plain bisection)
    unsigned bisect_tag = 0;    // per-launch tag of the requeue FIFO entries
    bool res_valid = false;     // dD / dZ still hold the inputs of the last solve (ah_resolve)
    int res_kind = 0;
    int64_t res_N = 0, res_nvec_u = 0, res_nvec_v = 0, res_kmax = 0;
    ah::Problem res_P{};
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t side_stream = nullptr;   // high priority: the short Remedy-3 round runs beside k_vec
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_early0 = nullptr, ev_early1 = nullptr;   // around the early k_full launch on the side stream
    std::vector<cudaEvent_t> side_ev;     // 4 timing events per column block on the side stream (Remedy-3 round, k_full)
    cudaStream_t stream = nullptr;  // the stream kernels run on (own_stream or caller's)
    cudaEvent_t ev[12] = {};
    cudaEvent_t ev_stage[2] = {};
    std::string err;
    int mode = 0;
    ah_stats stats = {};
    DevBuf dD, dW, dW2, dKs, dRList, dZ, dMap1, dMap2, dScale, dOut, dList, dCount, dQueue, dStage[2];
    void *hPinned = nullptr;  // pinned bounce buffer for the per-k outputs
    size_t hPinnedCap = 0;
    double *hIn = nullptr;      // pinned staging of the padded inputs D | w | w^2
    size_t hInCap = 0;
    bool async_utils = false;            // ah_set_async: the device utilities enqueue and return (ah_wait synchronises)
    char *hArena = nullptr;              // pinned staging for the small host arrays of the utilities (pointer / index vectors)
    size_t hArenaCap = 0, hArenaUsed = 0;
    DevBuf dPtrA, dPtrB, dPtrC, dIdx, dBatchA;
    std::vector<cudaEvent_t> chunk_ev;   // 5 timing events per column block of the current call
    unsigned long long *hCounts = nullptr;  // pinned, 2 per column block: hand-over / Remedy-3 counts of the last call
    size_t hCountsCap = 0;
    void *nccl_comm = nullptr;              // ncclComm_t of ah_nccl_init (one process per GPU); see ah_multi.inl
    int nccl_rank = 0, nccl_nranks = 0;
};

namespace {

int fail(ah_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(ah_ctx *c, cudaError_t e, const char *what) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    cudaGetLastError();
    return fail(c, AH_ERR_CUDA_BASE + (int)e, m);
}
#define AH_CUDA(call)                                          \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

int ensure(ah_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return AH_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = std::max(bytes, (size_t)256);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMalloc failed for " + std::to_string(want) + " bytes"); }
    b.cap = want;
    return AH_OK;
}
int ensure_pinned(ah_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->hPinnedCap) return AH_OK;
    if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
    ctx->hPinned = nullptr; ctx->hPinnedCap = 0;
    cudaError_t e = cudaMallocHost(&ctx->hPinned, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMallocHost failed"); }
    ctx->hPinnedCap = bytes;
    return AH_OK;
}

// Host arrays of the device utilities go through a pinned arena so that the copy is truly asynchronous (a
// cudaMemcpyAsync from pageable memory synchronises the stream first).  Slots are recycled at every synchronisation.
int h2d_staged(ah_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (need > ctx->hArenaCap - ctx->hArenaUsed) {
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_CUDA_BASE, "stream synchronize failed"); }
        ctx->hArenaUsed = 0;
        if (need > ctx->hArenaCap) {
            if (ctx->hArena) cudaFreeHost(ctx->hArena);
            ctx->hArena = nullptr; ctx->hArenaCap = 0;
            const size_t cap = std::max<size_t>(need, (size_t)8 << 20);
            if (cudaMallocHost((void **)&ctx->hArena, cap) != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMallocHost failed"); }
            ctx->hArenaCap = cap;
        }
    }
    char *slot = ctx->hArena + ctx->hArenaUsed;
    ctx->hArenaUsed += need;
    std::memcpy(slot, src_host, bytes);
    cudaError_t e = cudaMemcpyAsync(dst_dev, slot, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemcpyAsync (staged host array)");
    return AH_OK;
}
// end of a device utility: synchronise unless the caller asked for enqueue-only behaviour (ah_set_async)
int util_done(ah_ctx *ctx) {
    if (ctx->async_utils) return AH_OK;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    ctx->hArenaUsed = 0;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
    return AH_OK;
}

// ---- host-side copies of the few scalar reductions the reference does once per problem -------------
double jl_pairwise_abs_sum(const double *x, int64_t lo, int64_t hi) {  // Julia sum(abs, x)
    if (hi < lo) return 0.0;
    if (lo == hi) return std::fabs(x[lo]);
    if (hi - lo < 1024) {
        double v = std::fabs(x[lo]) + std::fabs(x[lo + 1]);
        for (int64_t i = lo + 2; i <= hi; ++i) v += std::fabs(x[i]);
        return v;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    return jl_pairwise_abs_sum(x, lo, mid) + jl_pairwise_abs_sum(x, mid + 1, hi);
}
struct hdd { double hi, lo; };
hdd h_renorm(double u, double v) { volatile double w = u + v; volatile double t = u - w; return {w, t + v}; }
hdd h_mul12(double x, double y) {  // Dekker/Veltkamp, DoubleDouble.jl:34-42,137-142
    volatile double px = x * 1.34217729e8, hx = (x - px) + px, lx = x - hx;
    volatile double py = y * 1.34217729e8, hy = (y - py) + py, ly = y - hy;
    volatile double z = x * y;
    volatile double e = (((hx * hy - z) + hx * ly) + lx * hy) + lx * ly;
    return {z, e};
}
hdd h_add(hdd x, hdd y) {  // DoubleDouble.jl:112-116
    volatile double r = x.hi + y.hi;
    volatile double s = std::fabs(x.hi) > std::fabs(y.hi) ? (((x.hi - r) + y.hi) + y.lo) + x.lo
                                                         : (((y.hi - r) + x.hi) + x.lo) + y.lo;
    return h_renorm(r, s);
}

}  // namespace
void ah_release_nccl(ah_ctx *ctx);   // ah_multi.inl
namespace {

struct Timer {
    cudaEvent_t a, b;
};

template <typename T>
T *carve(char *&p, size_t count) {
    T *r = reinterpret_cast<T *>(p);
    p += ((count * sizeof(T) + 255) / 256) * 256;
    return r;
}
size_t out_bytes(int64_t cnt) {
    auto al = [](size_t b) { return ((b + 255) / 256) * 256; };
    return al(cnt * 8) * 7 + al(cnt * 4) * 2;
}

struct Launch {
    int kind;
    int64_t N;       // poles
    int64_t nvec_u;  // rows of U vectors
    int64_t nvec_v;  // rows of V vectors (half)
    int64_t kmax;
};

int validate_common(ah_ctx *ctx, int64_t N, const double *D, const double *w, int64_t nw, int64_t k0, int64_t k1,
                    int64_t kmax, const ah_result *out, bool positiveD) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    if (!D || !w || !out || !out->lambda) return fail(ctx, AH_ERR_INVALID_ARG, "null argument");
    if (N < 1) return fail(ctx, AH_ERR_INVALID_ARG, "need at least one pole");
    if (N > (int64_t)1 << 30) return fail(ctx, AH_ERR_INVALID_ARG, "more than 2^30 poles");
    if (k0 < 1 || k1 < k0 || k1 > kmax) return fail(ctx, AH_ERR_INVALID_ARG, "k-range outside 1.." + std::to_string(kmax));
    // branch-free passes (they vectorise; this runs on every drop-in call); which test failed is looked up afterwards
    int bad = 0;
    for (int64_t l = 0; l < N; ++l) bad |= !(std::fabs(D[l]) <= 1.7976931348623157e308);
    for (int64_t l = 1; l < N; ++l) bad |= !(D[l - 1] > D[l]);
    if (positiveD) bad |= !(D[N - 1] > 0.0);
    for (int64_t l = 0; l < nw; ++l) bad |= (w[l] == 0.0) | !(std::fabs(w[l]) <= 1.7976931348623157e308);
    if (!bad) return AH_OK;
    for (int64_t l = 0; l < N; ++l) {
        if (!std::isfinite(D[l])) return fail(ctx, AH_ERR_NOT_ORDERED, "non-finite pole");
        if (l && !(D[l - 1] > D[l])) return fail(ctx, AH_ERR_NOT_ORDERED, "D must be strictly decreasing (sort and deflate first)");
        if (positiveD && !(D[l] > 0.0)) return fail(ctx, AH_ERR_NOT_ORDERED, "HalfArrow needs D > 0 (pass abs(D))");
    }
    for (int64_t l = 0; l < nw; ++l)
        if (w[l] == 0.0 || !std::isfinite(w[l])) return fail(ctx, AH_ERR_REDUCIBLE, "zero or non-finite entry in z/u (deflate first)");
    return fail(ctx, AH_ERR_INVALID_ARG, "invalid input");
}

// End of a solve: optionally poll the stream instead of blocking in the driver (ARROWHEAD_B200_SPINWAIT=1).  Measured on
// 8 x B200 with one process per GPU: no difference (2.38 ms per step either way), so blocking is the default.
cudaError_t wait_stream(ah_ctx *ctx, cudaStream_t st) {
    if (ctx->spinwait) {
        const auto t0 = std::chrono::steady_clock::now();
        for (unsigned it = 0;; ++it) {
            const cudaError_t e = cudaStreamQuery(st);
            if (e != cudaErrorNotReady) return e;
            if ((it & 1023u) == 1023u && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(200)) break;   // long call: stop burning a core
        }
    }
    return cudaStreamSynchronize(st);
}

int run_solver(ah_ctx *ctx, const Launch &L, ah::Problem P, const double *hD, const double *hW, const double *hZ,
               int64_t k0, int64_t k1, ah_result *out) {
    using namespace ah;
    const int64_t N = L.N, cnt = k1 - k0 + 1;
    ctx->stats = ah_stats{};
    const auto t_enter = std::chrono::steady_clock::now();
    AH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if ((out->rowmap_u || out->rowmap_v) && ((out->U && out->U_mem != AH_MEM_DEVICE) || (out->V && out->V_mem != AH_MEM_DEVICE)))
        return fail(ctx, AH_ERR_INVALID_ARG, "rowmap needs device-resident U/V (ah_set_identity + AH_MEM_DEVICE)");
    if (out->U && out->ldu < (out->rowmap_u ? 1 : L.nvec_u)) return fail(ctx, AH_ERR_INVALID_ARG, "ldu too small");
    if (L.kind == KIND_HALF && out->V && out->ldv < (out->rowmap_v ? 1 : L.nvec_v)) return fail(ctx, AH_ERR_INVALID_ARG, "ldv too small");
    // Mapped stores (`U[zx,zx[k]] = v`): entry r of vector k goes to row rowmap[r] of COLUMN rowmap_u[k-1].  The column
    // index of V is taken from rowmap_u as well (`V[zxv,zx[k]] = v`, arrowhead6.jl:640), so a HalfArrow call maps
    // both factors or neither; every index must lie inside the leading dimension.
    if (L.kind == KIND_HALF && out->V && out->U && ((out->rowmap_u != nullptr) != (out->rowmap_v != nullptr)))
        return fail(ctx, AH_ERR_INVALID_ARG, "HalfArrow: rowmap_u and rowmap_v must be given together");
    if (L.kind == KIND_HALF && out->rowmap_v && !out->rowmap_u) return fail(ctx, AH_ERR_INVALID_ARG, "HalfArrow: rowmap_v needs rowmap_u (it supplies the column index)");
    if (out->rowmap_u && k1 > L.nvec_u) return fail(ctx, AH_ERR_INVALID_ARG, "rowmap_u has no entry for every k of the range");
    if (out->rowmap_u && out->U)
        for (int64_t r = 0; r < L.nvec_u; ++r)
            if (out->rowmap_u[r] < 0 || out->rowmap_u[r] >= out->ldu) return fail(ctx, AH_ERR_INVALID_ARG, "rowmap_u entry outside [0, ldu)");
    if (L.kind == KIND_HALF && out->rowmap_v && out->V)
        for (int64_t r = 0; r < L.nvec_v; ++r)
            if (out->rowmap_v[r] < 0 || out->rowmap_v[r] >= out->ldv) return fail(ctx, AH_ERR_INVALID_ARG, "rowmap_v entry outside [0, ldv)");

    int rc;
    if (hD) {
        // ---- inputs -> device, padded to whole tiles of the fast kernel's ring.  Pad poles lie strictly
        // below every real pole (so no shift or bisection point can hit them) and carry weight 0.
        const int64_t Np = ((N + ah::TILE_POLES - 1) / ah::TILE_POLES) * ah::TILE_POLES;
        // D | w | w^2 go through one pinned staging block and one copy (pageable sources are staged by the driver in
        // small pieces: 3 x 256 KB took 50 us + 100 us of host time at n = 32768)
        if ((size_t)(3 * Np * 8) > ctx->hInCap) {
            if (ctx->hIn) cudaFreeHost(ctx->hIn);
            ctx->hIn = nullptr; ctx->hInCap = 0;
            if (cudaMallocHost((void **)&ctx->hIn, (size_t)(3 * Np * 8)) != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMallocHost failed"); }
            ctx->hInCap = (size_t)(3 * Np * 8);
        }
        double *hPadD = ctx->hIn, *hPadW = ctx->hIn + Np, *hPadW2 = ctx->hIn + 2 * Np;
        std::memcpy(hPadD, hD, N * 8);
        std::memcpy(hPadW, hW, N * 8);
        for (int64_t l = 0; l < N; ++l) hPadW2[l] = hW[l] * hW[l];   // one rounding, as z[k]^2 in the reference
        double pad;
        if (L.kind == KIND_HALF) {
            pad = hD[N - 1] / 2.0;
        } else {
            double span = hD[0] - hD[N - 1];
            if (!(span > 0.0)) span = std::max(std::fabs(hD[N - 1]), 1.0);
            pad = hD[N - 1] - span;
            if (!std::isfinite(pad)) pad = -1.7976931348623157e308;
        }
        // A pad's term is w2 * (1/t) with w2 = 0: it must never see t = d (1 - m d) = 0 (0 * Inf = NaN).  1 - m d vanishes
        // exactly only if m = 1/d is representable, i.e. only if d = dshift(pad, sigma) is a power of two for some shift
        // sigma in D -- nudge the pad until it is not (checked against every pole).
        auto pad_shift = [&](double pd, double sg) { return L.kind == KIND_HALF ? (pd - sg) * (pd + sg) : pd - sg; };
        for (int tries = 0; tries < 64; ++tries) {
            uint64_t hit = 0;                       // (branch-free: the loop vectorises)
            for (int64_t l = 0; l < N; ++l) {
                const double d = pad_shift(pad, hD[l]);
                uint64_t bits;
                std::memcpy(&bits, &d, 8);
                hit |= (uint64_t)((bits & 0x000fffffffffffffull) == 0);
            }
            if (!hit) break;
            pad = (L.kind == KIND_HALF) ? std::nextafter(pad, 0.0) : std::nextafter(pad, -INFINITY);
        }
        for (int64_t l = N; l < Np; ++l) { hPadD[l] = pad; hPadW[l] = 0.0; hPadW2[l] = 0.0; }

        if ((rc = ensure(ctx, ctx->dD, 3 * Np * 8))) return rc;
        AH_CUDA(cudaEventRecord(ctx->ev[0], st));
        AH_CUDA(cudaMemcpyAsync(ctx->dD.p, ctx->hIn, 3 * Np * 8, cudaMemcpyHostToDevice, st));
        ctx->stats.h2d_bytes += 3 * Np * 8;
        P.D = (const double *)ctx->dD.p;
        P.w = P.D + Np;
        P.w2 = P.D + 2 * Np;
        P.Np = Np;
        P.pad = pad;
        P.z = nullptr;
        if (hZ) {
            if ((rc = ensure(ctx, ctx->dZ, N * 8))) return rc;
            AH_CUDA(cudaMemcpyAsync(ctx->dZ.p, hZ, N * 8, cudaMemcpyHostToDevice, st));
            ctx->stats.h2d_bytes += N * 8;
            P.z = (const double *)ctx->dZ.p;
        }
        ctx->res_valid = true; ctx->res_kind = L.kind; ctx->res_N = L.N; ctx->res_nvec_u = L.nvec_u; ctx->res_nvec_v = L.nvec_v;
        ctx->res_kmax = L.kmax; ctx->res_P = P;
    } else {
        // ah_resolve: the padded inputs of the previous solve are still on the device
        AH_CUDA(cudaEventRecord(ctx->ev[0], st));
    }
    Outputs O{};
    if (out->rowmap_u) {
        if ((rc = ensure(ctx, ctx->dMap1, L.nvec_u * 8))) return rc;
        AH_CUDA(cudaMemcpyAsync(ctx->dMap1.p, out->rowmap_u, L.nvec_u * 8, cudaMemcpyHostToDevice, st));
        O.rowmap_u = (const int64_t *)ctx->dMap1.p;
        ctx->stats.h2d_bytes += L.nvec_u * 8;
    }
    if (L.kind == KIND_HALF && out->rowmap_v) {
        if ((rc = ensure(ctx, ctx->dMap2, L.nvec_v * 8))) return rc;
        AH_CUDA(cudaMemcpyAsync(ctx->dMap2.p, out->rowmap_v, L.nvec_v * 8, cudaMemcpyHostToDevice, st));
        O.rowmap_v = (const int64_t *)ctx->dMap2.p;
        ctx->stats.h2d_bytes += L.nvec_v * 8;
    }
    if (L.kind == KIND_HALF && out->rowscale_v) {
        if ((rc = ensure(ctx, ctx->dScale, N * 8))) return rc;
        AH_CUDA(cudaMemcpyAsync(ctx->dScale.p, out->rowscale_v, N * 8, cudaMemcpyHostToDevice, st));
        O.rowscale_v = (const double *)ctx->dScale.p;
        ctx->stats.h2d_bytes += N * 8;
    }
    AH_CUDA(cudaEventRecord(ctx->ev[1], st));

    // ---- per-k outputs (device SoA, one allocation)
    if ((rc = ensure(ctx, ctx->dOut, out_bytes(cnt)))) return rc;
    {
        char *p = (char *)ctx->dOut.p;
        O.lambda = carve<double>(p, cnt);
        O.sind = carve<int64_t>(p, cnt);
        O.kb = carve<double>(p, cnt);
        O.kz = carve<double>(p, cnt);
        O.knu = carve<double>(p, cnt);
        O.krho = carve<double>(p, cnt);
        O.qout = carve<int64_t>(p, cnt);
        O.steps = carve<int32_t>(p, cnt);
        O.status = carve<int32_t>(p, cnt);
    }
    if ((rc = ensure(ctx, ctx->dList, cnt * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dKs, (size_t)cnt * (sizeof(ah::KState) + sizeof(ah::KExtra))))) return rc;   // KState[cnt] | KExtra[cnt]
    if ((rc = ensure(ctx, ctx->dRList, cnt * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dCount, 2048))) return rc;
    {   // requeue FIFO of k_bisect's time slicing; entries carry a per-launch tag, so it is cleared only when (re)allocated
        const size_t qbytes = ah::bisect_queue_entries(cnt) * 8;
        if (qbytes > ctx->dQueue.cap) {
            if ((rc = ensure(ctx, ctx->dQueue, qbytes))) return rc;
            AH_CUDA(cudaMemsetAsync(ctx->dQueue.p, 0, ctx->dQueue.cap, st));
            ctx->bisect_tag = 0;
        }
    }

    // ---- eigenvector destination: device-resident, or host through a double-buffered staging area
    const bool wantU = out->U != nullptr, wantV = (L.kind == KIND_HALF) && out->V != nullptr;
    const bool hostU = wantU && out->U_mem == AH_MEM_HOST, hostV = wantV && out->V_mem == AH_MEM_HOST;
    const bool staged = hostU || hostV;
    int64_t chunk = cnt;
    size_t colbytes = 0;
    if (staged) {
        colbytes = (size_t)((hostU ? L.nvec_u : 0) + (hostV ? L.nvec_v : 0)) * 8;
        const size_t budget = (size_t)256 << 20;
        chunk = std::max<int64_t>(1, std::min<int64_t>(cnt, (int64_t)(budget / colbytes)));
        // keep every launch wide enough to fill the machine
        chunk = std::max<int64_t>(chunk, std::min<int64_t>(cnt, 4096));
        if ((rc = ensure(ctx, ctx->dStage[0], chunk * colbytes))) return rc;
        if (chunk < cnt && (rc = ensure(ctx, ctx->dStage[1], chunk * colbytes))) return rc;
    }

    float ms_main = 0.f, ms_all = 0.f;
    int64_t n_full = 0;
    int nbuf = 0, n_count_reads = 0;
    // host-staged: a short first block gets the copy engine going early (the copies, not the kernels, bound this path)
    const int64_t first = (staged && chunk < cnt) ? std::min<int64_t>(chunk, 1024) : chunk;
    const int nchunks = (int)(1 + (cnt > first ? (cnt - first + chunk - 1) / chunk : 0));
    if ((size_t)(2 * nchunks) > ctx->hCountsCap) {   // per-block hand-over / Remedy-3 counts of this call (pinned)
        if (ctx->hCounts) cudaFreeHost(ctx->hCounts);
        ctx->hCounts = nullptr; ctx->hCountsCap = 0;
        const size_t want = std::max<size_t>(64, 2 * (size_t)nchunks);
        if (cudaMallocHost((void **)&ctx->hCounts, want * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMallocHost failed"); }
        ctx->hCountsCap = want;
    }
    for (int q = 0; q < 2 * nchunks; ++q) ctx->hCounts[q] = 0;
    while ((int)ctx->chunk_ev.size() < 5 * nchunks) {
        cudaEvent_t e;
        AH_CUDA(cudaEventCreate(&e));
        ctx->chunk_ev.push_back(e);
    }
    while ((int)ctx->side_ev.size() < 4 * nchunks) {
        cudaEvent_t e;
        AH_CUDA(cudaEventCreate(&e));
        ctx->side_ev.push_back(e);
    }
    // Single-block calls with eigenvectors: the per-k results leave on the side stream as soon as the Remedy-3 round and
    // k_full are through, and the caller's arrays are filled from them while the main k_vec is still writing U.
    if ((rc = ensure_pinned(ctx, out_bytes(cnt)))) return rc;
    const bool early_perk = ctx->mode == 0 && nchunks == 1 && (wantU || wantV);
    ctx->stats.host_pre_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
    for (int64_t c0 = 0, cc = 0; c0 < cnt; c0 += cc, ++nbuf) {
        cc = std::min(nbuf == 0 ? first : chunk, cnt - c0);
        const int slot = nbuf & 1;
        cudaEvent_t *E = &ctx->chunk_ev[5 * (size_t)nbuf];   // start | prep done | bisect done | vec done | full done
        Outputs Oc = O;
        // outputs are indexed by k - kbase with kbase = k0 + c0; shift the SoA pointers accordingly
        Oc.lambda += c0; Oc.sind += c0; Oc.kb += c0; Oc.kz += c0; Oc.knu += c0; Oc.krho += c0;
        Oc.qout += c0; Oc.steps += c0; Oc.status += c0;
        if (staged) {
            if (nbuf >= 2) AH_CUDA(cudaStreamWaitEvent(st, ctx->ev_stage[slot], 0));  // D2H of this slot finished
            char *sp = (char *)ctx->dStage[slot].p;
            if (hostU) { Oc.U = (double *)sp; Oc.ldu = L.nvec_u; sp += (size_t)cc * L.nvec_u * 8; }
            else { Oc.U = wantU ? out->U + c0 * out->ldu : nullptr; Oc.ldu = out->ldu; }
            if (hostV) { Oc.V = (double *)sp; Oc.ldv = L.nvec_v; }
            else { Oc.V = wantV ? out->V + c0 * out->ldv : nullptr; Oc.ldv = out->ldv; }
        } else {
            Oc.U = wantU ? out->U + (out->rowmap_u ? 0 : c0 * out->ldu) : nullptr; Oc.ldu = out->ldu;
            Oc.V = wantV ? out->V + (out->rowmap_u ? 0 : c0 * out->ldv) : nullptr; Oc.ldv = out->ldv;   // (column index: rowmap_u, see above)
        }
        const int64_t kb = k0 + c0;
        int64_t *dlist = (int64_t *)ctx->dList.p;
        unsigned long long *dcount = (unsigned long long *)ctx->dCount.p;
        AH_CUDA(cudaEventRecord(E[0], st));
        bool side_full = false;      // k_full ran on the side stream (beside the main k_vec)
        if (ctx->mode == 0) {
            KState *ks = (KState *)ctx->dKs.p;
            int64_t *rlist = (int64_t *)ctx->dRList.p;
#ifdef AH_DIAG_TIMING
            AH_CUDA(cudaMemsetAsync(dcount, 0, 2048, st));
#else
            AH_CUDA(cudaMemsetAsync(dcount, 0, 128 + 256 * 4, st));
#endif
            rc = ah::launch_prep(L.kind, P, ks, kb, cc, dlist, dcount, st);
            if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_prep launch");
            AH_CUDA(cudaEventRecord(E[1], st));
            // The eigenpairs k_prep hands over (double-double b: K_b / K_z test failed) are latency-bound work for k_full, one
            // CTA each, ~0.6 ms: they start NOW on the high-priority side stream and run beside k_bisect instead of after
            // everything else.  k_bisect's grid would fill every SM to the last register, so as many of its CTAs as there are
            // such eigenpairs (16 at most) return at once and leave their places to k_full's -- only when the list is not empty.
            // dcount[6] = length of the hand-over list at this point; the late k_full launch begins there.
            AH_CUDA(cudaMemcpyAsync(dcount + 6, dcount, 8, cudaMemcpyDeviceToDevice, st));
            AH_CUDA(cudaEventRecord(ctx->ev_early0, st));
            AH_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_early0, 0));
            {
                const int grid = (int)std::min<int64_t>(cc, (int64_t)ctx->sm_count);
                switch (L.kind) {
                    case KIND_ARROW: k_full<KIND_ARROW, 256><<<grid, 256, 0, ctx->side_stream>>>(P, Oc, dlist, kb, 0, dcount + 6); break;
                    case KIND_DPR1: k_full<KIND_DPR1, 256><<<grid, 256, 0, ctx->side_stream>>>(P, Oc, dlist, kb, 0, dcount + 6); break;
                    default: k_full<KIND_HALF, 256><<<grid, 256, 0, ctx->side_stream>>>(P, Oc, dlist, kb, 0, dcount + 6); break;
                }
                AH_CUDA(cudaGetLastError());
                ctx->stats.launches += 1;
            }
            AH_CUDA(cudaEventRecord(ctx->ev_early1, ctx->side_stream));
            if (++ctx->bisect_tag == 0) {   // tags wrapped: forget the old entries
                AH_CUDA(cudaMemsetAsync(ctx->dQueue.p, 0, ctx->dQueue.cap, st));
                ctx->bisect_tag = 1;
            }
            rc = ah::launch_bisect(L.kind, 0, P, Oc, ks, kb, cc, dlist, rlist, dcount, (unsigned long long *)ctx->dQueue.p,
                                   (long long)(ctx->dQueue.cap / 8), ctx->bisect_tag, ctx->quantum, ctx->window, ctx->spec, ctx->sm_count, st,
                                   (ah::KExtra *)(ks + cnt), ctx->accel);
            if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_bisect launch");
            ctx->stats.launches += 2;   // k_prep, k_bisect round 0
            AH_CUDA(cudaEventRecord(E[2], st));
            // Remedy 3 on the fast path: re-shift (k_rem3_prep) and bisect again (round 1) for the ~1 % of eigenpairs
            // on the list (its length lives on the device).  The round is short and latency-bound, so it runs on a
            // high-priority side stream BESIDE the eigenvector kernel of everything round 0 finished; the listed
            // eigenpairs get their vectors from a second, list-driven k_vec on the side stream, too.
            {
                const int64_t guess = std::max<int64_t>(1, std::min<int64_t>(cc, (int64_t)ctx->sm_count * 2 * ah::BJ));   // full grid; idle slots are skipped
                const bool overlap = wantU || wantV;
                cudaStream_t rs = overlap ? ctx->side_stream : st;
                cudaEvent_t *SE = &ctx->side_ev[4 * (size_t)nbuf];
                if (overlap) {
                    AH_CUDA(cudaEventRecord(ctx->ev_fork, st));
                    AH_CUDA(cudaStreamWaitEvent(rs, ctx->ev_fork, 0));
                }
                AH_CUDA(cudaEventRecord(SE[0], rs));
                rc = ah::launch_rem3_prep(L.kind, P, Oc, ks, kb, guess, rlist, dlist, dcount, ctx->sm_count, rs, ctx->accel);
                if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_rem3_prep launch");
                rc = ah::launch_bisect(L.kind, 1, P, Oc, ks, kb, guess, dlist, rlist, dcount, nullptr, 0, 0u, 0, 0, 0, ctx->sm_count, rs, nullptr, ctx->accel);
                if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_bisect round 1 launch");
                AH_CUDA(cudaEventRecord(SE[1], rs));
                ctx->stats.launches += 2;
                if (overlap) {
                    // the listed eigenpairs get their vectors on the side stream, too (other columns than the main k_vec's)
                    rc = ah::launch_vec(L.kind, P, Oc, ks, kb, guess, rlist, dcount + 2, ctx->sm_count, rs);
                    if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_vec (Remedy-3 list) launch");
                    // ... and so does everything else that does not depend on the main k_vec: k_full (its own columns) and the
                    // per-k results on their way to the host, which the caller's arrays are filled from while k_vec still runs
                    AH_CUDA(cudaEventRecord(SE[2], rs));
                    {
                        const int grid = (int)std::min<int64_t>(cc, (int64_t)ctx->sm_count * 4);
                        switch (L.kind) {
                            case KIND_ARROW: k_full<KIND_ARROW, 256><<<grid, 256, 0, rs>>>(P, Oc, dlist, kb, 0, dcount, dcount + 6); break;
                            case KIND_DPR1: k_full<KIND_DPR1, 256><<<grid, 256, 0, rs>>>(P, Oc, dlist, kb, 0, dcount, dcount + 6); break;
                            default: k_full<KIND_HALF, 256><<<grid, 256, 0, rs>>>(P, Oc, dlist, kb, 0, dcount, dcount + 6); break;
                        }
                        AH_CUDA(cudaGetLastError());
                        ctx->stats.launches += 1;
                    }
                    AH_CUDA(cudaEventRecord(SE[3], rs));
                    AH_CUDA(cudaMemcpyAsync(&ctx->hCounts[2 * (size_t)nbuf], dcount, 8, cudaMemcpyDeviceToHost, rs));
                    AH_CUDA(cudaMemcpyAsync(&ctx->hCounts[2 * (size_t)nbuf + 1], dcount + 2, 8, cudaMemcpyDeviceToHost, rs));
                    n_count_reads = nbuf + 1;
                    if (early_perk) {
                        AH_CUDA(cudaEventRecord(ctx->ev[6], rs));
                        AH_CUDA(cudaMemcpyAsync(ctx->hPinned, ctx->dOut.p, out_bytes(cnt), cudaMemcpyDeviceToHost, rs));
                        AH_CUDA(cudaEventRecord(ctx->ev[7], rs));
                    }
                    AH_CUDA(cudaEventRecord(ctx->ev_join, rs));
                    rc = ah::launch_vec(L.kind, P, Oc, ks, kb, cc, nullptr, nullptr, ctx->sm_count, st);
                    if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_vec launch");
                    AH_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
                    ctx->stats.launches += 2;
                }
                side_full = overlap;
            }
            AH_CUDA(cudaEventRecord(E[3], st));
            if (!side_full) {
                // the hand-over list is consumed straight from device memory; a few hundred CTAs cover the
                // ~1 % of eigenpairs that take a Remedy / double-double branch
                AH_CUDA(cudaStreamWaitEvent(st, ctx->ev_early1, 0));     // (the early k_full of this block, on the side stream)
                const int grid = (int)std::min<int64_t>(cc, (int64_t)ctx->sm_count * 4);
                switch (L.kind) {
                    case KIND_ARROW: k_full<KIND_ARROW, 256><<<grid, 256, 0, st>>>(P, Oc, dlist, kb, 0, dcount, dcount + 6); break;
                    case KIND_DPR1: k_full<KIND_DPR1, 256><<<grid, 256, 0, st>>>(P, Oc, dlist, kb, 0, dcount, dcount + 6); break;
                    default: k_full<KIND_HALF, 256><<<grid, 256, 0, st>>>(P, Oc, dlist, kb, 0, dcount, dcount + 6); break;
                }
                AH_CUDA(cudaGetLastError());
                ctx->stats.launches += 1;
                AH_CUDA(cudaMemcpyAsync(&ctx->hCounts[2 * (size_t)nbuf], dcount, 8, cudaMemcpyDeviceToHost, st));
                AH_CUDA(cudaMemcpyAsync(&ctx->hCounts[2 * (size_t)nbuf + 1], dcount + 2, 8, cudaMemcpyDeviceToHost, st));
                n_count_reads = nbuf + 1;
            }
            AH_CUDA(cudaEventRecord(E[4], st));
        } else {
            AH_CUDA(cudaEventRecord(E[3], st));
            const int grid = (int)std::min<int64_t>(cc, (int64_t)ctx->sm_count * 4);
            switch (L.kind) {
                case KIND_ARROW: k_full<KIND_ARROW, 256><<<grid, 256, 0, st>>>(P, Oc, nullptr, kb, cc, nullptr); break;
                case KIND_DPR1: k_full<KIND_DPR1, 256><<<grid, 256, 0, st>>>(P, Oc, nullptr, kb, cc, nullptr); break;
                default: k_full<KIND_HALF, 256><<<grid, 256, 0, st>>>(P, Oc, nullptr, kb, cc, nullptr); break;
            }
            AH_CUDA(cudaGetLastError());
            ctx->stats.launches += 1;
            n_full += cc;
            AH_CUDA(cudaEventRecord(E[4], st));
        }
        if (staged) {
            // hand the finished block to the copy stream; the next block computes meanwhile
            AH_CUDA(cudaEventRecord(ctx->ev[5], st));
            AH_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[5], 0));
            char *sp = (char *)ctx->dStage[slot].p;
            if (hostU) {
                if (out->ldu == L.nvec_u)
                    AH_CUDA(cudaMemcpyAsync(out->U + c0 * out->ldu, sp, (size_t)cc * L.nvec_u * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
                else
                    AH_CUDA(cudaMemcpy2DAsync(out->U + c0 * out->ldu, out->ldu * 8, sp, L.nvec_u * 8, L.nvec_u * 8, cc,
                                              cudaMemcpyDeviceToHost, ctx->copy_stream));
                sp += (size_t)cc * L.nvec_u * 8;
                ctx->stats.d2h_bytes += cc * L.nvec_u * 8;
            }
            if (hostV) {
                if (out->ldv == L.nvec_v)
                    AH_CUDA(cudaMemcpyAsync(out->V + c0 * out->ldv, sp, (size_t)cc * L.nvec_v * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
                else
                    AH_CUDA(cudaMemcpy2DAsync(out->V + c0 * out->ldv, out->ldv * 8, sp, L.nvec_v * 8, L.nvec_v * 8, cc,
                                              cudaMemcpyDeviceToHost, ctx->copy_stream));
                ctx->stats.d2h_bytes += cc * L.nvec_v * 8;
            }
            AH_CUDA(cudaEventRecord(ctx->ev_stage[slot], ctx->copy_stream));
        }
    }

#ifdef AH_DIAG_TIMING
    {
        unsigned long long tg[20];
        AH_CUDA(cudaMemcpyAsync(tg, (unsigned long long *)ctx->dCount.p + 152, sizeof tg, cudaMemcpyDeviceToHost, st));
        AH_CUDA(cudaStreamSynchronize(st));
        const double ctas = (double)std::max<unsigned long long>(1, tg[7]), sw = (double)std::max<unsigned long long>(1, tg[6]);
        fprintf(stderr, "[diag] k_bisect warp 0, cycles per sweep (mean over %llu CTAs, %.1f sweeps each): boundary %.0f | barrier2 wait %.0f | "
                        "param copy %.0f | tiles %.0f | reduce %.0f | barrier1 wait %.0f\n", tg[7], sw / ctas, tg[0] / sw, tg[1] / sw, tg[2] / sw,
                tg[3] / sw, tg[4] / sw, tg[5] / sw);
        fprintf(stderr, "[diag] slot 0's boundary, cycles per sweep: decide %.0f | finalize %.0f (%.3f/sweep) | swap check %.0f (%.3f) | requeue %.0f (%.3f) | "
                        "ticket %.0f (%.3f) | resolve/poll %.0f (%.3f) | load KState %.0f\n", tg[8] / sw, tg[9] / sw, tg[15] / sw, tg[10] / sw, tg[16] / sw,
                tg[11] / sw, tg[17] / sw, tg[12] / sw, tg[18] / sw, tg[13] / sw, tg[19] / sw, tg[14] / sw);
    }
#endif
    // ---- per-k outputs -> host
    auto unpack = [&]() {
        char *p = (char *)ctx->hPinned;
        double *lam = carve<double>(p, cnt);
        int64_t *sind = carve<int64_t>(p, cnt);
        double *kb = carve<double>(p, cnt), *kz = carve<double>(p, cnt), *knu = carve<double>(p, cnt), *krho = carve<double>(p, cnt);
        int64_t *qout = carve<int64_t>(p, cnt);
        int32_t *steps = carve<int32_t>(p, cnt), *status = carve<int32_t>(p, cnt);
        std::memcpy(out->lambda, lam, cnt * 8);
        if (out->sind) std::memcpy(out->sind, sind, cnt * 8);
        if (out->kb) std::memcpy(out->kb, kb, cnt * 8);
        if (out->kz) std::memcpy(out->kz, kz, cnt * 8);
        if (out->knu) std::memcpy(out->knu, knu, cnt * 8);
        if (out->krho) std::memcpy(out->krho, krho, cnt * 8);
        if (out->qout) std::memcpy(out->qout, qout, cnt * 8);
        if (out->steps) std::memcpy(out->steps, steps, cnt * 4);
        if (out->status) std::memcpy(out->status, status, cnt * 4);
        int64_t tot = 0;
        for (int64_t c = 0; c < cnt; ++c) tot += steps[c];
        ctx->stats.total_steps = tot;
    };
    if (early_perk) {
        AH_CUDA(cudaEventSynchronize(ctx->ev_join));     // recorded on the side stream after the copy of the per-k results
        unpack();                                        // (the main k_vec is still writing U)
    } else {
        AH_CUDA(cudaEventRecord(ctx->ev[6], st));
        AH_CUDA(cudaMemcpyAsync(ctx->hPinned, ctx->dOut.p, out_bytes(cnt), cudaMemcpyDeviceToHost, st));
        AH_CUDA(cudaEventRecord(ctx->ev[7], st));
    }
    AH_CUDA(wait_stream(ctx, st));
    if (staged) AH_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    const auto t_synced = std::chrono::steady_clock::now();
    if (!early_perk) unpack();
    for (int c = 0; c < nchunks; ++c) {   // kernel times of every column block
        cudaEvent_t *E = &ctx->chunk_ev[5 * (size_t)c];
        float t = 0.f;
        AH_CUDA(cudaEventElapsedTime(&t, E[0], E[4]));
        ms_all += t;
        if (ctx->mode == 0 && (wantU || wantV)) {
            AH_CUDA(cudaEventElapsedTime(&t, ctx->side_ev[4 * (size_t)c + 2], ctx->side_ev[4 * (size_t)c + 3]));
        } else {
            AH_CUDA(cudaEventElapsedTime(&t, E[3], E[4]));
        }
        ctx->stats.full_ms += t;
        if (ctx->mode == 0) {
            AH_CUDA(cudaEventElapsedTime(&t, E[0], E[1]));
            ctx->stats.prep_ms += t;
            AH_CUDA(cudaEventElapsedTime(&t, E[1], E[2]));
            ctx->stats.bisect_ms += t;
            ms_main += t;
            AH_CUDA(cudaEventElapsedTime(&t, ctx->side_ev[4 * (size_t)c], ctx->side_ev[4 * (size_t)c + 1]));
            ctx->stats.rem3_ms += t;          // runs beside k_vec: contained in the span below when vectors are wanted
            ms_main += t;
            AH_CUDA(cudaEventElapsedTime(&t, E[2], E[3]));
            ctx->stats.vec_ms += t;
        }
    }
    ctx->stats.d2h_bytes += cnt * (8 * 7 + 4 * 2);
    float th = 0.f, td = 0.f;
    AH_CUDA(cudaEventElapsedTime(&th, ctx->ev[0], ctx->ev[1]));
    AH_CUDA(cudaEventElapsedTime(&td, ctx->ev[6], ctx->ev[7]));
    ctx->stats.h2d_ms = th;
    ctx->stats.d2h_ms = td;
    ctx->stats.kernel_ms = ms_all;
    ctx->stats.main_kernel_ms = ms_main;
    for (int q = 0; q < n_count_reads; ++q) { n_full += (int64_t)ctx->hCounts[2 * q]; ctx->stats.n_rem3 += (int64_t)ctx->hCounts[2 * q + 1]; }
    ctx->stats.n_full = n_full;
    ctx->stats.n_fast = cnt - n_full;
    ctx->stats.host_post_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_synced).count();
    return AH_OK;
}

void fill_tau(ah::Problem &P, const double *tau, int64_t npoles) {
    if (tau) {
        for (int j = 0; j < 5; ++j) P.tau[j] = tau[j];
    } else {
        P.tau[0] = 1e3; P.tau[1] = 10.0 * (double)npoles; P.tau[2] = 1e3; P.tau[3] = 1e3; P.tau[4] = 1e3;
    }
}

// ---- utility kernels ---------------------------------------------------------------------------
__global__ void k_set_identity(double *U, int64_t nrows, int64_t ncols, int64_t ld) {
    const int64_t total = nrows * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = idx / nrows, r = idx - c * nrows;
        U[c * ld + r] = (r == c) ? 1.0 : 0.0;
    }
}
__global__ void k_givens_rows(double *U, int64_t ld, int64_t ncols, int64_t i1, int64_t i2, double c, double s) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += (int64_t)gridDim.x * blockDim.x) {
        double a1 = U[j * ld + i1], a2 = U[j * ld + i2];
        U[j * ld + i1] = __dadd_rn(__dmul_rn(c, a1), __dmul_rn(s, a2));
        U[j * ld + i2] = __dadd_rn(__dmul_rn(-s, a1), __dmul_rn(c, a2));
    }
}
__global__ void k_gather(double *dst, int64_t ldd, const double *src, int64_t lds, const int64_t *rows, int64_t nrows,
                         const int64_t *cols, int64_t ncols) {
    const int64_t total = nrows * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = idx / nrows, r = idx - c * nrows;
        dst[c * ldd + r] = src[cols[c] * lds + rows[r]];
    }
}
// nb problems of equal size in one launch: work item = (problem b, eigenpair k), one CTA each
template <int KIND, int BT>
__global__ void __launch_bounds__(BT) k_full_batched(ah::Problem P0, ah::Outputs O0, int64_t nb, const double *__restrict__ avals,
                                                      const double *__restrict__ maxabs, int64_t ustride) {
    __shared__ double red[2 * (BT / 32) + 2];
    const int64_t per = P0.N + 1, total = nb * per;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t b = w / per, k = w - b * per + 1;
        ah::Problem P = P0;
        P.D += b * P0.N; P.w += b * P0.N;
        P.a = avals[b]; P.maxabs_w = maxabs[b];
        ah::Outputs O = O0;
        const int64_t o = b * per;
        O.lambda += o; O.sind += o; O.kb += o; O.kz += o; O.knu += o; O.krho += o; O.qout += o; O.steps += o; O.status += o;
        if (O0.U) O.U = O0.U + b * ustride;
        ah::solve_full<KIND, BT>(P, O, k, k - 1, red);
        __syncthreads();
    }
}
__global__ void k_gather_elems(double *dst, const double *__restrict__ src, const int64_t *__restrict__ idx, int64_t count) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) dst[t] = src[idx[t]];
}
__global__ void k_gather_batched(double *dst, int64_t dstride, int64_t ldd, const double *__restrict__ src, int64_t sstride, int64_t lds,
                                 const int64_t *__restrict__ rows, int64_t nrows, const int64_t *__restrict__ cols, int64_t ncols, int64_t nb) {
    const int64_t per = nrows * ncols, total = per * nb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / per, e = t - b * per, c = e / nrows, r = e - c * nrows;
        dst[b * dstride + c * ldd + r] = src[b * sstride + cols[b * ncols + c] * lds + rows[b * nrows + r]];
    }
}
__global__ void k_copy2d_batched(const uint64_t *__restrict__ dptr, int64_t ldd, const uint64_t *__restrict__ sptr, int64_t lds, int64_t nrows,
                                 int64_t ncols, int64_t nb) {
    const int64_t per = nrows * ncols, total = per * nb;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = t / per, e = t - b * per, c = e / nrows, r = e - c * nrows;
        reinterpret_cast<double *>(dptr[b])[c * ldd + r] = reinterpret_cast<const double *>(sptr[b])[c * lds + r];
    }
}
// Exhaustive accuracy of the reciprocal seed (rcp.approx.ftz.f64 = MUFU.RCP64H looks at the upper 32 bits of its operand
// only): every pattern of the 20 mantissa bits of the high word x {low word 0, low word all ones} x both signs x a few
// exponents; out[0] = max |1 - x r0| (exact: one FMA), out[1] = max |1 - x rcp_fast(x)| (the corrected reciprocal of the
// secular term) as bit patterns of non-negative doubles.
__global__ void k_rcp_seed_error(unsigned long long *out) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;          // 2^20 mantissa patterns x 2 low words x 2 signs
    const unsigned hi20 = idx & 0xfffffu, lowsel = (idx >> 20) & 1u, neg = (idx >> 21) & 1u;
    const int exps[5] = {1023, 1023 - 500, 1023 + 700, 1, 2045};
    double worst0 = 0.0, worst1 = 0.0;
    for (int q = 0; q < 5; ++q) {
        const unsigned long long bits = ((unsigned long long)neg << 63) | ((unsigned long long)exps[q] << 52) | ((unsigned long long)hi20 << 32) |
                                        (lowsel ? 0xffffffffull : 0ull);
        const double x = __longlong_as_double((long long)bits);
        const double r0 = ah::rcp_seed(x);
        const double e0 = fabs(__fma_rn(-x, r0, 1.0));
        const double e1 = fabs(__fma_rn(-x, ah::rcp_fast(x), 1.0));
        if (q < 3 || (r0 != 0.0 && fabs(r0) <= 1.7976931348623157e308)) {   // (the extreme exponents only where the reciprocal is a normal number)
            worst0 = fmax(worst0, e0);
            worst1 = fmax(worst1, e1);
        }
    }
    atomicMax(out, (unsigned long long)__double_as_longlong(worst0));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(worst1));
}
// the arrow driver's final view(U,rows,es) and the Hermitian wrappers' Diagonal(c)*U in ONE pass over the eigenvectors
template <int NC>
__global__ void k_gather_phase(double *dst, int64_t ldd, const double *__restrict__ src, int64_t lds, const int64_t *__restrict__ rows,
                               int64_t nrows, const int64_t *__restrict__ cols, int64_t ncols, const double *__restrict__ phase) {
    const int64_t total = nrows * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / nrows, r = idx - c * nrows;
        const double v = src[cols[c] * lds + rows[r]];
        double *o = dst + (c * ldd + r) * NC;
        if (NC == 2) {
            *reinterpret_cast<double2 *>(o) = make_double2(__dmul_rn(phase[2 * r], v), __dmul_rn(phase[2 * r + 1], v));
        } else {
            reinterpret_cast<double2 *>(o)[0] = make_double2(__dmul_rn(phase[4 * r], v), __dmul_rn(phase[4 * r + 1], v));
            reinterpret_cast<double2 *>(o)[1] = make_double2(__dmul_rn(phase[4 * r + 2], v), __dmul_rn(phase[4 * r + 3], v));
        }
    }
}
template <int NC>
__global__ void k_scale_rows_phase(double *dst, int64_t ldd, const double *__restrict__ src, int64_t lds,
                                   const double *__restrict__ phase, int64_t nrows, int64_t ncols) {
    const int64_t total = nrows * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = idx / nrows, r = idx - c * nrows;
        const double v = src[c * lds + r];
        double *o = dst + (c * ldd + r) * NC;
        if (NC == 2) {
            *reinterpret_cast<double2 *>(o) = make_double2(__dmul_rn(phase[2 * r], v), __dmul_rn(phase[2 * r + 1], v));
        } else {
            reinterpret_cast<double2 *>(o)[0] = make_double2(__dmul_rn(phase[4 * r], v), __dmul_rn(phase[4 * r + 1], v));
            reinterpret_cast<double2 *>(o)[1] = make_double2(__dmul_rn(phase[4 * r + 2], v), __dmul_rn(phase[4 * r + 3], v));
        }
    }
}
__global__ void k_dfma_burst(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3;
    double x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
// The instruction mix of k_bisect's inner block from registers only: 4 poles x 8 (sigma, m) pairs = 32 independent
// secular terms per iteration (7 FP64-pipe instructions + 1 MUFU.RCP64H each), no memory traffic, no barriers.
__global__ void __launch_bounds__(256, 2) k_term_burst(double *out, int iters, double d0, double s0, double m0) {
    double Dv[4], sg[8], mm[8], acc[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) Dv[e] = d0 + 1e-3 * e + 1e-9 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sg[j] = s0 + 1e-2 * j; mm[j] = m0 * (1.0 + 1e-2 * j); acc[j] = 0.0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) Dv[e] = __dadd_rn(Dv[e], 1e-9);    // new "tile" (4 extra DADD per 224 instructions)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j] = ah::secular_term_acc(acc[j], __dadd_rn(Dv[e], -sg[j]), mm[j], 1.0 + 1e-3 * e);
    }
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += acc[j];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = t;
}
__global__ void k_store_stream(double2 *dst, int64_t n2, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = make_double2(v, v);
}

}  // namespace

// ================================================================================================
extern "C" {

int ah_version(void) { return AH_VERSION; }

int ah_create(ah_ctx **pctx, int device) {
    if (!pctx) return AH_ERR_INVALID_ARG;
    *pctx = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, AH_ERR_NO_DEVICE, "no CUDA device available (libarrowhead_b200 has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, AH_ERR_INVALID_ARG, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, AH_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                                   std::to_string(prop.minor) + "; this library is built for sm_100a only");
    ah_ctx *ctx = new (std::nothrow) ah_ctx();
    if (!ctx) return fail(nullptr, AH_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *q = getenv("ARROWHEAD_B200_QUANTUM")) ctx->quantum = atoi(q);
    if (const char *q = getenv("ARROWHEAD_B200_WINDOW")) ctx->window = std::max(1, atoi(q));
    if (const char *q = getenv("ARROWHEAD_B200_SPEC")) ctx->spec = atoi(q) != 0;
    if (const char *q = getenv("ARROWHEAD_B200_ACCEL")) ctx->accel = atoi(q) != 0;
    if (const char *q = getenv("ARROWHEAD_B200_SPINWAIT")) ctx->spinwait = atoi(q) != 0;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { delete ctx; return cuda_fail(nullptr, e, "cudaSetDevice"); }
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = highest priority
        ok = ok && cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_early0, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_early1, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
    }
    for (auto &ev : ctx->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    for (auto &ev : ctx->ev_stage) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    if (ok) { ok = cudaMallocHost((void **)&ctx->hCounts, 64 * sizeof(unsigned long long)) == cudaSuccess; if (ok) ctx->hCountsCap = 64; }
    if (ok) ok = ah::configure_pipe_kernels() == 0;
    if (!ok) {
        cudaError_t le = cudaGetLastError();
        ah_destroy(ctx);
        return cuda_fail(nullptr, le == cudaSuccess ? cudaErrorUnknown : le, "context setup");
    }
    ctx->stream = ctx->own_stream;
    *pctx = ctx;
    return AH_OK;
}

void ah_destroy(ah_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (DevBuf *b : {&ctx->dD, &ctx->dW, &ctx->dW2, &ctx->dKs, &ctx->dRList, &ctx->dZ, &ctx->dMap1, &ctx->dMap2, &ctx->dScale, &ctx->dOut, &ctx->dList,
                      &ctx->dCount, &ctx->dQueue, &ctx->dStage[0], &ctx->dStage[1]})
        if (b->p) cudaFree(b->p);
    if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
    if (ctx->hIn) cudaFreeHost(ctx->hIn);
    if (ctx->hCounts) cudaFreeHost(ctx->hCounts);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->ev_stage) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->chunk_ev) cudaEventDestroy(ev);
    if (ctx->hArena) cudaFreeHost(ctx->hArena);
    for (DevBuf *b : {&ctx->dPtrA, &ctx->dPtrB, &ctx->dPtrC, &ctx->dIdx, &ctx->dBatchA})
        if (b->p) cudaFree(b->p);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_early0) cudaEventDestroy(ctx->ev_early0);
    if (ctx->ev_early1) cudaEventDestroy(ctx->ev_early1);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (auto &ev : ctx->side_ev) cudaEventDestroy(ev);
    ah_release_nccl(ctx);
    delete ctx;
}

const char *ah_last_error(const ah_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ah_set_stream(ah_ctx *ctx, void *cuda_stream) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return AH_OK;
}
int ah_get_stats(const ah_ctx *ctx, ah_stats *out) {
    if (!ctx || !out) return AH_ERR_INVALID_ARG;
    *out = ctx->stats;
    return AH_OK;
}
int ah_set_mode(ah_ctx *ctx, int mode) {
    if (!ctx || (mode != 0 && mode != 1)) return AH_ERR_INVALID_ARG;
    ctx->mode = mode;
    return AH_OK;
}

int ah_symarrow_eigen(ah_ctx *ctx, int64_t N, const double *D, const double *z, double a, const double *tau,
                      int64_t k0, int64_t k1, ah_result *out) {
    int rc = validate_common(ctx, N, D, z, N, k0, k1, N + 1, out, false);
    if (rc) return rc;
    if (!std::isfinite(a)) return fail(ctx, AH_ERR_INVALID_ARG, "non-finite arrow top");
    ah::Problem P{};
    P.N = N;
    P.a = a;
    double mz = 0.0;
    for (int64_t l = 0; l < N; ++l) mz = std::max(mz, std::fabs(z[l]));
    P.maxabs_w = mz;
    fill_tau(P, tau, N);
    Launch L{ah::KIND_ARROW, N, N + 1, 0, N + 1};
    return run_solver(ctx, L, P, D, z, nullptr, k0, k1, out);
}

int ah_symdpr1_eigen(ah_ctx *ctx, int64_t n, const double *D, const double *u, double rho, const double *tau,
                     int64_t k0, int64_t k1, ah_result *out) {
    int rc = validate_common(ctx, n, D, u, n, k0, k1, n, out, false);
    if (rc) return rc;
    if (n < 2) return fail(ctx, AH_ERR_INVALID_ARG, "SymDPR1 needs n >= 2 (the 1x1 case is closed-form, arrowhead4.jl:456-471)");
    if (!(rho > 0.0) || !std::isfinite(rho)) return fail(ctx, AH_ERR_INVALID_ARG, "rho must be > 0 (negate the matrix first, arrowhead4.jl:443-449)");
    ah::Problem P{};
    P.N = n;
    P.a = rho;
    P.sumabs_w = jl_pairwise_abs_sum(u, 0, n - 1);
    fill_tau(P, tau, n);
    Launch L{ah::KIND_DPR1, n, n, 0, n};
    return run_solver(ctx, L, P, D, u, nullptr, k0, k1, out);
}

int ah_halfarrow_svd(ah_ctx *ctx, int64_t nd, int64_t nz, const double *D, const double *z, const double *tau,
                     int64_t k0, int64_t k1, ah_result *out) {
    if (ctx && nz != nd && nz != nd + 1) return fail(ctx, AH_ERR_INVALID_ARG, "length(z) must be length(D) or length(D)+1");
    int rc = validate_common(ctx, nd, D, z, nz, k0, k1, nz, out, true);
    if (rc) return rc;
    ah::Problem P{};
    P.N = nd;
    P.has_tip = nz == nd + 1;
    P.ztip = P.has_tip ? z[nd] : 0.0;
    std::vector<double> z1((size_t)nd);
    double mz = 0.0;
    for (int64_t l = 0; l < nd; ++l) {
        volatile double t = z[l] * D[l];
        z1[l] = t;
        mz = std::max(mz, std::fabs(z1[l]));
    }
    P.maxabs_w = mz;
    volatile double zz = 0.0;
    hdd acc{0.0, 0.0};
    for (int64_t l = 0; l < nz; ++l) {
        volatile double sq = z[l] * z[l];
        zz = zz + sq;
        acc = h_add(acc, h_mul12(z[l], z[l]));
    }
    P.zz = zz;
    P.zz_hi = acc.hi;
    P.zz_lo = acc.lo;
    fill_tau(P, tau, nd);
    Launch L{ah::KIND_HALF, nd, nz, nd + 1, nz};
    return run_solver(ctx, L, P, D, z1.data(), z, k0, k1, out);
}

int ah_resolve(ah_ctx *ctx, int64_t k0, int64_t k1, ah_result *out) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    if (!out || !out->lambda) return fail(ctx, AH_ERR_INVALID_ARG, "null argument");
    if (!ctx->res_valid) return fail(ctx, AH_ERR_INVALID_ARG, "ah_resolve: no inputs resident (call ah_symarrow_eigen / ah_symdpr1_eigen / ah_halfarrow_svd on this context first)");
    if (k0 < 1 || k1 < k0 || k1 > ctx->res_kmax) return fail(ctx, AH_ERR_INVALID_ARG, "k-range outside 1.." + std::to_string(ctx->res_kmax));
    Launch L{ctx->res_kind, ctx->res_N, ctx->res_nvec_u, ctx->res_nvec_v, ctx->res_kmax};
    return run_solver(ctx, L, ctx->res_P, nullptr, nullptr, nullptr, k0, k1, out);
}

int ah_device_malloc(ah_ctx *ctx, void **dptr, int64_t bytes) {
    if (!ctx || !dptr || bytes < 0) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMalloc failed"); }
    return AH_OK;
}
int ah_device_free(ah_ctx *ctx, void *dptr) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    AH_CUDA(cudaFree(dptr));
    return AH_OK;
}
int ah_host_alloc_pinned(ah_ctx *ctx, void **hptr, int64_t bytes) {
    if (!ctx || !hptr || bytes < 0) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMallocHost(hptr, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMallocHost failed"); }
    return AH_OK;
}
int ah_host_free_pinned(ah_ctx *ctx, void *hptr) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaFreeHost(hptr));
    return AH_OK;
}
int ah_memcpy_d2h(ah_ctx *ctx, void *dst_host, const void *src_dev, int64_t bytes) {
    if (!ctx || !dst_host || !src_dev || bytes < 0) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    AH_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    AH_CUDA(cudaStreamSynchronize(ctx->stream));
    return AH_OK;
}
int ah_memcpy_h2d(ah_ctx *ctx, void *dst_dev, const void *src_host, int64_t bytes) {
    if (!ctx || !dst_dev || !src_host || bytes < 0) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    AH_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    AH_CUDA(cudaStreamSynchronize(ctx->stream));
    return AH_OK;
}

int ah_set_identity(ah_ctx *ctx, double *U_dev, int64_t nrows, int64_t ncols, int64_t ld) {
    if (!ctx || !U_dev || nrows < 1 || ncols < 1 || ld < nrows) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    k_set_identity<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(U_dev, nrows, ncols, ld);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}
int ah_apply_givens_rows(ah_ctx *ctx, double *U_dev, int64_t ld, int64_t ncols, int64_t i1, int64_t i2, double c, double s) {
    if (!ctx || !U_dev || ncols < 1 || i1 < 0 || i2 < 0 || i1 >= ld || i2 >= ld || i1 == i2) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    k_givens_rows<<<(int)std::min<int64_t>((ncols + 255) / 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(U_dev, ld, ncols, i1, i2, c, s);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}
int ah_gather_permuted(ah_ctx *ctx, double *dst_dev, int64_t ld_dst, const double *src_dev, int64_t ld_src,
                       const int64_t *rows, int64_t nrows, const int64_t *cols, int64_t ncols) {
    if (!ctx || !dst_dev || !src_dev || !rows || !cols || nrows < 1 || ncols < 1 || ld_dst < nrows) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dMap1, nrows * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dMap2, ncols * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dMap1.p, rows, nrows * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dMap2.p, cols, ncols * 8))) return rc;
    k_gather<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dst_dev, ld_dst, src_dev, ld_src, (const int64_t *)ctx->dMap1.p, nrows,
                                                          (const int64_t *)ctx->dMap2.p, ncols);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}

int ah_symarrow_eigen_batched(ah_ctx *ctx, int64_t nb, int64_t N, const double *D, const double *z, const double *a,
                              const double *tau, ah_result *out, int64_t ustride) {
    using namespace ah;
    if (!ctx) return AH_ERR_INVALID_ARG;
    if (!D || !z || !a || !out || !out->lambda || nb < 1 || N < 1) return fail(ctx, AH_ERR_INVALID_ARG, "null argument or empty batch");
    if (out->U && (out->U_mem != AH_MEM_DEVICE || out->ldu < N + 1 || ustride < (N + 1) * out->ldu))
        return fail(ctx, AH_ERR_INVALID_ARG, "batched solve needs device-resident U blocks (ldu >= N+1, ustride >= (N+1)*ldu)");
    std::vector<double> maxabs((size_t)nb);
    for (int64_t b = 0; b < nb; ++b) {
        const double *Db = D + b * N, *zb = z + b * N;
        double mz = 0.0;
        for (int64_t l = 0; l < N; ++l) {
            if (!std::isfinite(Db[l]) || (l && !(Db[l - 1] > Db[l]))) return fail(ctx, AH_ERR_NOT_ORDERED, "batched: D must be strictly decreasing in every problem");
            if (zb[l] == 0.0 || !std::isfinite(zb[l])) return fail(ctx, AH_ERR_REDUCIBLE, "batched: zero or non-finite entry in z");
            mz = std::max(mz, std::fabs(zb[l]));
        }
        if (!std::isfinite(a[b])) return fail(ctx, AH_ERR_INVALID_ARG, "batched: non-finite arrow top");
        maxabs[b] = mz;
    }
    ctx->stats = ah_stats{};
    AH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t cnt = nb * (N + 1);
    int rc;
    ctx->res_valid = false;
    if ((rc = ensure(ctx, ctx->dD, nb * N * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dW, nb * N * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dBatchA, 2 * nb * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dOut, out_bytes(cnt)))) return rc;
    AH_CUDA(cudaMemcpyAsync(ctx->dD.p, D, nb * N * 8, cudaMemcpyHostToDevice, st));
    AH_CUDA(cudaMemcpyAsync(ctx->dW.p, z, nb * N * 8, cudaMemcpyHostToDevice, st));
    AH_CUDA(cudaMemcpyAsync(ctx->dBatchA.p, a, nb * 8, cudaMemcpyHostToDevice, st));
    AH_CUDA(cudaMemcpyAsync((double *)ctx->dBatchA.p + nb, maxabs.data(), nb * 8, cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes = 2 * nb * N * 8 + 2 * nb * 8;
    Problem P{};
    P.N = N; P.D = (const double *)ctx->dD.p; P.w = (const double *)ctx->dW.p;
    fill_tau(P, tau, N);
    Outputs O{};
    {
        char *p = (char *)ctx->dOut.p;
        O.lambda = carve<double>(p, cnt); O.sind = carve<int64_t>(p, cnt);
        O.kb = carve<double>(p, cnt); O.kz = carve<double>(p, cnt); O.knu = carve<double>(p, cnt); O.krho = carve<double>(p, cnt);
        O.qout = carve<int64_t>(p, cnt); O.steps = carve<int32_t>(p, cnt); O.status = carve<int32_t>(p, cnt);
    }
    O.U = out->U; O.ldu = out->ldu;
    AH_CUDA(cudaEventRecord(ctx->ev[0], st));
    const int grid = (int)std::min<int64_t>(cnt, (int64_t)ctx->sm_count * 32);
    const double *av = (const double *)ctx->dBatchA.p, *mx = av + nb;
    if (N <= 64) k_full_batched<KIND_ARROW, 32><<<grid, 32, 0, st>>>(P, O, nb, av, mx, ustride);
    else if (N <= 512) k_full_batched<KIND_ARROW, 128><<<grid, 128, 0, st>>>(P, O, nb, av, mx, ustride);
    else k_full_batched<KIND_ARROW, 256><<<grid, 256, 0, st>>>(P, O, nb, av, mx, ustride);
    AH_CUDA(cudaGetLastError());
    AH_CUDA(cudaEventRecord(ctx->ev[1], st));
    if ((rc = ensure_pinned(ctx, out_bytes(cnt)))) return rc;
    AH_CUDA(cudaMemcpyAsync(ctx->hPinned, ctx->dOut.p, out_bytes(cnt), cudaMemcpyDeviceToHost, st));
    AH_CUDA(cudaStreamSynchronize(st));
    {
        char *p = (char *)ctx->hPinned;
        double *lam = carve<double>(p, cnt);
        int64_t *sind = carve<int64_t>(p, cnt);
        double *kb = carve<double>(p, cnt), *kz = carve<double>(p, cnt), *knu = carve<double>(p, cnt), *krho = carve<double>(p, cnt);
        int64_t *qout = carve<int64_t>(p, cnt);
        int32_t *steps = carve<int32_t>(p, cnt), *status = carve<int32_t>(p, cnt);
        std::memcpy(out->lambda, lam, cnt * 8);
        if (out->sind) std::memcpy(out->sind, sind, cnt * 8);
        if (out->kb) std::memcpy(out->kb, kb, cnt * 8);
        if (out->kz) std::memcpy(out->kz, kz, cnt * 8);
        if (out->knu) std::memcpy(out->knu, knu, cnt * 8);
        if (out->krho) std::memcpy(out->krho, krho, cnt * 8);
        if (out->qout) std::memcpy(out->qout, qout, cnt * 8);
        if (out->steps) std::memcpy(out->steps, steps, cnt * 4);
        if (out->status) std::memcpy(out->status, status, cnt * 4);
    }
    float ms = 0.f;
    AH_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.kernel_ms = ms; ctx->stats.full_ms = ms; ctx->stats.launches = 1; ctx->stats.n_full = cnt;
    ctx->stats.d2h_bytes = cnt * (8 * 7 + 4 * 2);
    return AH_OK;
}

int ah_symarrow_eigvals_dd(ah_ctx *ctx, int64_t N, const double *D, const double *z_hi, const double *z_lo, double a_hi,
                           double a_lo, const double *tau, int64_t k0, int64_t k1, ah_result *out) {
    using namespace ah;
    int rc = validate_common(ctx, N, D, z_hi, N, k0, k1, N + 1, out, false);
    if (rc) return rc;
    if (!z_lo || !tau) return fail(ctx, AH_ERR_INVALID_ARG, "z_lo and tau are required (rootsah passes [1e2,1e2,1e2,1e2,1e2])");
    if (out->U || out->V) return fail(ctx, AH_ERR_INVALID_ARG, "eigenvalue-only entry point: U and V must be NULL");
    if (!std::isfinite(a_hi) || !std::isfinite(a_lo)) return fail(ctx, AH_ERR_INVALID_ARG, "non-finite arrow top");
    ctx->stats = ah_stats{};
    AH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t cnt = k1 - k0 + 1;
    ctx->res_valid = false;
    if ((rc = ensure(ctx, ctx->dD, N * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dW, N * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dW2, N * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dOut, out_bytes(cnt)))) return rc;
    AH_CUDA(cudaMemcpyAsync(ctx->dD.p, D, N * 8, cudaMemcpyHostToDevice, st));
    AH_CUDA(cudaMemcpyAsync(ctx->dW.p, z_hi, N * 8, cudaMemcpyHostToDevice, st));
    AH_CUDA(cudaMemcpyAsync(ctx->dW2.p, z_lo, N * 8, cudaMemcpyHostToDevice, st));
    ctx->stats.h2d_bytes = 3 * N * 8;
    Problem P{};
    P.N = N; P.D = (const double *)ctx->dD.p; P.w = (const double *)ctx->dW.p; P.a = a_hi;
    for (int j = 0; j < 5; ++j) P.tau[j] = tau[j];
    Outputs O{};
    {
        char *p = (char *)ctx->dOut.p;
        O.lambda = carve<double>(p, cnt); O.sind = carve<int64_t>(p, cnt);
        O.kb = carve<double>(p, cnt); O.kz = carve<double>(p, cnt); O.knu = carve<double>(p, cnt); O.krho = carve<double>(p, cnt);
        O.qout = carve<int64_t>(p, cnt); O.steps = carve<int32_t>(p, cnt); O.status = carve<int32_t>(p, cnt);
    }
    AH_CUDA(cudaEventRecord(ctx->ev[0], st));
    const int grid = (int)std::min<int64_t>(cnt, (int64_t)ctx->sm_count * 8);
    if (N <= 512) k_roots<64><<<grid, 64, 0, st>>>(P, (const double *)ctx->dW2.p, a_lo, O, k0, cnt);
    else k_roots<256><<<grid, 256, 0, st>>>(P, (const double *)ctx->dW2.p, a_lo, O, k0, cnt);
    AH_CUDA(cudaGetLastError());
    AH_CUDA(cudaEventRecord(ctx->ev[1], st));
    if ((rc = ensure_pinned(ctx, out_bytes(cnt)))) return rc;
    AH_CUDA(cudaMemcpyAsync(ctx->hPinned, ctx->dOut.p, out_bytes(cnt), cudaMemcpyDeviceToHost, st));
    AH_CUDA(cudaStreamSynchronize(st));
    {
        char *p = (char *)ctx->hPinned;
        double *lam = carve<double>(p, cnt);
        int64_t *sind = carve<int64_t>(p, cnt);
        double *kb = carve<double>(p, cnt), *kz = carve<double>(p, cnt), *knu = carve<double>(p, cnt), *krho = carve<double>(p, cnt);
        int64_t *qout = carve<int64_t>(p, cnt);
        int32_t *steps = carve<int32_t>(p, cnt), *status = carve<int32_t>(p, cnt);
        std::memcpy(out->lambda, lam, cnt * 8);
        if (out->sind) std::memcpy(out->sind, sind, cnt * 8);
        if (out->kb) std::memcpy(out->kb, kb, cnt * 8);
        if (out->kz) std::memcpy(out->kz, kz, cnt * 8);
        if (out->knu) std::memcpy(out->knu, knu, cnt * 8);
        if (out->krho) std::memcpy(out->krho, krho, cnt * 8);
        if (out->qout) std::memcpy(out->qout, qout, cnt * 8);
        if (out->steps) std::memcpy(out->steps, steps, cnt * 4);
        if (out->status) std::memcpy(out->status, status, cnt * 4);
    }
    float ms = 0.f;
    AH_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.kernel_ms = ms; ctx->stats.full_ms = ms; ctx->stats.launches = 1; ctx->stats.n_full = cnt;
    ctx->stats.d2h_bytes = cnt * (8 * 7 + 4 * 2);
    return AH_OK;
}

int ah_gather_elements(ah_ctx *ctx, const double *src_dev, const int64_t *idx, int64_t count, double *out) {
    if (!ctx || !src_dev || !idx || !out || count < 1) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dIdx, count * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dScale, count * 8))) return rc;
    AH_CUDA(cudaMemcpyAsync(ctx->dIdx.p, idx, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_gather_elems<<<(int)std::min<int64_t>((count + 255) / 256, ctx->sm_count * 8), 256, 0, ctx->stream>>>(
        (double *)ctx->dScale.p, src_dev, (const int64_t *)ctx->dIdx.p, count);
    AH_CUDA(cudaGetLastError());
    AH_CUDA(cudaMemcpyAsync(out, ctx->dScale.p, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    AH_CUDA(cudaStreamSynchronize(ctx->stream));
    return AH_OK;
}

int ah_gather_permuted_batched(ah_ctx *ctx, int64_t nb, double *dst_dev, int64_t dst_stride, int64_t ld_dst,
                               const double *src_dev, int64_t src_stride, int64_t ld_src, const int64_t *rows,
                               int64_t nrows, const int64_t *cols, int64_t ncols) {
    if (!ctx || !dst_dev || !src_dev || !rows || !cols || nb < 1 || nrows < 1 || ncols < 1 || ld_dst < nrows) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dMap1, nb * nrows * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dMap2, nb * ncols * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dMap1.p, rows, nb * nrows * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dMap2.p, cols, nb * ncols * 8))) return rc;
    k_gather_batched<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dst_dev, dst_stride, ld_dst, src_dev, src_stride, ld_src,
                                                                  (const int64_t *)ctx->dMap1.p, nrows, (const int64_t *)ctx->dMap2.p, ncols, nb);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}

int ah_copy2d_batched(ah_ctx *ctx, int64_t nb, int64_t nrows, int64_t ncols, const uint64_t *dst_ptrs, int64_t ld_dst,
                      const uint64_t *src_ptrs, int64_t ld_src) {
    if (!ctx || !dst_ptrs || !src_ptrs || nb < 1 || nrows < 1 || ncols < 1) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dPtrA, nb * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dPtrC, nb * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dPtrA.p, src_ptrs, nb * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dPtrC.p, dst_ptrs, nb * 8))) return rc;
    k_copy2d_batched<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const uint64_t *)ctx->dPtrC.p, ld_dst, (const uint64_t *)ctx->dPtrA.p, ld_src,
                                                                  nrows, ncols, nb);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}

int ah_dgemm_batched(ah_ctx *ctx, int64_t nb, int64_t m, int64_t n, int64_t k, const uint64_t *A_ptrs, int64_t lda,
                     const uint64_t *B_ptrs, int64_t ldb, const uint64_t *C_ptrs, int64_t ldc) {
    if (!ctx || !A_ptrs || !B_ptrs || !C_ptrs || nb < 1 || m < 1 || n < 1 || k < 1 || lda < m || ldb < k || ldc < m) return AH_ERR_INVALID_ARG;
    if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX || nb > 65535) return fail(ctx, AH_ERR_INVALID_ARG, "ah_dgemm_batched: dimension too large");
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    // the three pointer arrays travel in one staged copy
    if ((rc = ensure(ctx, ctx->dPtrA, 3 * nb * 8))) return rc;
    std::vector<uint64_t> ptrs((size_t)(3 * nb));
    std::memcpy(ptrs.data(), A_ptrs, nb * 8);
    std::memcpy(ptrs.data() + nb, B_ptrs, nb * 8);
    std::memcpy(ptrs.data() + 2 * nb, C_ptrs, nb * 8);
    if ((rc = h2d_staged(ctx, ctx->dPtrA.p, ptrs.data(), 3 * nb * 8))) return rc;
    const uint64_t *dp = (const uint64_t *)ctx->dPtrA.p;
    AH_CUDA(cudaEventRecord(ctx->ev[8], ctx->stream));
    rc = ah::launch_dgemm(dp, dp + nb, dp + 2 * nb, (int)nb, (int)m, (int)n, (int)k, lda, ldb, ldc, ctx->sm_count, ctx->stream);
    if (rc) return cuda_fail(ctx, (cudaError_t)rc, "k_dgemm_dmma launch");
    AH_CUDA(cudaEventRecord(ctx->ev[9], ctx->stream));
    return util_done(ctx);
}

int ah_scale_rows_phase(ah_ctx *ctx, double *dst_dev, int64_t ld_dst, const double *src_dev, int64_t ld_src,
                        const double *phase, int ncomp, int64_t nrows, int64_t ncols) {
    if (!ctx || !dst_dev || !src_dev || !phase || (ncomp != 2 && ncomp != 4) || nrows < 1 || ncols < 1 || ld_dst < nrows || ld_src < nrows)
        return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dScale, (size_t)nrows * ncomp * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dScale.p, phase, (size_t)nrows * ncomp * 8))) return rc;
    const int grid = ctx->sm_count * 8;
    if (ncomp == 2) k_scale_rows_phase<2><<<grid, 256, 0, ctx->stream>>>(dst_dev, ld_dst, src_dev, ld_src, (const double *)ctx->dScale.p, nrows, ncols);
    else k_scale_rows_phase<4><<<grid, 256, 0, ctx->stream>>>(dst_dev, ld_dst, src_dev, ld_src, (const double *)ctx->dScale.p, nrows, ncols);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}

int ah_gather_permuted_phase(ah_ctx *ctx, double *dst_dev, int64_t ld_dst, const double *src_dev, int64_t ld_src, const int64_t *rows,
                             int64_t nrows, const int64_t *cols, int64_t ncols, const double *phase, int ncomp) {
    if (!ctx || !dst_dev || !src_dev || !rows || !cols || !phase || (ncomp != 2 && ncomp != 4) || nrows < 1 || ncols < 1 || ld_dst < nrows)
        return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dMap1, nrows * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dMap2, ncols * 8))) return rc;
    if ((rc = ensure(ctx, ctx->dScale, (size_t)nrows * ncomp * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dMap1.p, rows, nrows * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dMap2.p, cols, ncols * 8))) return rc;
    if ((rc = h2d_staged(ctx, ctx->dScale.p, phase, (size_t)nrows * ncomp * 8))) return rc;
    const int grid = ctx->sm_count * 8;
    if (ncomp == 2)
        k_gather_phase<2><<<grid, 256, 0, ctx->stream>>>(dst_dev, ld_dst, src_dev, ld_src, (const int64_t *)ctx->dMap1.p, nrows,
                                                          (const int64_t *)ctx->dMap2.p, ncols, (const double *)ctx->dScale.p);
    else
        k_gather_phase<4><<<grid, 256, 0, ctx->stream>>>(dst_dev, ld_dst, src_dev, ld_src, (const int64_t *)ctx->dMap1.p, nrows,
                                                          (const int64_t *)ctx->dMap2.p, ncols, (const double *)ctx->dScale.p);
    AH_CUDA(cudaGetLastError());
    return util_done(ctx);
}

int ah_set_async(ah_ctx *ctx, int on) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    ctx->async_utils = on != 0;
    return AH_OK;
}
int ah_wait(ah_ctx *ctx) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    AH_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->hArenaUsed = 0;
    AH_CUDA(cudaGetLastError());
    return AH_OK;
}
int ah_last_gemm_ms(ah_ctx *ctx, double *ms) {
    if (!ctx || !ms) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaEventSynchronize(ctx->ev[9]));
    float t = 0.f;
    AH_CUDA(cudaEventElapsedTime(&t, ctx->ev[8], ctx->ev[9]));
    *ms = t;
    return AH_OK;
}
int ah_measure_dmma_peak(ah_ctx *ctx, double *gfma_per_s, double *elapsed_ms) {
    if (!ctx || !gfma_per_s) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    const int threads = 256, blocks = ctx->sm_count * 2, iters = 1 << 12;
    int rc;
    if ((rc = ensure(ctx, ctx->dStage[0], (size_t)threads * blocks * 8))) return rc;
    double *o = (double *)ctx->dStage[0].p;
    ah::k_dmma_burst<<<blocks, threads, 0, ctx->stream>>>(o, 16, 0.999999, 1e-7);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        AH_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        ah::k_dmma_burst<<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999999, 1e-7);
        AH_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        AH_CUDA(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0.f;
        AH_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        best = std::min(best, ms);
    }
    AH_CUDA(cudaGetLastError());
    // one DMMA.8x8x4 = 8*8*4 = 256 FMA per warp
    *gfma_per_s = (double)(threads / 32) * blocks * (double)iters * 16.0 * 256.0 / (best * 1e-3) / 1e9;
    if (elapsed_ms) *elapsed_ms = best;
    return AH_OK;
}
int ah_measure_fp64_peak(ah_ctx *ctx, double *ginstr_per_s, double *elapsed_ms) {
    if (!ctx || !ginstr_per_s) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    const int threads = 512, blocks = ctx->sm_count * 4, iters = 1 << 15;
    int rc;
    if ((rc = ensure(ctx, ctx->dStage[0], (size_t)threads * blocks * 8))) return rc;
    double *o = (double *)ctx->dStage[0].p;
    k_dfma_burst<<<blocks, threads, 0, ctx->stream>>>(o, 256, 0.999999, 1e-7);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        AH_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        k_dfma_burst<<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.999999, 1e-7);
        AH_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        AH_CUDA(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0.f;
        AH_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        best = std::min(best, ms);
    }
    AH_CUDA(cudaGetLastError());
    const double instr = (double)threads * blocks * (double)iters * 8.0;
    *ginstr_per_s = instr / (best * 1e-3) / 1e9;
    if (elapsed_ms) *elapsed_ms = best;
    return AH_OK;
}
int ah_measure_term_peak(ah_ctx *ctx, double *gterms_per_s, double *elapsed_ms) {
    if (!ctx || !gterms_per_s) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int blocks = ctx->sm_count * 2;
    if (const char *e = getenv("ARROWHEAD_B200_BURST_CTAS")) blocks = ctx->sm_count * std::max(1, atoi(e));   // experiment: 1 CTA (8 warps) per SM alone
    const int threads = 256, iters = 1 << 13;
    int rc;
    if ((rc = ensure(ctx, ctx->dStage[0], (size_t)threads * blocks * 8))) return rc;
    double *o = (double *)ctx->dStage[0].p;
    k_term_burst<<<blocks, threads, 0, ctx->stream>>>(o, 64, 0.5, 0.25, 0.125);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        AH_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        k_term_burst<<<blocks, threads, 0, ctx->stream>>>(o, iters, 0.5, 0.25, 0.125);
        AH_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        AH_CUDA(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0.f;
        AH_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        best = std::min(best, ms);
    }
    AH_CUDA(cudaGetLastError());
    *gterms_per_s = (double)threads * blocks * (double)iters * 32.0 / (best * 1e-3) / 1e9;
    if (elapsed_ms) *elapsed_ms = best;
    return AH_OK;
}
int ah_measure_rcp_seed_error(ah_ctx *ctx, double *seed_err, double *corrected_err) {
    if (!ctx || !seed_err) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->dCount, 2048))) return rc;
    AH_CUDA(cudaMemsetAsync(ctx->dCount.p, 0, 16, ctx->stream));
    k_rcp_seed_error<<<(1u << 22) / 256, 256, 0, ctx->stream>>>((unsigned long long *)ctx->dCount.p);
    AH_CUDA(cudaGetLastError());
    unsigned long long h[2];
    AH_CUDA(cudaMemcpyAsync(h, ctx->dCount.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    AH_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(seed_err, &h[0], 8);
    if (corrected_err) std::memcpy(corrected_err, &h[1], 8);
    return AH_OK;
}
int ah_measure_store_bw(ah_ctx *ctx, int64_t bytes, double *gb_per_s) {
    if (!ctx || !gb_per_s || bytes < 16) return AH_ERR_INVALID_ARG;
    AH_CUDA(cudaSetDevice(ctx->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, AH_ERR_NOMEM, "cudaMalloc failed"); }
    const int64_t n2 = bytes / 16;
    k_store_stream<<<ctx->sm_count * 16, 512, 0, ctx->stream>>>((double2 *)p, n2, 1.0);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(ctx->ev[0], ctx->stream);
        k_store_stream<<<ctx->sm_count * 16, 512, 0, ctx->stream>>>((double2 *)p, n2, (double)rep);
        cudaEventRecord(ctx->ev[1], ctx->stream);
        cudaEventSynchronize(ctx->ev[1]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        best = std::min(best, ms);
    }
    e = cudaGetLastError();
    cudaFree(p);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "store bandwidth kernel");
    *gb_per_s = (double)(n2 * 16) / (best * 1e-3) / 1e9;
    return AH_OK;
}

}  // extern "C"

#include "ah_multi.inl"
