// ah_multi.inl -- several GPUs behind the C ABI (included at the end of ah_api.cu; needs struct ah_ctx).
//
// The reference parallelises the eigenpair loop inside ONE process (`Threads.@threads for k`, arrowhead3.jl:641-644).
// ah_multi gives a single-threaded caller (Julia's ccall) the same: one context per device, one worker thread per
// context living inside the library, the k-range split into contiguous blocks (ah_k_partition -- the blocks of
// arrowhead.jl_b200/sharding.py::k_partition).  No data-path collective: every device gets D, z (0.5 MB) and owns its
// eigenpairs and its columns of U.  The optional all-gather of U (north_star: "NCCL allgather over NVLink runs only
// when the caller needs U resident on one device") is ncclAllGather / ncclBroadcast, with libnccl.so.2 dlopen()ed on
// first use so that the library has no link-time NCCL dependency and shares torch's copy when both live in a process.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("ARROWHEAD_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
#define AH_NCCL_SYM(F) api.F = reinterpret_cast<decltype(api.F)>(dlsym(api.handle, "nccl" #F)); if (!api.F) { api.err = "libnccl: missing symbol nccl" #F; return; }
        AH_NCCL_SYM(GetUniqueId) AH_NCCL_SYM(CommInitRank) AH_NCCL_SYM(CommInitAll) AH_NCCL_SYM(CommDestroy) AH_NCCL_SYM(AllGather)
        AH_NCCL_SYM(Broadcast) AH_NCCL_SYM(GroupStart) AH_NCCL_SYM(GroupEnd) AH_NCCL_SYM(GetErrorString)
#undef AH_NCCL_SYM
    });
    return api.err.empty() ? &api : nullptr;
}

void k_partition_host(int64_t nk, int world, int rank, int64_t *k0, int64_t *k1) {
    const int64_t base = nk / world, rem = nk % world;
    const int64_t cnt = base + (rank < rem ? 1 : 0);
    const int64_t start = (int64_t)rank * base + std::min<int64_t>(rank, rem);
    if (cnt == 0) { *k0 = 1; *k1 = 0; } else { *k0 = start + 1; *k1 = start + cnt; }
}

// in-place all-gather of the column blocks of `world` full-size buffers; bufs[i] / comms[i] / streams[i] for the ranks
// this process drives (all of them: ah_multi; one: ah_allgather_columns).  Equal blocks -> ncclAllGather, else one
// ncclBroadcast per block inside a group.
int nccl_gather_blocks(NcclApi *nc, int world, const int *ranks, int nlocal, double *const *bufs, ncclComm_t *comms,
                       cudaStream_t *streams, const int *devices, int64_t ld, int64_t kmax, std::string &err) {
    ncclResult_t r = ncclSuccess;
    const bool equal = kmax % world == 0;
    if ((r = nc->GroupStart()) != ncclSuccess) { err = nc->GetErrorString(r); return AH_ERR_CUDA_BASE; }
    for (int j = 0; j < nlocal && r == ncclSuccess; ++j) {
        cudaSetDevice(devices[j]);
        if (equal) {
            const int64_t blk = kmax / world;
            r = nc->AllGather(bufs[j] + (int64_t)ranks[j] * blk * ld, bufs[j], (size_t)(blk * ld), ncclDouble, comms[j], streams[j]);
        } else {
            for (int root = 0; root < world && r == ncclSuccess; ++root) {
                int64_t k0, k1;
                k_partition_host(kmax, world, root, &k0, &k1);
                if (k1 < k0) continue;
                double *p = bufs[j] + (k0 - 1) * ld;
                r = nc->Broadcast(p, p, (size_t)((k1 - k0 + 1) * ld), ncclDouble, root, comms[j], streams[j]);
            }
        }
    }
    ncclResult_t r2 = nc->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { err = std::string("NCCL: ") + nc->GetErrorString(r); return AH_ERR_CUDA_BASE; }
    return AH_OK;
}

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
};

}  // namespace

struct ah_multi {
    int ndev = 0;
    std::vector<int> devices;
    std::vector<ah_ctx *> ctx;
    std::vector<Worker *> workers;
    std::vector<ncclComm_t> comms;     // created on the first all-gather
    std::string err;
    double elapsed_ms = 0.0;           // host wall time of the last ah_multi_* solve
    double gather_ms = 0.0;            // device time of the last all-gather (max over devices)
};

namespace {

void worker_loop(Worker *w) {
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [w] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        const int rc = job();
        lk.lock();
        w->rc = rc;
        w->done = true;
        w->cv.notify_all();
    }
}

// run fn(r) on every device's worker thread, wait for all, return the first failure
int run_on_all(ah_multi *m, const std::function<int(int)> &fn) {
    for (int r = 0; r < m->ndev; ++r) {
        Worker *w = m->workers[r];
        std::lock_guard<std::mutex> lk(w->mu);
        w->job = [fn, r] { return fn(r); };
        w->has_job = true; w->done = false;
        w->cv.notify_all();
    }
    int rc = AH_OK;
    for (int r = 0; r < m->ndev; ++r) {
        Worker *w = m->workers[r];
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv.wait(lk, [w] { return w->done; });
        if (w->rc != AH_OK && rc == AH_OK) { rc = w->rc; m->err = "device " + std::to_string(m->devices[r]) + ": " + m->ctx[r]->err; }
    }
    return rc;
}

// the part of `out` that device r owns: per-k arrays offset by its first eigenpair; U / V either the caller's host
// matrix at the block's first column, or the device's own full-size buffer at the same column offset
ah_result shard_result(const ah_result &out, int64_t k0, double *U_dev, double *V_dev) {
    ah_result s = out;
    const int64_t c0 = k0 - 1;
    s.lambda += c0;
    if (s.sind) s.sind += c0;
    if (s.kb) s.kb += c0;
    if (s.kz) s.kz += c0;
    if (s.knu) s.knu += c0;
    if (s.krho) s.krho += c0;
    if (s.qout) s.qout += c0;
    if (s.steps) s.steps += c0;
    if (s.status) s.status += c0;
    if (U_dev) { s.U = U_dev + c0 * out.ldu; s.U_mem = AH_MEM_DEVICE; }
    else if (s.U) s.U += c0 * out.ldu;
    if (V_dev) { s.V = V_dev + c0 * out.ldv; s.V_mem = AH_MEM_DEVICE; }
    else if (s.V) s.V += c0 * out.ldv;
    return s;
}

int multi_check(ah_multi *m, const ah_result *out, double *const *U_full, double *const *V_full) {
    if (!m) return AH_ERR_INVALID_ARG;
    if (!out || !out->lambda) { m->err = "null argument"; return AH_ERR_INVALID_ARG; }
    if (out->rowmap_u || out->rowmap_v) { m->err = "ah_multi_*: mapped stores (deflation) are not sharded; use one context"; return AH_ERR_INVALID_ARG; }
    if (!U_full && out->U && out->U_mem != AH_MEM_HOST) { m->err = "ah_multi_*: device-resident U goes through U_full (one buffer per device)"; return AH_ERR_INVALID_ARG; }
    if (!V_full && out->V && out->V_mem != AH_MEM_HOST) { m->err = "ah_multi_*: device-resident V goes through V_full (one buffer per device)"; return AH_ERR_INVALID_ARG; }
    return AH_OK;
}

template <typename Solve>
int multi_solve(ah_multi *m, int64_t kmax, const ah_result *out, double *const *U_full, double *const *V_full, Solve solve) {
    int rc = multi_check(m, out, U_full, V_full);
    if (rc) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    rc = run_on_all(m, [&](int r) -> int {
        int64_t k0, k1;
        k_partition_host(kmax, m->ndev, r, &k0, &k1);
        if (k1 < k0) return AH_OK;
        ah_result s = shard_result(*out, k0, U_full ? U_full[r] : nullptr, V_full ? V_full[r] : nullptr);
        return solve(m->ctx[r], k0, k1, &s);
    });
    m->elapsed_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

}  // namespace

void ah_release_nccl(ah_ctx *ctx) {
    if (!ctx || !ctx->nccl_comm) return;
    if (NcclApi *nc = nccl_api()) nc->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

extern "C" {

void ah_k_partition(int64_t nk, int nparts, int r, int64_t *k0, int64_t *k1) {
    if (!k0 || !k1) return;
    if (nparts < 1 || r < 0 || r >= nparts || nk < 0) { *k0 = 1; *k1 = 0; return; }
    k_partition_host(nk, nparts, r, k0, k1);
}

int ah_multi_create(ah_multi **pm, const int *devices, int ndev) {
    if (!pm) return AH_ERR_INVALID_ARG;
    *pm = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        cudaGetLastError();
        return fail(nullptr, AH_ERR_NO_DEVICE, "no CUDA device available (libarrowhead_b200 has no CPU fallback)");
    }
    if (ndev <= 0) ndev = have;
    ah_multi *m = new (std::nothrow) ah_multi();
    if (!m) return fail(nullptr, AH_ERR_NOMEM, "out of host memory");
    m->ndev = ndev;
    for (int r = 0; r < ndev; ++r) {
        const int dev = devices ? devices[r] : r;
        ah_ctx *c = nullptr;
        const int rc = ah_create(&c, dev);
        if (rc) { ah_multi_destroy(m); return rc; }     // g_create_error holds the message
        m->devices.push_back(dev);
        m->ctx.push_back(c);
    }
    for (int r = 0; r < ndev; ++r) {
        Worker *w = new Worker();
        m->workers.push_back(w);
        w->th = std::thread(worker_loop, w);
    }
    *pm = m;
    return AH_OK;
}

void ah_multi_destroy(ah_multi *m) {
    if (!m) return;
    for (Worker *w : m->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    if (!m->comms.empty()) {
        NcclApi *nc = nccl_api();
        if (nc) for (ncclComm_t c : m->comms) nc->CommDestroy(c);
    }
    for (ah_ctx *c : m->ctx) ah_destroy(c);
    delete m;
}

int ah_multi_ndev(const ah_multi *m) { return m ? m->ndev : 0; }
ah_ctx *ah_multi_ctx(ah_multi *m, int r) { return (m && r >= 0 && r < m->ndev) ? m->ctx[r] : nullptr; }
const char *ah_multi_last_error(const ah_multi *m) { return m ? m->err.c_str() : g_create_error.c_str(); }
double ah_multi_elapsed_ms(const ah_multi *m) { return m ? m->elapsed_ms : 0.0; }

int ah_multi_symarrow_eigen(ah_multi *m, int64_t N, const double *D, const double *z, double a, const double *tau,
                            ah_result *out, double *const *U_full) {
    return multi_solve(m, N + 1, out, U_full, nullptr, [&](ah_ctx *c, int64_t k0, int64_t k1, ah_result *s) {
        return ah_symarrow_eigen(c, N, D, z, a, tau, k0, k1, s);
    });
}
int ah_multi_symdpr1_eigen(ah_multi *m, int64_t n, const double *D, const double *u, double rho, const double *tau,
                           ah_result *out, double *const *U_full) {
    return multi_solve(m, n, out, U_full, nullptr, [&](ah_ctx *c, int64_t k0, int64_t k1, ah_result *s) {
        return ah_symdpr1_eigen(c, n, D, u, rho, tau, k0, k1, s);
    });
}
int ah_multi_halfarrow_svd(ah_multi *m, int64_t nd, int64_t nz, const double *D, const double *z, const double *tau,
                           ah_result *out, double *const *U_full, double *const *V_full) {
    return multi_solve(m, nz, out, U_full, V_full, [&](ah_ctx *c, int64_t k0, int64_t k1, ah_result *s) {
        return ah_halfarrow_svd(c, nd, nz, D, z, tau, k0, k1, s);
    });
}

int ah_multi_allgather_columns(ah_multi *m, double *const *U_full, int64_t ld, int64_t kmax, double *elapsed_ms) {
    if (!m) return AH_ERR_INVALID_ARG;
    if (!U_full || ld < 1 || kmax < 1) { m->err = "ah_multi_allgather_columns: bad arguments"; return AH_ERR_INVALID_ARG; }
    NcclApi *nc = nccl_api();
    if (!nc) { m->err = "NCCL is not available: libnccl.so.2 could not be loaded"; return AH_ERR_NO_DEVICE; }
    if (m->comms.empty()) {
        m->comms.resize(m->ndev);
        ncclResult_t r = nc->CommInitAll(m->comms.data(), m->ndev, m->devices.data());
        if (r != ncclSuccess) { m->comms.clear(); m->err = std::string("ncclCommInitAll: ") + nc->GetErrorString(r); return AH_ERR_CUDA_BASE; }
    }
    std::vector<int> ranks(m->ndev);
    std::vector<cudaStream_t> streams(m->ndev);
    std::vector<cudaEvent_t> e0(m->ndev), e1(m->ndev);
    for (int r = 0; r < m->ndev; ++r) {
        ranks[r] = r; streams[r] = m->ctx[r]->stream;
        cudaSetDevice(m->devices[r]);
        cudaEventCreate(&e0[r]); cudaEventCreate(&e1[r]);
        cudaEventRecord(e0[r], streams[r]);
    }
    int rc = nccl_gather_blocks(nc, m->ndev, ranks.data(), m->ndev, U_full, m->comms.data(), streams.data(), m->devices.data(), ld, kmax, m->err);
    double worst = 0.0;
    for (int r = 0; r < m->ndev; ++r) {
        cudaSetDevice(m->devices[r]);
        cudaEventRecord(e1[r], streams[r]);
        if (cudaStreamSynchronize(streams[r]) != cudaSuccess && rc == AH_OK) { m->err = "all-gather: stream synchronize failed"; rc = AH_ERR_CUDA_BASE; cudaGetLastError(); }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0[r], e1[r]) == cudaSuccess) worst = std::max(worst, (double)ms);
        cudaEventDestroy(e0[r]); cudaEventDestroy(e1[r]);
    }
    m->gather_ms = worst;
    if (elapsed_ms) *elapsed_ms = worst;
    return rc;
}

// ---- one process per GPU (torchrun-style): the caller carries the 128-byte id from rank 0 to the others --------------
int ah_nccl_unique_id(void *id128) {
    if (!id128) return AH_ERR_INVALID_ARG;
    NcclApi *nc = nccl_api();
    if (!nc) return fail(nullptr, AH_ERR_NO_DEVICE, "NCCL is not available: libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = nc->GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128));
    if (r != ncclSuccess) return fail(nullptr, AH_ERR_CUDA_BASE, std::string("ncclGetUniqueId: ") + nc->GetErrorString(r));
    return AH_OK;
}

int ah_nccl_init(ah_ctx *ctx, int nranks, int rank, const void *id128) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, AH_ERR_INVALID_ARG, "ah_nccl_init: bad arguments");
    NcclApi *nc = nccl_api();
    if (!nc) return fail(ctx, AH_ERR_NO_DEVICE, "NCCL is not available: libnccl.so.2 could not be loaded");
    AH_CUDA(cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) { nc->CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    ncclResult_t r = nc->CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return fail(ctx, AH_ERR_CUDA_BASE, std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
    ctx->nccl_comm = comm; ctx->nccl_rank = rank; ctx->nccl_nranks = nranks;
    return AH_OK;
}

int ah_allgather_columns(ah_ctx *ctx, double *U_full, int64_t ld, int64_t kmax, double *elapsed_ms) {
    if (!ctx) return AH_ERR_INVALID_ARG;
    if (!U_full || ld < 1 || kmax < 1) return fail(ctx, AH_ERR_INVALID_ARG, "ah_allgather_columns: bad arguments");
    if (!ctx->nccl_comm) return fail(ctx, AH_ERR_INVALID_ARG, "ah_allgather_columns: call ah_nccl_init first");
    NcclApi *nc = nccl_api();
    AH_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    AH_CUDA(cudaEventRecord(ctx->ev[0], st));
    int rc = nccl_gather_blocks(nc, ctx->nccl_nranks, &ctx->nccl_rank, 1, &U_full, &comm, &st, &ctx->device, ld, kmax, ctx->err);
    if (rc) return rc;
    AH_CUDA(cudaEventRecord(ctx->ev[1], st));
    AH_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    AH_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    if (elapsed_ms) *elapsed_ms = ms;
    return AH_OK;
}

}  // extern "C"
