// ah_fast.cuh -- the fast path: standard (no-remedy, binary64) eigenpairs at FP64-pipe speed.
//
// Persistent kernel, one CTA per SM.  The two input vectors D (poles) and w (z | u | z.*D) are
// streamed over and over through a shared-memory ring by 1-D TMA bulk copies (cp.async.bulk +
// mbarrier), one tile = FTL consecutive poles.  Every pass of the ring over l = 0..N-1 is one "sweep";
// every O(N) loop of the reference's per-k solver is one sweep with its own per-term formula:
//
//   SHIFT   sum_l w_l^2 / ((D_l - D_k) - mid)                 arrowhead3.jl:437-441
//   INV     P,Q, sum|z'|, max(D'+|z'|), max(|D'|+|z'|)        arrowhead3.jl:152-189 + bracket :671-680 + nu_1 :460-464
//   BISECT  sum_l z'_l^2/(D'_l - m), ~54 times per eigenpair  arrowhead3.jl:686-704      <- ~95 % of all work
//   VNORM   sum_l v_l^2,  v_l = w_l/((D_l - sigma) - mu)      arrowhead3.jl:517-522
//   VWRITE  v_l / ||v||  -> U[:,k]  (recomputed, not re-read)  arrowhead3.jl:522, 643
//
// A CTA works on FJ eigenpairs ("slots") at once.  Each thread copies its FE elements of the current
// tile from shared memory into registers ONCE and evaluates the BISECT term for every bisecting slot
// from those registers, so shared-memory traffic per term drops by the number of bisecting slots and the
// FP64 pipe is the only saturated unit.  The inverse matrix (D', z') is never materialised: the
// BISECT term is evaluated on the original data as  w_l^2 / (d_l (1 - m d_l)),  d_l = D_l - sigma,
// with a MUFU.RCP64H seed + one cubic correction (ah_device.cuh), 7 FP64-pipe instructions per term.
// The other four sweeps use IEEE divisions in exactly the reference's operation order and need several
// accumulators; they run through one shared "aux" channel that a single slot owns at a time.
//
// All 8 warps split each tile, so an eigenpair's sweep is a CTA-wide reduction (fixed order:
// thread-sequential over its elements, xor-shuffle tree, warps 0..7) finished by warp 0, whose lane j
// owns slot j's scalar state machine.  Anything non-standard (double-double needed, Knu > tolnu ->
// Remedy 3, Remedy 1, |nu| = Inf) is appended to a list and solved by k_full (ah_full.cuh).
#pragma once
#include "ah_device.cuh"
#include "ah_full.cuh"

namespace ah {

constexpr int FT = 256;          // threads per CTA
constexpr int FW = FT / 32;      // warps
constexpr int FJ = 4;            // eigenpair slots per CTA
constexpr int FE = 8;            // elements per thread per tile
constexpr int FTL = FT * FE;     // poles per tile (2048)
constexpr int FS = 3;            // ring stages
constexpr int NAUX = 5;          // aux accumulators

enum : int { PH_FREE = 0, PH_WANT_START, PH_AUX, PH_BISECT, PH_WANT_VEC, PH_DONE };
enum : int { AUX_NONE = 0, AUX_SHIFT, AUX_INV, AUX_VNORM, AUX_VWRITE };

struct SlotState {   // cold per-slot state, touched by warp 0 only
    int64_t k;
    int64_t i;       // shift index (1-based)
    int phase;
    int sideR;
    int steps;
    int pad_;
    double sigma, wz, wz2, b;
    double left, right, middle;
    double nu1, Kb, Kz;
    double mu, lambda, inrm, nu;
};
struct HotState {    // what every thread needs during a sweep
    double sigma, m;
    int act, tile_i, owner, pad_;
};
struct AuxState {
    int owner, phase;        // slot index or -1
    int ii, pad_;            // 0-based skip index for INV
    double p0, p1, p2, p3;   // SHIFT: Dk, mid | INV: sigma, wz, sgn | VNORM/VWRITE: sigma, mu, inrm, lambda
    double *ucol, *vcol;
};
struct FastShared {
    alignas(128) double tileD[FS][FTL];
    alignas(128) double tileW[FS][FTL];
    alignas(8) unsigned long long full[FS];
    alignas(8) unsigned long long empty[FS];
    double part[FJ][FW];
    double apart[NAUX][FW];
    SlotState slot[FJ];
    HotState hot[FJ];
    AuxState aux;
    int alive;
};

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// element e (0..FE-1) of thread t in a tile sits at offset (e>>1)*2*FT + 2*t + (e&1):
// consecutive lanes read consecutive 16-byte pairs -> conflict-free LDS.128 and coalesced STG.128.
__device__ __forceinline__ int elem_offset(int t, int e) { return (e >> 1) * (2 * FT) + 2 * t + (e & 1); }

template <int KIND>
__device__ __forceinline__ void fast_outputs(const Outputs &O, const SlotState &S, int64_t col, double knu, int status) {
    O.lambda[col] = S.lambda;
    O.sind[col] = S.i;
    O.kb[col] = S.Kb; O.kz[col] = S.Kz; O.knu[col] = knu; O.krho[col] = 0.0;
    O.qout[col] = 0;
    O.steps[col] = S.steps;
    O.status[col] = status;
}

// D'_l = 1/d_l for a single pole (bracket ends are the neighbours of the shift, no reduction needed)
template <int KIND>
__device__ __forceinline__ double inv_pole(const double *D, int64_t l, double sigma) {
    return __ddiv_rn(1.0, dshift<KIND>(D[l], sigma));
}

template <int KIND>
__global__ void __launch_bounds__(FT, 1)
k_fast(Problem P, Outputs O, int64_t kbase, int64_t cnt, int64_t *__restrict__ dlist, unsigned long long *dcount) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FastShared &sh = *reinterpret_cast<FastShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t N = P.N;
    const int ntiles = (int)((N + FTL - 1) / FTL);
    const bool want_vec = (O.U != nullptr) || (KIND == KIND_HALF && O.V != nullptr);

    if (tid == 0) {
        for (int s = 0; s < FS; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], FW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sh.aux.owner = -1; sh.aux.phase = AUX_NONE;
        sh.alive = 1;
    }
    if (tid < FJ) { sh.slot[tid].phase = PH_FREE; sh.hot[tid].act = 0; }
    __syncthreads();

    long long g_issue = 0;   // tiles issued (thread 0 only)
    long long g = 0;         // tiles consumed
    auto issue_tile = [&](long long gi) {
        const int s = (int)(gi % FS);
        const long long use = gi / FS;
        if (use > 0) mbar_wait(&sh.empty[s], (uint32_t)((use - 1) & 1));
        const int tile = (int)(gi % ntiles);
        mbar_arrive_expect_tx(&sh.full[s], 2u * FTL * 8u);
        bulk_g2s(&sh.tileD[s][0], P.D + (int64_t)tile * FTL, FTL * 8u, &sh.full[s]);
        bulk_g2s(&sh.tileW[s][0], P.w + (int64_t)tile * FTL, FTL * 8u, &sh.full[s]);
    };

    bool first = true;
    while (true) {
        // ================= sweep end / start: warp 0 runs the scalar state machines =================
        if (warp == 0) {
            const int j = lane;
            bool want_aux = false;
            if (j < FJ) {
                SlotState &S = sh.slot[j];
                // ---- (1) bisection update
                if (!first && S.phase == PH_BISECT) {
                    double sum = sh.part[j][0];
                    for (int w = 1; w < FW; ++w) sum = __dadd_rn(sum, sh.part[j][w]);
                    sum = __dmul_rn(sum, S.wz2);
                    if (KIND != KIND_DPR1) sum = __dadd_rn(sum, __ddiv_rn(S.wz2, __dadd_rn(0.0, -S.middle)));
                    const double F = __dadd_rn(__dadd_rn(S.b, -S.middle), -sum);
                    if (F > 0.0) S.left = S.middle; else S.right = S.middle;
                    S.middle = (S.left + S.right) / 2.0;
                    S.steps += 1;
                }
                // ---- (2) aux completion for the owner
                if (!first && sh.aux.owner == j && S.phase == PH_AUX) {
                    double a[NAUX];
                    for (int q = 0; q < NAUX; ++q) {
                        double v = sh.apart[q][0];
                        for (int w = 1; w < FW; ++w) v = (q < 3) ? __dadd_rn(v, sh.apart[q][w]) : fmax(v, sh.apart[q][w]);
                        a[q] = v;
                    }
                    const int ap = sh.aux.phase;
                    if (ap == AUX_SHIFT) {
                        const int64_t k = S.k;
                        const double Dk = P.D[k - 1], mid = sh.aux.p1;
                        bool right;
                        if (KIND == KIND_ARROW) right = __dadd_rn(__dadd_rn(__dadd_rn(P.a, -Dk), -mid), -a[0]) < 0.0;
                        else if (KIND == KIND_HALF) right = __dadd_rn(__dadd_rn(__dadd_rn(P.zz, -__dmul_rn(Dk, Dk)), -mid), -a[0]) < 0.0;
                        else right = __dadd_rn(1.0, __dmul_rn(P.a, a[0])) > 0.0;
                        if (right) { S.i = k; S.sideR = 1; } else { S.i = k - 1; S.sideR = 0; }
                        // continue with INV, keeping the aux channel
                        const int64_t ii = S.i - 1;
                        S.sigma = P.D[ii];
                        S.wz = __ddiv_rn(1.0, P.w[ii]);
                        sh.aux.phase = AUX_INV; sh.aux.ii = (int)ii;
                        sh.aux.p0 = S.sigma; sh.aux.p1 = S.wz; sh.aux.p2 = S.sideR ? 1.0 : -1.0;
                    } else if (ap == AUX_INV) {
                        const int64_t ii = S.i - 1;
                        const double wz = S.wz, sigma = S.sigma;
                        double Ps = a[0], Qs = a[1], sabs = a[2], bnd = a[3], nu1m = a[4];
                        // neighbours of the shift give max D' / min D' (D strictly decreasing)
                        double mxD, mnD;
                        if (KIND == KIND_DPR1) {
                            mxD = ii > 0 ? inv_pole<KIND>(P.D, ii - 1, sigma) : inv_pole<KIND>(P.D, N - 1, sigma);
                            mnD = ii < N - 1 ? inv_pole<KIND>(P.D, ii + 1, sigma) : inv_pole<KIND>(P.D, 0, sigma);
                        } else {
                            mxD = ii > 0 ? inv_pole<KIND>(P.D, ii - 1, sigma) : 0.0;
                            mnD = ii < N - 1 ? inv_pole<KIND>(P.D, ii + 1, sigma) : 0.0;
                            const double az = fabs(wz);
                            sabs = __dadd_rn(sabs, az);
                            bnd = fmax(bnd, az);
                            nu1m = fmax(nu1m, az);
                        }
                        if (KIND == KIND_ARROW) {
                            double as_ = __dadd_rn(P.a, -sigma);
                            if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
                        } else if (KIND == KIND_HALF) {
                            dd ad = dd_add(dd{P.zz_hi, P.zz_lo}, dd_mul(dd_from(sigma), dd_from(-sigma)));
                            double as_ = __dadd_rn(ad.hi, ad.lo);
                            if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
                        } else {
                            double ir = __ddiv_rn(1.0, P.a);
                            if (P.a > 0) Ps = __dadd_rn(Ps, ir); else Qs = __dadd_rn(Qs, ir);
                        }
                        const double PQ = __dadd_rn(Ps, Qs);
                        S.Kb = __ddiv_rn(__dadd_rn(Ps, -Qs), fabs(PQ));
                        if (KIND == KIND_DPR1) S.Kz = __dmul_rn(__dadd_rn(P.sumabs_w, -fabs(P.w[ii])), fabs(wz));
                        else S.Kz = __dmul_rn(P.maxabs_w, fabs(wz));
                        sh.aux.owner = -1; sh.aux.phase = AUX_NONE;
                        if (S.Kb < P.tau[0] || S.Kz < P.tau[1]) {
                            S.b = __dmul_rn(__dmul_rn(PQ, wz), wz);
                            S.wz2 = __dmul_rn(wz, wz);
                            S.nu1 = fmax(__dadd_rn(sabs, fabs(S.b)), nu1m);
                            if (S.sideR) { S.left = mxD; S.right = fmax(bnd, __dadd_rn(S.b, sabs)); }
                            else { S.left = fmin(-bnd, __dadd_rn(S.b, -sabs)); S.right = mnD; }
                            S.middle = (S.left + S.right) / 2.0;
                            S.steps = 0;
                            S.phase = PH_BISECT;
                        } else {   // double-double (or BigFloat) needed: full kernel
                            dlist[atomicAdd(dcount, 1ull)] = S.k;
                            S.phase = PH_FREE;
                        }
                    } else if (ap == AUX_VNORM) {
                        double ss = a[0];
                        if (KIND != KIND_DPR1) ss = __dadd_rn(ss, 1.0);
                        S.inrm = __ddiv_rn(1.0, sqrt(ss));
                        sh.aux.phase = AUX_VWRITE;
                        sh.aux.p2 = S.inrm;
                    } else if (ap == AUX_VWRITE) {
                        fast_outputs<KIND>(O, S, S.k - kbase, __ddiv_rn(S.nu1, fabs(S.nu)), 0);
                        sh.aux.owner = -1; sh.aux.phase = AUX_NONE;
                        S.phase = PH_FREE;
                    }
                }
                // ---- (1b) bisection convergence test (also right after INV: the reference tests first)
                if (S.phase == PH_BISECT &&
                    !((S.right - S.left) > 2.0 * kEps * fmax(fabs(S.left), fabs(S.right)))) {
                    const double nu = S.right;
                    const int64_t ii = S.i - 1;
                    const double Knu = __ddiv_rn(S.nu1, fabs(nu));
                    bool bail = !isfinite(nu);
                    if (KIND == KIND_HALF) bail = bail || !(Knu < P.tau[2]);
                    else bail = bail || (Knu > P.tau[2]);
                    double mu = 0.0, lambda = 0.0;
                    if (!bail) {
                        mu = __ddiv_rn(1.0, nu);
                        const double Di = P.D[ii];
                        if (KIND == KIND_HALF) {
                            lambda = sqrt(__dadd_rn(mu, __dmul_rn(S.sigma, S.sigma)));
                            bail = (S.k == N + 1) &&
                                   (__ddiv_rn(__dadd_rn(__dmul_rn(Di, Di), fabs(mu)), fabs(__dmul_rn(lambda, lambda))) > P.tau[4]);
                        } else {
                            lambda = __dadd_rn(mu, S.sigma);
                            if (__ddiv_rn(__dadd_rn(fabs(Di), fabs(mu)), fabs(lambda)) > P.tau[4]) {
                                const int64_t k = S.k, i = S.i;
                                bool c1 = (k == 1 && P.D[0] < 0.0);
                                bool c2 = (KIND == KIND_ARROW) && (k == N + 1 && P.D[N - 1] > 0.0);
                                bool c3 = (KIND == KIND_ARROW ? i < N : true) && !S.sideR &&
                                          sign_jl(Di) + sign_jl(P.D[ii + 1 < N ? ii + 1 : ii]) == 0.0;
                                bool c4 = (i > 1 && S.sideR && sign_jl(Di) + sign_jl(P.D[ii > 0 ? ii - 1 : 0]) == 0.0);
                                bail = c1 || c2 || c3 || c4;
                            }
                        }
                    }
                    if (bail) {
                        dlist[atomicAdd(dcount, 1ull)] = S.k;
                        S.phase = PH_FREE;
                    } else {
                        S.nu = nu; S.mu = mu; S.lambda = lambda;
                        if (want_vec) S.phase = PH_WANT_VEC;
                        else { fast_outputs<KIND>(O, S, S.k - kbase, Knu, 0); S.phase = PH_FREE; }
                    }
                }
                // ---- (3) fetch new work
                if (S.phase == PH_FREE) {
                    const unsigned long long wi = atomicAdd(dcount + 1, 1ull);
                    if ((int64_t)wi < cnt) {
                        S.k = kbase + (int64_t)wi;
                        S.steps = 0;
                        S.phase = PH_WANT_START;
                    } else {
                        S.phase = PH_DONE;
                    }
                }
                want_aux = (S.phase == PH_WANT_START) || (S.phase == PH_WANT_VEC);
            }
            // ---- (4) aux arbitration: lowest waiting slot gets the free channel
            __syncwarp();
            const unsigned wants = __ballot_sync(0xffffffffu, want_aux);
            const int aux_owner_now = sh.aux.owner;
            __syncwarp();
            if (aux_owner_now < 0 && wants) {
                const int win = __ffs(wants) - 1;
                if (j == win) {
                    SlotState &S = sh.slot[j];
                    sh.aux.owner = j;
                    if (S.phase == PH_WANT_START) {
                        const int64_t k = S.k;
                        const bool ext_first = (k == 1);
                        const bool ext_last = (KIND != KIND_DPR1) && (k == N + 1);
                        if (ext_first || ext_last) {
                            S.i = ext_first ? 1 : N;
                            S.sideR = ext_first ? 1 : 0;
                            const int64_t ii = S.i - 1;
                            S.sigma = P.D[ii];
                            S.wz = __ddiv_rn(1.0, P.w[ii]);
                            sh.aux.phase = AUX_INV; sh.aux.ii = (int)ii;
                            sh.aux.p0 = S.sigma; sh.aux.p1 = S.wz; sh.aux.p2 = S.sideR ? 1.0 : -1.0;
                        } else {
                            const double Dk = P.D[k - 1];
                            sh.aux.phase = AUX_SHIFT; sh.aux.ii = -1;
                            sh.aux.p0 = Dk;
                            sh.aux.p1 = dshift<KIND>(P.D[k - 2], Dk) / 2.0;
                        }
                    } else {   // PH_WANT_VEC
                        const int64_t col = S.k - kbase;
                        const int64_t ucolidx = O.rowmap_u ? O.rowmap_u[S.k - 1] : col;
                        sh.aux.phase = AUX_VNORM; sh.aux.ii = -1;
                        sh.aux.p0 = S.sigma; sh.aux.p1 = S.mu; sh.aux.p2 = 0.0; sh.aux.p3 = S.lambda;
                        sh.aux.ucol = O.U ? O.U + ucolidx * O.ldu : nullptr;
                        sh.aux.vcol = (KIND == KIND_HALF && O.V) ? O.V + ucolidx * O.ldv : nullptr;
                    }
                    S.phase = PH_AUX;
                }
            }
            __syncwarp();
            // ---- publish hot state
            bool alive = false;
            if (j < FJ) {
                const SlotState &S = sh.slot[j];
                HotState &H = sh.hot[j];
                H.act = (S.phase == PH_BISECT);
                H.sigma = S.sigma; H.m = S.middle;
                const int64_t ii = S.i - 1;
                H.tile_i = (int)(ii / FTL);
                H.owner = (int)(((ii % FTL) % (2 * FT)) >> 1);
                alive = S.phase != PH_DONE;
            }
            const unsigned al = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) sh.alive = al != 0;
        }
        first = false;
        __syncthreads();
        if (!sh.alive) break;

        // ================= per-thread copies of the sweep parameters =================
        double sig[FJ], mm[FJ];
        bool act[FJ];
        int tile_i[FJ];
        bool own[FJ];
#pragma unroll
        for (int j = 0; j < FJ; ++j) {
            sig[j] = sh.hot[j].sigma; mm[j] = sh.hot[j].m; act[j] = sh.hot[j].act != 0;
            tile_i[j] = sh.hot[j].tile_i; own[j] = sh.hot[j].owner == tid;
        }
        const int aphase = sh.aux.phase;
        const int aux_ii = sh.aux.ii;
        const double ap0 = sh.aux.p0, ap1 = sh.aux.p1, ap2 = sh.aux.p2, ap3 = sh.aux.p3;
        double *const aucol = sh.aux.ucol;
        double *const avcol = sh.aux.vcol;
        double acc[FJ];
#pragma unroll
        for (int j = 0; j < FJ; ++j) acc[j] = 0.0;
        double xa0 = 0.0, xa1 = 0.0, xa2 = 0.0, xa3 = -INFINITY, xa4 = 0.0;

        // ================= the sweep =================
        for (int tile = 0; tile < ntiles; ++tile, ++g) {
            if (tid == 0) {
                if (g == 0) { for (; g_issue < FS - 1; ++g_issue) issue_tile(g_issue); }
                issue_tile(g_issue); ++g_issue;
            }
            const int s = (int)(g % FS);
            mbar_wait(&sh.full[s], (uint32_t)((g / FS) & 1));
            double Dv[FE], Wv[FE], W2[FE];
#pragma unroll
            for (int e = 0; e < FE; e += 2) {
                const double2 d2 = *reinterpret_cast<const double2 *>(&sh.tileD[s][elem_offset(tid, e)]);
                const double2 w2 = *reinterpret_cast<const double2 *>(&sh.tileW[s][elem_offset(tid, e)]);
                Dv[e] = d2.x; Dv[e + 1] = d2.y; Wv[e] = w2.x; Wv[e + 1] = w2.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[s]);
#pragma unroll
            for (int e = 0; e < FE; ++e) W2[e] = __dmul_rn(Wv[e], Wv[e]);
            const int64_t base = (int64_t)tile * FTL;

            // ---- aux channel (at most one slot; exact reference arithmetic)
            if (aphase == AUX_SHIFT) {
#pragma unroll
                for (int e = 0; e < FE; ++e) {
                    const int64_t l = base + elem_offset(tid, e);
                    if (l < N) xa0 = __dadd_rn(xa0, __ddiv_rn(W2[e], __dadd_rn(dshift<KIND>(Dv[e], ap0), -ap1)));
                }
            } else if (aphase == AUX_INV) {
#pragma unroll
                for (int e = 0; e < FE; ++e) {
                    const int64_t l = base + elem_offset(tid, e);
                    if (l < N && l != aux_ii) {
                        const double Dp = __ddiv_rn(1.0, dshift<KIND>(Dv[e], ap0));
                        const double zp = __dmul_rn(__dmul_rn(-Wv[e], Dp), ap1);
                        const double t = __dmul_rn(W2[e], Dp);
                        if (Dp > 0.0) xa0 = __dadd_rn(xa0, t); else xa1 = __dadd_rn(xa1, t);
                        const double az = fabs(zp);
                        xa2 = __dadd_rn(xa2, az);
                        xa3 = fmax(xa3, __dadd_rn(__dmul_rn(ap2, Dp), az));
                        xa4 = fmax(xa4, __dadd_rn(fabs(Dp), az));
                    }
                }
            } else if (aphase == AUX_VNORM) {
#pragma unroll
                for (int e = 0; e < FE; ++e) {
                    const int64_t l = base + elem_offset(tid, e);
                    if (l < N) {
                        const double v = __ddiv_rn(Wv[e], __dadd_rn(dshift<KIND>(Dv[e], ap0), -ap1));
                        xa0 = __dadd_rn(xa0, __dmul_rn(v, v));
                    }
                }
            } else if (aphase == AUX_VWRITE) {
                double *const vec = (KIND == KIND_HALF) ? avcol : aucol;
                const int64_t *vmap = (KIND == KIND_HALF) ? O.rowmap_v : O.rowmap_u;
#pragma unroll
                for (int e = 0; e < FE; e += 2) {
                    const int64_t l = base + elem_offset(tid, e);
                    double v[2], vo[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        v[h] = __dmul_rn(__ddiv_rn(Wv[e + h], __dadd_rn(dshift<KIND>(Dv[e + h], ap0), -ap1)), ap2);
                        vo[h] = v[h];
                        if (KIND == KIND_HALF && O.rowscale_v && l + h < N) vo[h] = __dmul_rn(v[h], O.rowscale_v[l + h]);
                    }
                    if (vec) {
                        if (!vmap && l + 1 < N && ((reinterpret_cast<uintptr_t>(vec + l) & 15) == 0)) {
                            *reinterpret_cast<double2 *>(vec + l) = make_double2(vo[0], vo[1]);
                        } else {
                            if (l < N) vec[vmap ? vmap[l] : l] = vo[0];
                            if (l + 1 < N) vec[vmap ? vmap[l + 1] : l + 1] = vo[1];
                        }
                    }
                    if (KIND == KIND_HALF && aucol) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            if (l + h < N) aucol[O.rowmap_u ? O.rowmap_u[l + h] : l + h] = __ddiv_rn(__dmul_rn(ap3, v[h]), Dv[e + h]);
                    }
                }
                if (KIND != KIND_DPR1 && tile == ntiles - 1 && tid == 0) {
                    const double vt = __dmul_rn(-1.0, ap2);
                    if (vec) vec[vmap ? vmap[N] : N] = vt;
                    if (KIND == KIND_HALF && aucol && P.has_tip) aucol[O.rowmap_u ? O.rowmap_u[N] : N] = __ddiv_rn(__dmul_rn(P.ztip, vt), ap3);
                }
            }

            // ---- bisection terms for every bisecting slot, from registers
#pragma unroll
            for (int j = 0; j < FJ; ++j) {
                if (act[j]) {
                    if (tile == tile_i[j] && own[j]) {
                        // this thread holds the shift pole itself: skip it
                        const int skip = (int)((sh.slot[j].i - 1) - base);
#pragma unroll
                        for (int e = 0; e < FE; ++e)
                            if (elem_offset(tid, e) != skip)
                                acc[j] = secular_term_acc(acc[j], dshift<KIND>(Dv[e], sig[j]), mm[j], W2[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < FE; ++e)
                            acc[j] = secular_term_acc(acc[j], dshift<KIND>(Dv[e], sig[j]), mm[j], W2[e]);
                    }
                }
            }
        }

        // ================= reduce this sweep's partial sums =================
#pragma unroll
        for (int j = 0; j < FJ; ++j) {
            if (act[j]) {
                const double v = warp_allreduce(acc[j], OpSum());
                if (lane == 0) sh.part[j][warp] = v;
            }
        }
        if (aphase == AUX_SHIFT || aphase == AUX_INV || aphase == AUX_VNORM) {
            xa0 = warp_allreduce(xa0, OpSum());
            xa1 = warp_allreduce(xa1, OpSum());
            xa2 = warp_allreduce(xa2, OpSum());
            xa3 = warp_allreduce(xa3, OpMax());
            xa4 = warp_allreduce(xa4, OpMax());
            if (lane == 0) {
                sh.apart[0][warp] = xa0; sh.apart[1][warp] = xa1; sh.apart[2][warp] = xa2;
                sh.apart[3][warp] = xa3; sh.apart[4][warp] = xa4;
            }
        }
        __syncthreads();
    }

    // drain the prefetched tiles before the CTA (and its shared memory) goes away
    if (tid == 0) {
        for (long long gi = g; gi < g_issue; ++gi) mbar_wait(&sh.full[(int)(gi % FS)], (uint32_t)((gi / FS) & 1));
    }
}

inline int configure_fast_kernels() {
    const int smem = (int)sizeof(FastShared);
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_fast<KIND_ARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_fast<KIND_DPR1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(k_fast<KIND_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return (int)e;
    return 0;
}

// P.D / P.w must be padded to a multiple of FTL (see pad_poles in ah_api.cu).
inline int launch_fast(int kind, const Problem &P, const Outputs &O, int64_t kbase, int64_t cnt, int64_t *dlist,
                       unsigned long long *dcount, int sm_count, cudaStream_t st) {
    const int smem = (int)sizeof(FastShared);
    // one CTA per SM, but never more CTAs than there are groups of FJ eigenpairs
    int64_t want = (cnt + FJ - 1) / FJ;
    const int grid = (int)(want < sm_count ? (want < 1 ? 1 : want) : sm_count);
    switch (kind) {
        case KIND_ARROW: k_fast<KIND_ARROW><<<grid, FT, smem, st>>>(P, O, kbase, cnt, dlist, dcount); break;
        case KIND_DPR1: k_fast<KIND_DPR1><<<grid, FT, smem, st>>>(P, O, kbase, cnt, dlist, dcount); break;
        default: k_fast<KIND_HALF><<<grid, FT, smem, st>>>(P, O, kbase, cnt, dlist, dcount); break;
    }
    return (int)cudaGetLastError();
}

}  // namespace ah
