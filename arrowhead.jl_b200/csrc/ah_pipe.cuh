// ah_pipe.cuh -- the fast path: binary64 eigenpairs (standard ones and Remedy 3) as a pipeline of kernels.
//
// Every O(N) loop of the reference's per-k solver is a "sweep" over the poles l = 0..N-1:
//
//   k_prep    SHIFT   sum_l w_l^2 / ((D_l - D_k) - mid)                 arrowhead3.jl:437-441
//             INV     P,Q, sum|z'|, max(D'+|z'|), max(|D'|+|z'|)        arrowhead3.jl:152-189, brackets :671-680, nu_1 :460-464
//   k_bisect  BISECT  sum_l z'_l^2/(D'_l - m), ~54 sweeps per eigenpair  arrowhead3.jl:686-704      <- ~90 % of all work
//   k_vec     VNORM   sum_l v_l^2,  v_l = w_l/((D_l - sigma) - mu)      arrowhead3.jl:517-522
//             VWRITE  v_l / ||v||  -> U[:,k]  (recomputed, not re-read)  arrowhead3.jl:522, 643
//
// (arrowhead4.jl / arrowhead6.jl are the same loops with the DPR1 / HalfArrow term formulas, selected
// by the KIND template parameter.)  The per-eigenpair scalars travel between the kernels in a 128-byte
// KState record; nothing of size N is ever materialised per eigenpair -- the inverse (D', z') of the
// reference exists only as the formula  z'_l^2/(D'_l - m) = wz^2 * w_l^2 / (d_l (1 - m d_l)),  d_l = D_l - sigma.
//
// k_bisect is the hot kernel: persistent, two CTAs of 8 warps per SM (while one CTA sits at its sweep
// boundary the other keeps the FP64 pipe busy).  D and w^2 stream over and over through a shared-memory
// ring filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier); one tile = 1024 poles.  A CTA bisects BJ = 8 eigenpairs ("slots") at once: each thread moves its 4 poles of the
// tile from shared memory into registers once and evaluates the term for all 8 slots from there, so
// L2 and shared-memory traffic per term drop 8x and the FP64 pipe is the only busy unit.  One term is
// 7.25 FP64-pipe instructions (DADD d, DFMA e=1-m d, DMUL t=d e, MUFU.RCP64H seed + 3 DFMA, DMUL/DFMA into the sum of the
// tile's 4 terms, one DADD per tile into the running sum).  Most of the reference's ~54 bisection steps are never evaluated:
// a sweep at any point yields a certified bound for the computed sums at all grid points on one side of it, so the
// kernel probes near the root (rational model of F), descends the dyadic grid for free wherever a certificate decides,
// and evaluates only the last few cells, where F is rounding noise -- ~14 sweeps, final bracket bit-identical
// ("accelerated bisection", see the comment above k_bisect and DESIGN.md section 3).
// The inner block is branch-free: the one pole per slot that must be skipped (the shift pole, d = 0)
// poisons that slot's accumulator in exactly one thread, which saves the accumulator before the block
// and redoes its four terms without the pole afterwards (a rare, warp-divergent side path).
// A sweep ends with a fixed-order reduction (thread-sequential, xor-shuffle tree, warps 0..7); warp j
// then advances slot j's bisection state, finalises a converged eigenpair (Knu test, Remedy-1 test) and
// takes the next work ticket; towards the end of a launch eigenpairs are time-sliced through a global
// FIFO so that all slots stay busy until (almost) the last sweep (see the comment above k_bisect).
// Knu > tolnu (Remedy 3, ~1 % of the eigenpairs) goes through k_rem3_prep and a second round of k_bisect
// on the re-shifted inverse; anything else that is non-standard (double-double needed, Remedy 1,
// |nu| = Inf, a Remedy 3 beyond its plain branch) is appended to a hand-over list and solved by k_full
// (ah_full.cuh).
#pragma once
#include "ah_device.cuh"
#include "ah_full.cuh"

#include <algorithm>
#include <cstdint>

namespace ah {

enum : int { KS_READY = 0, KS_HANDOFF = 1, KS_DONE = 2, KS_REM3 = 3, KS_REM3_READY = 4, KS_DONE1 = 5 };   // DONE1: finished in round 1
// counters (unsigned long long[16] + unsigned[256] per-SM arrival counts of k_bisect round 0): [0] hand-over list length, [1] k_bisect round-0 ticket counter,
//                                   [2] Remedy-3 list length,   [3] k_bisect round-1 ticket counter,
//                                   [4] round-0 requeue count,  [5] round-0 eigenpairs finished

struct KState {   // 128 bytes per eigenpair
    double sigma, wz, b, left, right, nu1, Kb, Kz;   // written by k_prep
    double nu, mu, lambda, knu;                       // written by k_bisect
    int64_t i;                                        // shift index, 1-based
    int32_t sideR, flag, steps, pad_;
    double krho;                                      // Remedy 3: K_rho of the re-shifted inverse
};
// Remedy-3 round (k_rem3_prep -> k_bisect round 1) re-uses the record: sigma := sigma_1 (HalfArrow: its
// square root; the squared shift is parked in `lambda`), b := rho, left/right := DPR1 bracket, nu1 := new nu_1.
static_assert(sizeof(KState) == 128, "KState layout");
// Certified-bracket state of the accelerated bisection (k_bisect<..., ACCEL>), parked here when an eigenpair is time-sliced
struct KExtra {   // 64 bytes per eigenpair
    double lo, hi;              // certified: every grid midpoint <= lo decides "left", every one >= hi decides "right"
    double hx1, hF1, hx2, hF2;  // the last two evaluated points and their F values (secant on (m - pole) F(m))
    double pole;                // the pole end of the initial bracket
    int8_t nhist, probing, nprobe, nfail;
    int32_t pad_;
};
static_assert(sizeof(KExtra) == 64, "KExtra layout");

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

__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// D'_l = 1/d_l for a single pole (bracket ends are the neighbours of the shift, no reduction needed)
template <int KIND>
__device__ __forceinline__ double inv_pole(const double *D, int64_t l, double sigma) {
    return __ddiv_rn(1.0, dshift<KIND>(D[l], sigma));
}

// =====================================================================================================
// k_prep: shift selection + shifted inverse, PJ consecutive eigenpairs per CTA
constexpr int PT = 256, PW = PT / 32, PJ = 4;

template <int KIND>
__global__ void __launch_bounds__(PT, 2)
k_prep(Problem P, KState *__restrict__ ks, int64_t kbase, int64_t cnt, int64_t *__restrict__ dlist,
       unsigned long long *dcount) {
    __shared__ double red[PJ * 5][PW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t N = P.N, Np = P.Np;
    const double *__restrict__ D = P.D;
    const double *__restrict__ W = P.w;
    const int64_t c0 = (int64_t)blockIdx.x * PJ;

    int64_t kk[PJ];
    bool valid[PJ], ext[PJ];
    double Dk[PJ], mid[PJ];
#pragma unroll
    for (int j = 0; j < PJ; ++j) {
        valid[j] = c0 + j < cnt;
        kk[j] = kbase + (valid[j] ? c0 + j : c0);
        const int64_t k = kk[j];
        ext[j] = (k == 1) || (KIND != KIND_DPR1 && k == N + 1);
        if (!ext[j]) {
            Dk[j] = D[k - 1];
            mid[j] = dshift<KIND>(D[k - 2], Dk[j]) / 2.0;
        } else {   // no shift test for the two exterior eigenvalues; harmless dummy sweep
            Dk[j] = D[0];
            mid[j] = 1.0;
        }
    }

    // ---- SHIFT sweep (pads carry weight 0 and lie below every pole: their terms are exact zeros).
    // Only the sign of F(mid) is used (arrowhead3.jl:441) and the reference forms this sum with Julia's
    // pairwise `sum`, so a <=1-ulp reciprocal (4 FP64-pipe instructions instead of ~10) changes nothing.
    double xs[PJ];
#pragma unroll
    for (int j = 0; j < PJ; ++j) xs[j] = 0.0;
    double2 d2n = ldg2(D + 2 * tid), w2n = ldg2(W + 2 * tid);
    for (int l0 = 2 * tid; l0 < (int)Np; l0 += 2 * PT) {
        const double2 d2 = d2n, w2 = w2n;
        if (l0 + 2 * PT < (int)Np) { d2n = ldg2(D + l0 + 2 * PT); w2n = ldg2(W + l0 + 2 * PT); }   // prefetch the next pair
        const double dv[2] = {d2.x, d2.y};
        const double w2v[2] = {__dmul_rn(w2.x, w2.x), __dmul_rn(w2.y, w2.y)};
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < PJ; ++j)
                xs[j] = __fma_rn(w2v[h], rcp_fast(__dadd_rn(dshift<KIND>(dv[h], Dk[j]), -mid[j])), xs[j]);
    }
#pragma unroll
    for (int j = 0; j < PJ; ++j) {
        const double v = warp_allreduce(xs[j], OpSum());
        if (lane == 0) red[j][warp] = v;
    }
    __syncthreads();
    int64_t ii[PJ];
    bool sideR[PJ];
    double sigma[PJ], wz[PJ], sgn[PJ];
#pragma unroll
    for (int j = 0; j < PJ; ++j) {
        double S = red[j][0];
        for (int w = 1; w < PW; ++w) S = __dadd_rn(S, red[j][w]);
        const int64_t k = kk[j];
        int64_t i;
        if (ext[j]) {
            i = (k == 1) ? 1 : N;
            sideR[j] = (k == 1);
        } else {
            bool right;
            if (KIND == KIND_ARROW) right = __dadd_rn(__dadd_rn(__dadd_rn(P.a, -Dk[j]), -mid[j]), -S) < 0.0;
            else if (KIND == KIND_HALF) right = __dadd_rn(__dadd_rn(__dadd_rn(P.zz, -__dmul_rn(Dk[j], Dk[j])), -mid[j]), -S) < 0.0;
            else right = __dadd_rn(1.0, __dmul_rn(P.a, S)) > 0.0;
            i = right ? k : k - 1;
            sideR[j] = right;
        }
        ii[j] = i - 1;
        sigma[j] = D[ii[j]];
        wz[j] = __ddiv_rn(1.0, W[ii[j]]);
        sgn[j] = sideR[j] ? 1.0 : -1.0;
    }
    __syncthreads();

    // ---- INV sweep: the reference's formulas for every entry of the inverse, in its operation order
    double a0[PJ], a1[PJ], a2[PJ], a3[PJ], a4[PJ];
#pragma unroll
    for (int j = 0; j < PJ; ++j) { a0[j] = 0.0; a1[j] = 0.0; a2[j] = 0.0; a3[j] = -INFINITY; a4[j] = 0.0; }
    d2n = ldg2(D + 2 * tid); w2n = ldg2(W + 2 * tid);
    int iis[PJ];                                           // 32-bit copies: the sweep index fits (N < 2^31 is checked on the host)
#pragma unroll
    for (int j = 0; j < PJ; ++j) iis[j] = (int)ii[j];
    const int Ni = (int)N, Npi = (int)Np;
    for (int l0 = 2 * tid; l0 < Npi; l0 += 2 * PT) {
        const double2 d2 = d2n, w2 = w2n;
        if (l0 + 2 * PT < Npi) { d2n = ldg2(D + l0 + 2 * PT); w2n = ldg2(W + l0 + 2 * PT); }   // prefetch the next pair
        const double dv[2] = {d2.x, d2.y};
        const double wv[2] = {w2.x, w2.y};
        // One thread per eigenpair holds that eigenpair's shift pole (d = 0: not an entry of the inverse), in one iteration.
        // Every other iteration runs without the per-slot test, so that the 8 independent chains of the pair interleave.
        bool hit = false;
#pragma unroll
        for (int j = 0; j < PJ; ++j) hit = hit || (l0 == (iis[j] & ~1));
        auto entry = [&](int h, int j) {
            const double W2 = __dmul_rn(wv[h], wv[h]);
            const double Dp = rcp_fast(dshift<KIND>(dv[h], sigma[j]));   // feeds sums and maxima over N entries only
            const double zp = __dmul_rn(__dmul_rn(-wv[h], Dp), wz[j]);
            const double t = __dmul_rn(W2, Dp);
            if (Dp > 0.0) a0[j] = __dadd_rn(a0[j], t); else a1[j] = __dadd_rn(a1[j], t);
            const double az = fabs(zp);
            a2[j] = __dadd_rn(a2[j], az);
            a3[j] = fmax(a3[j], __dadd_rn(__dmul_rn(sgn[j], Dp), az));
            a4[j] = fmax(a4[j], __dadd_rn(fabs(Dp), az));
        };
        if (!hit && l0 + 1 < Ni) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < PJ; ++j) entry(h, j);
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = l0 + h;
                if (l < Ni) {
#pragma unroll
                    for (int j = 0; j < PJ; ++j)
                        if (l != iis[j]) entry(h, j);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < PJ; ++j) {
        const double v0 = warp_allreduce(a0[j], OpSum()), v1 = warp_allreduce(a1[j], OpSum());
        const double v2 = warp_allreduce(a2[j], OpSum()), v3 = warp_allreduce(a3[j], OpMax());
        const double v4 = warp_allreduce(a4[j], OpMax());
        if (lane == 0) {
            red[j * 5 + 0][warp] = v0; red[j * 5 + 1][warp] = v1; red[j * 5 + 2][warp] = v2;
            red[j * 5 + 3][warp] = v3; red[j * 5 + 4][warp] = v4;
        }
    }
    __syncthreads();
    if (tid < PJ && valid[0]) {
        // one thread per slot finishes the scalars (the per-slot arrays above are fully unrolled
        // registers, so pick this thread's slot with a static loop)
#pragma unroll
        for (int j = 0; j < PJ; ++j) {
            if (j != tid || !valid[j]) continue;
            double Ps = red[j * 5][0], Qs = red[j * 5 + 1][0], sabs = red[j * 5 + 2][0], bnd = red[j * 5 + 3][0],
                   nu1m = red[j * 5 + 4][0];
            for (int w = 1; w < PW; ++w) {
                Ps = __dadd_rn(Ps, red[j * 5][w]); Qs = __dadd_rn(Qs, red[j * 5 + 1][w]);
                sabs = __dadd_rn(sabs, red[j * 5 + 2][w]);
                bnd = fmax(bnd, red[j * 5 + 3][w]); nu1m = fmax(nu1m, red[j * 5 + 4][w]);
            }
            const int64_t i0 = ii[j];
            const double sg = sigma[j], wzj = wz[j];
            // neighbours of the shift give max D' / min D' (D strictly decreasing)
            double mxD, mnD;
            if (KIND == KIND_DPR1) {
                mxD = i0 > 0 ? inv_pole<KIND>(D, i0 - 1, sg) : inv_pole<KIND>(D, N - 1, sg);
                mnD = i0 < N - 1 ? inv_pole<KIND>(D, i0 + 1, sg) : inv_pole<KIND>(D, 0, sg);
            } else {
                mxD = i0 > 0 ? inv_pole<KIND>(D, i0 - 1, sg) : 0.0;
                mnD = i0 < N - 1 ? inv_pole<KIND>(D, i0 + 1, sg) : 0.0;
                const double az = fabs(wzj);   // the slot (0, wz) of the inverse arrow
                sabs = __dadd_rn(sabs, az);
                bnd = fmax(bnd, az);
                nu1m = fmax(nu1m, az);
            }
            if (KIND == KIND_ARROW) {
                const double as_ = __dadd_rn(P.a, -sg);
                if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
            } else if (KIND == KIND_HALF) {
                const dd ad = dd_add(dd{P.zz_hi, P.zz_lo}, dd_mul(dd_from(sg), dd_from(-sg)));
                const double as_ = __dadd_rn(ad.hi, ad.lo);
                if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
            } else {
                const double ir = __ddiv_rn(1.0, P.a);
                if (P.a > 0) Ps = __dadd_rn(Ps, ir); else Qs = __dadd_rn(Qs, ir);
            }
            const double PQ = __dadd_rn(Ps, Qs);
            KState S;
            S.sigma = sg; S.wz = wzj; S.i = i0 + 1; S.sideR = sideR[j] ? 1 : 0;
            S.Kb = __ddiv_rn(__dadd_rn(Ps, -Qs), fabs(PQ));
            if (KIND == KIND_DPR1) S.Kz = __dmul_rn(__dadd_rn(P.sumabs_w, -fabs(W[i0])), fabs(wzj));
            else S.Kz = __dmul_rn(P.maxabs_w, fabs(wzj));
            S.nu = 0.0; S.mu = 0.0; S.lambda = 0.0; S.knu = 0.0; S.steps = 0; S.pad_ = 0; S.krho = 0.0;
            if (S.Kb < P.tau[0] || S.Kz < P.tau[1]) {
                S.b = __dmul_rn(__dmul_rn(PQ, wzj), wzj);
                S.nu1 = fmax(__dadd_rn(sabs, fabs(S.b)), nu1m);
                if (sideR[j]) { S.left = mxD; S.right = fmax(bnd, __dadd_rn(S.b, sabs)); }
                else { S.left = fmin(-bnd, __dadd_rn(S.b, -sabs)); S.right = mnD; }
                S.flag = KS_READY;
            } else {   // double-double (or BigFloat) needed: full kernel
                S.b = 0.0; S.nu1 = 0.0; S.left = 0.0; S.right = 0.0;
                S.flag = KS_HANDOFF;
                dlist[atomicAdd(dcount, 1ull)] = kk[j];
            }
            ks[c0 + j] = S;
        }
    }
}

// =====================================================================================================
// k_bisect
#ifndef AH_BT
#define AH_BT 256
#endif
#ifndef AH_BCTAS
#define AH_BCTAS 2
#endif
constexpr int BT = AH_BT;        // threads per CTA
constexpr int BCTAS = AH_BCTAS;  // CTAs per SM
constexpr int BW = BT / 32;      // warps
constexpr int BJ = 8;            // eigenpair slots per CTA
constexpr int BE = 4;            // poles per thread per tile
constexpr int BTL = BT * BE;     // poles per tile (2048)
#ifndef AH_BS
#define AH_BS 4
#endif
constexpr int BS = AH_BS;        // ring stages
#ifndef AH_R1_CTAS_PER_SM
#define AH_R1_CTAS_PER_SM 1
#endif
#ifndef AH_STAGGER
#define AH_STAGGER 1
#endif
#ifndef AH_STAGGER_NS_PER_TILE
#define AH_STAGGER_NS_PER_TILE 600   // half of the ~1.2 us a tile takes with 8 slots
#endif
#ifndef AH_QSHIFT
#define AH_QSHIFT 2
#endif
#ifndef AH_REM3_DIRECT_MAX
#define AH_REM3_DIRECT_MAX 128
#endif
constexpr int REM3_DIRECT_MAX = AH_REM3_DIRECT_MAX;   // Remedy-3 lists up to this length are bisected by k_rem3_prep itself, one CTA per eigenpair
constexpr int TILE_POLES = 2048; // inputs are padded to a multiple of this (a multiple of BTL)
static_assert(TILE_POLES % BTL == 0, "tile size");
static_assert((BS & (BS - 1)) == 0, "ring depth must be a power of two (stage = g & (BS-1); measured: BS = 3, 5, 6 cost 7 % in index arithmetic, BS = 2 and 4 perform alike)");

enum : int { PH_FREE = 0, PH_BISECT, PH_DONE, PH_PENDING };

struct BSlot {   // cold per-slot state, touched by warp j only
    int64_t k, i;
    int phase, sideR, steps, q;   // q: steps since the slot took this eigenpair (time slice)
    unsigned long long ticket;    // PH_PENDING: the work ticket this slot is waiting on
    double sigma, wz2, b, left, right, m, nu1, Kb, Kz;
    double knu0, krho, sig2;   // round 1: first K_nu (HalfArrow keeps it), K_rho, squared shift (HalfArrow)
    // accelerated bisection (ACCEL): the point the coming sweep evaluates (the grid midpoint m or a probe), certified bounds, history
    double x, lo, hi, hx1, hF1, hx2, hF2, pole;
    double hx0, hF0, Flo, Fhi;   // third point of the history; F at the certified bounds (regula falsi fallback of the probe model)
    int nhist, probing, nprobe, nfail, cert;
};
struct BisectShared {
    alignas(128) double tileD[BS][BTL];
    alignas(128) double tileW[BS][BTL];
    alignas(8) unsigned long long full[BS];
    alignas(8) unsigned long long empty[BS];
    alignas(16) double2 hot[BJ];   // (sigma, m) of every slot for the coming sweep
    double part[BJ][BW];
    int skip[BJ];                  // 0-based index of the slot's shift pole, -1: none (idle slot, or Remedy-3 round); N < 2^30
    int act[BJ];                   // coming sweep: 1 slot is bisecting, 2 slot waits for its ticket, 0 slot has retired
    double save[BJ];               // accumulator of slot j saved by the one thread that owns its shift pole
    // speculative tail (see k_bisect): helpers of leader slot j for the coming sweep and the bisection points they evaluate
    int grpn[BJ];                  // number of helper slots of slot j: 0, 2 (both next midpoints) or 6 (the next two levels)
    int grp[BJ][6];
    double grp_pt[BJ][6];
    BSlot slot[BJ];
};

// element e (0..BE-1) of thread t in a tile sits at offset (e>>1)*2*BT + 2*t + (e&1):
// consecutive lanes read consecutive 16-byte pairs -> conflict-free LDS.128.
__device__ __forceinline__ int elem_offset(int t, int e) { return (e >> 1) * (2 * BT) + 2 * t + (e & 1); }

// Remedy-1 trigger of the reference (arrowhead3.jl:526-534, arrowhead4.jl:400-408); the Remedy itself runs in k_full.
template <int KIND>
__device__ __forceinline__ bool wants_remedy1(const Problem &P, int64_t k, int64_t i, bool sideR, double mu, double lambda) {
    const int64_t N = P.N, ii = i - 1;
    const double Di = P.D[ii];
    if (!(__ddiv_rn(__dadd_rn(fabs(Di), fabs(mu)), fabs(lambda)) > P.tau[4])) return false;
    const bool c1 = (k == 1 && P.D[0] < 0.0);
    const bool c2 = (KIND == KIND_ARROW) && (k == N + 1 && P.D[N - 1] > 0.0);
    const bool c3 = (KIND == KIND_ARROW ? i < N : true) && !sideR && sign_jl(Di) + sign_jl(P.D[ii + 1 < N ? ii + 1 : ii]) == 0.0;
    const bool c4 = (i > 1 && sideR && sign_jl(Di) + sign_jl(P.D[ii > 0 ? ii - 1 : 0]) == 0.0);
    return c1 || c2 || c3 || c4;
}

// ROUND 0: nu of the inverse arrow (arrowhead3.jl:449-465).  Standard eigenpairs are finished here; K_nu > tolnu
// goes to the Remedy-3 list (or, where the fast path has no round 1 for it, to k_full); the rest to k_full.
// ROUND 1: nu of the re-shifted inverse (DPR1), arrowhead3.jl:497-511.
template <int KIND, int ROUND>
__device__ __forceinline__ void bisect_finalize(const Problem &P, const Outputs &O, KState *ks, BSlot &S, int64_t kbase,
                                                int64_t *dlist, int64_t *rlist, unsigned long long *dcount, int lane) {
    const int64_t N = P.N;
    double nu = S.right;
    const int64_t ii = S.i - 1;
    double Knu = __ddiv_rn(S.nu1, fabs(nu));
    enum { GO_DONE, GO_FULL, GO_REM3 } go = GO_DONE;
    double mu = 0.0, lambda = 0.0, nu1 = S.nu1;
    if (ROUND == 0) {
        const bool big = (KIND == KIND_HALF) ? !(Knu < P.tau[2]) : (Knu > P.tau[2]);
        if (!isfinite(nu)) go = GO_FULL;
        else if (big) go = (KIND == KIND_HALF && S.k == N + 1) ? GO_FULL : GO_REM3;
        else {
            mu = __ddiv_rn(1.0, nu);
            if (KIND == KIND_HALF) {
                const double Di = P.D[ii];
                lambda = sqrt(__dadd_rn(mu, __dmul_rn(S.sigma, S.sigma)));
                if ((S.k == N + 1) &&
                    (__ddiv_rn(__dadd_rn(__dmul_rn(Di, Di), fabs(mu)), fabs(__dmul_rn(lambda, lambda))) > P.tau[4])) go = GO_FULL;
            } else {
                lambda = __dadd_rn(mu, S.sigma);
                if (wants_remedy1<KIND>(P, S.k, S.i, S.sideR != 0, mu, lambda)) go = GO_FULL;
            }
        }
        if (go == GO_REM3) {   // sign conventions of arrowhead3.jl:469-471
            nu = S.sideR ? fabs(nu) : -fabs(nu);
            nu1 = -sign_jl(nu) * nu1;
        }
    } else {
        if (KIND == KIND_ARROW && S.sideR && P.D[0] > 0.0 && nu < 0.0) go = GO_FULL;      // 'R' -> 'L' retry
        if (KIND != KIND_HALF) {
            if (isnan(Knu) || Knu > P.tau[2]) go = GO_FULL;                              // the while loop would go on
        } else {
            Knu = S.knu0;                                                                // arrowhead6.jl keeps the first K_nu
        }
        if (!isfinite(nu)) go = GO_FULL;
        if (go == GO_DONE) {
            mu = __ddiv_rn(1.0, nu);
            if (KIND == KIND_HALF) lambda = sqrt(__dadd_rn(mu, S.sig2));
            else {
                lambda = __dadd_rn(mu, S.sigma);
                if (wants_remedy1<KIND>(P, S.k, S.i, S.sideR != 0, mu, lambda)) go = GO_FULL;
            }
        }
    }
    if (go == GO_DONE && !interlace_ok(P.D, N, S.k, lambda)) go = GO_FULL;   // k_full repeats the reference's scheme and flags what it gets
    if (lane == 0) {
        const int64_t col = S.k - kbase;
        KState &G = ks[col];
        if (go == GO_FULL) {
            G.flag = KS_HANDOFF;
            dlist[atomicAdd(dcount, 1ull)] = S.k;
        } else if (go == GO_REM3) {
            G.nu = nu; G.nu1 = nu1; G.knu = Knu; G.steps = S.steps;
            G.flag = KS_REM3;
            rlist[atomicAdd(dcount + 2, 1ull)] = S.k;
        } else {
            G.nu = nu; G.mu = mu; G.lambda = lambda; G.knu = Knu; G.steps = S.steps;
            G.flag = (ROUND == 0) ? KS_DONE : KS_DONE1;
            O.lambda[col] = lambda;
            O.sind[col] = S.i;
            O.kb[col] = S.Kb; O.kz[col] = S.Kz; O.knu[col] = Knu; O.krho[col] = (ROUND == 0) ? 0.0 : S.krho;
            O.qout[col] = 0;
            O.steps[col] = S.steps;
            O.status[col] = (ROUND == 0) ? 0 : AH_BR_REM3;   // round 1 = one plain Remedy-3 re-shift (anything more went to k_full)
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Accelerated bisection (see the comment above k_bisect): certificates, free descent of the dyadic grid, probe points.
// All of it is scalar work of one thread on a slot's state; the sweeps themselves do not change.
//
// The decision of one bisection step as a monotone function of the sweep's sum: goleft <=> value > 0, in exactly the
// reference's operation order (round 0: arrowhead3.jl:696-701 on the inverse arrow; round 1: :736-740 on the DPR1 inverse).
template <int KIND, int ROUND>
struct StepDecision {
    double bm, c, wz2, b, sgn;
    __device__ __forceinline__ StepDecision(const BSlot &S, double x) {
        if (ROUND == 0) {
            wz2 = S.wz2; b = S.b; sgn = 1.0;
            bm = __dadd_rn(S.b, -x);
            c = (KIND != KIND_DPR1) ? __ddiv_rn(wz2, __dadd_rn(0.0, -x)) : 0.0;
        } else {
            wz2 = 1.0; b = S.b; bm = 0.0; sgn = sign_jl(S.b);
            c = (KIND != KIND_DPR1) ? __ddiv_rn(1.0, __dadd_rn(0.0, -x)) : 0.0;
        }
    }
    __device__ __forceinline__ double operator()(double s) const {
        if (ROUND == 0) {
            double t = __dmul_rn(s, wz2);
            if (KIND != KIND_DPR1) t = __dadd_rn(t, c);
            return __dadd_rn(bm, -t);
        } else {
            if (KIND != KIND_DPR1) s = __dadd_rn(s, c);
            const double F = __dadd_rn(1.0, __dmul_rn(b, s));     // b holds rho
            return -(sgn * F);
        }
    }
};

// Preconditions of the certificates for one eigenpair: (1) the bracket lies outside ALL poles of the inverse, so that every
// term of a sweep has the same sign -- always true in round 0 (extreme eigenvalue of the inverse arrow); in round 1 only for
// the brackets [mx1, mx1 + rho uu] (rho > 0, 'R') and [mn1 + rho uu, mn1] (rho < 0, 'L') of arrowhead3.jl:722-726, the other
// two lie between the two outermost poles; (2) t = d (1 - m d) cannot overflow for any m of the bracket and any pole, pads
// included (dshift is monotone in the pole, so the two ends of the padded pole range bound |d|).
template <int KIND, int ROUND>
__device__ __forceinline__ int accel_certifiable(const Problem &P, const BSlot &S) {
    if (ROUND == 1 && !((S.b > 0.0) == (S.sideR != 0))) return 0;
    const double dmax = fmax(fabs(dshift<KIND>(P.D[0], S.sigma)), fabs(dshift<KIND>(P.pad, S.sigma)));
    return (fmax(fabs(S.left), fabs(S.right)) * dmax * dmax < 1.0e300) ? 1 : 0;
}
__device__ __forceinline__ void accel_reset(BSlot &S, int cert) {
    S.cert = cert;
    S.lo = -INFINITY; S.hi = INFINITY; S.hx1 = 0.0; S.hF1 = 0.0; S.hx2 = 0.0; S.hF2 = 0.0;
    S.hx0 = 0.0; S.hF0 = 0.0; S.Flo = INFINITY; S.Fhi = -INFINITY;
    S.pole = S.sideR ? S.left : S.right;
    S.nhist = 0; S.probing = cert; S.nprobe = 0; S.nfail = 0;
    S.x = S.m;
}
// After a sweep at S.x with the sum `sum` (not NaN): certificates, history, and -- at a grid point -- the bisection step itself.
template <int KIND, int ROUND>
__device__ __forceinline__ void accel_update(BSlot &S, double sum, double th2) {
    const double x = S.x, m = S.m;
    const bool grid = (x == m);
    const StepDecision<KIND, ROUND> dec(S, x);
    const double F = dec(sum);
    if (S.cert && fabs(sum) > 1.0e-250 && fabs(sum) <= 1.7976931348623157e308) {
        const bool la = dec(__dmul_rn(sum, 1.0 + th2)) > 0.0, lb = dec(__dmul_rn(sum, 1.0 - th2)) > 0.0;
        if (la && lb) { if (x > S.lo) { S.lo = x; S.Flo = F; } }
        else if (!la && !lb) { if (x < S.hi) { S.hi = x; S.Fhi = F; } }
        else if (!grid && ++S.nfail >= 3) S.probing = 0;    // (every failed probe widens the next probe's offset 4x)
    }
    if (fabs(F) <= 1.7976931348623157e308) {
        if (S.nhist == 0 || x != S.hx2) {
            S.hx0 = S.hx1; S.hF0 = S.hF1; S.hx1 = S.hx2; S.hF1 = S.hF2; S.hx2 = x; S.hF2 = F;
            if (S.nhist < 3) S.nhist += 1;
        }
    } else {
        S.probing = 0;
    }
    if (grid) { if (F > 0.0) S.left = m; else S.right = m; }   // the step plain bisection takes here (an infinite sum included)
}
// Grid points decided by the certificates cost nothing; S.m = the first grid midpoint that is not (or the converged bracket's).
__device__ __forceinline__ void accel_descend(BSlot &S) {
    double l = S.left, r = S.right;
    const double lo = S.lo, hi = S.hi;
    double m = (l + r) / 2.0;
    while ((r - l) > 2.0 * kEps * fmax(fabs(l), fabs(r))) {
        if (m <= lo) l = m;
        else if (m >= hi) r = m;
        else break;
        m = (l + r) / 2.0;
    }
    S.left = l; S.right = r; S.m = m;
}
// Where the coming sweep evaluates F: the grid midpoint, or a probe that tightens lo / hi.  Root estimate from the model
// value(m) = L(m) + alpha + gamma / (m - q) through the last evaluated points, L(m) = b - m in round 0 (the known linear part
// of the arrow's secular function), 0 in round 1: with three points q is fitted too (it lumps the poles behind the extreme
// one), with two q = the extreme pole; if neither lands inside the certified interval, regula falsi on its ends.  The probe
// sits delta = 1.25 x (certifiable distance) -- 4x more after every probe that could not be certified, never below 4 ulp -- to
// the side of the estimate whose certificate is further away.  Heuristic only: no influence on the result.
template <int ROUND>
__device__ __forceinline__ void accel_choose(BSlot &S, double th2) {
    double x = S.m;
    if (S.probing && S.nhist >= 2 && S.nprobe < 24) {
        const double x1 = S.hx1, F1 = S.hF1, x2 = S.hx2, F2 = S.hF2, p = S.pole;
        const double a_ = fmax(S.lo, S.left), b_ = fmin(S.hi, S.right);
        const double slope = fabs((F2 - F1) * rcp_fast(x2 - x1));
        const double L1 = (ROUND == 0) ? S.b - x1 : 0.0, L2 = (ROUND == 0) ? S.b - x2 : 0.0;
        const double R1 = F1 - L1, R2 = F2 - L2;
        // the uncertain part of F: wz^2 S (+ the tip term) in round 0, rho S in round 1 -- a point is certifiable once |F| exceeds
        // th2 times it; 1.25 x that distance (measured on B200: 1.5 x of |b - m| + |R| = 13.7 sweeps per eigenpair, this 12.9)
        const double scale = (ROUND == 0) ? fabs(R2) : fabs(sign_jl(S.b) * F2 + 1.0);
        double delta = 1.25 * th2 * scale * rcp_fast(slope) * (double)(1 << (2 * S.nfail));
        delta = fmax(delta, 4.0 * kEps * fabs(x2));
        if (!((b_ - a_) > 6.0 * delta)) S.probing = 0;          // (also when delta is NaN)
        else {
            auto root_with_pole = [&](double q) {
                const double i1 = rcp_fast(x1 - q), i2 = rcp_fast(x2 - q);
                const double gam = (R1 - R2) * rcp_fast(i1 - i2), alpha = R2 - gam * i2;
                if (ROUND == 0) {   // (b + alpha - m)(m - q) + gamma = 0
                    const double B = S.b + alpha + q, C = (S.b + alpha) * q - gam;
                    const double sq = sqrt(B * B - 4.0 * C);
                    const double ra = (B + sq) / 2.0, rb = (B - sq) / 2.0;
                    return (ra > a_ && ra < b_) ? ra : rb;
                }
                return q - gam * rcp_fast(alpha);
            };
            double xs = NAN;
            if (S.nhist >= 3) {
                const double x0 = S.hx0, R0 = S.hF0 - ((ROUND == 0) ? S.b - x0 : 0.0);
                const double rho = (R0 - R1) * (x2 - x1) * rcp_fast((R1 - R2) * (x1 - x0));
                const double qq = (x2 - rho * x0) * rcp_fast(1.0 - rho);
                const double off = S.sideR ? qq - p : p - qq;      // the fitted pole must not lie on the root's side of the extreme pole
                if (off <= 0.0) xs = root_with_pole(qq);
            }
            if (!(xs > a_ && xs < b_)) xs = root_with_pole(p);
            if (!(xs > a_ && xs < b_) && S.lo >= S.left && S.hi <= S.right) xs = S.lo + (S.hi - S.lo) * S.Flo * rcp_fast(S.Flo - S.Fhi);
            const double xp = ((xs - a_) > (b_ - xs)) ? xs - delta : xs + delta;
            if (xp > a_ && xp < b_) { x = xp; S.nprobe += 1; }
        }
    }
    S.x = x;
}

__device__ __forceinline__ bool bisect_converged(const BSlot &S) {
    return !((S.right - S.left) > 2.0 * kEps * fmax(fabs(S.left), fabs(S.right)));
}

__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Work distribution.  Eigenpairs are handed to slots by ticket: ticket t < cnt is eigenpair t of the block (round 1:
// entry t of the Remedy-3 list), taken with one atomicAdd.  A bisection is ~55 dependent sweeps, so with plain
// first-come-first-served the last "wave" of a launch leaves slots idle: half a wave at the end of a long launch, and
// most of a wave when the k-range is short (a 1/8 shard of n = 32768 is 1.7 eigenpairs per slot: 2 x 55 sweeps with
// 27 % of the slots idle in the second half).  Round 0 therefore TIME-SLICES once the eigenpairs that are still
// waiting would fit in two more waves: a slot that has spent `quantum` steps on its eigenpair while others wait
// writes the bracket back to its KState record, appends the eigenpair to a FIFO (ticket cnt + r, r = dcount[4]++;
// entry = launch tag << 32 | column, so the buffer never needs clearing) and takes the next ticket.  All eigenpairs
// in flight then advance at the same rate and finish together: the launch ends after ~(sum of steps) / slots sweeps.
// A slot whose ticket has no entry yet idles (PH_PENDING; it never blocks its CTA) until the entry appears or every
// eigenpair of the block is finished (dcount[5] == cnt).  Which slot advances an eigenpair, and when, has no
// influence on its arithmetic: results are bitwise independent of the schedule.
// NSLOT: slots per CTA.  All NSLOT slots are evaluated in every sweep (skipping idle ones with a branch breaks the
// interleaving of the chains and costs more than it saves), so a block of fewer than 8 eigenpairs per CTA runs with
// 7 or 6 slots.  Round 1 is the exception: its list is ~1 % of the block, slot j of a CTA is used only while
// j * CTAs < cnt (which spreads the list over all CTAs) and the sweep skips the unused slots.
//
// ACCEL (round 0): certified skipping of bisection steps.  The result of a bisection is the final bracket, i.e. the cell of
// the dyadic grid  m = fl((l + r)/2)  that the computed decisions  F7(m) > 0  lead to.  Most of these ~54 decisions are
// known without evaluating F7(m): F is strictly decreasing on the bracket (the bracket lies outside all poles of the
// inverse, so every secular term has the same sign), and a sweep at ANY point x bounds the computed sum at every grid
// point on one side of x:
//   * same sign  =>  the computed sum S7 and the exact sum S_X of the computed terms' exact values satisfy
//     |S7 - S_X| <= theta |S_X|, theta = (terms per thread + reduction depth + 2) * 2^-53  (<= 1.6 ulp per reciprocal, one rounding
//     per accumulation; no cancellation), and S_X(m) is monotone in m because every fl operation of a term is;
//   * the decision  fl(fl(b - m) - fl(fl(S wz^2) + fl(wz^2 / -m))) > 0  is monotone in S and in m for the same reason.
// Hence if the decision at x comes out "left" for S7(x) (1 +- 2.01 theta) both, F7(m) > 0 at every grid point m <= x ("lo"),
// and likewise "right" for every m >= x ("hi").  Grid points outside (lo, hi) are decided for free; grid points inside are
// evaluated and decided by their own F7(m) exactly as plain bisection would -- the final bracket is bit-identical.
// Where to evaluate is a heuristic with no influence on the result: the first two sweeps are grid points, then a secant
// iteration on (m - pole) F(m) (exact for F = alpha + beta/(m - pole)) aims delta = 4 x 2.01 theta scale / |F'| to the side
// of its root estimate whose certificate is further away, until lo and hi are ~8 delta apart; the rest (the last ~10 grid
// cells, where F7 is rounding noise) are plain steps.  ~19 sweeps instead of ~54.  Preconditions checked per eigenpair
// (else plain bisection): no overflow of d (1 - m d) anywhere in the bracket; per launch: N <= 2^21.
template <int KIND, int ROUND, int NSLOT, int ACCEL>
__global__ void __launch_bounds__(BT, BCTAS)
k_bisect(Problem P, Outputs O, KState *__restrict__ ks, int64_t kbase, int64_t cnt, int64_t *__restrict__ dlist,
         int64_t *__restrict__ rlist, unsigned long long *dcount, unsigned long long *__restrict__ queue, long long qcap,
         unsigned tag, int quantum, int window, int spec, KExtra *__restrict__ kx, double th2) {
    static_assert(NSLOT >= 1 && NSLOT <= BJ && NSLOT <= BW, "one warp per slot");
    // Round 0: dcount[6] eigenpairs were handed over by k_prep and are being solved by k_full on the side stream, one CTA
    // each; this grid alone fills every SM, so that many of its CTAs (16 at most, and never more than an eighth of the grid)
    // make room.  Work is taken by ticket, so the others simply do a little more.
    if (ROUND == 0 && (unsigned long long)blockIdx.x < min(dcount[6], (unsigned long long)min(16u, gridDim.x >> 3))) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BisectShared &sh = *reinterpret_cast<BisectShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = (int)(P.Np / BTL);

    if (tid == 0) {
        for (int s = 0; s < BS; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], BW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < NSLOT) { sh.slot[tid].phase = PH_FREE; sh.skip[tid] = -1; sh.act[tid] = 0; sh.hot[tid] = make_double2(0.0, 0.0); sh.grpn[tid] = 0; }
    if (ROUND == 1) { cnt = (int64_t)dcount[2]; quantum = 0; }   // round 1 works through the (short) Remedy-3 list
#if AH_STAGGER
    // Every sweep is ntiles tiles for every CTA, so the two CTAs of an SM would reach their sweep boundaries (reduce,
    // decide, refill: ~2 us without FP64 work) together for the whole launch.  The second CTA to arrive on an SM starts
    // half a sweep late instead, and each then computes through the other's boundary.
    if (ROUND == 0 && tid == 0 && ntiles >= 4) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned *const arrivals = reinterpret_cast<unsigned *>(dcount + 16);
        if (atomicAdd(&arrivals[smid & 255u], 1u) & 1u) {
            const unsigned long long wait_ns = (unsigned long long)ntiles * AH_STAGGER_NS_PER_TILE;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            do {
                __nanosleep(500);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            } while (t1 - t0 < wait_ns);
        }
    }
#endif
    __syncthreads();
    if (ROUND == 1 && cnt <= REM3_DIRECT_MAX) return;   // short list: k_rem3_prep has bisected it directly
    const unsigned long long ucnt = (unsigned long long)cnt;
    const unsigned long long slots_total = (unsigned long long)gridDim.x * NSLOT;
    unsigned long long *const tickets = dcount + (ROUND == 0 ? 1 : 3);
    // shortest slice near convergence: a switch costs ~2.5 us of latency, so not more often than every ~32 tiles, and
    // not shortened at all when the sweeps are short (measured: n = 8192 loses 3 %, n >= 16384 gains 1-3 %)
    const int qmin = ntiles >= 16 ? max(1, 32 / ntiles) : quantum;

    bool wrapped = false;    // g_issue has wrapped around 2^32 at least once
    uint32_t g_issue = 0;    // tiles issued (thread 0 only); wraps mod 2^32, a multiple of 2*BS -> parities stay consistent
    uint32_t g = 0;          // tiles consumed
    uint32_t issue_pos = 0;  // g_issue mod ntiles
    auto issue_tile = [&](uint32_t gi) {
        const int s = (int)(gi & (BS - 1));
        const uint32_t use = gi / BS;
        if (gi >= (uint32_t)BS || wrapped) mbar_wait(&sh.empty[s], (use - 1u) & 1u);
        const int tile = (int)issue_pos;
        issue_pos = (issue_pos + 1 == (uint32_t)ntiles) ? 0u : issue_pos + 1;
        mbar_arrive_expect_tx(&sh.full[s], 2u * BTL * 8u);
        bulk_g2s(&sh.tileD[s][0], P.D + (int64_t)tile * BTL, BTL * 8u, &sh.full[s]);
        bulk_g2s(&sh.tileW[s][0], P.w2 + (int64_t)tile * BTL, BTL * 8u, &sh.full[s]);
    };
    if (tid == 0) { for (; g_issue < (uint32_t)(BS - 1); ++g_issue) issue_tile(g_issue); }

#ifdef AH_DIAG_TIMING
    // diagnostic build only: where warp 0 spends its cycles (clock64 deltas summed per CTA, atomically added to dcount[152..159])
    long long tg_boundary = 0, tg_b2 = 0, tg_param = 0, tg_tiles = 0, tg_reduce = 0, tg_b1 = 0, tg_sweeps = 0, tg_mark = clock64();
    long long tb_decide = 0, tb_final = 0, tb_swapchk = 0, tb_requeue = 0, tb_ticket = 0, tb_resolve = 0, tb_load = 0, tb_n_final = 0,
              tb_n_swapchk = 0, tb_n_requeue = 0, tb_n_ticket = 0, tb_n_poll = 0, tb_mark = 0;
#define AH_TG(var) do { if (tid == 0) { const long long now__ = clock64(); var += now__ - tg_mark; tg_mark = now__; } } while (0)
#define AH_TB_START() do { if (tid == 0) tb_mark = clock64(); } while (0)
#define AH_TB(var, cnt) do { if (tid == 0) { const long long now__ = clock64(); var += now__ - tb_mark; tb_mark = now__; cnt += 1; } } while (0)
#else
#define AH_TG(var) do { } while (0)
#define AH_TB_START() do { } while (0)
#define AH_TB(var, cnt) do { } while (0)
#endif
    bool first = true;
    while (true) {
        AH_TG(tg_b1);
        // ================= sweep boundary: warp j advances slot j =================
        if (warp < NSLOT) {
            BSlot &S = sh.slot[warp];
            AH_TB_START();
            // one bisection step of this slot from the partial sums that slot `ps` accumulated in the sweep just finished
            // (ps == warp: the slot's own; otherwise a helper slot that evaluated the slot's NEXT point, see below);
            // returns 1 / 0 for the two halves, -1 when the eigenpair left the slot (NaN sweep: hand-over to k_full)
            auto bisection_step = [&](int ps) -> int {
                double sum = (lane < BW) ? sh.part[ps][lane] : 0.0;
#pragma unroll
                for (int o = BW / 2; o > 0; o >>= 1) {
                    const double p = __shfl_xor_sync(0xffffffffu, sum, o);
                    sum = (lane & o) ? __dadd_rn(p, sum) : __dadd_rn(sum, p);
                }
                sum = __shfl_sync(0xffffffffu, sum, 0);
                const double m = S.m;
                bool goleft;
                if (sum != sum) {
                    // NaN: a term of this sweep was 0*Inf inside the reciprocal correction -- the bisection point sits
                    // exactly on an inverse pole (possibly a pad pole's, for k = N+1) or d*(1-m*d) overflowed.  k_full
                    // redoes the eigenpair from scratch and evaluates such sweeps with exact divisions.
                    __syncwarp();
                    if (lane == 0) {
                        ks[S.k - kbase].flag = KS_HANDOFF;
                        dlist[atomicAdd(dcount, 1ull)] = S.k;
                        if (quantum > 0) atomicAdd(dcount + 5, 1ull);
                        S.phase = PH_FREE;
                    }
                    __syncwarp();
                    return -1;
                } else if (ROUND == 0) {
                    sum = __dmul_rn(sum, S.wz2);
                    if (KIND != KIND_DPR1) sum = __dadd_rn(sum, __ddiv_rn(S.wz2, __dadd_rn(0.0, -m)));
                    goleft = __dadd_rn(__dadd_rn(S.b, -m), -sum) > 0.0;                  // arrowhead3.jl:696-701
                } else {
                    if (KIND != KIND_DPR1) sum = __dadd_rn(sum, __ddiv_rn(1.0, __dadd_rn(0.0, -m)));
                    const double F = __dadd_rn(1.0, __dmul_rn(S.b, sum));                // b holds rho; arrowhead3.jl:736-740
                    goleft = sign_jl(S.b) * F < 0.0;
                }
                __syncwarp();
                if (lane == 0) {
                    if (goleft) S.left = m; else S.right = m;
                    S.m = (S.left + S.right) / 2.0;
                    S.steps += 1;
                    S.q += 1;
                }
                __syncwarp();
                return goleft ? 1 : 0;
            };
            // ACCEL: the sweep evaluated S.x (the grid midpoint S.m or a probe); ps != warp: helper slot ps evaluated the
            // grid midpoint S.m for this slot (speculative tail)
            auto accel_step = [&](int ps) -> int {
                double sum = (lane < BW) ? sh.part[ps][lane] : 0.0;
#pragma unroll
                for (int o = BW / 2; o > 0; o >>= 1) {
                    const double p = __shfl_xor_sync(0xffffffffu, sum, o);
                    sum = (lane & o) ? __dadd_rn(p, sum) : __dadd_rn(sum, p);
                }
                int rcode = 0;
                if (lane == 0) {
                    if (ps != warp) S.x = S.m;
                    if (sum != sum) {           // NaN sweep
                        if (S.x == S.m) {       // at a grid point, as in plain bisection: k_full redoes the eigenpair with exact divisions
                            ks[S.k - kbase].flag = KS_HANDOFF;
                            dlist[atomicAdd(dcount, 1ull)] = S.k;
                            if (quantum > 0) atomicAdd(dcount + 5, 1ull);
                            S.phase = PH_FREE;
                            rcode = -1;
                        } else {
                            S.probing = 0;      // a probe that cannot be used
                        }
                    } else {
                        accel_update<KIND, ROUND>(S, sum, th2);
                    }
                    S.steps += 1;
                    S.q += 1;
                }
                rcode = __shfl_sync(0xffffffffu, rcode, 0);
                __syncwarp();
                return rcode;
            };
            if (ACCEL) {
                if (!first && S.phase == PH_BISECT) {
                    int rc = accel_step(warp);
                    // Speculative tail: helper slots evaluated the midpoints of the next grid level(s) below the point this
                    // slot's own sweep decided.  Whatever the certificates do not decide anyway and a helper happens to
                    // have evaluated is a real bisection step taken here, from the helper's sum -- same point, same sum,
                    // same decision as the sweep this slot would have spent on it.
                    const int nh = (ROUND == 0) ? sh.grpn[warp] : 0;
                    for (int lvl = 0; lvl < 2 && nh >= 2 && rc >= 0 && (lvl == 0 || nh >= 6); ++lvl) {
                        int t = -1;
                        if (lane == 0) {
                            accel_descend(S);
                            if (!bisect_converged(S))
                                for (int q = 0; q < nh; ++q)
                                    if (sh.grp_pt[warp][q] == S.m) t = q;
                        }
                        t = __shfl_sync(0xffffffffu, t, 0);
                        if (t < 0) break;
                        rc = accel_step(sh.grp[warp][t]);
                    }
                }
            } else if (!first && S.phase == PH_BISECT) {
                const int g1 = bisection_step(warp);
                // Speculative tail: helper slots (idle slots of this CTA) evaluated F at the midpoints of both halves
                // (and, with six helpers, of all four quarters) in the same sweep, so the step(s) that the reference
                // would take next are decided here from their sums -- same points, same sums, same decisions, in the
                // same order; a level is used only if the bracket has not converged and the helper's point is bitwise
                // the slot's new midpoint.
                const int nh = (ROUND == 0) ? sh.grpn[warp] : 0;
                if (nh >= 2 && g1 >= 0 && !bisect_converged(S)) {
                    const int h1 = g1 ? 1 : 0;                       // goleft: bracket [m, right] -> point B, else point A
                    if (sh.grp_pt[warp][h1] == S.m) {
                        const int g2 = bisection_step(sh.grp[warp][h1]);
                        if (nh >= 6 && g2 >= 0 && !bisect_converged(S)) {
                            const int h2 = 2 + 2 * h1 + (g2 ? 1 : 0);
                            if (sh.grp_pt[warp][h2] == S.m) bisection_step(sh.grp[warp][h2]);
                        }
                    }
                }
            }
#ifdef AH_DIAG_TIMING
            long long tb_dummy = 0;
            AH_TB(tb_decide, tb_dummy);
#endif
            while (true) {
                if (S.phase == PH_BISECT) {
                    if (ACCEL) {
                        if (lane == 0) accel_descend(S);
                        __syncwarp();
                    }
                    if (bisect_converged(S)) {   // (the reference tests before the first step, too) -> finalise
                        bisect_finalize<KIND, ROUND>(P, O, ks, S, kbase, dlist, rlist, dcount, lane);
                        __syncwarp();
                        if (lane == 0) { if (quantum > 0) atomicAdd(dcount + 5, 1ull); S.phase = PH_FREE; }
                        __syncwarp();
                        AH_TB(tb_final, tb_n_final);
                    } else {
                        // time slice used up and others are waiting -> back of the queue
                        int swap = 0;
                        const int qv = S.q;
                        // The slice shrinks as the eigenpair nears convergence (steps left ~ exponent of the bracket width
                        // over 2 eps max|bracket|): time-sliced eigenpairs then finish within a step or two of each other
                        // instead of a quantum or two, which is what the launch waits for at its end.
                        const double wdt = S.right - S.left, mxb = fmax(fabs(S.left), fabs(S.right));
                        const int left_est = ((__double2hiint(wdt) >> 20) & 0x7ff) - ((__double2hiint(mxb) >> 20) & 0x7ff) + 51;
                        const int qeff = min(quantum, max(qmin, left_est >> AH_QSHIFT));
                        __syncwarp();
                        // (the exterior eigenpairs are never handed back: they need 2-3 times the sweeps of the others -- decades
                        // of bracket, then ~17 steps inside the rounding noise of F -- and would end the launch alone)
                        if (ROUND == 0 && quantum > 0 && qv >= qeff && S.k != 1 && S.k != P.N + 1) {
                            if (lane == 0) {
                                const unsigned long long head = ld_volatile(tickets), nreq = ld_volatile(dcount + 4);
                                const unsigned long long tail = ucnt + nreq;
                                swap = (head < tail && tail - head < (unsigned long long)window * slots_total &&
                                        nreq + slots_total < (unsigned long long)qcap) ? 1 : 0;
                                S.q = 0;
                            }
                            swap = __shfl_sync(0xffffffffu, swap, 0);
                            AH_TB(tb_swapchk, tb_n_swapchk);
                        }
                        if (!swap) {
                            if (ACCEL) {
                                if (lane == 0) accel_choose<ROUND>(S, th2);
                                __syncwarp();
                            }
                            break;
                        }
                        if (lane == 0) {
                            KState &G = ks[S.k - kbase];
                            G.left = S.left; G.right = S.right; G.steps = S.steps;
                            if (ACCEL) {
                                KExtra &X = kx[S.k - kbase];
                                X.lo = S.lo; X.hi = S.hi; X.hx1 = S.hx1; X.hF1 = S.hF1; X.hx2 = S.hx2; X.hF2 = S.hF2; X.pole = S.pole;
                                X.nhist = (int8_t)S.nhist; X.probing = (int8_t)S.probing; X.nprobe = (int8_t)S.nprobe; X.nfail = (int8_t)S.nfail;
                                G.nu = S.hx0; G.mu = S.hF0; G.lambda = S.Flo; G.knu = S.Fhi;   // (free until the finalize)
                            }
                            __threadfence();
                            const unsigned long long r = atomicAdd(dcount + 4, 1ull);
                            *reinterpret_cast<volatile unsigned long long *>(queue + r) =
                                ((unsigned long long)tag << 32) | (unsigned long long)(uint32_t)(S.k - kbase);
                            S.phase = PH_FREE;
                        }
                        __syncwarp();
                        AH_TB(tb_requeue, tb_n_requeue);
                    }
                }
                if (S.phase == PH_DONE) break;
                if (ROUND == 1 && (int64_t)warp * gridDim.x >= cnt) {   // spread a short list over all CTAs
                    __syncwarp();
                    if (lane == 0) S.phase = PH_DONE;
                    __syncwarp();
                    break;
                }
                if (S.phase == PH_FREE) {
                    if (lane == 0) { S.ticket = atomicAdd(tickets, 1ull); S.phase = PH_PENDING; }
                    __syncwarp();
                    AH_TB(tb_ticket, tb_n_ticket);
                }
                // PH_PENDING: which eigenpair does the ticket stand for?
                long long col = -1;      // >= 0: column of the block, -1: not there yet, -2: never
                if (lane == 0) {
                    const unsigned long long t = S.ticket;
                    // round 0: the last column first (if the block ends with k = N+1, that exterior eigenpair is one of the two
                    // longest bisections of the matrix; k = 1, the other one, is column 0 of its block), then columns 0, 1, ...
                    if (t < ucnt) col = (ROUND == 0) ? (t == 0 ? (long long)(ucnt - 1) : (long long)(t - 1)) : (long long)(rlist[t] - kbase);
                    else if (quantum == 0) col = -2;
                    else {
                        const unsigned long long r = t - ucnt;
                        const unsigned long long e = (r < (unsigned long long)qcap) ? ld_volatile(queue + r) : 0ull;
                        if ((unsigned)(e >> 32) == tag) { __threadfence(); col = (long long)(uint32_t)e; }
                        else if (ld_volatile(dcount + 5) >= ucnt) col = -2;   // everything is finished: nothing will be queued any more
                    }
                }
                col = __shfl_sync(0xffffffffu, col, 0);
                AH_TB(tb_resolve, tb_n_poll);
                if (col < 0) {
                    __syncwarp();
                    if (lane == 0 && col == -2) S.phase = PH_DONE;
                    __syncwarp();
                    break;
                }
                const KState &G = ks[col];
                if (G.flag != (ROUND == 0 ? KS_READY : KS_REM3_READY)) {   // k_prep handed it over: not for this kernel
                    __syncwarp();
                    if (lane == 0) { if (quantum > 0) atomicAdd(dcount + 5, 1ull); S.phase = PH_FREE; }
                    __syncwarp();
                    continue;
                }
                __syncwarp();
                if (lane == 0) {
                    S.k = kbase + col; S.i = G.i; S.sideR = G.sideR;
                    S.sigma = G.sigma; S.b = G.b;
                    S.left = __ldcg(&G.left); S.right = __ldcg(&G.right); S.m = (S.left + S.right) / 2.0;   // (a requeued bracket comes from another SM)
                    S.steps = __ldcg(&G.steps);
                    S.nu1 = G.nu1; S.Kb = G.Kb; S.Kz = G.Kz;
                    if (ROUND == 0) S.wz2 = __dmul_rn(G.wz, G.wz);
                    else { S.wz2 = 1.0; S.knu0 = G.knu; S.krho = G.krho; S.sig2 = G.lambda; }
                    S.q = 0;
                    S.phase = PH_BISECT;
                    if (ACCEL) {
                        accel_reset(S, accel_certifiable<KIND, ROUND>(P, S));
                        if (ROUND == 0 && S.steps > 0) {   // time-sliced before: its certificates and history come along
                            const KExtra &X = kx[col];
                            S.lo = __ldcg(&X.lo); S.hi = __ldcg(&X.hi); S.hx1 = __ldcg(&X.hx1); S.hF1 = __ldcg(&X.hF1);
                            S.hx2 = __ldcg(&X.hx2); S.hF2 = __ldcg(&X.hF2); S.pole = __ldcg(&X.pole);
                            const int pk = __ldcg(reinterpret_cast<const int *>(&X.nhist));
                            S.nhist = (int)(int8_t)(pk & 0xff); S.probing = (int)(int8_t)((pk >> 8) & 0xff);
                            S.nprobe = (int)(int8_t)((pk >> 16) & 0xff); S.nfail = (int)(int8_t)((pk >> 24) & 0xff);
                            S.hx0 = __ldcg(&G.nu); S.hF0 = __ldcg(&G.mu); S.Flo = __ldcg(&G.lambda); S.Fhi = __ldcg(&G.knu);
                        }
                    }
                }
                __syncwarp();
#ifdef AH_DIAG_TIMING
                AH_TB(tb_load, tb_dummy);
#endif
            }
            if (lane == 0) {
                const bool act = S.phase == PH_BISECT;
                sh.hot[warp] = act ? make_double2(S.sigma, ACCEL ? S.x : S.m) : make_double2(0.0, 0.0);
                sh.skip[warp] = (act && ROUND == 0) ? (int)(S.i - 1) : -1;   // the re-shifted sigma_1 is not a pole
                // 1: plain bisection steps (helper slots may evaluate the next grid levels), 3: (ACCEL) still probing -- the
                // sequence of evaluated points must not depend on whether helpers were around, so none are used --,
                // 2: waits for its ticket, 0: retired
                sh.act[warp] = act ? ((ACCEL && S.probing) ? 3 : 1) : (S.phase == PH_PENDING ? 2 : 0);
            }
        }
        first = false;
        AH_TG(tg_boundary);
        __syncthreads();
        AH_TG(tg_b2);

        // ================= speculative tail: idle slots help the slots that still bisect =================
        // Towards the end of a launch a CTA holds fewer eigenpairs than slots, and a bisection is a chain of ~55
        // DEPENDENT sweeps: the last eigenpairs (up to 73 steps) would run on alone while every sweep still evaluates all
        // NSLOT slots.  An idle slot therefore evaluates, for a slot that still bisects, the point that slot will need
        // NEXT: with two helpers both possible next midpoints (two bisection steps per sweep), with six the two next
        // levels (three steps per sweep).  The arithmetic of every F(m) is unchanged, so results stay bit-identical.
        if (ROUND == 0 && spec) {
            unsigned busy = 0, idle = 0;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) { const int aj = sh.act[j]; if (aj == 1) busy |= 1u << j; else if (aj != 3) idle |= 1u << j; }
            const int nb = __popc(busy), ni = __popc(idle);
            // two helpers for as many leaders as the idle slots allow, four more for the first of them where they last
            const int nb2 = min(nb, ni >> 1), nb6 = min(nb2, (ni - 2 * nb2) >> 2);
            if (nb2) {                                                  // CTA-uniform
                if (warp < NSLOT && lane == 0) {
                    auto nth_bit = [](unsigned mask, int r) { for (int q = 0; q < r; ++q) mask &= mask - 1; return __ffs(mask) - 1; };
                    if ((busy >> warp) & 1u) {                          // leader: its helpers, in ascending slot order
                        const int rank = __popc(busy & ((1u << warp) - 1u));
                        const int h = rank < nb6 ? 6 : (rank < nb2 ? 2 : 0);
                        const int off = rank < nb6 ? 6 * rank : 6 * nb6 + 2 * (rank - nb6);
                        for (int t = 0; t < h; ++t) sh.grp[warp][t] = nth_bit(idle, off + t);
                        sh.grpn[warp] = h;
                    } else {
                        const int ri = __popc(idle & ((1u << warp) - 1u));
                        if (((idle >> warp) & 1u) && ri < 6 * nb6 + 2 * (nb2 - nb6)) {     // helper number t of leader L
                            const int lr = ri < 6 * nb6 ? ri / 6 : nb6 + (ri - 6 * nb6) / 2;
                            const int t = ri < 6 * nb6 ? ri % 6 : (ri - 6 * nb6) % 2;
                            const int L = nth_bit(busy, lr);
                            const BSlot &Ls = sh.slot[L];
                            const double l = Ls.left, r = Ls.right, m = Ls.m;
                            const double mA = (l + m) / 2.0, mB = (m + r) / 2.0;
                            double pt;
                            switch (t) {
                                case 0: pt = mA; break;
                                case 1: pt = mB; break;
                                case 2: pt = (l + mA) / 2.0; break;
                                case 3: pt = (mA + m) / 2.0; break;
                                case 4: pt = (m + mB) / 2.0; break;
                                default: pt = (mB + r) / 2.0; break;
                            }
                            sh.grp_pt[L][t] = pt;
                            sh.hot[warp] = make_double2(Ls.sigma, pt);
                            sh.skip[warp] = (int)(Ls.i - 1);
                        }
                        sh.grpn[warp] = 0;
                    }
                }
                __syncthreads();
            } else if (warp < NSLOT && lane == 0) {
                sh.grpn[warp] = 0;
            }
        }

        // ================= per-thread copies of the sweep parameters =================
        unsigned actmask = 0;
        bool waiting = false;
        unsigned fixmask = 0;
        int fmin = 0x7fffffff, fmax_ = -1;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
            const int sk = sh.skip[j];
            const int aj = sh.act[j];
            if (aj & 1) actmask |= 1u << j;
            if (aj == 2) waiting = true;
            if (sk >= 0) {
                static_assert((BTL & (BTL - 1)) == 0 && BTL % (2 * BT) == 0, "tile / owner of a pole by shifts and masks");
                const int t_i = sk / BTL;
                const int owner = (sk & (2 * BT - 1)) >> 1;
                if (owner == tid) { fixmask |= 1u << j; fmin = min(fmin, t_i); fmax_ = max(fmax_, t_i); }
            }
        }
        if (actmask == 0) {
            if (!waiting) break;
            __nanosleep(2000);     // only slots waiting for queued work: look again in a moment
            __syncthreads();
            continue;
        }
        double acc[NSLOT];
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
        AH_TG(tg_param);

        // ================= the sweep =================
        for (int tile = 0; tile < ntiles; ++tile, ++g) {
            if (tid == 0) { issue_tile(g_issue); if (++g_issue == 0u) wrapped = true; }
            const int s = (int)(g & (BS - 1));
            mbar_wait(&sh.full[s], (g / BS) & 1u);
            double Dv[BE], W2[BE];
#pragma unroll
            for (int e = 0; e < BE; e += 2) {
                const double2 d2 = *reinterpret_cast<const double2 *>(&sh.tileD[s][elem_offset(tid, e)]);
                const double2 w2 = *reinterpret_cast<const double2 *>(&sh.tileW[s][elem_offset(tid, e)]);
                Dv[e] = d2.x; Dv[e + 1] = d2.y; W2[e] = w2.x; W2[e + 1] = w2.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[s]);

            // rare: this thread holds the shift pole (d = 0) of some slot(s) in this tile.  The main block
            // below turns that slot's accumulator into NaN; save it here and redo the slot's four terms
            // without the pole afterwards, in the same order as every other thread -> the summation order
            // of a slot never depends on which other eigenpairs share the CTA (bitwise reproducible).
            const bool fix = fixmask != 0 && tile >= fmin && tile <= fmax_;
            if (fix) {
#pragma unroll
                for (int j = 0; j < NSLOT; ++j)
                    if ((fixmask >> j) & 1u) sh.save[j] = acc[j];
            }

            // ---- the branch-free block: BE x NSLOT independent 8-deep chains.  (Idle slots -- end of the work list,
            // Remedy-3 round -- are evaluated on dummy parameters: skipping them with CTA-uniform branches was
            // measured and costs the full-occupancy case more than the tail gains.)
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
                // Few eigenpairs, spread thinly over the CTAs: skip idle slots with a CTA-uniform branch.
                // (The same test costs the full-occupancy loop 2-4 %: measured.)
                if (ROUND == 0 || ((actmask >> j) & 1u)) {
                    const double2 h = sh.hot[j];   // (sigma_j, m_j): broadcast LDS.128, reloaded per tile to save 32 registers
                    double t = secular_term_first(dshift<KIND>(Dv[0], h.x), h.y, W2[0]);    // tile sum first, then the running sum
#pragma unroll
                    for (int e = 1; e < BE; ++e)
                        t = secular_term_acc(t, dshift<KIND>(Dv[e], h.x), h.y, W2[e]);
                    acc[j] = __dadd_rn(acc[j], t);
                }
            }

            if (fix) {
                const int base = tile * BTL;
#pragma unroll
                for (int j = 0; j < NSLOT; ++j) {
                    if ((fixmask >> j) & 1u) {
                        const int sk = sh.skip[j];
                        if (sk / BTL == tile) {
                            const double2 h = sh.hot[j];
                            double t = 0.0;                    // the tile sum without the pole: its other three terms in order
                            bool have = false;
#pragma unroll
                            for (int e = 0; e < BE; ++e)
                                if (base + elem_offset(tid, e) != sk) {
                                    t = have ? secular_term_acc(t, dshift<KIND>(Dv[e], h.x), h.y, W2[e])
                                             : secular_term_first(dshift<KIND>(Dv[e], h.x), h.y, W2[e]);
                                    have = true;
                                }
                            acc[j] = __dadd_rn(sh.save[j], t);
                        }
                    }
                }
            }
        }

        AH_TG(tg_tiles);
        // ================= reduce this sweep's partial sums =================
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
            const double v = warp_allreduce(acc[j], OpSum());
            if (lane == 0) sh.part[j][warp] = v;
        }
        AH_TG(tg_reduce);
#ifdef AH_DIAG_TIMING
        if (tid == 0) ++tg_sweeps;
#endif
        __syncthreads();
    }
#ifdef AH_DIAG_TIMING
    if (tid == 0 && ROUND == 0) {
        atomicAdd(dcount + 152, (unsigned long long)tg_boundary); atomicAdd(dcount + 153, (unsigned long long)tg_b2);
        atomicAdd(dcount + 154, (unsigned long long)tg_param); atomicAdd(dcount + 155, (unsigned long long)tg_tiles);
        atomicAdd(dcount + 156, (unsigned long long)tg_reduce); atomicAdd(dcount + 157, (unsigned long long)tg_b1);
        atomicAdd(dcount + 158, (unsigned long long)tg_sweeps); atomicAdd(dcount + 159, 1ull);
        const long long tb[12] = {tb_decide, tb_final, tb_swapchk, tb_requeue, tb_ticket, tb_resolve, tb_load, tb_n_final, tb_n_swapchk, tb_n_requeue,
                                  tb_n_ticket, tb_n_poll};
        for (int q = 0; q < 12; ++q) atomicAdd(dcount + 160 + q, (unsigned long long)tb[q]);
    }
#endif

    // drain the prefetched tiles before the CTA (and its shared memory) goes away
    if (tid == 0) {
        for (uint32_t gi = g; gi != g_issue; ++gi) mbar_wait(&sh.full[(int)(gi & (BS - 1))], (gi / BS) & 1u);
    }
}

// =====================================================================================================
// k_rem3_prep: Remedy 3 (arrowhead3.jl:466-496, arrowhead4.jl:333-368, arrowhead6.jl:429-470), first half:
// the new shift sigma_1 between the poles, the inverse there (a DPR1: only rho, K_rho and the bisection
// bracket are formed, by inv_at_value of ah_full.cuh) -> KState for k_bisect round 1.  One CTA per listed
// eigenpair (~1 % of all).  Anything beyond the plain binary64 branch (sigma_1 hits a pole, double-double rho)
// is handed to k_full.
constexpr int RT = 256;
static_assert(RT == BT, "the direct bisection reproduces k_bisect's summation order thread by thread");

#ifndef AH_REM3_RU
#define AH_REM3_RU 4
#endif
#ifndef AH_REM3_MINB
#define AH_REM3_MINB 1
#endif
template <int KIND>
__global__ void __launch_bounds__(RT, AH_REM3_MINB)
k_rem3_prep(Problem P, Outputs O, KState *__restrict__ ks, int64_t kbase, int64_t *__restrict__ rlist, int64_t *__restrict__ dlist,
            unsigned long long *dcount, double th2) {
    __shared__ double red[2 * (RT / 32) + 2];
    __shared__ double part2[2][RT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nwork = (int64_t)dcount[2];
    // Short lists (small k-ranges: multi-GPU shards, host-staged blocks): the persistent 8-slot kernel would spend a
    // whole latency-bound wave on a handful of eigenpairs, so each CTA bisects its eigenpair itself with plain
    // global loads -- in EXACTLY k_bisect's summation order (same thread -> pole map, same reduction trees), so the
    // result does not depend on which of the two ways was taken (bitwise shard-invariance is tested).
    const bool direct = nwork <= REM3_DIRECT_MAX;
    const int ntiles = (int)(P.Np / BTL);
    for (int64_t wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const int64_t k = rlist[wi];
        KState &G = ks[k - kbase];
        const double nu = G.nu, nu1 = G.nu1, sigma = G.sigma;
        const bool sideR = G.sideR != 0;
        double sigma1, shift;
        if (KIND == KIND_HALF) {
            sigma1 = __dadd_rn(__ddiv_rn(__dadd_rn(nu1, nu), __dmul_rn(__dmul_rn(2.0, nu), nu1)), __dmul_rn(sigma, sigma));
            shift = sqrt(sigma1);
        } else {
            sigma1 = __dadd_rn(__ddiv_rn(__dadd_rn(nu1, nu), __dmul_rn(__dmul_rn(2.0, nu), nu1)), sigma);
            shift = sigma1;
        }
        bool full = is_pole(P.D, P.N, shift) || !isfinite(shift);
        InvValue R;
        int status = 0;
        if (!full) {   // CTA-uniform
            inv_at_value<KIND, RT>(P, shift, P.tau[3], R, status, red);
            full = (R.Qout != 0) || (status != 0) || !isfinite(R.rho);
        }
        double left = 0.0, right = 0.0, nu1new = 0.0;
        if (!full) {
            const double ruu = __dmul_rn(R.rho, R.uu);   // bracket of bisect(::SymDPR1,side), arrowhead3.jl:722-726
            if (R.rho > 0.0) {
                if (!sideR) { left = R.mn1; right = R.mn2; }
                else { left = R.mx1; right = __dadd_rn(R.mx1, ruu); }
            } else {
                if (!sideR) { left = __dadd_rn(R.mn1, ruu); right = R.mn1; }
                else { left = R.mx2; right = R.mx1; }
            }
            nu1new = __dadd_rn(R.maxabsD, __dmul_rn(fabs(R.rho), R.uu));
        }
        if (full) {
            if (tid == 0) {
                G.flag = KS_HANDOFF;
                dlist[atomicAdd(dcount, 1ull)] = k;
            }
        } else if (!direct) {
            if (tid == 0) {
                G.sigma = shift; G.b = R.rho; G.krho = R.Krho;
                G.left = left; G.right = right;
                G.nu1 = nu1new;
                G.lambda = sigma1;
                G.flag = KS_REM3_READY;
            }
        } else {
            BSlot S;                                   // identical copy in every thread
            S.k = k; S.i = G.i; S.phase = PH_BISECT; S.sideR = sideR ? 1 : 0; S.steps = G.steps; S.q = 0;
            S.sigma = shift; S.wz2 = 1.0; S.b = R.rho; S.left = left; S.right = right; S.m = (left + right) / 2.0;
            S.nu1 = nu1new; S.Kb = G.Kb; S.Kz = G.Kz; S.knu0 = G.knu; S.krho = R.Krho; S.sig2 = sigma1;
            __syncthreads();                           // every thread has read G before thread 0 rewrites it in the finalize
            int par = 0;
            bool nanfail = false;
            const bool accel = th2 > 0.0;
            accel_reset(S, accel ? accel_certifiable<KIND, 1>(P, S) : 0);    // (every thread: identical scalar work)
            while (true) {
                if (accel) accel_descend(S);
                if (bisect_converged(S)) break;
                if (accel) accel_choose<1>(S, th2); else S.x = S.m;
                // One CTA per eigenpair: a sweep is latency-bound on the L2 loads unless several tiles are in flight.  The
                // loads of RU tiles are issued together and the next group is fetched while this one is evaluated; the
                // terms are still accumulated tile by tile, pair by pair -- k_bisect's order, bit for bit.
                constexpr int RU = AH_REM3_RU, RH = BE / 2;
                double acc = 0.0;
                double2 dcur[RU][RH], wcur[RU][RH], dnxt[RU][RH], wnxt[RU][RH];
                auto fetch = [&](int tile0, double2 (&dd)[RU][RH], double2 (&ww)[RU][RH]) {
#pragma unroll
                    for (int u = 0; u < RU; ++u) {
                        const int tile = min(tile0 + u, ntiles - 1);      // (a clamped tile is loaded but never accumulated)
                        const double *Dt = P.D + (int64_t)tile * BTL, *Wt = P.w2 + (int64_t)tile * BTL;
#pragma unroll
                        for (int h = 0; h < RH; ++h) { dd[u][h] = ldg2(Dt + elem_offset(tid, 2 * h)); ww[u][h] = ldg2(Wt + elem_offset(tid, 2 * h)); }
                    }
                };
                fetch(0, dcur, wcur);
                for (int tile0 = 0; tile0 < ntiles; tile0 += RU) {
                    if (tile0 + RU < ntiles) fetch(tile0 + RU, dnxt, wnxt);
#pragma unroll
                    for (int u = 0; u < RU; ++u) {
                        if (tile0 + u < ntiles) {
                            static_assert(RH == 2, "tile sum of four terms, as in k_bisect");
                            double t = secular_term_first(dshift<KIND>(dcur[u][0].x, S.sigma), S.x, wcur[u][0].x);
                            t = secular_term_acc(t, dshift<KIND>(dcur[u][0].y, S.sigma), S.x, wcur[u][0].y);
                            t = secular_term_acc(t, dshift<KIND>(dcur[u][1].x, S.sigma), S.x, wcur[u][1].x);
                            t = secular_term_acc(t, dshift<KIND>(dcur[u][1].y, S.sigma), S.x, wcur[u][1].y);
                            acc = __dadd_rn(acc, t);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < RU; ++u)
#pragma unroll
                        for (int h = 0; h < RH; ++h) { dcur[u][h] = dnxt[u][h]; wcur[u][h] = wnxt[u][h]; }
                }
                const double v = warp_allreduce(acc, OpSum());
                if (lane == 0) part2[par][warp] = v;
                __syncthreads();
                double sum = (lane < BW) ? part2[par][lane] : 0.0;
#pragma unroll
                for (int o = BW / 2; o > 0; o >>= 1) {
                    const double p = __shfl_xor_sync(0xffffffffu, sum, o);
                    sum = (lane & o) ? __dadd_rn(p, sum) : __dadd_rn(sum, p);
                }
                sum = __shfl_sync(0xffffffffu, sum, 0);
                par ^= 1;
                if (accel) {
                    S.steps += 1;
                    if (sum != sum) {
                        if (S.x == S.m) { nanfail = true; break; }
                        S.probing = 0;
                    } else {
                        accel_update<KIND, 1>(S, sum, th2);
                    }
                    continue;
                }
                if (sum != sum) { nanfail = true; break; }   // as in k_bisect: k_full redoes this eigenpair
                const double m = S.m;
                if (KIND != KIND_DPR1) sum = __dadd_rn(sum, __ddiv_rn(1.0, __dadd_rn(0.0, -m)));
                const double F = __dadd_rn(1.0, __dmul_rn(S.b, sum));
                if (sign_jl(S.b) * F < 0.0) S.left = m; else S.right = m;
                S.m = (S.left + S.right) / 2.0;
                S.steps += 1;
            }
            if (nanfail) {
                if (tid == 0) { G.flag = KS_HANDOFF; dlist[atomicAdd(dcount, 1ull)] = k; }
            } else {
                if (tid == 0) { G.sigma = shift; G.krho = R.Krho; }   // k_vec reads sigma (the new shift), mu, lambda
                bisect_finalize<KIND, 1>(P, O, ks, S, kbase, dlist, rlist, dcount, tid == 0 ? 0 : 1);
            }
        }
        __syncthreads();
    }
}

// =====================================================================================================
// k_vec: eigenvector norm + normalised write, VJ consecutive eigenpairs per CTA
constexpr int VT = 256, VW = VT / 32, VJ = 4;

template <int KIND>
__global__ void __launch_bounds__(VT, 2)
k_vec(Problem P, Outputs O, const KState *__restrict__ ks, int64_t kbase, int64_t cnt, const int64_t *__restrict__ klist,
      const unsigned long long *nlist_dev, int want_flag) {
    // dense mode (klist == nullptr): group g = eigenpairs kbase + VJ*g .. +VJ-1 with flag == want_flag (the ones round 0
    // finished); list mode: the eigenpairs of klist (the Remedy-3 list, finished by round 1), grid-stride over groups.
    __shared__ double red[VJ][VW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t N = P.N, Np = P.Np;
    const double *__restrict__ D = P.D;
    const double *__restrict__ W = P.w;
    const int64_t nitems = klist ? (int64_t)*nlist_dev : cnt;
    for (int64_t c0 = (int64_t)blockIdx.x * VJ; c0 < nitems; c0 += (int64_t)gridDim.x * VJ) {
    if (c0 != (int64_t)blockIdx.x * VJ) __syncthreads();    // red[] is reused

    bool valid[VJ];
    int64_t cix[VJ];                                         // index of the eigenpair within the block (k - kbase)
    double sigma[VJ], mu[VJ], lambda[VJ];
    bool any = false;
#pragma unroll
    for (int j = 0; j < VJ; ++j) {
        valid[j] = c0 + j < nitems;
        cix[j] = valid[j] ? (klist ? klist[c0 + j] - kbase : c0 + j) : 0;
        valid[j] = valid[j] && ks[cix[j]].flag == want_flag;
        any = any || valid[j];
        if (valid[j]) { sigma[j] = ks[cix[j]].sigma; mu[j] = ks[cix[j]].mu; lambda[j] = ks[cix[j]].lambda; }
        else { sigma[j] = 0.0; mu[j] = -1.0e300; lambda[j] = 1.0; }
    }
    if (!any) continue;

    // ---- VNORM (the norm is a sum over N entries: <=1-ulp reciprocal; the entries themselves are divided exactly below)
    double ss[VJ];
#pragma unroll
    for (int j = 0; j < VJ; ++j) ss[j] = 0.0;
    double2 d2n = ldg2(D + 2 * tid), w2n = ldg2(W + 2 * tid);
    const int Ni = (int)N, Npi = (int)Np;
    for (int l0 = 2 * tid; l0 < Npi; l0 += 2 * VT) {
        const double2 d2 = d2n, w2 = w2n;
        if (l0 + 2 * VT < Npi) { d2n = ldg2(D + l0 + 2 * VT); w2n = ldg2(W + l0 + 2 * VT); }   // prefetch the next pair
        const double dv[2] = {d2.x, d2.y};
        const double wv[2] = {w2.x, w2.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (l0 + h < Ni) {
#pragma unroll
                for (int j = 0; j < VJ; ++j) {
                    const double v = __dmul_rn(wv[h], rcp_fast(__dadd_rn(dshift<KIND>(dv[h], sigma[j]), -mu[j])));
                    ss[j] = __fma_rn(v, v, ss[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < VJ; ++j) {
        const double v = warp_allreduce(ss[j], OpSum());
        if (lane == 0) red[j][warp] = v;
    }
    __syncthreads();
    double inrm[VJ];
#pragma unroll
    for (int j = 0; j < VJ; ++j) {
        double s = red[j][0];
        for (int w = 1; w < VW; ++w) s = __dadd_rn(s, red[j][w]);
        if (KIND != KIND_DPR1) s = __dadd_rn(s, 1.0);
        inrm[j] = __ddiv_rn(1.0, sqrt(s));
    }

    // ---- VWRITE
    const int64_t *umap = O.rowmap_u;
    const int64_t *vmap = (KIND == KIND_HALF) ? O.rowmap_v : O.rowmap_u;
    double *ucol[VJ], *vcol[VJ];
#pragma unroll
    for (int j = 0; j < VJ; ++j) {
        const int64_t k = kbase + cix[j];
        const int64_t colidx = (valid[j] && umap) ? umap[k - 1] : cix[j];
        ucol[j] = (valid[j] && O.U) ? O.U + colidx * O.ldu : nullptr;
        vcol[j] = (KIND == KIND_HALF && valid[j] && O.V) ? O.V + colidx * O.ldv : nullptr;
    }
    d2n = ldg2(D + 2 * tid); w2n = ldg2(W + 2 * tid);
    if (KIND != KIND_HALF && !umap) {
        // common case (no deflation row map): per-column store mode decided once -- 0: skip, 1: 16-byte stores (column
        // start 16-byte aligned; l0 is even), 2: two 8-byte stores (odd leading dimension -> every other column)
        int mode[VJ];
#pragma unroll
        for (int j = 0; j < VJ; ++j) mode[j] = !ucol[j] ? 0 : (((reinterpret_cast<uintptr_t>(ucol[j]) & 15) == 0) ? 1 : 2);
        for (int l0 = 2 * tid; l0 < Ni; l0 += 2 * VT) {
            const double2 d2 = d2n, w2 = w2n;
            if (l0 + 2 * VT < Npi) { d2n = ldg2(D + l0 + 2 * VT); w2n = ldg2(W + l0 + 2 * VT); }
            const bool pair = l0 + 1 < Ni;
            // all 2 x VJ quotients branch-free and interleaved; the exact division only where one leaves the normal range
            double q0[VJ], q1[VJ];
            bool ok = true;
#pragma unroll
            for (int j = 0; j < VJ; ++j) {
                q0[j] = div_fast(w2.x, __dadd_rn(dshift<KIND>(d2.x, sigma[j]), -mu[j]), ok);
                q1[j] = div_fast(w2.y, __dadd_rn(dshift<KIND>(d2.y, sigma[j]), -mu[j]), ok);
            }
            if (!ok) {
#pragma unroll
                for (int j = 0; j < VJ; ++j) {
                    q0[j] = __ddiv_rn(w2.x, __dadd_rn(dshift<KIND>(d2.x, sigma[j]), -mu[j]));
                    q1[j] = __ddiv_rn(w2.y, __dadd_rn(dshift<KIND>(d2.y, sigma[j]), -mu[j]));
                }
            }
#pragma unroll
            for (int j = 0; j < VJ; ++j) {
                const double v0 = __dmul_rn(q0[j], inrm[j]);
                const double v1 = __dmul_rn(q1[j], inrm[j]);
                if (mode[j] == 1 && pair) {
                    *reinterpret_cast<double2 *>(ucol[j] + l0) = make_double2(v0, v1);
                } else if (mode[j] != 0) {
                    ucol[j][l0] = v0;
                    if (pair) ucol[j][l0 + 1] = v1;
                }
            }
        }
    } else
    for (int64_t l0 = 2 * tid; l0 < Np; l0 += 2 * VT) {
        if (l0 >= N) break;
        const double2 d2 = d2n, w2 = w2n;
        if (l0 + 2 * VT < Np) { d2n = ldg2(D + l0 + 2 * VT); w2n = ldg2(W + l0 + 2 * VT); }
        const double dv[2] = {d2.x, d2.y};
        const double wv[2] = {w2.x, w2.y};
        double rs[2] = {1.0, 1.0};
        if (KIND == KIND_HALF && O.rowscale_v) {
            rs[0] = O.rowscale_v[l0];
            if (l0 + 1 < N) rs[1] = O.rowscale_v[l0 + 1];
        }
        double q[VJ][2];
        bool ok = true;
#pragma unroll
        for (int j = 0; j < VJ; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) q[j][h] = div_fast(wv[h], __dadd_rn(dshift<KIND>(dv[h], sigma[j]), -mu[j]), ok);
        if (!ok) {
#pragma unroll
            for (int j = 0; j < VJ; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) q[j][h] = __ddiv_rn(wv[h], __dadd_rn(dshift<KIND>(dv[h], sigma[j]), -mu[j]));
        }
#pragma unroll
        for (int j = 0; j < VJ; ++j) {
            double *const vec = (KIND == KIND_HALF) ? vcol[j] : ucol[j];
            double v[2], vo[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                v[h] = __dmul_rn(q[j][h], inrm[j]);
                vo[h] = (KIND == KIND_HALF && O.rowscale_v) ? __dmul_rn(v[h], rs[h]) : v[h];
            }
            if (vec) {
                if (!vmap && l0 + 1 < N && ((reinterpret_cast<uintptr_t>(vec + l0) & 15) == 0)) {
                    *reinterpret_cast<double2 *>(vec + l0) = make_double2(vo[0], vo[1]);
                } else {
                    vec[vmap ? vmap[l0] : l0] = vo[0];
                    if (l0 + 1 < N) vec[vmap ? vmap[l0 + 1] : l0 + 1] = vo[1];
                }
            }
            if (KIND == KIND_HALF && ucol[j]) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (l0 + h < N) ucol[j][umap ? umap[l0 + h] : l0 + h] = __ddiv_rn(__dmul_rn(lambda[j], v[h]), dv[h]);
            }
        }
    }
    if (KIND != KIND_DPR1 && tid < VJ) {
#pragma unroll
        for (int j = 0; j < VJ; ++j) {
            if (j != tid || !valid[j]) continue;
            double *const vec = (KIND == KIND_HALF) ? vcol[j] : ucol[j];
            const double vt = __dmul_rn(-1.0, inrm[j]);
            if (vec) vec[vmap ? vmap[N] : N] = vt;
            if (KIND == KIND_HALF && ucol[j] && P.has_tip) ucol[j][umap ? umap[N] : N] = __ddiv_rn(__dmul_rn(P.ztip, vt), lambda[j]);
        }
    }
    }   // groups
}

// =====================================================================================================
inline int configure_pipe_kernels() {
    const int smem = (int)sizeof(BisectShared);
    cudaError_t e;
#define AH_SET_SMEM(K) if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return (int)e
#define AH_SET_KIND(KD) AH_SET_SMEM((k_bisect<KD, 0, 8, 0>)); AH_SET_SMEM((k_bisect<KD, 0, 7, 0>)); AH_SET_SMEM((k_bisect<KD, 0, 6, 0>)); AH_SET_SMEM((k_bisect<KD, 1, 8, 0>)); \
    AH_SET_SMEM((k_bisect<KD, 0, 8, 1>)); AH_SET_SMEM((k_bisect<KD, 0, 7, 1>)); AH_SET_SMEM((k_bisect<KD, 0, 6, 1>)); AH_SET_SMEM((k_bisect<KD, 1, 8, 1>))
    AH_SET_KIND(KIND_ARROW); AH_SET_KIND(KIND_DPR1); AH_SET_KIND(KIND_HALF);
#undef AH_SET_KIND
#undef AH_SET_SMEM
    return 0;
}

// P.D / P.w / P.w2 must be padded to P.Np, a multiple of TILE_POLES (see run_solver in ah_api.cu).
inline int launch_prep(int kind, const Problem &P, KState *ks, int64_t kbase, int64_t cnt, int64_t *dlist,
                       unsigned long long *dcount, cudaStream_t st) {
    const int grid = (int)((cnt + PJ - 1) / PJ);
    switch (kind) {
        case KIND_ARROW: k_prep<KIND_ARROW><<<grid, PT, 0, st>>>(P, ks, kbase, cnt, dlist, dcount); break;
        case KIND_DPR1: k_prep<KIND_DPR1><<<grid, PT, 0, st>>>(P, ks, kbase, cnt, dlist, dcount); break;
        default: k_prep<KIND_HALF><<<grid, PT, 0, st>>>(P, ks, kbase, cnt, dlist, dcount); break;
    }
    return (int)cudaGetLastError();
}
// round 0: all READY eigenpairs of the block; round 1: the Remedy-3 list (its length is read on the device)
// requeue buffer entries k_bisect round 0 may need for a block of cnt eigenpairs (it stops time-slicing when the buffer is full)
inline size_t bisect_queue_entries(int64_t cnt) { return (size_t)cnt * 24 + 8192; }

// quantum: steps an eigenpair keeps its slot while others wait (0: no time slicing, < 0: chosen here);
// window: time slicing starts when fewer than window x (all slots) eigenpairs are waiting; tag: unique per launch, != 0
// relative bound used by the certificates of the accelerated bisection: 2.01 x (roundings per accumulator + 8 reduction
// levels + reciprocal (<= 1.6 ulp) + slack) x 2^-53; 0 = not available (N > 2^21: the bound would no longer be small)
inline double bisect_th2(const Problem &P) {
    if (P.N > ((int64_t)1 << 21)) return 0.0;
    // per accumulator: BE roundings inside a tile sum + one per tile for the running sum; 8 reduction levels; <= 1.6 ulp per
    // reciprocal; slack
    const double depth = (double)(P.Np / BTL) + BE;
    return 2.01 * (depth + 8.0 + 1.6 + 4.4) * 1.1102230246251565e-16;
}
inline int launch_bisect(int kind, int round, const Problem &P, const Outputs &O, KState *ks, int64_t kbase, int64_t cnt,
                         int64_t *dlist, int64_t *rlist, unsigned long long *dcount, unsigned long long *queue, long long qcap,
                         unsigned tag, int quantum, int window, int spec, int sm_count, cudaStream_t st, KExtra *kx = nullptr, int accel = 0) {
    const int smem = (int)sizeof(BisectShared);
    const int64_t cap = (int64_t)sm_count * BCTAS;
    // Round 0: 8 slots per CTA unless the block has fewer than 8 (7) eigenpairs per CTA of a full grid -- then 7 (6)
    // slots, and no more CTAs than groups of that size.  Round 1 (list length known to the device only): one
    // eigenpair per CTA where possible.
    const int nslot = (round == 1 || cnt > cap * (BJ - 1)) ? BJ : (cnt > cap * (BJ - 2) ? BJ - 1 : BJ - 2);
    const int64_t want = (round == 1) ? cnt : (cnt + nslot - 1) / nslot;
    // Round 1 runs beside k_vec (two CTAs of 30 K registers per SM): one CTA per SM leaves room for one of them
    // (measured: the k_vec span shrinks from 2.54 to 2.40 ms at n = 32768; half a CTA per SM makes the round itself too long)
    const int64_t gcap = (round == 1) ? (int64_t)sm_count * AH_R1_CTAS_PER_SM : cap;
    const int grid = (int)(want < gcap ? (want < 1 ? 1 : want) : gcap);
    if (quantum < 0) {   // automatic time slice
        const int64_t ntiles = P.Np / BTL;
        // a switch costs a few microseconds of latency per slot: keep it small against quantum sweeps of ntiles tiles;
        // and nothing ever waits when every eigenpair has a slot of its own from the start
        quantum = (cnt <= (int64_t)grid * nslot) ? 0 : (int)std::min<int64_t>(32, std::max<int64_t>(8, 128 / std::max<int64_t>(1, ntiles)));
    }
    const double th2 = (accel && (kx || round == 1)) ? bisect_th2(P) : 0.0;
#define AH_LAUNCH_BISECT(KD, RD, NS, AC) k_bisect<KD, RD, NS, AC><<<grid, BT, smem, st>>>(P, O, ks, kbase, cnt, dlist, rlist, dcount, queue, qcap, tag, quantum, window, spec, kx, th2)
#define AH_LAUNCH_KIND(RD, NS, AC)                                          \
    switch (kind) {                                                         \
        case KIND_ARROW: AH_LAUNCH_BISECT(KIND_ARROW, RD, NS, AC); break;   \
        case KIND_DPR1: AH_LAUNCH_BISECT(KIND_DPR1, RD, NS, AC); break;     \
        default: AH_LAUNCH_BISECT(KIND_HALF, RD, NS, AC); break;            \
    }
    if (round == 1) {
        if (th2 > 0.0) { AH_LAUNCH_KIND(1, 8, 1) }
        else { AH_LAUNCH_KIND(1, 8, 0) }
    } else if (th2 > 0.0) {
        if (nslot == 8) { AH_LAUNCH_KIND(0, 8, 1) }
        else if (nslot == 7) { AH_LAUNCH_KIND(0, 7, 1) }
        else { AH_LAUNCH_KIND(0, 6, 1) }
    }
    else if (nslot == 8) { AH_LAUNCH_KIND(0, 8, 0) }
    else if (nslot == 7) { AH_LAUNCH_KIND(0, 7, 0) }
    else { AH_LAUNCH_KIND(0, 6, 0) }
#undef AH_LAUNCH_KIND
#undef AH_LAUNCH_BISECT
    return (int)cudaGetLastError();
}
inline int launch_rem3_prep(int kind, const Problem &P, const Outputs &O, KState *ks, int64_t kbase, int64_t cnt, int64_t *rlist,
                            int64_t *dlist, unsigned long long *dcount, int sm_count, cudaStream_t st, int accel = 0) {
    const int grid = (int)std::min<int64_t>(cnt, (int64_t)sm_count * 4);
    const double th2 = accel ? bisect_th2(P) : 0.0;
    switch (kind) {
        case KIND_ARROW: k_rem3_prep<KIND_ARROW><<<grid, RT, 0, st>>>(P, O, ks, kbase, rlist, dlist, dcount, th2); break;
        case KIND_DPR1: k_rem3_prep<KIND_DPR1><<<grid, RT, 0, st>>>(P, O, ks, kbase, rlist, dlist, dcount, th2); break;
        default: k_rem3_prep<KIND_HALF><<<grid, RT, 0, st>>>(P, O, ks, kbase, rlist, dlist, dcount, th2); break;
    }
    return (int)cudaGetLastError();
}
// klist == nullptr: the eigenpairs round 0 finished (flag KS_DONE); else the Remedy-3 list (flag KS_DONE1, length on the device)
inline int launch_vec(int kind, const Problem &P, const Outputs &O, const KState *ks, int64_t kbase, int64_t cnt,
                      const int64_t *klist, const unsigned long long *nlist_dev, int sm_count, cudaStream_t st) {
    int grid = (int)((cnt + VJ - 1) / VJ);
    if (klist) grid = (int)std::min<int64_t>(grid, (int64_t)sm_count * 2);
    const int flag = klist ? KS_DONE1 : KS_DONE;
    switch (kind) {
        case KIND_ARROW: k_vec<KIND_ARROW><<<grid, VT, 0, st>>>(P, O, ks, kbase, cnt, klist, nlist_dev, flag); break;
        case KIND_DPR1: k_vec<KIND_DPR1><<<grid, VT, 0, st>>>(P, O, ks, kbase, cnt, klist, nlist_dev, flag); break;
        default: k_vec<KIND_HALF><<<grid, VT, 0, st>>>(P, O, ks, kbase, cnt, klist, nlist_dev, flag); break;
    }
    return (int)cudaGetLastError();
}

}  // namespace ah
