// ah_full.cuh -- the complete per-eigenpair solver, one CTA per eigenpair.
//
// This kernel implements every branch of the reference's per-k solvers that can be done in
// binary64 + double-double: shift selection, shifted inverse with the Kb/Kz test and the Dekker
// double-double recomputation of b, bisection, the Knu test with Remedy 3 (inverse at a shift
// between poles -> DPR1, double-double rho), eigenvector, Remedy 1 (Rayleigh quotient).
//   SymArrow : arrowhead3.jl:419-550 with inv! :152-244, inv(A,sigma) :260-332, bisect :667-744
//   SymDPR1  : arrowhead4.jl:289-422 with inv! :89-159,  inv(A,sigma) :173-225
//   HalfArrow: arrowhead6.jl:344-508 with inv! :144-213, inv(A,sigma) :227-275
// It is the path for the eigenpairs the fast kernel (ah_fast.cuh) hands over (Remedy 3 / 1,
// double-double, |nu|=Inf) and for ah_set_mode(ctx,1).  BigFloat branches (Qout 50/100) and the
// pole-return sub-path, which raise in the reference, are reported through the status word.
//
// Data layout: D[N], w[N] in global memory (L2-resident, read-only); nothing of size N is ever
// materialised per eigenpair -- the inverse's entries D'_l = 1/d_l, z'_l are recomputed from the
// pole differences d_l in every sweep.  All scalar state is replicated in registers of every thread
// and derived only from block-wide all-reductions, so control flow is CTA-uniform.
#pragma once
#include "../../include/arrowhead_b200.h"
#include "ah_device.cuh"

namespace ah {

enum : int { KIND_ARROW = 0, KIND_DPR1 = 1, KIND_HALF = 2 };
// status word: bits 0..5 = AH_ST_* conditions, bits 8..14 = AH_BR_* branch trace (include/arrowhead_b200.h)

struct Problem {
    int64_t N;         // number of poles
    const double *D;   // [N] strictly decreasing
    const double *w;   // [N] z (arrow) | u (dpr1) | z.*D (half)
    const double *w2;  // [Np] w_l^2 (fast path)
    int64_t Np;        // D, w, w2 are padded to Np poles (multiple of the fast path's tile); pads: weight 0, D = pad
    double pad;        // pad pole: strictly below every real pole; pad - D_l is never a power of two (1 - m d never cancels exactly there)
    const double *z;   // [N] half only: the original z (double-double branch needs it)
    double a;          // arrow: arrow top; dpr1: rho (>0)
    double maxabs_w;   // max_l |w_l|
    double sumabs_w;   // sum_l |w_l| (dpr1 Kz)
    double zz;         // half: dot(z,z) over all nz entries
    double zz_hi, zz_lo;  // half: double-double sum of z_l^2 (all nz entries, sequential)
    double ztip;       // half: z[nd] when has_tip
    int has_tip;       // half
    double tau[5];
};

struct Outputs {
    double *lambda;
    int64_t *sind;
    double *kb, *kz, *knu, *krho;
    int64_t *qout;
    int32_t *steps, *status;
    double *U;
    int64_t ldu;
    const int64_t *rowmap_u;
    double *V;
    int64_t ldv;
    const int64_t *rowmap_v;
    const double *rowscale_v;
};

template <int KIND>
__device__ __forceinline__ double dshift(double Dl, double s) {
    if (KIND == KIND_HALF) return __dmul_rn(__dadd_rn(Dl, -s), __dadd_rn(Dl, s));
    return __dadd_rn(Dl, -s);
}
__device__ __forceinline__ double sign_jl(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }

// Cauchy interlacing: eigenvalue k (descending) of an irreducible arrow / DPR1 (rho > 0) / the k-th singular
// value of a half-arrow lies in [D_k, D_{k-1}] (D_0 = +inf, D_{N+1} = -inf).  The reference does not test this;
// its two-stage scheme (nu from the inverse at a pole, then Remedy 3 from that nu) can return an eigenvalue of
// the wrong end of the spectrum when the first stage is pure rounding noise (K_nu ~ 1/eps), so a violated
// bound is reported to the caller (AH_ST_INTERLACE) instead of passing silently.
__device__ __forceinline__ bool interlace_ok(const double *D, int64_t N, int64_t k, double lambda) {
    if (!(lambda == lambda)) return false;
    if (k <= N && lambda < D[k - 1]) return false;
    if (k >= 2 && lambda > D[k - 2]) return false;
    return true;
}

// merge two (best, second-best) pairs; MAXI: maxima, else minima
template <bool MAXI>
__device__ __forceinline__ void top2_merge(double &a1, double &a2, double b1, double b2) {
    if (MAXI) {
        double n1 = fmax(a1, b1);
        double n2 = fmax(fmin(a1, b1), fmax(a2, b2));
        a1 = n1; a2 = n2;
    } else {
        double n1 = fmin(a1, b1);
        double n2 = fmin(fmax(a1, b1), fmin(a2, b2));
        a1 = n1; a2 = n2;
    }
}
template <bool MAXI>
__device__ __forceinline__ void top2_push(double &a1, double &a2, double x) {
    if (MAXI) {
        if (x > a1) { a2 = a1; a1 = x; } else if (x > a2) a2 = x;
    } else {
        if (x < a1) { a2 = a1; a1 = x; } else if (x < a2) a2 = x;
    }
}
template <bool MAXI>
__device__ __forceinline__ void block_allreduce_top2(double &m1, double &m2, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double p1 = __shfl_xor_sync(0xffffffffu, m1, o);
        double p2 = __shfl_xor_sync(0xffffffffu, m2, o);
        top2_merge<MAXI>(m1, m2, p1, p2);
    }
    if (lane == 0) { scratch[2 * warp] = m1; scratch[2 * warp + 1] = m2; }
    __syncthreads();
    double r1 = scratch[0], r2 = scratch[1];
    for (int w = 1; w < nw; ++w) top2_merge<MAXI>(r1, r2, scratch[2 * w], scratch[2 * w + 1]);
    __syncthreads();
    m1 = r1; m2 = r2;
}

// is x bitwise-equal to some pole?  D strictly decreasing -> binary search (isequal semantics).
__device__ __forceinline__ bool is_pole(const double *D, int64_t N, double x) {
    int64_t lo = 0, hi = N - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        double d = D[mid];
        if (d == x) return __double_as_longlong(d) == __double_as_longlong(x);
        if (d > x) lo = mid + 1; else hi = mid - 1;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// the inverse at a shift that is not a pole: DPR1 (D'',u,rho); only its scalars are formed.
struct InvValue {
    double rho, Krho;
    int64_t Qout;
    double uu, maxabsD;
    double mx1, mx2, mn1, mn2;
};

template <int KIND, int BT>
__device__ void inv_at_value(const Problem &P, double sig, double tol, InvValue &R, int &status, double *red) {
    const int64_t N = P.N;
    const double *__restrict__ D = P.D;
    const double *__restrict__ W = P.w;
    double Ps = 0.0, Qs = 0.0, uu = 0.0, mabs = 0.0;
    double mx1 = -INFINITY, mx2 = -INFINITY, mn1 = INFINITY, mn2 = INFINITY;
#pragma unroll 4
    for (int64_t l = threadIdx.x; l < N; l += BT) {
        double w = W[l];
        double D1 = __ddiv_rn(1.0, dshift<KIND>(D[l], sig));
        double u1 = __dmul_rn(w, D1);
        double t = __dmul_rn(__dmul_rn(w, w), D1);
        if (D1 > 0.0) Ps = __dadd_rn(Ps, t); else Qs = __dadd_rn(Qs, t);
        uu = __dadd_rn(uu, __dmul_rn(u1, u1));
        mabs = fmax(mabs, fabs(D1));
        top2_push<true>(mx1, mx2, D1);
        top2_push<false>(mn1, mn2, D1);
    }
    Ps = block_allreduce(Ps, OpSum(), red);
    Qs = block_allreduce(Qs, OpSum(), red);
    uu = block_allreduce(uu, OpSum(), red);
    mabs = block_allreduce(mabs, OpMax(), red);
    block_allreduce_top2<true>(mx1, mx2, red);
    block_allreduce_top2<false>(mn1, mn2, red);
    if (KIND != KIND_DPR1) {  // slot (0,-1) at the arrow-top position
        uu = __dadd_rn(uu, 1.0);
        top2_push<true>(mx1, mx2, 0.0);
        top2_push<false>(mn1, mn2, 0.0);
    }
    double as_ = 0.0;
    dd ad = dd_from(0.0);
    if (KIND == KIND_ARROW) {
        as_ = __dadd_rn(P.a, -sig);
        ad = dd_sub(dd_from(P.a), dd_from(sig));
    } else if (KIND == KIND_HALF) {
        ad = dd_add(dd{P.zz_hi, P.zz_lo}, dd_mul(dd_from(sig), dd_from(-sig)));
        as_ = __dadd_rn(ad.hi, ad.lo);
    }
    if (KIND == KIND_DPR1) {
        double ir = __ddiv_rn(1.0, P.a);
        if (P.a > 0) Ps = __dadd_rn(Ps, ir); else Qs = __dadd_rn(Qs, ir);
    } else {
        if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
    }
    double PQ = __dadd_rn(Ps, Qs);
    double Krho = __ddiv_rn(__dadd_rn(Ps, -Qs), fabs(PQ));
    double rho;
    int64_t Qout = 0;
    if (Krho < tol) {
        rho = __ddiv_rn(-1.0, PQ);
    } else {
        Qout = 1;
        status |= AH_BR_DD_RHO;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        dd s1 = dd_from(sig);
        for (int64_t l = threadIdx.x; l < N; l += BT) {
            dd t;
            bool pos;
            if (KIND == KIND_HALF) {
                dd Dk = dd_from(D[l]);
                dd Dd = dd_mul(dd_add(Dk, s1), dd_sub(Dk, s1));
                dd zd = dd_mul(dd_from(P.z[l]), Dk);
                t = dd_div(dd_mul(zd, zd), Dd);
                pos = Dd.hi > 0.0;
            } else {
                dd wl = dd_from(W[l]);
                t = dd_div(dd_mul(wl, wl), dd_sub(dd_from(D[l]), s1));
                pos = __ddiv_rn(1.0, dshift<KIND>(D[l], sig)) > 0.0;
            }
            if (pos) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        Pd = block_allreduce_dd(Pd, red);
        Qd = block_allreduce_dd(Qd, red);
        if (KIND == KIND_DPR1) {
            dd ir = dd_div(dd_from(1.0), dd_from(P.a));
            if (P.a > 0) Pd = dd_add(Pd, ir); else Qd = dd_add(Qd, ir);
        } else {
            if (__dadd_rn(ad.hi, ad.lo) < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        }
        dd S = dd_add(Pd, Qd);
        if (KIND == KIND_ARROW) {
            bool pluszero = (__double_as_longlong(S.hi) == 0) && (__double_as_longlong(S.lo) == 0);
            if (!pluszero) {
                Krho = dd_to_double(dd_div(dd_sub(Pd, Qd), dd_abs(S)));
                dd r = dd_div(dd_from(1.0), S);
                rho = -__dadd_rn(r.hi, r.lo);
            } else {
                Qout = 100;
                status |= 2;
                rho = __ddiv_rn(-1.0, PQ);
            }
        } else if (KIND == KIND_DPR1) {
            Krho = dd_to_double(dd_div(dd_sub(Pd, Qd), dd_abs(S)));
            if (!(Krho < tol / kEps)) {
                Qout = 100;
                status |= 2;
            }
            dd r = dd_div(dd_from(1.0), S);
            rho = -__dadd_rn(r.hi, r.lo);
        } else {
            dd r = dd_div(dd_from(1.0), S);
            rho = -__dadd_rn(r.hi, r.lo);
        }
    }
    R.rho = rho; R.Krho = Krho; R.Qout = Qout; R.uu = uu; R.maxabsD = mabs;
    R.mx1 = mx1; R.mx2 = mx2; R.mn1 = mn1; R.mn2 = mn2;
}

// bisect(::SymDPR1,side) on the implicit inverse (arrowhead3.jl:717-744); returns `right`.
template <int KIND, int BT>
__device__ double bisect_dpr1_dev(const Problem &P, double sig, const InvValue &R, bool sideR, int &steps,
                                  double *red) {
    const int64_t N = P.N;
    const double *__restrict__ D = P.D;
    const double *__restrict__ W = P.w;
    double left, right;
    double ruu = __dmul_rn(R.rho, R.uu);
    if (R.rho > 0.0) {
        if (!sideR) { left = R.mn1; right = R.mn2; }
        else { left = R.mx1; right = __dadd_rn(R.mx1, ruu); }
    } else {
        if (!sideR) { left = __dadd_rn(R.mn1, ruu); right = R.mn1; }
        else { left = R.mx2; right = R.mx1; }
    }
    double middle = (left + right) / 2.0;
    const double sr = sign_jl(R.rho);
    while ((right - left) > 2.0 * kEps * fmax(fabs(left), fabs(right))) {
        double acc = 0.0;
#pragma unroll 4
        for (int64_t l = threadIdx.x; l < N; l += BT) {
            double w = W[l];
            acc = secular_term_acc(acc, dshift<KIND>(D[l], sig), middle, __dmul_rn(w, w));
        }
        double S = block_allreduce(acc, OpSum(), red);
        if (S != S) {   // NaN from the reciprocal-seed term (bisection point on an inverse pole, overflow): exact divisions
            acc = 0.0;
            for (int64_t l = threadIdx.x; l < N; l += BT) {
                double w = W[l];
                acc = __dadd_rn(acc, secular_term_exact(dshift<KIND>(D[l], sig), middle, __dmul_rn(w, w)));
            }
            S = block_allreduce(acc, OpSum(), red);
        }
        if (KIND != KIND_DPR1) S = __dadd_rn(S, __ddiv_rn(1.0, __dadd_rn(0.0, -middle)));
        double F = __dadd_rn(1.0, __dmul_rn(R.rho, S));
        if (sr * F < 0.0) left = middle; else right = middle;
        middle = (left + right) / 2.0;
        ++steps;
    }
    return right;
}

// ------------------------------------------------------------------------------------------------
template <int KIND, int BT>
__device__ void solve_full(const Problem &P, const Outputs &O, int64_t k, int64_t col, double *red) {
    const int64_t N = P.N;
    const double *__restrict__ D = P.D;
    const double *__restrict__ W = P.w;
    const int tid = threadIdx.x;
    double Kb = 0, Kz = 0, Knu = 0, Krho = 0;
    int64_t Qout = 0;
    int steps = 0, status = 0;

    // ---- shift selection (arrowhead3.jl:431-442, arrowhead4.jl:302-310, arrowhead6.jl:373-384)
    int64_t i;
    bool sideR;
    if (k == 1) {
        i = 1; sideR = true;
    } else if (KIND != KIND_DPR1 && k == N + 1) {
        i = N; sideR = false;
    } else {
        const double Dk = D[k - 1], Dkm = D[k - 2];
        const double mid = dshift<KIND>(Dkm, Dk) / 2.0;
        double S = 0.0;
        for (int64_t l = tid; l < N; l += BT) {
            double w = W[l];
            S = __dadd_rn(S, __ddiv_rn(__dmul_rn(w, w), __dadd_rn(dshift<KIND>(D[l], Dk), -mid)));
        }
        S = block_allreduce(S, OpSum(), red);
        bool right;
        if (KIND == KIND_ARROW) {
            double F = __dadd_rn(__dadd_rn(__dadd_rn(P.a, -Dk), -mid), -S);
            right = F < 0.0;
        } else if (KIND == KIND_HALF) {
            double atemp = __dadd_rn(P.zz, -__dmul_rn(Dk, Dk));
            double F = __dadd_rn(__dadd_rn(atemp, -mid), -S);
            right = F < 0.0;
        } else {
            double F = __dadd_rn(1.0, __dmul_rn(P.a, S));
            right = F > 0.0;
        }
        if (right) { i = k; sideR = true; } else { i = k - 1; sideR = false; }
    }
    const int64_t ii = i - 1;
    double sigma = D[ii];
    const double wi = W[ii];
    const double wz = __ddiv_rn(1.0, wi);

    // ---- inverse of the matrix shifted by pole i (inv!), reduced to its scalars
    double Ps = 0.0, Qs = 0.0, sabs = 0.0, nu1m = 0.0;
    double mxp = -INFINITY, mnm = INFINITY, mxD = -INFINITY, mnD = INFINITY;
    for (int64_t l = tid; l < N; l += BT) {
        if (l == ii) continue;
        double w = W[l];
        double Dp = __ddiv_rn(1.0, dshift<KIND>(D[l], sigma));
        double zp = __dmul_rn(__dmul_rn(-w, Dp), wz);
        double t = __dmul_rn(__dmul_rn(w, w), Dp);
        if (Dp > 0.0) Ps = __dadd_rn(Ps, t); else Qs = __dadd_rn(Qs, t);
        double az = fabs(zp);
        sabs = __dadd_rn(sabs, az);
        mxp = fmax(mxp, __dadd_rn(Dp, az));
        mnm = fmin(mnm, __dadd_rn(Dp, -az));
        mxD = fmax(mxD, Dp);
        mnD = fmin(mnD, Dp);
        nu1m = fmax(nu1m, __dadd_rn(fabs(Dp), az));
    }
    Ps = block_allreduce(Ps, OpSum(), red);
    Qs = block_allreduce(Qs, OpSum(), red);
    sabs = block_allreduce(sabs, OpSum(), red);
    mxp = block_allreduce(mxp, OpMax(), red);
    mnm = block_allreduce(mnm, OpMin(), red);
    mxD = block_allreduce(mxD, OpMax(), red);
    mnD = block_allreduce(mnD, OpMin(), red);
    nu1m = block_allreduce(nu1m, OpMax(), red);
    if (KIND != KIND_DPR1) {  // the slot (0, wz)
        double az = fabs(wz);
        sabs = __dadd_rn(sabs, az);
        mxp = fmax(mxp, az);
        mnm = fmin(mnm, -az);
        mxD = fmax(mxD, 0.0);
        mnD = fmin(mnD, 0.0);
        nu1m = fmax(nu1m, az);
    }
    dd ad = dd_from(0.0);
    if (KIND == KIND_ARROW) {
        double as_ = __dadd_rn(P.a, -sigma);
        ad = dd_sub(dd_from(P.a), dd_from(sigma));
        if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
    } else if (KIND == KIND_HALF) {
        ad = dd_add(dd{P.zz_hi, P.zz_lo}, dd_mul(dd_from(sigma), dd_from(-sigma)));
        double as_ = __dadd_rn(ad.hi, ad.lo);
        if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
    } else {
        double ir = __ddiv_rn(1.0, P.a);
        if (P.a > 0) Ps = __dadd_rn(Ps, ir); else Qs = __dadd_rn(Qs, ir);
    }
    const double PQ = __dadd_rn(Ps, Qs);
    Kb = __ddiv_rn(__dadd_rn(Ps, -Qs), fabs(PQ));
    if (KIND == KIND_DPR1) Kz = __dmul_rn(__dadd_rn(P.sumabs_w, -fabs(wi)), fabs(wz));
    else Kz = __dmul_rn(P.maxabs_w, fabs(wz));
    double b;
    if (Kb < P.tau[0] || Kz < P.tau[1]) {
        b = __dmul_rn(__dmul_rn(PQ, wz), wz);
    } else if (KIND == KIND_HALF || Kz < 1.0 / (kEps * kEps)) {
        Qout = 1;
        status |= AH_BR_DD_B;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        const dd s1 = dd_from(sigma);
        for (int64_t l = tid; l < N; l += BT) {
            if (l == ii) continue;
            dd t;
            bool pos;
            if (KIND == KIND_HALF) {
                dd Dk = dd_from(D[l]);
                dd Dd = dd_mul(dd_add(Dk, s1), dd_sub(Dk, s1));
                dd zd = dd_mul(dd_from(P.z[l]), Dk);
                t = dd_div(dd_mul(zd, zd), Dd);
                pos = Dd.hi > 0.0;
            } else {
                dd Dd = dd_sub(dd_from(D[l]), s1);
                dd wl = dd_from(W[l]);
                t = dd_div(dd_mul(wl, wl), Dd);
                pos = dd_to_double(Dd) > 0.0;
            }
            if (pos) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        Pd = block_allreduce_dd(Pd, red);
        Qd = block_allreduce_dd(Qd, red);
        dd wzd;
        if (KIND == KIND_ARROW) {
            if (dd_to_double(ad) < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
            wzd = dd_from(wi);
        } else if (KIND == KIND_DPR1) {
            dd ir = dd_div(dd_from(1.0), dd_from(P.a));
            if (P.a > 0) Pd = dd_add(Pd, ir); else Qd = dd_add(Qd, ir);
            wzd = dd_from(wi);
        } else {
            if (ad.hi < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
            wzd = dd_mul(dd_from(P.z[ii]), s1);
        }
        dd bd = dd_div(dd_add(Pd, Qd), dd_mul(wzd, wzd));
        b = __dadd_rn(bd.hi, bd.lo);
    } else {
        Qout = 50;
        status |= 2;
        b = __dmul_rn(__dmul_rn(PQ, wz), wz);
    }

    // ---- bisection for the extreme eigenvalue of the inverse arrow (arrowhead3.jl:667-707)
    double nu;
    {
        double left, right;
        if (!sideR) { left = fmin(mnm, __dadd_rn(b, -sabs)); right = mnD; }
        else { left = mxD; right = fmax(mxp, __dadd_rn(b, sabs)); }
        const double wz2 = __dmul_rn(wz, wz);
        double middle = (left + right) / 2.0;
        while ((right - left) > 2.0 * kEps * fmax(fabs(left), fabs(right))) {
            double acc = 0.0;
#pragma unroll 4
            for (int64_t l = tid; l < N; l += BT) {
                if (l == ii) continue;
                double w = W[l];
                acc = secular_term_acc(acc, dshift<KIND>(D[l], sigma), middle, __dmul_rn(w, w));
            }
            double S = block_allreduce(acc, OpSum(), red);
            if (S != S) {   // NaN from the reciprocal-seed term: redo the sweep with exact divisions (see secular_term_exact)
                acc = 0.0;
                for (int64_t l = tid; l < N; l += BT) {
                    if (l == ii) continue;
                    double w = W[l];
                    acc = __dadd_rn(acc, secular_term_exact(dshift<KIND>(D[l], sigma), middle, __dmul_rn(w, w)));
                }
                S = block_allreduce(acc, OpSum(), red);
            }
            S = __dmul_rn(S, wz2);
            if (KIND != KIND_DPR1) S = __dadd_rn(S, __ddiv_rn(wz2, __dadd_rn(0.0, -middle)));
            double F = __dadd_rn(__dadd_rn(b, -middle), -S);
            if (F > 0.0) left = middle; else right = middle;
            middle = (left + right) / 2.0;
            ++steps;
        }
        nu = right;
    }

    double lambda;
    // deflation case `U[zx, zx[k]] = v`: rows AND the column go through the same map
    const int64_t ucolidx = O.rowmap_u ? O.rowmap_u[k - 1] : col;
    double *Ucol = O.U ? O.U + ucolidx * O.ldu : nullptr;
    double *Vcol = (KIND == KIND_HALF && O.V) ? O.V + ucolidx * O.ldv : nullptr;
    double *vec = (KIND == KIND_HALF) ? Vcol : Ucol;                 // the vector of the arrow problem
    const int64_t *vmap = (KIND == KIND_HALF) ? O.rowmap_v : O.rowmap_u;
    const int64_t nu_rows = (KIND == KIND_HALF) ? (P.has_tip ? N + 1 : N) : 0;

    if (isinf(nu)) {
        status |= 1;
        if (KIND == KIND_DPR1) status |= 8;
        lambda = sigma;
        const int64_t nv = (KIND == KIND_DPR1) ? N : N + 1;
        if (vec)
            for (int64_t l = tid; l < nv; l += BT) vec[vmap ? vmap[l] : l] = (l == ii) ? 1.0 : 0.0;
        if (KIND == KIND_HALF && Ucol)
            for (int64_t l = tid; l < nu_rows; l += BT) Ucol[O.rowmap_u ? O.rowmap_u[l] : l] = (l == ii) ? 1.0 : 0.0;
    } else {
        double nu1 = fmax(__dadd_rn(sabs, fabs(b)), nu1m);
        Knu = __ddiv_rn(nu1, fabs(nu));
        double sig_v = sigma, mu, lam2 = 0.0;
        if (KIND != KIND_HALF) {
            int iter = 0;
            while (Knu > P.tau[2]) {
                status |= (status & AH_BR_REM3) ? AH_BR_REM3_REPEAT : AH_BR_REM3;
                nu = sideR ? fabs(nu) : -fabs(nu);
                nu1 = -sign_jl(nu) * nu1;
                double sigma1 = __dadd_rn(__ddiv_rn(__dadd_rn(nu1, nu), __dmul_rn(__dmul_rn(2.0, nu), nu1)), sigma);
                if (is_pole(D, N, sigma1)) { status |= 4; break; }
                InvValue R;
                inv_at_value<KIND, BT>(P, sigma1, P.tau[3], R, status, red);
                Krho = R.Krho;
                nu = bisect_dpr1_dev<KIND, BT>(P, sigma1, R, sideR, steps, red);
                if (KIND == KIND_ARROW && sideR && D[0] > 0.0 && nu < 0.0) {
                    status |= AH_BR_RL_RETRY;
                    nu = bisect_dpr1_dev<KIND, BT>(P, sigma1, R, false, steps, red);
                }
                nu1 = __dadd_rn(R.maxabsD, __dmul_rn(fabs(R.rho), R.uu));
                Knu = __ddiv_rn(nu1, fabs(nu));
                sigma = sigma1;
                Qout += 2 * R.Qout;
                if (isnan(Knu)) break;
                if (++iter >= 32) { status |= 16; break; }
            }
            mu = __ddiv_rn(1.0, nu);
            sig_v = sigma;
            lambda = __dadd_rn(mu, sigma);
        } else {
            if (Knu < P.tau[2]) {
                mu = __ddiv_rn(1.0, nu);
                lam2 = __dadd_rn(mu, __dmul_rn(sigma, sigma));
            } else {
                status |= AH_BR_REM3;
                nu = sideR ? fabs(nu) : -fabs(nu);
                nu1 = -sign_jl(nu) * nu1;
                double sigma1 = __dadd_rn(__ddiv_rn(__dadd_rn(nu1, nu), __dmul_rn(__dmul_rn(2.0, nu), nu1)),
                                          __dmul_rn(sigma, sigma));
                double ss1 = sqrt(sigma1);
                if (is_pole(D, N, ss1)) {
                    status |= 4;
                    mu = __ddiv_rn(1.0, nu);
                    lam2 = __dadd_rn(mu, __dmul_rn(sigma, sigma));
                } else {
                    InvValue R;
                    inv_at_value<KIND, BT>(P, ss1, P.tau[3], R, status, red);
                    Krho = R.Krho;
                    double nub = bisect_dpr1_dev<KIND, BT>(P, ss1, R, sideR, steps, red);
                    mu = __ddiv_rn(1.0, nub);
                    sig_v = ss1;
                    lam2 = __dadd_rn(mu, sigma1);
                }
                if (Qout == 1) Qout += 2;
            }
            lambda = sqrt(lam2);
        }

        // ---- eigenvector: v_l = w_l / (d_l(sigma) - mu), tip -1, scaled by 1/||v||
        double ss = 0.0;
        for (int64_t l = tid; l < N; l += BT) {
            double v = __ddiv_rn(W[l], __dadd_rn(dshift<KIND>(D[l], sig_v), -mu));
            ss = __dadd_rn(ss, __dmul_rn(v, v));
        }
        ss = block_allreduce(ss, OpSum(), red);
        if (KIND != KIND_DPR1) ss = __dadd_rn(ss, 1.0);
        const double inrm = __ddiv_rn(1.0, sqrt(ss));
        if (vec) {
            for (int64_t l = tid; l < N; l += BT) {
                double v = __dmul_rn(__ddiv_rn(W[l], __dadd_rn(dshift<KIND>(D[l], sig_v), -mu)), inrm);
                double vo = v;
                if (KIND == KIND_HALF && O.rowscale_v) vo = __dmul_rn(v, O.rowscale_v[l]);
                vec[vmap ? vmap[l] : l] = vo;
                if (KIND == KIND_HALF && Ucol)
                    Ucol[O.rowmap_u ? O.rowmap_u[l] : l] = __ddiv_rn(__dmul_rn(lambda, v), D[l]);
            }
            if (KIND != KIND_DPR1 && tid == 0) {
                double vt = __dmul_rn(-1.0, inrm);
                vec[vmap ? vmap[N] : N] = vt;
                if (KIND == KIND_HALF && Ucol && P.has_tip)
                    Ucol[O.rowmap_u ? O.rowmap_u[N] : N] = __ddiv_rn(__dmul_rn(P.ztip, vt), lambda);
            }
        } else if (KIND == KIND_HALF && Ucol) {
            for (int64_t l = tid; l < N; l += BT) {
                double v = __dmul_rn(__ddiv_rn(W[l], __dadd_rn(dshift<KIND>(D[l], sig_v), -mu)), inrm);
                Ucol[O.rowmap_u ? O.rowmap_u[l] : l] = __ddiv_rn(__dmul_rn(lambda, v), D[l]);
            }
            if (tid == 0 && P.has_tip)
                Ucol[O.rowmap_u ? O.rowmap_u[N] : N] = __ddiv_rn(__dmul_rn(P.ztip, __dmul_rn(-1.0, inrm)), lambda);
        }

        // ---- Remedy 1 (arrowhead3.jl:526-544, arrowhead4.jl:400-417, arrowhead6.jl:489-503)
        bool rem1;
        if (KIND == KIND_HALF) {
            double Di = fabs(D[ii]);
            rem1 = (k == N + 1) &&
                   (__ddiv_rn(__dadd_rn(__dmul_rn(Di, Di), fabs(__ddiv_rn(1.0, nu))), fabs(__dmul_rn(lambda, lambda))) > P.tau[4]);
        } else {
            rem1 = __ddiv_rn(__dadd_rn(fabs(D[ii]), fabs(__ddiv_rn(1.0, nu))), fabs(lambda)) > P.tau[4];
            if (rem1) {
                bool c1 = (k == 1 && D[0] < 0.0);
                bool c2 = (KIND == KIND_ARROW) && (k == N + 1 && D[N - 1] > 0.0);
                bool c3, c4;
                if (KIND == KIND_ARROW) c3 = (i < N && !sideR && sign_jl(D[ii]) + sign_jl(D[ii + 1]) == 0.0);
                else c3 = (!sideR && sign_jl(D[ii]) + sign_jl(D[ii + 1]) == 0.0);
                c4 = (i > 1 && sideR && sign_jl(D[ii]) + sign_jl(D[ii - 1]) == 0.0);
                rem1 = c1 || c2 || c3 || c4;
            }
        }
        if (rem1) {
            InvValue R;
            inv_at_value<KIND, BT>(P, 0.0, P.tau[3], R, status, red);
            status |= AH_BR_REM1;
            Krho = R.Krho;
            if (KIND == KIND_HALF) { if (Qout == 1) Qout += 4; }
            else Qout += 4 * R.Qout;
            if (isinf(R.rho)) {
                status |= AH_BR_REM1_ZERO;
                lambda = 0.0;
            } else {
                double s1 = 0.0, s2 = 0.0;
                for (int64_t l = tid; l < N; l += BT) {
                    double w = W[l];
                    double v = __dmul_rn(__ddiv_rn(w, __dadd_rn(dshift<KIND>(D[l], sig_v), -mu)), inrm);
                    double D1 = __ddiv_rn(1.0, dshift<KIND>(D[l], 0.0));
                    double u1 = __dmul_rn(w, D1);
                    s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(v, v), D1));
                    s2 = __dadd_rn(s2, __dmul_rn(v, u1));
                }
                s1 = block_allreduce(s1, OpSum(), red);
                s2 = block_allreduce(s2, OpSum(), red);
                if (KIND != KIND_DPR1) s2 = __dadd_rn(s2, __dmul_rn(__dmul_rn(-1.0, inrm), -1.0));
                double nur = __dadd_rn(s1, __dmul_rn(R.rho, __dmul_rn(s2, s2)));
                lambda = (KIND == KIND_HALF) ? sqrt(__ddiv_rn(1.0, nur)) : __ddiv_rn(1.0, nur);
            }
        }
    }
    if (!(status & 1) && !interlace_ok(D, N, k, lambda)) status |= 32;   // AH_ST_INTERLACE
    if (tid == 0) {
        O.lambda[col] = lambda;
        O.sind[col] = i;
        O.kb[col] = Kb; O.kz[col] = Kz; O.knu[col] = Knu; O.krho[col] = Krho;
        O.qout[col] = Qout;
        O.steps[col] = steps;
        O.status[col] = status;
    }
}

// klist == nullptr: eigenpairs k0 .. k0+nwork-1.  Otherwise the work list written by the fast kernel,
// whose length is read from *nwork_dev (no host round trip between the two launches).
// first_dev (optional): the list entries below *first_dev were solved by an earlier launch (the eigenpairs k_prep hands over
// are started beside k_bisect, those handed over later follow after the Remedy-3 round).
template <int KIND, int BT>
__global__ void __launch_bounds__(BT) k_full(Problem P, Outputs O, const int64_t *__restrict__ klist,
                                             int64_t k0, int64_t nwork, const unsigned long long *nwork_dev,
                                             const unsigned long long *first_dev = nullptr) {
    __shared__ double red[2 * (BT / 32) + 2];
    if (klist) nwork = (int64_t)*nwork_dev;
    const int64_t first = first_dev ? (int64_t)*first_dev : 0;
    for (int64_t wi = first + blockIdx.x; wi < nwork; wi += gridDim.x) {
        const int64_t k = klist ? klist[wi] : k0 + wi;
        solve_full<KIND, BT>(P, O, k, k - k0, red);
    }
}

}  // namespace ah
