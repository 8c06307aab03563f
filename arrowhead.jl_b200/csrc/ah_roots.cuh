// ah_roots.cuh -- eigenvalue-only solver with a double-double border: the per-k solver of rootsah
// (src/arrowhead7.jl:351-495, `eigen(A,zD,alphaD,k,tau)`), one CTA per eigenvalue.
//
// rootsah turns a polynomial with real distinct roots into an arrowhead companion matrix in barycentric form whose
// border z and arrow top alpha are known in double-double (zD, alphaD; arrowhead7.jl:53-115), then calls this solver
// for k = 1..n in a SERIAL loop (:116-118).  Here the loop is the grid.  Differences to eigen(A,Ainv,k,tau)
// (ah_full.cuh), all kept: the shift selection sums in double-double (:381-389); the double-double recomputation of b
// and rho uses (zD, alphaD) (:163-188, :270-287) and there is a `Kb == Inf -> b = 0` branch but no BigFloat branch;
// Remedy 3 runs once, not in a loop, and leaves nu untouched for the Remedy-1 test (:415-459, :464); Remedy 1 inverts
// the PLAIN matrix (`inv(A,0.0,tau[4])`, :470); `Qout==1 && (Qout+=2|4)` bookkeeping; only lambda and Qout come back.
#pragma once
#include "ah_device.cuh"
#include "ah_full.cuh"

namespace ah {

template <int BT>
__device__ void solve_roots(const Problem &P, const double *__restrict__ Wlo, double a_lo, const Outputs &O, int64_t k,
                            int64_t col, double *red) {
    constexpr int KIND = KIND_ARROW;
    const int64_t N = P.N;
    const double *__restrict__ D = P.D;
    const double *__restrict__ W = P.w;
    const int tid = threadIdx.x;
    double Kb = 0, Kz = 0, Knu = 0, Krho = 0;
    int64_t Qout = 0;
    int steps = 0, status = 0;
    const dd aD{P.a, a_lo};

    // ---- shift selection in double-double (arrowhead7.jl:374-390)
    int64_t i;
    bool sideR;
    if (k == 1) { i = 1; sideR = true; }
    else if (k == N + 1) { i = N; sideR = false; }
    else {
        const dd Dk = dd_from(D[k - 1]);
        const dd mid = dd_div(dd_sub(dd_from(D[k - 2]), Dk), dd_from(2.0));
        dd S = dd_from(0.0);
        for (int64_t l = tid; l < N; l += BT) {
            const dd zl{W[l], Wlo[l]};
            S = dd_add(S, dd_div(dd_mul(zl, zl), dd_sub(dd_sub(dd_from(D[l]), Dk), mid)));
        }
        S = block_allreduce_dd(S, red);
        const dd F = dd_sub(dd_sub(dd_sub(aD, Dk), mid), S);
        if (__dadd_rn(F.hi, F.lo) < 0.0) { i = k; sideR = true; } else { i = k - 1; sideR = false; }
    }
    const int64_t ii = i - 1;
    const double sigma = D[ii];
    const double wi = W[ii];
    const double wz = __ddiv_rn(1.0, wi);

    // ---- inverse shifted by pole i, reduced to its scalars (arrowhead7.jl:132-225)
    double Ps = 0.0, Qs = 0.0, sabs = 0.0, nu1m = 0.0, mz = 0.0;
    double mxp = -INFINITY, mnm = INFINITY, mxD = -INFINITY, mnD = INFINITY;
    for (int64_t l = tid; l < N; l += BT) {
        const double w = W[l];
        mz = fmax(mz, fabs(w));
        if (l == ii) continue;
        const double Dp = __ddiv_rn(1.0, __dadd_rn(D[l], -sigma));
        const double zp = __dmul_rn(__dmul_rn(-w, Dp), wz);
        const double t = __dmul_rn(__dmul_rn(w, w), Dp);
        if (Dp > 0.0) Ps = __dadd_rn(Ps, t); else Qs = __dadd_rn(Qs, t);
        const double az = fabs(zp);
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
    mz = block_allreduce(mz, OpMax(), red);
    mxp = block_allreduce(mxp, OpMax(), red);
    mnm = block_allreduce(mnm, OpMin(), red);
    mxD = block_allreduce(mxD, OpMax(), red);
    mnD = block_allreduce(mnD, OpMin(), red);
    nu1m = block_allreduce(nu1m, OpMax(), red);
    {   // the slot (0, wz)
        const double az = fabs(wz);
        sabs = __dadd_rn(sabs, az);
        mxp = fmax(mxp, az); mnm = fmin(mnm, -az);
        mxD = fmax(mxD, 0.0); mnD = fmin(mnD, 0.0);
        nu1m = fmax(nu1m, az);
    }
    {
        const double as_ = __dadd_rn(P.a, -sigma);
        if (as_ < 0) Ps = __dadd_rn(Ps, -as_); else Qs = __dadd_rn(Qs, -as_);
    }
    const double PQ = __dadd_rn(Ps, Qs);
    Kb = __ddiv_rn(__dadd_rn(Ps, -Qs), fabs(PQ));
    Kz = __dmul_rn(mz, fabs(wz));
    double b;
    if (Kb < P.tau[0] || Kz < P.tau[1]) {
        b = __dmul_rn(__dmul_rn(PQ, wz), wz);
    } else if (isinf(Kb)) {
        b = 0.0;
    } else {
        Qout = 1;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        const dd s1 = dd_from(sigma);
        for (int64_t l = tid; l < N; l += BT) {
            if (l == ii) continue;
            const dd Dd = dd_sub(dd_from(D[l]), s1);
            const dd zl{W[l], Wlo[l]};
            const dd t = dd_div(dd_mul(zl, zl), Dd);
            if (Dd.hi > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        Pd = block_allreduce_dd(Pd, red);
        Qd = block_allreduce_dd(Qd, red);
        const dd ad = dd_sub(aD, s1);
        if (ad.hi < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        const dd wzd{wi, Wlo[ii]};
        const dd bd = dd_div(dd_add(Pd, Qd), dd_mul(wzd, wzd));
        b = __dadd_rn(bd.hi, bd.lo);
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
                const double w = W[l];
                acc = secular_term_acc(acc, __dadd_rn(D[l], -sigma), middle, __dmul_rn(w, w));
            }
            double S = __dmul_rn(block_allreduce(acc, OpSum(), red), wz2);
            S = __dadd_rn(S, __ddiv_rn(wz2, __dadd_rn(0.0, -middle)));
            const double F = __dadd_rn(__dadd_rn(b, -middle), -S);
            if (F > 0.0) left = middle; else right = middle;
            middle = (left + right) / 2.0;
            ++steps;
        }
        nu = right;
    }

    double lambda;
    if (isinf(nu)) {
        status |= 1;
        lambda = sigma;
    } else {
        double nu1 = fmax(__dadd_rn(sabs, fabs(b)), nu1m);
        Knu = __ddiv_rn(nu1, fabs(nu));
        double sig_v = sigma, mu;
        if (Knu < P.tau[2]) {
            mu = __ddiv_rn(1.0, nu);
            lambda = __dadd_rn(mu, sigma);
        } else {   // Remedy 3, once; nu keeps its (sign-adjusted) first-stage value
            nu = sideR ? fabs(nu) : -fabs(nu);
            nu1 = -sign_jl(nu) * nu1;
            const double sigma1 = __dadd_rn(__ddiv_rn(__dadd_rn(nu1, nu), __dmul_rn(__dmul_rn(2.0, nu), nu1)), sigma);
            if (is_pole(D, N, sigma1)) {
                status |= 4;
                mu = __ddiv_rn(1.0, nu);
                lambda = __dadd_rn(mu, sigma);
            } else {
                // inv(A,zD,alphaD,sigma1,tau[4]) (arrowhead7.jl:232-292): rho, K_rho and the DPR1 bisection bracket
                double P1 = 0.0, Q1 = 0.0, uu = 0.0, mabs = 0.0;
                double mx1 = -INFINITY, mx2 = -INFINITY, mn1 = INFINITY, mn2 = INFINITY;
                for (int64_t l = tid; l < N; l += BT) {
                    const double w = W[l];
                    const double D1 = __ddiv_rn(1.0, __dadd_rn(D[l], -sigma1));
                    const double u1 = __dmul_rn(w, D1);
                    const double t = __dmul_rn(__dmul_rn(w, w), D1);
                    if (D1 > 0.0) P1 = __dadd_rn(P1, t); else Q1 = __dadd_rn(Q1, t);
                    uu = __dadd_rn(uu, __dmul_rn(u1, u1));
                    mabs = fmax(mabs, fabs(D1));
                    top2_push<true>(mx1, mx2, D1);
                    top2_push<false>(mn1, mn2, D1);
                }
                P1 = block_allreduce(P1, OpSum(), red);
                Q1 = block_allreduce(Q1, OpSum(), red);
                uu = block_allreduce(uu, OpSum(), red);
                mabs = block_allreduce(mabs, OpMax(), red);
                block_allreduce_top2<true>(mx1, mx2, red);
                block_allreduce_top2<false>(mn1, mn2, red);
                uu = __dadd_rn(uu, 1.0);
                top2_push<true>(mx1, mx2, 0.0);
                top2_push<false>(mn1, mn2, 0.0);
                const double as_ = __dadd_rn(P.a, -sigma1);
                if (as_ < 0) P1 = __dadd_rn(P1, -as_); else Q1 = __dadd_rn(Q1, -as_);
                const double PQ1 = __dadd_rn(P1, Q1);
                Krho = __ddiv_rn(__dadd_rn(P1, -Q1), fabs(PQ1));
                InvValue R;
                if (Krho < P.tau[3]) {
                    R.rho = __ddiv_rn(-1.0, PQ1);
                } else {
                    dd Pd = dd_from(0.0), Qd = dd_from(0.0);
                    const dd s1 = dd_from(sigma1);
                    for (int64_t l = tid; l < N; l += BT) {
                        const dd zl{W[l], Wlo[l]};
                        const dd t = dd_div(dd_mul(zl, zl), dd_sub(dd_from(D[l]), s1));
                        if (__ddiv_rn(1.0, __dadd_rn(D[l], -sigma1)) > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
                    }
                    Pd = block_allreduce_dd(Pd, red);
                    Qd = block_allreduce_dd(Qd, red);
                    const dd ad = dd_sub(aD, s1);
                    if (__dadd_rn(ad.hi, ad.lo) < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
                    const dd r = dd_div(dd_from(1.0), dd_add(Pd, Qd));
                    R.rho = -__dadd_rn(r.hi, r.lo);
                }
                R.Krho = Krho; R.Qout = 0; R.uu = uu; R.maxabsD = mabs;
                R.mx1 = mx1; R.mx2 = mx2; R.mn1 = mn1; R.mn2 = mn2;
                const double nub = bisect_dpr1_dev<KIND, BT>(P, sigma1, R, sideR, steps, red);
                mu = __ddiv_rn(1.0, nub);
                sig_v = sigma1;
                lambda = __dadd_rn(mu, sigma1);
            }
            if (Qout == 1) Qout += 2;
        }

        // ---- Remedy 1 (arrowhead7.jl:464-487)
        if (__ddiv_rn(__dadd_rn(fabs(D[ii]), fabs(__ddiv_rn(1.0, nu))), fabs(lambda)) > P.tau[4]) {
            const bool c1 = (k == 1 && D[0] < 0.0);
            const bool c2 = (k == N + 1 && D[N - 1] > 0.0);
            const bool c3 = (i < N && !sideR && sign_jl(D[ii]) + sign_jl(D[ii + 1]) == 0.0);
            bool c4 = false;
            if (!(c1 || c2 || c3) && sideR) {
                if (i == 1) status |= 8;   // A.D[0]: BoundsError in the reference
                else c4 = (sign_jl(D[ii]) + sign_jl(D[ii - 1]) == 0.0);
            }
            if (c1 || c2 || c3 || c4) {
                InvValue R;
                inv_at_value<KIND, BT>(P, 0.0, P.tau[3], R, status, red);   // the plain matrix: inv(A,0.0,tau[4])
                Krho = R.Krho;
                if (Qout == 1) Qout += 4;
                if (isinf(R.rho)) {
                    lambda = 0.0;
                } else {
                    double ss = 0.0;
                    for (int64_t l = tid; l < N; l += BT) {
                        const double v = __ddiv_rn(W[l], __dadd_rn(__dadd_rn(D[l], -sig_v), -mu));
                        ss = __dadd_rn(ss, __dmul_rn(v, v));
                    }
                    ss = __dadd_rn(block_allreduce(ss, OpSum(), red), 1.0);
                    const double inrm = __ddiv_rn(1.0, sqrt(ss));
                    double s1 = 0.0, s2 = 0.0;
                    for (int64_t l = tid; l < N; l += BT) {
                        const double w = W[l];
                        const double v = __dmul_rn(__ddiv_rn(w, __dadd_rn(__dadd_rn(D[l], -sig_v), -mu)), inrm);
                        const double D1 = __ddiv_rn(1.0, D[l]);
                        s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(v, v), D1));
                        s2 = __dadd_rn(s2, __dmul_rn(v, __dmul_rn(w, D1)));
                    }
                    s1 = block_allreduce(s1, OpSum(), red);
                    s2 = block_allreduce(s2, OpSum(), red);
                    s2 = __dadd_rn(s2, __dmul_rn(__dmul_rn(-1.0, inrm), -1.0));
                    lambda = __ddiv_rn(1.0, __dadd_rn(s1, __dmul_rn(R.rho, __dmul_rn(s2, s2))));
                }
            }
        }
    }
    if (!(status & 1) && !interlace_ok(D, N, k, lambda)) status |= 32;
    if (tid == 0) {
        O.lambda[col] = lambda;
        O.sind[col] = i;
        O.kb[col] = Kb; O.kz[col] = Kz; O.knu[col] = Knu; O.krho[col] = Krho;
        O.qout[col] = Qout;
        O.steps[col] = steps;
        O.status[col] = status & AH_ST_MASK;   // rootsah keeps no branch trace (its per-k solver differs, arrowhead7.jl:351-495)
    }
}

template <int BT>
__global__ void __launch_bounds__(BT) k_roots(Problem P, const double *__restrict__ Wlo, double a_lo, Outputs O, int64_t k0,
                                              int64_t nwork) {
    __shared__ double red[2 * (BT / 32) + 2];
    for (int64_t wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        solve_roots<BT>(P, Wlo, a_lo, O, k0 + wi, wi, red);
        __syncthreads();
    }
}

}  // namespace ah
