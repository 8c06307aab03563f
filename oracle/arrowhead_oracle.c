/*
 * oracle/arrowhead_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE binary64, no FMA contraction: build with
 * -ffp-contract=off, no -ffast-math) of the per-eigenpair solvers of
 * ivanslapnicar/Arrowhead.jl.  The reference is pure Julia and there is no Julia
 * runtime in this image, so the reference cannot be compiled or executed here; this
 * file follows the reference's loops statement by statement instead (same evaluation
 * order, same branch tests, same returned bracket end) and is pinned against the
 * reference's own known-answer vectors (test/runtests.jl:22-42, 58-78) by
 * tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (arrowhead.jl_b200/) never does.
 *
 * Reference map (file:line relative to /root/reference/src):
 *   dd_*                      DoubleDouble.jl:34-42 (split), 93-96 (double), 112-116 (add2),
 *                             129-133 (sub), 137-149 (mul12, mul2), 152-157 (div2), 160-168 (sqrt2)
 *   jl_pairwise_sum           Julia Base.mapreduce_impl (pairwise, block 1024) used by `sum`
 *   bisect_arrow              arrowhead3.jl:667-707
 *   bisect_dpr1               arrowhead3.jl:717-744
 *   arrow_inv_shift_index     arrowhead3.jl:152-244  (inv!(B,A::SymArrow,i,tau))
 *   arrow_inv_shift_value     arrowhead3.jl:260-332  (inv(A::SymArrow,sigma::Float64,tau))
 *   ah_oracle_symarrow_k      arrowhead3.jl:419-550  (eigen(A,Ainv,k,tau))
 *   dpr1_inv_shift_index      arrowhead4.jl:89-159
 *   dpr1_inv_shift_value      arrowhead4.jl:173-225
 *   ah_oracle_symdpr1_k       arrowhead4.jl:289-422
 *   half_inv_shift_index      arrowhead6.jl:144-213, doubledot arrowhead6.jl:7-13
 *   half_inv_shift_value      arrowhead6.jl:227-275
 *   ah_oracle_halfarrow_k     arrowhead6.jl:344-508
 *   *_range                   the `Threads.@threads for k` loops, arrowhead3.jl:639-644,
 *                             arrowhead4.jl:529-535, arrowhead6.jl:635-642 (OpenMP static over k,
 *                             one scratch arrow per thread like Bx[threadid()])
 *
 * Branches of the reference that need 256-bit BigFloat (Qout 50/100) or that throw in the
 * reference (pole-return sub-paths, arrowhead3.jl:481, arrowhead4.jl:324,350, arrowhead6.jl:440)
 * are not computed; the eigenpair gets a status bit instead (same bits as include/arrowhead_b200.h).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define AH_EPS 2.220446049250313e-16

/* Accumulator type of every O(N) sum.  Default build: double, i.e. exactly the reference's rounding
 * sequence.  -DAH_ORACLE_HI: x87 long double (64-bit significand) -- the same algorithm with (nearly)
 * exactly rounded sums, used by the tests to tell the reference's own summation noise (which the
 * algorithm amplifies by up to Knu <= tolnu = 1e3) apart from an error of the implementation under test. */
#ifdef AH_ORACLE_HI
typedef long double acc_t;
#else
typedef double acc_t;
#endif

#define AH_ST_OK 0
#define AH_ST_NU_INF 1          /* |nu|==Inf deflation branch taken (arrowhead3.jl:451-455) */
#define AH_ST_NEEDS_BIGFLOAT 2  /* reference would switch to BigFloat (Qout 50 / 100) */
#define AH_ST_POLE_RETURN 4     /* "came back with a pole" sub-path (throws in the reference) */
#define AH_ST_REF_THROWS 8      /* reference raises (e.g. arrowhead4.jl:324 undefined `sigma`) */
/* branch trace (bits 8..14 of the status word; same meaning as AH_BR_* in include/arrowhead_b200.h): which
 * non-standard branches of the reference's per-k solver this eigenpair went through */
#define AH_BR_DD_B 0x100        /* b recomputed in double-double (inv!, Kb/Kz test failed) */
#define AH_BR_REM3 0x200        /* Remedy 3: re-shift between the poles (K_nu > tolnu) */
#define AH_BR_REM3_REPEAT 0x400 /* ... more than once (the while loop of arrowhead3.jl:466 went round again) */
#define AH_BR_RL_RETRY 0x800    /* 'R' -> 'L' retry of the DPR1 bisection (arrowhead3.jl:499-501) */
#define AH_BR_DD_RHO 0x1000     /* rho of an inverse at a non-pole shift in double-double (K_rho >= tolrho) */
#define AH_BR_REM1 0x2000       /* Remedy 1: Rayleigh quotient on the inverse of the unshifted matrix */
#define AH_BR_REM1_ZERO 0x4000  /* ... whose rho is infinite: lambda = 0 */

/* ------------------------------------------------------------------ Julia scalar semantics */
static inline double jl_max(double a, double b) {
    if (isnan(a)) return a;
    if (isnan(b)) return b;
    if (a > b) return a;
    if (b > a) return b;
    return signbit(a) ? b : a;
}
static inline double jl_min(double a, double b) {
    if (isnan(a)) return a;
    if (isnan(b)) return b;
    if (a < b) return a;
    if (b < a) return b;
    return signbit(a) ? a : b;
}
static inline double jl_sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : x); }
static inline int jl_isequal(double a, double b) {
    if (isnan(a) && isnan(b)) return 1;
    uint64_t ua, ub;
    memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
    return ua == ub;
}

/* Julia `sum` over an array: pairwise above 1024 elements, sequential inside a block. */
static double jl_pairwise_sum(const double *x, int64_t lo, int64_t hi /* inclusive */) {
#ifdef AH_ORACLE_HI
    acc_t v = 0.0;
    for (int64_t i = lo; i <= hi; ++i) v += x[i];
    return (double)v;
#endif
    if (hi < lo) return 0.0;
    if (lo == hi) return x[lo];
    if (hi - lo < 1024) {
        double v = x[lo] + x[lo + 1];
        for (int64_t i = lo + 2; i <= hi; ++i) v += x[i];
        return v;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    double v1 = jl_pairwise_sum(x, lo, mid);
    double v2 = jl_pairwise_sum(x, mid + 1, hi);
    return v1 + v2;
}

/* ------------------------------------------------------------------ Dekker double-double */
typedef struct { double hi, lo; } dd;

static inline dd dd_from(double x) { dd r = {x, 0.0}; return r; }
static inline double dd_to_double(dd x) { return x.hi + x.lo; }
static inline dd dd_renorm(double u, double v) { double w = u + v; dd r = {w, (u - w) + v}; return r; }
static inline dd dd_neg(dd x) { dd r = {-x.hi, -x.lo}; return r; }
static inline dd dd_add(dd x, dd y) {
    double r = x.hi + y.hi;
    double s = fabs(x.hi) > fabs(y.hi) ? (((x.hi - r) + y.hi) + y.lo) + x.lo
                                       : (((y.hi - r) + x.hi) + x.lo) + y.lo;
    return dd_renorm(r, s);
}
static inline dd dd_sub(dd x, dd y) {
    double r = x.hi - y.hi;
    double s = fabs(x.hi) > fabs(y.hi) ? (((x.hi - r) - y.hi) - y.lo) + x.lo
                                       : (((-y.hi - r) + x.hi) + x.lo) - y.lo;
    return dd_renorm(r, s);
}
static inline void dd_split(double x, double *h, double *l) {
    double p = x * 1.34217729e8;
    *h = (x - p) + p;
    *l = x - *h;
}
static inline dd dd_mul12(double x, double y) {
    double hx, lx, hy, ly;
    dd_split(x, &hx, &lx);
    dd_split(y, &hy, &ly);
    double z = x * y;
    dd r = {z, (((hx * hy - z) + hx * ly) + lx * hy) + lx * ly};
    return r;
}
static inline dd dd_mul(dd x, dd y) {
    dd c = dd_mul12(x.hi, y.hi);
    double cc = (x.hi * y.lo + x.lo * y.hi) + c.lo;
    return dd_renorm(c.hi, cc);
}
static inline dd dd_div(dd x, dd y) {
    double c = x.hi / y.hi;
    dd u = dd_mul12(c, y.hi);
    double cc = ((((x.hi - u.hi) - u.lo) + x.lo) - c * y.lo) / y.hi;
    return dd_renorm(c, cc);
}
static inline dd dd_abs(dd x) { return x.hi > 0 ? x : dd_neg(x); }
static inline int dd_is_pluszero(dd x) { return jl_isequal(x.hi, 0.0) && jl_isequal(x.lo, 0.0); }

/* exported for the DD unit tests */
void ah_oracle_dd_op(int op, const double *x, const double *y, double *out) {
    dd a = {x[0], x[1]}, b = {y[0], y[1]}, r = {0, 0};
    switch (op) {
        case 0: r = dd_add(a, b); break;
        case 1: r = dd_sub(a, b); break;
        case 2: r = dd_mul(a, b); break;
        case 3: r = dd_div(a, b); break;
        default: break;
    }
    out[0] = r.hi; out[1] = r.lo;
}

/* ------------------------------------------------------------------ bisection */
/* Extreme eigenvalue of the arrow [diag(D) z; z' a] (arrow top position irrelevant).
 * zsq is scratch of length n.  Returns `right`, like the reference. */
static double bisect_arrow(const double *D, const double *z, double a, int64_t n, char side,
                           double *zsq, int32_t *steps) {
    double left, right;
    for (int64_t k = 0; k < n; ++k) zsq[k] = fabs(z[k]);
    double sumz = jl_pairwise_sum(zsq, 0, n - 1);
    if (side == 'L') {
        left = D[0] - zsq[0];
        for (int64_t k = 1; k < n; ++k) left = jl_min(left, D[k] - zsq[k]);
        left = jl_min(left, a - sumz);
        right = D[0];
        for (int64_t k = 1; k < n; ++k) right = jl_min(right, D[k]);
    } else {
        left = D[0];
        for (int64_t k = 1; k < n; ++k) left = jl_max(left, D[k]);
        right = D[0] + zsq[0];
        for (int64_t k = 1; k < n; ++k) right = jl_max(right, D[k] + zsq[k]);
        right = jl_max(right, a + sumz);
    }
    double middle = (left + right) / 2.0;
    for (int64_t k = 0; k < n; ++k) zsq[k] = zsq[k] * zsq[k];
    int32_t cnt = 0;
    while ((right - left) > 2.0 * AH_EPS * jl_max(fabs(left), fabs(right))) {
        acc_t F = 0.0;
        for (int64_t k = 0; k < n; ++k) F += zsq[k] / (D[k] - middle);
        F = a - middle - F;
        if (F > 0.0) left = middle; else right = middle;
        middle = (left + right) / 2.0;
        ++cnt;
    }
    if (steps) *steps += cnt;
    return right;
}

/* Extreme eigenvalue of diag(D)+r*u*u' (arrowhead3.jl:717-744).  usq scratch of length n (n>=2). */
static double bisect_dpr1(const double *D, const double *u, double r, int64_t n, char side,
                          double *usq, int32_t *steps) {
    /* sortperm(D,rev=true): only positions 1,2,n-1,n of the sorted order are used */
    double mx1 = -INFINITY, mx2 = -INFINITY, mn1 = INFINITY, mn2 = INFINITY;
    for (int64_t k = 0; k < n; ++k) {
        double d = D[k];
        if (d > mx1) { mx2 = mx1; mx1 = d; } else if (d > mx2) mx2 = d;
        if (d < mn1) { mn2 = mn1; mn1 = d; } else if (d < mn2) mn2 = d;
    }
    acc_t uu = 0.0;
    for (int64_t k = 0; k < n; ++k) uu += u[k] * u[k];
    double left, right;
    if (r > 0.0) {
        if (side == 'L') { left = mn1; right = mn2; }
        else { left = mx1; right = mx1 + r * uu; }
    } else {
        if (side == 'L') { left = mn1 + r * uu; right = mn1; }
        else { left = mx2; right = mx1; }
    }
    double middle = (left + right) / 2.0;
    for (int64_t k = 0; k < n; ++k) usq[k] = u[k] * u[k];
    double sr = jl_sign(r);
    int32_t cnt = 0;
    while ((right - left) > 2.0 * AH_EPS * jl_max(fabs(left), fabs(right))) {
        acc_t F = 0.0;
        for (int64_t k = 0; k < n; ++k) F = F + usq[k] / (D[k] - middle);
        F = 1.0 + r * F;
        if (sr * F < 0.0) left = middle; else right = middle;
        middle = (left + right) / 2.0;
        ++cnt;
    }
    if (steps) *steps += cnt;
    return right;
}

/* ------------------------------------------------------------------ scratch */
typedef struct {
    double *bd, *bz;   /* inverse arrow (shift at a pole)       : n+1 entries */
    double *d1, *u1;   /* inverse DPR1 (shift between poles)     : n+1 entries */
    double *w;         /* squares / abs scratch                  : n+1 entries */
    double *t;         /* term scratch for pairwise sums         : n+1 entries */
    double *z1;        /* HalfArrow z.*D                         : n+1 entries */
} scratch_t;

static int scratch_alloc(scratch_t *s, int64_t n) {
    size_t m = (size_t)(n + 2);
    s->bd = (double *)calloc(m, 8); s->bz = (double *)calloc(m, 8);
    s->d1 = (double *)calloc(m, 8); s->u1 = (double *)calloc(m, 8);
    s->w = (double *)calloc(m, 8);  s->t = (double *)calloc(m, 8);
    s->z1 = (double *)calloc(m, 8);
    return s->bd && s->bz && s->d1 && s->u1 && s->w && s->t && s->z1;
}
static void scratch_free(scratch_t *s) {
    free(s->bd); free(s->bz); free(s->d1); free(s->u1); free(s->w); free(s->t); free(s->z1);
}

/* ------------------------------------------------------------------ SymArrow */
/* inv!(B,A,i,tau): A = arrow with N poles, arrow top last.  i is 1-based.
 * B gets N entries: the N-1 shifted/inverted poles in order, then the slot (0, wz). */
static void arrow_inv_shift_index(int64_t N, const double *D, const double *z, double a, int64_t i,
                                  double tolb, double tolz, double *Bd, double *Bz, double *b_out,
                                  double *Kb_out, double *Kz_out, int64_t *Qout, int32_t *status) {
    const int64_t ii = i - 1;
    double wz = 1.0 / z[ii];
    double sigma = D[ii];
    for (int64_t k = 0; k < ii; ++k) {
        Bd[k] = 1.0 / (D[k] - sigma);
        Bz[k] = -z[k] * Bd[k] * wz;
    }
    for (int64_t k = ii + 1; k < N; ++k) {
        Bd[k - 1] = 1.0 / (D[k] - sigma);
        Bz[k - 1] = -z[k] * Bd[k - 1] * wz;
    }
    double as = a - sigma;
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < ii; ++k) {
        if (Bd[k] > 0.0) P += z[k] * z[k] * Bd[k]; else Q += z[k] * z[k] * Bd[k];
    }
    for (int64_t k = ii; k < N - 1; ++k) {
        if (Bd[k] > 0.0) P += z[k + 1] * z[k + 1] * Bd[k]; else Q += z[k + 1] * z[k + 1] * Bd[k];
    }
    if (as < 0) P -= as; else Q -= as;
    double Kb = (P - Q) / fabs(P + Q);
    double mz = fabs(z[0]);
    for (int64_t k = 1; k < N; ++k) mz = jl_max(mz, fabs(z[k]));
    double Kz = mz * fabs(wz);
    double b;
    if (Kb < tolb || Kz < tolz) {
        b = (P + Q) * wz * wz;
    } else if (Kz < 1.0 / (AH_EPS * AH_EPS)) {
        *Qout = 1;
        dd s1 = dd_from(sigma);
        dd wzd = dd_from(z[ii]);
        dd ad = dd_sub(dd_from(a), s1);
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        for (int64_t k = 0; k < N; ++k) {
            if (k == ii) continue;
            dd Dd = dd_sub(dd_from(D[k]), s1);
            dd zk = dd_from(z[k]);
            dd t = dd_div(dd_mul(zk, zk), Dd);
            if (dd_to_double(Dd) > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        if (dd_to_double(ad) < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        dd bd = dd_div(dd_add(Pd, Qd), dd_mul(wzd, wzd));
        b = dd_to_double(bd);
    } else {
        *Qout = 50;
        *status |= AH_ST_NEEDS_BIGFLOAT;
        b = (P + Q) * wz * wz;
    }
    Bd[N - 1] = 0.0;
    Bz[N - 1] = wz;
    *b_out = b; *Kb_out = Kb; *Kz_out = Kz;
}

/* inv(A,sigma::Float64,tau): DPR1 of order N+1, slot (0,-1) at the arrow-top position (last). */
static void arrow_inv_shift_value(int64_t N, const double *D, const double *z, double a, double sigma,
                                  double tol, double *D1, double *u1, double *rho_out, double *Krho_out,
                                  int64_t *Qout, int32_t *status) {
    for (int64_t k = 0; k < N; ++k) {
        D1[k] = 1.0 / (D[k] - sigma);
        u1[k] = z[k] * D1[k];
    }
    D1[N] = 0.0;
    u1[N] = -1.0;
    double as = a - sigma;
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < N; ++k) {
        if (D1[k] > 0.0) P += z[k] * z[k] * D1[k]; else Q += z[k] * z[k] * D1[k];
    }
    if (as < 0) P -= as; else Q -= as;
    double Krho = (P - Q) / fabs(P + Q);
    double rho;
    if (Krho < tol) {
        rho = -1.0 / (P + Q);
    } else {
        *Qout = 1;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        dd s1 = dd_from(sigma);
        dd ad = dd_sub(dd_from(a), s1);
        for (int64_t k = 0; k < N; ++k) {
            dd zk = dd_from(z[k]);
            dd t = dd_div(dd_mul(zk, zk), dd_sub(dd_from(D[k]), s1));
            if (D1[k] > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        if (ad.hi + ad.lo < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        dd S = dd_add(Pd, Qd);
        if (!dd_is_pluszero(S)) {
            Krho = dd_to_double(dd_div(dd_sub(Pd, Qd), dd_abs(S)));
            dd r = dd_div(dd_from(1.0), S);
            rho = -(r.hi + r.lo);
        } else {
            *Qout = 100;
            *status |= AH_ST_NEEDS_BIGFLOAT;
            rho = -1.0 / (P + Q);
        }
    }
    *rho_out = rho; *Krho_out = Krho;
}

/* k-th eigenpair (1-based, descending) of an ordered irreducible arrow, arrow top last.
 * v has N+1 entries.  tau = [tolb,tolz,tolnu,tolrho,tollambda]. */
static void symarrow_k(int64_t N, const double *D, const double *z, double a, int64_t k,
                       const double *tau, scratch_t *s, double *lambda_out, double *v,
                       int64_t *sind, double *kb, double *kz, double *knu, double *krho,
                       int64_t *qout, int32_t *steps_out, int32_t *status_out) {
    const int64_t n = N + 1;
    double Kb = 0, Kz = 0, Knu = 0, Krho = 0;
    int64_t Qout = 0;
    int32_t steps = 0, status = 0;
    double sigma, lambda;
    int64_t i;
    char side;
    for (int64_t l = 0; l < n; ++l) v[l] = 0.0;

    if (k == 1) { sigma = D[0]; i = 1; side = 'R'; }
    else if (k == n) { sigma = D[N - 1]; i = N; side = 'L'; }
    else {
        double Dk = D[k - 1];
        double atemp = a - Dk;
        double middle = (D[k - 2] - Dk) / 2.0;
        for (int64_t l = 0; l < N; ++l) s->t[l] = (z[l] * z[l]) / ((D[l] - Dk) - middle);
        double Fmiddle = (atemp - middle) - jl_pairwise_sum(s->t, 0, N - 1);
        if (Fmiddle < 0.0) { sigma = Dk; i = k; side = 'R'; }
        else { sigma = D[k - 2]; i = k - 1; side = 'L'; }
    }

    double b;
    arrow_inv_shift_index(N, D, z, a, i, tau[0], tau[1], s->bd, s->bz, &b, &Kb, &Kz, &Qout, &status);
    if (Qout == 1) status |= AH_BR_DD_B;
    double nu = bisect_arrow(s->bd, s->bz, b, N, side, s->w, &steps);

    if (isinf(nu)) {
        v[i - 1] = 1.0;
        lambda = sigma;
        status |= AH_ST_NU_INF;
    } else {
        double nu1 = 0.0;
        for (int64_t l = 0; l < N; ++l) nu1 = jl_max(nu1, fabs(s->bd[l]) + fabs(s->bz[l]));
        for (int64_t l = 0; l < N; ++l) s->t[l] = fabs(s->bz[l]);
        nu1 = jl_max(jl_pairwise_sum(s->t, 0, N - 1) + fabs(b), nu1);
        Knu = nu1 / fabs(nu);
        while (Knu > tau[2]) {
            status |= (status & AH_BR_REM3) ? AH_BR_REM3_REPEAT : AH_BR_REM3;
            nu = side == 'R' ? fabs(nu) : -fabs(nu);
            nu1 = -jl_sign(nu) * nu1;
            double sigma1 = (nu1 + nu) / (2.0 * nu * nu1) + sigma;
            int pole = 0;
            for (int64_t l = 0; l < N; ++l) if (jl_isequal(sigma1, D[l])) { pole = 1; break; }
            if (pole) {
                /* arrowhead3.jl:474-496; line 481 raises in the reference */
                status |= AH_ST_POLE_RETURN;
                break;
            }
            int64_t Qout1 = 0;
            double rho;
            arrow_inv_shift_value(N, D, z, a, sigma1, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1, &status);
            if (Qout1 != 0) status |= AH_BR_DD_RHO;
            nu = bisect_dpr1(s->d1, s->u1, rho, N + 1, side, s->w, &steps);
            if (side == 'R' && D[0] > 0.0 && nu < 0.0) {
                status |= AH_BR_RL_RETRY;
                nu = bisect_dpr1(s->d1, s->u1, rho, N + 1, 'L', s->w, &steps);
            }
            double mD = fabs(s->d1[0]);
            for (int64_t l = 1; l <= N; ++l) mD = jl_max(mD, fabs(s->d1[l]));
            acc_t uu = 0.0;
            for (int64_t l = 0; l <= N; ++l) uu += s->u1[l] * s->u1[l];
            nu1 = mD + fabs(rho) * uu;
            Knu = nu1 / fabs(nu);
            sigma = sigma1;
            Qout = Qout + 2 * Qout1;
            if (isnan(Knu)) break; /* guard: `while NaN > tau` is false in Julia too */
        }
        double mu = 1.0 / nu;
        for (int64_t l = 0; l < N; ++l) v[l] = z[l] / ((D[l] - sigma) - mu);
        v[N] = -1.0;
        lambda = mu + sigma;
        acc_t ss = 0.0;
        for (int64_t l = 0; l < n; ++l) ss += v[l] * v[l];
        double inrm = 1.0 / sqrt(ss);
        for (int64_t l = 0; l < n; ++l) v[l] *= inrm;

        if ((fabs(D[i - 1]) + fabs(1.0 / nu)) / fabs(lambda) > tau[4]) {
            int c1 = (k == 1 && D[0] < 0.0);
            int c2 = (k == n && D[N - 1] > 0.0);
            int c3 = (i < n - 1 && side == 'L' && jl_sign(D[i - 1]) + jl_sign(D[i]) == 0);
            int c4 = (i > 1 && side == 'R' && jl_sign(D[i - 1]) + jl_sign(D[i - 2]) == 0);
            if (c1 || c2 || c3 || c4) {
                int64_t Qout1 = 0;
                double rho;
                arrow_inv_shift_value(N, D, z, a, 0.0, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1, &status);
                status |= AH_BR_REM1;
                if (Qout1 != 0) status |= AH_BR_DD_RHO;
                Qout = Qout + 4 * Qout1;
                if (isinf(rho)) {
                    status |= AH_BR_REM1_ZERO;
                    lambda = 0.0;
                } else {
                    for (int64_t l = 0; l < n; ++l) s->t[l] = v[l] * v[l] * s->d1[l];
                    double s1 = jl_pairwise_sum(s->t, 0, n - 1);
                    for (int64_t l = 0; l < n; ++l) s->t[l] = v[l] * s->u1[l];
                    double s2 = jl_pairwise_sum(s->t, 0, n - 1);
                    double nu1r = s1 + rho * (s2 * s2);
                    lambda = 1.0 / nu1r;
                }
            }
        }
    }
    *lambda_out = lambda;
    *sind = i; *kb = Kb; *kz = Kz; *knu = Knu; *krho = Krho; *qout = Qout;
    *steps_out = steps; *status_out = status;
}

/* ------------------------------------------------------------------ rootsah: eigenvalue-only, double-double border */
/* inv(A,zD,alphaD,i,tau) of arrowhead7.jl:132-225: as arrow_inv_shift_index, but the double-double recomputation of b
 * uses the supplied double-double border zD = (z, zlo) and arrow top alphaD = (a, alo), there is a `Kb == Inf -> b = 0`
 * branch and no BigFloat branch. */
static void roots_inv_shift_index(int64_t N, const double *D, const double *z, const double *zlo, double a, double alo,
                                  int64_t i, double tolb, double tolz, double *Bd, double *Bz, double *b_out,
                                  double *Kb_out, double *Kz_out, int64_t *Qout) {
    const int64_t ii = i - 1;
    double wz = 1.0 / z[ii];
    double sigma = D[ii];
    for (int64_t k = 0; k < ii; ++k) { Bd[k] = 1.0 / (D[k] - sigma); Bz[k] = -z[k] * Bd[k] * wz; }
    for (int64_t k = ii + 1; k < N; ++k) { Bd[k - 1] = 1.0 / (D[k] - sigma); Bz[k - 1] = -z[k] * Bd[k - 1] * wz; }
    double as = a - sigma;
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < ii; ++k) { if (Bd[k] > 0.0) P += z[k] * z[k] * Bd[k]; else Q += z[k] * z[k] * Bd[k]; }
    for (int64_t k = ii; k < N - 1; ++k) { if (Bd[k] > 0.0) P += z[k + 1] * z[k + 1] * Bd[k]; else Q += z[k + 1] * z[k + 1] * Bd[k]; }
    if (as < 0) P -= as; else Q -= as;
    double Kb = (P - Q) / fabs(P + Q);
    double mz = fabs(z[0]);
    for (int64_t k = 1; k < N; ++k) mz = jl_max(mz, fabs(z[k]));
    double Kz = mz * fabs(wz);
    double b;
    if (Kb < tolb || Kz < tolz) {
        b = (P + Q) * wz * wz;
    } else if (isinf(Kb)) {
        b = 0.0;
    } else {
        *Qout = 1;
        dd s1 = dd_from(sigma);
        dd wzd = {z[ii], zlo[ii]};
        dd aD = {a, alo};
        dd ad = dd_sub(aD, s1);
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        for (int64_t k = 0; k < N; ++k) {
            if (k == ii) continue;
            dd Dd = dd_sub(dd_from(D[k]), s1);
            dd zk = {z[k], zlo[k]};
            dd t = dd_div(dd_mul(zk, zk), Dd);
            if (Dd.hi > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        if (ad.hi < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        dd bd = dd_div(dd_add(Pd, Qd), dd_mul(wzd, wzd));
        b = dd_to_double(bd);
    }
    Bd[N - 1] = 0.0;
    Bz[N - 1] = wz;
    *b_out = b; *Kb_out = Kb; *Kz_out = Kz;
}

/* inv(A,zD,alphaD,sigma::Float64,tau) of arrowhead7.jl:232-292 */
static void roots_inv_shift_value(int64_t N, const double *D, const double *z, const double *zlo, double a, double alo,
                                  double sigma, double tol, double *D1, double *u1, double *rho_out, double *Krho_out,
                                  int64_t *Qout) {
    for (int64_t k = 0; k < N; ++k) { D1[k] = 1.0 / (D[k] - sigma); u1[k] = z[k] * D1[k]; }
    D1[N] = 0.0;
    u1[N] = -1.0;
    double as = a - sigma;
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < N; ++k) { if (D1[k] > 0.0) P += z[k] * z[k] * D1[k]; else Q += z[k] * z[k] * D1[k]; }
    if (as < 0) P -= as; else Q -= as;
    double Krho = (P - Q) / fabs(P + Q);
    double rho;
    if (Krho < tol) {
        rho = -1.0 / (P + Q);
    } else {
        *Qout = 1;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        dd s1 = dd_from(sigma);
        dd aD = {a, alo};
        dd ad = dd_sub(aD, s1);
        for (int64_t k = 0; k < N; ++k) {
            dd zk = {z[k], zlo[k]};
            dd t = dd_div(dd_mul(zk, zk), dd_sub(dd_from(D[k]), s1));
            if (D1[k] > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        if (ad.hi + ad.lo < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        dd r = dd_div(dd_from(1.0), dd_add(Pd, Qd));
        rho = -(r.hi + r.lo);
    }
    *rho_out = rho; *Krho_out = Krho;
}

/* eigen(A,zD,alphaD,k,tau) of arrowhead7.jl:351-495: k-th eigenvalue only.  status: AH_ST_POLE_RETURN where the
 * reference raises (:436 `map(Double,A.D)-sigma_v`), AH_ST_REF_THROWS for the unguarded A.D[i-1] with i == 1 (:474). */
static void rootsah_k(int64_t N, const double *D, const double *z, const double *zlo, double a, double alo, int64_t k,
                      const double *tau, scratch_t *s, double *lambda_out, int64_t *sind, double *kb, double *kz,
                      double *knu, double *krho, int64_t *qout, int32_t *steps_out, int32_t *status_out) {
    const int64_t n = N + 1;
    double Kb = 0, Kz = 0, Knu = 0, Krho = 0;
    int64_t Qout = 0;
    int32_t steps = 0, status = 0;
    double sigma, lambda;
    int64_t i;
    char side;
    if (k == 1) { sigma = D[0]; i = 1; side = 'R'; }
    else if (k == n) { sigma = D[N - 1]; i = N; side = 'L'; }
    else {   /* shift selection in double-double (:381-389) */
        dd Dk = dd_from(D[k - 1]);
        dd middle = dd_div(dd_sub(dd_from(D[k - 2]), Dk), dd_from(2.0));
        dd F = dd_from(0.0);
        for (int64_t l = 0; l < N; ++l) {
            dd zl = {z[l], zlo[l]};
            F = dd_add(F, dd_div(dd_mul(zl, zl), dd_sub(dd_sub(dd_from(D[l]), Dk), middle)));
        }
        dd aD = {a, alo};
        F = dd_sub(dd_sub(dd_sub(aD, Dk), middle), F);
        if (F.hi + F.lo < 0.0) { sigma = D[k - 1]; i = k; side = 'R'; }
        else { sigma = D[k - 2]; i = k - 1; side = 'L'; }
    }
    double b;
    roots_inv_shift_index(N, D, z, zlo, a, alo, i, tau[0], tau[1], s->bd, s->bz, &b, &Kb, &Kz, &Qout);
    double nu = bisect_arrow(s->bd, s->bz, b, N, side, s->w, &steps);
    if (isinf(nu)) {
        lambda = sigma;
        status |= AH_ST_NU_INF;
    } else {
        double nu1 = 0.0;
        for (int64_t l = 0; l < N; ++l) nu1 = jl_max(nu1, fabs(s->bd[l]) + fabs(s->bz[l]));
        for (int64_t l = 0; l < N; ++l) s->t[l] = fabs(s->bz[l]);
        nu1 = jl_max(jl_pairwise_sum(s->t, 0, N - 1) + fabs(b), nu1);
        Knu = nu1 / fabs(nu);
        double sig_v = sigma, mu;
        if (Knu < tau[2]) {
            mu = 1.0 / nu;
            lambda = mu + sigma;
        } else {   /* Remedy 3, once (:415-459); nu keeps its first-stage value for the Remedy-1 test below */
            nu = side == 'R' ? fabs(nu) : -fabs(nu);
            nu1 = -jl_sign(nu) * nu1;
            double sigma1 = (nu1 + nu) / (2.0 * nu * nu1) + sigma;
            int pole = 0;
            for (int64_t l = 0; l < N; ++l) if (jl_isequal(sigma1, D[l])) { pole = 1; break; }
            if (pole) {
                status |= AH_ST_POLE_RETURN;
                mu = 1.0 / nu;
                lambda = mu + sigma;
            } else {
                int64_t Qout1 = 0;
                double rho;
                roots_inv_shift_value(N, D, z, zlo, a, alo, sigma1, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1);
                double nub = bisect_dpr1(s->d1, s->u1, rho, N + 1, side, s->w, &steps);
                mu = 1.0 / nub;
                sig_v = sigma1;
                lambda = mu + sigma1;
            }
            if (Qout == 1) Qout = Qout + 2;
        }
        /* Remedy 1 (:464-487); the inverse at 0 is the generic inv(A,0.0,tau) of arrowhead3.jl (plain z) */
        if ((fabs(D[i - 1]) + fabs(1.0 / nu)) / fabs(lambda) > tau[4]) {
            int c1 = (k == 1 && D[0] < 0.0);
            int c2 = (k == n && D[N - 1] > 0.0);
            int c3 = (i < n - 1 && side == 'L' && jl_sign(D[i - 1]) + jl_sign(D[i]) == 0);
            int c4 = 0;
            if (!(c1 || c2 || c3) && side == 'R') {
                if (i == 1) status |= AH_ST_REF_THROWS;      /* A.D[0]: BoundsError in the reference */
                else c4 = (jl_sign(D[i - 1]) + jl_sign(D[i - 2]) == 0);
            }
            if (c1 || c2 || c3 || c4) {
                int64_t Qout1 = 0;
                double rho;
                arrow_inv_shift_value(N, D, z, a, 0.0, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1, &status);
                if (Qout == 1) Qout = Qout + 4;
                if (isinf(rho)) {
                    lambda = 0.0;
                } else {
                    /* v = normalize([z./((D-sig_v)-mu); -1]) */
                    double *v = s->t;
                    acc_t ss = 0.0;
                    for (int64_t l = 0; l < N; ++l) { v[l] = z[l] / ((D[l] - sig_v) - mu); ss += v[l] * v[l]; }
                    v[N] = -1.0; ss += 1.0;
                    double inrm = 1.0 / sqrt(ss);
                    acc_t s1 = 0.0, s2 = 0.0;
                    for (int64_t l = 0; l < n; ++l) { double vl = v[l] * inrm; s1 += vl * vl * s->d1[l]; s2 += vl * s->u1[l]; }
                    double nu1r = s1 + rho * (s2 * s2);
                    lambda = 1.0 / nu1r;
                }
            }
        }
    }
    *lambda_out = lambda;
    *sind = i; *kb = Kb; *kz = Kz; *knu = Knu; *krho = Krho; *qout = Qout;
    *steps_out = steps; *status_out = status;
}

/* ------------------------------------------------------------------ SymDPR1 */
/* inv!(B,A::SymDPR1,i,tau): B gets n-1 entries, no slot. */
static void dpr1_inv_shift_index(int64_t n, const double *D, const double *u, double r, int64_t i,
                                 double tolb, double tolz, double *Bd, double *Bz, double *b_out,
                                 double *Kb_out, double *Kz_out, int64_t *Qout, int32_t *status,
                                 double *tmp) {
    const int64_t ii = i - 1;
    double wz = 1.0 / u[ii];
    double sigma = D[ii];
    for (int64_t k = 0; k < ii; ++k) {
        Bd[k] = 1.0 / (D[k] - sigma);
        Bz[k] = -u[k] * Bd[k] * wz;
    }
    for (int64_t k = ii + 1; k < n; ++k) {
        Bd[k - 1] = 1.0 / (D[k] - sigma);
        Bz[k - 1] = -u[k] * Bd[k - 1] * wz;
    }
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < ii; ++k) {
        if (Bd[k] > 0.0) P += u[k] * u[k] * Bd[k]; else Q += u[k] * u[k] * Bd[k];
    }
    for (int64_t k = ii; k < n - 1; ++k) {
        if (Bd[k] > 0.0) P += u[k + 1] * u[k + 1] * Bd[k]; else Q += u[k + 1] * u[k + 1] * Bd[k];
    }
    if (r > 0) P += 1.0 / r; else Q += 1.0 / r;
    double Kb = (P - Q) / fabs(P + Q);
    for (int64_t k = 0; k < n; ++k) tmp[k] = fabs(u[k]);
    double Kz = (jl_pairwise_sum(tmp, 0, n - 1) - fabs(u[ii])) * fabs(wz);
    double b;
    if (Kb < tolb || Kz < tolz) {
        b = (P + Q) * wz * wz;
    } else if (Kz < 1.0 / (AH_EPS * AH_EPS)) {
        *Qout = 1;
        dd s1 = dd_from(sigma);
        dd wzd = dd_from(u[ii]);
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        for (int64_t k = 0; k < n; ++k) {
            if (k == ii) continue;
            dd Dd = dd_sub(dd_from(D[k]), s1);
            dd uk = dd_from(u[k]);
            dd t = dd_div(dd_mul(uk, uk), Dd);
            if (dd_to_double(Dd) > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        dd ir = dd_div(dd_from(1.0), dd_from(r));
        if (r > 0) Pd = dd_add(Pd, ir); else Qd = dd_add(Qd, ir);
        dd bd = dd_div(dd_add(Pd, Qd), dd_mul(wzd, wzd));
        b = dd_to_double(bd);
    } else {
        *Qout = 50;
        *status |= AH_ST_NEEDS_BIGFLOAT;
        b = (P + Q) * wz * wz;
    }
    *b_out = b; *Kb_out = Kb; *Kz_out = Kz;
}

/* inv(A::SymDPR1,sigma::Float64,tau): DPR1 of order n. */
static void dpr1_inv_shift_value(int64_t n, const double *D, const double *u, double r, double sigma,
                                 double tol, double *D1, double *u1, double *rho_out, double *Krho_out,
                                 int64_t *Qout, int32_t *status) {
    for (int64_t k = 0; k < n; ++k) {
        D1[k] = 1.0 / (D[k] - sigma);
        u1[k] = u[k] * D1[k];
    }
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < n; ++k) {
        if (D1[k] > 0.0) P += u[k] * u[k] * D1[k]; else Q += u[k] * u[k] * D1[k];
    }
    if (r > 0) P += 1.0 / r; else Q += 1.0 / r;
    double Krho = (P - Q) / fabs(P + Q);
    double rho;
    if (Krho < tol) {
        rho = -1.0 / (P + Q);
    } else {
        *Qout = 1;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        dd s1 = dd_from(sigma);
        for (int64_t k = 0; k < n; ++k) {
            dd uk = dd_from(u[k]);
            dd t = dd_div(dd_mul(uk, uk), dd_sub(dd_from(D[k]), s1));
            if (D1[k] > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        dd ir = dd_div(dd_from(1.0), dd_from(r));
        if (r > 0) Pd = dd_add(Pd, ir); else Qd = dd_add(Qd, ir);
        dd S = dd_add(Pd, Qd);
        Krho = dd_to_double(dd_div(dd_sub(Pd, Qd), dd_abs(S)));
        if (Krho < tol / AH_EPS) {
            dd rr = dd_div(dd_from(1.0), S);
            rho = -(rr.hi + rr.lo);
        } else {
            *Qout = 100;
            *status |= AH_ST_NEEDS_BIGFLOAT;
            dd rr = dd_div(dd_from(1.0), S);
            rho = -(rr.hi + rr.lo);
        }
    }
    *rho_out = rho; *Krho_out = Krho;
}

static void symdpr1_k(int64_t n, const double *D, const double *u, double r, int64_t k,
                      const double *tau, scratch_t *s, double *lambda_out, double *v,
                      int64_t *sind, double *kb, double *kz, double *knu, double *krho,
                      int64_t *qout, int32_t *steps_out, int32_t *status_out) {
    double Kb = 0, Kz = 0, Knu = 0, Krho = 0;
    int64_t Qout = 0;
    int32_t steps = 0, status = 0;
    double sigma, lambda;
    int64_t i;
    char side;
    for (int64_t l = 0; l < n; ++l) v[l] = 0.0;

    if (k == 1) { sigma = D[0]; i = 1; side = 'R'; }
    else {
        double Dk = D[k - 1];
        double middle = (D[k - 2] - Dk) / 2.0;
        for (int64_t l = 0; l < n; ++l) s->t[l] = (u[l] * u[l]) / ((D[l] - Dk) - middle);
        double Fmiddle = 1.0 + r * jl_pairwise_sum(s->t, 0, n - 1);
        if (Fmiddle > 0.0) { sigma = Dk; i = k; side = 'R'; }
        else { sigma = D[k - 2]; i = k - 1; side = 'L'; }
    }
    double b;
    dpr1_inv_shift_index(n, D, u, r, i, tau[0], tau[1], s->bd, s->bz, &b, &Kb, &Kz, &Qout, &status, s->t);
    if (Qout == 1) status |= AH_BR_DD_B;
    double nu = bisect_arrow(s->bd, s->bz, b, n - 1, side, s->w, &steps);

    if (isinf(nu)) {
        v[i - 1] = 1.0;
        lambda = sigma;                 /* arrowhead4.jl:324 reads an undefined name: reference throws */
        status |= AH_ST_NU_INF | AH_ST_REF_THROWS;
    } else {
        double nu1 = 0.0;
        for (int64_t l = 0; l < n - 1; ++l) nu1 = jl_max(nu1, fabs(s->bd[l]) + fabs(s->bz[l]));
        for (int64_t l = 0; l < n - 1; ++l) s->t[l] = fabs(s->bz[l]);
        nu1 = jl_max(jl_pairwise_sum(s->t, 0, n - 2) + fabs(b), nu1);
        Knu = nu1 / fabs(nu);
        while (Knu > tau[2]) {
            status |= (status & AH_BR_REM3) ? AH_BR_REM3_REPEAT : AH_BR_REM3;
            nu = side == 'R' ? fabs(nu) : -fabs(nu);
            nu1 = -jl_sign(nu) * nu1;
            double sigma1 = (nu1 + nu) / (2.0 * nu * nu1) + sigma;
            int pole = 0;
            for (int64_t l = 0; l < n; ++l) if (jl_isequal(sigma1, D[l])) { pole = 1; break; }
            if (pole) { status |= AH_ST_POLE_RETURN; break; }
            int64_t Qout1 = 0;
            double rho;
            dpr1_inv_shift_value(n, D, u, r, sigma1, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1, &status);
            if (Qout1 != 0) status |= AH_BR_DD_RHO;
            nu = bisect_dpr1(s->d1, s->u1, rho, n, side, s->w, &steps);
            double mD = fabs(s->d1[0]);
            for (int64_t l = 1; l < n; ++l) mD = jl_max(mD, fabs(s->d1[l]));
            acc_t uu = 0.0;
            for (int64_t l = 0; l < n; ++l) uu += s->u1[l] * s->u1[l];
            nu1 = mD + fabs(rho) * uu;
            Knu = nu1 / fabs(nu);
            sigma = sigma1;
            Qout = Qout + 2 * Qout1;
            if (isnan(Knu)) break;
        }
        double mu = 1.0 / nu;
        for (int64_t l = 0; l < n; ++l) v[l] = u[l] / ((D[l] - sigma) - mu);
        lambda = mu + sigma;
        acc_t ss = 0.0;
        for (int64_t l = 0; l < n; ++l) ss += v[l] * v[l];
        double inrm = 1.0 / sqrt(ss);
        for (int64_t l = 0; l < n; ++l) v[l] *= inrm;

        if ((fabs(D[i - 1]) + fabs(1.0 / nu)) / fabs(lambda) > tau[4]) {
            int c1 = (k == 1 && D[0] < 0.0);
            /* side=='L' implies i=k-1<=n-1, so D[i+1] (1-based) exists */
            int c2 = (side == 'L' && jl_sign(D[i - 1]) + jl_sign(D[i]) == 0);
            int c3 = (i > 1 && side == 'R' && jl_sign(D[i - 1]) + jl_sign(D[i - 2]) == 0);
            if (c1 || c2 || c3) {
                int64_t Qout1 = 0;
                double rho;
                dpr1_inv_shift_value(n, D, u, r, 0.0, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1, &status);
                status |= AH_BR_REM1;
                if (Qout1 != 0) status |= AH_BR_DD_RHO;
                Qout = Qout + 4 * Qout1;
                if (isinf(rho)) {
                    status |= AH_BR_REM1_ZERO;
                    lambda = 0.0;
                } else {
                    for (int64_t l = 0; l < n; ++l) s->t[l] = v[l] * v[l] * s->d1[l];
                    double s1 = jl_pairwise_sum(s->t, 0, n - 1);
                    for (int64_t l = 0; l < n; ++l) s->t[l] = v[l] * s->u1[l];
                    double s2 = jl_pairwise_sum(s->t, 0, n - 1);
                    lambda = 1.0 / (s1 + rho * (s2 * s2));
                }
            }
        }
    }
    *lambda_out = lambda;
    *sind = i; *kb = Kb; *kz = Kz; *knu = Knu; *krho = Krho; *qout = Qout;
    *steps_out = steps; *status_out = status;
}

/* ------------------------------------------------------------------ HalfArrow */
/* doubledot([z;s],[z;-s]) with nz entries of z */
static dd half_a_dd(int64_t nz, const double *z, double s) {
    dd acc = dd_from(0.0);
    for (int64_t k = 0; k < nz; ++k) acc = dd_add(acc, dd_mul(dd_from(z[k]), dd_from(z[k])));
    acc = dd_add(acc, dd_mul(dd_from(s), dd_from(-s)));
    return acc;
}

/* inv!(B,A::HalfArrow,i,tau): nd poles; B gets nd entries (nd-1 poles + slot). */
static void half_inv_shift_index(int64_t nd, int64_t nz, const double *D, const double *z,
                                 const double *z1, int64_t i, double tolb, double tolz, double *Bd,
                                 double *Bz, double *b_out, double *Kb_out, double *Kz_out,
                                 int64_t *Qout) {
    const int64_t ii = i - 1;
    double wz = 1.0 / z1[ii];
    double sigma = D[ii];
    for (int64_t k = 0; k < ii; ++k) {
        Bd[k] = 1.0 / ((D[k] - sigma) * (D[k] + sigma));
        Bz[k] = -z1[k] * Bd[k] * wz;
    }
    for (int64_t k = ii + 1; k < nd; ++k) {
        Bd[k - 1] = 1.0 / ((D[k] - sigma) * (D[k] + sigma));
        Bz[k - 1] = -z1[k] * Bd[k - 1] * wz;
    }
    dd ad = half_a_dd(nz, z, sigma);
    double a = ad.hi + ad.lo;
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    for (int64_t k = 0; k < ii; ++k) {
        if (Bd[k] > 0.0) P += z1[k] * z1[k] * Bd[k]; else Q += z1[k] * z1[k] * Bd[k];
    }
    for (int64_t k = ii; k < nd - 1; ++k) {
        if (Bd[k] > 0.0) P += z1[k + 1] * z1[k + 1] * Bd[k]; else Q += z1[k + 1] * z1[k + 1] * Bd[k];
    }
    if (a < 0) P -= a; else Q -= a;
    double Kb = (P - Q) / fabs(P + Q);
    double mz = fabs(z1[0]);
    for (int64_t k = 1; k < nd; ++k) mz = jl_max(mz, fabs(z1[k]));
    double Kz = mz * fabs(wz);
    double b;
    if (Kb < tolb || Kz < tolz) {
        b = (P + Q) * wz * wz;
    } else {
        *Qout = 1;
        dd s1 = dd_from(sigma);
        dd wzd = dd_mul(dd_from(z[ii]), s1);
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        for (int64_t k = 0; k < nd; ++k) {
            if (k == ii) continue;
            dd Dk = dd_from(D[k]);
            dd Dd = dd_mul(dd_add(Dk, s1), dd_sub(Dk, s1));
            dd zd = dd_mul(dd_from(z[k]), Dk);
            dd t = dd_div(dd_mul(zd, zd), Dd);
            if (Dd.hi > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        if (ad.hi < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        dd bd = dd_div(dd_add(Pd, Qd), dd_mul(wzd, wzd));
        b = bd.hi + bd.lo;
    }
    Bd[nd - 1] = 0.0;
    Bz[nd - 1] = wz;
    *b_out = b; *Kb_out = Kb; *Kz_out = Kz;
}

/* inv(A::HalfArrow,sigma::Float64,tau): DPR1 of order nd+1 with slot (0,-1) last. */
static void half_inv_shift_value(int64_t nd, int64_t nz, const double *D, const double *z,
                                 const double *z1, double sigma, double tol, double *D1, double *u1,
                                 double *rho_out, double *Krho_out, int64_t *Qout) {
    for (int64_t k = 0; k < nd; ++k) {
        D1[k] = 1.0 / ((D[k] - sigma) * (D[k] + sigma));
        u1[k] = z1[k] * D1[k];
    }
    D1[nd] = 0.0;
    u1[nd] = -1.0;
    acc_t P = 0.0, Q = 0.0;
    *Qout = 0;
    dd ad = half_a_dd(nz, z, sigma);
    double a = ad.hi + ad.lo;
    for (int64_t k = 0; k < nd; ++k) {
        if (D1[k] > 0.0) P += z1[k] * z1[k] * D1[k]; else Q += z1[k] * z1[k] * D1[k];
    }
    if (a < 0) P -= a; else Q -= a;
    double Krho = (P - Q) / fabs(P + Q);
    double rho;
    if (Krho < tol) {
        rho = -1.0 / (P + Q);
    } else {
        *Qout = 1;
        dd Pd = dd_from(0.0), Qd = dd_from(0.0);
        dd s1 = dd_from(sigma);
        for (int64_t k = 0; k < nd; ++k) {
            dd Dk = dd_from(D[k]);
            dd Dd = dd_mul(dd_add(Dk, s1), dd_sub(Dk, s1));
            dd zd = dd_mul(dd_from(z[k]), Dk);
            dd t = dd_div(dd_mul(zd, zd), Dd);
            if (Dd.hi > 0.0) Pd = dd_add(Pd, t); else Qd = dd_add(Qd, t);
        }
        if (ad.hi + ad.lo < 0) Pd = dd_sub(Pd, ad); else Qd = dd_sub(Qd, ad);
        dd r = dd_div(dd_from(1.0), dd_add(Pd, Qd));
        rho = -(r.hi + r.lo);
    }
    *rho_out = rho; *Krho_out = Krho;
}

/* k-th singular triple of an ordered irreducible HalfArrow [diag(D) z], D>0 decreasing.
 * nz = nd (no tip) or nd+1 (tip).  u has nz entries, v has nd+1 entries. */
static void halfarrow_k(int64_t nd, int64_t nz, const double *D, const double *z, int64_t k,
                        const double *tau, scratch_t *s, double *lambda_out, double *u, double *v,
                        int64_t *sind, double *kb, double *kz, double *knu, double *krho,
                        int64_t *qout, int32_t *steps_out, int32_t *status_out) {
    const int64_t n = nd + 1;
    double Kb = 0, Kz = 0, Knu = 0, Krho = 0;
    int64_t Qout = 0;
    int32_t steps = 0, status = 0;
    double sigma, lambda;
    int64_t i;
    char side;
    double *z1 = s->z1;
    for (int64_t l = 0; l < nz; ++l) u[l] = 0.0;
    for (int64_t l = 0; l < n; ++l) v[l] = 0.0;
    for (int64_t l = 0; l < nd; ++l) z1[l] = z[l] * D[l];

    if (k == 1) { sigma = D[0]; i = 1; side = 'R'; }
    else if (k == n) { sigma = D[nd - 1]; i = nd; side = 'L'; }
    else {
        double Dk = D[k - 1];
        acc_t zz = 0.0;
        for (int64_t l = 0; l < nz; ++l) zz += z[l] * z[l];
        double atemp = zz - Dk * Dk;
        double middle = ((D[k - 2] - Dk) * (D[k - 2] + Dk)) / 2.0;
        for (int64_t l = 0; l < nd; ++l)
            s->t[l] = (z1[l] * z1[l]) / ((D[l] - Dk) * (D[l] + Dk) - middle);
        double Fmiddle = (atemp - middle) - jl_pairwise_sum(s->t, 0, nd - 1);
        if (Fmiddle < 0.0) { sigma = Dk; i = k; side = 'R'; }
        else { sigma = D[k - 2]; i = k - 1; side = 'L'; }
    }
    double b;
    half_inv_shift_index(nd, nz, D, z, z1, i, tau[0], tau[1], s->bd, s->bz, &b, &Kb, &Kz, &Qout);
    if (Qout == 1) status |= AH_BR_DD_B;
    /* the reference's scratch arrow has one extra never-written slot in the tip shape
       (arrowhead6.jl:636 vs :206-207); it reads as zero in fresh memory and contributes nothing */
    double nu = bisect_arrow(s->bd, s->bz, b, nd, side, s->w, &steps);

    if (isinf(nu)) {
        u[i - 1] = 1.0;
        v[i - 1] = 1.0;
        lambda = sigma;
        status |= AH_ST_NU_INF;
    } else {
        double nu1 = 0.0;
        for (int64_t l = 0; l < nd; ++l) nu1 = jl_max(nu1, fabs(s->bd[l]) + fabs(s->bz[l]));
        for (int64_t l = 0; l < nd; ++l) s->t[l] = fabs(s->bz[l]);
        nu1 = jl_max(jl_pairwise_sum(s->t, 0, nd - 1) + fabs(b), nu1);
        Knu = nu1 / fabs(nu);
        double sig_v, mu_v, lam2;     /* shift, mu used for the vector; lambda^2 */
        int have_vec = 1;
        if (Knu < tau[2]) {
            mu_v = 1.0 / nu;
            sig_v = sigma;
            lam2 = mu_v + sigma * sigma;
        } else {
            status |= AH_BR_REM3;
            nu = side == 'R' ? fabs(nu) : -fabs(nu);
            nu1 = -jl_sign(nu) * nu1;
            double sigma1 = (nu1 + nu) / (2.0 * nu * nu1) + sigma * sigma;
            double ss1 = sqrt(sigma1);
            int pole = 0;
            for (int64_t l = 0; l < nd; ++l) if (jl_isequal(ss1, fabs(D[l]))) { pole = 1; break; }
            if (pole) {
                /* arrowhead6.jl:433-462 raises at :440 in the reference; keep the unremedied pair */
                status |= AH_ST_POLE_RETURN;
                sig_v = sigma; mu_v = 1.0 / nu; lam2 = mu_v + sigma * sigma;
            } else {
                int64_t Qout1 = 0;
                double rho;
                half_inv_shift_value(nd, nz, D, z, z1, ss1, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1);
                if (Qout1 != 0) status |= AH_BR_DD_RHO;
                double nu1b = bisect_dpr1(s->d1, s->u1, rho, nd + 1, side, s->w, &steps);
                mu_v = 1.0 / nu1b;
                sig_v = ss1;
                lam2 = mu_v + sigma1;
            }
            if (Qout == 1) Qout = Qout + 2;
        }
        if (have_vec) {
            for (int64_t l = 0; l < nd; ++l) v[l] = z1[l] / ((D[l] - sig_v) * (D[l] + sig_v) - mu_v);
            v[nd] = -1.0;
            acc_t ss = 0.0;
            for (int64_t l = 0; l < n; ++l) ss += v[l] * v[l];
            double inrm = 1.0 / sqrt(ss);
            for (int64_t l = 0; l < n; ++l) v[l] *= inrm;
            lambda = sqrt(lam2);
            for (int64_t l = 0; l < nd; ++l) u[l] = lambda * v[l] / D[l];
            if (nz == n) u[nd] = z[nd] * v[nd] / lambda;
        }
        if (k == n && (fabs(D[i - 1]) * fabs(D[i - 1]) + fabs(1.0 / nu)) / fabs(lambda * lambda) > tau[4]) {
            int64_t Qout1 = 0;
            double rho;
            half_inv_shift_value(nd, nz, D, z, z1, 0.0, tau[3], s->d1, s->u1, &rho, &Krho, &Qout1);
            status |= AH_BR_REM1;
            if (Qout1 != 0) status |= AH_BR_DD_RHO;
            if (Qout == 1) Qout = Qout + 4;
            if (isinf(rho)) {
                status |= AH_BR_REM1_ZERO;
                lambda = 0.0;
            } else {
                for (int64_t l = 0; l < n; ++l) s->t[l] = v[l] * v[l] * s->d1[l];
                double s1 = jl_pairwise_sum(s->t, 0, n - 1);
                for (int64_t l = 0; l < n; ++l) s->t[l] = v[l] * s->u1[l];
                double s2 = jl_pairwise_sum(s->t, 0, n - 1);
                lambda = sqrt(1.0 / (s1 + rho * (s2 * s2)));
            }
        }
    }
    *lambda_out = lambda;
    *sind = i; *kb = Kb; *kz = Kz; *knu = Knu; *krho = Krho; *qout = Qout;
    *steps_out = steps; *status_out = status;
}

/* ------------------------------------------------------------------ exported range drivers */
/* All take ordered irreducible inputs and a 1-based inclusive k-range; column c = k-k0 of the
 * column-major outputs (leading dimension ld) receives eigenvector k.  Info goes to SoA arrays of
 * length k1-k0+1.  U/V may be NULL (eigenvalues + info only).  Returns 0, or -1 on bad arguments,
 * -2 on allocation failure. */
int ah_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ah_oracle_symarrow_eigen(int64_t N, const double *D, const double *z, double a, const double *tau,
                             int64_t k0, int64_t k1, int nthreads, double *lambda, double *U, int64_t ldu,
                             int64_t *sind, double *kb, double *kz, double *knu, double *krho,
                             int64_t *qout, int32_t *steps, int32_t *status) {
    if (N < 1 || k0 < 1 || k1 > N + 1 || k1 < k0) return -1;
    if (U && ldu < N + 1) return -1;
    int err = 0;
    if (nthreads < 1) nthreads = ah_oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s;
        double *vloc = (double *)malloc((size_t)(N + 2) * 8);
        int ok = scratch_alloc(&s, N + 1) && vloc;
        if (!ok) {
#pragma omp atomic write
            err = -2;
        }
#pragma omp barrier
        if (!err) {
#pragma omp for schedule(static)
            for (int64_t k = k0; k <= k1; ++k) {
                int64_t c = k - k0;
                double *v = U ? U + c * ldu : vloc;
                symarrow_k(N, D, z, a, k, tau, &s, &lambda[c], v, &sind[c], &kb[c], &kz[c], &knu[c],
                           &krho[c], &qout[c], &steps[c], &status[c]);
            }
        }
        if (ok) scratch_free(&s);
        free(vloc);
    }
    return err;
}

int ah_oracle_symdpr1_eigen(int64_t n, const double *D, const double *u, double rho, const double *tau,
                            int64_t k0, int64_t k1, int nthreads, double *lambda, double *U, int64_t ldu,
                            int64_t *sind, double *kb, double *kz, double *knu, double *krho,
                            int64_t *qout, int32_t *steps, int32_t *status) {
    if (n < 2 || k0 < 1 || k1 > n || k1 < k0) return -1;
    if (U && ldu < n) return -1;
    int err = 0;
    if (nthreads < 1) nthreads = ah_oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s;
        double *vloc = (double *)malloc((size_t)(n + 2) * 8);
        int ok = scratch_alloc(&s, n + 1) && vloc;
        if (!ok) {
#pragma omp atomic write
            err = -2;
        }
#pragma omp barrier
        if (!err) {
#pragma omp for schedule(static)
            for (int64_t k = k0; k <= k1; ++k) {
                int64_t c = k - k0;
                double *v = U ? U + c * ldu : vloc;
                symdpr1_k(n, D, u, rho, k, tau, &s, &lambda[c], v, &sind[c], &kb[c], &kz[c], &knu[c],
                          &krho[c], &qout[c], &steps[c], &status[c]);
            }
        }
        if (ok) scratch_free(&s);
        free(vloc);
    }
    return err;
}

int ah_oracle_halfarrow_svd(int64_t nd, int64_t nz, const double *D, const double *z, const double *tau,
                            int64_t k0, int64_t k1, int nthreads, double *sigma, double *U, int64_t ldu,
                            double *V, int64_t ldv, int64_t *sind, double *kb, double *kz, double *knu,
                            double *krho, int64_t *qout, int32_t *steps, int32_t *status) {
    if (nd < 1 || (nz != nd && nz != nd + 1)) return -1;
    int64_t kmax = nz;   /* number of singular triples computed by the driver loop */
    if (k0 < 1 || k1 > kmax || k1 < k0) return -1;
    if (U && ldu < nz) return -1;
    if (V && ldv < nd + 1) return -1;
    int err = 0;
    if (nthreads < 1) nthreads = ah_oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s;
        double *uloc = (double *)malloc((size_t)(nd + 2) * 8);
        double *vloc = (double *)malloc((size_t)(nd + 2) * 8);
        int ok = scratch_alloc(&s, nd + 1) && uloc && vloc;
        if (!ok) {
#pragma omp atomic write
            err = -2;
        }
#pragma omp barrier
        if (!err) {
#pragma omp for schedule(static)
            for (int64_t k = k0; k <= k1; ++k) {
                int64_t c = k - k0;
                double *uu = U ? U + c * ldu : uloc;
                double *vv = V ? V + c * ldv : vloc;
                halfarrow_k(nd, nz, D, z, k, tau, &s, &sigma[c], uu, vv, &sind[c], &kb[c], &kz[c],
                            &knu[c], &krho[c], &qout[c], &steps[c], &status[c]);
            }
        }
        if (ok) scratch_free(&s);
        free(uloc); free(vloc);
    }
    return err;
}

/* rootsah's loop `for k=1:n; E[k],Qout[k] = eigen(A,zD,alphaD,k,tau); end` (arrowhead7.jl:116-118; serial in the reference) */
int ah_oracle_rootsah_eigvals(int64_t N, const double *D, const double *z, const double *zlo, double a, double alo,
                              const double *tau, int64_t k0, int64_t k1, int nthreads, double *lambda, int64_t *sind,
                              double *kb, double *kz, double *knu, double *krho, int64_t *qout, int32_t *steps,
                              int32_t *status) {
    if (N < 1 || k0 < 1 || k1 > N + 1 || k1 < k0) return -1;
    int err = 0;
    if (nthreads < 1) nthreads = ah_oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        scratch_t s;
        int ok = scratch_alloc(&s, N + 1);
        if (!ok) {
#pragma omp atomic write
            err = -2;
        }
#pragma omp barrier
        if (!err) {
#pragma omp for schedule(static)
            for (int64_t k = k0; k <= k1; ++k) {
                int64_t c = k - k0;
                rootsah_k(N, D, z, zlo, a, alo, k, tau, &s, &lambda[c], &sind[c], &kb[c], &kz[c], &knu[c], &krho[c],
                          &qout[c], &steps[c], &status[c]);
            }
        }
        if (ok) scratch_free(&s);
    }
    return err;
}
