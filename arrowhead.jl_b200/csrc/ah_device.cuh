// ah_device.cuh -- device-side building blocks shared by the kernels of libarrowhead_b200.so
//
//  * Dekker double-double arithmetic with the exact operation order of the reference's
//    DoubleDouble.jl (:93-96 renorm, :112-116 add2, :129-133 sub, :137-149 mul12/mul2, :152-157 div2);
//    the two-product uses one FMA, which returns the same (exact) error term as Dekker's split-based
//    mul12 whenever the Veltkamp split does not overflow.
//  * deterministic block-wide all-reductions (fixed shuffle tree, fixed cross-warp order).
//  * the secular-term evaluators used by the bisection sweeps.
//
// The translation unit is compiled with -fmad=false: every FMA in the binary is one that is written
// as fma()/__fma_rn() here, so the non-bisection arithmetic rounds exactly like the reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ah {

constexpr double kEps = 2.220446049250313e-16;

// ------------------------------------------------------------------------------------------------
// double-double
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd dd_from(double x) { return dd{x, 0.0}; }
__device__ __forceinline__ double dd_to_double(dd x) { return __dadd_rn(x.hi, x.lo); }
__device__ __forceinline__ dd dd_renorm(double u, double v) {
    double w = __dadd_rn(u, v);
    return dd{w, __dadd_rn(__dadd_rn(u, -w), v)};
}
__device__ __forceinline__ dd dd_neg(dd x) { return dd{-x.hi, -x.lo}; }
__device__ __forceinline__ dd dd_add(dd x, dd y) {
    double r = __dadd_rn(x.hi, y.hi);
    double s;
    if (fabs(x.hi) > fabs(y.hi))
        s = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(x.hi, -r), y.hi), y.lo), x.lo);
    else
        s = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(y.hi, -r), x.hi), x.lo), y.lo);
    return dd_renorm(r, s);
}
__device__ __forceinline__ dd dd_sub(dd x, dd y) {
    double r = __dadd_rn(x.hi, -y.hi);
    double s;
    if (fabs(x.hi) > fabs(y.hi))
        s = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(x.hi, -r), -y.hi), -y.lo), x.lo);
    else
        s = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(-y.hi, -r), x.hi), x.lo), -y.lo);
    return dd_renorm(r, s);
}
__device__ __forceinline__ dd dd_mul12(double x, double y) {
    double z = __dmul_rn(x, y);
    return dd{z, __fma_rn(x, y, -z)};
}
__device__ __forceinline__ dd dd_mul(dd x, dd y) {
    dd c = dd_mul12(x.hi, y.hi);
    double cc = __dadd_rn(__dadd_rn(__dmul_rn(x.hi, y.lo), __dmul_rn(x.lo, y.hi)), c.lo);
    return dd_renorm(c.hi, cc);
}
__device__ __forceinline__ dd dd_div(dd x, dd y) {
    double c = __ddiv_rn(x.hi, y.hi);
    dd u = dd_mul12(c, y.hi);
    double num = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(x.hi, -u.hi), -u.lo), x.lo), -__dmul_rn(c, y.lo));
    return dd_renorm(c, __ddiv_rn(num, y.hi));
}
__device__ __forceinline__ dd dd_abs(dd x) { return x.hi > 0 ? x : dd_neg(x); }

// ------------------------------------------------------------------------------------------------
// reductions.  Every thread gets the result; the combination order is fixed by (lane, warp) only.
struct OpSum {
    __device__ __forceinline__ double operator()(double a, double b) const { return __dadd_rn(a, b); }
};
struct OpMax {
    __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); }
};
struct OpMin {
    __device__ __forceinline__ double operator()(double a, double b) const { return fmin(a, b); }
};

template <typename Op>
__device__ __forceinline__ double warp_allreduce(double v, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double p = __shfl_xor_sync(0xffffffffu, v, o);
        // lower lane's value first, so both partners compute the identical expression
        v = (threadIdx.x & o) ? op(p, v) : op(v, p);
    }
    return v;
}
__device__ __forceinline__ dd warp_allreduce_dd(dd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd p;
        p.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        p.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        v = (threadIdx.x & o) ? dd_add(p, v) : dd_add(v, p);
    }
    return v;
}

// scratch: at least 2*(blockDim.x/32) doubles of shared memory
template <typename Op>
__device__ __forceinline__ double block_allreduce(double v, Op op, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_allreduce(v, op);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = scratch[0];
    for (int w = 1; w < nw; ++w) r = op(r, scratch[w]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ dd block_allreduce_dd(dd v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_allreduce_dd(v);
    if (lane == 0) {
        scratch[2 * warp] = v.hi;
        scratch[2 * warp + 1] = v.lo;
    }
    __syncthreads();
    dd r{scratch[0], scratch[1]};
    for (int w = 1; w < nw; ++w) r = dd_add(r, dd{scratch[2 * w], scratch[2 * w + 1]});
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------------------
// reciprocal seeds and the secular term
//
// rcp.approx.ftz.f64 is MUFU.RCP64H: it looks at the upper 32 bits of the operand only (relative
// error <= ~2^-20).  One cubic correction r1 = r0*(1 + e + e^2), e = 1 - t*r0, brings that to
// ~2^-60 before rounding, i.e. r1 is 1/t to about one ulp -- 3 DFMA instead of the 7 FP64-pipe
// instructions + slow-path test that nvcc emits for an IEEE division.
__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double rcp_fast(double t) {
    double r0 = rcp_seed(t);
    double e = __fma_rn(-t, r0, 1.0);
    double p = __fma_rn(e, e, e);
    return __fma_rn(r0, p, r0);
}

// w / d without nvcc's IEEE-division sequence (whose slow-path test is a branch per quotient and keeps the
// compiler from interleaving independent quotients): q0 = w * (1/d to ~1 ulp), one residual correction
// q = q0 + (w - d*q0)/d  (Markstein).  q is w/d to half an ulp + 2^-100: the correctly rounded quotient.
// Valid while d, 1/d and q stay in the normal range; `ok` is cleared otherwise (the caller divides exactly then).
__device__ __forceinline__ double div_fast(double w, double d, bool &ok) {
    const double y = rcp_fast(d);
    const double q0 = __dmul_rn(w, y);
    const double r = __fma_rn(-d, q0, w);
    const double ad = fabs(d), aq = fabs(q0);
    ok = ok && (ad > 1.0e-290) && (ad < 1.0e290) && (aq > 1.0e-290) && (aq < 1.0e290);
    return __fma_rn(r, y, q0);
}

// One term of the inverse-arrow / inverse-DPR1 secular sum, written on the ORIGINAL data:
//   z'_l^2 / (D'_l - m)  with  D'_l = 1/d,  z'_l^2 = w2 * wz^2 / d^2
//   = wz^2 * w2 / (d * (1 - m*d))
// (wz^2 is applied once to the reduced sum).  d = pole difference, m = bisection point, w2 = w_l^2.
// 1 - m*d is formed with one FMA, i.e. with a single rounding of the exact value.
__device__ __forceinline__ double secular_term_acc(double acc, double d, double m, double w2) {
    double e = __fma_rn(-m, d, 1.0);
    double t = __dmul_rn(d, e);
    return __fma_rn(w2, rcp_fast(t), acc);
}

// First term of a tile sum: the fast kernels add the terms of one tile (4 poles per thread) among themselves first and the
// tile sums to the running sum afterwards -- 4 + (tiles per sweep) roundings per accumulator instead of 4 x tiles, which is
// what makes the error bound of the bisection certificates (ah_pipe.cuh) three times tighter at n = 32768.
__device__ __forceinline__ double secular_term_first(double d, double m, double w2) {
    double e = __fma_rn(-m, d, 1.0);
    double t = __dmul_rn(d, e);
    return __dmul_rn(w2, rcp_fast(t));
}

// The same term with IEEE divisions only: w2/(d*(1-m*d)) is +-Inf when the bisection point sits exactly on an inverse
// pole (as z'^2/(D'-m) is in the reference) and 0 when the denominator overflows, where the reciprocal-seed version
// above produces NaN (0*Inf inside the correction).  Used to redo a sweep whose sum came out NaN.
__device__ __forceinline__ double secular_term_exact(double d, double m, double w2) {
    if (w2 == 0.0) return 0.0;
    double e = __fma_rn(-m, d, 1.0);
    return __ddiv_rn(w2, __dmul_rn(d, e));
}

}  // namespace ah
