// Branch-free FP64 exp / log / reciprocal for the latency-critical inner loops (Gram tiles, Gauss-Hermite nodes).
//
// CUDA's libdevice exp() / log() / division are accurate to 1 ulp but each call carries its own slow-path branch, so a
// sequence of independent calls compiles to a sequence of basic blocks and their ~20-deep DFMA chains run one after the
// other (measured on the C5 evaluator: 16 Gram entries per thread = 16 serialised chains, the FP64 pipe 3-4x under-used).
// The functions below are straight-line code (selects instead of branches) and come in a "vector" form that evaluates K
// independent arguments stage by stage, which hands the scheduler K independent chains.
//
// Accuracy (checked on the host against libm, tests/test_fastmath.py): exp <= 1 ulp on [-745, 709] with +inf / 0 / NaN
// handled like exp(); log <= 2 ulp for normal positive arguments, +inf -> +inf, NaN -> NaN; reciprocal <= 1 ulp.
#pragma once
#include <math.h>
#include <string.h>
#ifdef __CUDACC__
#define GPC_HD __host__ __device__ __forceinline__
#else
#define GPC_HD static inline
#endif

GPC_HD double gpc_bits_to_double(long long b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(b);
#else
    double d; memcpy(&d, &b, 8); return d;
#endif
}
GPC_HD long long gpc_double_to_bits(double d) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(d);
#else
    long long b; memcpy(&b, &d, 8); return b;
#endif
}

GPC_HD int gpc_lo_int(double d) {
#ifdef __CUDA_ARCH__
    return __double2loint(d);
#else
    return (int)(unsigned)(gpc_double_to_bits(d) & 0xffffffffLL);
#endif
}

// 1 / x for finite x != 0 (x = +-inf -> +-0, NaN -> NaN): hardware seed (MUFU.RCP64H, ~20 bits) + Newton steps.
GPC_HD double gpc_rcp(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#else
    return 1.0 / x;
#endif
}

// exp(x): n = rint(x log2 e), r = x - n ln2 (two-constant Cody-Waite), degree-13 Taylor polynomial on |r| <= ln2 / 2
// (truncation 4e-18 relative), scaled by 2^n in two steps so that overflow gives +inf and underflow degrades gradually.
template <int K>
GPC_HD void gpc_exp_v(const double (&x)[K], double (&out)[K]) {
    const double LOG2E = 1.4426950408889634, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double r[K], p[K], fn[K];
    int ni[K];
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < K; ++k) {
        double xc = x[k];
        if (xc < -1000.0) xc = -1000.0;          // selects: NaN compares false and passes through
        if (xc > 1000.0) xc = 1000.0;
        // round-to-nearest through the 2^52 + 2^51 shift (two adds instead of FRND + F2I, which issue at a quarter rate);
        // the integer n sits in the low word of the shifted value (|n| <= 1443)
        const double sh = fma(xc, LOG2E, 6755399441055744.0);
        ni[k] = gpc_lo_int(sh);
        fn[k] = sh - 6755399441055744.0;
        r[k] = fma(fn[k], -LN2_HI, xc);
        r[k] = fma(fn[k], -LN2_LO, r[k]);
        p[k] = 1.6059043836821613e-10;           // 1/13!
    }
    const double c[12] = {2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.755731922398589e-06,
                          2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
                          4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0};
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int j = 0; j < 12; ++j) {
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int k = 0; k < K; ++k) p[k] = fma(p[k], r[k], c[j]);
    }
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < K; ++k) {
        p[k] = fma(p[k], r[k], 1.0);
        const int n = ni[k];                     // garbage for NaN input; the product below stays NaN through p
        const int n1 = n >> 1, n2 = n - n1;
        const double s1 = gpc_bits_to_double((long long)(n1 + 1023) << 52);
        const double s2 = gpc_bits_to_double((long long)(n2 + 1023) << 52);
        out[k] = (p[k] * s1) * s2;
    }
}

GPC_HD double gpc_exp(double x) {
    double a[1] = {x}, o[1];
    gpc_exp_v<1>(a, o);
    return o[0];
}

// log(t) for t > 0 normal (or +inf / NaN): t = 2^e f with f in [sqrt(1/2), sqrt(2)), u = (f - 1) / (f + 1),
// log f = 2 u (1 + v/3 + v^2/5 + ... + v^11/23), v = u^2 <= 0.0295 (truncation 5e-19 relative);
// log t = e ln2_hi + (log f + e ln2_lo).
template <int K>
GPC_HD void gpc_log_v(const double (&t)[K], double (&out)[K]) {
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double u[K], v[K], q[K], fe[K];
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < K; ++k) {
        const long long b = gpc_double_to_bits(t[k]);
        int e = (int)((b >> 52) & 0x7ff) - 1023;
        double f = gpc_bits_to_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);   // [1, 2)
        const bool big = f > 1.4142135623730951;
        f = big ? 0.5 * f : f;
        e = big ? e + 1 : e;
        fe[k] = (double)e;
        const double sm1 = f - 1.0, den = f + 1.0, rc = gpc_rcp(den);
        const double u0 = sm1 * rc;
        u[k] = fma(fma(-u0, den, sm1), rc, u0);         // one correction step: the leading term 2u carries the result
        v[k] = u[k] * u[k];
        q[k] = 1.0 / 23.0;
    }
    const double c[10] = {1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0,
                          1.0 / 5.0, 1.0 / 3.0};
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int j = 0; j < 10; ++j) {
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int k = 0; k < K; ++k) q[k] = fma(q[k], v[k], c[j]);
    }
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < K; ++k) {
        const double two_u = u[k] + u[k];
        const double lf = fma(two_u * v[k], q[k], two_u);                 // 2u + 2u v q
        const double res = fma(fe[k], LN2_HI, fma(fe[k], LN2_LO, lf));
        out[k] = (fabs(t[k]) < INFINITY) ? res : t[k];                    // +inf -> +inf, NaN -> NaN
    }
}

GPC_HD double gpc_log(double t) {
    double a[1] = {t}, o[1];
    gpc_log_v<1>(a, o);
    return o[0];
}

// lgamma(x) and digamma(x) for x > 0, branch-free: shift x up to X = x + n >= 10 with the recurrences
//   P = x (x+1) ... (x+n-1),  P' = dP/dx   (lgamma(x) = lgamma(X) - log P,  digamma(x) = digamma(X) - P'/P),
// then the Stirling / asymptotic series at X (truncation < 1e-16 for X >= 10).  One log(X), one log(P), two reciprocals
// and ~50 multiply-adds per argument, against several hundred instructions for libdevice lgamma + a digamma loop.
template <int K>
GPC_HD void gpc_lgamma_digamma_v(const double (&x)[K], double (&lg)[K], double (&dg)[K]) {
    double X[K], P[K], Q[K];
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < K; ++k) { X[k] = x[k]; P[k] = 1.0; Q[k] = 0.0; }
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int it = 0; it < 10; ++it) {
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int k = 0; k < K; ++k) {
            const bool up = X[k] < 10.0;
            const double q1 = fma(Q[k], X[k], P[k]), p1 = P[k] * X[k];
            Q[k] = up ? q1 : Q[k];
            P[k] = up ? p1 : P[k];
            X[k] = up ? X[k] + 1.0 : X[k];
        }
    }
    double lX[K], lP[K];
    gpc_log_v<K>(X, lX);
    gpc_log_v<K>(P, lP);
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k = 0; k < K; ++k) {
        const double r = gpc_rcp(X[k]), r2 = r * r;
        // sum B_2m / (2m (2m-1) X^(2m-1)), m = 1..7
        double sl = 1.0 / 156.0;
        sl = fma(sl, r2, -691.0 / 360360.0);
        sl = fma(sl, r2, 1.0 / 1188.0);
        sl = fma(sl, r2, -1.0 / 1680.0);
        sl = fma(sl, r2, 1.0 / 1260.0);
        sl = fma(sl, r2, -1.0 / 360.0);
        sl = fma(sl, r2, 1.0 / 12.0);
        // sum B_2m / (2m X^(2m)), m = 1..7
        double sd = 1.0 / 12.0;
        sd = fma(sd, r2, -691.0 / 32760.0);
        sd = fma(sd, r2, 1.0 / 132.0);
        sd = fma(sd, r2, -1.0 / 240.0);
        sd = fma(sd, r2, 1.0 / 252.0);
        sd = fma(sd, r2, -1.0 / 120.0);
        sd = fma(sd, r2, 1.0 / 12.0);
        const double lgX = fma(X[k] - 0.5, lX[k], -X[k]) + 0.91893853320467274178 + sl * r;
        const double dgX = (lX[k] - 0.5 * r) - sd * r2;
        lg[k] = lgX - lP[k];
        dg[k] = dgX - Q[k] * gpc_rcp(P[k]);
    }
}

#ifdef __CUDACC__
// x^-1/2 without a slow-path branch: hardware seed (MUFU.RSQ64H) + two third-order Newton steps.
__device__ __forceinline__ double gpc_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {                       // third-order steps: 20 -> 60 -> 180 bits
        const double e = fma(-x * y, y, 1.0);
        y = fma(y * e, fma(0.375, e, 0.5), y);
    }
    return y;
}

#endif
