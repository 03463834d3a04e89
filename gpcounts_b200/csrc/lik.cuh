// Likelihood terms of the variational expectations: 20-point Gauss-Hermite quadrature of the NB / ZINB
// log-densities (GPcounts/NegativeBinomialLikelihood.py:37-48, 63-77), the Poisson closed form
// (gpflow.likelihoods.Poisson, GPcounts_Module.py:622-623), with the analytic derivatives
// d/d(fmean), d/d(fvar), d/d(alpha), d/d(km) that the reference obtains from TensorFlow autodiff.
#pragma once
#include <math.h>
#include "../../include/gpcounts_b200.h"
#include "fastmath.cuh"

#define GPC_NGH 20

// numpy.polynomial.hermite.hermgauss(20): nodes * sqrt(2), weights / sqrt(pi)
__constant__ double c_gh_z[GPC_NGH];
__constant__ double c_gh_w[GPC_NGH];

struct LikParams {
    int lik;
    double alpha, km;       // constrained values
    double k;               // 1 / alpha
    double log_alpha;
    double lgamma_k;        // lgamma(k)
    double digamma_k;       // psi(k)
};

struct LikPoint {
    double ve;              // E_q[log p(y|f)]
    double gm;              // d ve / d fmean
    double gv;              // d ve / d fvar
    double dalpha;          // d ve / d alpha
    double dkm;             // d ve / d km
};

// psi(x), x > 0: upward recurrence to x >= 10, then the asymptotic series (error < 1e-15).
__device__ __forceinline__ double digamma_pos(double x) {
    double r = 0.0;
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    const double xi = 1.0 / x, x2 = xi * xi;
    double s = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 - x2 * (1.0 / 132.0
               - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0)))))));
    return r + log(x) - 0.5 * xi - s;
}

// FAST: lgamma(k) / digamma(k) from the branch-free routine of fastmath.cuh.  An evaluator must take lgamma(k + y) and
// lgamma(k) from the SAME implementation: at alpha -> 0 (k = 1/alpha ~ 1e39 on wild line-search trial points) k + y == k
// in FP64 and the reference's lgamma(k + y) - lgamma(k) is exactly 0, whereas two implementations differ by
// ~1e-16 * k log k ~ 1e25 and hand the line search a garbage objective.
template <bool FAST = false>
__device__ __forceinline__ void lik_setup(LikParams& lp, int lik, double alpha, double km) {
    lp.lik = lik;
    lp.alpha = alpha;
    lp.km = km;
    lp.k = 1.0 / alpha;
    lp.log_alpha = log(alpha);
    if (FAST) {
        const double xa[1] = {lp.k};
        double lg[1], dg[1];
        gpc_lgamma_digamma_v<1>(xa, lg, dg);
        lp.lgamma_k = lg[0];
        lp.digamma_k = dg[0];
    } else {
        lp.lgamma_k = lgamma(lp.k);
        lp.digamma_k = digamma_pos(lp.k);
    }
}

// IEEE corner cases of the reference's NB log-density  Y log(m / (m + k)) - k log(1 + m alpha)
// (NegativeBinomialLikelihood.py:37-48) that the fused form  y (log m + log alpha - lt) - k lt  does not reproduce:
//   m = +inf (exp overflow):  log(inf / inf) = NaN                                   -> l = NaN (the gradients are NaN already)
//   m = 0    (exp underflow): Y log(0) = -inf for Y > 0, 0 * -inf = NaN for Y = 0;   TensorFlow's d/dF = (..) * m = inf * 0 = NaN
// They occur on wild line-search trial points and decide how SciPy's line search goes on (NaN objective: the step is
// extended until 20 evaluations are spent, the BFGS memory is reset and the fit continues; +inf objective with a NaN
// slope: retreat to the current iterate and stop), so they are reproduced rather than smoothed over.
__device__ __forceinline__ void nb_corner_cases(double m, double y, double& l, double& dl) {
    const bool m_inf = m > 1.7976931348623157e308, m_zero = m == 0.0;
    l = m_inf ? NAN : l;
    l = m_zero ? (y == 0.0 ? NAN : -INFINITY) : l;
    dl = m_zero ? NAN : dl;
}

// y: count, s: scale factor (1 if none), lgy1 = lgamma(y + 1)
__device__ LikPoint lik_point(const LikParams& lp, double fm, double fv, double y, double s) {
    LikPoint o;
    o.dalpha = 0.0;
    o.dkm = 0.0;
    if (lp.lik == GPC_POISSON) {
        const double e = exp(fm + 0.5 * fv);
        o.ve = y * fm - e - lgamma(y + 1.0);
        o.gm = y - e;
        o.gv = -0.5 * e;
        return o;
    }
    const double sd = sqrt(fv);
    const double alpha = lp.alpha, k = lp.k;
    const double ia = k, ia2 = k * k;
    double ve = 0.0, gm = 0.0, gz = 0.0, da = 0.0, dk = 0.0;
    if (lp.lik == GPC_NB) {
        const double logs = log(s);
#pragma unroll 4
        for (int i = 0; i < GPC_NGH; ++i) {
            const double z = c_gh_z[i], w = c_gh_w[i];
            const double logm = fm + sd * z + logs;
            const double m = exp(logm);
            const double t = 1.0 + m * alpha;
            const double lt = log(t);
            double l = y * (logm + lp.log_alpha - lt) - k * lt;
            double dl = (y - m) / t;
            const double dal = y * ia + lt * ia2 - (y + k) * m / t;
            nb_corner_cases(m, y, l, dl);
            ve = fma(w, l, ve);
            gm = fma(w, dl, gm);
            gz = fma(w * z, dl, gz);
            da = fma(w, dal, da);
        }
        const double dpsi = digamma_pos(k + y) - lp.digamma_k;
        ve += lgamma(k + y) - lgamma(y + 1.0) - lp.lgamma_k;
        da += -ia2 * dpsi;
    } else {  // GPC_ZINB (no scale in the reference)
        // The dropout probability is formed exactly as the reference forms it, psi = 1 - m/(km+m)
        // (NegativeBinomialLikelihood.py:69), including its cancellation when km << m.  When psi rounds to
        // exactly 0 the reference's autodiff gradient is NaN for every parameter upstream of psi (the
        // 0 * (1/psi) of d log(psi), for zero and non-zero counts alike because tf.where still back-propagates
        // through the unselected branch); SciPy's line search then retreats to the current iterate and
        // stops there with "success".  That behaviour decides where reference fits end, so it is kept.
        const double km = lp.km;
        bool poison = false;
        if (y == 0.0) {
#pragma unroll 2
            for (int i = 0; i < GPC_NGH; ++i) {
                const double z = c_gh_z[i], w = c_gh_w[i];
                const double F = fm + sd * z;
                const double m = exp(F);
                const double t = 1.0 + m * alpha;
                const double lt = log(t);
                const double kmm = km + m;
                const double q = m / kmm;
                const double psi = 1.0 - q;
                const double omp = 1.0 - psi;
                poison = poison || (psi == 0.0);
                const double a = log(psi);
                const double b = log(omp) - lt * ia;
                const double mx = fmax(a, b);
                const double ea = exp(a - mx), eb = exp(b - mx);
                const double se = ea + eb;
                const double l = mx + log(se);
                const double ra = ea / se, rb = eb / se;
                const double dq = q * (km / kmm);                  // dq/dF
                const double dqk = m / (kmm * kmm);                // -dq/dkm
                const double dl = -ra * dq / psi + rb * (dq / omp - m / t);
                ve = fma(w, l, ve);
                gm = fma(w, dl, gm);
                gz = fma(w * z, dl, gz);
                dk = fma(w, ra * dqk / psi - rb * dqk / omp, dk);
                da = fma(w, rb * (lt * ia2 - m * ia / t), da);
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < GPC_NGH; ++i) {
                const double z = c_gh_z[i], w = c_gh_w[i];
                const double F = fm + sd * z;
                const double m = exp(F);
                const double t = 1.0 + m * alpha;
                const double lt = log(t);
                const double kmm = km + m;
                const double q = m / kmm;
                const double psi = 1.0 - q;
                const double omp = 1.0 - psi;
                poison = poison || (psi == 0.0);
                const double l = log(omp) + y * (F + lp.log_alpha - lt) - k * lt;
                const double dl = q * (km / kmm) / omp + (y - m) / t;
                ve = fma(w, l, ve);
                gm = fma(w, dl, gm);
                gz = fma(w * z, dl, gz);
                dk = fma(w, -(m / (kmm * kmm)) / omp, dk);
                da = fma(w, y * ia + lt * ia2 - (y + k) * m / t, da);
            }
            const double dpsi = digamma_pos(k + y) - lp.digamma_k;
            ve += lgamma(k + y) - lgamma(y + 1.0) - lp.lgamma_k;
            da += -ia2 * dpsi;
        }
        if (poison) { gm = NAN; gz = NAN; dk = NAN; }
    }
    o.ve = ve;
    o.gm = gm;
    o.gv = gz / (2.0 * sd);
    o.dalpha = da;
    o.dkm = dk;
    return o;
}
