// Host-side accuracy check of gpcounts_b200/csrc/fastmath.cuh against long-double libm (tests/test_fastmath.py).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include "../../gpcounts_b200/csrc/fastmath.cuh"

static double ulp_err(double got, long double ref) {
    if (std::isinf((double)ref) || ref == 0.0L) return (got == (double)ref) ? 0.0 : 1e9;
    const double r = (double)ref;
    const double ulp = std::fabs(std::nextafter(r, INFINITY) - r);
    return (double)(fabsl((long double)got - ref) / ulp);
}

int main() {
    uint64_t s = 88172645463325252ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); };
    double worst_exp = 0, worst_log = 0, worst_exp_neg = 0;
    for (int i = 0; i < 2000000; ++i) {
        const double x = -745.0 + 1454.0 * rnd();
        const double e = ulp_err(gpc_exp(x), expl((long double)x));
        if (x > -708.0 && e > worst_exp) worst_exp = e;
        const double xn = -60.0 * rnd() * rnd();
        const double en = ulp_err(gpc_exp(xn), expl((long double)xn));
        if (en > worst_exp_neg) worst_exp_neg = en;
        const double t = std::exp(-30.0 + 740.0 * rnd());
        const double l = ulp_err(gpc_log(t), logl((long double)t));
        if (l > worst_log) worst_log = l;
        const double t1 = 1.0 + std::exp(-40.0 + 60.0 * rnd());
        const double l1 = ulp_err(gpc_log(t1), logl((long double)t1));
        if (l1 > worst_log) worst_log = l1;
    }
    // lgamma / digamma against long-double references (digamma: asymptotic series after a shift to X >= 60)
    double worst_lg = 0, worst_dg = 0;
    for (int i = 0; i < 400000; ++i) {
        const double x = (i & 1) ? std::exp(-7.0 + 20.0 * rnd()) : 1.0 + std::floor(2000.0 * rnd() * rnd());
        double xa[1] = {x}, lg[1], dg[1];
        gpc_lgamma_digamma_v<1>(xa, lg, dg);
        const long double lref = lgammal((long double)x);
        long double X = x, sh = 0;
        while (X < 60.0L) { sh += 1.0L / X; X += 1.0L; }
        const long double r2 = 1.0L / (X * X);
        const long double dref = logl(X) - 0.5L / X - r2 * (1.0L / 12 - r2 * (1.0L / 120 - r2 * (1.0L / 252 - r2 * (1.0L / 240 - r2 * (1.0L / 132))))) - sh;
        const double el = (double)(fabsl(lg[0] - lref) / fmaxl(1.0L, fabsl(lref)));
        const double ed = (double)(fabsl(dg[0] - dref) / fmaxl(1.0L, fabsl(dref)));
        if (el > worst_lg) worst_lg = el;
        if (ed > worst_dg) worst_dg = ed;
    }
    printf("lgamma_err %.3e digamma_err %.3e\n", worst_lg, worst_dg);
    double v4[4] = {-1.5, 0.0, 2.25, -30.0}, o4[4];
    gpc_exp_v<4>(v4, o4);
    int vec_ok = 1;
    for (int k = 0; k < 4; ++k) vec_ok &= (o4[k] == gpc_exp(v4[k]));
    printf("exp_ulp %.3f exp_nonpos_ulp %.3f log_ulp %.3f vec_ok %d\n", worst_exp, worst_exp_neg, worst_log, vec_ok);
    printf("special exp(800)=%g exp(-800)=%g exp(-720)=%g log(inf)=%g exp(0)=%.17g log(1)=%.17g\n", gpc_exp(800.0), gpc_exp(-800.0),
           gpc_exp(-720.0), gpc_log(INFINITY), gpc_exp(0.0), gpc_log(1.0));
    return 0;
}
