// Device L-BFGS-B (unbounded case) - one CTA minimises the negative ELBO of one gene-model.
//
// Replaces `gpflow.optimizers.Scipy().minimize(..., options=dict(maxiter=5000))`
// (GPcounts/GPcounts_Module.py:692-697) = SciPy L-BFGS-B v3.0 with SciPy defaults
// (m = 10, ftol = 2.22e-9, gtol = 1e-5, maxls = 20, maxfun = 15000).  The reference's results depend on
// where this optimiser stops, so the iteration is restated step for step
// (CPU twin: oracle/lbfgsb.py, itself checked against the installed SciPy):
//   * empty memory: generalised Cauchy point z = x - g/theta, first step min(1/|d|, 1e10), later steps 1
//   * otherwise the quasi-Newton direction through the compact representation (formk / subsm with all
//     variables free), d = z - x
//   * More'-Thuente dcsrch / dcstep with ftol 1e-3, gtol 0.9, xtol 0.1, at most maxls trial points
//   * convergence: max|g| <= gtol, or (fold - f) <= ftol * max(|fold|, |f|, 1) after an accepted step
//   * pair accepted iff dr > epsmch * ddum; theta = rr / dr; memory refreshed when a factorisation or
//     the line search fails; abnormal termination when that happens with an empty memory
// All scalar logic is executed redundantly by every thread on CTA-uniform values.
#pragma once
#include "model.cuh"
#include "fastmath.cuh"
#ifdef GPC_SMALL_ONLY
#include "vgp_small.cuh"
#else
#include "svgp_fused.cuh"
#endif

#define GPC_MMAX 10
#define GPC_EPSMCH 2.220446049250313e-16

enum { TASK_PGTOL = 0, TASK_FTOL = 1, TASK_LIMIT = 2, TASK_ABNORMAL = 3, TASK_CHOL = 4 };

struct OptSmem {
    double SS[GPC_MMAX][GPC_MMAX];     // s_i . s_j (symmetric, logical order = oldest first)
    double SY[GPC_MMAX][GPC_MMAX];     // lower incl. diagonal: s_i . y_j (i >= j), diagonal = dr   (matupd)
    double YY[GPC_MMAX][GPC_MMAX];     // y_i . y_j
    double RZ[GPC_MMAX][GPC_MMAX];     // upper incl. diagonal: s_i . y_j (i <= j) by explicit dots  (formk)
    double WN[2 * GPC_MMAX][2 * GPC_MMAX];   // upper-triangular LEL^T factor of the middle matrix
    double wv[2 * GPC_MMAX];
    double dots[GPC_MMAX][6];          // per memory vector: S.d, Y.d, S.r, Y.r, S.g, Y.g
#ifndef GPC_SMALL_ONLY
    double dotw[GPC_THREADS / 32][GPC_MMAX / 2][6];   // per-warp partial sums of one pass (summed over the warps in a fixed order)
#endif
    double T[GPC_MMAX][GPC_MMAX];      // formt scratch
    double rd[2 * GPC_MMAX];           // 1 / diag of the LEL^T factor in WN (formk_free), used by subsm_solve
    int info;
    int info_dir;                      // outcome of the direction solve (formk / subsm) done at the end of an iteration
    double red[64];                    // CTA reduction scratch of the optimiser
};

struct DcsrchState {
    bool brackt;
    int stage;
    double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax, stp;
};

__device__ __forceinline__ void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy,
                                       double& stp, double fp, double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf, stpc, stpq, theta, s, gamma, p, q, r;
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) < fabs(stpq - stx)) stpf = stpc; else stpf = stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc; else stpf = stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc; else stpf = stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc; else stpf = stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

#define LS_FTOL 1e-3
#define LS_GTOL 0.9
#define LS_XTOL 0.1
#define LS_STPMIN 0.0
#define LS_STPMAX 1e10

// returns false on "ERROR" (initial slope not negative / step out of range)
__device__ __forceinline__ bool dcsrch_start(DcsrchState& s, double f, double g, double stp) {
    if (stp < LS_STPMIN || stp > LS_STPMAX || !(g < 0.0)) return false;
    s.brackt = false;
    s.stage = 1;
    s.finit = f; s.ginit = g;
    s.gtest = LS_FTOL * g;
    s.width = LS_STPMAX - LS_STPMIN;
    s.width1 = s.width / 0.5;
    s.stx = 0.0; s.fx = f; s.gx = g;
    s.sty = 0.0; s.fy = f; s.gy = g;
    s.stmin = 0.0;
    s.stmax = stp + 4.0 * stp;
    s.stp = stp;
    return true;
}

// returns 0 = evaluate again at s.stp, 1 = converged, 2 = warning (also terminates the search)
__device__ __forceinline__ int dcsrch_step(DcsrchState& s, double f, double g) {
    double stp = s.stp;
    const double ftest = s.finit + stp * s.gtest;
    if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
    int task = 0;
    if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) task = 2;
    if (s.brackt && s.stmax - s.stmin <= LS_XTOL * s.stmax) task = 2;
    if (stp == LS_STPMAX && f <= ftest && g <= s.gtest) task = 2;
    if (stp == LS_STPMIN && (f > ftest || g >= s.gtest)) task = 2;
    if (f <= ftest && fabs(g) <= LS_GTOL * (-s.ginit)) task = 1;
    if (task != 0) return task;
    if (s.stage == 1 && f <= s.fx && f > ftest) {
        const double gt = s.gtest;
        const double fm = f - stp * gt;
        double fxm = s.fx - s.stx * gt, fym = s.fy - s.sty * gt;
        const double gm = g - gt;
        double gxm = s.gx - gt, gym = s.gy - gt;
        dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * gt;
        s.fy = fym + s.sty * gt;
        s.gx = gxm + gt;
        s.gy = gym + gt;
    } else {
        dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
    }
    if (s.brackt) {
        if (fabs(s.sty - s.stx) >= 0.66 * s.width1) stp = s.stx + 0.5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
    }
    if (s.brackt) {
        s.stmin = fmin(s.stx, s.sty);
        s.stmax = fmax(s.stx, s.sty);
    } else {
        s.stmin = stp + 1.1 * (stp - s.stx);
        s.stmax = stp + 4.0 * (stp - s.stx);
    }
    stp = fmax(stp, LS_STPMIN);
    stp = fmin(stp, LS_STPMAX);
    if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) || (s.brackt && s.stmax - s.stmin <= LS_XTOL * s.stmax))
        stp = s.stx;
    s.stp = stp;
    return 0;
}

// The limited-memory matrices are at most 20 x 20; warp 0 works on them cooperatively in shared memory (the rest
// of the CTA waits at the next barrier).  All routines must be called by every lane of one warp.

// Upper Cholesky A = R^T R of the leading n x n block, n <= GPC_MMAX (LINPACK dpofa semantics: 0 or the failing pivot
// index); rdiag, if given, receives 1 / R[k][k].  Right-looking; per pivot every lane fetches the operands of its trailing
// entries BEFORE the pivot's rsqrt chain (they do not depend on it), so a pivot step costs one shared-memory round trip
// plus the rsqrt chain instead of three round trips, a square root and a division.
__device__ int warp_chol_upper(double* A, int lda, int n, double* rdiag) {
    const int lane = threadIdx.x & 31;
    constexpr int NQ = ((GPC_MMAX - 1) * (GPC_MMAX - 1) + 31) / 32;     // trailing entries per lane (square enumeration)
    for (int k = 0; k < n; ++k) {
        __syncwarp();
        const double akk = A[k * lda + k];
        const int rem = n - k - 1;
        const unsigned inv = rem > 0 ? (65536u + rem - 1) / rem : 0u;   // exact idx / rem for idx < rem^2, rem <= 20
        double aki[NQ], akj[NQ], aij[NQ];
        int off[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int idx = lane + 32 * q;
            const int d = (int)(((unsigned)idx * inv) >> 16);
            const int i = k + 1 + d, j = k + 1 + (idx - d * rem);
            const bool on = idx < rem * rem && j >= i;
            off[q] = on ? i * lda + j : -1;
            aki[q] = on ? A[k * lda + i] : 0.0;
            akj[q] = on ? A[k * lda + j] : 0.0;
            aij[q] = on ? A[i * lda + j] : 0.0;
        }
        const double akl = lane < rem ? A[k * lda + k + 1 + lane] : 0.0;
        if (!(akk > 0.0)) return k + 1;
        const double rs = gpc_rsqrt(akk);
        double r = akk * rs;
        r = fma(fma(-r, r, akk), 0.5 * rs, r);                          // sqrt(akk)
        __syncwarp();                                                   // every lane holds its operands: row k may be rewritten
        if (lane == 0) { A[k * lda + k] = r; if (rdiag) rdiag[k] = rs; }
        if (lane < rem) A[k * lda + k + 1 + lane] = akl * rs;
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (off[q] >= 0) A[off[q]] = fma(-(aki[q] * rs), akj[q] * rs, aij[q]);
    }
    __syncwarp();
    return 0;
}

// formk for the all-free case: builds and factors WN (upper triangle). 0 on success.  os->rd receives 1 / diag.
__device__ int formk_free(OptSmem* os, int col, double theta) {
    const int L2 = 2 * GPC_MMAX, lane = threadIdx.x & 31;
    double* WN = &os->WN[0][0];
    const double ith = 1.0 / theta;
    for (int idx = lane; idx < col * col; idx += 32) {
        const int iy = idx / col, jy = idx % col;
        if (jy <= iy) {
            WN[jy * L2 + iy] = os->YY[iy][jy] * ith + (jy == iy ? os->SY[iy][iy] : 0.0);     // Y'Y / theta + D
            WN[(col + jy) * L2 + col + iy] = 0.0;                                            // theta * S'AA'S = 0
        }
        WN[jy * L2 + col + iy] = (jy >= iy) ? os->RZ[iy][jy] : 0.0;                           // R_z' (and -L_a' = 0)
    }
    __syncwarp();
    int info = warp_chol_upper(WN, L2, col, os->rd);
    if (info != 0) return -1;
    // columns of block (1,2): solve R1^T x = b  (dtrsl job 11), one lane per column, the column in registers
    if (lane < col) {
        const int js = col + lane;
        double x[GPC_MMAX];
#pragma unroll
        for (int i = 0; i < GPC_MMAX; ++i) {
            if (i < col) {
                double t = WN[i * L2 + js];
#pragma unroll
                for (int k = 0; k < i; ++k) t = fma(-WN[k * L2 + i], x[k], t);
                x[i] = t * os->rd[i];
                WN[i * L2 + js] = x[i];
            }
        }
    }
    __syncwarp();
    for (int idx = lane; idx < col * col; idx += 32) {
        const int a = idx / col, b = idx % col;
        if (b >= a) {
            double t = 0.0;
            for (int k = 0; k < col; ++k) t += WN[k * L2 + col + a] * WN[k * L2 + col + b];
            WN[(col + a) * L2 + col + b] += t;
        }
    }
    __syncwarp();
    info = warp_chol_upper(WN + col * L2 + col, L2, col, os->rd + col);
    if (info != 0) return -2;
    return 0;
}

// formt: T = theta*SS + L D^-1 L^T must be positive definite. 0 on success.
__device__ int formt_check(OptSmem* os, int col, double theta) {
    const int lane = threadIdx.x & 31;
    double* T = &os->T[0][0];
    for (int idx = lane; idx < col * col; idx += 32) {
        const int i = idx / col, j = idx % col;
        if (j >= i) {
            double ddum = 0.0;
            for (int k = 0; k < i; ++k) ddum += os->SY[i][k] * os->SY[j][k] / os->SY[k][k];
            T[i * GPC_MMAX + j] = ddum + theta * os->SS[i][j];
        }
    }
    __syncwarp();
    return warp_chol_upper(T, GPC_MMAX, col, nullptr);
}

// subsm middle solve: wv <- K^-1 wv through the LEL^T factor (column-oriented substitutions).  Lane k keeps wv[k] in a
// register (2 col <= 20 < 32), the pivot value travels by shuffle and the matrix row of the next step is fetched while the
// current one is applied: a step is shuffle + multiply-add instead of two warp barriers, a division and two round trips
// through shared memory.  os->rd holds 1 / diag of the factor (formk_free).
__device__ void subsm_solve(OptSmem* os, int col) {
    const int L2 = 2 * GPC_MMAX, c2 = 2 * col, lane = threadIdx.x & 31;
    const double* WN = &os->WN[0][0];
    const bool in = lane < c2;
    double w = in ? os->wv[lane] : 0.0;
    const double rdl = in ? os->rd[lane] : 0.0;
    double un = in ? WN[lane] : 0.0;                      // U[0][lane]
    for (int i = 0; i < c2; ++i) {                       // U^T x = b
        const double u = un;
        if (i + 1 < c2) un = in ? WN[(i + 1) * L2 + lane] : 0.0;
        const double xi = __shfl_sync(0xffffffffu, w * rdl, i);
        if (lane == i) w = xi;
        else if (lane > i) w = fma(-u, xi, w);
    }
    if (lane < col) w = -w;
    un = in ? WN[lane * L2 + c2 - 1] : 0.0;               // U[lane][c2 - 1]
    for (int i = c2 - 1; i >= 0; --i) {                  // U x = b
        const double u = un;
        if (i > 0) un = in ? WN[lane * L2 + i - 1] : 0.0;
        const double xi = __shfl_sync(0xffffffffu, w * rdl, i);
        if (lane == i) w = xi;
        else if (lane < i) w = fma(-u, xi, w);
    }
    if (in) os->wv[lane] = w;
    __syncwarp();
}

struct OptResult {
    int task;          // TASK_*
    int nit, nfev;
    double f;          // final loss (= -ELBO)
    double gnorm;      // max |g|
    long long eval_cycles;   // SM cycles spent inside model_eval (profiling aid, GPC_S_EVAL_CYCLES)
};

__device__ __forceinline__ double cta_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < GPC_THREADS / 32; ++w) t = fmax(t, red[w]);
    __syncthreads();
    return t;
}

// Two sums and a maximum over the CTA with one barrier pair; identical results in every thread.
__device__ __forceinline__ void cta_sum2_max(double& s0, double& s1, double& mx, double* red) {
    const int NW = GPC_THREADS / 32;
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp] = s0; red[NW + warp] = s1; red[2 * NW + warp] = mx; }
    __syncthreads();
    double a = 0.0, b = 0.0, c = red[2 * NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { a += red[w]; b += red[NW + w]; c = fmax(c, red[2 * NW + w]); }
    __syncthreads();
    s0 = a; s1 = b; mx = c;
}

// Optional evaluation trace (debug hook, see gpc_dbg_set_trace): 8 doubles per evaluation.
struct OptTrace {
    double* buf;
    int cap;
    int n;
};

__device__ __forceinline__ void trace_eval(OptTrace* tr, double f, double stp, double gd, const double* x,
                                           const double* g, int P, double* red) {
    if (!tr || !tr->buf) return;
    double gmax = 0.0;
    int nan = 0;
    for (int i = threadIdx.x; i < P; i += GPC_THREADS) { const double v = fabs(g[i]); if (v != v) nan = 1; else gmax = fmax(gmax, v); }
    gmax = cta_max(gmax, red);
    const double nn = cta_sum((double)nan, red);
    if (threadIdx.x == 0 && tr->n < tr->cap) {
        double* r = tr->buf + (size_t)tr->n * 8;
        r[0] = f; r[1] = stp; r[2] = gd; r[3] = gmax; r[4] = x[0]; r[5] = x[1]; r[6] = x[2] ; r[7] = nn > 0 ? -x[3] - 1e300 : x[3];
    }
    tr->n += 1;
}

// Evaluate loss = -ELBO and its gradient at ws.x; returns status. g is written negated into ws.g.
__device__ __forceinline__ int eval_loss(const ModelDev& md, const double* Z, const double* y, const double* scale,
                                         const SlotWs& ws, CtaSmem* sm, double& loss, long long& cycles) {
    double elbo;
    const long long t0 = clock64();
    const int st = model_eval(md, Z, y, scale, ws.x, ws.g, &elbo, ws, sm);
    cycles += clock64() - t0;
    if (st != GPC_ST_OK) return st;
    loss = -elbo;
    return GPC_ST_OK;
}

// Minimise from ws.x (in/out). Memory slots: WS / WY rows of length P, slot (head + j) % m holds pair j.
__device__ OptResult lbfgsb_minimize(const ModelDev& md, const double* Z, const double* y, const double* scale,
                                     const gpc_opt_t& opt, const SlotWs& ws, CtaSmem* sm, OptSmem* os,
                                     OptTrace* tr = nullptr) {
    const int P = md.P, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = opt.m;
    OptResult res;
    res.nit = 0; res.nfev = 0; res.task = TASK_ABNORMAL; res.f = NAN; res.gnorm = NAN; res.eval_cycles = 0;
    double f;
    if (eval_loss(md, Z, y, scale, ws, sm, f, res.eval_cycles) != GPC_ST_OK) { res.task = TASK_CHOL; res.nfev = 1; return res; }
    int nfev = 1, nit = 0;
    trace_eval(tr, f, 0.0, 0.0, ws.x, ws.g, P, os->red);
    // g <- -grad(ELBO)
    double gmax = 0.0;
    for (int i = tid; i < P; i += GPC_THREADS) { const double v = -ws.g[i]; ws.g[i] = v; gmax = fmax(gmax, fabs(v)); }
    double sbgnrm = cta_max(gmax, os->red);
    if (sbgnrm <= opt.gtol) { res.task = TASK_PGTOL; res.nfev = nfev; res.f = f; res.gnorm = sbgnrm; return res; }
    int col = 0, head = 0;
    double theta = 1.0;
    bool updated = false;      // a new pair was stored since the last formk
#if defined(GPC_PHASE_PROF) && !defined(GPC_DENSE_ONLY) && !defined(GPC_SMALL_ONLY)
    long long op_t = clock64(), op_acc[6] = {0, 0, 0, 0, 0, 0};
#define OPROF(k) do { const long long t_ = clock64(); op_acc[k] += t_ - op_t; op_t = t_; } while (0)
#define OPROF_RESET do { op_t = clock64(); } while (0)
#define OPROF_FLUSH do { if (threadIdx.x == 0) { for (int q_ = 0; q_ < 6; ++q_) atomicAdd(&g_opt_prof[q_], (unsigned long long)op_acc[q_]); atomicAdd(&g_opt_prof[7], (unsigned long long)nit); } } while (0)
#else
#define OPROF(k)
#define OPROF_RESET
#define OPROF_FLUSH
#endif
    while (true) {
        OPROF_RESET;
        // ------------------------------------------------------------------ search direction (+ line-search set-up)
        // one pass: d = z - x, dtd, g.d, save (x, g) as (xold, gold); after the first iteration the first trial
        // point is z itself (stp = 1), so it is written in the same pass
        double acc2[2] = {0.0, 0.0};
        const bool first_is_z = nit > 0;
        if (col == 0) {
            for (int i = tid; i < P; i += GPC_THREADS) {
                const double xi = ws.x[i], gi = ws.g[i];
                const double zi = xi - gi / theta;
                const double di = zi - xi;
                ws.z[i] = zi;
                ws.d[i] = di;
                ws.xold[i] = xi;
                ws.gold[i] = gi;
                if (first_is_z) ws.x[i] = zi;
                acc2[0] = fma(di, di, acc2[0]);
                acc2[1] = fma(gi, di, acc2[1]);
            }
        } else {
            // the middle-matrix solve for this direction was done at the end of the previous iteration (warp 1, next
            // to formt on warp 0): every path that reaches this point with col > 0 comes from there
            const int info = os->info_dir;
            if (info != 0) { col = 0; head = 0; theta = 1.0; updated = false; continue; }   // refresh the memory
            updated = false;
            const double ith = 1.0 / theta;
            // eight elements per thread at a time, memory pairs in the outer loop: the 16 loads of a pair are independent
            // (a per-element j-loop serialises one L2 round trip per pair); per element the arithmetic keeps its order
            for (int i0 = tid; i0 < P; i0 += 8 * GPC_THREADS) {
                double xi[8], gi[8], di[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * GPC_THREADS;
                    xi[u] = i < P ? ws.x[i] : 0.0;
                    gi[u] = i < P ? ws.g[i] : 0.0;
                    di[u] = -gi[u];
                }
                for (int j = 0; j < col; ++j) {
                    const size_t slot = (size_t)((head + j) % m) * P;
                    const double cy = os->wv[j] / theta, cs = os->wv[col + j];
                    double vy[8], vs[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * GPC_THREADS;
                        vy[u] = i < P ? ws.WY[slot + i] : 0.0;
                        vs[u] = i < P ? ws.WS[slot + i] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) di[u] += vy[u] * cy + vs[u] * cs;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * GPC_THREADS;
                    if (i < P) {
                        double d = di[u] * ith;
                        const double zi = xi[u] + d;
                        d = zi - xi[u];
                        ws.z[i] = zi;
                        ws.d[i] = d;
                        ws.xold[i] = xi[u];
                        ws.gold[i] = gi[u];
                        if (first_is_z) ws.x[i] = zi;
                        acc2[0] = fma(d, d, acc2[0]);
                        acc2[1] = fma(gi[u], d, acc2[1]);
                    }
                }
            }
        }
        cta_sum_n<2>(acc2, os->red);
        OPROF(0);
        // ------------------------------------------------------------------ line search (lnsrlb)
        const double dtd = acc2[0];
        const double dnorm = sqrt(dtd);
        double gd = acc2[1];
        const double gdold = gd;
        const double fold = f;
        double stp = (nit == 0) ? fmin(1.0 / dnorm, LS_STPMAX) : 1.0;
        bool ls_failed = false;
        double gmax = 0.0, rr = 0.0;
        DcsrchState ls;
        if (!(gd < 0.0) || !dcsrch_start(ls, f, gd, stp)) {
            ls_failed = true;                      // ascent direction in projection (info = -4)
        } else {
            int ifun = 0;
            while (true) {
                stp = ls.stp;
                ++ifun;
                ++nfev;
                if (!(ifun == 1 && first_is_z)) {
                    if (stp == 1.0) {
                        for (int i = tid; i < P; i += GPC_THREADS) ws.x[i] = ws.z[i];
                    } else {
                        for (int i = tid; i < P; i += GPC_THREADS) ws.x[i] = stp * ws.d[i] + ws.xold[i];
                    }
                    __syncthreads();
                }
                OPROF(1);
                if (eval_loss(md, Z, y, scale, ws, sm, f, res.eval_cycles) != GPC_ST_OK) {
                    res.task = TASK_CHOL; res.nit = nit; res.nfev = nfev; return res;
                }
                OPROF_RESET;
                // one pass: g <- -grad, g.d, max|g|, |g - gold|^2
                double a3[2] = {0.0, 0.0};
                double mx = 0.0;
                // four elements per thread at a time: the twelve loads are independent (one L2 round trip instead of four);
                // the per-element arithmetic and the order of the partial sums are unchanged
                for (int i0 = tid; i0 < P; i0 += 4 * GPC_THREADS) {
                    double gv[4], dv[4], ov[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * GPC_THREADS;
                        gv[u] = i < P ? ws.g[i] : 0.0;
                        dv[u] = i < P ? ws.d[i] : 0.0;
                        ov[u] = i < P ? ws.gold[i] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * GPC_THREADS;
                        if (i < P) {
                            const double v = -gv[u];
                            ws.g[i] = v;
                            a3[0] = fma(v, dv[u], a3[0]);
                            const double r = v - ov[u];
                            a3[1] = fma(r, r, a3[1]);
                            mx = fmax(mx, fabs(v));
                        }
                    }
                }
                cta_sum2_max(a3[0], a3[1], mx, os->red);
                gd = a3[0]; rr = a3[1]; gmax = mx;
                trace_eval(tr, f, stp, gd, ws.x, ws.g, P, os->red);
                const int t = dcsrch_step(ls, f, gd);
                OPROF(2);
                if (t != 0) break;
                if (ifun >= opt.maxls) { ls_failed = true; break; }
            }
        }
        if (ls_failed) {
            for (int i = tid; i < P; i += GPC_THREADS) { ws.x[i] = ws.xold[i]; ws.g[i] = ws.gold[i]; }
            f = fold;
            __syncthreads();
            if (col == 0) { res.task = TASK_ABNORMAL; ++nit; break; }
            col = 0; head = 0; theta = 1.0; updated = false;
            continue;
        }
        stp = ls.stp;
        ++nit;
        // ------------------------------------------------------------------ limits and convergence tests
        sbgnrm = gmax;
        if (nit >= opt.maxiter || nfev > opt.maxfun) { res.task = TASK_LIMIT; break; }
        if (sbgnrm <= opt.gtol) { res.task = TASK_PGTOL; break; }
        if ((fold - f) <= opt.ftol * fmax(fmax(fabs(fold), fabs(f)), 1.0)) { res.task = TASK_FTOL; break; }
        // ------------------------------------------------------------------ BFGS update (matupd)
        double dr, ddum;
        if (stp == 1.0) {
            dr = gd - gdold;
            ddum = -gdold;
        } else {
            dr = (gd - gdold) * stp;
            ddum = -gdold * stp;
        }
        const bool skip = dr <= GPC_EPSMCH * ddum;
        if (!skip) {
            if (col == m) {       // drop the oldest pair: shift the small matrices
                __syncthreads();
                {
                    // read (i + 1, j + 1), barrier, write (i, j): (m - 1)^2 <= 81 elements, at most two per thread
                    constexpr int NP = (81 + GPC_THREADS - 1) / GPC_THREADS;
                    double v0[NP], v1[NP], v2[NP], v3[NP];
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const int e = tid + q * GPC_THREADS, i = e / GPC_MMAX, j = e % GPC_MMAX;
                        const bool mine = i < m - 1 && j < m - 1;
                        v0[q] = v1[q] = v2[q] = v3[q] = 0.0;
                        if (mine) { v0[q] = os->SS[i + 1][j + 1]; v1[q] = os->SY[i + 1][j + 1]; v2[q] = os->YY[i + 1][j + 1]; v3[q] = os->RZ[i + 1][j + 1]; }
                    }
                    __syncthreads();
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const int e = tid + q * GPC_THREADS, i = e / GPC_MMAX, j = e % GPC_MMAX;
                        if (i < m - 1 && j < m - 1) { os->SS[i][j] = v0[q]; os->SY[i][j] = v1[q]; os->YY[i][j] = v2[q]; os->RZ[i][j] = v3[q]; }
                    }
                }
                head = (head + 1) % m;
                col = m - 1;
            }
            // one pass: s = stp * d, y = g - gold, stored in the memory and kept in (d, gold) for the dot products
            const size_t slot = (size_t)((head + col) % m) * P;
            for (int i0 = tid; i0 < P; i0 += 4 * GPC_THREADS) {
                double dv[4], gv[4], ov[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * GPC_THREADS;
                    dv[u] = i < P ? ws.d[i] : 0.0;
                    gv[u] = i < P ? ws.g[i] : 0.0;
                    ov[u] = i < P ? ws.gold[i] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * GPC_THREADS;
                    if (i < P) {
                        const double si = (stp == 1.0) ? dv[u] : dv[u] * stp;
                        const double yi = gv[u] - ov[u];
                        ws.d[i] = si;
                        ws.gold[i] = yi;
                        ws.WS[slot + i] = si;
                        ws.WY[slot + i] = yi;
                    }
                }
            }
            ++col;
            theta = rr / dr;
            updated = true;
        }
        __syncthreads();
        OPROF(3);
        // dots of every memory pair with the new s (= d), y (= gold) and the new gradient g
#ifdef GPC_SMALL_ONLY
        // small-N instantiation (P <= 1228, two warps, vectors in shared memory): one warp per pair
        for (int j = warp; j < col; j += GPC_THREADS / 32) {
            const size_t slot = (size_t)((head + j) % m) * P;
            double a[6] = {0, 0, 0, 0, 0, 0};
            for (int i0 = lane; i0 < P; i0 += 4 * 32) {           // twenty independent loads per step
                double sv[4], yv[4], dv[4], rv[4], gv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32;
                    const bool in = i < P;
                    sv[u] = in ? ws.WS[slot + i] : 0.0; yv[u] = in ? ws.WY[slot + i] : 0.0;
                    dv[u] = in ? ws.d[i] : 0.0; rv[u] = in ? ws.gold[i] : 0.0; gv[u] = in ? ws.g[i] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (i0 + u * 32 < P) {
                        a[0] = fma(sv[u], dv[u], a[0]); a[1] = fma(yv[u], dv[u], a[1]);
                        a[2] = fma(sv[u], rv[u], a[2]); a[3] = fma(yv[u], rv[u], a[3]);
                        a[4] = fma(sv[u], gv[u], a[4]); a[5] = fma(yv[u], gv[u], a[5]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) a[q] = warp_sum(a[q]);
            if (lane == 0)
                for (int q = 0; q < 6; ++q) os->dots[j][q] = a[q];
        }
#else
        // Element-parallel: thread t owns the elements t, t + 256, ...; the memory pairs are processed in two passes of five, so
        // the three shared vectors are read twice instead of once per pair (the vectors live in the slot behind L2, this phase
        // is bound by L2 traffic: 0.45 MB instead of 0.86 MB per iteration at P = 2148) and a pass is three L2 round trips of 39
        // independent loads.  Partial sums: warp shuffle, then over the warps in a fixed order.
        {
            constexpr int HP = GPC_MMAX / 2, NW = GPC_THREADS / 32;
            for (int j0 = 0; j0 < col; j0 += HP) {
                const int nj = min(HP, col - j0);
                size_t slot[HP];
#pragma unroll
                for (int jj = 0; jj < HP; ++jj) slot[jj] = (size_t)((head + j0 + (jj < nj ? jj : 0)) % m) * P;
                double a[HP][6];
#pragma unroll
                for (int jj = 0; jj < HP; ++jj)
#pragma unroll
                    for (int q = 0; q < 6; ++q) a[jj][q] = 0.0;
                for (int i0 = tid; i0 < P; i0 += 3 * GPC_THREADS) {
                    double dv[3], rv[3], gv[3], sv[HP][3], yv[HP][3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int i = i0 + u * GPC_THREADS;
                        const bool in = i < P;
                        dv[u] = in ? ws.d[i] : 0.0; rv[u] = in ? ws.gold[i] : 0.0; gv[u] = in ? ws.g[i] : 0.0;
#pragma unroll
                        for (int jj = 0; jj < HP; ++jj) {
                            sv[jj][u] = in ? ws.WS[slot[jj] + i] : 0.0;
                            yv[jj][u] = in ? ws.WY[slot[jj] + i] : 0.0;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
#pragma unroll
                        for (int jj = 0; jj < HP; ++jj) {
                            a[jj][0] = fma(sv[jj][u], dv[u], a[jj][0]); a[jj][1] = fma(yv[jj][u], dv[u], a[jj][1]);
                            a[jj][2] = fma(sv[jj][u], rv[u], a[jj][2]); a[jj][3] = fma(yv[jj][u], rv[u], a[jj][3]);
                            a[jj][4] = fma(sv[jj][u], gv[u], a[jj][4]); a[jj][5] = fma(yv[jj][u], gv[u], a[jj][5]);
                        }
                    }
                }
#pragma unroll
                for (int jj = 0; jj < HP; ++jj)
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const double v = warp_sum(a[jj][q]);
                        if (lane == 0) os->dotw[warp][jj][q] = v;
                    }
                __syncthreads();
                if (tid < 6 * nj) {
                    const int jj = tid / 6, q = tid % 6;
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) v += os->dotw[w][jj][q];
                    os->dots[j0 + jj][q] = v;
                }
                if (j0 + HP < col) __syncthreads();           // dotw is reused by the second pass
            }
        }
#endif
        __syncthreads();
        OPROF(4);
        if (!skip && warp == 0) {
            const int c = col - 1;
            if (lane < col) {
                const int j = lane;
                os->SS[j][c] = os->SS[c][j] = os->dots[j][0];
                os->YY[j][c] = os->YY[c][j] = os->dots[j][3];
                os->SY[c][j] = os->dots[j][1];        // s_new . y_j
                os->RZ[j][c] = os->dots[j][2];        // s_j . y_new (diagonal included)
            }
            __syncwarp();
            if (lane == 0) {
                os->SS[c][c] = (stp == 1.0) ? dtd : stp * stp * dtd;
                os->SY[c][c] = dr;
            }
        }
        __syncthreads();
        // formt (acceptance of the new pair) on warp 0 and, concurrently on warp 1, the direction solve of the next
        // iteration: formk (only when a pair was stored) and subsm with the dot products of the new gradient
        // (A register-resident variant of these three - lane per column, pivots by shuffle, fully unrolled 20-step
        // elimination - was measured in round 2: 13 k SASS instructions that gained 1 % on the 256-thread kernel and cost the
        // 64-thread small-N kernel most of its time in instruction-cache misses (stall_no_inst 44-69 %).  The compact
        // shared-memory loops below are what ships.)
        if (!skip && warp == 0) {
            const int info = formt_check(os, col, theta);
            if (lane == 0) os->info = info;
        }
        if (warp == 1 && col > 0) {
            int info = 0;
            if (updated) info = formk_free(os, col, theta);
            if (info == 0) {
                if (lane < col) {
                    os->wv[lane] = -os->dots[lane][5];                 // Y_j . r, r = -g
                    os->wv[col + lane] = -theta * os->dots[lane][4];   // theta * S_j . r
                }
                __syncwarp();
                subsm_solve(os, col);
            }
            if (lane == 0) os->info_dir = info;
        }
        __syncthreads();
        if (!skip) {
            const int info = os->info;
            __syncthreads();
            if (info != 0) { col = 0; head = 0; theta = 1.0; updated = false; }
        }
        OPROF(5);
    }
    OPROF_FLUSH;
    res.nit = nit;
    res.nfev = nfev;
    res.f = f;
    res.gnorm = sbgnrm;
    return res;
}
