// On-chip full-VGP evaluation for small N (N <= GPC_SMALL_NMAX; the reference's first configuration: fission yeast bulk time
// course, N = 18 / 36, full VGP - GPcounts_Module.py:685-686): ONE small CTA (GPC_THREADS = 64) per gene-model with every
// N x N factor in shared memory, so that 8-9 gene-models share an SM.  This regime is latency-bound (15.6 kFLOP per
// evaluation at N = 18, thousands of sequential evaluations per gene - SURVEY.md section 6): what counts is the length of
// the dependent chains and the number of CTAs in flight, not tensor throughput, so there is no DMMA, no cp.async ring and
// no blocked factorisation here.
//
// Math: SURVEY.md appendix B.3 (reverse pass of gpflow VGP.elbo, whitened):
//   L = chol(K + eps I), C = L Lq, fmean = L q_mu, fvar = rowsum(C o C), quadrature -> (g_m, g_v)
//   Cbar = 2 diag(g_v) C,  Lbar = tril(Cbar Lq^T + g_m q_mu^T),  grad Lq = tril(L^T Cbar) - (Lq - diag(1/Lq_ii)),
//   grad q_mu = L^T g_m - q_mu,  Kbar = L^-T Phi(L^T Lbar) L^-1 (through the explicit inverse of L),
//   grad var = <Kbar, dK/dvar>, grad ls = <Kbar, dK/dls>   (Gram entries recomputed, vectorised exp).
// Kernels: RBF, Constant, BranchKernel.  Likelihoods: NB (+ scale), ZINB, Poisson.
#pragma once
#include "model.cuh"
#include "fastmath.cuh"

#define GPC_SMALL_NMAX 48
#define GPC_SMALL_LPP 4                          // lanes per data point in the quadrature
#define GPC_SMALL_PPP (GPC_THREADS / GPC_SMALL_LPP)   // points per pass

// Shared-memory arena of the small evaluator: this header, then five N x ld matrices (ld = N | 1).
struct SmallHdr {
    KernParams kp;
    LikParams lp;
    Hyper h;
    double red[80];
    double fmean[GPC_SMALL_NMAX], fvar[GPC_SMALL_NMAX], gm[GPC_SMALL_NMAX], gv[GPC_SMALL_NMAX], qmu[GPC_SMALL_NMAX];
    double dinv[GPC_SMALL_NMAX];                 // 1 / L_ii
    double yv[GPC_SMALL_NMAX], sv[GPC_SMALL_NMAX];   // counts and scale factors of the gene
    double Xs[2 * GPC_SMALL_NMAX];               // inputs (d <= 2)
    int ok;
};

__host__ __device__ constexpr size_t small_hdr_bytes() { return (sizeof(SmallHdr) + 127) & ~(size_t)127; }
__host__ __device__ inline size_t small_arena_bytes(int N) {
    const size_t ld = (size_t)(N | 1);
    return small_hdr_bytes() + 5 * (((size_t)N * ld * sizeof(double) + 127) & ~(size_t)127);
}

// e / n for 0 <= e < n^2, 2 <= n <= 64 (or n == 1, e == 0), with inv = ceil(2^20 / n)
__device__ __forceinline__ int small_div(int e, unsigned inv) { return (int)(((unsigned)e * inv) >> 20); }

// Quadrature of one data point over the lane's nodes; generic (ZINB) version of lik_nodes_nb with the reference's psi
// arithmetic (see lik.cuh).  Returns partial sums over nodes i0, i0 + ISTEP, ...
struct SmallPart { double ve, gm, gz, da, dk; bool poison; };

template <int ISTEP>
__device__ __forceinline__ SmallPart small_nodes_nb(const LikParams& lp, double fm, double sd, double y, double s, int i0) {
    constexpr int NK = GPC_NGH / ISTEP;
    SmallPart o{0.0, 0.0, 0.0, 0.0, 0.0, false};
    const double alpha = lp.alpha, k = lp.k, ia2 = k * k;
    double F[NK], m[NK], t[NK], lt[NK], z[NK], w[NK];
#pragma unroll
    for (int q = 0; q < NK; ++q) {
        z[q] = c_gh_z[i0 + q * ISTEP];
        w[q] = c_gh_w[i0 + q * ISTEP];
        F[q] = fma(sd, z[q], fm);
    }
    gpc_exp_v<NK>(F, m);
#pragma unroll
    for (int q = 0; q < NK; ++q) { m[q] *= s; t[q] = fma(m[q], alpha, 1.0); }
    gpc_log_v<NK>(t, lt);
#pragma unroll
    for (int q = 0; q < NK; ++q) {
        const double it = gpc_rcp(t[q]);
        double l = fma(y, (F[q] + lp.log_alpha) - lt[q], -k * lt[q]);
        double dl = (y - m[q]) * it;
        const double dal = fma(lt[q], ia2, y * k) - (y + k) * (m[q] * it);
        nb_corner_cases(m[q], y, l, dl);
        o.ve = fma(w[q], l, o.ve); o.gm = fma(w[q], dl, o.gm); o.gz = fma(w[q] * z[q], dl, o.gz); o.da = fma(w[q], dal, o.da);
    }
    return o;
}

// (out of line: the NB path is the hot one, and this kernel lives or dies by its instruction-cache footprint)
template <int ISTEP>
__device__ __noinline__ SmallPart small_nodes_zinb(const LikParams& lp, double fm, double sd, double y, int i0) {
    SmallPart o{0.0, 0.0, 0.0, 0.0, 0.0, false};
    const double alpha = lp.alpha, k = lp.k, ia = k, ia2 = k * k, km = lp.km;
    for (int i = i0; i < GPC_NGH; i += ISTEP) {
        const double z = c_gh_z[i], w = c_gh_w[i];
        const double F = fm + sd * z;
        const double m = exp(F);
        const double t = 1.0 + m * alpha;
        const double lt = log(t);
        const double kmm = km + m;
        const double q = m / kmm;
        const double psi = 1.0 - q;                 // NegativeBinomialLikelihood.py:69, cancellation kept (lik.cuh)
        const double omp = 1.0 - psi;
        o.poison = o.poison || (psi == 0.0);
        if (y == 0.0) {
            const double a = log(psi);
            const double b = log(omp) - lt * ia;
            const double mx = fmax(a, b);
            const double ea = exp(a - mx), eb = exp(b - mx);
            const double se = ea + eb;
            const double l = mx + log(se);
            const double ra = ea / se, rb = eb / se;
            const double dq = q * (km / kmm);
            const double dqk = m / (kmm * kmm);
            const double dl = -ra * dq / psi + rb * (dq / omp - m / t);
            o.ve = fma(w, l, o.ve); o.gm = fma(w, dl, o.gm); o.gz = fma(w * z, dl, o.gz);
            o.dk = fma(w, ra * dqk / psi - rb * dqk / omp, o.dk);
            o.da = fma(w, rb * (lt * ia2 - m * ia / t), o.da);
        } else {
            const double l = log(omp) + y * (F + lp.log_alpha - lt) - k * lt;
            const double dl = q * (km / kmm) / omp + (y - m) / t;
            o.ve = fma(w, l, o.ve); o.gm = fma(w, dl, o.gm); o.gz = fma(w * z, dl, o.gz);
            o.dk = fma(w, -(m / (kmm * kmm)) / omp, o.dk);
            o.da = fma(w, y * ia + lt * ia2 - (y + k) * m / t, o.da);
        }
    }
    return o;
}

// Kernel value from a squared distance through the branch-free exp (RBF / BRANCH; CONST handled by the caller).
__device__ __forceinline__ double small_kval(const KernParams& kp, double r2, bool cross) {
    return (cross ? kp.cross_scale : kp.var) * gpc_exp(-0.5 * r2 * kp.inv_ls2);
}

// ELBO + gradient of one full-VGP gene-model with N <= GPC_SMALL_NMAX.  `arena` = shared memory of small_arena_bytes(N).
__device__ __noinline__ int eval_vgp_small(const ModelDev& md, const double* y, const double* scale, const double* theta,
                                           double* grad, double* f_out, unsigned char* arena) {
    const int N = md.N, tid = threadIdx.x, ld = N | 1;
    const unsigned ninv = ((1u << 20) + (unsigned)N - 1u) / (unsigned)N;
    SmallHdr* hd = reinterpret_cast<SmallHdr*>(arena);
    const size_t msz = (((size_t)N * ld * sizeof(double) + 127) & ~(size_t)127) / sizeof(double);
    double* Lm = reinterpret_cast<double*>(arena + small_hdr_bytes());   // K + shift I -> L
    double* Lq = Lm + msz;                                               // tril(q_sqrt)          (later T = Linv^T P)
    double* Cm = Lq + msz;                                               // C = L Lq -> Cbar       (later P)
    double* Gm = Cm + msz;                                               // Lbar
    double* Li = Gm + msz;                                               // L^-1
    if (tid == 0) {
        hd->h = hyper_from_theta(md, theta);
        kern_setup(hd->kp, md, hd->h.var, hd->h.ls);
        lik_setup<true>(hd->lp, md.lik, hd->h.alpha, hd->h.km);
        hd->ok = 1;
    }
    const double* q_mu = theta + GPC_NHYP;
    const double* lq_packed = theta + GPC_NHYP + N;
    // everything the evaluation reads from global memory is staged once, with independent loads (every later access is
    // shared memory: at this size a dependent L2 round trip costs as much as a whole phase of the evaluation)
    for (int i = tid; i < N; i += GPC_THREADS) {
        hd->qmu[i] = q_mu[i];
        hd->yv[i] = y[i];
        hd->sv[i] = scale ? scale[i] : 1.0;
    }
    for (int i = tid; i < N * md.d; i += GPC_THREADS) hd->Xs[i] = md.X[i];
    for (int e0 = tid; e0 < N * N; e0 += 4 * GPC_THREADS) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * GPC_THREADS;
            const int i = e < N * N ? small_div(e, ninv) : 0, j = e - i * N;
            v[u] = (e < N * N && j <= i) ? lq_packed[tri_index(i, j)] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * GPC_THREADS;
            if (e < N * N) { const int i = small_div(e, ninv), j = e - i * N; Lq[i * ld + j] = v[u]; }
        }
    }
    __syncthreads();
    const KernParams& kp = hd->kp;
    const LikParams& lp = hd->lp;
    const double shift = md.jitter + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    const bool need_kg = (md.train_mask & 3) != 0;
    const int dd = md.d;
    // ---- Gram (lower part; the factorisation reads only it) and q_sqrt ------------------------------------------------
    for (int e = tid; e < N * N; e += GPC_THREADS) {
        const int i = small_div(e, ninv), j = e - i * N;
        double kv = 0.0;
        if (j <= i) {
            if (md.kernel == GPC_CONST) kv = kp.var;
            else {
                bool cross;
                const double r2 = kern_r2(kp, hd->Xs + i * dd, hd->Xs + j * dd, cross);
                kv = small_kval(kp, r2, cross);
            }
            if (i == j) kv += shift;
        }
        Lm[i * ld + j] = kv;
    }
    __syncthreads();
    // ---- L = chol(K): left-looking, thread i owns row i; every thread recomputes the pivot from row j (one barrier per
    //      column); four interleaved partial sums per dot product
    {
        const int i = tid;
        for (int j = 0; j < N; ++j) {
            // (scalar accumulators: an array indexed by k & 3 would live in local memory)
            double pj0 = 0.0, pj1 = 0.0, pj2 = 0.0, pj3 = 0.0, pi0 = 0.0, pi1 = 0.0, pi2 = 0.0, pi3 = 0.0;
            const bool mine = i < N && i > j;
            const double* rj = Lm + j * ld;
            const double* ri = Lm + (mine ? i : j) * ld;
            int k = 0;
            for (; k + 4 <= j; k += 4) {
                const double a0 = rj[k], a1 = rj[k + 1], a2 = rj[k + 2], a3 = rj[k + 3];
                pj0 = fma(a0, a0, pj0); pj1 = fma(a1, a1, pj1); pj2 = fma(a2, a2, pj2); pj3 = fma(a3, a3, pj3);
                pi0 = fma(ri[k], a0, pi0); pi1 = fma(ri[k + 1], a1, pi1); pi2 = fma(ri[k + 2], a2, pi2); pi3 = fma(ri[k + 3], a3, pi3);
            }
            for (; k < j; ++k) { const double a0 = rj[k]; pj0 = fma(a0, a0, pj0); pi0 = fma(ri[k], a0, pi0); }
            const double pj[4] = {pj0, pj1, pj2, pj3}, pi[4] = {pi0, pi1, pi2, pi3};
            const double djj = Lm[j * ld + j] - ((pj[0] + pj[1]) + (pj[2] + pj[3]));
            if (!(djj > 0.0)) { if (tid == 0) hd->ok = 0; break; }      // uniform: every thread sees the same pivot
            const double rinv = gpc_rsqrt(djj);
            double sq = djj * rinv;
            sq = fma(fma(-sq, sq, djj), 0.5 * rinv, sq);
            const double lij = mine ? (Lm[i * ld + j] - ((pi[0] + pi[1]) + (pi[2] + pi[3]))) * rinv : 0.0;
            __syncthreads();                                             // all pivots read before a_jj is replaced
            if (mine) Lm[i * ld + j] = lij;
            if (i == j) { Lm[j * ld + j] = sq; hd->dinv[j] = rinv; }
            __syncthreads();
        }
    }
    __syncthreads();
    if (!hd->ok) return GPC_ST_CHOL;
    // ---- C = L Lq (lower) -----------------------------------------------------------------------------------------------
    for (int e = tid; e < N * N; e += GPC_THREADS) {
        const int i = small_div(e, ninv), j = e - i * N;
        if (j <= i) {
            double s0 = 0.0, s1 = 0.0;
            int k = j;
            for (; k + 1 <= i; k += 2) {
                s0 = fma(Lm[i * ld + k], Lq[k * ld + j], s0);
                s1 = fma(Lm[i * ld + k + 1], Lq[(k + 1) * ld + j], s1);
            }
            if (k <= i) s0 = fma(Lm[i * ld + k], Lq[k * ld + j], s0);
            Cm[i * ld + j] = s0 + s1;
        }
    }
    __syncthreads();
    // ---- fmean = L q_mu, fvar = rowsum(C o C): GPC_SMALL_LPP lanes per row ----------------------------------------------
    double sum_ve = 0.0, sum_da = 0.0, sum_dk = 0.0;
    {
        const int part = tid % GPC_SMALL_LPP;
        for (int n0 = 0; n0 < N; n0 += GPC_SMALL_PPP) {
            const int n = n0 + tid / GPC_SMALL_LPP;
            const bool v = n < N;
            double s1 = 0.0, s2 = 0.0;
            if (v) {
                for (int j = part; j <= n; j += GPC_SMALL_LPP) {
                    s1 = fma(Lm[n * ld + j], hd->qmu[j], s1);
                    const double c = Cm[n * ld + j];
                    s2 = fma(c, c, s2);
                }
            }
#pragma unroll
            for (int o = 1; o < GPC_SMALL_LPP; o <<= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            // ---- quadrature of point n on its GPC_SMALL_LPP lanes --------------------------------------------------------
            const double fm = s1, fv = s2;
            const double yy = v ? hd->yv[n] : 0.0;
            const double sc = v ? hd->sv[n] : 1.0;
            double ve = 0.0, gm = 0.0, gz = 0.0, da = 0.0, dk = 0.0;
            bool poison = false;
            if (md.lik == GPC_POISSON) {
                if (v && part == 0) {
                    const double e = gpc_exp(fm + 0.5 * fv);
                    ve = yy * fm - e - lgamma(yy + 1.0);
                    gm = yy - e;
                    gz = -0.5 * e;                               // holds g_v directly for Poisson
                }
            } else {
                const double sd = sqrt(fv);
                if (md.lik == GPC_NB) {
                    const SmallPart o = small_nodes_nb<GPC_SMALL_LPP>(lp, fm, sd, yy, sc, part);
                    ve = o.ve; gm = o.gm; gz = o.gz; da = o.da;
                } else {
                    const SmallPart o = small_nodes_zinb<GPC_SMALL_LPP>(lp, fm, sd, yy, part);
                    ve = o.ve; gm = o.gm; gz = o.gz; da = o.da; dk = o.dk; poison = o.poison;
                }
                if (v && part == 0 && (md.lik == GPC_NB || yy != 0.0)) {       // node-independent terms, once per point
                    const double xa[2] = {lp.k + yy, yy + 1.0};
                    double lg[2], dg[2];
                    gpc_lgamma_digamma_v<2>(xa, lg, dg);
                    ve += (lg[0] - lg[1]) - lp.lgamma_k;        // the reference's order: at large k it absorbs lgamma(y + 1)
                    da += -(lp.k * lp.k) * (dg[0] - lp.digamma_k);
                    if (md.lik == GPC_NB && scale) ve += yy * gpc_log(sc);      // log m = F + log s
                }
            }
            if (poison) { gm = NAN; gz = NAN; dk = NAN; }
            if (v) { sum_ve += ve; sum_da += da; sum_dk += dk; }
#pragma unroll
            for (int o = 1; o < GPC_SMALL_LPP; o <<= 1) {
                gm += __shfl_xor_sync(0xffffffffu, gm, o);
                gz += __shfl_xor_sync(0xffffffffu, gz, o);
            }
            if (v && part == 0) {
                hd->gm[n] = gm;
                hd->gv[n] = (md.lik == GPC_POISSON) ? gz : gz / (2.0 * sqrt(fv));
            }
        }
    }
    __syncthreads();
    // ---- Cbar = 2 diag(gv) C ---------------------------------------------------------------------------------------------
    for (int e = tid; e < N * N; e += GPC_THREADS) {
        const int i = small_div(e, ninv), j = e - i * N;
        if (j <= i) Cm[i * ld + j] *= 2.0 * hd->gv[i];
    }
    __syncthreads();
    // ---- Lbar = tril(Cbar Lq^T + gm q_mu^T) -> Gm;  grad Lq = tril(L^T Cbar) - (Lq - diag(1/Lq_ii));  KL sums ----------------
    double klq = 0.0, kll = 0.0;
    double* glq = grad + GPC_NHYP + N;
    for (int e = tid; e < N * N; e += GPC_THREADS) {
        const int i = small_div(e, ninv), j = e - i * N;
        if (j <= i) {
            double s0 = 0.0, s1 = 0.0;
            for (int k = 0; k <= j; ++k) s0 = fma(Cm[i * ld + k], Lq[j * ld + k], s0);        // Cbar[i][k] Lq[j][k], k <= j
            for (int k = i; k < N; ++k) s1 = fma(Lm[k * ld + i], Cm[k * ld + j], s1);         // L[k][i] Cbar[k][j], k >= i
            Gm[i * ld + j] = fma(hd->gm[i], hd->qmu[j], s0);
            const double l = Lq[i * ld + j];
            glq[tri_index(i, j)] = s1 - (i == j ? l - 1.0 / l : l);
            klq = fma(l, l, klq);
            if (i == j) kll += log(l * l);
        }
    }
    // grad q_mu = L^T gm - q_mu
    double klm = 0.0;
    for (int j = tid; j < N; j += GPC_THREADS) {
        double s = 0.0;
        for (int i = j; i < N; ++i) s = fma(Lm[i * ld + j], hd->gm[i], s);
        grad[GPC_NHYP + j] = s - hd->qmu[j];
        klm = fma(hd->qmu[j], hd->qmu[j], klm);
    }
    double acc6[6] = {sum_ve, sum_da, sum_dk, klm, klq, kll};
    cta_sum_n<6>(acc6, hd->red);
    const double kl = 0.5 * (acc6[3] - (double)N + acc6[4] - acc6[5]);
    double gvar = 0.0, gls = 0.0;
    if (need_kg) {
        __syncthreads();
        // ---- P = Phi(tril(L^T Lbar)) -> Cm;  Linv (column c by thread c, forward substitution) -> Li ------------------------
        for (int e = tid; e < N * N; e += GPC_THREADS) {
            const int i = small_div(e, ninv), j = e - i * N;
            if (j <= i) {
                double s = 0.0;
                for (int k = i; k < N; ++k) s = fma(Lm[k * ld + i], Gm[k * ld + j], s);
                Cm[i * ld + j] = (i == j) ? 0.5 * s : s;
            }
        }
        for (int c = tid; c < N; c += GPC_THREADS) {
            for (int i = 0; i < c; ++i) Li[i * ld + c] = 0.0;
            for (int i = c; i < N; ++i) {
                double s = (i == c) ? 1.0 : 0.0;
                for (int k = c; k < i; ++k) s = fma(-Lm[i * ld + k], Li[k * ld + c], s);
                Li[i * ld + c] = s * hd->dinv[i];
            }
        }
        __syncthreads();
        // ---- T = Linv^T P -> Lq storage:  T[i][j] = sum_{k >= max(i, j)} Linv[k][i] P[k][j] -------------------------------------
        for (int e = tid; e < N * N; e += GPC_THREADS) {
            const int i = small_div(e, ninv), j = e - i * N;
            double s = 0.0;
            for (int k = max(i, j); k < N; ++k) s = fma(Li[k * ld + i], Cm[k * ld + j], s);
            Lq[i * ld + j] = s;
        }
        __syncthreads();
        // ---- Kbar = T Linv, contracted on the fly with dK/dvar, dK/dls (Gram entries recomputed) --------------------------------
        double a2[2] = {0.0, 0.0};
        for (int e = tid; e < N * N; e += GPC_THREADS) {
            const int i = small_div(e, ninv), j = e - i * N;
            double s = 0.0;
            for (int k = j; k < N; ++k) s = fma(Lq[i * ld + k], Li[k * ld + j], s);
            if (md.kernel == GPC_CONST) a2[0] += s;
            else {
                bool cross;
                const double r2 = kern_r2(kp, hd->Xs + i * dd, hd->Xs + j * dd, cross);
                const double k0 = small_kval(kp, r2, cross);
                kern_grad_acc(kp, s, k0, r2, cross, a2[0], a2[1]);
            }
        }
        cta_sum_n<2>(a2, hd->red);
        gvar = a2[0];
        gls = a2[1];
    }
    store_hyper_grad(md, hd->h, gvar, gls, acc6[1], acc6[2], grad);
    *f_out = acc6[0] - kl;
    __syncthreads();
    return GPC_ST_OK;
}
