// One ELBO + analytic-gradient evaluation of a gene-model by one CTA (sm_100a, FP64).
//
// Restates, with hand-derived reverse passes, what the reference evaluates through
// GPflow + TensorFlow autodiff on every L-BFGS-B step (GPcounts/GPcounts_Module.py:692-697):
//   VGP   gpflow.models.VGP.elbo                        (call site GPcounts_Module.py:685-686)
//   SVGP  gpflow.models.SVGP.elbo, whitened conditional (call site GPcounts_Module.py:679-680)
//   GPR   gpflow.models.GPR.log_marginal_likelihood     (call site GPcounts_Module.py:667)
//   SGPR  gpflow.models.SGPR.elbo (Titsias bound)       (call site GPcounts_Module.py:661-663)
// with RBF / Constant kernels (:574, :601-604) and the BranchKernel (GPcounts/branchingKernel.py:31-74).
#pragma once
#include "dense.cuh"
#include "lik.cuh"

#define GPC_BRANCH_NOISE 1e-6     // BranchKernel(noise_level=1e-6), branchingKernel.py:23
#define GPC_GAUSS_VAR_LOWER 1e-6  // gpflow Gaussian likelihood variance lower bound

struct ModelDev {
    int kind, kernel, lik, N, M, d, n_off, train_mask, n_zsets, handoff_alpha;
    double xp, jitter;
    const double *X, *Z, *Zgene, *restart;
    int Mv;          // rows of q_mu / q_sqrt
    int P;           // parameter count
};

// Per-slot workspace carved from the caller's buffer.
struct SlotWs {
    double *K0, *L, *dinv, *Lq, *G1, *G2, *G3, *G4, *dinv2;   // square buffers (dinv: nblk x 64 x 64)
    double *Kuf, *A, *B, *S;                      // M x N   (sparse models)
    double *v0, *v1, *v2, *v3;                    // N
    double *x, *g, *d, *xold, *gold, *z, *WS, *WY;   // optimiser vectors (P) and memory (m x P)
    // TMA descriptors (CUtensorMap in kernel-parameter space) of the gene-major count / scale matrices and the row of the
    // gene being evaluated, or nullptr: then the evaluators read y / scale through their pointer arguments
    const void *tmY, *tmS;
    int row;
    // 1 on the first evaluation of a fit attempt (set by the callers, cleared by the evaluator): a model whose kernel
    // hyper-parameters are frozen (branching grid, train_mask & 3 == 0 - GPcounts_Module.py:608-614) builds and factors its
    // Gram matrix then and reuses the factor in ws.L for every further evaluation of the attempt
    int fresh;
    int arena_off;      // small instantiation: byte offset of the evaluator's arena in the dynamic shared memory
};

struct KernParams {
    int kernel, d;
    double var, ls, inv_ls2, inv_ls3, xp;
    double cross_scale;   // var^2 / (var + jitter)              (BranchKernel cross-branch prefactor)
    double cross_dvar;    // (var + 2 jit) / (var (var + jit))   (d log Kcross / d var)
};

__device__ __forceinline__ double softplus_d(double x) { return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

__device__ __forceinline__ void kern_setup(KernParams& kp, const ModelDev& md, double var, double ls) {
    kp.kernel = md.kernel;
    kp.d = md.d;
    kp.var = var;
    kp.ls = ls;
    kp.inv_ls2 = 1.0 / (ls * ls);
    kp.inv_ls3 = kp.inv_ls2 / ls;
    kp.xp = md.xp;
    const double jit = 1e-6;   // default_jitter() inside BranchKernel.K, branchingKernel.py:51-55
    kp.cross_scale = var * var / (var + jit);
    kp.cross_dvar = (var + 2.0 * jit) / (var * (var + jit));
}

// Effective squared distance and branch flag of a pair of inputs.
__device__ __forceinline__ double kern_r2(const KernParams& kp, const double* a, const double* b, bool& cross) {
    cross = false;
    if (kp.kernel == GPC_CONST) return 0.0;
    if (kp.kernel == GPC_BRANCH) {
        if (a[1] == b[1]) { const double t = a[0] - b[0]; return t * t; }
        cross = true;
        const double u = a[0] - kp.xp, v = b[0] - kp.xp;
        return u * u + v * v;
    }
    double r2 = 0.0;
    for (int q = 0; q < kp.d; ++q) { const double t = a[q] - b[q]; r2 = fma(t, t, r2); }
    return r2;
}

__device__ __forceinline__ double kern_val(const KernParams& kp, double r2, bool cross) {
    if (kp.kernel == GPC_CONST) return kp.var;
    const double e = exp(-0.5 * r2 * kp.inv_ls2);
    return (cross ? kp.cross_scale : kp.var) * e;
}

// Accumulate <G, dK/dvar> and <G, dK/dls> contributions of one Gram entry with value k0.
__device__ __forceinline__ void kern_grad_acc(const KernParams& kp, double g, double k0, double r2, bool cross,
                                              double& dvar, double& dls) {
    if (kp.kernel == GPC_CONST) { dvar += g; return; }
    const double gk = g * k0;
    dvar = fma(gk, cross ? kp.cross_dvar : 1.0 / kp.var, dvar);
    dls = fma(gk * r2, kp.inv_ls3, dls);
}

struct Hyper {
    double var, ls, alpha, km;       // constrained values (alpha = noise variance for Gaussian)
    double dvar, dls, dalpha, dkm;   // d(constrained)/d(theta)
};

__device__ __forceinline__ Hyper hyper_from_theta(const ModelDev& md, const double* theta) {
    Hyper h;
    h.var = softplus_d(theta[0]);   h.dvar = sigmoid_d(theta[0]);
    h.ls = softplus_d(theta[1]);    h.dls = sigmoid_d(theta[1]);
    h.alpha = softplus_d(theta[2]); h.dalpha = sigmoid_d(theta[2]);
    h.km = softplus_d(theta[3]);    h.dkm = sigmoid_d(theta[3]);
    if (md.lik == GPC_GAUSS) h.alpha += GPC_GAUSS_VAR_LOWER;
    return h;
}

// Write the four hyper-parameter gradient slots (chain rule through softplus, frozen / unused slots = 0).
__device__ __forceinline__ void store_hyper_grad(const ModelDev& md, const Hyper& h, double gvar, double gls,
                                                 double galpha, double gkm, double* grad) {
    if (threadIdx.x == 0) {
        const bool use_ls = md.kernel != GPC_CONST;
        const bool use_al = md.lik == GPC_NB || md.lik == GPC_ZINB || md.lik == GPC_GAUSS;
        const bool use_km = md.lik == GPC_ZINB;
        grad[0] = (md.train_mask & 1) ? gvar * h.dvar : 0.0;
        grad[1] = (use_ls && (md.train_mask & 2)) ? gls * h.dls : 0.0;
        grad[2] = (use_al && (md.train_mask & 4)) ? galpha * h.dalpha : 0.0;
        grad[3] = (use_km && (md.train_mask & 8)) ? gkm * h.dkm : 0.0;
    }
}

// K0 (n x n, full symmetric, no jitter) and L = K0 + shift*I from inputs Xs (n x d).
__device__ void build_gram_square(const KernParams& kp, int n, const double* Xs, double shift, double* K0, double* L) {
    GPC_FOR_2D(n, n, i, j) {
        bool cross;
        const double r2 = kern_r2(kp, Xs + (size_t)i * kp.d, Xs + (size_t)j * kp.d, cross);
        const double k = kern_val(kp, r2, cross);
        K0[idx] = k;
        L[idx] = (i == j) ? k + shift : k;
    }
    __syncthreads();
}

// <G, dK/dvar>, <G, dK/dls> over the full n x n matrix G against the stored Gram K0.
__device__ void gram_square_grad(const KernParams& kp, int n, const double* Xs, const double* K0, const double* G,
                                 double& dvar, double& dls, CtaSmem* sm) {
    sm = cta_arena();
    double acc[2] = {0.0, 0.0};
    GPC_FOR_2D(n, n, i, j) {
        bool cross;
        const double r2 = kern_r2(kp, Xs + (size_t)i * kp.d, Xs + (size_t)j * kp.d, cross);
        kern_grad_acc(kp, G[idx], K0[idx], r2, cross, acc[0], acc[1]);
    }
    cta_sum_n<2>(acc, sm->red);
    dvar = acc[0];
    dls = acc[1];
}

__device__ __forceinline__ size_t tri_index(int i, int j) { return (size_t)i * (i + 1) / 2 + j; }

// Unpack tril(q_sqrt) (row-major packed) into a full Mv x Mv matrix with a zero upper triangle.
__device__ void unpack_lq(int mv, const double* packed, double* Lq) {
    GPC_FOR_2D(mv, mv, i, j) {
        Lq[idx] = (j <= i) ? packed[tri_index(i, j)] : 0.0;
    }
    __syncthreads();
}

// gauss_kl (whitened): 0.5 (|q_mu|^2 - Mv + |tril(Lq)|_F^2 - sum log Lq_ii^2)
__device__ double kl_white(int mv, const double* q_mu, const double* packed, CtaSmem* sm) {
    sm = cta_arena();
    double acc[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < mv; i += GPC_THREADS) {
        acc[0] = fma(q_mu[i], q_mu[i], acc[0]);
        const double dii = packed[tri_index(i, i)];
        acc[2] += log(dii * dii);
    }
    const size_t nt = (size_t)mv * (mv + 1) / 2;
    for (size_t t = threadIdx.x; t < nt; t += GPC_THREADS) acc[1] = fma(packed[t], packed[t], acc[1]);
    cta_sum_n<3>(acc, sm->red);
    return 0.5 * (acc[0] - (double)mv + acc[1] - acc[2]);
}

// Likelihood pass over the N data points: fills gm / gv and returns sum VE, d/dalpha, d/dkm.
__device__ void lik_pass(const ModelDev& md, const Hyper& h, int N, const double* fmean, const double* fvar,
                         const double* y, const double* scale, double* gm, double* gv, double& ve, double& dalpha,
                         double& dkm, CtaSmem* sm) {
    sm = cta_arena();
    LikParams lp;
    lik_setup(lp, md.lik, h.alpha, h.km);
    double acc[3] = {0.0, 0.0, 0.0};
    for (int n = threadIdx.x; n < N; n += GPC_THREADS) {
        const LikPoint o = lik_point(lp, fmean[n], fvar[n], y[n], scale ? scale[n] : 1.0);
        gm[n] = o.gm;
        gv[n] = o.gv;
        acc[0] += o.ve;
        acc[1] += o.dalpha;
        acc[2] += o.dkm;
    }
    cta_sum_n<3>(acc, sm->red);
    ve = acc[0];
    dalpha = acc[1];
    dkm = acc[2];
}

// ------------------------------------------------------------------------------------------------ VGP
__device__ __noinline__ int eval_vgp(const ModelDev& md, const double* y, const double* scale, const double* theta, double* grad,
                        double* f_out, const SlotWs& ws, CtaSmem* sm) {
    sm = cta_arena();
    const int N = md.N, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Hyper h = hyper_from_theta(md, theta);
    KernParams kp;
    kern_setup(kp, md, h.var, h.ls);
    const double* q_mu = theta + GPC_NHYP;
    const double* lq_packed = theta + GPC_NHYP + N;
    double* fmean = ws.v0; double* fvar = ws.v1; double* gm = ws.v2; double* gv = ws.v3;

    const double shift = md.jitter + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    const bool need_kernel_grad = (md.train_mask & 3) != 0;
    if (need_kernel_grad || ws.fresh) {            // frozen kernel: the factor of the first evaluation stays in ws.L
        build_gram_square(kp, N, md.X, shift, ws.K0, ws.L);
        if (!cta_potrf(N, ws.L, N, ws.dinv, sm)) return GPC_ST_CHOL;
    }
    __syncthreads();
    if (tid == 0) const_cast<SlotWs&>(ws).fresh = 0;
    unpack_lq(N, lq_packed, ws.Lq);
    // C = L * Lq (lower)
    cta_gemm<false, false>(N, N, N, 1.0, ws.L, N, ws.Lq, N, 0.0, ws.G1, N, F_A_LOWER | F_B_LOWER | F_C_LOWER, sm);
    for (int row = warp; row < N; row += GPC_THREADS / 32) {
        double s1 = 0.0, s2 = 0.0;
        for (int j = lane; j <= row; j += 32) {
            s1 = fma(ws.L[(size_t)row * N + j], q_mu[j], s1);
            const double c = ws.G1[(size_t)row * N + j];
            s2 = fma(c, c, s2);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) { fmean[row] = s1; fvar[row] = s2; }
    }
    __syncthreads();
    double ve, dalpha, dkm;
    lik_pass(md, h, N, fmean, fvar, y, scale, gm, gv, ve, dalpha, dkm, sm);
    // Cbar = 2 diag(gv) C
    GPC_FOR_2D(N, N, i, j) {
        if (j <= i) ws.G1[idx] *= 2.0 * gv[i];
    }
    __syncthreads();
    // Lbar = tril(Cbar Lq^T + gm q_mu^T)
    cta_gemm<false, true>(N, N, N, 1.0, ws.G1, N, ws.Lq, N, 0.0, ws.G2, N, F_A_LOWER | F_B_UPPER | F_C_LOWER, sm);
    GPC_FOR_2D(N, N, i, j) {
        if (j <= i) ws.G2[idx] = fma(gm[i], q_mu[j], ws.G2[idx]);
    }
    // grad Lq = tril(L^T Cbar) - (Lq - diag(1/Lq_ii))
    cta_gemm<true, false>(N, N, N, 1.0, ws.L, N, ws.G1, N, 0.0, ws.G3, N, F_A_UPPER | F_B_LOWER | F_C_LOWER, sm);
    double* glq = grad + GPC_NHYP + N;
    GPC_FOR_2D(N, N, i, j) {
        if (j <= i) {
            const double l = ws.Lq[idx];
            glq[tri_index(i, j)] = ws.G3[idx] - (i == j ? l - 1.0 / l : l);
        }
    }
    // grad q_mu = L^T gm - q_mu
    for (int j = tid; j < N; j += GPC_THREADS) {
        double s = 0.0;
        for (int i = j; i < N; ++i) s = fma(ws.L[(size_t)i * N + j], gm[i], s);
        grad[GPC_NHYP + j] = s - q_mu[j];
    }
    __syncthreads();
    const double kl = kl_white(N, q_mu, lq_packed, sm);
    double gvar = 0.0, gls = 0.0;
    if (need_kernel_grad) {
        cta_chol_backward(N, ws.L, N, ws.dinv, ws.G2, N, ws.G3, N, sm);
        gram_square_grad(kp, N, md.X, ws.K0, ws.G2, gvar, gls, sm);
    }
    store_hyper_grad(md, h, gvar, gls, dalpha, dkm, grad);
    *f_out = ve - kl;
    __syncthreads();
    return GPC_ST_OK;
}

// ------------------------------------------------------------------------------------------------ SVGP
__device__ __noinline__ int eval_svgp(const ModelDev& md, const double* Z, const double* y, const double* scale,
                         const double* theta, double* grad, double* f_out, const SlotWs& ws, CtaSmem* sm) {
    sm = cta_arena();
    const int N = md.N, M = md.M, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Hyper h = hyper_from_theta(md, theta);
    KernParams kp;
    kern_setup(kp, md, h.var, h.ls);
    const double* q_mu = theta + GPC_NHYP;
    const double* lq_packed = theta + GPC_NHYP + M;
    double* fmean = ws.v0; double* fvar = ws.v1; double* gm = ws.v2; double* gv = ws.v3;

    const double shift = md.jitter + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    const double kdiag = h.var + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    build_gram_square(kp, M, Z, shift, ws.K0, ws.L);
    // Kuf and A (= copy of Kuf, solved in place)
    GPC_FOR_2D(M, N, i, n) {
        bool cross;
        const double r2 = kern_r2(kp, Z + (size_t)i * kp.d, md.X + (size_t)n * kp.d, cross);
        const double k = kern_val(kp, r2, cross);
        ws.Kuf[idx] = k;
        ws.A[idx] = k;
    }
    __syncthreads();
    if (!cta_potrf(M, ws.L, M, ws.dinv, sm)) return GPC_ST_CHOL;
    unpack_lq(M, lq_packed, ws.Lq);
    cta_trsm_left(M, N, ws.L, M, ws.dinv, ws.A, N, sm);                       // A = Lm^-1 Kuf
    cta_gemm<true, false>(M, N, M, 1.0, ws.Lq, M, ws.A, N, 0.0, ws.B, N, F_A_UPPER, sm);   // B = Lq^T A
    for (int n = tid; n < N; n += GPC_THREADS) {
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int i = 0; i < M; ++i) {
            const double a = ws.A[(size_t)i * N + n], b = ws.B[(size_t)i * N + n];
            s1 = fma(a, q_mu[i], s1);
            s2 = fma(a, a, s2);
            s3 = fma(b, b, s3);
        }
        fmean[n] = s1;
        fvar[n] = (kdiag - s2) + s3;
    }
    __syncthreads();
    double ve, dalpha, dkm;
    lik_pass(md, h, N, fmean, fvar, y, scale, gm, gv, ve, dalpha, dkm, sm);
    // Bbar = 2 B diag(gv) (in place);  S = q_mu gm^T - 2 A diag(gv)
    double gv_sum = 0.0;
    for (int n = tid; n < N; n += GPC_THREADS) gv_sum += gv[n];
    gv_sum = cta_sum(gv_sum, sm->red);
    GPC_FOR_2D(M, N, i, n) {
        const double g2 = 2.0 * gv[n];
        ws.B[idx] *= g2;
        ws.S[idx] = fma(q_mu[i], gm[n], -g2 * ws.A[idx]);
    }
    __syncthreads();
    cta_gemm<false, false>(M, N, M, 1.0, ws.Lq, M, ws.B, N, 1.0, ws.S, N, F_A_LOWER, sm);   // S += Lq Bbar  (= Abar)
    // grad q_mu = A gm - q_mu
    for (int row = warp; row < M; row += GPC_THREADS / 32) {
        double s = 0.0;
        for (int n = lane; n < N; n += 32) s = fma(ws.A[(size_t)row * N + n], gm[n], s);
        s = warp_sum(s);
        if (lane == 0) grad[GPC_NHYP + row] = s - q_mu[row];
    }
    // grad Lq = tril(A Bbar^T) - (Lq - diag(1/Lq_ii))
    cta_gemm<false, true>(M, M, N, 1.0, ws.A, N, ws.B, N, 0.0, ws.G1, M, F_C_LOWER, sm);
    double* glq = grad + GPC_NHYP + M;
    GPC_FOR_2D(M, M, i, j) {
        if (j <= i) {
            const double l = ws.Lq[idx];
            glq[tri_index(i, j)] = ws.G1[idx] - (i == j ? l - 1.0 / l : l);
        }
    }
    __syncthreads();
    const double kl = kl_white(M, q_mu, lq_packed, sm);
    double gvar = 0.0, gls = 0.0;
    const bool need_kernel_grad = (md.train_mask & 3) != 0;
    if (need_kernel_grad) {
        cta_trsm_left_t(M, N, ws.L, M, ws.dinv, ws.S, N, sm);                  // S = Lm^-T Abar  (= Kuf bar)
        // Lm bar = -tril(S A^T)
        cta_gemm<false, true>(M, M, N, -1.0, ws.S, N, ws.A, N, 0.0, ws.G2, M, F_C_LOWER, sm);
        double acc[2] = {0.0, 0.0};
        GPC_FOR_2D(M, N, i, n) {
            bool cross;
            const double r2 = kern_r2(kp, Z + (size_t)i * kp.d, md.X + (size_t)n * kp.d, cross);
            kern_grad_acc(kp, ws.S[idx], ws.Kuf[idx], r2, cross, acc[0], acc[1]);
        }
        cta_sum_n<2>(acc, sm->red);
        cta_chol_backward(M, ws.L, M, ws.dinv, ws.G2, M, ws.G1, M, sm);
        double guu_var, guu_ls;
        gram_square_grad(kp, M, Z, ws.K0, ws.G2, guu_var, guu_ls, sm);
        gvar = gv_sum + acc[0] + guu_var;
        gls = acc[1] + guu_ls;
    }
    store_hyper_grad(md, h, gvar, gls, dalpha, dkm, grad);
    *f_out = ve - kl;
    __syncthreads();
    return GPC_ST_OK;
}

// ------------------------------------------------------------------------------------------------ GPR
// log N(y | 0, K + noise I);  theta = (var, ls, noise, -)
__device__ __noinline__ int eval_gpr(const ModelDev& md, const double* y, const double* theta, double* grad, double* f_out,
                        const SlotWs& ws, CtaSmem* sm) {
    sm = cta_arena();
    const int N = md.N, tid = threadIdx.x;
    const Hyper h = hyper_from_theta(md, theta);
    KernParams kp;
    kern_setup(kp, md, h.var, h.ls);
    double* v = ws.v0; double* w = ws.v1;
    const double shift = h.alpha + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    build_gram_square(kp, N, md.X, shift, ws.K0, ws.L);
    if (!cta_potrf(N, ws.L, N, ws.dinv, sm)) return GPC_ST_CHOL;
    for (int n = tid; n < N; n += GPC_THREADS) v[n] = y[n];
    __syncthreads();
    cta_trsm_left(N, 1, ws.L, N, ws.dinv, v, 1, sm);
    double acc[2] = {0.0, 0.0};
    for (int n = tid; n < N; n += GPC_THREADS) {
        w[n] = v[n];
        acc[0] = fma(v[n], v[n], acc[0]);
        acc[1] += log(ws.L[(size_t)n * N + n]);
    }
    cta_sum_n<2>(acc, sm->red);
    cta_trsm_left_t(N, 1, ws.L, N, ws.dinv, w, 1, sm);
    // Lbar = tril(w v^T) - diag(1/L_ii)
    GPC_FOR_2D(N, N, i, j) {
        if (j <= i) ws.G2[idx] = w[i] * v[j] - (i == j ? 1.0 / ws.L[idx] : 0.0);
    }
    __syncthreads();
    cta_chol_backward(N, ws.L, N, ws.dinv, ws.G2, N, ws.G3, N, sm);
    double gvar, gls;
    gram_square_grad(kp, N, md.X, ws.K0, ws.G2, gvar, gls, sm);
    double tr = 0.0;
    for (int n = tid; n < N; n += GPC_THREADS) tr += ws.G2[(size_t)n * N + n];
    tr = cta_sum(tr, sm->red);
    store_hyper_grad(md, h, gvar, gls, tr, 0.0, grad);
    *f_out = -0.5 * N * 1.8378770664093453 - acc[1] - 0.5 * acc[0];
    __syncthreads();
    return GPC_ST_OK;
}


// ------------------------------------------------------------------------------------------------ SGPR
// Titsias collapsed bound (gpflow.models.SGPR.elbo; reference call site GPcounts_Module.py:661-663):
//   V = L^-1 Kuf,  B = V V^T / s2 + I = LB LB^T,  c = LB^-1 V y / s2
//   F = -N/2 log 2pi - sum log LB_ii - N/2 log s2 - y'y/(2 s2) + c'c/2 - sum(kdiag)/(2 s2) + tr(V V^T)/(2 s2)
// theta = (var, ls, noise s2, -); reverse pass through both Cholesky factors.
__device__ __noinline__ int eval_sgpr(const ModelDev& md, const double* Z, const double* y, const double* theta,
                                      double* grad, double* f_out, const SlotWs& ws, CtaSmem* sm) {
    sm = cta_arena();
    const int N = md.N, M = md.M, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Hyper h = hyper_from_theta(md, theta);
    KernParams kp;
    kern_setup(kp, md, h.var, h.ls);
    const double s2 = h.alpha, is2 = 1.0 / s2;
    const double shift = md.jitter + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    const double kdiag = h.var + (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
    double* V = ws.A; double* Vb = ws.B; double* Bm = ws.Lq; double* VVt = ws.G4;
    double* cvec = ws.v0; double* wvec = ws.v1; double* uvec = ws.v2;
    build_gram_square(kp, M, Z, shift, ws.K0, ws.L);
    GPC_FOR_2D(M, N, i, n) {
        bool cross;
        const double r2 = kern_r2(kp, Z + (size_t)i * kp.d, md.X + (size_t)n * kp.d, cross);
        const double k = kern_val(kp, r2, cross);
        ws.Kuf[idx] = k;
        V[idx] = k;
    }
    __syncthreads();
    if (!cta_potrf(M, ws.L, M, ws.dinv, sm)) return GPC_ST_CHOL;
    cta_trsm_left(M, N, ws.L, M, ws.dinv, V, N, sm);
    // VVt = V V^T / s2 (full), B = VVt + I
    cta_gemm<false, true>(M, M, N, is2, V, N, V, N, 0.0, VVt, M, 0, sm);
    double acc[2] = {0.0, 0.0};
    GPC_FOR_2D(M, M, i, j) {
        const double v = VVt[idx];
        Bm[idx] = (i == j) ? v + 1.0 : v;
        if (i == j) acc[0] += v;                       // tr(V V^T) / s2
    }
    for (int n = tid; n < N; n += GPC_THREADS) acc[1] = fma(y[n], y[n], acc[1]);
    cta_sum_n<2>(acc, sm->red);
    const double trvv = acc[0], yy = acc[1];
    if (!cta_potrf(M, Bm, M, ws.dinv2, sm)) return GPC_ST_CHOL;
    // u = V y;  a = u / s2;  c = LB^-1 a;  w = LB^-T c
    for (int row = warp; row < M; row += GPC_THREADS / 32) {
        double s = 0.0;
        for (int n = lane; n < N; n += 32) s = fma(V[(size_t)row * N + n], y[n], s);
        s = warp_sum(s);
        if (lane == 0) { uvec[row] = s; cvec[row] = s * is2; }
    }
    __syncthreads();
    cta_trsm_left(M, 1, Bm, M, ws.dinv2, cvec, 1, sm);
    double a3[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < M; i += GPC_THREADS) {
        wvec[i] = cvec[i];
        a3[0] = fma(cvec[i], cvec[i], a3[0]);
        a3[1] += log(Bm[(size_t)i * M + i]);
    }
    cta_sum_n<3>(a3, sm->red);
    const double cc = a3[0], logdetLB = a3[1];
    cta_trsm_left_t(M, 1, Bm, M, ws.dinv2, wvec, 1, sm);
    // LB-bar = -tril(w c^T) - diag(1 / LB_ii)  -> G2;  reverse pass of chol(B)
    GPC_FOR_2D(M, M, i, j) {
        if (j <= i) ws.G2[idx] = -wvec[i] * cvec[j] - (i == j ? 1.0 / Bm[idx] : 0.0);
    }
    __syncthreads();
    cta_chol_backward(M, Bm, M, ws.dinv2, ws.G2, M, ws.G3, M, sm);
    // W = Bbar + Bbar^T -> G1;  <Bbar, VVt>;  w . a
    double a2[2] = {0.0, 0.0};
    GPC_FOR_2D(M, M, i, j) {
        ws.G1[idx] = ws.G2[idx] + ws.G2[(size_t)j * M + i];
        a2[0] = fma(ws.G2[idx], VVt[idx], a2[0]);
    }
    for (int i = tid; i < M; i += GPC_THREADS) a2[1] = fma(wvec[i], uvec[i] * is2, a2[1]);
    cta_sum_n<2>(a2, sm->red);
    // Vbar = (W V + w y^T + V) / s2
    cta_gemm<false, false>(M, N, M, is2, ws.G1, M, V, N, 0.0, Vb, N, 0, sm);
    GPC_FOR_2D(M, N, i, n) {
        Vb[idx] += (wvec[i] * y[n] + V[idx]) * is2;
    }
    __syncthreads();
    cta_trsm_left_t(M, N, ws.L, M, ws.dinv, Vb, N, sm);                         // S = L^-T Vbar = Kuf-bar
    cta_gemm<false, true>(M, M, N, -1.0, Vb, N, V, N, 0.0, ws.G2, M, F_C_LOWER, sm);   // L-bar = -tril(S V^T)
    double ak[2] = {0.0, 0.0};
    GPC_FOR_2D(M, N, i, n) {
        bool cross;
        const double r2 = kern_r2(kp, Z + (size_t)i * kp.d, md.X + (size_t)n * kp.d, cross);
        kern_grad_acc(kp, Vb[idx], ws.Kuf[idx], r2, cross, ak[0], ak[1]);
    }
    cta_sum_n<2>(ak, sm->red);
    cta_chol_backward(M, ws.L, M, ws.dinv, ws.G2, M, ws.G3, M, sm);
    double guu_var, guu_ls;
    gram_square_grad(kp, M, Z, ws.K0, ws.G2, guu_var, guu_ls, sm);
    const double gvar = guu_var + ak[0] - 0.5 * N * is2;
    const double gls = guu_ls + ak[1];
    // d/d s2: explicit terms + through B (= -<Bbar, VVt> / s2) + through a (= -w.a / s2) + tr term
    const double gs2 = -0.5 * N * is2 + 0.5 * yy * is2 * is2 + 0.5 * N * kdiag * is2 * is2 - 0.5 * trvv * is2
                       - a2[0] * is2 - a2[1] * is2;
    store_hyper_grad(md, h, gvar, gls, gs2, 0.0, grad);
    *f_out = -0.5 * N * 1.8378770664093453 - logdetLB - 0.5 * N * log(s2) - 0.5 * yy * is2 + 0.5 * cc
             - 0.5 * N * kdiag * is2 + 0.5 * trvv;
    __syncthreads();
    return GPC_ST_OK;
}

struct FusedSmem;
__device__ __noinline__ int eval_svgp_fused(const ModelDev& md, const double* Z, const double* y, const double* scale,
                               const double* theta, double* grad, double* f_out, const SlotWs& ws, FusedSmem* fs);

// The on-chip evaluator covers SVGP with M <= 64 inducing points, d <= 2, RBF / Constant kernels.
__device__ __forceinline__ bool use_fused(const ModelDev& md) {
    return md.kind == GPC_SVGP && md.M <= 64 && md.d <= 2 && md.kernel != GPC_BRANCH;
}

#define GPC_SMALL_ARENA_OFF 8704     // small instantiation: [OptSmem | optimiser vectors | arena of eval_vgp_small]
#define GPC_SMALL_PVEC(n) ((((size_t)(4 + (n) + (n) * ((n) + 1) / 2) + 15) & ~(size_t)15))   // padded parameter count
#define GPC_SMALL_VEC_BYTES(n) (6 * GPC_SMALL_PVEC(n) * sizeof(double))
__device__ __noinline__ int eval_vgp_small(const ModelDev& md, const double* y, const double* scale, const double* theta,
                                           double* grad, double* f_out, unsigned char* arena);

__device__ __noinline__ int model_eval_direct(const ModelDev& md, const double* Z, const double* y, const double* scale,
                          const double* theta, double* grad, double* f_out, const SlotWs& ws, CtaSmem* sm) {
#if defined(GPC_SMALL_ONLY)
    return eval_vgp_small(md, y, scale, theta, grad, f_out, g_smem + ws.arena_off);
#elif defined(GPC_FUSED_ONLY)
    return eval_svgp_fused(md, Z, y, scale, theta, grad, f_out, ws, reinterpret_cast<FusedSmem*>(sm));
#else
#ifndef GPC_DENSE_ONLY
    if (use_fused(md))
        return eval_svgp_fused(md, Z, y, scale, theta, grad, f_out, ws, reinterpret_cast<FusedSmem*>(sm));
#endif
    switch (md.kind) {
        case GPC_VGP: return eval_vgp(md, y, scale, theta, grad, f_out, ws, sm);
        case GPC_SVGP: return eval_svgp(md, Z, y, scale, theta, grad, f_out, ws, sm);
        case GPC_GPR: return eval_gpr(md, y, theta, grad, f_out, ws, sm);
        case GPC_SGPR: return eval_sgpr(md, Z, y, theta, grad, f_out, ws, sm);
        default: return GPC_ST_CHOL;
    }
#endif
}

// The kernels call the evaluator through a device function pointer.  An indirect call is compiled to the plain ABI
// (callee-saved registers stored once on entry, restored on exit), which makes the evaluators' register allocation
// independent of the calling kernel.  Called directly, ptxas' interprocedural allocation gave every kernel its own
// clone of each evaluator whose quality depended on unrelated code of that kernel: `cta_gemm` reloaded ~10 spilled
// values per k-panel inside `k_fit` under the 128-register cap of the dense instantiation (L2 round trips), and the
// fused evaluator ran 20-35 % slower inside `k_fit` than in `k_elbo_grad`.  The price of the ABI is that address
// spaces are no longer inferred across the call: evaluators re-derive their shared-memory arena from `g_smem`
// (a generic `sm` / `fs` argument would turn every LDS/STS into a generic load / store).
typedef int (*ModelEvalFn)(const ModelDev&, const double*, const double*, const double*, const double*, double*, double*,
                           const SlotWs&, CtaSmem*);
__device__ ModelEvalFn g_model_eval = model_eval_direct;

__device__ __forceinline__ int model_eval(const ModelDev& md, const double* Z, const double* y, const double* scale,
                          const double* theta, double* grad, double* f_out, const SlotWs& ws, CtaSmem* sm) {
    ModelEvalFn fn = *(volatile ModelEvalFn*)&g_model_eval;
    return fn(md, Z, y, scale, theta, grad, f_out, ws, sm);
}
