// Persistent per-gene kernels (fit and fixed-theta evaluation) and their launchers.
//
// This header is compiled three times with different macros (GPC_KNAME() gives the instantiations distinct names):
//   f256  (gpc_fused.cu) GPC_FUSED_ONLY: only the fused on-chip SVGP evaluator (+ optimiser).  Used when every model of
//         a call takes the fused path - the headline C5 workload.  Keeping the dense call graph out of this
//         instantiation matters: with it in the same kernel ptxas spilled ~1.5-2 KB per thread in the evaluator
//         (730 instead of 430 us per evaluation) depending on unrelated changes to the dense code.
//   d128  (gpc_dense.cu) GPC_DENSE_ONLY, __launch_bounds__(256, 2): the dense evaluators need only ~60 KB of shared
//         memory, so capping them at 128 registers lets two CTAs share an SM and hide each other's global-load
//         latency; used when no model of a call takes the fused path.
//   t256  (gpc_api.cu)   every evaluator, one CTA per SM: calls that mix fused and dense models.
// (A 512-thread, fused-only instantiation - GPC_THREADS=512, GPC_FUSED_ONLY - was measured slower: 494 vs 452 us per
// C5 evaluation, its 128-register cap spills the evaluator's accumulators.)
#pragma once
#include <algorithm>
#include <math.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/gpcounts_b200.h"

// launcher table of one instantiation (argument structs are passed as void* because each translation unit has its
// own, anonymous-namespace copy of the device code and of the argument types)
struct VariantOps {
    int (*upload_constants)(const double* z, const double* w);
    size_t (*smem_bytes)(int nmax);             // nmax: largest N of the models (sizes the arena of the small instantiation)
    int (*occupancy)(int nmax);
    cudaError_t (*launch_fit)(const void* fit_args, int grid, int nmax, cudaStream_t s);
    cudaError_t (*launch_eval)(const void* eval_args, int grid, int nmax, cudaStream_t s);
    int (*read_prof)(unsigned long long* out16);   // profiling build only (GPC_PHASE_PROF), else nullptr
    cudaError_t (*launch_predict)(const void* predict_args, int grid, cudaStream_t s);   // dense instantiations only, else nullptr
};

namespace {   // everything below is private to the including translation unit
#include "lbfgsb.cuh"

#define GPC_CAT2(a, b) a##_##b
#define GPC_CAT(a, b) GPC_CAT2(a, b)
#define GPC_KNAME(name) GPC_CAT(name, GPC_VARIANT)

// ---------------------------------------------------------------------------------------------- layout
struct WsLayout {
    size_t sq;          // doubles per square buffer
    size_t dinv;        // doubles of diagonal-block inverses
    size_t rect;        // doubles per M x N buffer
    size_t vec;         // doubles per N vector
    size_t pvec;        // doubles per parameter vector (padded)
    int m;              // L-BFGS memory
    size_t slot;        // doubles per slot
};

static inline size_t pad8(size_t x) { return (x + 7) & ~(size_t)7; }

static int model_mv(const gpc_model_t& md) {
    return md.kind == GPC_VGP ? md.N : (md.kind == GPC_SVGP ? md.M : 0);
}

static WsLayout make_layout(const gpc_model_t* models, int n_models, int m, bool small_only = false) {
    WsLayout L{};
    L.m = m;
    for (int k = 0; k < n_models; ++k) {
        const gpc_model_t& md = models[k];
        const bool sparse = md.kind == GPC_SVGP || md.kind == GPC_SGPR;
        const size_t sdim = sparse ? md.M : md.N;
        const size_t mv = model_mv(md);
        const size_t P = GPC_NHYP + mv + mv * (mv + 1) / 2;
        L.sq = std::max(L.sq, std::max(pad8(sdim * sdim), (size_t)4096));   // >= 64 x 64: scratch of the fused evaluator
        L.dinv = std::max(L.dinv, (size_t)((sdim + GPC_BLK - 1) / GPC_BLK) * GPC_BLK * GPC_BLK);
        if (sparse) L.rect = std::max(L.rect, pad8((size_t)md.M * md.N));
        L.vec = std::max(L.vec, pad8((size_t)std::max(md.N, md.M)));
        L.pvec = std::max(L.pvec, pad8(P));
    }
    if (small_only) { L.sq = 0; L.dinv = 0; L.rect = 0; L.vec = 0; }   // the on-chip small-N evaluator keeps its factors in shared memory
    L.slot = 7 * L.sq + 2 * L.dinv + 4 * L.rect + 4 * L.vec + (size_t)(6 + 2 * m) * L.pvec;
    return L;
}

__device__ __forceinline__ SlotWs carve(double* base, const WsLayout& L) {
    SlotWs w;
    double* p = base;
    w.K0 = p; p += L.sq; w.L = p; p += L.sq; w.Lq = p; p += L.sq;
    w.G1 = p; p += L.sq; w.G2 = p; p += L.sq; w.G3 = p; p += L.sq;
    w.G4 = p; p += L.sq;
    w.dinv = p; p += L.dinv; w.dinv2 = p; p += L.dinv;
    w.Kuf = p; p += L.rect; w.A = p; p += L.rect; w.B = p; p += L.rect; w.S = p; p += L.rect;
    w.v0 = p; p += L.vec; w.v1 = p; p += L.vec; w.v2 = p; p += L.vec; w.v3 = p; p += L.vec;
    w.x = p; p += L.pvec; w.g = p; p += L.pvec; w.d = p; p += L.pvec;
    w.xold = p; p += L.pvec; w.gold = p; p += L.pvec; w.z = p; p += L.pvec;
    w.WS = p; p += (size_t)L.m * L.pvec; w.WY = p;
    w.tmY = nullptr; w.tmS = nullptr; w.row = 0; w.fresh = 1; w.arena_off = GPC_SMALL_ARENA_OFF;
    return w;
}

__host__ __device__ static inline ModelDev to_dev(const gpc_model_t& md) {
    ModelDev d;
    d.kind = md.kind; d.kernel = md.kernel; d.lik = md.lik; d.N = md.N; d.M = md.M; d.d = md.d;
    d.n_off = md.n_off; d.train_mask = md.train_mask; d.n_zsets = md.n_zsets; d.handoff_alpha = md.handoff_alpha;
    d.xp = md.xp; d.jitter = md.jitter; d.X = md.X; d.Z = md.Z; d.Zgene = md.Zgene; d.restart = md.restart;
    d.Mv = md.kind == GPC_VGP ? md.N : (md.kind == GPC_SVGP ? md.M : 0);
    d.P = GPC_NHYP + d.Mv + d.Mv * (d.Mv + 1) / 2;
    return d;
}

// ---------------------------------------------------------------------------------------------- kernels
#define GPC_MAX_MODELS_CONST 64
#ifdef GPC_SMALL_ONLY
#define GPC_SMALL_NMAX_OR_0 GPC_SMALL_NMAX
#else
#define GPC_SMALL_NMAX_OR_0 0
#endif
struct FitArgs {
    const ModelDev* models;   // device array
    int n_models, chain, D;
    gpc_opt_t opt;
    const double *Y, *scale, *hyp0;
    long long ldy, ld_theta;
    double *stats, *theta_out;
    double* ws;
    int* queue;
    WsLayout lay;
    double* trace;            // debug: evaluation trace of (trace_gene, trace_model), 8 doubles per evaluation
    int trace_cap, trace_gene, trace_model;
    int nmax;                 // largest N of the models (sizes the shared-memory carve of the small instantiation)
    int use_tma;              // tmY (and tmS when scale != nullptr) are valid
    alignas(64) CUtensorMap tmY;   // Y     as a 2-D FP64 tensor [D][ldy], box 64 x 1 (one column tile of one gene)
    alignas(64) CUtensorMap tmS;   // scale likewise
};


// The dense (CtaSmem) and fused (FusedSmem) evaluators overlay the same shared-memory region.
#ifndef GPC_MIN_BLOCKS
#define GPC_MIN_BLOCKS 1
#endif

__host__ __device__ constexpr size_t eval_smem_bytes() {
#if defined(GPC_SMALL_ONLY)
    return GPC_SMALL_ARENA_OFF;      // [OptSmem | evaluator arena sized by N at launch]; see small_total_smem()
#elif defined(GPC_DENSE_ONLY)
    return (sizeof(CtaSmem) + 15) & ~(size_t)15;
#else
    return ((sizeof(CtaSmem) > sizeof(FusedSmem) ? sizeof(CtaSmem) : sizeof(FusedSmem)) + 127) & ~(size_t)127;
#endif
}
#if defined(GPC_SMALL_ONLY)
static_assert(sizeof(OptSmem) <= GPC_SMALL_ARENA_OFF, "OptSmem must fit in front of the small evaluator's arena");
#else
static_assert(eval_smem_bytes() + sizeof(OptSmem) + 1024 <= 232448, "shared memory of the fit kernel exceeds 227 KB");
#endif

__device__ __forceinline__ double softplus_inv_d(double p) { return p + log(-expm1(-p)); }

// Fit one (gene, model); returns true when the fit succeeded. prev_alpha < 0: no hand-off.
__device__ bool fit_one(const FitArgs& a, int g, int k, double prev_alpha, const SlotWs& ws, CtaSmem* sm, OptSmem* os,
                        double& alpha_out) {
    // the model descriptor lives in shared memory: it is passed by reference into out-of-line evaluators, and a
    // per-thread copy would sit in local memory behind an L1 that the 225 KB shared-memory carve-out leaves tiny
    __shared__ ModelDev s_md;
    __syncthreads();
    if (threadIdx.x == 0) s_md = a.models[k];
    __syncthreads();
    const ModelDev& md = s_md;
    const int tid = threadIdx.x;
    const double* y = a.Y + (size_t)g * a.ldy + md.n_off;
    const double* scale = a.scale ? a.scale + (size_t)g * a.ldy + md.n_off : nullptr;
    const double* h0 = a.hyp0 + ((size_t)g * a.n_models + k) * 4;
    double* st = a.stats + ((size_t)g * a.n_models + k) * GPC_NSTAT;
    const int max_restart = md.restart ? GPC_MAX_RESTART : 0;
    int nfev_total = 0;
    long long eval_cycles = 0;
    const long long t_start = clock64();
    OptResult r;
    int attempt = a.opt.attempt0 ? a.opt.attempt0[(size_t)g * a.n_models + k] : 0;      // host-driven restarts (safe_mode)
    if (attempt > max_restart) attempt = max_restart;
    bool ok = false;
    for (;; ++attempt) {
        double var, ls, alpha, km;
        if (attempt == 0) {
            var = h0[0]; ls = h0[1]; alpha = h0[2]; km = h0[3];
            if (md.handoff_alpha && prev_alpha > 0.0) alpha = prev_alpha;
        } else {
            const double* t = md.restart + (size_t)(attempt - 1) * 4;
            ls = t[0]; var = t[1]; alpha = t[2]; km = t[3];
        }
        if (md.kernel == GPC_CONST) ls = 1.0;
        {
            // gpflow refuses a Parameter whose unconstrained value is not finite (e.g. variance 0 of an all-zero gene:
            // softplus^-1(0) = -inf) with tf.errors.InvalidArgumentError, which fit_GP answers like a failed Cholesky
            // factorisation: random restart (:547-556).  CTA-uniform values.
            const bool use_al = md.lik == GPC_NB || md.lik == GPC_ZINB || md.lik == GPC_GAUSS;
            const double al_eff = md.lik == GPC_GAUSS ? alpha - GPC_GAUSS_VAR_LOWER : alpha;
            bool finite0 = isfinite(softplus_inv_d(var)) && isfinite(softplus_inv_d(ls));
            if (use_al) finite0 = finite0 && isfinite(softplus_inv_d(al_eff));
            if (md.lik == GPC_ZINB) finite0 = finite0 && isfinite(softplus_inv_d(km));
            if (!finite0) {
                r.task = TASK_CHOL; r.nit = 0; r.nfev = 0; r.f = NAN; r.gnorm = NAN; r.eval_cycles = 0;
                ok = false;
                if (attempt >= max_restart) break;
                continue;
            }
        }
        // theta0: hyper-parameters through softplus^-1, q_mu = 0, q_sqrt = I
        for (int i = tid; i < md.P; i += GPC_THREADS) ws.x[i] = 0.0;
        __syncthreads();
        if (tid == 0) {
            ws.x[0] = softplus_inv_d(var);
            ws.x[1] = softplus_inv_d(ls);
            ws.x[2] = softplus_inv_d(md.lik == GPC_GAUSS ? alpha - GPC_GAUSS_VAR_LOWER : alpha);
            ws.x[3] = md.lik == GPC_ZINB ? softplus_inv_d(km) : (km > 0.0 && isfinite(softplus_inv_d(km)) ? softplus_inv_d(km) : 0.0);
        }
        for (int i = tid; i < md.Mv; i += GPC_THREADS) ws.x[GPC_NHYP + md.Mv + tri_index(i, i)] = 1.0;
        __syncthreads();
        const double* Z = md.Z ? md.Z + (size_t)(md.n_zsets > 1 ? attempt : 0) * md.M * md.d : nullptr;
        if (attempt == 0 && md.Zgene) Z = md.Zgene + (size_t)g * md.M * md.d;
        if (tid == 0) const_cast<SlotWs&>(ws).fresh = 1;      // the slot workspace lives in shared memory (s_ws)
        __syncthreads();
        OptTrace tr{nullptr, 0, 0};
        if (a.trace && g == a.trace_gene && k == a.trace_model && attempt == 0) { tr.buf = a.trace; tr.cap = a.trace_cap; }
        r = lbfgsb_minimize(md, Z, y, scale, a.opt, ws, sm, os, &tr);
        nfev_total += r.nfev;
        eval_cycles += r.eval_cycles;
        ok = (r.task == TASK_PGTOL || r.task == TASK_FTOL);
        if (ok || attempt >= max_restart) break;
    }
    __syncthreads();
    const Hyper h = hyper_from_theta(md, ws.x);
    alpha_out = h.alpha;
    if (tid == 0) {
        st[GPC_S_LL] = ok ? -r.f : NAN;
        st[GPC_S_STATUS] = ok ? GPC_ST_OK : (r.task == TASK_CHOL ? GPC_ST_CHOL : GPC_ST_OPT);
        st[GPC_S_NIT] = r.nit;
        st[GPC_S_NFEV] = r.nfev;
        st[GPC_S_RESTARTS] = attempt;
        st[GPC_S_VAR] = h.var; st[GPC_S_LS] = h.ls; st[GPC_S_ALPHA] = h.alpha; st[GPC_S_KM] = h.km;
        st[GPC_S_NFEV_TOTAL] = nfev_total;
        st[GPC_S_GNORM] = r.gnorm;
        st[GPC_S_TASK] = r.task;
        st[GPC_S_EVAL_CYCLES] = (double)eval_cycles;
        st[GPC_S_FIT_CYCLES] = (double)(clock64() - t_start);
        for (int q = 14; q < GPC_NSTAT; ++q) st[q] = 0.0;
    }
    if (a.theta_out) {
        double* out = a.theta_out + ((size_t)g * a.n_models + k) * a.ld_theta;
        for (int i = tid; i < md.P; i += GPC_THREADS) out[i] = ws.x[i];
    }
    __syncthreads();
    return ok;
}

__global__ void __launch_bounds__(GPC_THREADS, GPC_MIN_BLOCKS) GPC_KNAME(k_fit)(const __grid_constant__ FitArgs a) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
#if defined(GPC_SMALL_ONLY)
    OptSmem* os = reinterpret_cast<OptSmem*>(g_smem);           // the evaluator's arena follows at GPC_SMALL_ARENA_OFF
#else
    OptSmem* os = reinterpret_cast<OptSmem*>(g_smem + eval_smem_bytes());
#endif
    __shared__ int s_job;
    __shared__ SlotWs s_ws;
    if (threadIdx.x == 0) {
        s_ws = carve(a.ws + (size_t)blockIdx.x * a.lay.slot, a.lay);
        if (a.use_tma) { s_ws.tmY = &a.tmY; s_ws.tmS = a.scale ? &a.tmS : nullptr; }
#if defined(GPC_SMALL_ONLY)
        {
            // the six optimiser vectors (P <= 1228 doubles each) sit in shared memory in front of the evaluator's arena:
            // at N = 18 an L-BFGS iteration is a chain of ~20 dependent vector accesses, each an L2 round trip otherwise
            const int n = a.nmax < 1 ? 1 : (a.nmax > GPC_SMALL_NMAX ? GPC_SMALL_NMAX : a.nmax);
            double* v = reinterpret_cast<double*>(g_smem + GPC_SMALL_ARENA_OFF);
            const size_t pv = GPC_SMALL_PVEC(n);
            s_ws.x = v; s_ws.g = v + pv; s_ws.d = v + 2 * pv; s_ws.xold = v + 3 * pv; s_ws.gold = v + 4 * pv; s_ws.z = v + 5 * pv;
            s_ws.arena_off = GPC_SMALL_ARENA_OFF + (int)GPC_SMALL_VEC_BYTES(n);
        }
#endif
    }
    __syncthreads();
    const SlotWs& ws = s_ws;
    const int n_jobs = a.chain ? a.D : a.D * a.n_models;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_job = atomicAdd(a.queue, 1);
        __syncthreads();
        const int job = s_job;
        if (job >= n_jobs) break;
        if (threadIdx.x == 0) s_ws.row = a.chain ? job : job / a.n_models;
        __syncthreads();
        if (a.chain) {
            double prev_alpha = -1.0;
            bool alive = true;
            for (int k = 0; k < a.n_models; ++k) {
                if (!alive) {
                    if (threadIdx.x == 0) {
                        double* st = a.stats + ((size_t)job * a.n_models + k) * GPC_NSTAT;
                        for (int q = 0; q < GPC_NSTAT; ++q) st[q] = 0.0;
                        st[GPC_S_LL] = NAN;
                        st[GPC_S_STATUS] = GPC_ST_SKIPPED;
                    }
                    continue;
                }
                double alpha_out;
                alive = fit_one(a, job, k, prev_alpha, ws, sm, os, alpha_out);
                prev_alpha = alpha_out;
            }
        } else {
            double alpha_out;
            fit_one(a, job / a.n_models, job % a.n_models, -1.0, ws, sm, os, alpha_out);
        }
    }
}

struct EvalArgs {
    ModelDev md;
    int D, zset;
    const double *Y, *scale, *theta;
    long long ldy;
    double *elbo, *grad;
    int* status;
    double* ws;
    int* queue;
    WsLayout lay;
    int use_tma;
    alignas(64) CUtensorMap tmY;
    alignas(64) CUtensorMap tmS;
};

__global__ void __launch_bounds__(GPC_THREADS, GPC_MIN_BLOCKS) GPC_KNAME(k_elbo_grad)(const __grid_constant__ EvalArgs a) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
    __shared__ int s_job;
    __shared__ SlotWs s_ws;
    if (threadIdx.x == 0) {
        s_ws = carve(a.ws + (size_t)blockIdx.x * a.lay.slot, a.lay);
        if (a.use_tma) { s_ws.tmY = &a.tmY; s_ws.tmS = a.scale ? &a.tmS : nullptr; }
    }
    const SlotWs& ws = s_ws;
    // the descriptor is passed by reference into out-of-line evaluators: a per-thread copy would live in local memory
    // (L2 round trips behind the small L1 that the shared-memory carve-out leaves), so it sits in shared memory as in k_fit
    __shared__ ModelDev s_md;
    if (threadIdx.x == 0) s_md = a.md;
    const ModelDev& md = s_md;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) { s_job = atomicAdd(a.queue, 1); s_ws.row = s_job; s_ws.fresh = 1; }
        __syncthreads();
        const int g = s_job;
        if (g >= a.D) break;
        const double* y = a.Y + (size_t)g * a.ldy + md.n_off;
        const double* scale = a.scale ? a.scale + (size_t)g * a.ldy + md.n_off : nullptr;
        const double* Z = md.Z ? md.Z + (size_t)a.zset * md.M * md.d : nullptr;
        if (md.Zgene) Z = md.Zgene + (size_t)g * md.M * md.d;
        double* grad = a.grad + (size_t)g * md.P;
        double f;
        const int st = model_eval(md, Z, y, scale, a.theta + (size_t)g * md.P, grad, &f, ws, sm);
        if (st != GPC_ST_OK)
            for (int i = threadIdx.x; i < md.P; i += GPC_THREADS) grad[i] = NAN;
        if (threadIdx.x == 0) {
            a.elbo[g] = st == GPC_ST_OK ? f : NAN;
            a.status[g] = st;
        }
    }
}

// ---------------------------------------------------------------------------------------------- posterior prediction
// Replaces `model.predict_f(Xnew)` / `predict_y` / the mean and covariance behind `predict_f_samples` of a fitted gpflow model
// (reference: load_predict_models GPcounts_Module.py:930-945, infer_branching :305-317, test_local_optima_case1 :962-980),
// batched over genes from the fitted parameter vectors theta-hat that gpc_fit returns:
//   VGP / SVGP (whitened):  A = Lm^-1 K(Zr, Xnew), Zr = X | Z;  mean = A^T q_mu;  cov = K(Xnew, Xnew) - A^T A + (Lq^T A)^T (Lq^T A)
//   GPR:                    A = L^-1 K(X, Xnew), L = chol(K + s2 I), v = L^-1 y;  mean = A^T v;  cov = K(Xnew, Xnew) - A^T A
// `var` is the diagonal of cov (latent f; predict_y adds the Gaussian noise on the host); `cov` (T x T) is optional.
struct PredictArgs {
    ModelDev md;
    int D, T, zset;
    const double *Y, *theta, *Xnew;
    long long ldy, ld_theta;
    double *mean, *var, *cov;
    double* ws;
    int* queue;
    size_t sq, dinv, rect, vec, slot;     // doubles
};

#if !defined(GPC_FUSED_ONLY) && !defined(GPC_SMALL_ONLY)
__global__ void __launch_bounds__(GPC_THREADS, GPC_MIN_BLOCKS) GPC_KNAME(k_predict)(const __grid_constant__ PredictArgs a) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
    __shared__ int s_job;
    __shared__ ModelDev s_md;
    if (threadIdx.x == 0) s_md = a.md;
    const ModelDev& md = s_md;
    const int tid = threadIdx.x, T = a.T;
    double* base = a.ws + (size_t)blockIdx.x * a.slot;
    double* K0 = base; double* L = K0 + a.sq; double* Lq = L + a.sq; double* dinv = Lq + a.sq;
    double* A = dinv + a.dinv; double* B = A + a.rect; double* v = B + a.rect;
    while (true) {
        __syncthreads();
        if (tid == 0) s_job = atomicAdd(a.queue, 1);
        __syncthreads();
        const int g = s_job;
        if (g >= a.D) break;
        const double* theta = a.theta + (size_t)g * a.ld_theta;
        const bool sparse = md.kind == GPC_SVGP || md.kind == GPC_SGPR;
        const int Mr = sparse ? md.M : md.N;                       // rows of the factor
        const double* Zr = sparse ? md.Z + (size_t)a.zset * md.M * md.d : md.X;
        const Hyper h = hyper_from_theta(md, theta);
        KernParams kp;
        kern_setup(kp, md, h.var, h.ls);
        const double noise = (md.kernel == GPC_BRANCH ? GPC_BRANCH_NOISE : 0.0);
        const double shift = (md.kind == GPC_GPR ? h.alpha : md.jitter) + noise;
        const double kdiag = h.var + noise;
        double* mean = a.mean + (size_t)g * T;
        double* var = a.var + (size_t)g * T;
        double* cov = a.cov ? a.cov + (size_t)g * T * T : nullptr;
        bool ok = md.kind != GPC_SGPR;                               // SGPR prediction is not provided (fit only)
        if (ok) {
            build_gram_square(kp, Mr, Zr, shift, K0, L);
            ok = cta_potrf(Mr, L, Mr, dinv, sm);
        }
        if (!ok) {
            for (int t = tid; t < T; t += GPC_THREADS) { mean[t] = NAN; var[t] = NAN; }
            if (cov) for (int t = tid; t < T * T; t += GPC_THREADS) cov[t] = NAN;
            continue;
        }
        GPC_FOR_2D(Mr, T, i, t) {
            bool cross;
            const double r2 = kern_r2(kp, Zr + (size_t)i * kp.d, a.Xnew + (size_t)t * kp.d, cross);
            A[idx] = kern_val(kp, r2, cross);
        }
        __syncthreads();
        cta_trsm_left(Mr, T, L, Mr, dinv, A, T, sm);               // A = L^-1 K(Zr, Xnew)
        const double* wvec;                                          // mean = A^T wvec
        if (md.kind == GPC_GPR) {
            const double* y = a.Y + (size_t)g * a.ldy + md.n_off;
            for (int n = tid; n < Mr; n += GPC_THREADS) v[n] = y[n];
            __syncthreads();
            cta_trsm_left(Mr, 1, L, Mr, dinv, v, 1, sm);
            wvec = v;
        } else {
            wvec = theta + GPC_NHYP;                                 // q_mu
            unpack_lq(Mr, theta + GPC_NHYP + Mr, Lq);
            cta_gemm<true, false>(Mr, T, Mr, 1.0, Lq, Mr, A, T, 0.0, B, T, F_A_UPPER, sm);   // B = Lq^T A
        }
        for (int t = tid; t < T; t += GPC_THREADS) {
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int i = 0; i < Mr; ++i) {
                const double av = A[(size_t)i * T + t];
                s1 = fma(av, wvec[i], s1);
                s2 = fma(av, av, s2);
                if (md.kind != GPC_GPR) { const double bv = B[(size_t)i * T + t]; s3 = fma(bv, bv, s3); }
            }
            mean[t] = s1;
            var[t] = (kdiag - s2) + s3;
        }
        if (cov) {
            // cov = K(Xnew, Xnew) - A^T A (+ B^T B); the BranchKernel adds its noise only on the diagonal of a square Gram
            GPC_FOR_2D(T, T, i, t) {
                bool cross;
                const double r2 = kern_r2(kp, a.Xnew + (size_t)i * kp.d, a.Xnew + (size_t)t * kp.d, cross);
                cov[idx] = kern_val(kp, r2, cross) + (i == t ? noise : 0.0);
            }
            __syncthreads();
            cta_gemm<true, false>(T, T, Mr, -1.0, A, T, A, T, 1.0, cov, T, 0, sm);
            if (md.kind != GPC_GPR) cta_gemm<true, false>(T, T, Mr, 1.0, B, T, B, T, 1.0, cov, T, 0, sm);
        }
    }
}

static cudaError_t GPC_KNAME(v_launch_predict)(const void* a, int grid, cudaStream_t s) {
    GPC_KNAME(k_predict)<<<grid, GPC_THREADS, eval_smem_bytes() + sizeof(OptSmem), s>>>(*static_cast<const PredictArgs*>(a));
    return cudaGetLastError();
}
#endif

static size_t GPC_KNAME(v_smem_bytes)(int nmax) {
#if defined(GPC_SMALL_ONLY)
    const int n = nmax < 1 ? 1 : (nmax > GPC_SMALL_NMAX ? GPC_SMALL_NMAX : nmax);
    return GPC_SMALL_ARENA_OFF + GPC_SMALL_VEC_BYTES(n) + small_arena_bytes(n);
#else
    (void)nmax;
    return eval_smem_bytes() + sizeof(OptSmem);
#endif
}

static int GPC_KNAME(v_upload_constants)(const double* z, const double* w) {
    if (cudaMemcpyToSymbol(c_gh_z, z, sizeof(double) * GPC_NGH) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_gh_w, w, sizeof(double) * GPC_NGH) != cudaSuccess) return -1;
    const int sb = (int)GPC_KNAME(v_smem_bytes)(GPC_SMALL_NMAX_OR_0);
    if (cudaFuncSetAttribute(GPC_KNAME(k_fit), cudaFuncAttributeMaxDynamicSharedMemorySize, sb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(GPC_KNAME(k_elbo_grad), cudaFuncAttributeMaxDynamicSharedMemorySize, sb) != cudaSuccess) return -1;
#if !defined(GPC_FUSED_ONLY) && !defined(GPC_SMALL_ONLY)
    if (cudaFuncSetAttribute(GPC_KNAME(k_predict), cudaFuncAttributeMaxDynamicSharedMemorySize, sb) != cudaSuccess) return -1;
#endif
    return 0;
}

static int GPC_KNAME(v_occupancy)(int nmax) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, GPC_KNAME(k_fit), GPC_THREADS, GPC_KNAME(v_smem_bytes)(nmax));
    return per_sm;
}

static cudaError_t GPC_KNAME(v_launch_fit)(const void* a, int grid, int nmax, cudaStream_t s) {
    GPC_KNAME(k_fit)<<<grid, GPC_THREADS, GPC_KNAME(v_smem_bytes)(nmax), s>>>(*static_cast<const FitArgs*>(a));
    return cudaGetLastError();
}

static cudaError_t GPC_KNAME(v_launch_eval)(const void* a, int grid, int nmax, cudaStream_t s) {
    GPC_KNAME(k_elbo_grad)<<<grid, GPC_THREADS, GPC_KNAME(v_smem_bytes)(nmax), s>>>(*static_cast<const EvalArgs*>(a));
    return cudaGetLastError();
}

#if defined(GPC_PHASE_PROF) && !defined(GPC_DENSE_ONLY) && !defined(GPC_SMALL_ONLY)
static int GPC_KNAME(v_read_prof)(unsigned long long* out16) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, g_phase_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out16 + 16, g_opt_prof, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;   // out16: 24 entries
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_prof, z, sizeof(z));
    cudaMemcpyToSymbol(g_opt_prof, z, sizeof(unsigned long long) * 8);
    return 0;
}
#endif

}  // anonymous namespace

// exported (C++ linkage) accessor of this instantiation
VariantOps GPC_KNAME(gpc_variant_ops)() {
    VariantOps o;
    o.upload_constants = GPC_KNAME(v_upload_constants);
    o.smem_bytes = GPC_KNAME(v_smem_bytes);
    o.occupancy = GPC_KNAME(v_occupancy);
    o.launch_fit = GPC_KNAME(v_launch_fit);
    o.launch_eval = GPC_KNAME(v_launch_eval);
    o.read_prof = nullptr;
    o.launch_predict = nullptr;
#if !defined(GPC_FUSED_ONLY) && !defined(GPC_SMALL_ONLY)
    o.launch_predict = GPC_KNAME(v_launch_predict);
#endif
#if defined(GPC_PHASE_PROF) && !defined(GPC_DENSE_ONLY) && !defined(GPC_SMALL_ONLY)
    o.read_prof = GPC_KNAME(v_read_prof);
#endif
    return o;
}
