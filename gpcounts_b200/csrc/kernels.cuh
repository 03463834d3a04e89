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
#include <cuda_runtime.h>
#include "../../include/gpcounts_b200.h"

// launcher table of one instantiation (argument structs are passed as void* because each translation unit has its
// own, anonymous-namespace copy of the device code and of the argument types)
struct VariantOps {
    int (*upload_constants)(const double* z, const double* w);
    size_t (*smem_bytes)();
    int (*occupancy)();
    cudaError_t (*launch_fit)(const void* fit_args, int grid, cudaStream_t s);
    cudaError_t (*launch_eval)(const void* eval_args, int grid, cudaStream_t s);
    int (*read_prof)(unsigned long long* out16);   // profiling build only (GPC_PHASE_PROF), else nullptr
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

static WsLayout make_layout(const gpc_model_t* models, int n_models, int m) {
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
};


// The dense (CtaSmem) and fused (FusedSmem) evaluators overlay the same shared-memory region.
#ifndef GPC_MIN_BLOCKS
#define GPC_MIN_BLOCKS 1
#endif

__host__ __device__ constexpr size_t eval_smem_bytes() {
#ifdef GPC_DENSE_ONLY
    return (sizeof(CtaSmem) + 15) & ~(size_t)15;
#else
    return ((sizeof(CtaSmem) > sizeof(FusedSmem) ? sizeof(CtaSmem) : sizeof(FusedSmem)) + 15) & ~(size_t)15;
#endif
}

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
    int attempt = 0;
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

__global__ void __launch_bounds__(GPC_THREADS, GPC_MIN_BLOCKS) GPC_KNAME(k_fit)(FitArgs a) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
    OptSmem* os = reinterpret_cast<OptSmem*>(g_smem + eval_smem_bytes());
    __shared__ int s_job;
    __shared__ SlotWs s_ws;
    if (threadIdx.x == 0) s_ws = carve(a.ws + (size_t)blockIdx.x * a.lay.slot, a.lay);
    __syncthreads();
    const SlotWs& ws = s_ws;
    const int n_jobs = a.chain ? a.D : a.D * a.n_models;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_job = atomicAdd(a.queue, 1);
        __syncthreads();
        const int job = s_job;
        if (job >= n_jobs) break;
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
};

__global__ void __launch_bounds__(GPC_THREADS, GPC_MIN_BLOCKS) GPC_KNAME(k_elbo_grad)(EvalArgs a) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
    __shared__ int s_job;
    const SlotWs ws = carve(a.ws + (size_t)blockIdx.x * a.lay.slot, a.lay);
    const ModelDev md = a.md;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_job = atomicAdd(a.queue, 1);
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

static size_t GPC_KNAME(v_smem_bytes)() { return eval_smem_bytes() + sizeof(OptSmem); }

static int GPC_KNAME(v_upload_constants)(const double* z, const double* w) {
    if (cudaMemcpyToSymbol(c_gh_z, z, sizeof(double) * GPC_NGH) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_gh_w, w, sizeof(double) * GPC_NGH) != cudaSuccess) return -1;
    const int sb = (int)GPC_KNAME(v_smem_bytes)();
    cudaFuncSetAttribute(GPC_KNAME(k_fit), cudaFuncAttributeMaxDynamicSharedMemorySize, sb);
    cudaFuncSetAttribute(GPC_KNAME(k_elbo_grad), cudaFuncAttributeMaxDynamicSharedMemorySize, sb);
    return 0;
}

static int GPC_KNAME(v_occupancy)() {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, GPC_KNAME(k_fit), GPC_THREADS, GPC_KNAME(v_smem_bytes)());
    return per_sm;
}

static cudaError_t GPC_KNAME(v_launch_fit)(const void* a, int grid, cudaStream_t s) {
    GPC_KNAME(k_fit)<<<grid, GPC_THREADS, GPC_KNAME(v_smem_bytes)(), s>>>(*static_cast<const FitArgs*>(a));
    return cudaGetLastError();
}

static cudaError_t GPC_KNAME(v_launch_eval)(const void* a, int grid, cudaStream_t s) {
    GPC_KNAME(k_elbo_grad)<<<grid, GPC_THREADS, GPC_KNAME(v_smem_bytes)(), s>>>(*static_cast<const EvalArgs*>(a));
    return cudaGetLastError();
}

#if defined(GPC_PHASE_PROF) && !defined(GPC_DENSE_ONLY)
static int GPC_KNAME(v_read_prof)(unsigned long long* out16) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out16, g_phase_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_prof, z, sizeof(z));
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
#if defined(GPC_PHASE_PROF) && !defined(GPC_DENSE_ONLY)
    o.read_prof = GPC_KNAME(v_read_prof);
#endif
    return o;
}
