// C-ABI entry points and the persistent per-gene kernels of gpcounts_b200 (sm_100a).
//
// Execution model: one CTA fits (or evaluates) one gene-model at a time.  A launch starts `n_slots`
// persistent CTAs (a multiple of the 148 SMs); each owns one workspace slot in HBM/L2 and pulls jobs
// from an atomic queue until it is empty, so genes whose optimisation takes 20 iterations and genes that
// take 2000 share the machine without any inter-CTA synchronisation and without host round trips.
// Genes never communicate: sharding over GPUs is a partition of the job list.
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../include/gpcounts_b200.h"
#include "lbfgsb.cuh"

static thread_local char g_err[512] = "";
static double* g_trace = nullptr;
static int g_trace_cap = 0, g_trace_gene = 0, g_trace_model = 0;
static long long g_launches = 0;

static int set_err(int code, const char* fmt, const char* a = "") {
    snprintf(g_err, sizeof(g_err), fmt, a);
    return code;
}

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

extern __shared__ __align__(16) unsigned char g_smem[];

// The dense (CtaSmem) and fused (FusedSmem) evaluators overlay the same shared-memory region.
__host__ __device__ constexpr size_t eval_smem_bytes() {
    return ((sizeof(CtaSmem) > sizeof(FusedSmem) ? sizeof(CtaSmem) : sizeof(FusedSmem)) + 15) & ~(size_t)15;
}

__device__ __forceinline__ double softplus_inv_d(double p) { return p + log(-expm1(-p)); }

// Fit one (gene, model); returns true when the fit succeeded. prev_alpha < 0: no hand-off.
__device__ bool fit_one(const FitArgs& a, int g, int k, double prev_alpha, const SlotWs& ws, CtaSmem* sm, OptSmem* os,
                        double& alpha_out) {
    const ModelDev md = a.models[k];
    const int tid = threadIdx.x;
    const double* y = a.Y + (size_t)g * a.ldy + md.n_off;
    const double* scale = a.scale ? a.scale + (size_t)g * a.ldy + md.n_off : nullptr;
    const double* h0 = a.hyp0 + ((size_t)g * a.n_models + k) * 4;
    double* st = a.stats + ((size_t)g * a.n_models + k) * GPC_NSTAT;
    const int max_restart = md.restart ? GPC_MAX_RESTART : 0;
    int nfev_total = 0;
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
        // theta0: hyper-parameters through softplus^-1, q_mu = 0, q_sqrt = I
        for (int i = tid; i < md.P; i += GPC_THREADS) ws.x[i] = 0.0;
        __syncthreads();
        if (tid == 0) {
            ws.x[0] = softplus_inv_d(var);
            ws.x[1] = softplus_inv_d(ls);
            ws.x[2] = softplus_inv_d(md.lik == GPC_GAUSS ? alpha - GPC_GAUSS_VAR_LOWER : alpha);
            ws.x[3] = softplus_inv_d(km);
        }
        for (int i = tid; i < md.Mv; i += GPC_THREADS) ws.x[GPC_NHYP + md.Mv + tri_index(i, i)] = 1.0;
        __syncthreads();
        const double* Z = md.Z ? md.Z + (size_t)(md.n_zsets > 1 ? attempt : 0) * md.M * md.d : nullptr;
        if (attempt == 0 && md.Zgene) Z = md.Zgene + (size_t)g * md.M * md.d;
        OptTrace tr{nullptr, 0, 0};
        if (a.trace && g == a.trace_gene && k == a.trace_model && attempt == 0) { tr.buf = a.trace; tr.cap = a.trace_cap; }
        r = lbfgsb_minimize(md, Z, y, scale, a.opt, ws, sm, os, &tr);
        nfev_total += r.nfev;
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
        for (int q = 12; q < GPC_NSTAT; ++q) st[q] = 0.0;
    }
    if (a.theta_out) {
        double* out = a.theta_out + ((size_t)g * a.n_models + k) * a.ld_theta;
        for (int i = tid; i < md.P; i += GPC_THREADS) out[i] = ws.x[i];
    }
    __syncthreads();
    return ok;
}

__global__ void __launch_bounds__(GPC_THREADS) k_fit(FitArgs a) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
    OptSmem* os = reinterpret_cast<OptSmem*>(g_smem + eval_smem_bytes());
    __shared__ int s_job;
    const SlotWs ws = carve(a.ws + (size_t)blockIdx.x * a.lay.slot, a.lay);
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

__global__ void __launch_bounds__(GPC_THREADS) k_elbo_grad(EvalArgs a) {
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

__global__ void __launch_bounds__(GPC_THREADS) k_dbg_linalg(int op, int n, int nrhs, double* A, double* B,
                                                           int* status, double* scratch) {
    CtaSmem* sm = reinterpret_cast<CtaSmem*>(g_smem);
    const int b = blockIdx.x;
    const int nblk = (n + GPC_BLK - 1) / GPC_BLK;
    double* Ab = A + (size_t)b * n * n;
    double* dinv = scratch + (size_t)b * ((size_t)nblk * GPC_BLK * GPC_BLK + (size_t)n * n);
    double* T = dinv + (size_t)nblk * GPC_BLK * GPC_BLK;
    if (op == 0) {
        const bool ok = cta_potrf(n, Ab, n, dinv, sm);
        if (threadIdx.x == 0) status[b] = ok ? 0 : 1;
        return;
    }
    // ops 1-4 expect A = SPD matrix; factor first
    const bool ok = cta_potrf(n, Ab, n, dinv, sm);
    if (threadIdx.x == 0) status[b] = ok ? 0 : 1;
    if (!ok) return;
    if (op == 1) cta_trsm_left(n, nrhs, Ab, n, dinv, B + (size_t)b * n * nrhs, nrhs, sm);
    else if (op == 2) cta_trsm_left_t(n, nrhs, Ab, n, dinv, B + (size_t)b * n * nrhs, nrhs, sm);
    else if (op == 3) cta_trsm_right(nrhs, n, Ab, n, dinv, B + (size_t)b * n * nrhs, n, sm);
    else if (op == 4) cta_chol_backward(n, Ab, n, dinv, B + (size_t)b * n * n, n, T, n, sm);
}

// ---------------------------------------------------------------------------------------------- host side
static const double H_GH_X[GPC_NGH] = {
    -5.3874808900112328, -4.6036824495507442, -3.9447640401156252, -3.3478545673832163, -2.7888060584281305,
    -2.2549740020892757, -1.7385377121165861, -1.2340762153953231, -0.73747372854539439, -0.24534070830090124,
    0.24534070830090124, 0.73747372854539439, 1.2340762153953231, 1.7385377121165861, 2.2549740020892757,
    2.7888060584281305, 3.3478545673832163, 3.9447640401156252, 4.6036824495507442, 5.3874808900112328};
static const double H_GH_W[GPC_NGH] = {
    2.2293936455341505e-13, 4.3993409922731809e-10, 1.0860693707692815e-07, 7.8025564785320632e-06,
    2.2833863601635403e-04, 3.2437733422378567e-03, 2.4810520887463643e-02, 1.0901720602002331e-01,
    2.8667550536283415e-01, 4.6224366960061009e-01, 4.6224366960061009e-01, 2.8667550536283415e-01,
    1.0901720602002331e-01, 2.4810520887463643e-02, 3.2437733422378567e-03, 2.2833863601635403e-04,
    7.8025564785320632e-06, 1.0860693707692815e-07, 4.3993409922731809e-10, 2.2293936455341505e-13};

static size_t smem_bytes() { return eval_smem_bytes() + sizeof(OptSmem); }

static int ensure_init() {
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_err(GPC_E_LAUNCH, "no CUDA device%s");
    if (dev < 64 && done[dev]) return 0;
    double z[GPC_NGH], w[GPC_NGH];
    const double s2 = 1.4142135623730951, sp = 1.7724538509055159;
    for (int i = 0; i < GPC_NGH; ++i) { z[i] = H_GH_X[i] * s2; w[i] = H_GH_W[i] / sp; }
    if (cudaMemcpyToSymbol(c_gh_z, z, sizeof(z)) != cudaSuccess || cudaMemcpyToSymbol(c_gh_w, w, sizeof(w)) != cudaSuccess)
        return set_err(GPC_E_LAUNCH, "constant upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    const int sb = (int)smem_bytes();
    cudaFuncSetAttribute(k_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, sb);
    cudaFuncSetAttribute(k_elbo_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, sb);
    cudaFuncSetAttribute(k_dbg_linalg, cudaFuncAttributeMaxDynamicSharedMemorySize, sb);
    if (dev < 64) done[dev] = true;
    return 0;
}

static int check_model(const gpc_model_t& md) {
    if (md.kind < GPC_VGP || md.kind > GPC_SGPR) return set_err(GPC_E_ARG, "bad model kind%s");
    if (md.kernel < GPC_RBF || md.kernel > GPC_BRANCH) return set_err(GPC_E_ARG, "bad kernel%s");
    if (md.lik < GPC_NB || md.lik > GPC_GAUSS) return set_err(GPC_E_ARG, "bad likelihood%s");
    if ((md.kind == GPC_GPR || md.kind == GPC_SGPR) != (md.lik == GPC_GAUSS))
        return set_err(GPC_E_ARG, "Gaussian likelihood goes with GPR/SGPR and only with them%s");
    if (md.N <= 0 || md.d <= 0 || !md.X) return set_err(GPC_E_ARG, "model needs N > 0, d > 0 and X%s");
    if ((md.kind == GPC_SVGP || md.kind == GPC_SGPR) && (md.M <= 0 || !md.Z || md.n_zsets < 1))
        return set_err(GPC_E_ARG, "sparse model needs M > 0 and Z%s");
    if (md.kernel == GPC_BRANCH && md.d != 2) return set_err(GPC_E_ARG, "BRANCH kernel needs d = 2 ([t, label])%s");
    return 0;
}

extern "C" {

void gpc_default_opt(gpc_opt_t* opt) {
    opt->m = 10; opt->maxiter = 5000; opt->maxfun = 15000; opt->maxls = 20;
    opt->ftol = 2.220446049250313e-09; opt->gtol = 1e-5;
}

int64_t gpc_param_count(const gpc_model_t* model) {
    const int64_t mv = model_mv(*model);
    return GPC_NHYP + mv + mv * (mv + 1) / 2;
}

int32_t gpc_default_slots(const gpc_model_t* models, int32_t n_models) {
    (void)models; (void)n_models;
    int dev = 0, sms = 148, per_sm = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (ensure_init() == 0)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fit, GPC_THREADS, smem_bytes());
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    return sms * per_sm;
}

size_t gpc_workspace_bytes(const gpc_model_t* models, int32_t n_models, const gpc_opt_t* opt, int32_t n_slots) {
    const WsLayout L = make_layout(models, n_models, opt ? opt->m : 0);
    return 256 + pad8(sizeof(ModelDev) * (size_t)n_models) + (size_t)n_slots * L.slot * sizeof(double);
}

int gpc_elbo_grad(const gpc_model_t* model, int32_t D, const double* Y, const double* scale, int64_t ldy,
                  const double* theta, int32_t zset, double* out_elbo, double* out_grad, int32_t* out_status,
                  void* workspace, size_t workspace_bytes, int32_t n_slots, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!model || D < 0 || !Y || !theta || !out_elbo || !out_grad || !out_status || !workspace || n_slots < 1)
        return set_err(GPC_E_ARG, "gpc_elbo_grad: null or invalid argument%s");
    if ((rc = check_model(*model))) return rc;
    if (zset < 0 || (model->Z && zset >= model->n_zsets)) return set_err(GPC_E_ARG, "gpc_elbo_grad: zset out of range%s");
    const WsLayout L = make_layout(model, 1, 0);
    if (workspace_bytes < 256 + (size_t)n_slots * L.slot * sizeof(double))
        return set_err(GPC_E_WORKSPACE, "gpc_elbo_grad: workspace too small%s");
    if (D == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    EvalArgs a;
    a.md = to_dev(*model);
    a.D = D; a.zset = zset; a.Y = Y; a.scale = scale; a.theta = theta; a.ldy = ldy;
    a.elbo = out_elbo; a.grad = out_grad; a.status = out_status;
    a.queue = (int*)workspace;
    a.ws = (double*)((char*)workspace + 256);
    a.lay = L;
    cudaMemsetAsync(a.queue, 0, 256, s);
    const int grid = n_slots < D ? n_slots : D;
    k_elbo_grad<<<grid, GPC_THREADS, smem_bytes(), s>>>(a);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(GPC_E_LAUNCH, "gpc_elbo_grad launch: %s", cudaGetErrorString(e));
    return 0;
}

int gpc_fit(const gpc_model_t* models, int32_t n_models, int32_t chain, const gpc_opt_t* opt, int32_t D,
            const double* Y, const double* scale, int64_t ldy, const double* hyp0, double* out_stats,
            double* out_theta, int64_t ld_theta, void* workspace, size_t workspace_bytes, int32_t n_slots,
            void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!models || n_models < 1 || n_models > GPC_MAX_MODELS_CONST || !opt || D < 0 || !Y || !hyp0 || !out_stats ||
        !workspace || n_slots < 1)
        return set_err(GPC_E_ARG, "gpc_fit: null or invalid argument%s");
    if (opt->m < 1 || opt->m > GPC_MMAX) return set_err(GPC_E_ARG, "gpc_fit: L-BFGS memory must be 1..10%s");
    int64_t pmax = 0;
    for (int k = 0; k < n_models; ++k) {
        if ((rc = check_model(models[k]))) return rc;
        pmax = std::max(pmax, gpc_param_count(&models[k]));
    }
    if (out_theta && ld_theta < pmax) return set_err(GPC_E_ARG, "gpc_fit: ld_theta smaller than the parameter count%s");
    const WsLayout L = make_layout(models, n_models, opt->m);
    const size_t hdr = 256 + pad8(sizeof(ModelDev) * (size_t)n_models);
    if (workspace_bytes < hdr + (size_t)n_slots * L.slot * sizeof(double))
        return set_err(GPC_E_WORKSPACE, "gpc_fit: workspace too small%s");
    if (D == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    ModelDev h_models[GPC_MAX_MODELS_CONST];
    for (int k = 0; k < n_models; ++k) h_models[k] = to_dev(models[k]);
    ModelDev* d_models = (ModelDev*)((char*)workspace + 256);
    // pageable host source: cudaMemcpyAsync stages it before returning, so the stack array may go away
    cudaMemcpyAsync(d_models, h_models, sizeof(ModelDev) * n_models, cudaMemcpyHostToDevice, s);
    FitArgs a;
    a.models = d_models; a.n_models = n_models; a.chain = chain; a.D = D; a.opt = *opt;
    a.Y = Y; a.scale = scale; a.hyp0 = hyp0; a.ldy = ldy; a.ld_theta = ld_theta;
    a.stats = out_stats; a.theta_out = out_theta;
    a.queue = (int*)workspace;
    a.ws = (double*)((char*)workspace + hdr);
    a.lay = L;
    a.trace = g_trace; a.trace_cap = g_trace_cap; a.trace_gene = g_trace_gene; a.trace_model = g_trace_model;
    cudaMemsetAsync(a.queue, 0, 256, s);
    const int n_jobs = chain ? D : D * n_models;
    const int grid = n_slots < n_jobs ? n_slots : n_jobs;
    k_fit<<<grid, GPC_THREADS, smem_bytes(), s>>>(a);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(GPC_E_LAUNCH, "gpc_fit launch: %s", cudaGetErrorString(e));
    return 0;
}

int gpc_dbg_linalg(int32_t op, int32_t batch, int32_t n, int32_t nrhs, double* A, double* B, int32_t* status,
                   void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (op < 0 || op > 4 || batch < 1 || n < 1 || !A || !status) return set_err(GPC_E_ARG, "gpc_dbg_linalg: bad argument%s");
    const int nblk = (n + GPC_BLK - 1) / GPC_BLK;
    double* scratch = nullptr;
    const size_t per = (size_t)nblk * GPC_BLK * GPC_BLK + (size_t)n * n;
    if (cudaMalloc(&scratch, per * batch * sizeof(double)) != cudaSuccess) return set_err(GPC_E_WORKSPACE, "gpc_dbg_linalg: cudaMalloc failed%s");
    k_dbg_linalg<<<batch, GPC_THREADS, smem_bytes(), (cudaStream_t)stream>>>(op, n, nrhs, A, B, status, scratch);
    ++g_launches;
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(scratch);
    if (e != cudaSuccess) return set_err(GPC_E_LAUNCH, "gpc_dbg_linalg: %s", cudaGetErrorString(e));
    return 0;
}

void gpc_dbg_set_trace(double* buf, int32_t cap, int32_t gene, int32_t model) {
    g_trace = buf; g_trace_cap = cap; g_trace_gene = gene; g_trace_model = model;
}

int64_t gpc_launch_count(void) { return g_launches; }
const char* gpc_last_error_string(void) { return g_err; }
const char* gpc_version(void) { return "gpcounts_b200 0.1 (sm_100a)"; }

}  // extern "C"
