// C-ABI entry points and the persistent per-gene kernels of gpcounts_b200 (sm_100a).
//
// Execution model: one CTA fits (or evaluates) one gene-model at a time.  A launch starts `n_slots`
// persistent CTAs (a multiple of the 148 SMs); each owns one workspace slot in HBM/L2 and pulls jobs
// from an atomic queue until it is empty, so genes whose optimisation takes 20 iterations and genes that
// take 2000 share the machine without any inter-CTA synchronisation and without host round trips.
// Genes never communicate: sharding over GPUs is a partition of the job list.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/gpcounts_b200.h"
#define GPC_VARIANT t256
#include "kernels.cuh"


static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int set_err(int code, const char* fmt, const char* a = "") {
    snprintf(g_err, sizeof(g_err), fmt, a);
    return code;
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
#define GPC_API_SMALL_NMAX 48     // = GPC_SMALL_NMAX of vgp_small.cuh (compiled into gpc_small.cu only)

// TMA descriptor of a gene-major FP64 matrix [D][ld]: box = 64 columns x 1 row, i.e. one column tile of one gene's counts
// (or scale factors); columns past `ld` read as zeros.  Needs a 16-byte aligned base and row pitch (ld even); returns false
// otherwise (the kernels then read the rows with ordinary loads).  cuTensorMapEncodeTiled is resolved through the runtime
// (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_row_tile_map(CUtensorMap* tm, const double* base, int64_t D, int64_t ld) {
    static EncodeTiledFn fn = nullptr;
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!fn) {
            void* p = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
            fn = (EncodeTiledFn)p;
        }
    }
    if (((uintptr_t)base & 15) || (ld & 1) || D < 1 || ld < 1) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)D};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {64, 1}, estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Kernel instantiations (kernels.cuh): [0] every evaluator; [1] dense evaluators only, two CTAs per SM;
// [2] fused SVGP evaluator only; [3] on-chip small-N VGP evaluator, 64-thread CTAs, 8-9 per SM.
#define GPC_NVARIANTS 4
static VariantOps g_ops[GPC_NVARIANTS];
VariantOps gpc_variant_ops_d128();     // gpc_dense.cu
VariantOps gpc_variant_ops_f256();     // gpc_fused.cu
VariantOps gpc_variant_ops_s64();      // gpc_small.cu

static bool host_use_fused(const gpc_model_t& md) {
    return md.kind == GPC_SVGP && md.M <= 64 && md.d <= 2 && md.kernel != GPC_BRANCH;
}

static bool host_use_small(const gpc_model_t& md) {
    return md.kind == GPC_VGP && md.N <= GPC_API_SMALL_NMAX;
}

static int models_nmax(const gpc_model_t* models, int n_models) {
    int n = 1;
    for (int k = 0; k < n_models; ++k) n = std::max(n, (int)models[k].N);
    return n;
}

static int pick_variant(const gpc_model_t* models, int n_models, int forced = -1) {
    bool any_fused = false, all_fused = true, all_small = true;
    for (int k = 0; k < n_models; ++k) {
        const bool f = host_use_fused(models[k]);
        any_fused = any_fused || f;
        all_fused = all_fused && f;
        all_small = all_small && host_use_small(models[k]);
    }
    if (forced == 0) return 0;       // instantiation [0] holds every evaluator (A/B measurements through gpc_opt_t.variant)
    if (forced == 1 && !any_fused) return 1;
    if (all_small) return 3;
    return all_fused ? 2 : (any_fused ? 0 : 1);
}

static int ensure_init() {
    static bool done[64] = {false};
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_err(GPC_E_LAUNCH, "no CUDA device%s");
    if (dev < 64 && done[dev]) return 0;
    double z[GPC_NGH], w[GPC_NGH];
    const double s2 = 1.4142135623730951, sp = 1.7724538509055159;
    for (int i = 0; i < GPC_NGH; ++i) { z[i] = H_GH_X[i] * s2; w[i] = H_GH_W[i] / sp; }
    g_ops[0] = gpc_variant_ops_t256();
    g_ops[1] = gpc_variant_ops_d128();
    g_ops[2] = gpc_variant_ops_f256();
    g_ops[3] = gpc_variant_ops_s64();
    for (int v = 0; v < GPC_NVARIANTS; ++v)
        if (g_ops[v].upload_constants(z, w) != 0)
            return set_err(GPC_E_LAUNCH, "constant upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFuncSetAttribute(k_dbg_linalg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes());
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
    if ((md.kind == GPC_SVGP || md.kind == GPC_SGPR) && md.restart && md.n_zsets != 1 && md.n_zsets != 1 + GPC_MAX_RESTART)
        return set_err(GPC_E_ARG, "a model with a restart table needs n_zsets = 1 or 1 + GPC_MAX_RESTART inducing sets%s");
    if (md.kernel == GPC_BRANCH && md.d != 2) return set_err(GPC_E_ARG, "BRANCH kernel needs d = 2 ([t, label])%s");
    return 0;
}

extern "C" {

void gpc_default_opt(gpc_opt_t* opt) {
    opt->m = 10; opt->maxiter = 5000; opt->maxfun = 15000; opt->maxls = 20;
    opt->ftol = 2.220446049250313e-09; opt->gtol = 1e-5;
    opt->variant = -1; opt->trace_cap = 0; opt->trace_gene = -1; opt->trace_model = -1; opt->trace = nullptr;
    opt->attempt0 = nullptr;
}

int64_t gpc_param_count(const gpc_model_t* model) {
    const int64_t mv = model_mv(*model);
    return GPC_NHYP + mv + mv * (mv + 1) / 2;
}

int32_t gpc_default_slots(const gpc_model_t* models, int32_t n_models) {
    int dev = 0, sms = 148, per_sm = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int v = models ? pick_variant(models, n_models) : 0;
    if (ensure_init() == 0) per_sm = g_ops[v].occupancy(models ? models_nmax(models, n_models) : 1);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > (v == 3 ? 12 : 2)) per_sm = (v == 3 ? 12 : 2);
    return sms * per_sm;
}

size_t gpc_workspace_bytes(const gpc_model_t* models, int32_t n_models, const gpc_opt_t* opt, int32_t n_slots) {
    const WsLayout L = make_layout(models, n_models, opt ? opt->m : 0, pick_variant(models, n_models, opt ? opt->variant : -1) == 3);
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
    const WsLayout L = make_layout(model, 1, 0, pick_variant(model, 1) == 3);
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
    a.use_tma = make_row_tile_map(&a.tmY, Y, D, ldy) && (!scale || make_row_tile_map(&a.tmS, scale, D, ldy));
    cudaMemsetAsync(a.queue, 0, 256, s);
    const int grid = n_slots < D ? n_slots : D;
    cudaError_t e = g_ops[pick_variant(model, 1)].launch_eval(&a, grid, model->N, s);
    ++g_launches;
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
    const int variant = pick_variant(models, n_models, opt->variant);
    const WsLayout L = make_layout(models, n_models, opt->m, variant == 3);
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
    a.nmax = models_nmax(models, n_models);
    a.trace = opt->trace; a.trace_cap = opt->trace_cap; a.trace_gene = opt->trace_gene; a.trace_model = opt->trace_model;
    a.use_tma = make_row_tile_map(&a.tmY, Y, D, ldy) && (!scale || make_row_tile_map(&a.tmS, scale, D, ldy));
    cudaMemsetAsync(a.queue, 0, 256, s);
    const int n_jobs = chain ? D : D * n_models;
    const int grid = n_slots < n_jobs ? n_slots : n_jobs;
    cudaError_t e = g_ops[variant].launch_fit(&a, grid, models_nmax(models, n_models), s);
    ++g_launches;
    if (e != cudaSuccess) return set_err(GPC_E_LAUNCH, "gpc_fit launch: %s", cudaGetErrorString(e));
    return 0;
}

size_t gpc_dbg_linalg_scratch_bytes(int32_t batch, int32_t n) {
    const size_t nblk = (size_t)(n + GPC_BLK - 1) / GPC_BLK;
    return ((size_t)nblk * GPC_BLK * GPC_BLK + (size_t)n * n) * (size_t)batch * sizeof(double);
}

static void predict_layout(const gpc_model_t& md, int T, PredictArgs& a) {
    const bool sparse = md.kind == GPC_SVGP || md.kind == GPC_SGPR;
    const size_t mr = sparse ? md.M : md.N;
    a.sq = pad8(mr * mr);
    a.dinv = (size_t)((mr + GPC_BLK - 1) / GPC_BLK) * GPC_BLK * GPC_BLK;
    a.rect = pad8(mr * (size_t)T);
    a.vec = pad8(mr);
    a.slot = 3 * a.sq + a.dinv + 2 * a.rect + a.vec;
}

size_t gpc_predict_workspace_bytes(const gpc_model_t* model, int32_t T, int32_t n_slots) {
    PredictArgs a;
    predict_layout(*model, T, a);
    return 256 + (size_t)n_slots * a.slot * sizeof(double);
}

int gpc_predict(const gpc_model_t* model, int32_t D, const double* Y, int64_t ldy, const double* theta, int64_t ld_theta,
                int32_t zset, const double* Xnew, int32_t T, double* out_mean, double* out_var, double* out_cov,
                void* workspace, size_t workspace_bytes, int32_t n_slots, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!model || D < 0 || !theta || !Xnew || T < 1 || !out_mean || !out_var || !workspace || n_slots < 1)
        return set_err(GPC_E_ARG, "gpc_predict: null or invalid argument%s");
    if ((rc = check_model(*model))) return rc;
    if (model->kind == GPC_SGPR) return set_err(GPC_E_UNSUPPORTED, "gpc_predict: SGPR prediction is not provided%s");
    if (model->kind == GPC_GPR && !Y) return set_err(GPC_E_ARG, "gpc_predict: GPR prediction needs the data Y%s");
    if (ld_theta < gpc_param_count(model)) return set_err(GPC_E_ARG, "gpc_predict: ld_theta smaller than the parameter count%s");
    if (zset < 0 || (model->Z && zset >= model->n_zsets)) return set_err(GPC_E_ARG, "gpc_predict: zset out of range%s");
    PredictArgs a;
    predict_layout(*model, T, a);
    if (workspace_bytes < 256 + (size_t)n_slots * a.slot * sizeof(double))
        return set_err(GPC_E_WORKSPACE, "gpc_predict: workspace too small%s");
    if (D == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    a.md = to_dev(*model);
    a.D = D; a.T = T; a.zset = zset; a.Y = Y; a.theta = theta; a.Xnew = Xnew; a.ldy = ldy; a.ld_theta = ld_theta;
    a.mean = out_mean; a.var = out_var; a.cov = out_cov;
    a.queue = (int*)workspace;
    a.ws = (double*)((char*)workspace + 256);
    cudaMemsetAsync(a.queue, 0, 256, s);
    const int grid = n_slots < D ? n_slots : D;
    cudaError_t e = g_ops[1].launch_predict(&a, grid, s);
    ++g_launches;
    if (e != cudaSuccess) return set_err(GPC_E_LAUNCH, "gpc_predict launch: %s", cudaGetErrorString(e));
    return 0;
}

int gpc_dbg_linalg(int32_t op, int32_t batch, int32_t n, int32_t nrhs, double* A, double* B, int32_t* status,
                   void* scratch, size_t scratch_bytes, void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    if (op < 0 || op > 4 || batch < 1 || n < 1 || !A || !status || !scratch)
        return set_err(GPC_E_ARG, "gpc_dbg_linalg: bad argument%s");
    if (scratch_bytes < gpc_dbg_linalg_scratch_bytes(batch, n))
        return set_err(GPC_E_WORKSPACE, "gpc_dbg_linalg: scratch too small%s");
    k_dbg_linalg<<<batch, GPC_THREADS, smem_bytes(), (cudaStream_t)stream>>>(op, n, nrhs, A, B, status, (double*)scratch);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(GPC_E_LAUNCH, "gpc_dbg_linalg: %s", cudaGetErrorString(e));
    return 0;
}

int64_t gpc_launch_count(void) { return g_launches.load(); }

#ifdef GPC_PHASE_PROF
// profiling build only (tools/phase_prof.py): read and reset the phase timers of the fused evaluator of one instantiation
int gpc_dbg_phase_prof(int variant, unsigned long long* out16) {
    if (ensure_init() != 0 || variant < 0 || variant >= GPC_NVARIANTS || !g_ops[variant].read_prof) return -1;
    return g_ops[variant].read_prof(out16);
}
#endif
const char* gpc_last_error_string(void) { return g_err; }
const char* gpc_version(void) { return "gpcounts_b200 0.1 (sm_100a)"; }

}  // extern "C"
