/*
 * gpcounts_b200 - C ABI of the B200-native per-gene GP fitting hot path.
 *
 * This is the drop-in boundary for the path SURVEY.md section 8 scopes: everything that happens in the
 * reference between `Fit_GPcounts.run_test` picking up a gene (GPcounts/GPcounts_Module.py:419-424)
 * and `fit_model` returning its log-likelihood (GPcounts_Module.py:506-537), for all genes at once.
 * The reference has no native interface (it is Python over GPflow/TensorFlow/SciPy); the entry points
 * below are what a ctypes binding inside `Fit_GPcounts` binds instead of the per-gene GPflow model +
 * SciPy L-BFGS-B loop (see INTEGRATION.md for the stub).
 *
 * Conventions: plain C, device pointers + sizes + an opaque CUDA stream handle (cudaStream_t passed as
 * void*); the caller owns every buffer including the workspace.  State kept by the library between calls: the
 * thread-local last-error string, a mutex-guarded per-device "constants uploaded" latch (Gauss-Hermite table) and an
 * atomic launch counter - nothing else (no environment variables, no allocations, no host synchronisation).
 * Every function returns 0 on success and a negative GPC_E_* code
 * on argument / launch errors.  Per-gene numerical failures are not errors: they are reported in the
 * per-(gene, model) statistics rows exactly like the reference reports them (NaN log-likelihood).
 *
 * All floating point data is FP64 (`gpflow.config.set_default_float(np.float64)`, GPcounts_Module.py:800).
 */
#ifndef GPCOUNTS_B200_H
#define GPCOUNTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* model kinds  - gpflow.models.{VGP,SVGP,GPR,SGPR}, GPcounts_Module.py:661-686 */
enum { GPC_VGP = 0, GPC_SVGP = 1, GPC_GPR = 2, GPC_SGPR = 3 };
/* kernels      - gpflow.kernels.{RBF,Constant} GPcounts_Module.py:574,601-604; BranchKernel branchingKernel.py:22-74 */
enum { GPC_RBF = 0, GPC_CONST = 1, GPC_BRANCH = 2 };
/* likelihoods  - GPcounts_Module.py:622-648, NegativeBinomialLikelihood.py:11-89 */
enum { GPC_NB = 0, GPC_ZINB = 1, GPC_POISSON = 2, GPC_GAUSS = 3 };

/* per-(gene, model) status, mirrors the reference's failure semantics */
enum {
    GPC_ST_OK = 0,          /* optimiser converged (res.success)                                  */
    GPC_ST_CHOL = 1,        /* Gram not positive definite  (tf.errors.InvalidArgumentError, :547) */
    GPC_ST_OPT = 2,         /* optimiser success == False  (:699)                                 */
    GPC_ST_SKIPPED = 3      /* not fitted because an earlier model of the chain failed (:442)     */
};

enum { GPC_E_ARG = -1, GPC_E_LAUNCH = -2, GPC_E_WORKSPACE = -3, GPC_E_UNSUPPORTED = -4 };

#define GPC_NHYP 4          /* theta[0..3] = softplus^-1 of (variance, lengthscale, alpha|noise, km) */
#define GPC_NSTAT 16        /* doubles per (gene, model) statistics row, see GPC_S_* */
#define GPC_MAX_RESTART 10  /* random restarts per fit, GPcounts_Module.py:549, :700 */

/* layout of one statistics row */
enum {
    GPC_S_LL = 0,        /* ELBO / log marginal at the optimum; NaN if the fit failed (:533-535) */
    GPC_S_STATUS = 1,    /* GPC_ST_*                                                             */
    GPC_S_NIT = 2,       /* L-BFGS-B iterations of the successful (or last) attempt             */
    GPC_S_NFEV = 3,      /* ELBO+gradient evaluations of that attempt                            */
    GPC_S_RESTARTS = 4,  /* count_fix: number of random restarts used                            */
    GPC_S_VAR = 5, GPC_S_LS = 6, GPC_S_ALPHA = 7, GPC_S_KM = 8,   /* fitted hyper-parameters     */
    GPC_S_NFEV_TOTAL = 9,/* evaluations over all attempts (FLOP accounting)                      */
    GPC_S_GNORM = 10,    /* ||grad||_inf at the returned point                                   */
    GPC_S_TASK = 11,     /* optimiser exit: 0 pgtol, 1 ftol, 2 maxiter/maxfun, 3 abnormal line search, 4 Cholesky */
    GPC_S_EVAL_CYCLES = 12, /* SM clock cycles spent in ELBO+gradient evaluations (all attempts)            */
    GPC_S_FIT_CYCLES = 13   /* SM clock cycles of the whole fit of this (gene, model)                       */
};

/*
 * One model of the per-gene chain (e.g. One_sample_test = {RBF, Constant}; Two_samples_test = three RBF
 * models on index ranges; branching grid = one BRANCH model per candidate branching time).
 * X / Z are shared by all genes.
 */
typedef struct gpc_model {
    int32_t kind;          /* GPC_VGP ...                                                        */
    int32_t kernel;        /* GPC_RBF ...                                                        */
    int32_t lik;           /* GPC_NB ...                                                         */
    int32_t N;             /* data points of this model                                          */
    int32_t M;             /* inducing points (SVGP / SGPR), else 0                              */
    int32_t d;             /* input dimension (BRANCH: 2 = [t, label])                           */
    int32_t n_off;         /* first column of Y / scale this model reads (Two_samples_test halves, :469,:477) */
    int32_t train_mask;    /* bit i set = hyper-parameter i trainable (set_trainable, :608-614)  */
    int32_t n_zsets;       /* number of inducing sets behind Z: 1, or 1 + GPC_MAX_RESTART (restart r uses set r) */
    int32_t handoff_alpha; /* 1: initial alpha := fitted alpha of the previous model of the chain on the first attempt (:785-787) */
    double xp;             /* branching point (BranchKernel.xp)                                  */
    double jitter;         /* gpflow default_jitter() = 1e-6                                     */
    const double* X;       /* device, N x d row-major                                            */
    const double* Z;       /* device, n_zsets x M x d, or NULL                                   */
    const double* Zgene;   /* device, D x M x d per-gene inducing points for the first attempt, or NULL (then set 0 of Z).
                              RobustGP's greedy selection runs under the gene's own initial variance (:673-677) and its
                              near-ties are decided by rounding, so attempt 0 can have gene-specific Z.       */
    const double* restart; /* device, GPC_MAX_RESTART x 4 (ls, var, alpha, km) restart table (:769-779), or NULL = no restarts */
} gpc_model_t;

typedef struct gpc_opt {
    int32_t m;             /* L-BFGS memory (SciPy maxcor = 10)        */
    int32_t maxiter;       /* 5000, GPcounts_Module.py:696            */
    int32_t maxfun;        /* 15000 (SciPy default)                   */
    int32_t maxls;         /* 20 (SciPy default)                      */
    double ftol;           /* 2.220446049250313e-09 (SciPy default)   */
    double gtol;           /* 1e-5 (SciPy default)                    */
    /* debug / measurement knobs (gpc_default_opt leaves them off) */
    int32_t variant;       /* -1: the library picks the kernel instantiation; >= 0 forces one (A/B measurements) */
    int32_t trace_cap;     /* capacity of `trace` in records of 8 doubles                                        */
    int32_t trace_gene;    /* record (loss, step, slope, max|g|, theta[0..3]) of every evaluation of attempt 0   */
    int32_t trace_model;   /*   of this (gene, model) ...                                                        */
    double* trace;         /*   ... into this device buffer; NULL = off                                          */
    /* host-driven restarts (safe_mode, GPcounts_Module.py:518-525, :962-1006): device int32 [D x n_models] or NULL.  A fit
     * whose entry is c > 0 starts at restart c (hyper-parameters of restart-table row c - 1, inducing set c), exactly as
     * if attempts 0 .. c-1 had failed; its GPC_S_RESTARTS counts from there. */
    const int32_t* attempt0;
} gpc_opt_t;

/* Default optimiser settings = gpflow.optimizers.Scipy / SciPy L-BFGS-B defaults with maxiter=5000. */
void gpc_default_opt(gpc_opt_t* opt);

/* Length of the parameter vector of a model: 4 + Mv + Mv(Mv+1)/2, Mv = N (VGP) | M (SVGP) | 0 (GPR, SGPR). */
int64_t gpc_param_count(const gpc_model_t* model);

/* Number of CTA slots (concurrent gene fits) the library would launch on the current device for these models. */
int32_t gpc_default_slots(const gpc_model_t* models, int32_t n_models);

/* Bytes of caller-provided device workspace for n_slots concurrent fits over the given models. */
size_t gpc_workspace_bytes(const gpc_model_t* models, int32_t n_models, const gpc_opt_t* opt, int32_t n_slots);

/*
 * Batched fixed-theta evaluation: ELBO and dELBO/dtheta of `model` for D genes.
 * Replaces: one `training_loss` + autodiff evaluation of a gpflow model (GPcounts_Module.py:680, 686;
 * gpflow VGP.elbo / SVGP.elbo), i.e. the callable SciPy drives at :692-697.
 *   Y      D x ldy   counts (gene-major); the model reads columns [n_off, n_off+N)
 *   scale  D x ldy   per-spot scale factors or NULL (GPcounts_Module.py:627-636)
 *   theta  D x P     unconstrained parameters (layout in include header / DESIGN.md)
 *   zset   which inducing set of model->Z to use
 *   out_elbo  D      ELBO (NaN where the Gram matrix is not positive definite)
 *   out_grad  D x P  gradient of the ELBO (frozen / unused hyper-parameter slots are zero)
 *   out_status D     int32: GPC_ST_OK or GPC_ST_CHOL
 */
int gpc_elbo_grad(const gpc_model_t* model, int32_t D, const double* Y, const double* scale, int64_t ldy,
                  const double* theta, int32_t zset, double* out_elbo, double* out_grad, int32_t* out_status,
                  void* workspace, size_t workspace_bytes, int32_t n_slots, void* stream);

/*
 * Batched fit: for every gene run the chain of `n_models` models through L-BFGS-B with the reference's
 * restart protocol.  Replaces `fit_single_gene` / `fit_model` / `fit_GP` / `fit_GP_with_likelihood`
 * (GPcounts_Module.py:431-707) for all genes of `run_test` (:419-424).
 *   chain      1: models are fitted in order within one job and a failed model skips the rest
 *                 (One_sample_test :441-459); 0: every (gene, model) is an independent job
 *   hyp0       D x n_models x 4  initial (variance, lengthscale, alpha|noise, km) per gene and model
 *                 (init_hyper_parameters :726-753); lengthscale is ignored for GPC_CONST
 *   out_stats  D x n_models x GPC_NSTAT
 *   out_theta  D x n_models x ld_theta fitted parameter vectors, or NULL
 */
int gpc_fit(const gpc_model_t* models, int32_t n_models, int32_t chain, const gpc_opt_t* opt, int32_t D,
            const double* Y, const double* scale, int64_t ldy, const double* hyp0,
            double* out_stats, double* out_theta, int64_t ld_theta,
            void* workspace, size_t workspace_bytes, int32_t n_slots, void* stream);

/*
 * Batched posterior prediction of the latent function at new inputs from fitted parameter vectors.
 * Replaces `model.predict_f(Xnew)` (and the mean / covariance behind `predict_y`, `predict_f_samples`) of the fitted gpflow
 * models in load_predict_models (GPcounts_Module.py:930-945), infer_branching (:305-317) and the safe-mode check (:962-980).
 *   theta     D x ld_theta  fitted parameter vectors (gpc_fit out_theta rows of this model)
 *   Y         D x ldy       data; only read for GPC_GPR (the Gaussian posterior mean depends on y), may be NULL otherwise
 *   Xnew      T x d         new inputs (BRANCH: [t, label])
 *   out_mean  D x T,  out_var  D x T   mean and variance of f(Xnew);  out_cov  D x T x T full covariance or NULL
 * Rows of genes whose Gram matrix is not positive definite are NaN.  GPC_SGPR is not supported (GPC_E_UNSUPPORTED).
 */
size_t gpc_predict_workspace_bytes(const gpc_model_t* model, int32_t T, int32_t n_slots);
int gpc_predict(const gpc_model_t* model, int32_t D, const double* Y, int64_t ldy, const double* theta, int64_t ld_theta,
                int32_t zset, const double* Xnew, int32_t T, double* out_mean, double* out_var, double* out_cov,
                void* workspace, size_t workspace_bytes, int32_t n_slots, void* stream);

/*
 * Test hooks for the CTA-level dense primitives (batched over `batch` matrices, one CTA each).
 *   op 0: potrf      A (n x n) -> lower Cholesky factor in place; status[b] = 0 | 1
 *   op 1: trsm_l     B (n x nrhs) <- L^-1 B        (A holds L)
 *   op 2: trsm_lt    B (n x nrhs) <- L^-T B
 *   op 3: trsm_r     B (nrhs x n) <- B L^-1
 *   op 4: chol_bwd   B (n x n, lower = Lbar) <- L^-T Phi(L^T Lbar) L^-1   (unsymmetrised)
 */
int gpc_dbg_linalg(int32_t op, int32_t batch, int32_t n, int32_t nrhs, double* A, double* B, int32_t* status,
                   void* scratch, size_t scratch_bytes, void* stream);
/* Bytes of caller-provided device scratch gpc_dbg_linalg needs (stream-ordered, no allocation inside the library). */
size_t gpc_dbg_linalg_scratch_bytes(int32_t batch, int32_t n);

/* Number of kernel launches issued by this library since load (bench.py `gpu_launches`). */
int64_t gpc_launch_count(void);

const char* gpc_last_error_string(void);
const char* gpc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPCOUNTS_B200_H */
