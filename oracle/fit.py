"""Oracle: CPU restatement of the per-gene fit driver of GPcounts (test infrastructure).

Follows ``GPcounts/GPcounts_Module.py``:
  run_test :393-428, fit_single_gene :431-503, fit_model :506-537, fit_GP :539-566,
  fit_GP_with_likelihood :569-707, initialize_hyper_parameters / init_hyper_parameters :726-806,
  Infer_branching_location / infer_branching / CalculateBranchingEvidence :243-366.
The optimiser of record is the installed SciPy's L-BFGS-B with SciPy defaults and ``maxiter=5000``
(``gpflow.optimizers.Scipy().minimize(..., options=dict(maxiter=5000))``, :692-697).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.optimize

from . import gp_math as gm
from .gp_math import CholeskyFailure, ModelSpec
from .robustgp import select_Z

LIK_CODE = {"Negative_binomial": "NB", "Zero_inflated_negative_binomial": "ZINB",
            "Poisson": "POISSON", "Gaussian": "GAUSS"}


@dataclass
class FitResult:
    log_likelihood: float = np.nan
    theta: np.ndarray | None = None
    spec: ModelSpec | None = None
    nit: int = 0
    nfev: int = 0
    restarts: int = 0
    ok: bool = False
    message: str = ""

    def hyper(self):
        p = np.logaddexp(self.theta[:4], 0.0)
        return dict(var=p[0], ls=p[1], alpha=p[2], km=p[3])


PERTURB = 0.0     # test knob: relative size of a random perturbation of the start point (stability probes)


def minimize_lbfgsb(spec, theta0, y, scale, options=None):
    """One ``gpflow.optimizers.Scipy().minimize`` call on the trainable variables."""
    idx = spec.trainable_index()
    theta = np.array(theta0, dtype=np.float64)
    if PERTURB:
        theta[idx] += PERTURB * np.random.default_rng(12345).standard_normal(idx.size)

    def fun(x):
        theta[idx] = x
        f, g = gm.elbo_and_grad(theta, spec, y, scale)
        return -f, -g[idx]

    opts = dict(maxiter=5000)
    if options:
        opts.update(options)
    res = scipy.optimize.minimize(fun, theta[idx].copy(), jac=True, method="L-BFGS-B", options=opts)
    theta[idx] = res.x
    return res, theta


class OracleFit:
    """Numpy-array analogue of ``Fit_GPcounts``: X (N,d), Y (D,N), scale (N,D) or None."""

    def __init__(self, X, Y, scale=None, sparse=False, M=0, alpha_policy="handoff", lbfgs_options=None,
                 inducing_variance=None):
        self.X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
        self.Y = np.asarray(Y, dtype=np.float64)
        self.scale = None if scale is None else np.asarray(scale, dtype=np.float64)
        self.sparse = sparse
        self.M = M
        if sparse and M == 0:
            self.M = int((5 * len(self.X)) / 100)                      # :121-125
        self.user = [None, None, None, None]                           # :75-80
        self.transform = True
        self.alpha_policy = alpha_policy    # "handoff" = Fit_GPcounts (:785-787); "one" = notebook goldens
        self.lbfgs_options = lbfgs_options
        # RobustGP's greedy selection ties to the last bit while the inducing set is sparse, and the winner then
        # depends on how (var0_g + 1e-12) - e^2 rounds (and on the BLAS summation order of its dot product): the
        # reference's Z is not reproducible across machines at that level.  None = the gene's own initial
        # variance as in the reference (:673-677); a number = use it for attempt 0 (fixes the tie-break so two
        # implementations can be compared on identical Z).
        self.inducing_variance = inducing_variance
        self.branching = False
        self.fix = False
        self.xp = -1000.0
        self.branch_var = self.branch_ls = None

    def initialize_hyper_parameters(self, length_scale=None, variance=None, alpha=None, km=None):
        self.user = [length_scale, variance, alpha, km]

    # ---- one fit with the restart protocol ---------------------------------------------------
    def fit_model(self, X, y_raw, scale, lik, model_index, models_number, lik_alpha=None, y_for_var=None):
        """fit_model/fit_GP/init_hyper_parameters/fit_GP_with_likelihood for one (gene, model)."""
        rng_range = np.max(X) - np.min(X)
        constant = (model_index == 2 and models_number == 2)
        count_fix, seed = 0, 0
        np.random.seed(seed)                                            # :758-761
        reset = False
        yv = y_raw if y_for_var is None else y_for_var
        while True:
            ls = (5 * rng_range) / 100 if self.user[0] is None else self.user[0]          # :729-732
            if self.user[1] is None:                                                     # :734-741
                if lik == "GAUSS" and not self.transform:
                    var = np.mean(yv + 1 ** 2)
                else:
                    var = np.mean(np.log(yv + 1) ** 2)
            else:
                var = self.user[1]
            alpha = 1.0 if self.user[2] is None else self.user[2]
            km = 35.0 if self.user[3] is None else self.user[3]
            if reset:                                                                    # :769-779
                count_fix += 1
                seed += 1
                np.random.seed(seed)
                ls = np.random.uniform((0.25 * rng_range) / 100, (30.0 * rng_range) / 100)
                var = np.random.uniform(0.0, 10.0)
                alpha = np.random.uniform(0.0, 10.0)
                km = np.random.uniform(0.0, 100.0)
            kernel = "RBF"
            if constant:                                                                 # :782-787
                kernel = "CONST"
                if count_fix == 0 and lik == "NB" and self.alpha_policy == "handoff" and lik_alpha is not None:
                    alpha = lik_alpha
            train = (True, True, True, True)
            xp = -1000.0
            if self.branching:                                                           # :606-617
                kernel = "BRANCH"
                xp = self.xp
                if self.fix:
                    var, ls = self.branch_var, self.branch_ls
                    train = (False, False, True, True)
                else:
                    var, ls = 1.0, 1.0
            y = y_raw
            if lik == "GAUSS":
                if self.transform:
                    y = np.log(y_raw + 1)                                                # :650-652
                alpha = 1.0                         # gpflow Gaussian likelihood default variance
            Z = None
            if self.sparse:                                                              # :655-659, 673-677
                zvar = var if (self.inducing_variance is None or count_fix > 0) else self.inducing_variance
                Z = select_Z(X, self.M, kernel, zvar, ls)
            kind = {(False, False): "VGP", (True, False): "SVGP",
                    (False, True): "GPR", (True, True): "SGPR"}[(self.sparse, lik == "GAUSS")]
            spec = ModelSpec(kind=kind, kernel=kernel, lik=lik, X=X, Z=Z, xp=xp, train=train)
            theta0 = gm.initial_theta(spec, var, ls, alpha, km)
            out = FitResult(spec=spec, restarts=count_fix)
            try:
                # gpflow.Parameter rejects a non-finite unconstrained value (variance 0 of an all-zero gene gives
                # softplus^-1(0) = -inf) with tf.errors.InvalidArgumentError, caught at :547 like a Cholesky failure
                used = [True, kernel != "CONST", lik in ("NB", "ZINB", "GAUSS"), lik == "ZINB"]
                if not all(np.isfinite(theta0[i]) for i in range(4) if used[i]):
                    raise CholeskyFailure("non-finite unconstrained initial value")
                res, theta = minimize_lbfgsb(spec, theta0, y, scale, self.lbfgs_options)
                out.nit, out.nfev, out.message = res.nit, res.nfev, str(res.message)
                if res.success:                                                          # :699
                    out.ok = True
                    out.theta = theta
                    out.log_likelihood = -float(res.fun)                                 # :510-515
                    return out, y
            except CholeskyFailure as e:                                                 # :547
                out.message = "cholesky: " + str(e)
            if count_fix < 10:                                                           # :549-550, :700-702
                reset = True
                continue
            return out, y                                                                # NaN, :533-535

    # ---- tests ----------------------------------------------------------------------------------
    def _prep(self, transform):
        self.transform = transform
        Y = self.Y
        if transform:
            Y = Y.astype(int).astype(float)                                              # :153-155
        return Y

    def _scale_col(self, g, sl=slice(None)):
        return None if self.scale is None else self.scale[sl, g]                        # :629

    def infer_trajectory(self, lik_name="Negative_binomial", transform=True, genes=None):
        Y = self._prep(transform)
        lik = LIK_CODE[lik_name]
        rows = []
        for g in (range(Y.shape[0]) if genes is None else genes):
            r1, _ = self.fit_model(self.X, Y[g], self._scale_col(g), lik, 1, 1)
            rows.append(dict(gene=g, fits=[r1], ll=[r1.log_likelihood]))
        return rows

    def one_sample_test(self, lik_name="Negative_binomial", transform=True, genes=None):
        Y = self._prep(transform)
        lik = LIK_CODE[lik_name]
        rows = []
        for g in (range(Y.shape[0]) if genes is None else genes):
            s = self._scale_col(g)
            r1, y1 = self.fit_model(self.X, Y[g], s, lik, 1, 2)
            ll1, ll2, llr, r2 = r1.log_likelihood, np.nan, np.nan, None
            if not np.isnan(ll1):                                                        # :442-453
                lik_alpha = r1.hyper()["alpha"] if lik in ("NB", "ZINB") else None
                # Gaussian quirk: self.y still holds log(y+1) when the constant model's
                # variance is initialised (:652 then :738 before :802)
                yv = y1 if lik == "GAUSS" else None
                r2, _ = self.fit_model(self.X, Y[g], s, lik, 2, 2, lik_alpha=lik_alpha, y_for_var=yv)
                ll2 = r2.log_likelihood
                if not np.isnan(ll2):
                    llr = ll1 - ll2
            if np.isnan(ll1) or np.isnan(ll2):                                           # :455-457
                ll2, llr = np.nan, np.nan
            rows.append(dict(gene=g, fits=[r1, r2], ll=[ll1, ll2, llr]))
        return rows

    def two_samples_test(self, lik_name="Negative_binomial", transform=True, genes=None):
        Y = self._prep(transform)
        lik = LIK_CODE[lik_name]
        N = self.X.shape[0]
        h = int(N / 2)                                                                   # :469, :477
        rows = []
        for g in (range(Y.shape[0]) if genes is None else genes):
            r1, _ = self.fit_model(self.X, Y[g], self._scale_col(g), lik, 1, 3)
            r2, _ = self.fit_model(self.X[:h], Y[g, :h], self._scale_col(g, slice(0, h)), lik, 2, 3)
            r3, _ = self.fit_model(self.X[h:], Y[g, h:], self._scale_col(g, slice(h, None)), lik, 3, 3)
            lls = [r.log_likelihood for r in (r1, r2, r3)]
            llr = np.nan if any(np.isnan(lls)) else (lls[1] + lls[2]) - lls[0]           # :486-495
            rows.append(dict(gene=g, fits=[r1, r2, r3], ll=lls + [llr]))
        return rows

    def infer_branching_location(self, cell_labels, bins_num=50, lik_name="Negative_binomial",
                                 branching_point=-1000, transform=True, gene=-1):
        """:243-366 for one gene (the reference keeps only the last gene's model)."""
        Y = self._prep(transform)
        lik = LIK_CODE[lik_name]
        g = range(Y.shape[0])[gene]
        Xb = np.c_[self.X[:, :1], np.asarray(cell_labels, dtype=np.float64)[:, None]]   # :256-257
        self.branching, self.fix, self.xp = True, False, float(branching_point)
        r0, _ = self.fit_model(Xb, Y[g], self._scale_col(g), lik, 1, 1)                  # :262
        h = r0.hyper()
        self.branch_var, self.branch_ls = h["var"], h["ls"]                              # :264-265
        testTimes = np.linspace(min(Xb[:, 0]), max(Xb[:, 0]), bins_num, endpoint=True)   # :271-273
        self.fix = True
        fits, ll = [], np.zeros(bins_num)
        for i in range(bins_num):                                                        # :279-290
            self.xp = testTimes[i]
            Xi = Xb.copy()
            Xi[np.where(Xi[:, 0] <= testTimes[i]), 1] = 1
            ri, _ = self.fit_model(Xi, Y[g], self._scale_col(g), lik, 1, 1)
            fits.append(ri)
            ll[i] = ri.log_likelihood
        self.branching, self.fix = False, False
        ev = calculate_branching_evidence(ll, testTimes)
        return dict(free_fit=r0, fits=fits, loglik=ll, test_times=testTimes,
                    branching_probability=ev["posteriorBranching"], logBayesFactor=ev["logBayesFactor"],
                    branching_location=testTimes[int(np.argmax(ev["posteriorBranching"]))])


def calculate_branching_evidence(loglik, Bsearch):
    """CalculateBranchingEvidence, GPcounts_Module.py:335-366."""
    o = np.asarray(loglik, dtype=np.float64)
    pn = np.exp(o - np.max(o))
    p = pn / pn.sum()
    Nb = o.size - 1
    if Nb != len(Bsearch) - 1:
        raise NameError("Passed in wrong length of Bsearch is %g- should be %g" % (len(Bsearch), Nb))
    obj = o[:-1]
    illmax = np.argmax(obj)
    llmax = obj[illmax]
    lratiostable = (llmax + np.log(1 + np.exp(obj[np.arange(obj.size) != illmax] - llmax).sum())
                    - o[-1] - np.log(Nb))
    return {"posteriorBranching": p, "logBayesFactor": lratiostable}
