"""Drop-in ``Fit_GPcounts`` (reference: ``GPcounts/GPcounts_Module.py:37-1063``) whose per-gene fit path runs
as one batched launch of the sm_100a kernels in ``libgpcounts_b200.so``.

Same constructor, methods, keyword arguments, returned DataFrames / dict keys and failure semantics (NaN rows)
as the reference class for the hot path SURVEY.md section 8 scopes:

    Fit_GPcounts(X, Y, scale=None, sparse=False, M=0, safe_mode=False)
        .Infer_trajectory / .One_sample_test / .Two_samples_test (lik_name, transform)
        .Infer_branching_location(cell_labels, bins_num, lik_name, branching_point, transform)
        .calculate_FDR(df) / .qvalue(pv) / .initialize_hyper_parameters(...)

What replaces what: the serial loop of ``run_test`` (:419-424) with a fresh GPflow model and a SciPy
L-BFGS-B run per (gene, model) becomes a single ``gpc_fit`` call over all genes; hyper-parameter
initialisation (:726-806), the restart table and RobustGP inducing-point selection (:655-659, :673-677) are
gene-independent and computed once on the host.  There is no CPU fallback: without the CUDA library and a
GPU the fit methods raise.

Under ``torchrun`` (torch.distributed initialised, world size G) every rank fits genes ``rank::G`` on its own
GPU and the per-gene statistics are all-gathered, so each rank returns the full table.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import host
from ._lib import (GPC_NSTAT, S_ALPHA, S_KM, S_LL, S_LS, S_NFEV, S_NFEV_TOTAL, S_NIT, S_RESTARTS, S_STATUS, S_VAR)

_LIK = {"Negative_binomial": "NB", "Zero_inflated_negative_binomial": "ZINB", "Poisson": "POISSON",
        "Gaussian": "GAUSS"}


class _Value:
    """Stands for a gpflow Parameter in result objects: ``.numpy()`` returns the constrained value."""

    def __init__(self, v):
        self._v = np.asarray(v)

    def numpy(self):
        return self._v

    def __repr__(self):
        return repr(self._v)


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class FittedModel:
    """Light stand-in for the gpflow model kept in ``Fit_GPcounts.model`` (attribute paths used by the
    reference and ``demo_notebooks/helper.py``: ``.kernel.variance``, ``.kernel.lengthscales``,
    ``.kernel.kern.*`` / ``.kernel.xp`` for the branching kernel, ``.likelihood.alpha`` / ``.km``,
    ``.inducing_variable.Z``, ``.data``)."""

    def __init__(self, stats_row, theta, X, y, Z=None, kernel="RBF", xp=None, lik_name=""):
        kern = _NS(variance=_Value(stats_row[S_VAR]), lengthscales=_Value(stats_row[S_LS]))
        if kernel == "BRANCH":
            self.kernel = _NS(kern=kern, xp=xp)
        else:
            self.kernel = kern
        self.likelihood = _NS(alpha=_Value(stats_row[S_ALPHA]), km=_Value(stats_row[S_KM]),
                              variance=_Value(stats_row[S_ALPHA]))
        self.inducing_variable = _NS(Z=_Value(Z)) if Z is not None else None
        self.data = (X, y)
        self.theta = theta
        self.lik_name = lik_name
        self._ll = float(stats_row[S_LL])
        self.n_iter = int(stats_row[S_NIT])
        self.n_feval = int(stats_row[S_NFEV])

    def log_posterior_density(self, *a):
        return _Value(self._ll)


class Fit_GPcounts(object):
    def __init__(self, X=None, Y=None, scale=None, sparse=False, M=0, safe_mode=False, device=None):
        self.safe_mode = safe_mode
        self.folder_name = "GPcounts_models/"
        self.transform = True
        self.sparse = sparse
        self.X = None
        self.M = M
        self.Z = None
        self.conditional_variance = False
        self.Y = None
        self.Y_copy = None
        self.D = None
        self.N = None
        self.scale = scale
        self.genes_name = None
        self.cells_name = None
        self.kernel = None
        self.bic = None
        self.lik_name = None
        self.models_number = None
        self.model_index = None
        self.hyper_parameters = {}
        self.user_hyper_parameters = [None, None, None, None]
        self.model = None
        self.var = None
        self.mean = None
        self.fix = False
        self.lik_alpha = None
        self.lik_km = None
        self.optimize = True
        self.branching = None
        self.xp = -1000.0
        self.y = None
        self.index = None
        self.seed_value = 0
        self.count_fix = 0
        # additions of this implementation
        self.device = device
        self.fit_stats = None           # DataFrame: per (gene, model) status / iterations / hyper-parameters
        self.optimizer_options = {}     # overrides of the L-BFGS-B settings (m, maxiter, maxfun, maxls, ftol, gtol)
        self.last_fit_seconds = None
        self.distributed = True         # shard genes over torch.distributed ranks when a process group is initialised
        self._z_cache = {}
        self.inducing_per_gene = False  # True: attempt 0 uses RobustGP points selected under each gene's own initial
                                        # variance, as the reference does; default: one shared set
        if safe_mode:
            raise NotImplementedError(
                "safe_mode=True (stochastic posterior-predictive local-optimum checks, reference :558-565, "
                ":808-857, :962-1006) is outside the B200 hot path; use safe_mode=False (the reference default).")
        if (X is None) or (Y is None):
            print("TypeError: GPcounts() missing 2 required positional arguments: X and Y")
        else:
            self.set_X_Y(X, Y)

    # ------------------------------------------------------------------------------------------ ingest
    def set_X_Y(self, X, Y):
        self.seed_value = 0
        np.random.seed(self.seed_value)
        if X.shape[0] == Y.shape[1]:
            self.cells_name = list(map(str, list(X.index.values)))
            self.X = X.values.astype(float)
            if len(self.X.shape) > 1:
                self.X = self.X.reshape([-1, self.X.shape[1]])
            else:
                self.X = self.X.reshape([-1, 1])
            if self.sparse:
                if self.M == 0:
                    self.M = int((5 * (len(self.X))) / 100)
                self.conditional_variance = True
            self.genes_name = Y.index.values.tolist()
            self.Y = Y.values
            self.Y_copy = self.Y
            self.D = Y.shape[0]
            self.N = Y.shape[1]
        else:
            print("InvalidArgumentError: Dimension 0 in X shape must be equal to Dimension 1 in Y, but shapes are %d and %d."
                  % (X.shape[0], Y.shape[1]))

    def initialize_hyper_parameters(self, length_scale=None, variance=None, alpha=None, km=None):
        """Reference :726-753 - user overrides of the initial hyper-parameters (None = default rule)."""
        self.user_hyper_parameters = [length_scale, variance, alpha, km]

    # ------------------------------------------------------------------------------------------ tests
    def _truncate(self, transform):
        self.transform = transform
        if transform:
            self.Y = self.Y.astype(int)
            self.Y = self.Y.astype(float)
        self.Y_copy = self.Y

    def Infer_trajectory(self, lik_name="Negative_binomial", transform=True):
        self._truncate(transform)
        return self.run_test(lik_name, 1, range(self.D))

    def One_sample_test(self, lik_name="Negative_binomial", transform=True):
        self._truncate(transform)
        return self.run_test(lik_name, 2, range(self.D))

    def Two_samples_test(self, lik_name="Negative_binomial", transform=True):
        self._truncate(transform)
        return self.run_test(lik_name, 3, range(self.D))

    def Model_selection_test(self, lik_name="Negative_binomial", kernel=None, transform=True):
        raise NotImplementedError("Model_selection_test (Linear / Periodic kernels + BIC, reference :163-229) is "
                                  "outside the B200 hot path (SURVEY.md section 8f, rank 3).")

    def load_predict_models(self, genes_name, test_name, likelihood="Negative_binomial", predict=True):
        raise NotImplementedError("load_predict_models (TF checkpoints + posterior-predictive sampling, reference "
                                  ":859-960) is outside the B200 hot path (SURVEY.md section 8f, rank 2).")

    def kmean_algorithm_inducing_points(self, M=0):
        """Reference :382-390.  As in the reference the k-means centres are stored in ``self.Z`` but, because the
        reference sets the misspelt attribute ``ConditionalVariance``, RobustGP selection still decides Z at fit time."""
        from sklearn.cluster import KMeans
        if M != 0:
            self.M = M
        self.ConditionalVariance = False
        kmeans = KMeans(n_clusters=self.M).fit(self.X)
        self.Z = np.sort(kmeans.cluster_centers_, axis=None).reshape([self.M, 1])

    def calculate_FDR(self, genes_results):
        genes_results["p_value"] = host.p_values(genes_results["log_likelihood_ratio"])
        genes_results["q_value"] = self.qvalue(genes_results["p_value"])
        return genes_results

    def qvalue(self, pv, pi0=None):
        return host.qvalue(np.asarray(pv, dtype=np.float64), pi0)

    # ------------------------------------------------------------------------------------------ engine glue
    def _column_names(self, models_number):
        if models_number == 1:
            return ["Dynamic_model_log_likelihood"]
        if models_number == 2:
            return ["Dynamic_model_log_likelihood", "Constant_model_log_likelihood", "log_likelihood_ratio"]
        return ["Shared_log_likelihood", "model_1_log_likelihood", "model_2_log_likelihood", "log_likelihood_ratio"]

    def _device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("gpcounts_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        if self.device is not None:
            return torch.device(self.device)
        return torch.device("cuda", torch.cuda.current_device())

    def _opt(self):
        from . import engine
        return engine.default_opt(**self.optimizer_options)

    def _restart(self, Xm, kernel, fixed=None):
        """(10, 4) restart table of a model; Gaussian noise / branching kernel columns follow the reference."""
        tab, _ = host.restart_table(Xm)
        tab = tab.copy()
        if self.lik_name == "Gaussian":
            tab[:, 2] = 1.0                      # gpflow Gaussian likelihood starts at variance 1 on every attempt
        if kernel == "BRANCH":                   # kernel re-created as RBF() or frozen on every attempt (:606-617)
            tab[:, 0] = 1.0 if fixed is None else fixed[1]
            tab[:, 1] = 1.0 if fixed is None else fixed[0]
        return tab

    def _hyp0(self, Ym, Xm, kernel="RBF", fixed=None, y_for_var=None):
        """(D, 4) initial (variance, lengthscale, alpha, km) of one model, reference :726-753."""
        u = self.user_hyper_parameters
        D = Ym.shape[0]
        h = np.zeros((D, 4))
        yv = Ym if y_for_var is None else y_for_var
        h[:, 0] = host.default_variance(yv, self.lik_name, self.transform) if u[1] is None else u[1]
        h[:, 1] = host.default_lengthscale(Xm) if u[0] is None else u[0]
        h[:, 2] = 1.0 if u[2] is None else u[2]
        h[:, 3] = 35.0 if u[3] is None else u[3]
        if self.lik_name == "Gaussian":
            h[:, 2] = 1.0
        if kernel == "BRANCH":
            h[:, 0], h[:, 1] = (1.0, 1.0) if fixed is None else (fixed[0], fixed[1])
        return h

    def _make_model(self, kernel, Xm, n_off=0, handoff=False, xp=-1000.0, train=(True, True, True, True),
                    fixed=None, restarts=True):
        from . import engine
        lik = _LIK[self.lik_name]
        kind = {(False, False): "VGP", (True, False): "SVGP", (False, True): "GPR", (True, True): "SGPR"}[
            (bool(self.sparse), lik == "GAUSS")]
        Z = None
        if self.sparse:
            ls0 = host.default_lengthscale(Xm) if self.user_hyper_parameters[0] is None else self.user_hyper_parameters[0]
            # gene-independent: one selection per (X, M, ls0), shared by the models of a test
            key = (Xm.shape, hash(Xm.tobytes()), int(self.M), float(ls0), bool(restarts))
            if key not in self._z_cache:
                self._z_cache = {key: host.inducing_sets(Xm, self.M, ls0, with_restarts=restarts)}
            Z = self._z_cache[key]
            self.Z = Z[0]
        return engine.Model(kind, kernel, lik, Xm, Z=Z, n_off=n_off, train=train, handoff_alpha=handoff, xp=xp,
                            restart=self._restart(Xm, kernel, fixed) if restarts else None, device=self._device())

    def _run(self, models, Yfit, hyp0, chain, genes_index):
        """Shard genes over ranks (if torch.distributed is up), run gpc_fit, gather the statistics."""
        from . import engine, sharding
        dev = self._device()
        genes_index = np.asarray(list(genes_index), dtype=np.int64)
        rank, world = sharding.rank_world() if self.distributed else (0, 1)
        mine = sharding.shard_indices(len(genes_index), rank, world)
        gsel = genes_index[mine]
        K = len(models)
        scale = None
        if self.scale is not None and _LIK[self.lik_name] == "NB":
            sc = self.scale.values if hasattr(self.scale, "values") else np.asarray(self.scale)
            scale = np.ascontiguousarray(sc[:, gsel].T, dtype=np.float64)           # (D, N): column per gene (:629)
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record(torch.cuda.current_stream(dev))
        if len(gsel):
            if self.sparse and self.inducing_per_gene:
                from . import inducing
                for k, m in enumerate(models):
                    if m.kernel == "RBF" and self.user_hyper_parameters[1] is None:
                        Xm = m.X.cpu().numpy()
                        m.Zgene = inducing.select_per_gene(Xm, m.M, float(hyp0[0, k, 1]), hyp0[gsel, k, 0], dev).contiguous()
            stats, theta = engine.fit(models, np.ascontiguousarray(Yfit[gsel]), scale, hyp0[gsel], chain=chain,
                                      opt=self._opt(), return_theta=True)
        else:
            stats = torch.zeros(0, K, GPC_NSTAT, dtype=torch.float64, device=dev)
            theta = None
        t1.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        self.last_fit_seconds = t0.elapsed_time(t1) * 1e-3
        stats_all = sharding.gather_rows(stats, len(genes_index), rank, world)
        self._last_theta = theta            # local shard only
        self._last_local_pos = mine         # positions (within genes_index) of the local shard
        return stats_all.cpu().numpy()

    def _record_stats(self, stats, genes_index, names):
        """Per (gene, model) fit statistics as a DataFrame (built column-wise: no per-row Python objects)."""
        stats = np.asarray(stats)
        G, K = stats.shape[0], len(names)
        flat = stats[:, :K].reshape(G * K, -1)
        gene_names = np.asarray(self.genes_name, dtype=object)[np.asarray(list(genes_index), dtype=np.int64)]
        self.fit_stats = pd.DataFrame({
            "gene": np.repeat(gene_names, K), "model": np.tile(np.asarray(names, dtype=object), G),
            "log_likelihood": flat[:, S_LL], "status": flat[:, S_STATUS].astype(np.int64),
            "n_iter": flat[:, S_NIT].astype(np.int64), "n_feval": flat[:, S_NFEV].astype(np.int64),
            "n_feval_total": flat[:, S_NFEV_TOTAL].astype(np.int64), "restarts": flat[:, S_RESTARTS].astype(np.int64),
            "variance": flat[:, S_VAR], "lengthscale": flat[:, S_LS], "alpha": flat[:, S_ALPHA], "km": flat[:, S_KM]})

    def _set_last_model(self, models, stats, k, Xm, y, kernel="RBF", xp=None):
        """``self.model`` = model of the last gene fitted, as the reference leaves it after the loop (:264-265)."""
        theta = None
        if self._last_theta is not None and len(self._last_local_pos) and self._last_local_pos[-1] == len(stats) - 1:
            theta = self._last_theta[-1, k, :models[k].P].cpu().numpy()
        row = stats[-1, k]
        if np.isnan(row[S_LL]):
            self.model = np.nan                                  # reference :533-535
        else:
            self.model = FittedModel(row, theta, Xm, y, Z=self.Z if self.sparse else None, kernel=kernel, xp=xp,
                                     lik_name=self.lik_name)

    # ------------------------------------------------------------------------------------------ run_test
    def run_test(self, lik_name, models_number, genes_index, branching=False):
        """Reference :393-428 with the gene loop replaced by one batched device fit."""
        self.Y = self.Y_copy
        self.models_number = models_number
        self.lik_name = lik_name
        self.optimize = True
        if lik_name not in _LIK:
            raise ValueError("unknown likelihood %r" % (lik_name,))
        if branching:
            return self._run_branching_stage(genes_index)
        column_name = self._column_names(models_number)
        genes_index = list(genes_index)
        Y = np.ascontiguousarray(self.Y, dtype=np.float64)
        Yfit = np.log(Y + 1) if (lik_name == "Gaussian" and self.transform) else Y      # :650-652
        X = self.X
        if models_number == 1:
            models = [self._make_model("RBF", X)]
            hyp0 = self._hyp0(Y, X)[:, None, :]
            stats = self._run(models, Yfit, hyp0, True, genes_index)
            out = stats[:, :, S_LL]
            names = ["dynamic"]
        elif models_number == 2:
            handoff = lik_name == "Negative_binomial"                                  # :785-787
            models = [self._make_model("RBF", X), self._make_model("CONST", X, handoff=handoff)]
            # Gaussian quirk: the constant model's initial variance is computed from the already
            # log-transformed y of the dynamic fit (:652 runs before :738 of the next model)
            yv2 = Yfit if (lik_name == "Gaussian" and self.transform) else None
            hyp0 = np.stack([self._hyp0(Y, X), self._hyp0(Y, X, kernel="CONST", y_for_var=yv2)], axis=1)
            stats = self._run(models, Yfit, hyp0, True, genes_index)
            ll = stats[:, :, S_LL].copy()
            bad = np.isnan(ll[:, 0]) | np.isnan(ll[:, 1])
            ll[bad, 1] = np.nan                                                         # :455-457
            out = np.c_[ll, ll[:, 0] - ll[:, 1]]
            names = ["dynamic", "constant"]
        else:
            h = int(self.N / 2)                                                         # :469, :477
            X1, X2 = X[:h], X[h:]
            models = [self._make_model("RBF", X), self._make_model("RBF", X1, n_off=0),
                      self._make_model("RBF", X2, n_off=h)]
            hyp0 = np.stack([self._hyp0(Y, X), self._hyp0(Y[:, :h], X1), self._hyp0(Y[:, h:], X2)], axis=1)
            stats = self._run(models, Yfit, hyp0, False, genes_index)
            ll = stats[:, :, S_LL]
            out = np.c_[ll, (ll[:, 1] + ll[:, 2]) - ll[:, 0]]                          # NaN propagates (:486-495)
            names = ["shared", "model_1", "model_2"]
        self._record_stats(stats, genes_index, names)
        self.index = genes_index[-1] if genes_index else None
        if genes_index:
            self.y = Y[genes_index[-1]].reshape([-1, 1])
            k_last = len(models) - 1
            if models_number == 3:          # the reference leaves the model of the second half behind (:477-484)
                self._set_last_model(models, stats, k_last, X[int(self.N / 2):], self.y[int(self.N / 2):])
            else:
                self._set_last_model(models, stats, k_last, X, self.y)
        return pd.DataFrame(out, index=[self.genes_name[g] for g in genes_index], columns=column_name)

    # ------------------------------------------------------------------------------------------ branching
    def Infer_branching_location(self, cell_labels, bins_num=50, lik_name="Negative_binomial", branching_point=-1000,
                                 transform=True):
        """Reference :243-333: free fit of the branching kernel at ``branching_point``, then one frozen-kernel fit
        per candidate branching time, posterior over the candidates and log Bayes factor."""
        self._truncate(transform)
        if self.sparse:
            raise NotImplementedError("sparse inference with the branching kernel is not part of the reference workflow")
        cell_labels = np.array(cell_labels)
        self.X = np.c_[self.X, cell_labels[:, None]]
        self.branching = True
        self.xp = branching_point
        genes_index = range(self.D)
        self.fix = False
        self.run_test(lik_name, 1, genes_index, branching=True)
        self.branching_kernel_var = self.model.kernel.kern.variance.numpy()
        self.branching_kernel_ls = self.model.kernel.kern.lengthscales.numpy()
        return self.infer_branching(lik_name, bins_num)

    def _run_branching_stage(self, genes_index):
        """Free fit (fix=False): BranchKernel(RBF(1, 1), xp) for every gene; ``self.model`` = last gene (:262-265)."""
        genes_index = list(genes_index)
        Y = np.ascontiguousarray(self.Y, dtype=np.float64)
        Yfit = np.log(Y + 1) if (self.lik_name == "Gaussian" and self.transform) else Y
        models = [self._make_model("BRANCH", self.X, xp=float(self.xp))]
        hyp0 = self._hyp0(Y, self.X, kernel="BRANCH")[:, None, :]
        stats = self._run(models, Yfit, hyp0, True, genes_index)
        self._record_stats(stats, genes_index, ["branching_free"])
        self.y = Y[genes_index[-1]].reshape([-1, 1])
        self._set_last_model(models, stats, 0, self.X, self.y, kernel="BRANCH", xp=self.xp)
        if isinstance(self.model, float):
            raise RuntimeError("branching kernel fit failed for the last gene")
        return pd.DataFrame(stats[:, :, S_LL], index=[self.genes_name[g] for g in genes_index],
                            columns=self._column_names(1))

    def infer_branching(self, lik_name, bins_num):
        testTimes = np.linspace(min(self.X[:, 0]), max(self.X[:, 0]), bins_num, endpoint=True)      # :271-273
        self.fix = True
        X = self.X
        g_last = self.D - 1                 # the reference keeps only the last gene's models (:289-290)
        Y = np.ascontiguousarray(self.Y, dtype=np.float64)
        Yfit = np.log(Y + 1) if (lik_name == "Gaussian" and self.transform) else Y
        fixed = (float(self.branching_kernel_var), float(self.branching_kernel_ls))
        models, Xs = [], []
        for i in range(bins_num):                                                                  # :279-288
            Xi = X.copy()
            Xi[np.where(Xi[:, 0] <= testTimes[i]), 1] = 1
            Xs.append(Xi)
            models.append(self._make_model("BRANCH", Xi, xp=float(testTimes[i]), train=(False, False, True, True),
                                           fixed=fixed))
        hyp0 = np.stack([self._hyp0(Y, X, kernel="BRANCH", fixed=fixed)] * bins_num, axis=1)
        stats = self._run(models, Yfit, hyp0, False, [g_last])
        self._record_stats(stats, [g_last], ["bin_%d" % i for i in range(bins_num)])
        log_ll = stats[0, :, S_LL].copy()
        p = self.CalculateBranchingEvidence({"loglik": log_ll}, testTimes)
        ll = p["posteriorBranching"]
        iMAP = int(np.argmax(ll))
        self.xp = testTimes[iMAP]
        self.y = Y[g_last].reshape([-1, 1])
        theta = self._last_theta[0, iMAP, :models[iMAP].P].cpu().numpy() if self._last_theta is not None else None
        self.model = FittedModel(stats[0, iMAP], theta, Xs[iMAP], self.y, kernel="BRANCH", xp=testTimes[iMAP],
                                 lik_name=lik_name)
        Xnew = np.linspace(min(self.X[:, 0]), max(self.X[:, 0]), 100).reshape(-1)[:, None]
        self.branching = False
        return {
            "geneName": self.genes_name,
            "branching_probability": ll,
            "branching_location": self.model.kernel.xp,
            # posterior-predictive mean / variance need Monte-Carlo sampling (:305-317): outside the hot path
            "mean": None,
            "variance": None,
            "Xnew": Xnew,
            "test_times": testTimes,
            "MAP_model": self.model,
            "loglik": log_ll,
            "logBayesFactor": p["logBayesFactor"],
            "likelihood": self.lik_name,
        }

    def CalculateBranchingEvidence(self, d, Bsearch):
        return host.branching_evidence(d["loglik"], Bsearch)
