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


def _expand(a0, pending, D, genes_index):
    """attempt0 table indexed by gene position in Y (what _run slices with the global gene indices)."""
    full = np.zeros((D, a0.shape[1]), dtype=np.int32)
    full[genes_index[pending]] = a0[pending]
    return full


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
        self._fitted = {}               # (test_name, lik_name) -> fitted parameter vectors on the device + model metadata
        self.persist_models = False     # True: also write the fitted parameters under folder_name (reference :528-531)
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
        """Reference :163-229: the one-sample test once per entry of ``["Linear", "Periodic", "RBF"]``, BIC per gene and
        kernel (``-2 LL_dynamic + K log N``) and BIC-based model probabilities (SpatialDE recipe), returned as the merged
        table of the best kernel per gene.

        What the reference actually fits: its kernel switch compares lower-case names (``"linear" in self.kernel``,
        ``"periodic" in self.kernel``, :576-598) with the capitalised names of this list, so all three passes build the RBF
        kernel with K = 4 - three identical fits, equal BICs, probabilities 1/3 each.  This method reproduces that
        behaviour (the ``kernel`` argument is ignored there too); a Linear / Periodic Gram builder is not part of the
        reference's effective path."""
        from scipy.special import logsumexp
        self._truncate(transform)
        ker_list = ["Linear", "Periodic", "RBF"]
        frames = []
        base = None
        for word in ker_list:
            self.kernel = word
            self.K = 4                                    # the RBF branch of :592-598 for every name (see above)
            if base is None:                              # the three passes are the same computation: fit once
                base = self.run_test(lik_name, 2, range(self.D))
            results = base.copy()
            results["BIC"] = -2 * results["Dynamic_model_log_likelihood"] + self.K * np.log(self.X.shape[0])
            results["Gene"] = self.genes_name
            results["Model"] = word
            results["p_value"] = host.p_values(results["log_likelihood_ratio"])
            results["q_value"] = self.qvalue(results["p_value"])
            frames.append(results)
        selection_results = pd.concat(frames, ignore_index=True)
        bic_values = -selection_results.pivot_table(values="BIC", index="Gene", columns="Model")
        with np.errstate(over="ignore", under="ignore"):
            log_v = logsumexp(bic_values.values, axis=1)
            model_prob = np.exp(bic_values.sub(log_v, axis=0)).add_suffix("_probability")
        best = selection_results.groupby("Gene")["BIC"].transform("min") == selection_results["BIC"]
        out = selection_results[best].join(model_prob, on="Gene")
        return out.reset_index(drop=True)

    # ------------------------------------------------------------------------------------------ fitted models
    _TEST_MODELS = {"Infer_trajectory": 1, "One_sample_test": 2, "Two_samples_test": 3}

    def get_file_name(self, test_name=None, lik_name=None):
        """Reference :709-723 names one TF checkpoint per (gene, model): <lik>_[sparse_][tst_]<gene>_model_<k>.  Here the
        fitted parameter vectors of ALL genes and models of a test live in one file with the same prefix."""
        import os
        if not os.path.exists(self.folder_name):
            os.mkdir(self.folder_name)
        lik_name = lik_name or self.lik_name
        n = self.models_number if test_name is None else self._TEST_MODELS.get(test_name, 1)
        return self.folder_name + lik_name + "_" + ("sparse_" if self.sparse else "") + ("tst_" if n == 3 else "") + \
            "models_%d.npz" % n

    def _remember(self, test_name, models, names, stats, genes_index, Xs, Ys, kernel_names):
        """Keep what load_predict_models needs: theta-hat of the local shard stays on the device, statistics on the host."""
        key = (test_name, self.lik_name)
        self._fitted[key] = dict(models=models, names=names, stats=np.asarray(stats), genes=[self.genes_name[g] for g in genes_index],
                                 theta=self._last_theta, local_pos=np.asarray(self._last_local_pos), Xs=Xs, Ys=Ys,
                                 kernels=kernel_names, Z=self.Z, sparse=self.sparse)
        if self.persist_models and self._last_theta is not None:
            pos = np.asarray(self._last_local_pos)
            np.savez_compressed(self.get_file_name(test_name), genes=np.array(self._fitted[key]["genes"], dtype=object)[pos],
                                theta=self._last_theta.cpu().numpy(), stats=np.asarray(stats)[pos], names=np.array(names),
                                kernels=np.array(kernel_names), lik=self.lik_name, X=self.X,
                                Z=self.Z if self.Z is not None else np.zeros(0), P=np.array([m.P for m in models]))

    def _latent_at(self, model, theta, Xnew, Y=None, full_cov=True):
        from . import engine
        return engine.predict(model, theta, Xnew, Y=Y, full_cov=full_cov)

    def load_predict_models(self, genes_name, test_name, likelihood="Negative_binomial", predict=True):
        """Reference :859-960: fitted models of the listed genes and, with ``predict``, their posterior predictive at 100
        test inputs - ``predict_y`` for the Gaussian likelihood, the Monte-Carlo recipe of
        ``samples_posterior_predictive_distribution`` otherwise (means: smoothed per-point averages, vars: the count samples).
        The models come from the last run of ``test_name`` with this likelihood in this object (or, with
        ``persist_models``, from the file written then)."""
        from . import engine, predictive
        self.Y = self.Y_copy
        self.lik_name = likelihood
        params = {"test_name": test_name, "likelihood": likelihood}
        key = (test_name, likelihood)
        if key not in self._fitted:
            raise RuntimeError("no fitted models for %s / %s in this object: run the test first" % key)
        fit = self._fitted[key]
        models, K = fit["models"], len(fit["models"])
        self.models_number = self._TEST_MODELS.get(test_name, 1)
        X1 = self.X[:, :1] if self.X.shape[1] > 1 else self.X
        xtest = np.linspace(np.min(self.X) - 0.1, np.max(self.X) + 0.1, 100)[:, None]          # :880
        if self.X.shape[1] > 1:
            xtest = self.X                                                                       # 2-D inputs: predict at the data
        rows = [fit["genes"].index(g) for g in genes_name]
        local = {int(p): i for i, p in enumerate(fit["local_pos"])}
        missing = [genes_name[i] for i, r in enumerate(rows) if r not in local]
        if missing:
            raise RuntimeError("fitted parameters of %s live on another rank" % missing)
        sel = [local[r] for r in rows]
        genes_means, genes_vars, genes_models = [], [], []
        per_model = []
        for k, m in enumerate(models):
            theta = fit["theta"][sel, k, :m.P].contiguous()
            st = fit["stats"][rows, k]
            Yk = fit["Ys"][k]
            mean = var = None
            if predict:
                Yd = np.ascontiguousarray(Yk[rows]) if likelihood == "Gaussian" else None
                mf, vf, cf = self._latent_at(m, theta, xtest, Y=Yd, full_cov=likelihood != "Gaussian")
                if likelihood == "Gaussian":
                    noise = torch.as_tensor(st[:, S_ALPHA], device=mf.device)[:, None]           # predict_y adds the noise variance
                    mean, var = mf.cpu().numpy()[:, :, None], (vf + noise).cpu().numpy()[:, :, None]
                else:
                    dev = mf.device
                    mean, var = predictive.posterior_predictive(
                        likelihood, mf, cf, torch.as_tensor(st[:, S_ALPHA], device=dev), torch.as_tensor(st[:, S_KM], device=dev),
                        branching=False, keep_samples=True)
            per_model.append((theta, st, mean, var))
        for i, g in enumerate(genes_name):
            means, variances, mods = [], [], []
            for k, m in enumerate(models):
                theta, st, mean, var = per_model[k]
                ok = not np.isnan(st[i, S_LL])
                means.append(mean[i] if (predict and ok) else 0)
                variances.append(var[i] if (predict and ok) else 0)
                mods.append(FittedModel(st[i], theta[i].cpu().numpy(), fit["Xs"][k], fit["Ys"][k][rows[i]].reshape(-1, 1),
                                        Z=fit["Z"] if fit["sparse"] else None, kernel=fit["kernels"][k], lik_name=likelihood)
                            if ok else np.nan)
            genes_means.append(means); genes_vars.append(variances); genes_models.append(mods)
        params["means"], params["vars"], params["models"] = genes_means, genes_vars, genes_models
        return params

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

    def _run(self, models, Yfit, hyp0, chain, genes_index, attempt0=None):
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
            opt = self._opt()
            a0 = None
            if attempt0 is not None:                                   # host-driven restarts of the safe mode
                a0 = torch.as_tensor(np.ascontiguousarray(attempt0[gsel], dtype=np.int32), device=dev)
                opt.attempt0 = a0.data_ptr()
            stats, theta = engine.fit(models, np.ascontiguousarray(Yfit[gsel]), scale, hyp0[gsel], chain=chain,
                                      opt=opt, return_theta=True)
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

    # ------------------------------------------------------------------------------------------ safe mode
    def _predictive_mean(self, model, theta, st, xtest, Yk):
        """Mean of the posterior predictive at xtest for the rows of theta (reference :962-977)."""
        from . import predictive
        out = np.zeros((theta.shape[0], len(xtest)))
        for c0 in range(0, theta.shape[0], 512):                       # bounded sampling buffers
            sl = slice(c0, c0 + 512)
            full = self.lik_name != "Gaussian"
            mf, vf, cf = self._latent_at(model, theta[sl], xtest, Y=np.ascontiguousarray(Yk[sl]) if not full else None, full_cov=full)
            if not full:
                out[sl] = mf.cpu().numpy()                              # predict_y mean
            else:
                dev = mf.device
                out[sl], _ = predictive.posterior_predictive(self.lik_name, mf, cf, torch.as_tensor(st[sl, S_ALPHA], device=dev),
                                                             torch.as_tensor(st[sl, S_KM], device=dev), seed=c0)
        return out

    def _safe_fit(self, model, Yfit, hyp0, genes_index, constant_of_one_sample=False):
        """``safe_mode=True`` for one model of the test (reference fit_GP :558-565, test_local_optima_case1 :962-1006,
        positive log-likelihood in fit_model :518-525): after a successful fit whose restart count is below 5 the posterior
        predictive mean is compared with the data range; a fit that fails the check (or has a positive log-likelihood) is
        repeated from the next seeded random restart.  Host-driven: rounds of batched fits of the flagged genes
        (gpc_opt_t.attempt0) and batched predictions.  The check is Monte-Carlo: parity with the reference is statistical."""
        genes_index = np.asarray(list(genes_index), dtype=np.int64)
        G = len(genes_index)
        Xm = model.X.cpu().numpy()
        x = (self.Z if self.sparse else Xm)
        xtest = np.linspace(np.min(x), np.max(x), 100)[:, None] if Xm.shape[1] == 1 else Xm          # :969-972
        Yk = Yfit[:, model.n_off:model.n_off + model.N]
        a0 = np.zeros((G, 1), dtype=np.int32)
        stats = np.zeros((G, 1, GPC_NSTAT))
        theta = None
        pending = np.arange(G)
        diff = 0 if model.N < 100 else 1                                                         # :989-992
        for _round in range(12):
            if len(pending) == 0:
                break
            st = self._run([model], Yfit, hyp0, True, genes_index[pending], attempt0=_expand(a0, pending, len(Yfit), genes_index))
            th = self._last_theta
            if theta is None:
                theta = torch.zeros(G, 1, model.P, dtype=torch.float64, device=th.device)
            theta[torch.as_tensor(pending, device=th.device)] = th[:, :, :model.P]
            stats[pending] = st
            ok = (st[:, 0, S_STATUS] == 0)
            c = st[:, 0, S_RESTARTS].astype(int)
            chk = np.nonzero(ok & (c < 5))[0]
            flag = np.zeros(len(pending), dtype=bool)
            if len(chk):
                mean = self._predictive_mean(model, th[torch.as_tensor(chk, device=th.device), 0, :model.P].contiguous(),
                                             st[chk, 0], xtest, Yk[genes_index[pending[chk]]])
                y = Yk[genes_index[pending[chk]]]
                m_mean, m_max, m_min = mean.mean(axis=1), mean.max(axis=1), mean.min(axis=1)
                y_mean, y_max, y_min = y.mean(axis=1), y.max(axis=1), y.min(axis=1)
                bad = np.zeros(len(chk), dtype=bool)
                if constant_of_one_sample:                                                       # :993-995
                    bad |= (m_min < y_min) | (m_max > y_max) | (m_mean == 0.0)
                with np.errstate(divide="ignore", invalid="ignore"):
                    dm = np.abs(np.round((m_mean - y_mean) / y_mean))
                bad |= (y_mean > 0.0) & (((dm > diff) & ((m_min < y_min) | (m_max > y_max))) | (m_mean == 0.0))   # :997-1006
                flag[chk] = bad
            pos = ok & (st[:, 0, S_LL] > 0) & (c < 10) & (self.lik_name != "Gaussian")           # :518-525
            redo = flag | pos
            a0[pending[redo], 0] = np.minimum(c[redo] + 1 + pos[redo].astype(int), 10)
            pending = pending[redo]
        return stats, theta

    def _run_safe(self, models, Yfit, hyp0, genes_index, models_number):
        """The test with ``safe_mode=True``: model by model (the checks need each model's own fitted state)."""
        genes_index = list(genes_index)
        G, K = len(genes_index), len(models)
        dist_saved, self.distributed = self.distributed, False          # the host-driven rounds run unsharded
        try:
            Pmax = max(m.P for m in models)
            stats = np.zeros((G, K, GPC_NSTAT))
            theta = None
            gi = np.asarray(genes_index, dtype=np.int64)
            for k, m in enumerate(models):
                h = hyp0[:, k:k + 1].copy()
                sub = np.arange(G)
                if models_number == 2 and k == 1:
                    alive = ~np.isnan(stats[:, 0, S_LL])                                   # :441-442
                    stats[~alive, 1, S_LL], stats[~alive, 1, S_STATUS] = np.nan, 3
                    sub = np.nonzero(alive)[0]
                    if m.handoff_alpha:
                        h[gi[sub], 0, 2] = stats[sub, 0, S_ALPHA]                          # :785-787 (first attempt only)
                if len(sub) == 0:
                    continue
                st, th = self._safe_fit(m, Yfit, h, gi[sub], constant_of_one_sample=(models_number == 2 and k == 1))
                stats[sub, k] = st[:, 0]
                if theta is None:
                    theta = torch.zeros(G, K, Pmax, dtype=torch.float64, device=th.device)
                theta[torch.as_tensor(sub, device=th.device), k, :m.P] = th[:, 0]
        finally:
            self.distributed = dist_saved
        self._last_theta, self._last_local_pos = theta, np.arange(G)
        return stats

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
            stats = self._run_safe(models, Yfit, hyp0, genes_index, 1) if self.safe_mode else self._run(models, Yfit, hyp0, True, genes_index)
            out = stats[:, :, S_LL]
            names = ["dynamic"]
        elif models_number == 2:
            handoff = lik_name == "Negative_binomial"                                  # :785-787
            models = [self._make_model("RBF", X), self._make_model("CONST", X, handoff=handoff)]
            # Gaussian quirk: the constant model's initial variance is computed from the already
            # log-transformed y of the dynamic fit (:652 runs before :738 of the next model)
            yv2 = Yfit if (lik_name == "Gaussian" and self.transform) else None
            hyp0 = np.stack([self._hyp0(Y, X), self._hyp0(Y, X, kernel="CONST", y_for_var=yv2)], axis=1)
            stats = self._run_safe(models, Yfit, hyp0, genes_index, 2) if self.safe_mode else self._run(models, Yfit, hyp0, True, genes_index)
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
            stats = self._run_safe(models, Yfit, hyp0, genes_index, 3) if self.safe_mode else self._run(models, Yfit, hyp0, False, genes_index)
            ll = stats[:, :, S_LL]
            out = np.c_[ll, (ll[:, 1] + ll[:, 2]) - ll[:, 0]]                          # NaN propagates (:486-495)
            names = ["shared", "model_1", "model_2"]
        self._record_stats(stats, genes_index, names)
        test_name = {1: "Infer_trajectory", 2: "One_sample_test", 3: "Two_samples_test"}[models_number]
        Xs = [m.X.cpu().numpy() for m in models]
        Ys = [Yfit[:, m.n_off:m.n_off + m.N] for m in models]
        self._remember(test_name, models, names, stats, genes_index, Xs, Ys, [m.kernel for m in models])
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
        # posterior predictive of the MAP model on both branches (reference :305-317)
        x1 = np.c_[Xnew, np.ones(len(Xnew))[:, None]]
        x2 = np.c_[Xnew, (np.ones(len(Xnew)) * 2)[:, None]]
        Xtest = np.concatenate((x1, x2))
        Xtest[np.where(Xtest[:, 0] <= self.model.kernel.xp), 1] = 1
        mu = var = None
        if theta is not None:
            from . import predictive
            th = torch.as_tensor(theta[None, :])
            gauss = self.lik_name == "Gaussian"
            mf, vf, cf = self._latent_at(models[iMAP], th, Xtest, Y=np.ascontiguousarray(Yfit[[g_last]]) if gauss else None,
                                         full_cov=not gauss)
            if gauss:
                mu = mf.cpu().numpy()[0][:, None]
                var = (vf.cpu().numpy()[0] + stats[0, iMAP, S_ALPHA])[:, None]            # predict_y: + noise variance
            else:
                dev = mf.device
                m_, v_ = predictive.posterior_predictive(self.lik_name, mf, cf, torch.as_tensor(stats[0, iMAP, S_ALPHA:S_ALPHA + 1], device=dev),
                                                         torch.as_tensor(stats[0, iMAP, S_KM:S_KM + 1], device=dev), branching=True,
                                                         keep_samples=True)
                mu, var = m_[0], v_[0]
        self.branching = False
        return {
            "geneName": self.genes_name,
            "branching_probability": ll,
            "branching_location": self.model.kernel.xp,
            "mean": mu,
            "variance": var,
            "Xnew": Xnew,
            "test_times": testTimes,
            "MAP_model": self.model,
            "loglik": log_ll,
            "logBayesFactor": p["logBayesFactor"],
            "likelihood": self.lik_name,
        }

    def CalculateBranchingEvidence(self, d, Bsearch):
        return host.branching_evidence(d["loglik"], Bsearch)
