"""Generate the golden fit fixtures under tests/golden/ with the CPU oracle (run in the build container):

    python tests/golden/make_fixtures.py [case ...]

Each fixture holds the inputs of one Fit_GPcounts workflow and the oracle's outputs (log-likelihoods per
model, iteration counts), once from the exact initial point and once from a start perturbed by 1e-12
(``oracle.fit.PERTURB``).  A fit whose two runs disagree is *roundoff-unstable in the reference itself*
(SURVEY.md section 4.3 / DESIGN.md "what parity means"); the GPU tests apply the 1e-4 LLR tolerance to the
stable fits and only the DE-call criterion to the rest.

The MouseOB fixture additionally stores the numbers the reference itself published for those genes
(``demo_notebooks/GPcounts_spatial.ipynb:1080,1086`` and ``data/MouseOB/SE_genes_comparison.csv``); its
inputs are columns of ``data/MouseOB/{mouse_ob_SV_genes,scale_nb_model_sel,locations}.csv``.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

torch.set_num_threads(1)
from joblib import Parallel, delayed  # noqa: E402

from oracle import fit as ofit  # noqa: E402
from tests.helpers import synth_counts  # noqa: E402

REF = "/root/reference/data/"


def _rows_to_arrays(rows):
    K = len(rows[0]["fits"])
    D = len(rows)
    ll = np.full((D, len(rows[0]["ll"])), np.nan)
    nit = np.zeros((D, K)); nfev = np.zeros((D, K)); rst = np.zeros((D, K)); hyp = np.full((D, K, 4), np.nan)
    for i, r in enumerate(rows):
        ll[i] = r["ll"]
        for k, f in enumerate(r["fits"]):
            if f is None:
                continue
            nit[i, k], nfev[i, k], rst[i, k] = f.nit, f.nfev, f.restarts
            if f.theta is not None:
                h = f.hyper()
                hyp[i, k] = [h["var"], h["ls"], h["alpha"], h["km"]]
    return dict(ll=ll, nit=nit, nfev=nfev, restarts=rst, hyper=hyp)


def _one_gene(kw, method, margs, g, perturb):
    torch.set_num_threads(1)
    ofit.PERTURB = perturb
    o = ofit.OracleFit(**kw)
    rows = getattr(o, method)(*margs, genes=[g])
    return rows[0]


def run_tests(kw, method, margs, D):
    out = {}
    for tag, pert in (("", 0.0), ("_pert", 1e-12)):
        rows = Parallel(n_jobs=8)(delayed(_one_gene)(kw, method, margs, g, pert) for g in range(D))
        for k, v in _rows_to_arrays(rows).items():
            out[k + tag] = v
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def case_one_sample_sparse(lik_name, tag, D=6, N=120, M=10, seed=5):
    X, Y = synth_counts(D, N, seed=seed)
    # inducing_variance=1.0: identical inducing points in oracle and product (see oracle/fit.py)
    res = run_tests(dict(X=X, Y=Y, sparse=True, M=M, inducing_variance=1.0), "one_sample_test", (lik_name,), D)
    save("one_sample_sparse_" + tag, X=X, Y=Y, M=M, lik=lik_name, **res)


def case_one_sample_vgp_scaled():
    D, N = 4, 60
    X, Y = synth_counts(D, N, seed=9)
    rng = np.random.default_rng(3)
    X = np.c_[X, rng.uniform(0, 1, N)]
    scale = rng.uniform(0.5, 2.0, (N, D))
    res = run_tests(dict(X=X, Y=Y, scale=scale), "one_sample_test", ("Negative_binomial",), D)
    save("one_sample_vgp_scaled", X=X, Y=Y, scale=scale, lik="Negative_binomial", **res)


def case_two_samples():
    D, N = 3, 40
    X, Y = synth_counts(D, N, seed=21)
    res = run_tests(dict(X=X, Y=Y), "two_samples_test", ("Negative_binomial",), D)
    save("two_samples_vgp_nb", X=X, Y=Y, lik="Negative_binomial", **res)


def case_gauss():
    D, N = 4, 50
    X, Y = synth_counts(D, N, seed=33)
    res = run_tests(dict(X=X, Y=Y), "one_sample_test", ("Gaussian",), D)
    save("one_sample_gpr_gauss", X=X, Y=Y, lik="Gaussian", **res)


def case_gauss_sparse():
    D, N, M = 4, 80, 8
    X, Y = synth_counts(D, N, seed=35)
    res = run_tests(dict(X=X, Y=Y, sparse=True, M=M, inducing_variance=1.0), "one_sample_test", ("Gaussian",), D)
    save("one_sample_sgpr_gauss", X=X, Y=Y, M=M, lik="Gaussian", **res)


def case_c5_sample():
    """48 genes of the headline benchmark workload (bench.py generator, N = 1000, M = 64)."""
    sys.path.insert(0, ROOT)
    from bench import synth_batch
    D, N, M = 48, 1000, 64
    X, Y = synth_batch(D, N, seed=1234)
    res = run_tests(dict(X=X, Y=Y, sparse=True, M=M, inducing_variance=1.0), "one_sample_test",
                    ("Negative_binomial",), D)
    save("c5_sample_sparse_nb", X=X, Y=Y, M=M, lik="Negative_binomial", **res)


def case_branching():
    N, bins = 40, 5
    rng = np.random.default_rng(7)
    t = np.sort(rng.uniform(0, 1, N))
    labels = rng.integers(1, 3, N).astype(float)
    f = np.log(20.0) + np.where((labels == 2) & (t > 0.5), 2.0 * (t - 0.5), 0.0)
    Y = rng.negative_binomial(5.0, 5.0 / (5.0 + np.exp(f)))[None, :].astype(float)
    out = {}
    for tag, pert in (("", 0.0), ("_pert", 1e-12)):
        ofit.PERTURB = pert
        o = ofit.OracleFit(t[:, None], Y)
        r = o.infer_branching_location(labels, bins_num=bins)
        out["loglik" + tag] = r["loglik"]
        out["free_ll" + tag] = r["free_fit"].log_likelihood
        out["free_hyper" + tag] = np.array([r["free_fit"].hyper()[k] for k in ("var", "ls", "alpha", "km")])
        out["branching_probability" + tag] = r["branching_probability"]
        out["logBayesFactor" + tag] = r["logBayesFactor"]
    ofit.PERTURB = 0.0
    save("branching_nb", t=t, labels=labels, Y=Y, bins=bins, **out)


def case_mouseob():
    import pandas as pd
    loc = pd.read_csv(REF + "MouseOB/locations.csv", index_col=0)
    Yall = pd.read_csv(REF + "MouseOB/mouse_ob_SV_genes.csv", index_col=0)
    Sall = pd.read_csv(REF + "MouseOB/scale_nb_model_sel.csv", index_col=0)
    paper = pd.read_csv(REF + "MouseOB/SE_genes_comparison.csv", index_col=0).set_index("gene")
    genes = ["2010300C02Rik", "Slc1a3", "Mcf2l", "Dab2", "Ptpro", "Pmepa1", "Fgfr1", "Bmp7"]
    X = loc.iloc[:, :2].values.astype(float)
    Y = Yall[genes].values.T.astype(float)
    scale = Sall[genes].values.astype(float)
    res = run_tests(dict(X=X, Y=Y, scale=scale), "one_sample_test", ("Negative_binomial",), len(genes))
    # the notebook goldens were produced with alpha0 = 1 for the constant model
    res_one = {}
    rows = Parallel(n_jobs=8)(delayed(_one_gene)(dict(X=X, Y=Y, scale=scale, alpha_policy="one"),
                                                  "one_sample_test", ("Negative_binomial",), g, 0.0)
                              for g in range(2))
    res_one["ll_alpha_one"] = _rows_to_arrays(rows)["ll"]
    notebook = np.array([[-660.911524, -683.227119, 22.315595], [-1112.007845, -1123.880571, 11.872727]])
    paper_llr = np.array([paper.loc[g, "LLR_gpcounts"] if g in paper.index else np.nan for g in genes])
    save("mouseob_vgp_scaled", X=X, Y=Y, scale=scale, genes=np.array(genes), lik="Negative_binomial",
         notebook_ll=notebook, paper_llr=paper_llr, **res, **res_one)


CASES = {
    "sparse_nb": lambda: case_one_sample_sparse("Negative_binomial", "nb"),
    "sparse_poisson": lambda: case_one_sample_sparse("Poisson", "poisson"),
    "sparse_zinb": lambda: case_one_sample_sparse("Zero_inflated_negative_binomial", "zinb", D=4, N=100, M=8, seed=6),
    "vgp_scaled": case_one_sample_vgp_scaled,
    "two_samples": case_two_samples,
    "gauss": case_gauss,
    "gauss_sparse": case_gauss_sparse,
    "c5_sample": case_c5_sample,
    "branching": case_branching,
    "mouseob": case_mouseob,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        t = time.time()
        CASES[n]()
        print("%s done in %.1fs" % (n, time.time() - t))
