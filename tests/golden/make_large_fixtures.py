"""Large-sample oracle fixtures (CPU oracle = restatement of the reference's GPflow + SciPy path; run in the build
container, takes about two hours of 7 cores in total):

    python tests/golden/make_large_fixtures.py c5_2000 mouseob_455 c5_2000_refpolicy c5_runs fission

  c5_2000.npz            2000 genes of the headline workload (bench.synth_batch(2000, 1000, seed=4321); N = 1000, M = 64,
                         sparse NB One_sample_test): log-likelihoods / iteration counts of the oracle from the exact
                         start and from a start perturbed by 1e-12 (the oracle's own run-to-run distribution).
  c5_2000_refpolicy.npz  the same genes with the reference's inducing-point policy (RobustGP selection under each gene's
                         own initial variance, GPcounts_Module.py:673-677) instead of one shared set.
  mouseob_455.npz        the 455 MouseOB genes of data_mouseob.npz (full VGP, N = 260, 2-D RBF, scaled NB): default run,
                         perturbed run and the constant model re-fitted with ftol=1e-15, gtol=1e-8 ("tight") - a gene is
                         *well-posed* when default and tight LLR agree to 1e-4 relative, else *plateau-stopped*
                         (SURVEY.md section 4.3).
  c5_sample_sparse_nb_runs.npz   32 perturbed oracle runs (k * 1e-13) of the dynamic fit of the 48-gene C5 sample.
  fission.npz            25 genes of the fission time course: One_sample_test on the 18 wild-type samples and
                         Two_samples_test on all 36 (full VGP, N = 18 / 36), default + perturbed runs.
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

NJ = int(os.environ.get("NJ", "7"))


def _fit_summary(f):
    if f is None:
        return [np.nan, 0, 0, 0, np.nan, np.nan, np.nan]
    h = f.hyper() if f.theta is not None else dict(var=np.nan, ls=np.nan, alpha=np.nan)
    return [f.log_likelihood, f.nit, f.nfev, f.restarts, h["var"], h["ls"], h["alpha"]]


def _one_sample(kw, g, perturb):
    torch.set_num_threads(1)
    ofit.PERTURB = perturb
    o = ofit.OracleFit(**kw)
    r = o.one_sample_test("Negative_binomial", genes=[g])[0]
    return np.array([_fit_summary(f) for f in r["fits"]])          # (2, 7)


def _pack(res):
    a = np.stack(res)                                              # (D, 2, 7)
    ll = np.c_[a[:, :, 0], a[:, 0, 0] - a[:, 1, 0]]
    return dict(ll=ll, nit=a[:, :, 1], nfev=a[:, :, 2], restarts=a[:, :, 3], hyper=a[:, :, 4:7])


def case_c5_2000(refpolicy=False):
    from bench import synth_batch
    D, N, M = 2000, 1000, 64
    X, Y = synth_batch(D, N, seed=4321)
    kw = dict(X=X, Y=Y, sparse=True, M=M, inducing_variance=None if refpolicy else 1.0)
    out = dict(D=D, N=N, M=M, seed=4321, y_checksum=float(Y.sum()), y_checksum2=float((Y * np.arange(N)).sum()))
    for tag, pert in ((("", 0.0),) if refpolicy else (("", 0.0), ("_pert", 1e-12))):
        res = Parallel(n_jobs=NJ, batch_size=4)(delayed(_one_sample)(kw, g, pert) for g in range(D))
        for k, v in _pack(res).items():
            out[k + tag] = v
    np.savez_compressed(os.path.join(HERE, "c5_2000_refpolicy.npz" if refpolicy else "c5_2000.npz"), **out)


def _mouseob_gene(X, Y, scale, g):
    torch.set_num_threads(1)
    out = {}
    for tag, pert in (("", 0.0), ("_pert", 1e-12)):
        ofit.PERTURB = pert
        o = ofit.OracleFit(X=X, Y=Y, scale=scale)
        r = o.one_sample_test("Negative_binomial", genes=[g])[0]
        out["fits" + tag] = np.array([_fit_summary(f) for f in r["fits"]])
    # constant model from the same start (alpha handed over from the default dynamic fit) with tight tolerances
    ofit.PERTURB = 0.0
    alpha = out["fits"][0, 6]
    tight = np.full(7, np.nan)
    if np.isfinite(alpha):
        ot = ofit.OracleFit(X=X, Y=Y, scale=scale, lbfgs_options=dict(ftol=1e-15, gtol=1e-8, maxiter=5000))
        ot.transform = True
        rt, _ = ot.fit_model(ot.X, Y[g].astype(int).astype(float), scale[:, g], "NB", 2, 2, lik_alpha=alpha)
        tight = np.array(_fit_summary(rt))
    out["tight_const"] = tight
    return out


def case_mouseob_455():
    d = np.load(os.path.join(HERE, "data_mouseob.npz"), allow_pickle=True)
    X, Y, scale = d["X"], d["Y"].astype(np.float64), d["scale"]
    D = Y.shape[0]
    res = Parallel(n_jobs=NJ)(delayed(_mouseob_gene)(X, Y, scale, g) for g in range(D))
    out = {}
    for tag in ("", "_pert"):
        for k, v in _pack([r["fits" + tag] for r in res]).items():
            out[k + tag] = v
    tc = np.stack([r["tight_const"] for r in res])
    out["ll_const_tight"] = tc[:, 0]
    out["nit_const_tight"] = tc[:, 1]
    out["llr_tight"] = out["ll"][:, 0] - tc[:, 0]
    np.savez_compressed(os.path.join(HERE, "mouseob_455.npz"), genes=d["genes"], **out)


def _traj_run(X, Y, M, g, k):
    torch.set_num_threads(1)
    ofit.PERTURB = k * 1e-13
    o = ofit.OracleFit(X=X, Y=Y, sparse=True, M=M, inducing_variance=1.0)
    r = o.infer_trajectory("Negative_binomial", genes=[g])[0]
    return g, k, r["ll"][0], r["fits"][0].nit


def case_c5_runs():
    fx = np.load(os.path.join(HERE, "c5_sample_sparse_nb.npz"), allow_pickle=True)
    X, Y, M = fx["X"], fx["Y"], int(fx["M"])
    R = 32
    res = Parallel(n_jobs=NJ, batch_size=4)(delayed(_traj_run)(X, Y, M, g, k) for g in range(Y.shape[0]) for k in range(R))
    ll = np.zeros((Y.shape[0], R)); nit = np.zeros((Y.shape[0], R), dtype=int)
    for g, k, v, n in res:
        ll[g, k], nit[g, k] = v, n
    np.savez_compressed(os.path.join(HERE, "c5_sample_sparse_nb_runs.npz"), ll_runs=ll, nit_runs=nit, perturb_step=1e-13)


def _fission_gene(X, Y, g, perturb, which):
    torch.set_num_threads(1)
    ofit.PERTURB = perturb
    if which == "one":
        o = ofit.OracleFit(X=X[:18], Y=Y[:, :18])
        r = o.one_sample_test("Negative_binomial", genes=[g])[0]
    else:
        o = ofit.OracleFit(X=X, Y=Y)
        r = o.two_samples_test("Negative_binomial", genes=[g])[0]
    return np.array(r["ll"]), np.array([_fit_summary(f) for f in r["fits"]])


def case_fission():
    d = np.load(os.path.join(HERE, "data_fission.npz"), allow_pickle=True)
    X, Y = d["minute"][:, None], d["Y"]
    D = Y.shape[0]
    out = {}
    for which in ("one", "two"):
        for tag, pert in (("", 0.0), ("_pert", 1e-12)):
            res = Parallel(n_jobs=NJ)(delayed(_fission_gene)(X, Y, g, pert, which) for g in range(D))
            out["ll_%s%s" % (which, tag)] = np.stack([r[0] for r in res])
            out["fits_%s%s" % (which, tag)] = np.stack([r[1] for r in res])
    np.savez_compressed(os.path.join(HERE, "fission.npz"), genes=d["genes"], **out)


CASES = {"c5_2000": case_c5_2000, "c5_2000_refpolicy": lambda: case_c5_2000(True), "mouseob_455": case_mouseob_455,
         "c5_runs": case_c5_runs, "fission": case_fission}

if __name__ == "__main__":
    for n in sys.argv[1:] or list(CASES):
        t = time.time()
        CASES[n]()
        print("%s done in %.1fs" % (n, time.time() - t), flush=True)
