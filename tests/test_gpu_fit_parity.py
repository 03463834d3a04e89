"""Optimised-fit parity through the drop-in API: ``Fit_GPcounts`` on the GPU vs the golden fixtures the CPU
oracle produced (``tests/golden/make_fixtures.py``; SciPy L-BFGS-B from identical initialisation).

Tolerances (BASELINE.json north_star): per-gene log-likelihood ratios within 1e-4 relative.  Both optimisers
stop on SciPy's relative-decrease rule with gradients of 1e-3..1e-2 and the ELBO carries ~1e-10 relative
evaluation noise, so even the reference is only reproducible to ~1e-6..1e-5 in the ELBO, and some fits are
chaotic at round-off level.  Every fixture therefore carries a second oracle run from a
start perturbed by 1e-12: fits whose two oracle runs agree are *stable* and must match tightly; the others
are only required to land on one of the oracle's answers or between them.

Even a fit whose two oracle runs agree can stop early on a plateau: SciPy's rule ends the run the first time one iteration
gains less than 2.2e-9 |f|, and along a creeping valley whether that happens at iteration 240 or 450 is decided by round-off.
A fit therefore has a small *set* of outcomes - `<fixture>_runs.npz` holds the oracle's outcomes over 32 starts perturbed by
k * 1e-13 (`tests/golden/make_perturbed_runs.py`, `make_large_fixtures.py`).  A GPU log-likelihood is accepted when it coincides
with ANY of the oracle's recorded outcomes (primary run, 1e-12 run, the 32 runs); there are no retries and no substituted
values.  That the GPU visits those outcomes with the oracle's frequencies is tested in tests/test_gpu_fit_large.py.
"""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LL_RTOL = 5e-6      # per-fit ELBO, relative (stable fits)
LLR_RTOL = 1e-4     # north_star
LLR_ATOL = 2e-3     # for LLR ~ 0 (constant genes)


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=True)


def _frames(fx):
    X, Y = fx["X"], fx["Y"]
    cells = ["c%d" % i for i in range(X.shape[0])]
    genes = ["g%d" % i for i in range(Y.shape[0])]
    Xdf = pd.DataFrame(X, index=cells)
    Ydf = pd.DataFrame(Y, index=genes, columns=cells)
    sdf = pd.DataFrame(fx["scale"], index=cells) if "scale" in fx.files else None
    return Xdf, Ydf, sdf


def _oracle_outcomes(name, g):
    """Dynamic-model log-likelihoods of gene g over the oracle's 32 perturbed starts, if the fixture has them."""
    path = os.path.join(GOLD, name + "_runs.npz")
    return np.load(path)["ll_runs"][g] if os.path.exists(path) else None


def _check_table(df, fx, K, skip=None, name=None):
    """Every per-model log-likelihood must coincide with one of the oracle's outcomes for that fit; the LLR of genes whose
    oracle runs agree must match to 1e-4 relative (north_star), shifted by the outcome the dynamic fit landed on."""
    got = df.values.copy()
    ref, ref2 = fx["ll"], fx["ll_pert"]
    assert got.shape == ref.shape
    n_stable = 0
    for g in range(ref.shape[0]):
        shift = 0.0           # oracle outcome the dynamic fit coincides with, minus the oracle's primary outcome
        for k in range(K):
            a, b = ref[g, k], ref2[g, k]
            stable = abs(a - b) <= 1e-6 * abs(a)
            n_stable += int(stable)
            if skip is not None and (g, k) in skip:
                continue
            outcomes = [a, b]
            if k == 0 and name is not None:
                runs = _oracle_outcomes(name, g)
                if runs is not None:
                    outcomes += list(runs)
            outcomes = np.array(outcomes)
            # unstable fits (the oracle's two runs disagree): plateau stops, chaotic at round-off level - the envelope
            tol = LL_RTOL * abs(a) if stable else max(10.0 * abs(a - b), 3e-4 * abs(a))
            j = int(np.abs(outcomes - got[g, k]).argmin())
            assert abs(got[g, k] - outcomes[j]) <= tol, (g, k, got[g, k], a, b)
            if k == 0:
                shift = outcomes[j] - a
        a, b = ref[g, K], ref2[g, K]
        if skip is not None and any(gg == g for gg, _ in skip):
            continue
        if abs(a - b) <= 1e-5 * abs(a) + 1e-5:
            assert abs(got[g, K] - (a + shift)) <= LLR_RTOL * abs(a) + LLR_ATOL, (g, got[g, K], a, shift)
    assert n_stable >= 0.7 * ref.shape[0] * K


@pytest.mark.parametrize("name", ["one_sample_sparse_nb", "one_sample_sparse_poisson", "one_sample_sparse_zinb"])
def test_one_sample_sparse(name):
    from gpcounts_b200 import Fit_GPcounts
    fx = _load(name)
    Xdf, Ydf, _ = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=int(fx["M"]))
    df = gp.One_sample_test(str(fx["lik"]))
    assert list(df.columns) == ["Dynamic_model_log_likelihood", "Constant_model_log_likelihood", "log_likelihood_ratio"]
    assert list(df.index) == list(Ydf.index)
    _check_table(df, fx, 2, name=name)
    assert (gp.fit_stats["status"] == 0).all()
    out = gp.calculate_FDR(df)
    assert {"p_value", "q_value"} <= set(out.columns)


def test_one_sample_vgp_scaled_2d():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("one_sample_vgp_scaled")
    Xdf, Ydf, sdf = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf, scale=sdf)
    df = gp.One_sample_test("Negative_binomial")
    _check_table(df, fx, 2)
    # the model object of the last gene exposes the attribute paths the reference uses
    assert np.isfinite(gp.model.likelihood.alpha.numpy())
    assert np.isfinite(gp.model.kernel.variance.numpy())


def test_infer_trajectory_matches_dynamic_column():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("one_sample_sparse_nb")
    Xdf, Ydf, _ = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=int(fx["M"]))
    df = gp.Infer_trajectory("Negative_binomial")
    assert list(df.columns) == ["Dynamic_model_log_likelihood"]
    ref = fx["ll"][:, 0]
    stable = np.abs(ref - fx["ll_pert"][:, 0]) <= 1e-6 * np.abs(ref)
    for g in np.nonzero(stable)[0]:
        outcomes = np.r_[ref[g], _oracle_outcomes("one_sample_sparse_nb", g)]
        assert np.abs(outcomes - df.values[g, 0]).min() <= LL_RTOL * abs(ref[g]), (g, df.values[g, 0], ref[g])


def test_two_samples():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("two_samples_vgp_nb")
    Xdf, Ydf, _ = _frames(fx)
    Xdf.columns = ["times"]
    gp = Fit_GPcounts(Xdf, Ydf)
    df = gp.Two_samples_test("Negative_binomial")
    assert list(df.columns) == ["Shared_log_likelihood", "model_1_log_likelihood", "model_2_log_likelihood",
                                "log_likelihood_ratio"]
    _check_table(df, fx, 3)


def test_gaussian_gpr():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("one_sample_gpr_gauss")
    Xdf, Ydf, _ = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf)
    df = gp.One_sample_test("Gaussian")
    _check_table(df, fx, 2)


def test_gaussian_sgpr():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("one_sample_sgpr_gauss")
    Xdf, Ydf, _ = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=int(fx["M"]))
    df = gp.One_sample_test("Gaussian")
    _check_table(df, fx, 2)


def test_branching():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("branching_nb")
    t, labels, Y = fx["t"], fx["labels"], fx["Y"]
    cells = ["c%d" % i for i in range(len(t))]
    gp = Fit_GPcounts(pd.DataFrame(t, index=cells), pd.DataFrame(Y, index=["gene"], columns=cells))
    res = gp.Infer_branching_location(labels, bins_num=int(fx["bins"]))
    assert set(res) == {"geneName", "branching_probability", "branching_location", "mean", "variance", "Xnew",
                        "test_times", "MAP_model", "loglik", "logBayesFactor", "likelihood"}
    ref, ref2 = fx["loglik"], fx["loglik_pert"]
    # these small fits stop on the relative-decrease rule while still drifting: the oracle's own two runs differ
    # by up to 4e-3, which bounds what any second implementation can reproduce
    tol = np.maximum(3.0 * np.abs(ref - ref2), 2e-5 * np.abs(ref))
    assert np.all(np.abs(res["loglik"] - ref) <= tol), (res["loglik"], ref, ref2)
    assert np.allclose(res["branching_probability"], fx["branching_probability"], atol=5e-3)
    assert abs(res["logBayesFactor"] - float(fx["logBayesFactor"])) < 2e-2
    assert int(np.argmax(res["branching_probability"])) == int(np.argmax(fx["branching_probability"]))


def test_mouseob_golden():
    """The numbers the reference itself published (GPcounts_spatial.ipynb:1080,1086) for the two genes whose
    inputs are on disk, plus six genes of data/MouseOB/SE_genes_comparison.csv."""
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("mouseob_vgp_scaled")
    Xdf, Ydf, sdf = _frames(fx)
    Ydf.index = list(fx["genes"])
    sdf.columns = list(fx["genes"])
    gp = Fit_GPcounts(Xdf, Ydf, scale=sdf)
    df = gp.One_sample_test("Negative_binomial")
    got = df.values
    nb = fx["notebook_ll"]
    # 2010300C02Rik: dynamic and constant ELBO published by the reference
    assert abs(got[0, 0] - nb[0, 0]) < 5e-5, got[0]
    # Fit_GPcounts hands alpha to the constant model (:785-787); the notebook API did not -> compare with the oracle
    # Slc1a3's dynamic fit is chaotic at round-off level in the reference itself: SciPy lands on -1112.0078 or
    # on -1118.958 depending on a 1e-13 perturbation of the start; accept either optimum
    _check_table(df, fx, 2, skip={(1, 0), (1, 1)})
    assert min(abs(got[1, 0] + 1112.00784), abs(got[1, 0] + 1118.9577)) < 2e-3, got[1]
    paper = fx["paper_llr"]
    ok = ~np.isnan(paper)
    stable = np.abs(fx["ll"][:, 2] - fx["ll_pert"][:, 2]) <= 1e-5 * np.abs(fx["ll"][:, 2])
    sel = ok & stable
    sel[1] = False      # Slc1a3's dynamic fit is chaotic at round-off level in the reference itself
    assert np.all(np.abs(got[sel, 2] - paper[sel]) <= 2e-3), (got[:, 2], paper)


def test_c5_headline_sample():
    """48 genes of the benchmark workload itself (N = 1000, M = 64, sparse NB One_sample_test).

    On this workload more than half of the *dynamic* fits are round-off-unstable in the oracle itself (constant genes
    whose RBF fit drifts along a flat valley for 300-1000 iterations: the oracle's two runs differ by up to 1 nat,
    tests/golden/make_fixtures.py), so the criteria are: (1) constant fits and short, well-posed dynamic fits match
    tightly (LLR 1e-4 relative, north_star); (2) every LLR stays inside the envelope of the oracle's own
    variability; (3) >= 99 % agreement of the differential-expression calls at q < 0.05 (north_star)."""
    from gpcounts_b200 import Fit_GPcounts, host
    fx = _load("c5_sample_sparse_nb")
    Xdf, Ydf, _ = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=int(fx["M"]))
    df = gp.One_sample_test("Negative_binomial")
    got, ref, ref2 = df.values, fx["ll"], fx["ll_pert"]
    assert (gp.fit_stats["status"] == 0).all()
    # (1a) constant fits
    cst = np.abs(ref[:, 1] - ref2[:, 1]) <= 1e-6 * np.abs(ref[:, 1])
    assert cst.sum() >= 45
    assert np.all(np.abs(got[cst, 1] - ref[cst, 1]) <= LL_RTOL * np.abs(ref[cst, 1]))
    # (1b) well-posed dynamic fits: short in both oracle runs and reproduced by the oracle itself
    well = (fx["nit"][:, 0] <= 150) & (fx["nit_pert"][:, 0] <= 150) & cst & \
           (np.abs(ref[:, 0] - ref2[:, 0]) <= 1e-6 * np.abs(ref[:, 0]))
    assert well.sum() >= 8, well.sum()
    assert np.all(np.abs(got[well, 0] - ref[well, 0]) <= LL_RTOL * np.abs(ref[well, 0])), (got[well, 0], ref[well, 0])
    assert np.all(np.abs(got[well, 2] - ref[well, 2]) <= LLR_RTOL * np.abs(ref[well, 2]) + LLR_ATOL)
    # (2) all genes: the dynamic fit ends inside the set of outcomes the oracle itself produces for that gene over 32 starts
    # perturbed by k * 1e-13 (c5_sample_sparse_nb_runs.npz), and so does the LLR.  For 18 of the 48 genes that set is a
    # continuum up to 2 nats wide (constant genes whose RBF fit creeps along a flat valley until SciPy's relative-decrease
    # rule fires; where it fires is decided by round-off).  A fresh draw from a continuous distribution falls outside the range
    # of 34 samples with probability 2 / 35, i.e. about one of the 17 continuum genes per run (P(>= 4) = 2 %): up to three
    # continuum genes may lie outside their envelope, by no more than the widest envelope of the fixture (2.2 nats - the scale
    # of the valley noise; a restart or a premature stop lands 10+ nats away); every other gene lies inside.  The
    # distributional criteria on 2000 genes are in test_gpu_fit_large.py.
    runs = np.c_[_load("c5_sample_sparse_nb_runs")["ll_runs"], ref[:, 0], ref2[:, 0]]
    tol0 = 1e-4 * np.abs(ref[:, 0])
    width = runs.max(axis=1) - runs.min(axis=1)
    lo, hi = runs.min(axis=1) - tol0, runs.max(axis=1) + tol0
    continuum = width > 0.05
    assert 10 <= continuum.sum() <= 24, continuum.sum()
    outside = (got[:, 0] < lo) | (got[:, 0] > hi)
    assert not (outside & ~continuum).any(), np.c_[got[:, 0], lo, hi][outside & ~continuum]
    assert outside.sum() <= 3, np.c_[got[:, 0], lo, hi][outside]
    assert np.all((got[:, 0] >= lo - width.max()) & (got[:, 0] <= hi + width.max())), np.c_[got[:, 0], lo, hi, width][outside]
    cst_err = np.abs(got[:, 1] - ref[:, 1])
    ins = ~outside
    assert np.all((got[ins, 2] >= (lo - ref[:, 1] - cst_err - 1e-9)[ins]) & (got[ins, 2] <= (hi - ref[:, 1] + cst_err + 1e-9)[ins]))
    spread = np.abs(ref[:, 2] - ref2[:, 2])
    err = np.abs(got[:, 2] - ref[:, 2])
    assert np.median(err) <= max(3 * np.median(spread), 2e-3), (np.median(err), np.median(spread))
    # (3) differential-expression calls: >= 99 % is the north_star figure for whole-genome runs; on 48 genes a single
    # basin flip is already 2 %, so at most one disagreeing call is allowed here
    q = host.qvalue(host.p_values(np.maximum(got[:, 2], 0)))
    q_ref = host.qvalue(host.p_values(np.maximum(ref[:, 2], 0)))
    assert ((q < 0.05) != (q_ref < 0.05)).sum() <= 1
    assert (q_ref < 0.05).sum() >= 10


def test_per_gene_inducing_mode_runs_and_matches_reference_selection():
    """inducing_per_gene=True reproduces the reference's gene-specific RobustGP selection (oracle without the
    inducing_variance knob) on the genes whose selection is tie-free."""
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("one_sample_sparse_nb")
    Xdf, Ydf, _ = _frames(fx)
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=int(fx["M"]))
    gp.inducing_per_gene = True
    df = gp.One_sample_test(str(fx["lik"]))
    assert (gp.fit_stats["status"] == 0).all() and np.isfinite(df.values).all()
    # shared-Z and per-gene-Z results differ at most at the level of two equally good inducing sets
    ref = fx["ll"]
    assert np.all(np.abs(df.values[:, 1] - ref[:, 1]) <= 1e-3 * np.abs(ref[:, 1]))


def test_failed_fits_follow_the_restart_protocol():
    """maxiter=1 can never end in CONVERGENCE, so every attempt is `success == False`: the reference restarts ten times
    from the seeded random hyper-parameters (:699-706, :769-779) and then reports NaN (:533-535, :455-457)."""
    from gpcounts_b200 import Fit_GPcounts
    from oracle.fit import OracleFit
    from tests.helpers import synth_counts
    X, Y = synth_counts(3, 30, seed=8)
    cells = ["c%d" % i for i in range(30)]
    gp = Fit_GPcounts(pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["a", "b", "c"], columns=cells))
    gp.optimizer_options = {"maxiter": 1}
    df = gp.One_sample_test("Negative_binomial")
    assert df.isna().all().all()
    st = gp.fit_stats
    dyn = st[st.model == "dynamic"]
    assert (dyn.status == 2).all() and (dyn.restarts == 10).all()
    assert (st[st.model == "constant"].status == 3).all()          # skipped: the dynamic fit failed (:442)
    o = OracleFit(X, Y, lbfgs_options=dict(maxiter=1))
    r, _ = o.fit_model(X, Y[0].astype(int).astype(float), None, "NB", 1, 2)
    assert r.restarts == 10 and np.isnan(r.log_likelihood)


def test_branching_gaussian_runs():
    from gpcounts_b200 import Fit_GPcounts
    fx = _load("branching_nb")
    t, labels, Y = fx["t"], fx["labels"], fx["Y"]
    cells = ["c%d" % i for i in range(len(t))]
    gp = Fit_GPcounts(pd.DataFrame(t, index=cells), pd.DataFrame(Y, index=["gene"], columns=cells))
    res = gp.Infer_branching_location(labels, bins_num=5, lik_name="Gaussian")
    assert np.isfinite(res["loglik"]).all() and abs(res["branching_probability"].sum() - 1) < 1e-12
