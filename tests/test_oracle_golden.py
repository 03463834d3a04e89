"""CPU: the oracle against the reference's own published numbers (SURVEY.md section 8c pins)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pq_table_reproduces_notebook():
    """20-row p/q table of demo_notebooks/GPcounts_spatial.ipynb:1334-1376 (LLR -> p -> Storey q, pi0 = 1)."""
    from oracle.qvalue import calculate_FDR
    fx = np.load(os.path.join(GOLD, "spatial_pq_table.npz"))
    p, q = calculate_FDR(fx["llr"])
    assert np.allclose(p, fx["p"], rtol=5e-6, atol=0)      # LLR is printed with 6 decimals
    assert np.allclose(q, fx["q"], rtol=5e-6, atol=0)


def test_product_qvalue_equals_oracle():
    from gpcounts_b200 import host
    from oracle.qvalue import calculate_FDR, qvalue
    rng = np.random.default_rng(0)
    for n in (7, 99, 100, 101, 2500):
        llr = np.abs(rng.standard_normal(n)) * 3
        p, q = calculate_FDR(llr)
        assert np.array_equal(host.p_values(llr), p)
        assert np.allclose(host.qvalue(p), q, rtol=1e-13, atol=0)
    pv = rng.uniform(0, 1, 300)
    assert np.allclose(host.qvalue(pv, pi0=0.7), qvalue(pv, pi0=0.7), rtol=1e-13)


def test_robustgp_first_inducing_point():
    """Z[0] = 0.71943149 on data/alpha_time_points.csv, seed 0 (scRNA-Seq_time_series.ipynb:670), M = 16 and M = 4."""
    from gpcounts_b200 import host
    from oracle.robustgp import select_Z
    X = np.load(os.path.join(GOLD, "alpha_time_points.npy"))
    ls = 0.05 * (X.max() - X.min())
    for M in (16, 4):
        np.random.seed(0)
        Z = select_Z(X, M, "RBF", 2.5, ls)
        assert abs(Z[0, 0] - 0.71943149) < 5e-9
        Zp = host.inducing_sets(X, M, ls, with_restarts=False)[0]
        assert np.array_equal(Z, Zp)


def test_golden_mouseob_constant_fits():
    """Two of the four log-likelihoods the reference published for genes whose inputs are on disk
    (GPcounts_spatial.ipynb:1080,1086): the constant-model fits (the dynamic fits take minutes on the CPU and are
    checked when the fixtures are generated, tests/golden/make_fixtures.py)."""
    from oracle.fit import OracleFit
    fx = np.load(os.path.join(GOLD, "mouseob_vgp_scaled.npz"), allow_pickle=True)
    o = OracleFit(fx["X"], fx["Y"], scale=fx["scale"], alpha_policy="one")
    Y = fx["Y"].astype(int).astype(float)
    for g in (0, 1):
        r, _ = o.fit_model(fx["X"], Y[g], fx["scale"][:, g], "NB", 2, 2)
        assert abs(r.log_likelihood - fx["notebook_ll"][g, 1]) < 2e-5, (g, r.log_likelihood)
    # and the stored oracle runs reproduce the published dynamic values
    assert abs(fx["ll_alpha_one"][0, 0] - fx["notebook_ll"][0, 0]) < 2e-5
    assert abs(fx["ll_alpha_one"][1, 0] - fx["notebook_ll"][1, 0]) < 2e-5
    assert abs(fx["ll_alpha_one"][0, 2] - fx["notebook_ll"][0, 2]) < 5e-5


def test_svgp_with_all_points_matches_vgp():
    """Consistency of restatements 3.4.3 vs 3.4.4: SVGP with Z = X equals VGP up to the jitter placement."""
    from oracle import gp_math as gm
    from tests.helpers import synth_counts
    X, Y = synth_counts(1, 25, seed=2)
    sv = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=X.copy())
    vg = gm.ModelSpec(kind="VGP", kernel="RBF", lik="NB", X=X)
    th = gm.initial_theta(vg, 2.0, 0.3, 0.7)
    th[4:4 + 25] = 0.1 * np.random.default_rng(0).standard_normal(25)
    f1, _ = gm.elbo_and_grad(th, vg, Y[0])
    f2, _ = gm.elbo_and_grad(th, sv, Y[0])
    assert abs(f1 - f2) < 1e-4 * abs(f1)


def test_poisson_quadrature_equals_closed_form():
    from oracle import gp_math as gm
    import torch
    fm = torch.tensor([0.3, -1.2, 2.0]); fv = torch.tensor([0.5, 0.1, 0.02]); y = torch.tensor([1.0, 0.0, 9.0])
    closed = (y * fm - torch.exp(fm + fv / 2) - torch.lgamma(y + 1)).sum()
    F = fm[None] + torch.sqrt(fv)[None] * gm.GH_Z[:, None]
    quad = (gm.GH_W[:, None] * (y[None] * F - torch.exp(F) - torch.lgamma(y + 1)[None])).sum()
    assert abs(float(closed - quad)) < 1e-10


def test_perturbed_oracle_runs_are_consistent_with_the_fixture():
    """tests/golden/make_perturbed_runs.py: run k = 0 is the fixture's primary run, run k = 10 (1e-12) its perturbed one;
    the outcome set of a fit is small (a few plateau values), which is what the GPU parity tests rely on."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fx = np.load(os.path.join(gold, "one_sample_sparse_nb.npz"), allow_pickle=True)
    runs = np.load(os.path.join(gold, "one_sample_sparse_nb_runs.npz"))["ll_runs"]
    assert runs.shape == (fx["ll"].shape[0], 32)
    assert np.array_equal(runs[:, 0], fx["ll"][:, 0])
    assert np.array_equal(runs[:, 10], fx["ll_pert"][:, 0])
    spread = runs.max(1) - runs.min(1)
    assert (spread < 1e-9).sum() >= 2           # well-posed fits reproduce to all digits
    assert spread.max() < 0.01                  # the others move by millinats: early plateau stops


def test_nb_log_density_ieee_classes_follow_the_reference_formula():
    """The oracle evaluates NegativeBinomialLikelihood.negative_binomial() literally (reference :37-48), so that exp(F)
    overflow / underflow and alpha -> 0 give the IEEE class the reference's TensorFlow graph gives; the CUDA evaluators are
    held to these classes by tests/test_gpu_elbo_parity.py.  NumPy restatement of the same line:"""
    import torch
    from oracle.gp_math import nb_logp
    with np.errstate(all="ignore"):
        def ref(m, y, alpha):
            from scipy.special import gammaln
            k = 1.0 / alpha
            return gammaln(k + y) - gammaln(y + 1) - gammaln(k) + y * np.log(m / (m + k)) - k * np.log(1 + m * alpha)
        cases = [(np.inf, 3.0, 0.5, "nan"), (np.inf, 0.0, 0.5, "nan"), (0.0, 3.0, 0.5, "-inf"), (0.0, 0.0, 0.5, "nan"),
                 (5.0, 3.0, 3e-40, "finite"), (5.0, 0.0, 3e-40, "finite")]
        for m, y, alpha, want in cases:
            a = float(nb_logp(torch.tensor(m, dtype=torch.float64), torch.tensor(y, dtype=torch.float64),
                              torch.tensor(alpha, dtype=torch.float64)))
            b = float(ref(np.float64(m), np.float64(y), np.float64(alpha)))
            cls = lambda v: "nan" if np.isnan(v) else ("-inf" if v == -np.inf else ("+inf" if v == np.inf else "finite"))
            assert cls(a) == cls(b) == want, (m, y, alpha, a, b)
        # alpha -> 0: k + y == k, the lgamma terms cancel exactly and lgamma(y + 1) is absorbed (left-to-right evaluation)
        a = float(nb_logp(torch.tensor(5.0, dtype=torch.float64), torch.tensor(3.0, dtype=torch.float64),
                          torch.tensor(3e-40, dtype=torch.float64)))
        assert abs(a - 3.0 * np.log(5.0 * 3e-40)) < 1e-9 * abs(a)
