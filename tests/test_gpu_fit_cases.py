"""Fit-path cases the round-1 suite left uncovered (VERDICT r1 "what's weak" item 14): restart triggers
inside ``gpc_fit`` (non-finite start, Cholesky failure), sparse ``Two_samples_test``, VGP with ZINB / Poisson,
per-spot scale with SVGP, job counts that are not a multiple of the slot count with ``chain=False``.

Small problems, so the CPU oracle (SciPy L-BFGS-B on the torch-FP64 restatement) is run inside the test from the exact
start and from a start perturbed by 1e-12: where its two runs agree the GPU must match to 5e-6 relative, elsewhere to
the oracle's own spread (see tests/test_gpu_fit_parity.py)."""
import numpy as np
import pandas as pd
import pytest

from tests.helpers import synth_counts

pytestmark = pytest.mark.gpu
LL_RTOL = 5e-6


def _frames(X, Y, scale=None):
    cells = ["c%d" % i for i in range(X.shape[0])]
    genes = ["g%d" % i for i in range(Y.shape[0])]
    return (pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=genes, columns=cells),
            None if scale is None else pd.DataFrame(scale, index=cells))


def _oracle(method, lik, X, Y, scale=None, user=None, **kw):
    from oracle import fit as ofit
    out = []
    for pert in (0.0, 1e-12):
        ofit.PERTURB = pert
        o = ofit.OracleFit(X, Y, scale=scale, **kw)
        if user:
            o.initialize_hyper_parameters(**user)
        out.append(getattr(o, method)(lik))
    ofit.PERTURB = 0.0
    return out


NOISE_ALPHA = 1e-9


def _compare(df, rows, rows_pert, K, stats=None):
    """`stats` (gp.fit_stats): fits that end at alpha < 1e-9 are in the regime where the reference's own objective is
    rounding noise - lgamma(k + y) - lgamma(k) at k = 1 / alpha > 1e9 is a difference of numbers above 1e10 (DESIGN.md
    section 4.6; a "Poisson-like" half of a two-sample split drifts there).  The optimiser then climbs the noise - the
    reference's safe_mode exists to catch the positive log-likelihoods this produces - and neither implementation
    reproduces the other or itself.  At most one such fit per test is tolerated, and it must not end below the oracle."""
    got = df.values
    noisy = 0
    for g, (r, rp) in enumerate(zip(rows, rows_pert)):
        for k in range(K):
            a, b = r["ll"][k], rp["ll"][k]
            tol = LL_RTOL * abs(a) if abs(a - b) <= 1e-6 * abs(a) else max(10.0 * abs(a - b), 3e-4 * abs(a))
            if stats is not None and abs(got[g, k] - a) > tol:
                alpha = float(stats.alpha.values[g * K + k])
                if alpha < NOISE_ALPHA and got[g, k] >= a - tol:
                    noisy += 1
                    continue
            assert abs(got[g, k] - a) <= tol, (g, k, got[g, k], a, b)
    assert noisy <= 1, noisy


def test_all_zero_gene_restarts_like_the_reference():
    """var0 = mean(log(y+1)^2) = 0 for an all-zero gene -> softplus^-1(0) = -inf: gpflow raises InvalidArgumentError and
    fit_GP takes random restart 1 (:547-556).  Poisson used to report PGTOL "success" at the start point."""
    from gpcounts_b200 import Fit_GPcounts
    X, Y = synth_counts(3, 40, seed=4)
    Y[1] = 0.0
    Xdf, Ydf, _ = _frames(X, Y)
    for lik in ("Poisson", "Negative_binomial"):
        gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=6)
        df = gp.One_sample_test(lik)
        st = gp.fit_stats
        z = st[st.gene == "g1"]
        assert (z[z.status != 3].restarts >= 1).all(), z
        assert (st[st.gene != "g1"].restarts == 0).all()
        if lik == "Poisson":
            rows, rows_p = _oracle("one_sample_test", lik, X, Y, sparse=True, M=6, inducing_variance=1.0)
            assert rows[1]["fits"][0].restarts == int(z.restarts.values[0]) == 1
            _compare(df, rows, rows_p, 2)
        # NB on an all-zero gene has no finite optimum (f -> -inf): the oracle's ten restarts all end in a failed
        # Cholesky factorisation and the row is NaN; only the restart protocol is asserted for it


def test_cholesky_failure_inside_the_fit_triggers_a_restart():
    """Initial variance 1e14 with a flat kernel: K + 1e-6 I has an exactly zero second pivot in FP64, the reference's
    Cholesky raises (tf.errors.InvalidArgumentError, :547) and the fit restarts from the seeded random draw."""
    from gpcounts_b200 import Fit_GPcounts
    X, Y = synth_counts(2, 30, seed=11)
    Xdf, Ydf, _ = _frames(X, Y)
    user = dict(variance=1e14, length_scale=1e9)
    gp = Fit_GPcounts(Xdf, Ydf)
    gp.initialize_hyper_parameters(**user)
    df = gp.One_sample_test("Negative_binomial")
    st = gp.fit_stats
    assert (st.restarts == 1).all() and (st.status == 0).all(), st
    assert np.isfinite(df.values).all()
    rows, rows_p = _oracle("one_sample_test", "Negative_binomial", X, Y, user=user)
    assert all(f.restarts == 1 for r in rows for f in r["fits"])
    _compare(df, rows, rows_p, 2)


def test_two_samples_sparse():
    from gpcounts_b200 import Fit_GPcounts
    X, Y = synth_counts(3, 60, seed=13)
    Xdf, Ydf, _ = _frames(X, Y)
    Xdf.columns = ["times"]
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=6)
    df = gp.Two_samples_test("Negative_binomial")
    rows, rows_p = _oracle("two_samples_test", "Negative_binomial", X, Y, sparse=True, M=6, inducing_variance=1.0)
    _compare(df, rows, rows_p, 3, stats=gp.fit_stats)
    # the model left behind is the one of the second half (:477-484)
    assert gp.model.data[0].shape[0] == 30 and gp.model.data[1].shape[0] == 30


@pytest.mark.parametrize("lik", ["Zero_inflated_negative_binomial", "Poisson"])
def test_one_sample_vgp_other_likelihoods(lik):
    from gpcounts_b200 import Fit_GPcounts
    X, Y = synth_counts(3, 30, seed=17)
    Xdf, Ydf, _ = _frames(X, Y)
    gp = Fit_GPcounts(Xdf, Ydf)
    df = gp.One_sample_test(lik)
    rows, rows_p = _oracle("one_sample_test", lik, X, Y)
    _compare(df, rows, rows_p, 2)


def test_one_sample_sparse_with_scale():
    from gpcounts_b200 import Fit_GPcounts
    X, Y = synth_counts(3, 60, seed=19)
    scale = np.random.default_rng(1).uniform(0.5, 2.0, (60, 3))
    Xdf, Ydf, sdf = _frames(X, Y, scale)
    gp = Fit_GPcounts(Xdf, Ydf, scale=sdf, sparse=True, M=8)
    df = gp.One_sample_test("Negative_binomial")
    rows, rows_p = _oracle("one_sample_test", "Negative_binomial", X, Y, scale=scale, sparse=True, M=8,
                           inducing_variance=1.0)
    _compare(df, rows, rows_p, 2)


def test_independent_jobs_not_a_multiple_of_the_slots():
    """chain=False (Two_samples_test): 3 D jobs over the persistent CTAs with 3 D not a multiple of the slot count; a
    gene's result does not depend on the batch it is fitted in (no cross-gene arithmetic), bit for bit."""
    from gpcounts_b200 import Fit_GPcounts, engine
    X, Y = synth_counts(101, 20, seed=23)
    Xdf, Ydf, _ = _frames(X, Y)
    Xdf.columns = ["times"]
    gp = Fit_GPcounts(Xdf, Ydf)
    df = gp.Two_samples_test("Negative_binomial")
    assert df.shape == (101, 4) and np.isfinite(df.values).all()
    for g in (0, 57, 100):
        gp1 = Fit_GPcounts(Xdf, Ydf.iloc[[g]])
        d1 = gp1.Two_samples_test("Negative_binomial")
        assert np.array_equal(d1.values[0], df.values[g]), (g, d1.values[0], df.values[g])


def test_restart_table_needs_matching_inducing_sets():
    """A C-ABI caller passing a restart table with 1 < n_zsets < 11 is rejected instead of reading out of bounds."""
    from gpcounts_b200 import engine, host
    from gpcounts_b200._lib import GpcError
    X, Y = synth_counts(2, 30, seed=2)
    Z = host.inducing_sets(X, 5, 0.05)[:3]
    tab, _ = host.restart_table(X)
    m = engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=tab)
    h = np.ones((2, 1, 4))
    with pytest.raises(GpcError, match="n_zsets"):
        engine.fit([m], Y, None, h)
