"""CPU: host-side logic of the drop-in (hyper-parameter tables, sharding, API surface)."""
import os
import sys

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restart_table_equals_reference_draws():
    """Restart r reseeds NumPy with r and draws ls, var, alpha, km (GPcounts_Module.py:769-779)."""
    from gpcounts_b200 import host
    X = np.random.default_rng(1).uniform(3, 9, (50, 2))
    tab, _ = host.restart_table(X)
    rng = X.max() - X.min()
    for r in range(1, 11):
        np.random.seed(r)
        ls = np.random.uniform(0.25 * rng / 100, 30.0 * rng / 100)
        var, alpha, km = np.random.uniform(0, 10), np.random.uniform(0, 10), np.random.uniform(0, 100)
        assert np.array_equal(tab[r - 1], [ls, var, alpha, km])


def test_restart_table_matches_oracle_restart_path():
    """Force the oracle into restarts (maxiter=1 is never 'success') and compare the hyper-parameters it starts from."""
    from gpcounts_b200 import host
    from oracle.fit import OracleFit
    from tests.helpers import synth_counts
    X, Y = synth_counts(1, 12, seed=0)
    o = OracleFit(X, Y, lbfgs_options=dict(maxiter=1))
    r, _ = o.fit_model(X, Y[0], None, "NB", 1, 1)
    assert r.restarts == 10 and not r.ok
    tab, _ = host.restart_table(X)
    assert np.isnan(r.log_likelihood)
    assert tab.shape == (10, 4)


def test_default_hyper_parameters():
    from gpcounts_b200 import host
    X = np.c_[np.linspace(7.888, 28.047, 20), np.linspace(8, 20, 20)]
    assert abs(host.default_lengthscale(X) - 0.05 * (28.047 - 7.888)) < 1e-12     # joint range quirk (:730)
    Y = np.array([[0.0, 1.0, 5.0], [2.0, 2.0, 2.0]])
    v = host.default_variance(Y, "Negative_binomial", True)
    assert np.allclose(v, np.mean(np.log(Y + 1) ** 2, axis=1))
    assert np.allclose(host.default_variance(Y, "Gaussian", False), np.mean(Y + 1, axis=1))   # :736 precedence quirk


def test_branching_evidence_matches_oracle():
    from gpcounts_b200 import host
    from oracle.fit import calculate_branching_evidence
    ll = np.array([-152.6, -152.5, -153.3, -153.9, -153.8])
    t = np.linspace(0, 1, 5)
    a, b = host.branching_evidence(ll, t), calculate_branching_evidence(ll, t)
    assert np.allclose(a["posteriorBranching"], b["posteriorBranching"]) and a["logBayesFactor"] == b["logBayesFactor"]
    with pytest.raises(NameError):
        host.branching_evidence(ll, t[:-1])


def test_shard_indices_partition():
    from gpcounts_b200 import sharding
    for n in (0, 1, 7, 20000):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.array([])
            assert np.array_equal(allidx, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _gather_worker(rank, world, port, n, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from gpcounts_b200 import sharding
    idx = sharding.shard_indices(n, rank, world)
    local = torch.tensor(idx, dtype=torch.float64)[:, None, None] * torch.ones(1, 2, 16, dtype=torch.float64)
    full = sharding.gather_rows(local, n, rank, world)
    ok = bool(torch.equal(full[:, 0, 0], torch.arange(n, dtype=torch.float64))) and full.shape == (n, 2, 16)
    if rank == 0:
        open(out, "w").write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 4])
def test_gather_rows_world_size_2_gloo(tmp_path, n):
    """The only collective of the path (final statistics gather), world_size 2 on the gloo backend."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.txt")
    port = 29500 + (os.getpid() % 1000) + n
    mp.spawn(_gather_worker, args=(2, port, n, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_api_surface_matches_reference():
    from gpcounts_b200 import Fit_GPcounts
    import inspect
    sig = inspect.signature(Fit_GPcounts.__init__)
    assert list(sig.parameters)[:7] == ["self", "X", "Y", "scale", "sparse", "M", "safe_mode"]
    for name in ["Infer_trajectory", "One_sample_test", "Two_samples_test", "Infer_branching_location",
                 "calculate_FDR", "qvalue", "initialize_hyper_parameters", "run_test", "set_X_Y",
                 "CalculateBranchingEvidence", "kmean_algorithm_inducing_points"]:
        assert callable(getattr(Fit_GPcounts, name)), name
    assert inspect.signature(Fit_GPcounts.One_sample_test).parameters["lik_name"].default == "Negative_binomial"
    assert inspect.signature(Fit_GPcounts.Infer_branching_location).parameters["bins_num"].default == 50
    X = pd.DataFrame(np.linspace(0, 1, 40), index=["c%d" % i for i in range(40)])
    Y = pd.DataFrame(np.ones((2, 40)) * 2.7, index=["a", "b"], columns=X.index)
    gp = Fit_GPcounts(X, Y, sparse=True)
    assert gp.M == 2 and gp.N == 40 and gp.D == 2 and gp.genes_name == ["a", "b"]      # M = 5 % of N (:121-125)
    df = pd.DataFrame({"log_likelihood_ratio": [0.253854, 22.315595]})
    out = gp.calculate_FDR(df)
    assert abs(out["p_value"][1] - 2.378719e-11) < 1e-16
    assert Fit_GPcounts(X, Y, safe_mode=True).safe_mode is True
    for name in ["load_predict_models", "get_file_name", "Model_selection_test"]:
        assert callable(getattr(Fit_GPcounts, name)), name


def test_per_gene_inducing_selection_matches_host_recurrence():
    """gpcounts_b200.inducing (batched torch recurrence) vs the host recurrence on a tie-free configuration."""
    from gpcounts_b200 import host, inducing
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 1, (60, 1))
    ls = 0.3
    var = np.array([0.7, 2.9, 11.3])
    Z = inducing.select_per_gene(X, 12, ls, var, "cpu").numpy()
    for g in range(3):
        np.random.seed(0)
        Zh = host.greedy_conditional_variance(X, 12, ls, var[g])
        assert np.array_equal(np.sort(Z[g].ravel()), np.sort(Zh.ravel()))


def test_qvalue_with_nan_rows_follows_the_reference():
    """A failed fit gives a NaN LLR (:455-457); the reference's qvalue receives a pandas Series whose min()/max() skip
    NaN, keeps m = len(pv), sorts NaN last and returns NaN only for those rows."""
    from gpcounts_b200 import Fit_GPcounts, host
    llr = np.array([12.0, 0.3, np.nan, 3.1, 0.0])
    df = pd.DataFrame({"log_likelihood_ratio": llr}, index=list("abcde"))
    gp = Fit_GPcounts.__new__(Fit_GPcounts)
    out = Fit_GPcounts.calculate_FDR(gp, df)
    p = out["p_value"].values
    assert np.isnan(p[2]) and np.isnan(out["q_value"].values[2])
    ok = ~np.isnan(p)
    order = np.argsort(p[ok])
    ps = p[ok][order]
    q = np.zeros(4)
    q[-1] = 5.0 * ps[-1] / 4.0            # m = 5 (NaN counted), rank 4 is the last finite one
    for i in range(2, -1, -1):
        q[i] = min(5.0 * ps[i] / (i + 1.0), q[i + 1])
    assert np.allclose(out["q_value"].values[ok][order], q)
    assert np.isnan(host.qvalue(np.array([np.nan, np.nan]))).all()


def test_default_variance_is_independent_of_batch_shape_and_memory_order():
    """`DataFrame.values` is Fortran-ordered; the per-gene initial variance must still be the reference's per-gene
    contiguous reduction (:421-422, :738), bit for bit, whatever the batch a gene arrives in."""
    from gpcounts_b200 import host
    rng = np.random.default_rng(0)
    Y = rng.poisson(7.0, (101, 20)).astype(float)
    ref = np.array([np.mean(np.log(Y[g].astype(float).reshape([-1, 1]) + 1) ** 2) for g in range(101)])
    Yf = pd.DataFrame(Y).values if False else np.asfortranarray(Y)
    assert np.array_equal(host.default_variance(Yf, "Negative_binomial", True), ref)
    assert np.array_equal(host.default_variance(Y, "Negative_binomial", True), ref)
    assert host.default_variance(Yf[5:6], "Negative_binomial", True)[0] == ref[5]
