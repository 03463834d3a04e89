"""The callers either side of the fit path: fitted-model store + ``load_predict_models``, the
posterior-predictive half of ``Infer_branching_location``, ``safe_mode`` and ``Model_selection_test`` - run through the
drop-in API on the GPU.  Predictions of the latent function are checked against the oracle at the fitted parameters; the
Monte-Carlo predictive quantities only statistically (they are stochastic in the reference too)."""
import os

import numpy as np
import pandas as pd
import pytest

from tests.helpers import synth_counts

pytestmark = pytest.mark.gpu


def _frames(D, N, seed):
    X, Y = synth_counts(D, N, seed=seed)
    cells = ["c%d" % i for i in range(N)]
    return X, Y, pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(D)], columns=cells)


def test_load_predict_models_after_one_sample_test(tmp_path):
    from gpcounts_b200 import Fit_GPcounts, predictive
    X, Y, Xdf, Ydf = _frames(4, 60, 3)
    gp = Fit_GPcounts(Xdf, Ydf, sparse=True, M=8)
    gp.folder_name = str(tmp_path) + "/"
    gp.persist_models = True
    res = gp.One_sample_test("Negative_binomial")
    assert os.path.exists(gp.get_file_name("One_sample_test"))
    params = gp.load_predict_models(["g2", "g0"], "One_sample_test", "Negative_binomial")
    assert params["test_name"] == "One_sample_test" and params["likelihood"] == "Negative_binomial"
    assert len(params["means"]) == 2 and len(params["means"][0]) == 2                      # genes x models
    for gi, g in enumerate((2, 0)):
        for k in range(2):
            mean, var, model = params["means"][gi][k], params["vars"][gi][k], params["models"][gi][k]
            assert np.asarray(mean).shape == (100,) and var.shape == (predictive.ROUNDS * predictive.Y_PER_POINT, 100)
            assert np.all(np.asarray(mean) >= 0) and np.all(var == np.round(var))
            # the predictive mean tracks the data: same order of magnitude as the gene's average count
            assert 0.3 * (Y[g].mean() + 0.1) < np.mean(mean) < 3.0 * (Y[g].mean() + 0.1), (g, k, np.mean(mean), Y[g].mean())
            assert abs(float(model.log_posterior_density().numpy()) - res.values[g, k]) < 1e-12
            assert np.isfinite(model.likelihood.alpha.numpy()) and model.inducing_variable.Z.numpy().shape == (8, 1)
    with pytest.raises(RuntimeError):
        gp.load_predict_models(["g0"], "Two_samples_test", "Negative_binomial")


def test_gaussian_prediction_matches_oracle_at_the_fitted_parameters():
    from gpcounts_b200 import Fit_GPcounts
    from oracle import gp_math as gm
    X, Y, Xdf, Ydf = _frames(3, 50, 5)
    gp = Fit_GPcounts(Xdf, Ydf)
    gp.Infer_trajectory("Gaussian")
    params = gp.load_predict_models(["g1"], "Infer_trajectory", "Gaussian")
    mean, var, model = params["means"][0][0], params["vars"][0][0], params["models"][0][0]
    assert mean.shape == (100, 1) and var.shape == (100, 1)
    spec = gm.ModelSpec(kind="GPR", kernel="RBF", lik="GAUSS", X=X)
    xtest = np.linspace(X.min() - 0.1, X.max() + 0.1, 100)[:, None]
    m_ref, v_ref = gm.predict_f(model.theta, spec, xtest, y=np.log(Y[1].astype(int) + 1.0))
    noise = float(model.likelihood.variance.numpy())
    assert np.abs(mean[:, 0] - m_ref).max() <= 1e-7 * np.abs(m_ref).max()
    assert np.abs(var[:, 0] - (v_ref + noise)).max() <= 1e-7 * np.abs(v_ref + noise).max()


def test_branching_returns_the_predictive_of_the_map_model():
    from gpcounts_b200 import Fit_GPcounts, predictive
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "branching_nb.npz"), allow_pickle=True)
    t, labels, Y = fx["t"], fx["labels"], fx["Y"]
    cells = ["c%d" % i for i in range(len(t))]
    gp = Fit_GPcounts(pd.DataFrame(t, index=cells), pd.DataFrame(Y, index=["gene"], columns=cells))
    res = gp.Infer_branching_location(labels, bins_num=int(fx["bins"]))
    assert np.asarray(res["mean"]).shape == (200,) and res["variance"].shape == (predictive.ROUNDS * predictive.Y_PER_POINT, 200)
    assert 0.3 * Y.mean() < np.mean(res["mean"]) < 3.0 * Y.mean()
    gp = Fit_GPcounts(pd.DataFrame(t, index=cells), pd.DataFrame(Y, index=["gene"], columns=cells))
    res = gp.Infer_branching_location(labels, bins_num=5, lik_name="Gaussian")
    assert res["mean"].shape == (200, 1) and res["variance"].shape == (200, 1) and np.all(res["variance"] > 0)


def test_safe_mode_restarts_only_the_fits_that_fail_the_predictive_check():
    from gpcounts_b200 import Fit_GPcounts
    X, Y, Xdf, Ydf = _frames(24, 40, 11)
    plain = Fit_GPcounts(Xdf, Ydf)
    ref = plain.One_sample_test("Negative_binomial")
    gp = Fit_GPcounts(Xdf, Ydf, safe_mode=True)
    df = gp.One_sample_test("Negative_binomial")
    assert list(df.columns) == list(ref.columns) and df.shape == ref.shape
    st = gp.fit_stats
    dyn = st[st.model == "dynamic"].reset_index(drop=True)
    # fits that passed the check at the first attempt are the plain fits, bit for bit
    same = (dyn.restarts == 0).values
    assert same.sum() >= 12
    assert np.array_equal(df.values[same, 0], ref.values[same, 0])
    assert np.isfinite(df.values[:, 0]).all()
    assert (st.restarts <= 10).all()


def test_model_selection_table():
    """Reference :163-229 (see the method's docstring for what the reference's kernel switch effectively fits)."""
    from gpcounts_b200 import Fit_GPcounts
    X, Y, Xdf, Ydf = _frames(3, 40, 13)
    gp = Fit_GPcounts(Xdf, Ydf)
    out = gp.Model_selection_test("Negative_binomial")
    assert {"Gene", "Model", "BIC", "p_value", "q_value", "Dynamic_model_log_likelihood", "Constant_model_log_likelihood",
            "log_likelihood_ratio", "Linear_probability", "Periodic_probability", "RBF_probability"} <= set(out.columns)
    ref = Fit_GPcounts(Xdf, Ydf).One_sample_test("Negative_binomial")
    rbf = out[out.Model == "RBF"].set_index("Gene").loc[list(ref.index)]
    assert np.array_equal(rbf["log_likelihood_ratio"].values, ref["log_likelihood_ratio"].values)
    assert np.allclose(rbf["BIC"].values, -2 * ref["Dynamic_model_log_likelihood"].values + 4 * np.log(40))
    assert np.allclose(out[["Linear_probability", "Periodic_probability", "RBF_probability"]].sum(axis=1), 1.0)


def test_sharded_fit_equals_single_gpu_bit_for_bit():
    """genes r::G per rank, no collective on the fit path, one all_gather of the statistics - the table of a
    2-GPU run (torchrun, NCCL) equals the single-GPU table bit for bit.  Needs two visible GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "check_sharded.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded == single: True" in out.stdout, out.stdout[-2000:]
