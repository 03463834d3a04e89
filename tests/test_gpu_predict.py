"""Batched posterior prediction (gpc_predict) vs the oracle's restatement of gpflow ``predict_f``: mean, variance and full
covariance at new inputs for VGP / SVGP (RBF, Constant, BranchKernel) and GPR, 1e-8 relative."""
import numpy as np
import pytest

from tests.helpers import random_theta, synth_counts

pytestmark = pytest.mark.gpu
TOL = 1e-8


@pytest.mark.parametrize("kind,kernel,lik,N,M", [("VGP", "RBF", "NB", 70, 0), ("VGP", "CONST", "NB", 30, 0),
                                                   ("SVGP", "RBF", "NB", 150, 20), ("SVGP", "RBF", "ZINB", 300, 70),
                                                   ("VGP", "BRANCH", "NB", 50, 0), ("GPR", "RBF", "GAUSS", 70, 0)])
def test_predict_matches_oracle(kind, kernel, lik, N, M):
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    rng = np.random.default_rng(N + M)
    D, T = 3, 41
    X, Y = synth_counts(D, N, seed=N)
    xp = -1000.0
    if kernel == "BRANCH":
        X = np.c_[X[:, 0], rng.integers(1, 3, N).astype(float)]
        xp = 0.4
        X[X[:, 0] <= xp, 1] = 1.0
        Xnew = np.c_[np.linspace(0, 1, T), np.where(np.arange(T) % 2 == 0, 1.0, 2.0)]
        Xnew[Xnew[:, 0] <= xp, 1] = 1.0
    else:
        Xnew = np.linspace(X.min() - 0.1, X.max() + 0.1, T)[:, None]
    Z = X[np.sort(rng.choice(N, M, replace=False))].copy() if kind == "SVGP" else None
    if lik == "GAUSS":
        Y = np.log(Y + 1)
    spec = gm.ModelSpec(kind=kind, kernel=kernel, lik=lik, X=X, Z=Z, xp=xp)
    theta = random_theta(spec, rng, D)
    model = engine.Model(kind, kernel, lik, X, Z=Z, xp=xp)
    mean, var, cov = engine.predict(model, theta, Xnew, Y=Y, full_cov=True)
    mean, var, cov = mean.cpu().numpy(), var.cpu().numpy(), cov.cpu().numpy()
    for g in range(D):
        m_ref, c_ref = gm.predict_f(theta[g], spec, Xnew, y=Y[g], full_cov=True)
        _, v_ref = gm.predict_f(theta[g], spec, Xnew, y=Y[g], full_cov=False)
        assert np.abs(mean[g] - m_ref).max() <= TOL * max(np.abs(m_ref).max(), 1e-300)
        assert np.abs(var[g] - v_ref).max() <= TOL * np.abs(v_ref).max()
        assert np.abs(cov[g] - c_ref).max() <= TOL * np.abs(c_ref).max()
        assert np.abs(np.diag(cov[g]) - var[g]).max() <= 1e-12 * np.abs(var[g]).max()
