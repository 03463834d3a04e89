"""Fixed-theta parity at the shapes of BASELINE.json's configurations (1e-8 relative, north_star), on the reference's
own inputs where they are on disk (tests/golden/data_*.npz, extracted by tests/golden/make_data_fixtures.py):

  C1  fission yeast time course      full VGP, N = 18 (wild type) and N = 36 (both strains), counts up to ~1e6
  C2  MouseHSC pseudotime            sparse SVGP, N = 2660, M = 133 (the reference's default 5 % of N), NB and ZINB
  C3  MouseOB spatial                full VGP, N = 260, 2-D RBF, per-spot scale factors
  C4  MouseHSC branching             full VGP + BranchKernel, N = 532 (every 5th cell), free and frozen kernel
  C5  synthetic sweep                sparse NB, (N, M) = (1000, 64), (1000, 256), (10000, 64)

Hyper-parameters are drawn around the reference's initialisation (lengthscale = 5 % of the input range, :730), q_mu and
q_sqrt at random; additionally C5 is checked at a fitted optimum returned by the device optimiser.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-8
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _theta(spec, rng, D, ls0, var_c=1.0):
    from oracle.gp_math import softplus_inv
    P, mv = spec.P, spec.Mv
    th = np.zeros((D, P))
    th[:, 0] = softplus_inv(var_c * rng.uniform(0.5, 3.0, D))
    th[:, 1] = softplus_inv(ls0 * rng.uniform(0.6, 2.5, D))
    th[:, 2] = rng.uniform(-2.0, 1.0, D)
    th[:, 3] = rng.uniform(1.0, 4.0, D)
    th[:, 4:4 + mv] = 0.5 * rng.standard_normal((D, mv))
    tri = (0.3 / np.sqrt(mv)) * rng.standard_normal((D, mv * (mv + 1) // 2))
    diag = np.array([i * (i + 1) // 2 + i for i in range(mv)])
    tri[:, diag] = rng.uniform(0.3, 1.2, (D, mv))
    th[:, 4 + mv:] = tri
    return th


def _check(kind, kernel, lik, X, Y, Z=None, scale=None, theta=None, xp=-1000.0, train=(True, True, True, True),
           ls0=None, seed=0, tol=TOL):
    from gpcounts_b200 import engine, host
    from oracle import gp_math as gm
    rng = np.random.default_rng(seed)
    D = Y.shape[0]
    spec = gm.ModelSpec(kind=kind, kernel=kernel, lik=lik, X=X, Z=Z, xp=xp, train=train)
    if theta is None:
        theta = _theta(spec, rng, D, host.default_lengthscale(X[:, :1] if kernel == "BRANCH" else X) if ls0 is None else ls0)
    model = engine.Model(kind, kernel, lik, X, Z=Z, train=train, xp=xp)
    elbo, grad, status = engine.elbo_grad(model, Y, scale, theta)
    elbo, grad, status = elbo.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy()
    assert (status == 0).all(), status
    mv = spec.Mv
    for g in range(D):
        f_ref, g_ref = gm.elbo_and_grad(theta[g], spec, Y[g], None if scale is None else scale[g])
        assert abs(elbo[g] - f_ref) <= tol * abs(f_ref), (g, elbo[g], f_ref)
        for b in (slice(0, 4), slice(4, 4 + mv), slice(4 + mv, spec.P)):
            ref = g_ref[b]
            err = np.abs(grad[g][b] - ref).max() / max(np.abs(ref).max(), 1e-300)
            assert err <= tol, (kind, kernel, lik, g, b, err)


@pytest.mark.parametrize("N", [18, 36])
@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_c1_fission_vgp(N, kernel):
    d = np.load(os.path.join(GOLD, "data_fission.npz"), allow_pickle=True)
    X = d["minute"][:N, None]
    Y = d["Y"][[0, 7, 24], :N].astype(int).astype(float)
    _check("VGP", kernel, "NB", X, Y, seed=N)


@pytest.mark.parametrize("lik", ["NB", "ZINB"])
@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_c2_mousehsc_sparse(lik, kernel):
    from gpcounts_b200 import host
    d = np.load(os.path.join(GOLD, "data_mousehsc.npz"), allow_pickle=True)
    X = d["pseudotime"][:, None]
    M = int(5 * len(X) / 100)                                    # the reference's default, :121-125
    assert (len(X), M) == (2660, 133)
    Z = host.inducing_sets(X, M, host.default_lengthscale(X), with_restarts=False)[0]
    Y = d["Y"][[0, 10]]
    _check("SVGP", kernel, lik, X, Y, Z=Z, seed=3)


@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_c3_mouseob_vgp_scaled_2d(kernel):
    d = np.load(os.path.join(GOLD, "data_mouseob.npz"), allow_pickle=True)
    X, Y, scale = d["X"], d["Y"][[0, 100, 300]].astype(float), np.ascontiguousarray(d["scale"][:, [0, 100, 300]].T)
    assert X.shape == (260, 2)
    _check("VGP", kernel, "NB", X, Y, scale=scale, seed=5)


@pytest.mark.parametrize("frozen", [False, True])
def test_c4_branching_vgp_532(frozen):
    d = np.load(os.path.join(GOLD, "data_mousehsc.npz"), allow_pickle=True)
    label = np.where(d["curve1"] > 0.8, 1.0, 2.0)                # Branching_GPcounts.ipynb cell 5
    t, lab = d["pseudotime"][1::5], label[1::5].copy()
    assert len(t) == 532
    xp = -1000.0
    if frozen:
        xp = float(np.linspace(t.min(), t.max(), 25)[9])         # one candidate of the 25-bin grid (:271-286)
        lab[t <= xp] = 1.0
    X = np.c_[t, lab]
    Y = d["Y"][[0, 4], 1::5]                                     # Mpo, Prtn3
    train = (False, False, True, True) if frozen else (True, True, True, True)
    _check("VGP", "BRANCH", "NB", X, Y, xp=xp, train=train, ls0=1.0, seed=7)


@pytest.mark.parametrize("N,M", [(1000, 64), (1000, 256), (10000, 64)])
@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_c5_sparse_nb(N, M, kernel):
    from bench import synth_batch
    from gpcounts_b200 import host
    X, Y = synth_batch(2 if N > 1000 or M > 64 else 3, N, seed=77)
    Z = host.inducing_sets(X, M, host.default_lengthscale(X), with_restarts=False)[0]
    _check("SVGP", kernel, "NB", X, Y, Z=Z, seed=N + M)


def test_c5_at_a_fitted_optimum():
    """ELBO / gradient parity where the optimiser ends (structured q_sqrt, small gradients): the gradient is compared
    on the scale of the ELBO's natural units, |g_gpu - g_ref| <= 1e-8 max(|g_ref|_inf, 1)."""
    from bench import synth_batch
    from gpcounts_b200 import engine, host
    from oracle import gp_math as gm
    N, M = 1000, 64
    X, Y = synth_batch(4, N, seed=99)
    ls0 = host.default_lengthscale(X)
    Z = host.inducing_sets(X, M, ls0, with_restarts=False)
    model = engine.Model("SVGP", "RBF", "NB", X, Z=Z)
    h = np.zeros((4, 1, 4))
    h[:, 0, 0] = host.default_variance(Y, "Negative_binomial", True)
    h[:, 0, 1], h[:, 0, 2], h[:, 0, 3] = ls0, 1.0, 35.0
    stats, theta = engine.fit([model], Y, None, h, chain=True)
    theta = theta[:, 0, :model.P].cpu().numpy()
    spec = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Z[0])
    elbo, grad, status = engine.elbo_grad(model, Y, None, theta)
    for g in range(4):
        f_ref, g_ref = gm.elbo_and_grad(theta[g], spec, Y[g])
        assert abs(float(elbo[g]) - f_ref) <= TOL * abs(f_ref)
        assert abs(float(stats[g, 0, 0]) - f_ref) <= TOL * abs(f_ref)          # the LL the fit reports is this ELBO
        assert np.abs(grad[g].cpu().numpy() - g_ref).max() <= TOL * max(np.abs(g_ref).max(), 1.0)
