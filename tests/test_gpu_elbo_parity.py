"""Fixed-theta parity: ELBO and gradient of the CUDA path vs the CPU oracle, tolerance 1e-8 relative
(BASELINE.json north_star), for every model x kernel x likelihood variant of the hot path."""
import numpy as np
import pytest
import torch

from tests.helpers import random_theta, synth_counts

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _compare(kind, kernel, lik, N, M=0, d=1, use_scale=False, D=3, seed=0, train=(True, True, True, True), xp=None):
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    rng = np.random.default_rng(seed)
    X, Y = synth_counts(D, N, seed=seed + 10)
    if d == 2:
        X = np.c_[X, rng.uniform(0, 1, N)]
    if kernel == "BRANCH":
        X = np.c_[X[:, 0], rng.integers(1, 3, N).astype(float)]
        xpv = 0.4 if xp is None else xp
        X[X[:, 0] <= xpv, 1] = 1.0
    else:
        xpv = -1000.0
    Z = None
    if kind in ("SVGP", "SGPR"):
        Z = X[np.sort(rng.choice(N, M, replace=False))].copy()
    if lik == "GAUSS":
        Y = np.log(Y + 1)
    scale = rng.uniform(0.5, 2.0, (D, N)) if use_scale else None
    spec = gm.ModelSpec(kind=kind, kernel=kernel, lik=lik, X=X, Z=Z, xp=xpv, train=train)
    theta = random_theta(spec, rng, D)
    model = engine.Model(kind, kernel, lik, X, Z=Z, train=train, xp=xpv)
    elbo, grad, status = engine.elbo_grad(model, Y, scale, theta)
    elbo, grad, status = elbo.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy()
    assert (status == 0).all()
    for g in range(D):
        f_ref, g_ref = gm.elbo_and_grad(theta[g], spec, Y[g], None if scale is None else scale[g])
        assert abs(elbo[g] - f_ref) <= TOL * abs(f_ref), (g, elbo[g], f_ref)
        # gradient blocks are compared on the scale of the largest entry of each block
        mv = spec.Mv
        blocks = [slice(0, 4), slice(4, 4 + mv), slice(4 + mv, spec.P)] if mv else [slice(0, 4)]
        for b in blocks:
            ref = g_ref[b]
            if ref.size == 0:
                continue
            err = np.abs(grad[g][b] - ref).max() / max(np.abs(ref).max(), 1e-300)
            assert err <= TOL, (kind, kernel, lik, g, b, err)


@pytest.mark.parametrize("lik", ["NB", "ZINB", "POISSON"])
@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_svgp(lik, kernel):
    _compare("SVGP", kernel, lik, N=150, M=20)


@pytest.mark.parametrize("M", [5, 13, 24, 33, 41, 50, 57, 64])
def test_svgp_fused_every_row_block_count(M):
    """The fused evaluator splits its rank update between paired warps differently for every number of 8-row blocks
    (1..8); N = 150 also leaves a ragged last column tile."""
    _compare("SVGP", "RBF", "NB", N=150, M=M, D=2, seed=M)
    _compare("SVGP", "CONST", "ZINB", N=150, M=M, D=2, seed=M + 1)


@pytest.mark.parametrize("lik", ["NB", "ZINB", "POISSON"])
@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_vgp(lik, kernel):
    _compare("VGP", kernel, lik, N=40)


def test_vgp_blocked_sizes():
    _compare("VGP", "RBF", "NB", N=130, D=2)        # three 64-blocks, ragged last


def test_svgp_blocked_sizes():
    _compare("SVGP", "RBF", "NB", N=300, M=70, D=2)


def test_svgp_scale_2d():
    _compare("SVGP", "RBF", "NB", N=120, M=16, d=2, use_scale=True)


def test_vgp_scale_2d():
    _compare("VGP", "RBF", "NB", N=60, d=2, use_scale=True)


@pytest.mark.parametrize("xp", [-1000.0, 0.4])
def test_vgp_branch(xp):
    _compare("VGP", "BRANCH", "NB", N=50, xp=xp)


def test_vgp_branch_frozen_kernel():
    _compare("VGP", "BRANCH", "NB", N=50, xp=0.4, train=(False, False, True, True))


@pytest.mark.parametrize("kernel", ["RBF", "CONST"])
def test_gpr(kernel):
    _compare("GPR", kernel, "GAUSS", N=70)


@pytest.mark.parametrize("kernel,M", [("RBF", 12), ("CONST", 12), ("RBF", 70)])
def test_sgpr(kernel, M):
    _compare("SGPR", kernel, "GAUSS", N=150, M=M)


def test_gpr_branch_kernel():
    """Gaussian likelihood with the branching kernel (Branching_GPcounts notebook, lik_name="Gaussian")."""
    _compare("GPR", "BRANCH", "GAUSS", N=60, xp=0.4)


def test_per_gene_inducing_points():
    """gpc_model_t.Zgene: every gene evaluates with its own inducing set."""
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    rng = np.random.default_rng(4)
    D, N, M = 3, 90, 12
    X, Y = synth_counts(D, N, seed=21)
    Zg = np.stack([X[np.sort(rng.choice(N, M, replace=False))] for _ in range(D)])
    model = engine.Model("SVGP", "RBF", "NB", X, Z=Zg[0], Zgene=Zg)
    spec0 = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Zg[0])
    theta = random_theta(spec0, rng, D)
    elbo, grad, status = engine.elbo_grad(model, Y, None, theta)
    for g in range(D):
        spec = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Zg[g])
        f_ref, g_ref = gm.elbo_and_grad(theta[g], spec, Y[g])
        assert abs(float(elbo[g]) - f_ref) <= TOL * abs(f_ref)
        assert np.abs(grad[g].cpu().numpy() - g_ref).max() <= TOL * np.abs(g_ref).max()


def test_not_positive_definite_is_reported():
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    X, Y = synth_counts(2, 30)
    spec = gm.ModelSpec(kind="VGP", kernel="RBF", lik="NB", X=X)
    theta = random_theta(spec, np.random.default_rng(0), 2)
    theta[1, 0] = float("nan")
    model = engine.Model("VGP", "RBF", "NB", X)
    elbo, grad, status = engine.elbo_grad(model, Y, None, theta)
    assert status.tolist() == [0, 1]
    assert torch.isnan(elbo[1]) and torch.isfinite(elbo[0])


@pytest.mark.parametrize("kind,N,M", [("SVGP", 150, 20), ("VGP", 30, 0), ("VGP", 80, 0)])
@pytest.mark.parametrize("case", ["overflow", "underflow", "underflow_positive_counts", "alpha_to_zero"])
def test_nb_objective_at_wild_trial_points_has_the_references_ieee_class(kind, N, M, case):
    """Line searches of the reference visit trial points where exp(F) over- or underflows, or alpha -> 0.  There the
    reference's NB log-density Y log(m / (m + k)) - k log(1 + m alpha) is NaN or -inf by IEEE rules (TensorFlow / the
    oracle evaluate it literally) and lgamma(k + y) - lgamma(k) is exactly 0 once k + y == k; SciPy's line search
    reacts differently to NaN, inf and finite objectives, so the device must return the same class (fused on-chip,
    small-N on-chip and dense evaluators):
      overflow   q_mu pushes F above 709.8 at some nodes: m = inf -> NaN
      underflow  F below -745.2: m = 0 -> NaN where a count is 0 (0 * -inf), -inf when every count is positive
      alpha -> 0 k = 1 / alpha ~ 1e39: finite objective (two different lgamma implementations would differ by ~1e25)."""
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    rng = np.random.default_rng(5)
    X, Y = synth_counts(2, N, seed=21)
    if case == "underflow_positive_counts":
        Y = Y + 1.0
    Z = X[np.sort(rng.choice(N, M, replace=False))].copy() if kind == "SVGP" else None
    spec = gm.ModelSpec(kind=kind, kernel="RBF", lik="NB", X=X, Z=Z)
    theta = random_theta(spec, rng, 2)
    mv = spec.Mv
    if case == "overflow":
        theta[:, 4:4 + mv] = 900.0
    elif case.startswith("underflow"):
        theta[:, 4:4 + mv] = -900.0
    else:
        theta[:, 2] = -91.0
    model = engine.Model(kind, "RBF", "NB", X, Z=Z)
    elbo, grad, status = engine.elbo_grad(model, Y, None, theta)
    elbo, status = elbo.cpu().numpy(), status.cpu().numpy()
    for g in range(2):
        f_ref, _ = gm.elbo_and_grad(theta[g], spec, Y[g], None)
        cls = lambda v: "nan" if np.isnan(v) else ("-inf" if v == -np.inf else ("+inf" if v == np.inf else "finite"))
        assert status[g] == 0
        assert cls(elbo[g]) == cls(f_ref), (case, g, elbo[g], f_ref)
        if np.isfinite(f_ref):
            assert abs(elbo[g] - f_ref) <= 1e-6 * abs(f_ref), (case, g, elbo[g], f_ref)
