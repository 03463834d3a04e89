"""Size-independent properties of the CUDA path at the benchmark's full shape (N = 1000 spots, M = 64 inducing points),
where the CPU oracle is too slow to be run on many genes."""
import numpy as np
import pytest
import torch

from tests.helpers import random_theta

pytestmark = pytest.mark.gpu


def _c5(D, seed=7):
    from bench import synth_batch
    from gpcounts_b200 import host
    X, Y = synth_batch(D, 1000, seed=seed)
    ls0 = host.default_lengthscale(X)
    Z = host.inducing_sets(X, 64, ls0)
    rt, _ = host.restart_table(X)
    h = np.zeros((D, 2, 4))
    h[:, :, 0] = host.default_variance(Y, "Negative_binomial", True)[:, None]
    h[:, :, 1], h[:, :, 2], h[:, :, 3] = ls0, 1.0, 35.0
    return X, Y, Z, rt, h


def test_fit_is_deterministic_and_order_independent():
    """Genes never interact: the result of a gene is bit-identical whatever the batch order, the number of CTA slots
    or the previous contents of the workspace (this is also what makes multi-GPU sharding exact)."""
    from gpcounts_b200 import engine
    D = 40
    X, Y, Z, rt, h = _c5(D)
    models = [engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=rt),
              engine.Model("SVGP", "CONST", "NB", X, Z=Z, restart=rt, handoff_alpha=True)]
    opt = engine.default_opt(maxiter=60)
    ref, _ = engine.fit(models, Y, None, h, chain=True, opt=opt, n_slots=148)
    perm = np.random.default_rng(0).permutation(D)
    for buf in engine._ws_cache.values():
        if buf is not None:
            buf.view(torch.float64)[: buf.numel() // 8].fill_(float("nan"))       # poison the workspace
    out, _ = engine.fit(models, Y[perm], None, h[perm], chain=True, opt=opt, n_slots=7)
    a, b = ref.cpu().numpy()[perm][:, :, :12], out.cpu().numpy()[:, :, :12]
    assert np.array_equal(a, b, equal_nan=True)


def test_gradient_matches_finite_differences_at_full_size():
    """d ELBO / d theta from the analytic reverse pass vs central differences of the ELBO, N = 1000, M = 64."""
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    D = 3
    X, Y, Z, rt, h = _c5(D, seed=11)
    model = engine.Model("SVGP", "RBF", "NB", X, Z=Z[0])
    spec = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Z[0])
    rng = np.random.default_rng(2)
    theta = random_theta(spec, rng, D)
    _, grad, st = engine.elbo_grad(model, Y, None, theta)
    assert int(st.abs().sum()) == 0
    grad = grad.cpu().numpy()
    # the ELBO carries ~1e-10 relative evaluation noise (cond(Kuu) ~ 1e8), so the step must not be too small
    eps = 3e-5
    for trial in range(6):
        v = rng.standard_normal(theta.shape)
        if trial < 3:                       # also probe the hyper-parameter directions alone
            v[:, 4:] = 0.0
        fp, _, _ = engine.elbo_grad(model, Y, None, theta + eps * v)
        fm, _, _ = engine.elbo_grad(model, Y, None, theta - eps * v)
        fd = (fp - fm).cpu().numpy() / (2 * eps)
        an = (grad * v).sum(1)
        assert np.all(np.abs(fd - an) <= 2e-4 * np.maximum(np.abs(an), 1.0)), (trial, fd, an)


def test_sparse_with_all_points_equals_full_vgp():
    """SVGP with Z = X is the VGP up to the placement of the jitter: the two CUDA evaluators
    (fused on-chip vs dense) must agree to ~1e-6 relative."""
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    from tests.helpers import synth_counts
    X, Y = synth_counts(4, 48, seed=3)
    vg = gm.ModelSpec(kind="VGP", kernel="RBF", lik="NB", X=X)
    th = random_theta(vg, np.random.default_rng(5), 4)
    th[:, 1] = 0.5                      # a lengthscale that keeps Kuu well conditioned
    e1, _, _ = engine.elbo_grad(engine.Model("VGP", "RBF", "NB", X), Y, None, th)
    e2, _, _ = engine.elbo_grad(engine.Model("SVGP", "RBF", "NB", X, Z=X.copy()), Y, None, th)
    assert torch.allclose(e1, e2, rtol=1e-4)


def test_elbo_grad_is_reproducible_bitwise():
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    X, Y, Z, rt, h = _c5(8, seed=13)
    model = engine.Model("SVGP", "RBF", "NB", X, Z=Z[0])
    spec = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Z[0])
    theta = random_theta(spec, np.random.default_rng(1), 8)
    e1, g1, _ = engine.elbo_grad(model, Y, None, theta, n_slots=8)
    e2, g2, _ = engine.elbo_grad(model, Y[::-1].copy(), None, theta[::-1].copy(), n_slots=3)
    assert torch.equal(e1, e2.flip(0)) and torch.equal(g1, g2.flip(0))


def test_dense_path_is_reproducible_under_contention():
    """Two resident CTAs per SM skew the warps of a CTA against each other; a missing barrier in the shared-memory
    factorisations shows up as run-to-run differences (or a hang).  Dense VGP (two diagonal blocks) and dense SVGP
    (M > 64), every slot busy vs few slots: bit-identical."""
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    from tests.helpers import synth_counts
    D = 600
    X, Y = synth_counts(D, 70, seed=8)
    rng = np.random.default_rng(4)
    for kind, Z in (("VGP", None), ("SVGP", np.sort(rng.uniform(X.min(), X.max(), (72, 1)), axis=0))):
        spec = gm.ModelSpec(kind=kind, kernel="RBF", lik="NB", X=X, Z=Z)
        theta = random_theta(spec, rng, D)
        model = engine.Model(kind, "RBF", "NB", X, Z=Z)
        e1, g1, s1 = engine.elbo_grad(model, Y, None, theta, n_slots=296)
        e2, g2, s2 = engine.elbo_grad(model, Y, None, theta, n_slots=37)
        assert torch.equal(s1, s2)
        bits = lambda t: t.contiguous().view(torch.int64)      # NaN-safe bitwise comparison
        assert torch.equal(bits(e1), bits(e2)) and torch.equal(bits(g1), bits(g2))
