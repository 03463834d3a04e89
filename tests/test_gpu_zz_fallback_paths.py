"""Paths of the fused evaluator that the main parity files do not reach: the count / scale tiles read with ordinary loads
when the gene-major matrices cannot be described by a TMA tensor map (odd row pitch: `make_row_tile_map`, gpc_api.cu), for
both clones of the evaluator (M > 56: compile-time row-block count; M <= 56: run-time).  Sorted last on purpose."""
import numpy as np
import pytest

from tests.helpers import random_theta, synth_counts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [20, 64])
@pytest.mark.parametrize("kernel,use_scale", [("RBF", True), ("RBF", False), ("CONST", False)])
def test_fused_evaluator_without_tma_descriptors(M, kernel, use_scale):
    from gpcounts_b200 import engine
    from oracle import gp_math as gm
    N, D = 151, 2                                   # odd row pitch: no tensor map, y / scale come through plain loads
    rng = np.random.default_rng(M)
    X, Y = synth_counts(D, N, seed=31)
    Z = X[np.sort(rng.choice(N, M, replace=False))].copy()
    scale = rng.uniform(0.5, 2.0, (D, N)) if use_scale else None
    spec = gm.ModelSpec(kind="SVGP", kernel=kernel, lik="NB", X=X, Z=Z)
    theta = random_theta(spec, rng, D)
    model = engine.Model("SVGP", kernel, "NB", X, Z=Z)
    elbo, grad, status = engine.elbo_grad(model, Y, scale, theta)
    elbo, grad, status = elbo.cpu().numpy(), grad.cpu().numpy(), status.cpu().numpy()
    assert (status == 0).all()
    for g in range(D):
        f_ref, g_ref = gm.elbo_and_grad(theta[g], spec, Y[g], None if scale is None else scale[g])
        assert abs(elbo[g] - f_ref) <= 1e-8 * abs(f_ref), (g, elbo[g], f_ref)
        for b in (slice(0, 4), slice(4, 4 + M), slice(4 + M, spec.P)):
            err = np.abs(grad[g][b] - g_ref[b]).max() / max(np.abs(g_ref[b]).max(), 1e-300)
            assert err <= 1e-8, (g, b, err)
