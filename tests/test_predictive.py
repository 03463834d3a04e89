"""CPU: the Monte-Carlo posterior-predictive recipe (gpcounts_b200/predictive.py, reference GPcounts_Module.py:808-857) on
torch CPU tensors - its averages must agree with the analytic expectation E[exp(f)] = exp(mu + s^2 / 2) of the same
latent Gaussian within Monte-Carlo error, for every count likelihood."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("lik", ["Negative_binomial", "Poisson", "Zero_inflated_negative_binomial"])
def test_posterior_predictive_average_matches_expectation(lik):
    from gpcounts_b200 import predictive
    T, G = 40, 3
    x = np.linspace(0, 1, T)
    mu = torch.as_tensor(np.stack([1.0 + np.sin(3 * x), 2.0 + 0 * x, 0.5 + x]))
    K = 0.05 * np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / 0.1 ** 2) + 1e-8 * np.eye(T)
    cov = torch.as_tensor(np.stack([K, 0.5 * K, 2 * K]))
    alpha = torch.tensor([0.3, 1.0, 0.05], dtype=torch.float64)
    km = torch.tensor([1e-9, 1e-9, 1e-9], dtype=torch.float64)         # psi = km / (km + mean) ~ 0: no dropout, ZINB reduces to NB
    mean, samples = predictive.posterior_predictive(lik, mu, cov, alpha, km, keep_samples=True, seed=1)
    assert mean.shape == (G, T) and samples.shape == (G, predictive.ROUNDS * predictive.Y_PER_POINT, T)
    expect = np.exp(mu.numpy() + 0.5 * np.stack([np.diag(K), 0.5 * np.diag(K), 2 * np.diag(K)]))
    assert (mean >= 0).all()
    rel = np.abs(mean - expect) / expect
    assert np.median(rel) < 0.03 and rel.max() < 0.15, (np.median(rel), rel.max())
    # branching variant: plain average of exp(f) over the latent samples (:848-849)
    mean_b, _ = predictive.posterior_predictive(lik, mu, cov, alpha, km, branching=True, seed=2)
    assert np.abs(mean_b - expect).max() / expect.max() < 0.1
    # count samples are integers with (about) the NB / Poisson variance
    assert np.all(samples == np.round(samples))
    if lik == "Poisson":
        assert abs(samples[1].var(axis=0).mean() / samples[1].mean(axis=0).mean() - 1.0) < 0.25


def test_gamma_sampler_moments():
    from gpcounts_b200 import predictive
    gen = torch.Generator().manual_seed(0)
    for shape in (0.3, 1.0, 4.5, 50.0):
        g = predictive._gamma(torch.full((200000,), shape, dtype=torch.float64), gen)
        assert abs(float(g.mean()) / shape - 1.0) < 0.02 and abs(float(g.var()) / shape - 1.0) < 0.05
