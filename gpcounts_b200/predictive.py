"""Posterior-predictive sampling behind ``load_predict_models`` and the ``safe_mode`` checks of the reference
(``samples_posterior_predictive_distribution`` / ``generate_Samples_from_distribution``, GPcounts_Module.py:808-857),
batched over gene-models.

The latent mean and covariance at the test inputs come from the CUDA library (``gpc_predict``); what happens here is the
reference's Monte-Carlo recipe on top of them - 20 rounds of 5 latent samples, 500 count draws per test point per round
from the likelihood at the running average of exp(f), Savitzky-Golay smoothing of the per-point average - with torch's
generators on the device instead of NumPy / scipy.stats on the host.  It is stochastic by construction: parity with the
reference is statistical, not numerical.  Not part of the fit hot path.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.signal import savgol_filter

ROUNDS, F_PER_ROUND, Y_PER_POINT = 20, 5, 500        # reference :837-842 (20 x predict_f_samples(xtest, 5)), :812-836 (size=500)


def _count_samples(lik_name, mean, alpha, km, gen):
    """(G, Y_PER_POINT, T) count draws at per-point means ``mean`` (G, T); reference :808-834."""
    G, T = mean.shape
    m = mean[:, None, :].expand(G, Y_PER_POINT, T)
    if lik_name == "Poisson":
        return torch.poisson(m, generator=gen)
    r = (1.0 / alpha)[:, None, None].expand(G, Y_PER_POINT, T)           # number of failures, :821, :826
    lam = _gamma(r.contiguous(), gen) * (m / r)
    y = torch.poisson(lam, generator=gen)                                 # NB(r, p = r / (mean + r)) as a Gamma-Poisson mixture
    if lik_name == "Negative_binomial":
        zero_alpha = (alpha == 0)[:, None, None]
        if bool(zero_alpha.any()):
            y = torch.where(zero_alpha, torch.poisson(m, generator=gen), y)      # :816-818
        return y
    # Zero_inflated_negative_binomial: ONE Bernoulli(1 - psi) draw per test point decides between zeros and NB draws (:828-834)
    psi = 1.0 - mean / (km[:, None] + mean)
    keep = torch.rand(G, T, generator=gen, device=mean.device, dtype=mean.dtype) < (1.0 - psi)
    return y * keep[:, None, :].to(y.dtype)


def _gamma(shape, gen):
    """Gamma(shape, 1) draws with an explicit generator (Marsaglia-Tsang, vectorised; torch's sampler takes no generator)."""
    a = torch.where(shape < 1.0, shape + 1.0, shape)
    d = a - 1.0 / 3.0
    c = 1.0 / torch.sqrt(9.0 * d)
    out = torch.zeros_like(a)
    todo = torch.ones_like(a, dtype=torch.bool)
    for _ in range(64):
        if not bool(todo.any()):
            break
        x = torch.randn(a.shape, generator=gen, device=a.device, dtype=a.dtype)
        v = (1.0 + c * x) ** 3
        u = torch.rand(a.shape, generator=gen, device=a.device, dtype=a.dtype)
        ok = (v > 0) & (torch.log(u) < 0.5 * x * x + d - d * v + d * torch.log(torch.where(v > 0, v, torch.ones_like(v))))
        take = todo & ok
        out = torch.where(take, d * v, out)
        todo = todo & ~ok
    boost = torch.where(shape < 1.0, torch.rand(a.shape, generator=gen, device=a.device, dtype=a.dtype) ** (1.0 / shape),
                        torch.ones_like(a))
    return out * boost


def posterior_predictive(lik_name, mean_f, cov_f, alpha, km, branching=False, keep_samples=False, seed=0):
    """Reference :837-857 for G gene-models at once.

    mean_f (G, T), cov_f (G, T, T): latent posterior at the test inputs; alpha, km (G,).
    Returns (mean (G, T) NumPy, samples (G, ROUNDS * Y_PER_POINT, T) NumPy or None)."""
    dev = mean_f.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    G, T = mean_f.shape
    eye = torch.eye(T, dtype=mean_f.dtype, device=dev)
    Lc = torch.linalg.cholesky(cov_f + 1e-6 * eye)                       # gpflow sample_mvn adds default_jitter()
    f_sum = torch.zeros(G, T, dtype=mean_f.dtype, device=dev)
    y_sum = torch.zeros(G, T, dtype=mean_f.dtype, device=dev)
    blocks = []
    n_f = 0
    for _ in range(ROUNDS):
        eps = torch.randn(G, F_PER_ROUND, T, generator=gen, device=dev, dtype=mean_f.dtype)
        f = mean_f[:, None, :] + eps @ Lc.transpose(1, 2)
        f_sum += torch.exp(f).sum(dim=1)
        n_f += F_PER_ROUND
        link_mean = f_sum / n_f                                          # np.mean(link_f, 0) over all samples so far (:843-845)
        y = _count_samples(lik_name, link_mean, alpha, km, gen)
        y_sum += y.sum(dim=1)
        if keep_samples:
            blocks.append(y.cpu())
    if branching:
        mean = (f_sum / n_f).cpu().numpy()                               # :848-849
    else:
        avg = (y_sum / (ROUNDS * Y_PER_POINT)).cpu().numpy()
        mean = savgol_filter(avg, int(T / 2) + 1, 3, axis=1)             # :851-852
        mean = np.where(mean > 0, mean, 0.0)
    samples = torch.cat(blocks, dim=1).numpy() if keep_samples else None
    return mean, samples
