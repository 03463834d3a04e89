"""Shared test helpers: synthetic inputs (SURVEY.md section 8d generator) and random parameter vectors."""
import numpy as np


def synth_counts(D, N, seed=1234, x_seed=0):
    """C5 generator: sorted uniform pseudotime, NB counts from sinusoidal / constant log-means."""
    X = np.sort(np.random.default_rng(x_seed).uniform(0, 1, N))[:, None]
    rng = np.random.default_rng(seed)
    Y = np.zeros((D, N))
    for g in range(D):
        b = rng.uniform(np.log(0.5), np.log(50))
        a = np.exp(rng.uniform(np.log(0.02), np.log(1.5))) if g % 2 == 0 else 0.0
        om = rng.uniform(0.5, 2)
        ph = rng.uniform(0, 2 * np.pi)
        alpha = np.exp(rng.uniform(np.log(0.05), np.log(2)))
        f = b + a * np.sin(2 * np.pi * om * X[:, 0] + ph)
        r = 1.0 / alpha
        Y[g] = rng.negative_binomial(r, r / (r + np.exp(f)))
    return X, Y


def random_theta(spec, rng, D):
    """Random but well-conditioned parameter vectors in the oracle/CUDA layout."""
    P, mv = spec.P, spec.Mv
    th = np.zeros((D, P))
    th[:, 0] = rng.uniform(-1.0, 2.0, D)       # variance
    th[:, 1] = rng.uniform(-2.5, 0.5, D)       # lengthscale
    th[:, 2] = rng.uniform(-2.0, 1.5, D)       # alpha / noise
    th[:, 3] = rng.uniform(0.0, 4.0, D)        # km
    if mv:
        th[:, 4:4 + mv] = 0.5 * rng.standard_normal((D, mv))
        tri = 0.1 * rng.standard_normal((D, mv * (mv + 1) // 2))
        diag = np.array([i * (i + 1) // 2 + i for i in range(mv)])
        tri[:, diag] = rng.uniform(0.3, 1.2, (D, mv))
        th[:, 4 + mv:] = tri
    return th
