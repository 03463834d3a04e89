"""Host-side pieces of the fit path that are gene-independent or O(D): hyper-parameter initialisation and
restart tables, inducing-point selection, p/q-values.  Pure NumPy; nothing here is on the per-gene hot path
(that is the CUDA library) and nothing here imports the test oracle.
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.stats as ss
from scipy import interpolate

N_RESTART = 10


def default_lengthscale(X):
    """5 % of the joint range of all entries of X (reference: GPcounts_Module.py:730)."""
    return (5 * (np.max(X) - np.min(X))) / 100


def default_variance(Y, lik_name, transform):
    """Per-gene initial kernel variance (reference: GPcounts_Module.py:734-738), Y is (D, N)."""
    # The reference reduces one gene at a time over a contiguous copy of its row (:421-422, :738): NumPy's pairwise
    # summation.  A (D, N) array taken from a DataFrame is Fortran-ordered, and reducing it along axis 1 sums in a
    # different order (1-ulp differences that decide where a creeping L-BFGS-B run stops) - so force row-major here.
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    if lik_name == "Gaussian" and not transform:
        return np.mean(Y + 1 ** 2, axis=1)            # operator-precedence quirk of the reference kept (:736)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.mean(np.log(Y + 1) ** 2, axis=1)


def restart_table(X):
    """The four uniform draws of restart r = 1..10 (reference: GPcounts_Module.py:769-779).

    The reference reseeds the global NumPy RNG with r before drawing, so the table is the same for every
    gene.  Returns (table (10, 4) = (ls, var, alpha, km), list of RNG states right after the draws) - the
    states are needed because inducing-point selection continues from them (:673-677)."""
    rng_range = np.max(X) - np.min(X)
    tab = np.zeros((N_RESTART, 4))
    states = []
    saved = np.random.get_state()
    for r in range(1, N_RESTART + 1):
        np.random.seed(r)
        tab[r - 1, 0] = np.random.uniform((0.25 * rng_range) / 100, (30.0 * rng_range) / 100)
        tab[r - 1, 1] = np.random.uniform(0.0, 10.0)
        tab[r - 1, 2] = np.random.uniform(0.0, 10.0)
        tab[r - 1, 3] = np.random.uniform(0.0, 100.0)
        states.append(np.random.get_state())
    np.random.set_state(saved)
    return tab, states


def _rbf(A, B, ls):
    A = A / ls
    B = B / ls
    d2 = -2.0 * A @ B.T + (A * A).sum(-1)[:, None] + (B * B).sum(-1)[None, :]
    return np.exp(-0.5 * d2)


def greedy_conditional_variance(X, M, ls, var=1.0):
    """Greedy conditional-variance inducing-point selection under an RBF kernel (what RobustGP's
    ``ConditionalVariance().compute_initialisation(X, M, kernel)`` does for GPcounts, :655-659, :673-677).
    Consumes one ``np.random.permutation(N)`` from the global RNG, exactly like the call it replaces."""
    N = X.shape[0]
    perm = np.random.permutation(N)
    Xp = X[perm]
    idx = np.zeros(M, dtype=int)
    d = np.full(N, var, dtype=np.float64) + 1e-12
    idx[0] = int(np.argmax(d))
    C = np.zeros((max(M - 1, 0), N))
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for m in range(M - 1):
            j = idx[m]
            L = np.round(var * _rbf(Xp, Xp[j:j + 1], ls)[:, 0], 20)
            L[j] += 1e-12
            e = (L - C[:m, j] @ C[:m]) / np.sqrt(d[j])
            C[m] = e
            d = np.clip(d - e * e, 0, None)
            idx[m + 1] = int(np.argmax(d))
    return Xp[idx]


def inducing_sets(X, M, ls0, with_restarts=True):
    """Inducing points for attempt 0 (seed 0, ls0) and restarts 1..10 (seed r, ls_r), gene-independent.
    Returns (n_sets, M, d)."""
    saved = np.random.get_state()
    sets = []
    np.random.seed(0)
    sets.append(greedy_conditional_variance(X, M, ls0))
    if with_restarts:
        tab, states = restart_table(X)
        for r in range(N_RESTART):
            np.random.set_state(states[r])
            sets.append(greedy_conditional_variance(X, M, tab[r, 0], tab[r, 1]))
    np.random.set_state(saved)
    return np.stack(sets)


def qvalue(pv, pi0=None):
    """Storey q-values (reference: GPcounts_Module.py:1008-1063, SciPy aliases replaced by NumPy)."""
    pv = np.asarray(pv, dtype=np.float64)
    # the reference receives a pandas Series, whose min() / max() skip the NaN p-values of failed fits (:455-457);
    # NaN rows sort last, keep m = len(pv) and come out as NaN q-values, exactly as there
    if np.isfinite(pv).any():
        assert np.nanmin(pv) >= 0 and np.nanmax(pv) <= 1, "p-values should be between 0 and 1"
    shape = pv.shape
    pv = pv.ravel()
    m = float(len(pv))
    if len(pv) < 100 and pi0 is None:
        pi0 = 1.0
    elif pi0 is None:
        lam = np.arange(0, 0.90, 0.01)
        counts = np.array([(pv > i).sum() for i in lam])
        pi0s = counts / (m * (1 - lam))
        tck = interpolate.splrep(lam, pi0s, k=3)
        pi0 = float(interpolate.splev(lam[-1], tck))
        if pi0 > 1:
            pi0 = 1.0
    assert 0 <= pi0 <= 1, "pi0 is not between 0 and 1: %f" % pi0
    order = np.argsort(pv)
    pv = pv[order]
    qv = pi0 * m / len(pv) * pv
    qv[-1] = min(qv[-1], 1.0)
    for i in range(len(pv) - 2, -1, -1):
        qv[i] = min(pi0 * m * pv[i] / (i + 1.0), qv[i + 1])
    out = np.zeros_like(qv)
    out[order] = qv
    return out.reshape(shape)


def p_values(llr):
    """1 - chi2_1.cdf(2 LLR) (reference: GPcounts_Module.py:369-371)."""
    return 1 - ss.chi2.cdf(2 * np.asarray(llr, dtype=np.float64), df=1)


def branching_evidence(loglik, Bsearch):
    """Posterior over candidate branching times and log Bayes factor (reference: :335-366)."""
    o = np.asarray(loglik, dtype=np.float64)
    pn = np.exp(o - np.max(o))
    p = pn / pn.sum()
    Nb = o.size - 1
    if Nb != len(Bsearch) - 1:
        raise NameError("Passed in wrong length of Bsearch is %g- should be %g" % (len(Bsearch), Nb))
    obj = o[:-1]
    imax = int(np.argmax(obj))
    llmax = obj[imax]
    lr = llmax + np.log(1 + np.exp(obj[np.arange(obj.size) != imax] - llmax).sum()) - o[-1] - np.log(Nb)
    return {"posteriorBranching": p, "logBayesFactor": lr}
