"""Oracle: restatement of RobustGP ``ConditionalVariance.compute_initialisation`` (test infrastructure).

RobustGP (github.com/markvdw/RobustGP, git master, un-vendored and unpinned - ``README.md:19-23`` of the
reference) is absent from ``/root/reference``; GPcounts calls it at ``GPcounts_Module.py:655-659, 673-677``
with the defaults ``sample=False, threshold=0.0``.  Published algorithm (Burt et al. 2020, greedy
conditional-variance / pivoted-Cholesky selection):

    perm = np.random.permutation(N)          # global NumPy RNG - GPcounts seeds it per fit (:758-761, 772)
    d    = Kdiag(X[perm]) + 1e-12
    idx0 = argmax d
    for m in 0..M-2:  j = idx[m];  L = round(K(X, x_j), 20);  L[j] += 1e-12
                      e = (L - c[:m, j] @ c[:m]) / sqrt(d[j]);  c[m] = e;  d = clip(d - e^2, 0)
                      idx[m+1] = argmax d
    Z = X[perm][idx]

Pinned by ``Z[0] = 0.71943149`` on ``data/alpha_time_points.csv`` with seed 0
(``demo_notebooks/scRNA-Seq_time_series.ipynb:670``), see ``tests/test_oracle_golden.py``.
"""
import warnings

import numpy as np


def conditional_variance(X, M, kernel_fn, kdiag_fn):
    """kernel_fn(A, B) -> (len(A), len(B)) Gram; kdiag_fn(A) -> diag.  Uses the *global* NumPy RNG."""
    N = X.shape[0]
    perm = np.random.permutation(N)
    Xp = X[perm]
    indices = np.zeros(M, dtype=int) + N
    di = np.array(kdiag_fn(Xp), dtype=np.float64) + 1e-12
    indices[0] = np.argmax(di)
    if M == 1:
        return Xp[indices], perm[indices]
    ci = np.zeros((M - 1, N))
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for m in range(M - 1):
            j = int(indices[m])
            dj = np.sqrt(di[j])
            cj = ci[:m, j]
            L = np.round(np.squeeze(np.array(kernel_fn(Xp, Xp[j:j + 1]))), 20)
            L[j] += 1e-12
            ei = (L - np.dot(cj, ci[:m])) / dj
            ci[m, :] = ei
            di = di - ei ** 2
            di = np.clip(di, 0, None)
            indices[m + 1] = np.argmax(di)
    indices = indices.astype(int)
    return Xp[indices], perm[indices]


def rbf_np(A, B, var, ls):
    A = A / ls
    B = B / ls
    d2 = -2.0 * A @ B.T + (A * A).sum(-1)[:, None] + (B * B).sum(-1)[None, :]
    return var * np.exp(-0.5 * d2)


def select_Z(X, M, kernel, var, ls):
    """Z for a gpflow RBF / Constant kernel (what GPcounts passes at :657, :675)."""
    if kernel == "CONST":
        kf = lambda A, B: var * np.ones((A.shape[0], B.shape[0]))
    else:
        kf = lambda A, B: rbf_np(A, B, var, ls)
    kd = lambda A: var * np.ones(A.shape[0])
    return conditional_variance(X, M, kf, kd)[0]
