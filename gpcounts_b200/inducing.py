"""Per-gene inducing-point selection, batched over genes on the GPU.

The reference recomputes RobustGP's greedy conditional-variance selection for every gene-fit under that fit's
initial kernel (GPcounts_Module.py:673-677).  In exact arithmetic the choice does not depend on the kernel
variance, but while the inducing set is still sparse many candidates tie to the last bit and the winner is
decided by how ``(var + 1e-12) - e^2`` rounds - which depends on the gene's initial variance.  So attempt 0 of
a sparse fit has *gene-specific* inducing points.  This module runs the same recurrence for all genes at once
with elementwise FP64 torch ops (IEEE mul / add / sqrt / div, accumulated in the index order of the
reference's dot product), so the selected indices are reproducible and agree with the host recurrence in
``host.greedy_conditional_variance``.
"""
from __future__ import annotations

import numpy as np
import torch

from . import host


def _first_argmax(d):
    """Index of the first maximum along dim 1 (np.argmax tie rule)."""
    mx = d.max(dim=1, keepdim=True).values
    n = d.shape[1]
    ar = torch.arange(n, device=d.device).expand_as(d)
    return torch.where(d == mx, ar, torch.full_like(ar, n)).min(dim=1).values


def select_per_gene(X, M, ls, var, device, chunk=2048):
    """Z (D, M, d) for genes with initial variances ``var`` (D,), seed-0 permutation (reference :758-761)."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    saved = np.random.get_state()
    np.random.seed(0)
    perm = np.random.permutation(N)
    np.random.set_state(saved)
    Xp = X[perm]
    K1 = torch.as_tensor(host._rbf(Xp, Xp, ls), device=device)            # unit-variance Gram, host exp
    var = torch.as_tensor(np.asarray(var, dtype=np.float64), device=device)
    D = var.shape[0]
    out = torch.empty(D, M, dtype=torch.long, device=device)
    for c0 in range(0, D, chunk):
        v = var[c0:c0 + chunk]
        G = v.shape[0]
        ar = torch.arange(G, device=device)
        d = v[:, None] + torch.zeros(G, N, dtype=torch.float64, device=device) + 1e-12
        idx = torch.empty(G, M, dtype=torch.long, device=device)
        idx[:, 0] = _first_argmax(d)
        Cm = torch.zeros(G, max(M - 1, 1), N, dtype=torch.float64, device=device)
        for m in range(M - 1):
            j = idx[:, m]
            L = v[:, None] * K1[j]
            L = torch.round(L * 1e20) / 1e20                              # np.round(L, 20)
            L[ar, j] += 1e-12
            acc = torch.zeros(G, N, dtype=torch.float64, device=device)
            for k in range(m):
                acc = acc + Cm[ar, k, j][:, None] * Cm[:, k, :]
            e = (L - acc) / torch.sqrt(d[ar, j])[:, None]
            Cm[:, m, :] = e
            d = torch.clamp(d - e * e, min=0.0)
            idx[:, m + 1] = _first_argmax(d)
        out[c0:c0 + G] = idx
    Xp_t = torch.as_tensor(Xp, device=device)
    return Xp_t[out]                                                     # (D, M, d)
