"""Round-off sensitivity of the oracle itself: dynamic-model fits of a fixture's genes from 32 starts perturbed by
k * 1e-13 (``oracle.fit.PERTURB``), stored next to the fixture as ``<name>_runs.npz`` (ll_runs[g, k], nit_runs[g, k]).

    python tests/golden/make_perturbed_runs.py one_sample_sparse_nb

SciPy's L-BFGS-B stops the first time one iteration gains less than 2.2e-9 |f|; on a creeping valley round-off decides
when that happens, so a fit has a small set of possible outcomes rather than one.  The GPU parity tests accept a fitted
log-likelihood that coincides with any of these outcomes."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from joblib import Parallel, delayed

name = sys.argv[1] if len(sys.argv) > 1 else "one_sample_sparse_nb"
fx = np.load(os.path.join(HERE, name + ".npz"), allow_pickle=True)
X, Y, M, lik = fx["X"], fx["Y"], int(fx["M"]), str(fx["lik"])
R = 32


def one(g, k):
    import torch
    torch.set_num_threads(1)
    from oracle import fit as ofit
    ofit.PERTURB = k * 1e-13
    o = ofit.OracleFit(X=X, Y=Y, sparse=True, M=M, inducing_variance=1.0)
    rows = o.infer_trajectory(lik, genes=[g]) if hasattr(o, "infer_trajectory") else o.one_sample_test(lik, genes=[g])
    return g, k, rows[0]["ll"][0], rows[0]["fits"][0].nit


res = Parallel(n_jobs=8)(delayed(one)(g, k) for g in range(Y.shape[0]) for k in range(R))
ll = np.zeros((Y.shape[0], R)); nit = np.zeros((Y.shape[0], R), dtype=int)
for g, k, v, n in res:
    ll[g, k], nit[g, k] = v, n
np.savez_compressed(os.path.join(HERE, name + "_runs.npz"), ll_runs=ll, nit_runs=nit, perturb_step=1e-13)
for g in range(Y.shape[0]):
    print("gene %d: spread %.3g, outcomes %s" % (g, ll[g].max() - ll[g].min(), np.array2string(np.sort(ll[g])[[0, 1, 2, R // 2, R - 1]], precision=5)))
