"""Throughput of the dense path: full VGP One_sample_test (MouseOB-like: N spots in 2-D, scaled NB)."""
import sys, time, numpy as np, pandas as pd, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import Fit_GPcounts
N = int(sys.argv[1]) if len(sys.argv) > 1 else 260
D = int(sys.argv[2]) if len(sys.argv) > 2 else 296
rng = np.random.default_rng(0)
X = np.c_[rng.uniform(8, 28, N), rng.uniform(9, 30, N)]
f = rng.normal(0, 1, (D, 1)) + 0.5 * np.sin(X[:, 0][None, :] / 3 + rng.uniform(0, 6, (D, 1))) * (np.arange(D)[:, None] % 2)
scale = rng.uniform(0.5, 2.0, (N, D))
Y = rng.negative_binomial(5.0, 5.0 / (5.0 + np.exp(f + 1.5) * scale.T)).astype(float)
cells = ["c%d" % i for i in range(N)]
import os
gp = Fit_GPcounts(pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(D)], columns=cells),
                  scale=pd.DataFrame(scale, index=cells))
if os.environ.get("MAXITER"): gp.optimizer_options = {"maxiter": int(os.environ["MAXITER"])}
t = time.time(); df = gp.One_sample_test("Negative_binomial"); dt = time.time() - t
st = gp.fit_stats
nfev = st.n_feval_total.sum()
flops = nfev * ((7.0 / 3.0) * N ** 3 + 6.0 * N * N)
print("VGP N=%d D=%d: %.2f s wall, kernel %.2f s -> %.1f genes/s, %.0f evals/s, %.2f TFLOP/s algorithmic, ok %.3f, evals/gene %.0f"
      % (N, D, dt, gp.last_fit_seconds, D / gp.last_fit_seconds, nfev / gp.last_fit_seconds, flops / gp.last_fit_seconds * 1e-12,
         (st.status == 0).mean(), nfev / D))
