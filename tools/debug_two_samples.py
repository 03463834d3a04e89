import sys, numpy as np, pandas as pd, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import Fit_GPcounts
fx = np.load('/root/repo/tests/golden/two_samples_vgp_nb.npz', allow_pickle=True)
X, Y = fx["X"], fx["Y"]
cells = ["c%d" % i for i in range(X.shape[0])]
for variant in (-1, 1, 0):
    gp = Fit_GPcounts(pd.DataFrame(X, index=cells, columns=["times"]), pd.DataFrame(Y, index=["g%d" % i for i in range(Y.shape[0])], columns=cells))
    gp.optimizer_options = dict(variant=variant)
    df = gp.Two_samples_test("Negative_binomial")
    print("variant", variant)
    print(np.c_[df.values, fx["ll"]])
    print(gp.fit_stats[["gene", "model", "log_likelihood", "status", "n_iter", "n_feval", "restarts", "variance", "lengthscale", "alpha"]].to_string())
