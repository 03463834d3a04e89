"""Is a gene's fit independent of the batch it is fitted in, and reproducible run to run?  (no cross-gene arithmetic exists)"""
import sys, numpy as np, pandas as pd
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import Fit_GPcounts
from tests.helpers import synth_counts
D = int(sys.argv[1]) if len(sys.argv) > 1 else 101
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
X, Y = synth_counts(D, N, seed=23)
cells = ["c%d" % i for i in range(N)]
Xdf = pd.DataFrame(X, index=cells, columns=["times"])
Ydf = pd.DataFrame(Y, index=["g%d" % i for i in range(D)], columns=cells)
def run(Yd, test):
    gp = Fit_GPcounts(Xdf, Yd)
    return getattr(gp, test)("Negative_binomial").values, gp.fit_stats
for test in ("Two_samples_test", "One_sample_test"):
    a, sa = run(Ydf, test)
    b, sb = run(Ydf, test)
    print(test, "batch vs batch: rows differing", int((a != b).any(axis=1).sum()), "of", D)
    bad = []
    for g in range(min(D, 12)):
        c, sc = run(Ydf.iloc[[g]], test)
        c2, _ = run(Ydf.iloc[[g]], test)
        if not np.array_equal(c[0], a[g]) or not np.array_equal(c[0], c2[0]):
            bad.append(g)
            print("  gene", g, "batch", a[g], "single", c[0], "single again", c2[0])
            print(sa[sa.gene == "g%d" % g][["model", "n_iter", "n_feval", "restarts"]].values.tolist(), sc[["n_iter", "n_feval"]].values.tolist())
    print(test, "single vs batch mismatches among first 12 genes:", bad)
