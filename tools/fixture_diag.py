"""Per-gene diagnosis of a sparse One_sample_test fixture against the oracle's outcome set (tests/golden/<name>_runs.npz):
prints the genes whose dynamic fit ends outside the envelope with their fit statistics and the tail of the evaluation trace.

    python tools/fixture_diag.py c5_sample_sparse_nb
"""
import sys, numpy as np, pandas as pd, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import Fit_GPcounts
name = sys.argv[1] if len(sys.argv) > 1 else "c5_sample_sparse_nb"
fx = np.load('/root/repo/tests/golden/%s.npz' % name, allow_pickle=True)
runs = np.load('/root/repo/tests/golden/%s_runs.npz' % name)["ll_runs"]
X, Y = fx["X"], fx["Y"]
cells = ["c%d" % i for i in range(X.shape[0])]
def fit(trace_gene=None):
    gp = Fit_GPcounts(pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(Y.shape[0])], columns=cells),
                      sparse=True, M=int(fx["M"]))
    buf = None
    if trace_gene is not None:
        buf = torch.zeros(4000 * 8, dtype=torch.float64, device="cuda")
        gp.optimizer_options = dict(trace=buf.data_ptr(), trace_cap=4000, trace_gene=trace_gene, trace_model=0)
    df = gp.One_sample_test(str(fx["lik"]))
    torch.cuda.synchronize()
    return gp, df, buf
gp, df, _ = fit()
got = df.values
ref = np.c_[runs, fx["ll"][:, 0], fx["ll_pert"][:, 0]]
lo, hi = ref.min(axis=1) - 1e-4 * np.abs(fx["ll"][:, 0]), ref.max(axis=1) + 1e-4 * np.abs(fx["ll"][:, 0])
bad = np.where((got[:, 0] < lo) | (got[:, 0] > hi))[0]
print("genes outside the oracle's envelope:", bad)
st = gp.fit_stats
for g in bad:
    print("gene", g, "gpu", got[g, 0], "oracle envelope", lo[g], hi[g], "oracle nit", fx["nit"][g], fx["nit_pert"][g])
    print(st[st.gene == "g%d" % g].to_string())
    _, _, buf = fit(trace_gene=int(g))
    t = buf.cpu().numpy().reshape(4000, 8)
    n = int((t[:, 0] != 0).sum())
    print("evaluations traced (attempt 0):", n)
    for r in t[max(0, n - 10):n]:
        print("  f %.10f stp %.4g gd %.4g gmax %.4g  theta %.6f %.6f %.6f" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6]))
