"""Evaluation trace of one (gene, model) fit of a golden fixture (debug fields of gpc_opt_t: trace, trace_cap, trace_gene, trace_model):

    python tools/trace_gene.py one_sample_sparse_nb 0 0

Prints the fit statistics and the last evaluations (f, step, g.d, max|g|)."""
import sys, numpy as np, pandas as pd, torch, ctypes as C
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import Fit_GPcounts, _lib
name, gene, model = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
fx = np.load('/root/repo/tests/golden/%s.npz' % name, allow_pickle=True)
X, Y = fx["X"], fx["Y"]
cells = ["c%d" % i for i in range(X.shape[0])]
gp = Fit_GPcounts(pd.DataFrame(X, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(Y.shape[0])], columns=cells),
                  sparse="M" in fx.files, M=int(fx["M"]) if "M" in fx.files else 0)
lib = _lib.load()
cap = 4000
buf = torch.zeros(cap * 8, dtype=torch.float64, device="cuda")
gp.optimizer_options = dict(trace=buf.data_ptr(), trace_cap=cap, trace_gene=gene, trace_model=model)
df = gp.One_sample_test(str(fx["lik"]))
torch.cuda.synchronize()
print(df.iloc[gene].values, "oracle", fx["ll"][gene])
st = gp.fit_stats
print(st[(st.gene == st.gene.unique()[gene])].to_string())
t = buf.cpu().numpy().reshape(cap, 8)
n = int((t[:, 0] != 0).sum())
print("evaluations traced:", n)
for r in t[max(0, n - 12):n]:
    print("f %.10f stp %.4g gd %.4g gmax %.4g  hyp %.6f %.6f %.6f" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6]))
