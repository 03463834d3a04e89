"""Round-off sensitivity of one fit on the GPU: the same gene 32 times, initial variance scaled by (1 + k * 1e-13).

    python tools/perturb_gene.py one_sample_sparse_nb 0
"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import engine, host
name, gene = sys.argv[1], int(sys.argv[2])
fx = np.load('/root/repo/tests/golden/%s.npz' % name, allow_pickle=True)
X, Y, M = fx["X"], fx["Y"], int(fx["M"])
D = 32
ls0 = host.default_lengthscale(X)
Z = host.inducing_sets(X, M, ls0); rt, _ = host.restart_table(X)
m = engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=rt)
Yg = np.repeat(Y[gene:gene + 1], D, axis=0)
h = np.zeros((D, 1, 4))
h[:, 0, 0] = host.default_variance(Yg, "Negative_binomial", True) * (1.0 + np.arange(D) * 1e-13)
h[:, 0, 1], h[:, 0, 2], h[:, 0, 3] = ls0, 1.0, 35.0
st, _ = engine.fit([m], torch.as_tensor(Yg, device="cuda"), None, h)
st = st.cpu().numpy()[:, 0]
ll, nit = st[:, 0], st[:, 2]
print("oracle", fx["ll"][gene, 0], fx["ll_pert"][gene, 0])
print("GPU ll  min %.6f  median %.6f  max %.6f" % (ll.min(), np.median(ll), ll.max()))
print("ll :", np.array2string(np.sort(ll), precision=5, max_line_width=200))
print("nit:", np.array2string(nit[np.argsort(ll)].astype(int), max_line_width=200))
