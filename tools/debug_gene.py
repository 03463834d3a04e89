import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import engine, host
fx = np.load('/root/repo/tests/golden/two_samples_vgp_nb.npz', allow_pickle=True)
X, Y = fx["X"], fx["Y"].astype(int).astype(float)
h = X.shape[0] // 2
X1, Y1 = X[:h], Y[:, :h]
m = [engine.Model("VGP", "RBF", "NB", X1)]           # no restart table: attempt 0 only
hyp = np.zeros((3, 1, 4)); hyp[:, 0, 0] = host.default_variance(Y1, "Negative_binomial", True); hyp[:, 0, 1] = host.default_lengthscale(X1); hyp[:, 0, 2] = 1; hyp[:, 0, 3] = 35
for variant in (-1, 1):
    cap = 400
    buf = torch.zeros(cap * 8, dtype=torch.float64, device="cuda")
    st, th = engine.fit(m, Y1, None, hyp, opt=engine.default_opt(variant=variant, trace=buf.data_ptr(), trace_cap=cap, trace_gene=2, trace_model=0))
    s = st.cpu().numpy()[:, 0]
    print("variant", variant, "LL", s[:, 0], "status", s[:, 1], "nit", s[:, 2], "nfev", s[:, 3], "task", s[:, 11], "gnorm", s[:, 10], "alpha", s[:, 7])
    t = buf.cpu().numpy().reshape(cap, 8); n = int((t[:, 0] != 0).sum())
    for r in t[max(0, n - 14):n]:
        print("   f %.10f stp %.4g gd %.4g gmax %.4g  hyp %.6f %.6f %.6f" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6]))
