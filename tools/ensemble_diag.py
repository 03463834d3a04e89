import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import engine, host
GOLD = '/root/repo/tests/golden'
for name in sys.argv[1:] or ["one_sample_sparse_zinb"]:
    fx = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=True)
    runs = np.load(os.path.join(GOLD, name + "_runs.npz"))["ll_runs"]
    X, Y, M, lik_name = fx["X"], fx["Y"].astype(int).astype(float), int(fx["M"]), str(fx["lik"])
    lik = {"Negative_binomial": "NB", "Poisson": "POISSON", "Zero_inflated_negative_binomial": "ZINB"}[lik_name]
    D, K = Y.shape[0], 16
    ls0 = host.default_lengthscale(X); Z = host.inducing_sets(X, M, ls0); rt, _ = host.restart_table(X)
    model = engine.Model("SVGP", "RBF", lik, X, Z=Z, restart=rt)
    var0 = host.default_variance(Y, lik_name, True)
    h = np.zeros((D * K, 1, 4)); h[:, 0, 0] = np.repeat(var0, K) * (1.0 + np.tile(np.arange(K), D) * 1e-13); h[:, 0, 1] = ls0; h[:, 0, 2] = 1; h[:, 0, 3] = 35
    st, _ = engine.fit([model], np.repeat(Y, K, axis=0), None, h, chain=True, return_theta=False)
    s = st.cpu().numpy()[:, 0]
    ll = s[:, 0].reshape(D, K); nit = s[:, 2].reshape(D, K)
    for g in range(D):
        print(name, "gene", g, "oracle [%.5f, %.5f] med %.5f | gpu sorted" % (runs[g].min(), runs[g].max(), np.median(runs[g])), np.array2string(np.sort(ll[g]), precision=5), "nit", nit[g].astype(int))
