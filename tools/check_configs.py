"""Sanity + timing of the non-headline BASELINE.json shapes on synthetic data:
C2 MouseHSC-like sparse (N=2660, M=133, NB and ZINB) and C4 branching (N=532, 25 candidate branching times)."""
import sys, time, numpy as np, pandas as pd, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import Fit_GPcounts
rng = np.random.default_rng(0)
which = sys.argv[1:] or ["c2", "c4"]
if "c2" in which:
    N, D = 2660, 19
    t = np.sort(rng.uniform(0, 1, N))
    f = rng.normal(1, 1, (D, 1)) + np.sin(4 * t)[None, :] * (np.arange(D)[:, None] % 2)
    Y = rng.negative_binomial(2.0, 2.0 / (2.0 + np.exp(f))).astype(float)
    cells = ["c%d" % i for i in range(N)]
    for lik in ("Negative_binomial", "Zero_inflated_negative_binomial"):
        gp = Fit_GPcounts(pd.DataFrame(t, index=cells), pd.DataFrame(Y, index=["g%d" % i for i in range(D)], columns=cells), sparse=True)
        t0 = time.time(); df = gp.One_sample_test(lik); dt = time.time() - t0
        st = gp.fit_stats
        print("C2 %s: M=%d, %.1f s (kernel %.1f s), %d/%d fits ok, evals %d, LLR range %.2f..%.2f" % (
            lik, gp.M, dt, gp.last_fit_seconds, (st.status == 0).sum(), len(st), st.n_feval_total.sum(),
            np.nanmin(df.values[:, 2]), np.nanmax(df.values[:, 2])))
if "c4" in which:
    N = 532
    t = np.sort(rng.uniform(0, 1, N)); labels = rng.integers(1, 3, N).astype(float)
    f = np.log(20.0) + np.where((labels == 2) & (t > 0.6), 2.5 * (t - 0.6), 0.0)
    Y = rng.negative_binomial(5.0, 5.0 / (5.0 + np.exp(f)))[None, :].astype(float)
    cells = ["c%d" % i for i in range(N)]
    gp = Fit_GPcounts(pd.DataFrame(t, index=cells), pd.DataFrame(Y, index=["gene"], columns=cells))
    t0 = time.time(); res = gp.Infer_branching_location(labels, bins_num=25); dt = time.time() - t0
    print("C4 branching N=%d, 25 bins: %.1f s; MAP branching time %.3f (truth 0.6), logBF %.2f, max prob %.3f" % (
        N, dt, res["branching_location"], res["logBayesFactor"], res["branching_probability"].max()))
