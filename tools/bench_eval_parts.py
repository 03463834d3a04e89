"""Which part of theta makes evaluations slow?  Mix fitted and random parameter blocks."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from bench import synth_batch
from gpcounts_b200 import engine, host
from tests.helpers import random_theta
from oracle import gp_math as gm
N, M, D = 1000, 64, 592
X, Y = synth_batch(D, N, seed=1)
ls0 = host.default_lengthscale(X)
Z = host.inducing_sets(X, M, ls0); rt, _ = host.restart_table(X)
models = [engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=rt)]
h = np.zeros((D, 1, 4)); h[:, :, 0] = host.default_variance(Y, "Negative_binomial", True)[:, None]; h[:, :, 1] = ls0; h[:, :, 2] = 1; h[:, :, 3] = 35
Yd = torch.as_tensor(Y, device="cuda")
st, th = engine.fit(models, Yd, None, h, chain=True, opt=engine.default_opt(maxiter=30))
fit = th[:, 0].cpu().numpy()
spec = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Z[0])
rnd = random_theta(spec, np.random.default_rng(0), D)
m1 = engine.Model("SVGP", "RBF", "NB", X, Z=Z[0])
def timeit(theta, label):
    t = torch.as_tensor(theta, device="cuda")
    for _ in range(2): engine.elbo_grad(m1, Yd, None, t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): e, g, s = engine.elbo_grad(m1, Yd, None, t)
    e1.record(); torch.cuda.synchronize()
    print("%-34s %.1f us/eval   (finite elbo: %d/%d)" % (label, e0.elapsed_time(e1) / 3 * 1e3 * 148 / D, int(torch.isfinite(e).sum()), D))
timeit(rnd, "random")
timeit(fit, "fitted(30 its)")
for name, sl in [("hyper", slice(0, 4)), ("q_mu", slice(4, 4 + M)), ("Lq", slice(4 + M, None))]:
    t = rnd.copy(); t[:, sl] = fit[:, sl]; timeit(t, "random + fitted " + name)
    t = fit.copy(); t[:, sl] = rnd[:, sl]; timeit(t, "fitted + random " + name)
print("fitted hyper medians (var, ls, alpha):", np.median(np.logaddexp(fit[:, :3], 0), axis=0), "Lq |offdiag| median", np.median(np.abs(fit[:, 4 + M:])))
