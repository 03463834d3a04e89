"""Eval-only timing at *fitted* parameters (vs random parameters in bench_eval.py): isolates data dependence."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from bench import synth_batch
from gpcounts_b200 import engine, host
N, M, D = 1000, 64, 592
X, Y = synth_batch(D, N, seed=1)
ls0 = host.default_lengthscale(X)
Z = host.inducing_sets(X, M, ls0); rt, _ = host.restart_table(X)
models = [engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=rt), engine.Model("SVGP", "CONST", "NB", X, Z=Z, restart=rt, handoff_alpha=True)]
h = np.zeros((D, 2, 4)); h[:, :, 0] = host.default_variance(Y, "Negative_binomial", True)[:, None]; h[:, :, 1] = ls0; h[:, :, 2] = 1; h[:, :, 3] = 35
Yd = torch.as_tensor(Y, device="cuda")
for maxiter in [int(a) for a in sys.argv[1:]] or [5000, 30, 3]:
    st, th = engine.fit(models, Yd, None, h, chain=True, opt=engine.default_opt(maxiter=maxiter))
    torch.cuda.synchronize()
    s = st.cpu().numpy()
    print("maxiter %d: in-fit cycles per eval: dyn %.0f const %.0f (us at 1965 MHz: %.1f / %.1f); eval share %.3f"
          % (maxiter, s[:, 0, 12].sum() / s[:, 0, 9].sum(), s[:, 1, 12].sum() / s[:, 1, 9].sum(),
             s[:, 0, 12].sum() / s[:, 0, 9].sum() / 1965, s[:, 1, 12].sum() / s[:, 1, 9].sum() / 1965, s[:, :, 12].sum() / s[:, :, 13].sum()))
    m1 = engine.Model("SVGP", "RBF", "NB", X, Z=Z[0])
    theta = th[:, 0, :m1.P].contiguous()
    for _ in range(2): engine.elbo_grad(m1, Yd, None, theta)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): engine.elbo_grad(m1, Yd, None, theta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("   eval-only at these thetas: %.1f us per eval per SM-slot" % (ms * 1e3 * 148 / D))
