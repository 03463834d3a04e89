"""Evaluation noise of the ELBO near an optimum: scatter of ELBO(theta_hat * (1 + k 1e-15)), k = 0..31, relative to |ELBO|.
SciPy's stopping rule compares successive values at 2.2e-9 |f|, so the noise must stay well below that.

    python tools/elbo_noise.py            # GPU evaluator (and saves theta_hat to gpurun_out/theta_hat.npz)
    python tools/elbo_noise.py oracle     # CPU oracle at the saved theta_hat
"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from bench import synth_batch
mode = sys.argv[1] if len(sys.argv) > 1 else "gpu"
N, M, D = 1000, 64, 8
X, Y = synth_batch(D, N, seed=77)
if mode == "gpu":
    import torch
    from gpcounts_b200 import engine, host
    ls0 = host.default_lengthscale(X)
    Z = host.inducing_sets(X, M, ls0); rt, _ = host.restart_table(X)
    m = engine.Model("SVGP", "RBF", "NB", X, Z=Z, restart=rt)
    h = np.zeros((D, 1, 4)); h[:, 0, 0] = host.default_variance(Y, "Negative_binomial", True); h[:, 0, 1] = ls0; h[:, 0, 2] = 1; h[:, 0, 3] = 35
    Yd = torch.as_tensor(Y, device="cuda")
    st, th = engine.fit([m], Yd, None, h, return_theta=True)
    th = th[:, 0].cpu().numpy()
    np.savez("/root/repo/gpurun_out/theta_hat.npz", theta=th, Z=Z[0])
    m1 = engine.Model("SVGP", "RBF", "NB", X, Z=Z[0])
    out = []
    for k in range(32):
        e, g, s = engine.elbo_grad(m1, Yd, None, torch.as_tensor(th * (1.0 + k * 1e-15), device="cuda"))
        out.append(e.cpu().numpy())
    out = np.array(out)
else:
    import torch
    torch.set_num_threads(8)
    from oracle import gp_math as gm
    d = np.load("/root/repo/gpurun_out/theta_hat.npz")
    th, Z0 = d["theta"], d["Z"]
    spec = gm.ModelSpec(kind="SVGP", kernel="RBF", lik="NB", X=X, Z=Z0)
    out = np.array([[float(gm.elbo(torch.as_tensor(th[g] * (1.0 + k * 1e-15)), spec, Y[g], None)) for g in range(D)] for k in range(32)])
sd = out.std(axis=0) / np.abs(out.mean(axis=0))
print(mode, "relative ELBO scatter per gene:", np.array2string(sd, precision=2))
print(mode, "max |ELBO_k - ELBO_0| / |ELBO|:", np.array2string(np.abs(out - out[0]).max(axis=0) / np.abs(out[0]), precision=2))
