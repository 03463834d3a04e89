"""Micro-benchmark: ELBO+grad evaluations per second of the eval-only kernel (C5 shape) and of a fit."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from bench import synth_batch, svgp_flops
from gpcounts_b200 import engine, host
from tests.helpers import random_theta
from oracle import gp_math as gm
N, M, D = 1000, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 2960
X, Y = synth_batch(D, N, seed=1)
Z = host.inducing_sets(X, M, host.default_lengthscale(X), with_restarts=False)
for kern in ["RBF", "CONST"]:
    model = engine.Model("SVGP", kern, "NB", X, Z=Z)
    spec = gm.ModelSpec(kind="SVGP", kernel=kern, lik="NB", X=X, Z=Z[0])
    theta = torch.as_tensor(random_theta(spec, np.random.default_rng(0), D), device="cuda")
    Yd = torch.as_tensor(Y, device="cuda")
    for _ in range(2):
        engine.elbo_grad(model, Yd, None, theta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        engine.elbo_grad(model, Yd, None, theta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%s eval-only: %d evals in %.2f ms -> %.0f evals/s, %.1f us per eval per SM-slot, %.2f TFLOP/s algorithmic"
          % (kern, D, ms, D / ms * 1e3, ms * 1e3 * 148 / D, D * svgp_flops(M, N) / ms * 1e-9))
