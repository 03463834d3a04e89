"""Tiny end-to-end run of every kernel family, sized for `compute-sanitizer --tool racecheck` / memcheck:

    python tools/racecheck_small.py                                     # must exit 0 on its own first
    compute-sanitizer --tool racecheck python tools/racecheck_small.py  # where the sanitizer is available

(The sanitizer is closed on the build pool of round 1; the shared-memory race this script was written to hunt - the
diagonal overwrite in the left-looking Cholesky - was found by inspection and is covered by
tests/test_gpu_properties.py::test_dense_path_is_reproducible_under_contention.)

Fused SVGP evaluation + short fit (M = 64), dense SVGP (M = 72 > 64: two diagonal blocks), dense VGP (N = 70), GPR, SGPR."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import engine, host
from tests.helpers import synth_counts, random_theta
from oracle import gp_math as gm

rng = np.random.default_rng(0)
def run(kind, kernel, lik, N, M, D=2, fit=False):
    X, Y = synth_counts(D, N, seed=3)
    if lik == "GAUSS": Y = np.log(Y + 1.0)
    Z = None
    if kind in ("SVGP", "SGPR"):
        Z = np.sort(rng.uniform(X.min(), X.max(), (M, 1)), axis=0)
    spec = gm.ModelSpec(kind=kind, kernel=kernel, lik=lik, X=X, Z=Z)
    th = random_theta(spec, rng, D)
    m = engine.Model(kind, kernel, lik, X, Z=Z)
    Yd = torch.as_tensor(Y, device="cuda")
    e, g, s = engine.elbo_grad(m, Yd, None, torch.as_tensor(th, device="cuda"))
    msg = "%s/%s/%s N=%d M=%s: elbo finite %d/%d" % (kind, kernel, lik, N, M, int(torch.isfinite(e).sum()), D)
    if fit:
        h = np.zeros((D, 1, 4)); h[:, :, 0] = 1.0; h[:, :, 1] = 0.2; h[:, :, 2] = 1.0; h[:, :, 3] = 35.0
        st, _ = engine.fit([m], Yd, None, h, opt=engine.default_opt(maxiter=12))
        msg += "; fit status %s nit %s" % (st[:, 0, 1].cpu().numpy(), st[:, 0, 2].cpu().numpy())
    print(msg)

run("SVGP", "RBF", "NB", 130, 64, fit=True)
run("SVGP", "RBF", "ZINB", 100, 40)
run("SVGP", "RBF", "NB", 100, 72, fit=True)
run("VGP", "RBF", "NB", 70, None, fit=True)
run("GPR", "RBF", "GAUSS", 70, None)
run("SGPR", "RBF", "GAUSS", 100, 20)
print("done")
