"""Tiny dense-path fit (full VGP, N = 20, chain=False like Two_samples_test) for compute-sanitizer initcheck / racecheck."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import engine
from tests.helpers import synth_counts
X, Y = synth_counts(3, 20, seed=23)
m = [engine.Model("VGP", "RBF", "NB", X), engine.Model("VGP", "RBF", "NB", X[:10], n_off=0), engine.Model("VGP", "RBF", "NB", X[10:], n_off=10)]
h = np.zeros((3, 3, 4)); h[:, :, 0] = 1.3; h[:, :, 1] = 0.05; h[:, :, 2] = 1.0; h[:, :, 3] = 35.0
st, th = engine.fit(m, Y, None, h, chain=False, opt=engine.default_opt(maxiter=int(sys.argv[1]) if len(sys.argv) > 1 else 25))
torch.cuda.synchronize()
print(st[:, :, 0].cpu().numpy())
