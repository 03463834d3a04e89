"""Separate host-side from device-side causes of batch-dependent fits: same hyp0 bits, engine.fit on the batch vs on one gene."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gpcounts_b200 import engine, host
from tests.helpers import synth_counts
D, N = 101, 20
X, Y = synth_counts(D, N, seed=23)
v_batch = host.default_variance(Y, "Negative_binomial", True)
v_single = np.array([host.default_variance(Y[g:g + 1], "Negative_binomial", True)[0] for g in range(D)])
print("host var0 differing between batch and single shapes:", int((v_batch != v_single).sum()), "of", D)
rt, _ = host.restart_table(X)
models = [engine.Model("VGP", "RBF", "NB", X, restart=rt), engine.Model("VGP", "CONST", "NB", X, restart=rt, handoff_alpha=True)]
h = np.zeros((D, 2, 4)); h[:, :, 0] = v_batch[:, None]; h[:, :, 1] = host.default_lengthscale(X); h[:, :, 2] = 1.0; h[:, :, 3] = 35.0
def fit(sel, n_slots=None):
    st, _ = engine.fit(models, np.ascontiguousarray(Y[sel]), None, np.ascontiguousarray(h[sel]), chain=True, return_theta=False, n_slots=n_slots)
    return st.cpu().numpy()
a = fit(np.arange(D))
for g in (0, 2, 6, 9):
    for ns in (None, 1, 2, 50):
        b = fit(np.array([g, (g + 1) % D, (g + 2) % D]), ns)
        print("gene", g, "n_slots", ns, "equal", np.array_equal(a[g, :, :12], b[0, :, :12]), a[g, :, 0], b[0, :, 0], a[g, :, 2], b[0, :, 2])
a2 = fit(np.arange(D), 37)
print("batch with 37 slots vs default slots: genes differing", int((a[:, :, 0] != a2[:, :, 0]).any(axis=1).sum()))
a3 = fit(np.arange(D)[::-1].copy())
print("batch reversed: genes differing", int((a[:, :, 0] != a3[::-1, :, 0]).any(axis=1).sum()))
