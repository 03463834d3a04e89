import numpy as np, sys, torch
sys.path.insert(0,'/root/repo')
from gpcounts_b200 import engine
f=np.load('/root/repo/tests/golden/one_sample_sparse_zinb.npz', allow_pickle=True)
X,Y=f['X'],f['Y']
th0=np.load('tools/tmp/z_th0.npy'); g0=np.load('tools/tmp/z_g0.npy'); Z=np.load('tools/tmp/z_Z.npy')
m=engine.Model('SVGP','RBF','ZINB',X,Z=Z)
e,g,s=engine.elbo_grad(m, Y[:1], None, th0[None])
g=g.cpu().numpy()[0]
print(e.item(), s.item())
print(g[:8]); print(g0[:8])
d=np.abs(g-g0); i=np.argsort(-d)[:10]; print(i, d[i], g[i], g0[i])
