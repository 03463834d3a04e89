import numpy as np, pandas as pd, sys, torch
sys.path.insert(0,'/root/repo')
from gpcounts_b200 import Fit_GPcounts, _lib
f=np.load('/root/repo/tests/golden/one_sample_sparse_zinb.npz', allow_pickle=True)
X,Y=f['X'],f['Y']
cells=['c%d'%i for i in range(X.shape[0])]
gene=int(sys.argv[1]) if len(sys.argv)>1 else 0
buf=torch.zeros(600,8,dtype=torch.float64,device='cuda')
_lib.load().gpc_dbg_set_trace(buf.data_ptr(), 600, gene, 0)
gp=Fit_GPcounts(pd.DataFrame(X,index=cells), pd.DataFrame(Y,index=['g%d'%i for i in range(Y.shape[0])],columns=cells), sparse=True, M=int(f['M']))
df=gp.One_sample_test('Zero_inflated_negative_binomial')
np.set_printoptions(precision=8, suppress=False, linewidth=220)
b=buf.cpu().numpy()
for i in range(45): print(i, '%.6f stp=%.4e gd=%.4e gmax=%.3e'%tuple(b[i,:4]), np.logaddexp(b[i,4:8].clip(-700,700),0), 'NAN' if b[i,7]<-1e299 else '')
