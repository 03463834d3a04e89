import numpy as np, pandas as pd, sys, torch
sys.path.insert(0,'/root/repo')
from gpcounts_b200 import Fit_GPcounts, _lib
fx=np.load('/root/repo/tests/golden/c5_sample_sparse_nb.npz', allow_pickle=True)
X,Y=fx['X'],fx['Y']
g=int(sys.argv[1])
cells=['c%d'%i for i in range(X.shape[0])]
buf=torch.zeros(1500,8,dtype=torch.float64,device='cuda')
_lib.load().gpc_dbg_set_trace(buf.data_ptr(), 1500, g, 0)
gp=Fit_GPcounts(pd.DataFrame(X,index=cells), pd.DataFrame(Y,index=['g%d'%i for i in range(48)],columns=cells), sparse=True, M=int(fx['M']))
df=gp.One_sample_test('Negative_binomial')
print(df.values[g])
b=buf.cpu().numpy()
n=int((b[:,0]!=0).sum())
np.set_printoptions(precision=6, linewidth=200)
for i in list(range(0,3))+list(range(max(3,n-16),n)): print(i, '%.8f stp=%.6e gd=%.4e gmax=%.3e'%tuple(b[i,:4]), np.logaddexp(b[i,4:8].clip(-700,700),0), 'NAN' if b[i,7]<-1e299 else '')
