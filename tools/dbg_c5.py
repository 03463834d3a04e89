import numpy as np, pandas as pd, sys
sys.path.insert(0,'/root/repo')
from gpcounts_b200 import Fit_GPcounts, host
fx=np.load('/root/repo/tests/golden/c5_sample_sparse_nb.npz', allow_pickle=True)
X,Y=fx['X'],fx['Y']
cells=['c%d'%i for i in range(X.shape[0])]
gp=Fit_GPcounts(pd.DataFrame(X,index=cells), pd.DataFrame(Y,index=['g%d'%i for i in range(Y.shape[0])],columns=cells), sparse=True, M=int(fx['M']))
df=gp.One_sample_test('Negative_binomial')
got,ref,ref2=df.values,fx['ll'],fx['ll_pert']
np.set_printoptions(precision=4, suppress=True, linewidth=200)
st=gp.fit_stats
nit=st[st.model=='dynamic'].n_iter.values
for g in range(48):
    print(g, 'dyn got %.4f ref %.4f ref2 %.4f | const got %.4f ref %.4f | LLR got %.4f ref %.4f ref2 %.4f | nit %d ref %d ref2 %d' % (got[g,0],ref[g,0],ref2[g,0],got[g,1],ref[g,1],got[g,2],ref[g,2],ref2[g,2], nit[g], fx['nit'][g,0], fx['nit_pert'][g,0]))
q=host.qvalue(host.p_values(np.maximum(got[:,2],0))); qr=host.qvalue(host.p_values(np.maximum(ref[:,2],0)))
print('DE agree', ((q<0.05)==(qr<0.05)).mean(), (q<0.05).sum(), (qr<0.05).sum())
