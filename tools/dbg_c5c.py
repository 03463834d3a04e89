import numpy as np, pandas as pd, sys, torch
sys.path.insert(0,'/root/repo')
from gpcounts_b200 import Fit_GPcounts, engine, host
fx=np.load('/root/repo/tests/golden/c5_sample_sparse_nb.npz', allow_pickle=True)
X,Y=fx['X'],fx['Y']; M=int(fx['M'])
cap={}
orig=engine.fit
def spy(models,Yd,scale,hyp0,**kw):
    cap.update(models=models,Y=np.array(Yd),hyp0=np.array(hyp0),kw=kw)
    return orig(models,Yd,scale,hyp0,**kw)
engine.fit=spy
cells=['c%d'%i for i in range(X.shape[0])]
gp=Fit_GPcounts(pd.DataFrame(X,index=cells), pd.DataFrame(Y,index=['g%d'%i for i in range(48)],columns=cells), sparse=True, M=M)
df=gp.One_sample_test('Negative_binomial')
print('api', df.values[12])
ls0=host.default_lengthscale(X)
Z=host.inducing_sets(X,M,ls0); rt,_=host.restart_table(X)
h=np.zeros((48,2,4)); h[:,:,0]=host.default_variance(Y,"Negative_binomial",True)[:,None]; h[:,:,1]=ls0; h[:,:,2]=1; h[:,:,3]=35
m=cap['models']
print('Y equal', np.array_equal(cap['Y'],Y), 'hyp0 equal', np.array_equal(cap['hyp0'],h), np.abs(cap['hyp0']-h).max())
print('Z equal', np.array_equal(m[0].Z.cpu().numpy(),Z), 'X equal', np.array_equal(m[0].X.cpu().numpy(),X), 'rt', np.array_equal(m[0].restart.cpu().numpy(),rt), 'kw', cap['kw'].keys(), [ (mm.kind,mm.kernel,mm.lik,mm.train_mask,mm.handoff_alpha,mm.n_zsets,mm.M,mm.N,mm.d) for mm in m])
st,_=orig(m,cap['Y'],None,cap['hyp0'],chain=True); print('replay captured', st[12,:,0].cpu().numpy())
st,_=orig(m,cap['Y'],None,cap['hyp0'],chain=True,opt=cap['kw']['opt']); print('replay captured+opt', st[12,:,0].cpu().numpy())
o=cap['kw']['opt']; print('opt', o.m,o.maxiter,o.maxfun,o.maxls,o.ftol,o.gtol)
