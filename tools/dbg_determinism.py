import numpy as np, sys, torch
sys.path.insert(0,'/root/repo')
from gpcounts_b200 import engine, host
fx=np.load('/root/repo/tests/golden/c5_sample_sparse_nb.npz', allow_pickle=True)
X,Y=fx['X'],fx['Y']; M=int(fx['M'])
ls0=host.default_lengthscale(X)
Z=host.inducing_sets(X,M,ls0); rt,_=host.restart_table(X)
models=[engine.Model("SVGP","RBF","NB",X,Z=Z,restart=rt), engine.Model("SVGP","CONST","NB",X,Z=Z,restart=rt,handoff_alpha=True)]
h=np.zeros((48,2,4)); h[:,:,0]=host.default_variance(Y,"Negative_binomial",True)[:,None]; h[:,:,1]=ls0; h[:,:,2]=1; h[:,:,3]=35
res=[]
for trial,(ns) in enumerate([148,148,48,7,148]):
    # poison the workspace between runs to expose reads of uninitialised memory
    for k,b in engine._ws_cache.items():
        if b is not None: b.view(torch.float64)[: b.numel()//8].fill_(float('nan') if trial%2 else 1e300)
    st,_=engine.fit(models,Y,None,h,chain=True,n_slots=ns)
    res.append(st.cpu().numpy()[:,:,0])
for i in range(1,len(res)):
    d=np.abs(res[i]-res[0]); print('run',i,'max diff vs run0', np.nanmax(d), 'genes differing', (d.max(axis=1)>0).sum(), np.where(d.max(axis=1)>1e-3)[0])
print(res[0][12], res[1][12], res[2][12], res[3][12])
