import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from syngular.tensor import _sweeps as sw
from syngular_b200 import ops
from oracle import ref_numpy as R, svd_numpy as S
for (n,chi,chiw) in ((16,64,2),(16,32,8),(14,16,4)):
    X,W=bench.make_chain(5 if chiw!=2 else 5,n=n,chi=chi,chiw=chiw)
    ref,spectra,disc=S.apply_round_svd(X,W,chi)
    dref=R.to_dense(ref); scale=np.abs(dref).max()
    Xd=[sw.as_core(c) for c in X]; Wd=[sw.as_core(c) for c in W]
    orig=ops.dominant_subspace
    log=[]
    def spy(A,ne,*a,**k):
        U,info=orig(A,ne,*a,**k)
        h=info.cpu().numpy()
        lam=torch.linalg.eigvalsh(A).flip(0).cpu().numpy()
        log.append((A.shape[0],ne,int(h[7])%1000000 + 0*1,lam[ne-1]/lam[0],(lam[ne-1]-lam[ne])/lam[0], int(h[7])//1000000))
        return U,info
    ops.dominant_subspace=spy
    sw.PURIFY_MIN_N=2*chi
    out,_=sw.apply_round_dm(Xd,Wd,chi)
    ops.dominant_subspace=orig
    sw.PURIFY_MIN_N=0
    outj,_=sw.apply_round_dm(Xd,Wd,chi)
    sw.PURIFY_MIN_N=256
    e1=np.abs(R.to_dense([c.cpu().numpy() for c in out])-dref).max()/scale
    e2=np.abs(R.to_dense([c.cpu().numpy() for c in outj])-dref).max()/scale
    print(sw.PURIFY_STATS); print(n,chi,chiw,"purify err %.2e jacobi err %.2e"%(e1,e2))
    for l in log: print("    n=%d ne=%d iters=%d lam_cut/lam0 %.1e gap/lam0 %.1e lift %d"%l)
