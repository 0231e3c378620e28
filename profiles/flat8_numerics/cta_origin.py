import sys, os, time, numpy as np
ROOT='/root/repo'; sys.path.insert(0, ROOT)
from oracle import flat_gmm
CH=272
def morton_sort(X, bits):
    lo=X.min(0); hi=X.max(0); ext=(hi-lo).max()
    g=np.minimum(((X-lo)/ext*(1<<bits)).astype(np.int64),(1<<bits)-1)
    key=np.zeros(len(X),np.int64)
    for b in range(bits):
        for a in range(3):
            key|=((g[:,a]>>b)&1)<<(3*b+a)
    return np.argsort(key,kind='stable')
def em(X, mu0, iters, s0, sort_bits, mode='unc32'):
    N,J=len(X),len(mu0)
    if sort_bits: X=X[morton_sort(X,sort_bits)]
    mu=np.array(mu0,np.float64); cov=np.tile(np.eye(3)*s0,(J,1,1)); logpi=np.full(J,-np.log(J))
    X64=X.astype(np.float64)
    for _ in range(iters):
        P=np.linalg.inv(cov); ld=np.log(np.linalg.det(cov))
        S0=np.zeros(J); S1=np.zeros((J,3)); S2=np.zeros((J,3,3))
        mu32=mu.astype(np.float32).astype(np.float64)   # kernel's means are fp32
        for a in range(0,N,CH):
            b=min(N,a+CH)
            d=X64[a:b,None,:]-mu32[None]
            q=logpi-0.5*(3*np.log(2*np.pi)+ld)-0.5*np.einsum('nja,jab,njb->nj',d,P,d)
            q-=q.max(1,keepdims=True); G=np.exp(q); G/=G.sum(1,keepdims=True)
            G=G.astype(np.float32)
            o=X[(a+b)//2].astype(np.float32)
            u=(X[a:b]-o).astype(np.float32)
            x,y,z=u[:,0],u[:,1],u[:,2]
            Phi=np.stack([np.ones_like(x),x,y,z,x*x,x*y,x*z,y*y,y*z,z*z],1).astype(np.float32)
            # sequential fp32 accumulation in blocks of 8 points
            o64=o.astype(np.float64)
            if mode=='unc32':
                S=np.zeros((J,10),np.float32)
                for s in range(0,b-a,4):
                    S=(S+ (G[s:s+4].T@Phi[s:s+4]).astype(np.float32)).astype(np.float32)
                S=S.astype(np.float64)
            elif mode=='unc64':
                S=G.astype(np.float64).T@Phi.astype(np.float64)
            elif mode=='cen32':
                d32=(X[a:b,None,:]-mu32.astype(np.float32)[None]).astype(np.float32)   # n,J,3
                g=G[:,:,None]*d32                                                      # gamma*d
                cols=[G, g[:,:,0],g[:,:,1],g[:,:,2], g[:,:,0]*d32[:,:,0], g[:,:,0]*d32[:,:,1], g[:,:,0]*d32[:,:,2], g[:,:,1]*d32[:,:,1], g[:,:,1]*d32[:,:,2], g[:,:,2]*d32[:,:,2]]
                T=np.stack(cols,2).astype(np.float32)                                  # n,J,10
                Sc=np.zeros((J,10),np.float32)
                for s in range(0,b-a,4):
                    Sc=(Sc+T[s:s+4].sum(0,dtype=np.float32)).astype(np.float32)
                Sc=Sc.astype(np.float64)
                S0+=Sc[:,0]; S1+=Sc[:,1:4]
                idx=[(0,0),(0,1),(0,2),(1,1),(1,2),(2,2)]
                for k,(i,j) in enumerate(idx):
                    S2[:,i,j]+=Sc[:,4+k]
                    if i!=j: S2[:,j,i]+=Sc[:,4+k]
                continue
            dl=mu32-o64[None]           # delta = m - o
            s0_=S[:,0]; s1_=S[:,1:4]
            s2_=np.empty((J,3,3)); idx=[(0,0),(0,1),(0,2),(1,1),(1,2),(2,2)]
            for k,(i,j) in enumerate(idx): s2_[:,i,j]=s2_[:,j,i]=S[:,4+k]
            M1=s1_-dl*s0_[:,None]
            M2=s2_-dl[:,:,None]*s1_[:,None,:]-s1_[:,:,None]*dl[:,None,:]+dl[:,:,None]*dl[:,None,:]*s0_[:,None,None]
            # kernel stores centred partial rows in fp32
            M1=M1.astype(np.float32).astype(np.float64); M2=M2.astype(np.float32).astype(np.float64)
            S0+=s0_; S1+=M1; S2+=M2
        pi=S0/S0.sum(); dm=S1/S0[:,None]
        cov=S2/S0[:,None,None]-dm[:,:,None]*dm[:,None,:]
        mu=mu32+dm; logpi=np.log(pi)
    return pi,mu,cov
X=np.load(os.path.join(ROOT,'tests/golden/bun000_xyz.npy')).astype(np.float32)
J,iters,s0=int(sys.argv[1]) if len(sys.argv)>1 else 800,10,1e-4
mu0=X[np.random.default_rng(1).choice(len(X),J,replace=False)]
cf='/tmp/oracle_%d.npz'%J
if os.path.exists(cf):
    d=np.load(cf); ow,omu,ocov=d['w'],d['mu'],d['cov']
else:
    ow,omu,ocov,_=flat_gmm.cpp_fit(X,mu0,iters,sigma0_sq=s0); np.savez(cf,w=ow,mu=omu,cov=ocov)
sb=int(sys.argv[2]); mode=sys.argv[3]
if True:
    t0=time.time(); pi,mu,cov=em(X,mu0,iters,s0,sb,mode)
    print(mode,'sort_bits',sb,'pi %.2e mu %.2e cov %.2e'%(flat_gmm.rel_fro(pi,ow),flat_gmm.rel_fro(mu,omu),flat_gmm.rel_fro(cov,ocov)),'%.0fs'%(time.time()-t0),flush=True)
