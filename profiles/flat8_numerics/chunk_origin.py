import sys, os, time, numpy as np
ROOT='/root/repo'; sys.path.insert(0, ROOT)
from oracle import flat_gmm
f32=np.float32
def morton_key(X,bits):
    lo=X.min(0); hi=X.max(0); ext=(hi-lo).max()
    g=np.minimum(((X-lo)/ext*(1<<bits)).astype(np.int64),(1<<bits)-1)
    key=np.zeros(len(X),np.int64)
    for b in range(bits):
        for a in range(3): key|=((g[:,a]>>b)&1)<<(3*b+a)
    return key
def em(X, mu0, iters, s0, sort_bits, CHK, origin='mid'):
    N,J=len(X),len(mu0)
    if sort_bits: X=X[np.argsort(morton_key(X,sort_bits),kind='stable')]
    mu=np.array(mu0,np.float64); cov=np.tile(np.eye(3)*s0,(J,1,1)); logpi=np.full(J,-np.log(J))
    X64=X.astype(np.float64)
    CTA=272
    for _ in range(iters):
        P=np.linalg.inv(cov); ld=np.log(np.linalg.det(cov))
        mu32=mu.astype(f32)
        A=np.zeros((J,10),np.float64)   # fp64 sum over CTAs of fp32 centred partial rows
        for c0 in range(0,N,CTA):
            c1=min(N,c0+CTA)
            acc=np.zeros((J,10),f32)    # centred fp32 accumulators of the CTA
            for a in range(c0,c1,CHK):
                b=min(c1,a+CHK)
                d=X64[a:b,None,:]-mu32.astype(np.float64)[None]
                q=logpi-0.5*(3*np.log(2*np.pi)+ld)-0.5*np.einsum('nja,jab,njb->nj',d,P,d)
                q-=q.max(1,keepdims=True); G=np.exp(q); G/=G.sum(1,keepdims=True)
                # kernel: e (fp32), inv (fp32): psi = inv*phi ; sum e*psi
                E=G.astype(f32); inv=np.ones(b-a,f32)
                o=X[(a+b)//2]
                u=(X[a:b]-o).astype(f32)
                iu=(inv[:,None]*u).astype(f32)
                psi=np.stack([inv,iu[:,0],iu[:,1],iu[:,2],iu[:,0]*u[:,0],iu[:,0]*u[:,1],iu[:,0]*u[:,2],iu[:,1]*u[:,1],iu[:,1]*u[:,2],iu[:,2]*u[:,2]],1).astype(f32)
                S=np.zeros((J,10),f32)
                for p in range(b-a):    # sequential FFMA (fma: product exact) -> emulate in float64 then round
                    S=(S.astype(np.float64)+E[p].astype(np.float64)[:,None]*psi[p].astype(np.float64)[None,:]).astype(f32)
                dl=(mu32-o[None]).astype(f32)      # delta = m - o, fp32
                def fma(a_,b_,c_): return (a_.astype(np.float64)*b_.astype(np.float64)+c_.astype(np.float64)).astype(f32)
                S0=S[:,0]; M1=np.stack([fma(-dl[:,k],S0,S[:,1+k]) for k in range(3)],1)
                idx=[(0,0),(0,1),(0,2),(1,1),(1,2),(2,2)]
                M2=[]
                for k,(i,j) in enumerate(idx):
                    t=fma(-dl[:,i],S[:,1+j],S[:,4+k])
                    t=fma(-dl[:,j],M1[:,i],t)
                    M2.append(t)
                Mc=np.concatenate([S0[:,None],M1,np.stack(M2,1)],1).astype(f32)
                acc=(acc+Mc).astype(f32)
            A+=acc.astype(np.float64)
        S0=A[:,0]; pi=S0/S0.sum(); dm=A[:,1:4]/S0[:,None]
        S2=np.empty((J,3,3))
        for k,(i,j) in enumerate([(0,0),(0,1),(0,2),(1,1),(1,2),(2,2)]): S2[:,i,j]=S2[:,j,i]=A[:,4+k]
        cov=S2/S0[:,None,None]-dm[:,:,None]*dm[:,None,:]
        mu=mu32.astype(np.float64)+dm; logpi=np.log(pi)
    return pi,mu,cov
X=np.load(os.path.join(ROOT,'tests/golden/bun000_xyz.npy')).astype(f32)
J,iters,s0=int(sys.argv[1]),10,1e-4
mu0=X[np.random.default_rng(1).choice(len(X),J,replace=False)]
d=np.load('/tmp/oracle_%d.npz'%J); ow,omu,ocov=d['w'],d['mu'],d['cov']
sb=int(sys.argv[2]); CHK=int(sys.argv[3])
t0=time.time(); pi,mu,cov=em(X,mu0,iters,s0,sb,CHK)
print('chunk',CHK,'sort_bits',sb,'pi %.2e mu %.2e cov %.2e'%(flat_gmm.rel_fro(pi,ow),flat_gmm.rel_fro(mu,omu),flat_gmm.rel_fro(cov,ocov)),'%.0fs'%(time.time()-t0),flush=True)
