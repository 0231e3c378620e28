import sys, os, time, numpy as np
ROOT='/root/repo'; sys.path.insert(0, ROOT)
from oracle import flat_gmm
f32=np.float32; f64=np.float64
def fma(a,b,c): return (a.astype(f64)*b.astype(f64)+c.astype(f64)).astype(f32)
def mul(a,b): return (a.astype(f64)*b.astype(f64)).astype(f32)
def morton_key(X,bits):
    lo=X.min(0); hi=X.max(0); ext=(hi-lo).max()
    g=np.minimum(((X-lo)/ext*(1<<bits)).astype(np.int64),(1<<bits)-1)
    key=np.zeros(len(X),np.int64)
    for b in range(bits):
        for a in range(3): key|=((g[:,a]>>b)&1)<<(3*b+a)
    return key
LOG2E=1.4426950408889634
def em(X, mu0, iters, s0, mode):
    N,J=len(X),len(mu0)
    X=X[np.argsort(morton_key(X,4),kind='stable')]
    mu=np.array(mu0,f64); cov=np.tile(np.eye(3)*s0,(J,1,1)); logpi=np.full(J,-np.log(J))
    CTA=272; CHK=32
    for _ in range(iters):
        P=np.linalg.inv(cov); ld=np.log(np.linalg.det(cov))
        mu32=mu.astype(f32)
        c2=(LOG2E*(logpi-0.5*ld-1.5*np.log(2*np.pi)))
        cref=c2.max()
        h=-0.5*LOG2E
        A32={k:(h*P[:,i,j]*(1 if i==j else 2)).astype(f32) for k,(i,j) in {'xx':(0,0),'yy':(1,1),'zz':(2,2),'xy':(0,1),'xz':(0,2),'yz':(1,2)}.items()}
        c2s=(c2.astype(f32)-f32(cref)).astype(f32)
        if mode=='chol':
            # -A (fp32 values) -> fp64 cholesky, U upper: r = U d ; |r|^2 = d^T(-A)d
            M=np.zeros((J,3,3))
            M[:,0,0]=-A32['xx']; M[:,1,1]=-A32['yy']; M[:,2,2]=-A32['zz']
            M[:,0,1]=M[:,1,0]=-A32['xy']/2; M[:,0,2]=M[:,2,0]=-A32['xz']/2; M[:,1,2]=M[:,2,1]=-A32['yz']/2
            Lc=np.linalg.cholesky(M)           # lower: M = Lc Lc^T ; r = Lc^T d
            U=np.transpose(Lc,(0,2,1)).astype(f32)   # r_i = sum_j U[i,j] d_j, upper triangular
            k4=(-c2s).astype(f32)
        A=np.zeros((J,10),f64)
        for c0 in range(0,N,CTA):
            c1=min(N,c0+CTA); acc=np.zeros((J,10),f32)
            for a in range(c0,c1,CHK):
                b=min(c1,a+CHK); o=X[(a+b)//2]
                x=X[a:b]
                if mode=='quad':
                    dx=(x[:,0:1]-mu32[None,:,0]).astype(f32); dy=(x[:,1:2]-mu32[None,:,1]).astype(f32); dz=(x[:,2:3]-mu32[None,:,2]).astype(f32)
                    t0=mul(A32['xz'][None],dz); t0=fma(A32['xy'][None],dy,t0); t0=fma(A32['xx'][None],dx,t0)
                    t1=mul(A32['yz'][None],dz); t1=fma(A32['yy'][None],dy,t1)
                    t2=mul(A32['zz'][None],dz)
                    q=fma(dz,t2,np.broadcast_to(c2s[None],dz.shape)); q=fma(dy,t1,q); q=fma(dx,t0,q)
                else:
                    u=(x-o).astype(f32)
                    om=(o[None]-mu32).astype(f32)                      # o - m
                    b3=mul(U[:,2,2],om[:,2])
                    b2=fma(U[:,1,1],om[:,1],mul(U[:,1,2],om[:,2]))
                    b1=fma(U[:,0,0],om[:,0],fma(U[:,0,1],om[:,1],mul(U[:,0,2],om[:,2])))
                    ux,uy,uz=u[:,0:1],u[:,1:2],u[:,2:3]
                    sh=(len(u),J)
                    r3=fma(np.broadcast_to(U[None,:,2,2],sh),np.broadcast_to(uz,sh),np.broadcast_to(b3[None],sh))
                    r2=fma(np.broadcast_to(U[None,:,1,1],sh),np.broadcast_to(uy,sh),fma(np.broadcast_to(U[None,:,1,2],sh),np.broadcast_to(uz,sh),np.broadcast_to(b2[None],sh)))
                    r1=fma(np.broadcast_to(U[None,:,0,0],sh),np.broadcast_to(ux,sh),fma(np.broadcast_to(U[None,:,0,1],sh),np.broadcast_to(uy,sh),fma(np.broadcast_to(U[None,:,0,2],sh),np.broadcast_to(uz,sh),np.broadcast_to(b1[None],sh))))
                    s=fma(r1,r1,np.broadcast_to(k4[None],sh)); s=fma(r2,r2,s); s=fma(r3,r3,s)
                    q=-s
                E=np.exp2(q.astype(f64)).astype(f32)
                inv=(1.0/E.astype(f64).sum(1)).astype(f32)
                u=(x-o).astype(f32)
                iu=(inv[:,None]*u).astype(f32)
                psi=np.stack([inv,iu[:,0],iu[:,1],iu[:,2],iu[:,0]*u[:,0],iu[:,0]*u[:,1],iu[:,0]*u[:,2],iu[:,1]*u[:,1],iu[:,1]*u[:,2],iu[:,2]*u[:,2]],1).astype(f32)
                S=np.zeros((J,10),f32)
                for p in range(b-a):
                    S=(S.astype(f64)+E[p].astype(f64)[:,None]*psi[p].astype(f64)[None,:]).astype(f32)
                dl=(mu32-o[None]).astype(f32)
                S0=S[:,0]; M1=np.stack([fma(-dl[:,k],S0,S[:,1+k]) for k in range(3)],1)
                idx=[(0,0),(0,1),(0,2),(1,1),(1,2),(2,2)]; M2=[]
                for k,(i,j) in enumerate(idx):
                    t=fma(-dl[:,i],S[:,1+j],S[:,4+k]); t=fma(-dl[:,j],M1[:,i],t); M2.append(t)
                Mc=np.concatenate([S0[:,None],M1,np.stack(M2,1)],1).astype(f32)
                acc=(acc+Mc).astype(f32)
            A+=acc.astype(f64)
        S0=A[:,0]; pi=S0/S0.sum(); dm=A[:,1:4]/S0[:,None]
        S2=np.empty((J,3,3))
        for k,(i,j) in enumerate([(0,0),(0,1),(0,2),(1,1),(1,2),(2,2)]): S2[:,i,j]=S2[:,j,i]=A[:,4+k]
        cov=S2/S0[:,None,None]-dm[:,:,None]*dm[:,None,:]
        mu=mu32.astype(f64)+dm; logpi=np.log(pi)
    return pi,mu,cov
X=np.load(os.path.join(ROOT,'tests/golden/bun000_xyz.npy')).astype(f32)
J,iters,s0=int(sys.argv[1]),10,float(sys.argv[3]) if len(sys.argv)>3 else 1e-4
mu0=X[np.random.default_rng(1).choice(len(X),J,replace=False)]
cf='/tmp/oracle_%d_%g.npz'%(J,s0)
if os.path.exists(cf):
    d=np.load(cf); ow,omu,ocov=d['w'],d['mu'],d['cov']
else:
    ow,omu,ocov,_=flat_gmm.cpp_fit(X,mu0,iters,sigma0_sq=s0); np.savez(cf,w=ow,mu=omu,cov=ocov)
t0=time.time(); pi,mu,cov=em(X,mu0,iters,s0,sys.argv[2])
print(sys.argv[2],'s0',s0,'pi %.2e mu %.2e cov %.2e'%(flat_gmm.rel_fro(pi,ow),flat_gmm.rel_fro(mu,omu),flat_gmm.rel_fro(cov,ocov)),'%.0fs'%(time.time()-t0),flush=True)
