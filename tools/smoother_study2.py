import sys; sys.path.insert(0,'/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
from gravo_mg_b200 import synth
import gravomg
def study(name,V,F,kind,lb,cloud=None):
    if cloud is None:
        V,S,M,neigh=synth.mesh_operators(V,F)
    else:
        S,M=cloud; neigh=gravomg.neighbors_from_stiffness(S)
    if kind=="poisson": lhs,rhs=synth.poisson_system(S,M)
    else: lhs,rhs=synth.smoothing_system(V,S,M)
    s=gravomg.MultigridSolver(V,neigh,M,lower_bound=lb)
    U=[u.tocsr() for u in s.prolongation_matrices]
    A=[lhs]
    for u in U: A.append((u.T@A[-1]@u).tocsr())
    R=[u.T.tocsr() for u in U]
    coarse=sla.splu(A[-1].tocsc())
    m=M.diagonal()
    def resn(x):
        r=lhs@x-rhs
        return np.sqrt((m[:,None]*r*r).sum(0)/(m[:,None]*rhs*rhs).sum(0)).max()
    Dinv=[1/a.diagonal() for a in A]
    rho=[float((np.asarray(abs(a).sum(1)).ravel()*d).max()) for a,d in zip(A,Dinv)]  # Gershgorin
    def jacobi(k,b,x,omegas):
        for w in omegas: x=x+w*Dinv[k][:,None]*(b-A[k]@x)
        return x
    def cheb(k,deg,alpha):
        b_=rho[k]; a_=b_/alpha; j=np.arange(deg)
        return list(1/((a_+b_)/2+(b_-a_)/2*np.cos(np.pi*(2*j+1)/(2*deg))))
    def solve(pre,post,maxit=100,tol=1e-6):
        def cyc(k,b,x):
            x=pre(k,b,x); r=b-A[k]@x; rc=R[k]@r
            e=coarse.solve(rc) if k==len(U)-1 else cyc(k+1,rc,np.zeros_like(rc))
            return post(k,b,x+U[k]@e)
        x=rhs.copy()
        for it in range(maxit):
            x=cyc(0,rhs,x); r=resn(x)
            if not np.isfinite(r): return (it+1,"DIVERGED")
            if r<=tol: break
        return it+1
    out={}
    out["jac2/3"]=solve(lambda k,b,x:jacobi(k,b,x,[2/3]*2),lambda k,b,x:jacobi(k,b,x,[2/3]*2))
    for deg in (2,3):
        for alpha in (6,8,12,16,24):
            out[f"cheb{deg}a{alpha}"]=solve(lambda k,b,x:jacobi(k,b,x,cheb(k,deg,alpha)),lambda k,b,x:jacobi(k,b,x,cheb(k,deg,alpha)[::-1]))
    print(name,[A_.shape[0] for A_ in A],"gersh rho",[f"{r:.2f}" for r in rho]); print("   ",out,flush=True)
which=sys.argv[1]
if which=="ico": 
    study("ico10k poisson",*synth.icosphere(5),"poisson",1000); study("ico10k smoothing",*synth.icosphere(5),"smoothing",1000)
    study("ico40k poisson lb200",*synth.icosphere(6),"poisson",200)
if which=="torus": study("torus%s"%sys.argv[2],*synth.torus_grid(int(sys.argv[2]),int(sys.argv[2])),sys.argv[3],500)
if which=="cloud":
    n=int(sys.argv[2]); P=synth.torus_cloud(n,seed=0); nbr=synth.knn_grid(P,n,k=8); L,M=synth.knn_graph_laplacian(nbr)
    study("cloud%d"%n,P,None,"poisson",500,cloud=(L,M))
