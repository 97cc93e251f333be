"""Prototype: cycles-to-1e-6 of the V-cycle with different SpMV-based smoothers (scipy, CPU)."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
from gravo_mg_b200 import synth
import gravomg
from oracle import oracle
n=int(sys.argv[1]) if len(sys.argv)>1 else 300
V,F=synth.torus_grid(n,n)
V,S,M,neigh=synth.mesh_operators(V,F)
lhs,rhs=synth.poisson_system(S,M)
s=gravomg.MultigridSolver(V,neigh,M,lower_bound=500)
U=[u.tocsr() for u in s.prolongation_matrices]
A=[lhs]
for u in U: A.append((u.T@A[-1]@u).tocsr())
R=[u.T.tocsr() for u in U]
coarse=sla.splu(A[-1].tocsc())
m=M.diagonal()
def resn(x): 
    r=lhs@x-rhs
    return np.sqrt((m[:,None]*r*r).sum()/(m[:,None]*rhs*rhs).sum())
Dinv=[1/a.diagonal() for a in A]
# estimate rho(D^-1 A) per level
rho=[]
for a,d in zip(A,Dinv):
    v=np.random.default_rng(0).standard_normal(a.shape[0])
    for _ in range(30):
        v=d*(a@v); lam=np.linalg.norm(v); v/=lam
    rho.append(lam)
print("rho(D^-1A):",[f"{r:.3f}" for r in rho])
def jacobi(k,b,x,omegas):
    for w in omegas:
        x=x+w*Dinv[k][:,None]*(b-A[k]@x)
    return x
def cheb_omegas(k,deg,alpha):
    # Chebyshev roots on [rho/alpha, rho*1.05] -> sequence of Jacobi damping factors (Richardson form)
    b_=rho[k]*1.05; a_=b_/alpha
    j=np.arange(deg)
    roots=(a_+b_)/2+(b_-a_)/2*np.cos(np.pi*(2*j+1)/(2*deg))
    return list(1/roots)
def l1dinv(k):
    a=A[k]; return 1/np.asarray(abs(a).sum(1)).ravel()
def solve(smooth_pre,smooth_post,maxit=100,tol=1e-6):
    def cyc(k,b,x):
        x=smooth_pre(k,b,x)
        r=b-A[k]@x
        rc=R[k]@r
        if k==len(U)-1: e=coarse.solve(rc)
        else: e=cyc(k+1,rc,np.zeros_like(rc))
        x=x+U[k]@e
        return smooth_post(k,b,x)
    x=rhs.copy(); hist=[]
    for it in range(maxit):
        x=cyc(0,rhs,x); hist.append(resn(x))
        if hist[-1]<=tol: break
    return len(hist),hist[-1]
for w in (0.5,0.6,2/3,0.7,0.8,0.9):
    print("jacobi w=%.3f 2+2"%w, solve(lambda k,b,x:jacobi(k,b,x,[w,w]),lambda k,b,x:jacobi(k,b,x,[w,w])))
for alpha in (3,4,6,8,12):
    print("cheb2 alpha",alpha, solve(lambda k,b,x:jacobi(k,b,x,cheb_omegas(k,2,alpha)),lambda k,b,x:jacobi(k,b,x,cheb_omegas(k,2,alpha)[::-1])))
for alpha in (4,8,12,20):
    print("cheb3 alpha",alpha, solve(lambda k,b,x:jacobi(k,b,x,cheb_omegas(k,3,alpha)),lambda k,b,x:jacobi(k,b,x,cheb_omegas(k,3,alpha)[::-1])))
for alpha in (8,16,30):
    print("cheb4 alpha",alpha, solve(lambda k,b,x:jacobi(k,b,x,cheb_omegas(k,4,alpha)),lambda k,b,x:jacobi(k,b,x,cheb_omegas(k,4,alpha)[::-1])))
og=oracle.OracleSolver(M,[u.tocsc() for u in U],tolerance=1e-6,smoother="gs"); og.solve(lhs,rhs); print("GS 2+2 (reference):",og.solver_timing["iterations"])
