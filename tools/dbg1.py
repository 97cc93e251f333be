import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import gravomg
from gravo_mg_b200 import synth
from oracle import oracle
V,F=synth.icosphere(4)
V,S,M,neigh=synth.mesh_operators(V,F)
lhs,rhs=synth.poisson_system(S,M)
for K in (1,3):
  for path in (0,1):
    s=gravomg.MultigridSolver(V,neigh,M,lower_bound=40).solver
    s.set_option("kernel_path",path)
    n=lhs.shape[0]
    s.stage(lhs,np.zeros((n,K)))
    print("staged",K,path,flush=True)
    rng=np.random.default_rng(5)
    s.level_op("residual",0,np.zeros((n,K)),np.zeros((n,K)))
    print("reduced",flush=True)
    info=s.level_info(); print(info,flush=True)
    for k in range(len(info)):
        A=s.level_matrix(k)
        x=rng.standard_normal((A.shape[0],K)); b=rng.standard_normal((A.shape[0],K))
        r=s.level_op("residual",k,x,b); print("res",k,np.abs(r-oracle.residual(A,b,x)).max(),flush=True)
        if k<len(info)-1:
            j=s.level_op("jacobi",k,x,b,sweeps=2); print("jac",k,np.abs(j-oracle.jacobi(A.T.tocsc(),b,x,2,2/3)).max(),flush=True)
            rr=s.level_op("restrict",k,x); print("restrict",k,rr.shape,flush=True)
            e=rng.standard_normal((info[k+1]["rows"],K))
            pp=s.level_op("prolong_add",k,e,x); print("prolong",k,pp.shape,flush=True)
        else:
            c=s.level_op("coarse",k,b); print("coarse",np.abs(A@c-b).max(),flush=True)
