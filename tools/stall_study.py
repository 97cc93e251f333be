"""Why the Chebyshev-Jacobi V-cycle stalls at ~1e-6 on >= 4 M-vertex Poisson systems (tau = 1e-6) and
what fixes it. numpy restatement of the device cycle with two evaluations of the row product A x:

  naive   sum_j A_ij x_j                                   (rounding ~ eps |A| |x|, x ~ 1/(tau sqrt N))
  diff    sum_{j != i} A_ij (x_j - x_i) + s_i x_i          (s_i = exact row sum; rounding ~ eps |A| |dx|)

usage: python tools/stall_study.py <n_side> [maxit] [area]
  area: total surface area the mesh is scaled to before the operators are built (default 1, the
        reference's normalisation; bench.py's weak-scaling runs use area = vertices / 1e6, i.e. the
        mesh spacing of the 1 M-vertex system)
"""
import sys
import time

sys.path.insert(0, "/root/repo")
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import gravomg
from gravo_mg_b200 import synth

n_side = int(sys.argv[1])
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 30
V, F = synth.torus_grid(n_side, n_side)
area = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
V, S, M, neigh = synth.mesh_operators(V, F)
if area != 1.0:
    V = V * np.sqrt(area)
    V, S, M, neigh = synth.mesh_operators(V, F, normalize=False)
lhs, rhs = synth.poisson_system(S, M)
t0 = time.time()
s = gravomg.MultigridSolver(V, neigh, M, lower_bound=500)
U = [u.tocsr() for u in s.prolongation_matrices]
print("hierarchy", [lhs.shape[0]] + [u.shape[1] for u in U], f"{time.time() - t0:.1f}s", flush=True)
A = [lhs.tocsr()]
for u in U:
    A.append((u.T @ A[-1] @ u).tocsr())
R = [u.T.tocsr() for u in U]
coarse = sla.splu(A[-1].tocsc())
m = M.diagonal()
Dinv = [1 / a.diagonal() for a in A]
rho = [float((np.asarray(abs(a).sum(1)).ravel() * d).max()) for a, d in zip(A, Dinv)]

# exact row sums of level 0 (long double accumulate per row), and the off-diagonal part
A0 = A[0]
rows = np.repeat(np.arange(A0.shape[0]), np.diff(A0.indptr))
rs = np.zeros(A0.shape[0], dtype=np.longdouble)
np.add.at(rs, rows, A0.data.astype(np.longdouble))
rowsum = rs.astype(np.float64)
off = A0.copy()
off.setdiag(0.0)
off.eliminate_zeros()
offsum = np.asarray(off.sum(1)).ravel()  # only used as a cross-check
print("rowsum range", rowsum.min(), rowsum.max(), "tau*m", 1e-6 * m.min(), 1e-6 * m.max())


def ax_naive(x):
    return A0 @ x


def ax_diff(x):
    # sum_{j != i} A_ij x_j - (sum_{j != i} A_ij) x_i is NOT the same rounding as sum A_ij (x_j - x_i);
    # do the real thing entry-wise
    d = x[A0.indices, 0] - x[rows, 0]
    acc = np.zeros(A0.shape[0])
    np.add.at(acc, rows, A0.data * d)  # diagonal entries contribute A_ii * 0
    return (acc + rowsum * x[:, 0])[:, None]


def cheb(k, deg, alpha=10.0):
    b_ = rho[k]
    a_ = b_ / alpha
    j = np.arange(deg)
    return list(1 / ((a_ + b_) / 2 + (b_ - a_) / 2 * np.cos(np.pi * (2 * j + 1) / (2 * deg))))


def run(ax0, label):
    def apply(k, x):
        return ax0(x) if k == 0 else A[k] @ x

    def jac(k, b, x, om):
        for w in om:
            x = x + w * Dinv[k][:, None] * (b - apply(k, x))
        return x

    def cyc(k, b, x):
        x = jac(k, b, x, cheb(k, 2))
        r = b - apply(k, x)
        rc = R[k] @ r
        e = coarse.solve(rc) if k == len(U) - 1 else cyc(k + 1, rc, np.zeros_like(rc))
        return jac(k, b, x + U[k] @ e, cheb(k, 2)[::-1])

    def resn(x, ax):
        r = ax(x) - rhs
        return float(np.sqrt((m[:, None] * r * r).sum() / (m[:, None] * rhs * rhs).sum()))

    x = rhs.copy()
    hist = []
    for it in range(maxit):
        x = cyc(0, rhs, x)
        rn, rd = resn(x, ax_naive), resn(x, ax_diff)
        hist.append((rn, rd))
        print(f"{label} cycle {it + 1}: residue naive-eval {rn:.4e}  diff-eval {rd:.4e}", flush=True)
        if max(rn, rd) <= 1e-7:
            break
    return x


print("|x| scale: 1/(tau sqrt N) =", 1 / (1e-6 * n_side))
run(ax_naive, "naive-cycle")
run(ax_diff, "diff-cycle ")
