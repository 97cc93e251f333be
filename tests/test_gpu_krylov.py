"""Conjugate gradients around the V-cycle and direct_solve, through the C ABI.

Reference points: direct_solve core.cpp:74-78 -> multigrid_solver.cpp:1287-1321 (sparse LLT of the whole
system, timing keys direct_factor / direct_solve / direct_residual); plain CG multigrid_solver.cpp:1453-1477.
The checker is a sparse direct solve (scipy splu); tolerances are stated per assertion.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from oracle import oracle

pytestmark = pytest.mark.gpu


def test_pcg_needs_fewer_iterations_than_the_cycle_loop(torus_mid):
    p = torus_mid
    plain = p.new_solver(tolerance=1e-6)
    x_plain = plain.solve(p.lhs, p.rhs)
    pcg = p.new_solver(tolerance=1e-6, krylov="pcg")
    x_pcg = pcg.solve(p.lhs, p.rhs)
    it_plain, it_pcg = int(plain.solver_timing["iterations"]), int(pcg.solver_timing["iterations"])
    assert it_pcg < it_plain, (it_pcg, it_plain)
    # the reported residue is the TRUE residual b - A x, judged again by the oracle's residualCheck
    assert pcg.solver_timing["residue"] <= 1e-6
    assert oracle.residual_check(p.lhs, p.rhs, x_pcg, 2, p.m, diff=True) <= 1e-6 * (1 + 1e-6)
    hist = [r for _, r in pcg.convergence]
    assert len(hist) == it_pcg
    # same solution as the cycle loop, through the operator (the system is nearly singular)
    assert np.abs(p.lhs @ (x_pcg - x_plain)).max() <= 4e-6 * np.abs(p.rhs).max()
    x2 = pcg.solve(p.lhs, p.rhs)
    np.testing.assert_array_equal(x_pcg, x2)  # deterministic reductions
    print(f"\n[krylov] torus 90k Poisson to 1e-6: cycle loop {it_plain} cycles, PCG {it_pcg} iterations")


def test_pcg_multi_column_and_plain_cg(ico_smoothing):
    p = ico_smoothing
    xd = sla.splu(sp.csc_matrix(p.lhs)).solve(p.rhs)
    pcg = p.new_solver(tolerance=1e-10, krylov="pcg")
    x = pcg.solve(p.lhs, p.rhs)  # K = 3, one alpha / beta per column
    assert p.mnorm(x - xd) <= 1e-8 * p.mnorm(xd)
    cg = p.new_solver(tolerance=1e-8, krylov="cg", max_iter=2000)
    xc = cg.solve(p.lhs, p.rhs)
    assert cg.solver_timing["residue"] <= 1e-8 and int(cg.solver_timing["iterations"]) > int(pcg.solver_timing["iterations"])
    assert p.mnorm(xc - xd) <= 1e-6 * p.mnorm(xd)
    with pytest.raises(RuntimeError, match="1..4 right-hand sides"):
        pcg.solve(p.lhs, np.ones((p.lhs.shape[0], 5)))
    pcg.solver.set_option("krylov", 0)
    x0 = pcg.solve(p.lhs, p.rhs)  # back to the reference loop on the same handle
    assert p.mnorm(x0 - xd) <= 1e-8 * p.mnorm(xd)


def test_direct_solve_small_system_is_a_dense_factorisation(ico_small):
    p = ico_small
    lhs = (p.M + 1e-3 * p.S).tocsr()
    rhs = p.M @ p.V
    s = p.new_solver()
    x = s.direct_solve(lhs, rhs)
    xd = sla.splu(sp.csc_matrix(lhs)).solve(rhs)
    assert p.mnorm(x - xd) <= 1e-12 * p.mnorm(xd)
    t = s.solver_timing
    assert t["direct_factor"] > 0 and t["direct_solve"] > 0 and t["direct_residual"] <= 1e-13
    x = s.direct_solve(p.lhs, p.rhs, pardiso=True)  # Poisson, nearly singular: backward error
    floor = np.finfo(float).eps * (abs(p.lhs) @ np.abs(x)).max()  # x carries a constant ~4e4: rounding floor of A x
    assert np.abs(p.lhs @ x - p.rhs).max() <= 50 * floor
    assert s.solve(lhs, rhs).shape == rhs.shape  # the V-cycle path of the same handle is untouched


def test_direct_solve_large_system_runs_cg_to_the_rounding_floor(torus_mid):
    p = torus_mid
    lhs = (p.M + 1e-3 * p.S).tocsr()
    rhs = p.M @ p.V
    s = p.new_solver(tolerance=1e-4)
    x = s.direct_solve(lhs, rhs)
    xd = sla.splu(sp.csc_matrix(lhs)).solve(rhs)
    assert p.mnorm(x - xd) <= 1e-11 * p.mnorm(xd)
    assert s.solver_timing["direct_residual"] <= 1e-12
    assert s.solver.get_option("krylov") == 0 and s.solver.get_option("tolerance") == 1e-4  # settings restored


def test_solve_with_page_locked_caller_buffers(ico10k):
    """lhs.data, rhs and ``out`` in page-locked memory (views of torch pinned tensors): the copy engine reads and
    writes them directly (host_xfer.h); same bits as the staged pageable path."""
    import torch

    p = ico10k
    s = p.new_solver(tolerance=1e-6)
    x_ref = s.solve(p.lhs, p.rhs)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        v = t.numpy()
        v[...] = a
        return v, t

    data, k1 = pinned(p.lhs.data)
    rhs, k2 = pinned(p.rhs)
    out, k3 = pinned(np.zeros_like(p.rhs))
    lhs = p.lhs.copy()
    lhs.data = data
    x = s.solve(lhs, rhs, out=out)
    assert x is out
    np.testing.assert_array_equal(x, x_ref)
    assert s.solver.transfer_timing()["pattern_reused"] == 1.0
    with pytest.raises(ValueError):
        s.solve(lhs, rhs, out=np.zeros((3, 1)))
