"""CUDA path against the CPU oracle, through the C ABI (ctypes -> libgravomg_b200.so).

Parity levels:
  P1 op level     each kernel vs the oracle's same operator on identical inputs. The staged
                  (TMA) kernels sum a row's products in CSR order with no fused multiply-add, so
                  fp64 results are BIT-EXACT; the direct kernels reduce across lanes: <= 1e-13.
  P2 cycle level  one device V-cycle vs the oracle's Jacobi V-cycle (same omega, U, operators).
  P3 solver level device solve vs the oracle's lexicographic Gauss-Seidel solve (the reference
                  algorithm): both reach the tolerance; cycle counts and the residual-norm ratio
                  are recorded; both agree with a sparse direct solve.
  P4 invariants   Galerkin symmetry, monotone residual history, independence of launch mode.

On the finest level the device evaluates A x in the cancellation-free form sum_{j != i} A_ij (x_j - x_i)
+ s_i x_i (option diff_form, on by default). The oracle's Jacobi variant has the same
form (``diff=True`` / ``row_product="diff"``); the Gauss-Seidel reference path never uses it.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from oracle import oracle

pytestmark = pytest.mark.gpu

EPS = np.finfo(float).eps
OMEGA = 2.0 / 3.0


def _staged(p, K, lhs=None, rhs=None, **kw):
    """A fresh solver with a system staged on the device; returns (binding object, lhs, rhs)."""
    s = p.new_solver(**kw).solver
    lhs = p.lhs if lhs is None else lhs
    if rhs is None:
        rng = np.random.default_rng(11)
        rhs = rng.standard_normal((lhs.shape[0], K))
    s.stage(lhs, rhs)
    return s, lhs, rhs


def _weights(binding, n_levels):
    """{level: (pre, post)} Jacobi dampings the device used (after a solve or a level_op)."""
    out = {}
    for k in range(n_levels):
        _, pre, post = binding.smoother_weights(k)
        out[k] = (pre, post)
    return out


def _jacobi_oracle(p, solver, **kw):
    """The oracle running the device's own cycle: Jacobi sweeps with the device's dampings."""
    diff = bool(solver.solver.get_option("diff_form"))
    return oracle.OracleSolver(p.M, p.U, smoother="jacobi", weights=_weights(solver.solver, len(p.U)),
                               row_product="diff" if diff else "plain", **kw)


def _device_levels(s, n_levels):
    return [s.level_matrix(k) for k in range(n_levels + 1)]


def _rows(A):
    """The oracle's smoothers walk column k of a CSC matrix as row k (multigrid_solver.cpp:1202);
    Galerkin operators are symmetric only to rounding, so hand them over transposed to compare
    the same stored entries the device reads."""
    return sp.csr_matrix(A).T


def _rel_inf(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ------------------------------------------------------------------------------ P1: operators
@pytest.mark.parametrize("mode", ["staged-exact", "staged-auto", "direct", "staged-exact-plain"])
@pytest.mark.parametrize("K", [1, 3])
@pytest.mark.parametrize("smoother", ["chebyshev", "jacobi"])
def test_p1_operators_match_the_oracle(ico_small, mode, K, smoother):
    p = ico_small
    s, lhs, _ = _staged(p, K, smoother=smoother)
    diff_form = not mode.endswith("-plain")
    s.set_option("diff_form", diff_form)
    mode = mode.replace("-plain", "")
    s.set_option("kernel_path", 1 if mode == "direct" else 0)
    s.set_option("lanes", 1 if mode == "staged-exact" else 0)
    s.stage(lhs, np.zeros((lhs.shape[0], K)))  # re-plan with the chosen kernel path
    L = len(p.U)
    rng = np.random.default_rng(5)
    s.level_op("residual", 0, np.zeros((lhs.shape[0], K)), np.zeros((lhs.shape[0], K)))  # triggers the reduction
    A = _device_levels(s, L)
    W = _weights(s, L)
    exact = mode == "staged-exact"

    def check(got, want):
        if exact:
            np.testing.assert_array_equal(got, want)
        else:
            assert _rel_inf(got, want) <= 1e-13

    for k in range(L + 1):
        n = A[k].shape[0]
        x = rng.standard_normal((n, K))
        b = rng.standard_normal((n, K))
        diff = diff_form and k == 0 and L > 0  # the finest level uses the cancellation-free row product
        check(s.level_op("residual", k, x, b), oracle.residual(_rows(A[k]) if diff else A[k], b, x, diff=diff))
        if k < L:
            for sweeps in (1, 2, 3):  # level_op cycles through the pre-smoothing dampings
                om = [W[k][0][i % len(W[k][0])] for i in range(sweeps)]
                check(s.level_op("jacobi", k, x, b, sweeps=sweeps), oracle.jacobi(_rows(A[k]), b, x, sweeps, om, diff=diff))
            U = p.U[k]
            e = rng.standard_normal((U.shape[1], K))
            check(s.level_op("restrict", k, x), oracle.restrict(U, x))
            check(s.level_op("prolong_add", k, e, x), oracle.prolong_add(U, e, x))


@pytest.mark.parametrize("path", [0, 1])
def test_p1_operators_on_a_larger_mesh(torus_mid, path):
    p = torus_mid
    s, lhs, _ = _staged(p, 1)
    s.set_option("kernel_path", path)
    s.set_option("lanes", 1)
    s.stage(lhs, np.zeros((lhs.shape[0], 1)))
    rng = np.random.default_rng(6)
    n = lhs.shape[0]
    x = rng.standard_normal((n, 1))
    b = rng.standard_normal((n, 1))
    got = s.level_op("jacobi", 0, x, b, sweeps=2)
    want = oracle.jacobi(lhs, b, x, 2, _weights(s, 1)[0][0], diff=True)
    if path == 0:
        np.testing.assert_array_equal(got, want)
    else:
        assert _rel_inf(got, want) <= 1e-13
    U = p.U[0]
    got = s.level_op("restrict", 0, x)
    want = oracle.restrict(U, x)
    if path == 0:
        np.testing.assert_array_equal(got, want)
    else:
        assert _rel_inf(got, want) <= 1e-13


def test_p1_galerkin_operators(ico_small):
    """Abar[k+1] = U^T Abar[k] U (multigrid_solver.cpp:1387-1392) on the device vs the oracle."""
    p = ico_small
    s, lhs, _ = _staged(p, 1)
    s.level_op("residual", 0, np.zeros((lhs.shape[0], 1)), np.zeros((lhs.shape[0], 1)))
    dev = _device_levels(s, len(p.U))
    o = oracle.OracleSolver(p.M, p.U)
    o.setup(lhs)
    ref = o.level_matrices()
    assert abs(dev[0] - lhs).max() == 0
    for got, want in zip(dev[1:], ref):
        assert got.shape == want.shape
        assert abs(got - want).max() <= 1e-13 * abs(want).max()
        assert abs(got - got.T).max() <= 1e-12 * abs(got).max()  # P4
    info = s.level_info()
    assert [lv["rows"] for lv in info] == [lhs.shape[0]] + [u.shape[1] for u in p.U]
    assert [lv["nnz_u"] for lv in info[:-1]] == [u.nnz for u in p.U]


def test_p1_galerkin_product_plans_are_bit_identical(ico10k, torus_mid):
    """The per-pattern index-pair lists (build_spgemm_plan) replace the search of the per-solve
    Galerkin kernels; the products are summed in the same order, so every level operator has the
    same bits either way."""
    for p in (ico10k, torus_mid):
        mats = []
        for plan in (1, 0):
            s = p.new_solver().solver
            s.set_option("spgemm_plan", plan)
            s.stage(p.lhs, p.rhs)
            s.solve_staged()
            mats.append(_device_levels(s, len(p.U)))
            if plan:
                assert s.transfer_timing()["galerkin_plan_pairs"] > 0
        for a, b in zip(*mats):
            np.testing.assert_array_equal(a.indptr, b.indptr)
            np.testing.assert_array_equal(a.indices, b.indices)
            np.testing.assert_array_equal(a.data, b.data)


def test_p1_coarse_solve(ico_small):
    """Dense Cholesky + explicit inverse on the device vs the oracle's sparse LDL^T."""
    p = ico_small
    rng = np.random.default_rng(8)
    L = len(p.U)
    # well conditioned: forward error
    lhs = (p.M + 1e-3 * p.S).tocsr()
    s, _, _ = _staged(p, 2, lhs=lhs)
    o = oracle.OracleSolver(p.M, p.U)
    o.setup(lhs)
    nc = p.U[-1].shape[1]
    b = rng.standard_normal((nc, 2))
    got = s.level_op("coarse", L, b)
    want = o.coarse_solve(b)
    assert np.linalg.norm(got - want) <= 1e-11 * np.linalg.norm(want)
    # Poisson (nearly singular): backward error
    s, lhs, _ = _staged(p, 2)
    got = s.level_op("coarse", L, b)
    Ac = s.level_matrix(L).toarray()
    backward = np.linalg.norm(Ac @ got - b) / (np.linalg.norm(Ac, 2) * np.linalg.norm(got) + np.linalg.norm(b))
    assert backward <= 1e-13


@pytest.mark.parametrize("fixture,expect_rows", [("ico_small", 54), ("ico10k", 1800), ("torus_mid", 2700)])
def test_p1_coarse_factor_dataflow_kernel(request, fixture, expect_rows):
    """The coarse factor as one dataflow kernel (tiles + flags, dense_factor.cuh) against the
    kernel-per-phase version and against numpy on the device's own coarse operator: one tile
    (54 rows), ~29 and ~43 tile rows (more tasks than resident CTAs)."""
    p = request.getfixturevalue(fixture)
    L = len(p.U)
    nc = p.U[-1].shape[1]
    assert 0.5 * expect_rows <= nc <= 1.6 * expect_rows
    rng = np.random.default_rng(5)
    b = rng.standard_normal((nc, 3))
    lhs = (p.M + 1e-3 * p.S).tocsr()
    got = {}
    for flag in (1, 0):
        s, _, _ = _staged(p, 3, lhs=lhs)
        s.set_option("coarse_dataflow", flag)
        got[flag] = s.level_op("coarse", L, b)
        got[flag, "again"] = s.level_op("coarse", L, b)   # second factorisation: next flag epoch
        Ac = s.level_matrix(L).toarray()
    want = np.linalg.solve(Ac, b)
    for flag in (1, 0):
        assert np.linalg.norm(got[flag] - want) <= 1e-11 * np.linalg.norm(want), flag
        np.testing.assert_array_equal(got[flag], got[flag, "again"])
    # Poisson (nearly singular operator): backward error of the dataflow factor
    s, _, _ = _staged(p, 3)
    x = s.level_op("coarse", L, b)
    Ac = s.level_matrix(L).toarray()
    backward = np.linalg.norm(Ac @ x - b) / (np.linalg.norm(Ac, 2) * np.linalg.norm(x) + np.linalg.norm(b))
    assert backward <= 1e-13


def test_smoother_weights_are_chebyshev_roots_on_the_gershgorin_band(ico_small):
    p = ico_small
    s, lhs, _ = _staged(p, 1, pre_iters=3, post_iters=2, cheb_alpha=8.0)
    s.level_op("residual", 0, np.zeros((lhs.shape[0], 1)), np.zeros((lhs.shape[0], 1)))
    A = _device_levels(s, len(p.U))
    for k in range(len(p.U)):
        rho, pre, post = s.smoother_weights(k)
        assert rho == pytest.approx(oracle.gershgorin_rho(A[k]), rel=1e-14)
        np.testing.assert_allclose(pre, oracle.chebyshev_weights(rho, 8.0, 3), rtol=1e-13)
        np.testing.assert_allclose(post, oracle.chebyshev_weights(rho, 8.0, 2)[::-1], rtol=1e-13)
        assert (pre > 0).all() and (pre * rho < 8.0 + 1e-9).all() and (pre * rho >= 1.0 - 1e-12).all()
    s2, _, _ = _staged(p, 1, smoother="jacobi", omega=0.7)
    s2.level_op("residual", 0, np.zeros((lhs.shape[0], 1)), np.zeros((lhs.shape[0], 1)))
    _, pre, post = s2.smoother_weights(0)
    np.testing.assert_array_equal(np.concatenate([pre, post]), 0.7)


def test_chebyshev_smoother_needs_fewer_cycles_than_fixed_jacobi(torus_mid):
    """Cycle counts to 1e-6 next to the reference algorithm's (Gauss-Seidel) count."""
    p = torus_mid
    counts = {}
    for smoother in ("jacobi", "chebyshev"):
        s = p.new_solver(tolerance=1e-6, smoother=smoother)
        s.solve(p.lhs, p.rhs)
        counts[smoother] = int(s.solver_timing["iterations"])
    og = oracle.OracleSolver(p.M, p.U, tolerance=1e-6, smoother="gs")
    og.solve(p.lhs, p.rhs)
    counts["reference gs"] = int(og.solver_timing["iterations"])
    print("\ncycles to 1e-6:", counts)
    assert counts["chebyshev"] < counts["jacobi"]
    assert counts["chebyshev"] <= 2 * counts["reference gs"]


# ------------------------------------------------------------------------------ P2: one cycle
@pytest.mark.parametrize("K", [1, 3])
@pytest.mark.parametrize("smoother", ["chebyshev", "jacobi"])
def test_p2_one_vcycle_matches_the_oracle_jacobi_cycle(ico_small, K, smoother):
    p = ico_small
    lhs = (p.M + 1e-3 * p.S).tocsr()
    rhs = (p.M @ p.V)[:, :K]
    solver = p.new_solver(max_iter=1, smoother=smoother)
    x = solver.solve(lhs, rhs)
    o = _jacobi_oracle(p, solver, max_iter=1)
    want = o.solve(lhs, rhs)
    assert np.linalg.norm(x - want) <= 1e-12 * np.linalg.norm(want)
    assert solver.solver_timing["iterations"] == 1
    assert solver.solver_timing["residue"] == pytest.approx(o.solver_timing["residue"], rel=1e-9)


@pytest.mark.parametrize("K", [1, 3])
@pytest.mark.parametrize("cycle_type", [1, 2])
def test_p2_f_and_w_cycles_match_the_oracle(ico_small, cycle_type, K):
    """cycle_type 1 / 2 (multigrid_solver.cpp:1091-1192, with the coarsest-level test fixed): one device
    cycle and a whole solve against the oracle running the same cycle with the device's sweeps."""
    p = ico_small
    lhs = (p.M + 1e-3 * p.S).tocsr()
    rhs = (p.M @ p.V)[:, :K]
    one = p.new_solver(max_iter=1, cycle_type=cycle_type)
    x = one.solve(lhs, rhs)
    o = _jacobi_oracle(p, one, max_iter=1, cycle_type=cycle_type)
    want = o.solve(lhs, rhs)
    assert np.linalg.norm(x - want) <= 1e-12 * np.linalg.norm(want)
    v = p.new_solver(max_iter=1)
    assert np.linalg.norm(x - v.solve(lhs, rhs)) > 1e-9 * np.linalg.norm(want)  # not a V-cycle
    full = p.new_solver(tolerance=1e-8, cycle_type=cycle_type)
    xf = full.solve(lhs, rhs)
    of = _jacobi_oracle(p, full, tolerance=1e-8, cycle_type=cycle_type)
    of.solve(lhs, rhs)
    vf = p.new_solver(tolerance=1e-8)
    vf.solve(lhs, rhs)
    assert full.solver_timing["iterations"] == of.solver_timing["iterations"] <= vf.solver_timing["iterations"]
    assert oracle.residual_check(lhs, rhs, xf, 2, p.m) <= 1e-8
    # launch modes agree
    host = p.new_solver(tolerance=1e-8, cycle_type=cycle_type)
    host.solver.set_option("loop_mode", 0)
    np.testing.assert_array_equal(host.solve(lhs, rhs), xf)


def test_p2_one_vcycle_poisson(ico_small):
    p = ico_small
    solver = p.new_solver(max_iter=1)
    x = solver.solve(p.lhs, p.rhs)
    o = _jacobi_oracle(p, solver, max_iter=1)
    want = o.solve(p.lhs, p.rhs)
    # rounding floor of A x for |x| ~ mean(rhs)/tau (see tests/test_oracle.py)
    bound = 4 * EPS * abs(p.lhs).sum(0).max() * np.linalg.norm(want)
    assert np.linalg.norm(p.lhs @ (x - want)) <= bound


# ------------------------------------------------------------------------------ P3: whole solve
@pytest.mark.parametrize("fixture,tol", [("ico10k", 1e-4), ("ico10k", 1e-6), ("torus_mid", 1e-6), ("cloud_mid", 1e-6)])
def test_p3_solve_reaches_the_tolerance_like_the_reference(request, fixture, tol):
    p = request.getfixturevalue(fixture)
    solver = p.new_solver(tolerance=tol)
    x = solver.solve(p.lhs, p.rhs)
    t = solver.solver_timing
    hist = [r for _, r in solver.convergence]

    # the same cycle on the CPU (Jacobi variant): same residual history down to the rounding
    # floor of forming A x (|x| ~ mean(rhs)/tau for the Poisson system, see tests/test_oracle.py)
    oj = _jacobi_oracle(p, solver, tolerance=tol)
    xj = oj.solve(p.lhs, p.rhs)
    hist_j = [r for _, r in oj.convergence]
    assert int(t["iterations"]) == len(hist)
    assert abs(len(hist) - len(hist_j)) <= 1
    m = min(len(hist), len(hist_j))
    floor = 10 * EPS * abs(p.lhs).sum(0).max() * np.linalg.norm(xj) / np.linalg.norm(p.rhs)
    np.testing.assert_allclose(hist[:m], hist_j[:m], rtol=1e-6, atol=floor)

    # the reference algorithm (lexicographic Gauss-Seidel): both terminate below the tolerance
    og = oracle.OracleSolver(p.M, p.U, tolerance=tol, smoother="gs")
    x_gs = og.solve(p.lhs, p.rhs)
    rhs = p.rhs
    res_gpu = oracle.residual_check(p.lhs, rhs, x, 2, p.m)  # judged by the oracle's norm
    res_ref = og.solver_timing["residue"]
    assert res_gpu <= tol and res_ref <= tol and abs(t["residue"] - res_gpu) <= 1e-6 * res_gpu + floor
    assert all(a > b for a, b in zip(hist, hist[1:]))  # P4: monotone
    print(f"\n[{fixture} tol={tol:g}] cycles gpu(chebyshev-jacobi)={len(hist)} ref(gs)={int(og.solver_timing['iterations'])} "
          f"residual-norm ratio gpu/ref={res_gpu / res_ref:.3f}")

    # both agree with a sparse direct solve in the norm the residual controls
    xd = sla.splu(sp.csc_matrix(p.lhs)).solve(rhs)
    err_gpu = p.mnorm(p.lhs @ (x - xd)) / p.mnorm(rhs)
    err_ref = p.mnorm(p.lhs @ (x_gs - xd)) / p.mnorm(rhs)
    assert err_gpu <= 1.01 * tol and err_ref <= 1.01 * tol


def test_p3_smoothing_system_K3(ico_smoothing):
    """demos/smoothing.py:43-50: (M + 1e-3 S) x = M V with three right-hand sides."""
    p = ico_smoothing
    solver = p.new_solver(tolerance=1e-8)
    x = solver.solve(p.lhs, p.rhs)
    assert x.shape == (p.lhs.shape[0], 3)
    xd = sla.splu(sp.csc_matrix(p.lhs)).solve(p.rhs)
    assert p.mnorm(x - xd) <= 1e-6 * p.mnorm(xd)
    assert solver.residual(p.lhs, p.rhs, x) <= 1e-8
    og = oracle.OracleSolver(p.M, p.U, tolerance=1e-8)
    x_gs = og.solve(p.lhs, p.rhs)
    assert p.mnorm(x - x_gs) <= 1e-6 * p.mnorm(x_gs)


@pytest.mark.parametrize("type_", [0, 1, 2, 3])
@pytest.mark.parametrize("K", [1, 3, 6])
def test_residual_types(ico_small, type_, K):
    """residualCheck (multigrid_solver.cpp:1228-1277), all four norms, max over columns."""
    p = ico_small
    rng = np.random.default_rng(9)
    n = p.lhs.shape[0]
    x = rng.standard_normal((n, K))
    b = rng.standard_normal((n, K)) * np.linspace(0.5, 2.0, K)
    got = p.solver.residual(p.lhs, b, x, type_)
    want = oracle.residual_check(p.lhs, b, x, type_, p.m)
    assert got == pytest.approx(want, rel=1e-12)
    assert got == pytest.approx(oracle.residual_check(p.lhs, b, x, type_, p.m, diff=True), rel=1e-13)
    p.solver.solver.set_option("diff_form", 0)
    try:
        assert p.solver.residual(p.lhs, b, x, type_) == pytest.approx(want, rel=1e-13)
    finally:
        p.solver.solver.set_option("diff_form", 1)


# ------------------------------------------------------------------------------ P4 / behaviour
def test_launch_modes_give_identical_results(ico10k):
    """Host loop over graph replays, one device-side while-graph, and plain launches."""
    p = ico10k
    results = []
    for use_graph, loop_mode in [(1, 0), (1, 1), (0, 0)]:
        solver = p.new_solver(tolerance=1e-6)
        solver.solver.set_option("use_graph", use_graph)
        solver.solver.set_option("loop_mode", loop_mode)
        x = solver.solve(p.lhs, p.rhs)
        results.append((x, solver.solver_timing["iterations"], solver.solver_timing["residue"]))
    for x, it, res in results[1:]:
        np.testing.assert_array_equal(x, results[0][0])
        assert it == results[0][1] and res == results[0][2]


def test_fused_coarse_tail_matches_per_operator_kernels(ico10k, torus_mid):
    """Option tail_rows: levels below the threshold run inside one persistent kernel with grid
    barriers; same arithmetic up to the lane split of the row sums."""
    for p in (ico10k, torus_mid):
        xs = []
        for tail in (0, 100000):
            solver = p.new_solver(tolerance=1e-6)
            solver.solver.set_option("tail_rows", tail)
            xs.append(solver.solve(p.lhs, p.rhs))
            iters = solver.solver_timing["iterations"]
        bound = 50 * EPS * abs(p.lhs).sum(0).max() * np.linalg.norm(xs[0])
        assert np.linalg.norm(p.lhs @ (xs[0] - xs[1])) <= bound
        o = _jacobi_oracle(p, solver, tolerance=1e-6)
        o.solve(p.lhs, p.rhs)
        assert abs(int(o.solver_timing["iterations"]) - int(iters)) <= 1


@pytest.mark.parametrize("kind", ["poisson", "smoothing_K3", "float32"])
def test_cluster_tail_matches_per_operator_kernels(torus_mid, kind):
    """Option cluster_tail_rows: the levels from the first one with <= that many rows down run as ONE thread-block
    cluster (operators staged in shared memory by bulk copies, hardware cluster barriers between operators,
    cluster_tail.cuh). Same arithmetic as the per-operator kernels up to the lane split of the row sums."""
    p = torus_mid
    lhs, rhs = (p.lhs, p.rhs) if kind == "poisson" else ((p.M + 1e-3 * p.S).tocsr(), p.M @ p.V)
    kw = dict(lower_bound=50, tolerance=1e-6 if kind != "float32" else 1e-4)
    if kind == "float32":
        kw["dtype"] = "float32"
    xs, iters, launches = [], [], []
    for rows in (0, 8192, 30000):
        solver = p.new_solver(**kw)
        solver.solver.set_option("cluster_tail_rows", rows)
        solver.solver.set_option("trace", 1)
        xs.append(solver.solve(lhs, rhs))
        iters.append(int(solver.solver_timing["iterations"]))
        _, tags = solver.solver.trace()
        launches.append(int((tags == 105).sum()))
        sizes = [lv["rows"] for lv in solver.solver.level_info()]
    assert len(sizes) >= 5 and sizes[2] <= 8192 < sizes[1] <= 30000
    # 0: off. 8192: levels 2.. in the cluster kernel, once per cycle. 30000: level 1 would join, but its operators
    # do not fit the shared memory of one cluster -> the per-operator kernels stay (same bits as off)
    assert launches[0] == 0 and launches[1] == iters[1] and launches[2] == 0
    np.testing.assert_array_equal(xs[0], xs[2])
    assert abs(iters[0] - iters[1]) <= 1
    if kind == "float32":  # fp32 levels correct an fp64 iterate: both runs end below the tolerance, judged by the oracle
        assert oracle.residual_check(lhs, rhs, xs[1], 2, p.m) <= 1e-4 * (1 + 1e-3)
        assert p.mnorm(xs[0] - xs[1]) <= 2e-4 * p.mnorm(xs[0])
    else:
        bound = 50 * EPS * abs(lhs).sum(0).max() * np.linalg.norm(xs[0])
        assert np.linalg.norm(lhs @ (xs[0] - xs[1])) <= bound


def test_kernel_paths_agree(ico10k):
    p = ico10k
    xs = []
    for path in (0, 1):
        solver = p.new_solver(tolerance=1e-6)
        solver.solver.set_option("kernel_path", path)
        xs.append(solver.solve(p.lhs, p.rhs))
    bound = 50 * EPS * abs(p.lhs).sum(0).max() * np.linalg.norm(xs[0])
    assert np.linalg.norm(p.lhs @ (xs[0] - xs[1])) <= bound


def test_at_least_one_cycle_max_iter_and_initial_guess(ico_small):
    p = ico_small
    s = p.new_solver(tolerance=1e30)
    s.solve(p.lhs, p.rhs)
    assert s.solver_timing["iterations"] == 1  # do { } while
    s = p.new_solver(tolerance=0.0, max_iter=3)
    x3 = s.solve(p.lhs, p.rhs)
    assert s.solver_timing["iterations"] == 3 and len(s.convergence) == 3
    # x0 = rhs (core.cpp:69): three oracle cycles from x0 = rhs land on the same iterate
    o = _jacobi_oracle(p, s, tolerance=0.0, max_iter=3)
    want = o.solve(p.lhs, p.rhs)
    bound = 8 * EPS * abs(p.lhs).sum(0).max() * np.linalg.norm(want)
    assert np.linalg.norm(p.lhs @ (x3 - want)) <= bound


@pytest.mark.parametrize("K", [1, 2, 5, 9])
def test_any_number_of_right_hand_sides(ico_small, K):
    """K = 2 and K > 3 are out-of-bounds reads upstream; here any K >= 1 works."""
    p = ico_small
    lhs = (p.M + 1e-3 * p.S).tocsr()
    rng = np.random.default_rng(K)
    rhs = p.M @ rng.standard_normal((lhs.shape[0], K))
    s = p.new_solver(tolerance=1e-9)
    x = s.solve(lhs, rhs if K > 1 else rhs[:, 0])  # 1-D rhs loads as N x 1
    assert x.shape == (lhs.shape[0], K)
    xd = sla.splu(sp.csc_matrix(lhs)).solve(rhs)
    assert np.linalg.norm(x - xd) <= 1e-7 * np.linalg.norm(xd)
    # columns are independent (multigrid_solver.cpp:1213): column 0 alone gives the same iterate
    s1 = p.new_solver(tolerance=0.0, max_iter=int(s.solver_timing["iterations"]))
    x0 = s1.solve(lhs, rhs[:, 0])
    np.testing.assert_array_equal(x0[:, 0], x[:, 0])


def test_float32_smoother_levels(ico10k):
    p = ico10k
    lhs = (p.M + 1e-3 * p.S).tocsr()
    rhs = p.M @ p.V
    s32 = p.new_solver(dtype="float32", tolerance=1e-4)
    x32 = s32.solve(lhs, rhs)
    assert s32.solver_timing["residue"] <= 1e-4
    assert oracle.residual_check(lhs, rhs, x32, 2, p.m) <= 2e-4
    s64 = p.new_solver(tolerance=1e-4)
    x64 = s64.solve(lhs, rhs)
    assert p.mnorm(x32 - x64) <= 1e-4 * p.mnorm(x64)
    # op level: fp32 Jacobi sweep vs the fp64 oracle, 2e-6 relative
    b = s32.solver
    b.stage(lhs, rhs)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(rhs.shape)
    got = b.level_op("jacobi", 0, x, rhs, sweeps=1)
    want = oracle.jacobi(lhs, rhs, x, 1, b.smoother_weights(0)[1][:1])
    assert _rel_inf(got, want) <= 2e-6


def test_no_hierarchy_falls_back_to_the_direct_solve():
    """N <= lower_bound: upstream indexes U[0] out of range (undefined); here the whole system
    goes to the dense direct solver."""
    import gravomg
    from gravo_mg_b200 import synth

    V, F = synth.icosphere(3)
    V, S, M, neigh = synth.mesh_operators(V, F)
    lhs, rhs = synth.smoothing_system(V, S, M)
    s = gravomg.MultigridSolver(V, neigh, M)
    assert s.prolongation_matrices == []
    x = s.solve(lhs, rhs)
    xd = sla.splu(sp.csc_matrix(lhs)).solve(rhs)
    assert np.linalg.norm(x - xd) <= 1e-10 * np.linalg.norm(xd)
    assert s.solver_timing["iterations"] == 1


def test_injected_hierarchy_and_repeated_solves(ico10k):
    """set_prolongation_matrices (core.cpp:86-88) and the repeated-solve pattern of
    demos/conformal_flow.py:49-58 (same hierarchy, new values each step)."""
    p = ico10k
    a = p.new_solver(build_hierarchy=False, tolerance=1e-7)
    a.set_prolongation_matrices(p.U)
    b = p.new_solver(tolerance=1e-7)
    xa = a.solve(p.lhs, p.rhs)
    xb = b.solve(p.lhs, p.rhs)
    np.testing.assert_array_equal(xa, xb)
    lhs2 = (p.M + 0.01 * p.S).tocsr()
    rhs2 = p.M @ p.V
    x2 = b.solve(lhs2, rhs2)          # same pattern, new values, K changes 1 -> 3
    fresh = p.new_solver(tolerance=1e-7)
    np.testing.assert_array_equal(x2, fresh.solve(lhs2, rhs2))
    x1_again = b.solve(p.lhs, p.rhs)  # and back
    np.testing.assert_array_equal(x1_again, xb)


def test_host_transfer_paths_and_pattern_change(torus_mid):
    """gmg_solve with caller-owned host arrays: the pinned-chunk worker threads (host_xfer.h) and
    the plain pageable copies move the same bytes; a changed sparsity pattern of the same shape
    is detected by the threaded comparison and re-staged."""
    p = torus_mid
    rng = np.random.default_rng(3)
    rhs3 = rng.standard_normal((p.lhs.shape[0], 3))
    ref = None
    for threads in (0, 1, 3, 8):
        solver = p.new_solver(tolerance=1e-6)
        solver.solver.set_option("xfer_threads", threads)
        x1 = solver.solve(p.lhs, p.rhs)
        x1b = solver.solve(p.lhs, p.rhs)                 # second call: speculative upload + pattern compare
        tt = solver.solver.transfer_timing()
        assert tt["pattern_reused"] == 1.0 and tt["transfer_threads"] == threads
        assert tt["h2d_bytes"] == p.lhs.data.nbytes + p.rhs.nbytes and tt["d2h_bytes"] == x1.nbytes
        x3 = solver.solve(p.lhs, rhs3)                   # K 1 -> 3 on the staged pattern
        # same shape and nnz, different pattern: swap two column indices inside one row and keep the
        # matrix the same operator by swapping the values too -> different arrays, same solution
        lhs2 = p.lhs.copy()
        r = 17
        a, b = lhs2.indptr[r], lhs2.indptr[r] + 1
        lhs2.indices[[a, b]] = lhs2.indices[[b, a]]
        lhs2.data[[a, b]] = lhs2.data[[b, a]]
        lhs2.has_sorted_indices = False
        x2 = solver.solve(lhs2, p.rhs)
        assert solver.solver.transfer_timing()["pattern_reused"] == 0.0
        np.testing.assert_array_equal(x1, x1b)
        # same operator, one row summed in another order: equal up to rounding of that row
        assert np.abs(x2 - x1).max() <= 1e-9 * np.abs(x1).max()
        if ref is None:
            ref = (x1, x3)
        np.testing.assert_array_equal(x1, ref[0])
        np.testing.assert_array_equal(x3, ref[1])


def test_timing_maps_and_csv_writers(ico_small, tmp_path):
    p = ico_small
    s = p.new_solver(tolerance=1e-6)
    s.solve(p.lhs, p.rhs)
    t = s.solver_timing
    assert sorted(t) == ["coarsest_solve", "cycles", "iterations", "reduction", "residue", "solver_total"]
    assert t["solver_total"] >= t["cycles"] > 0 and t["reduction"] > 0 and t["coarsest_solve"] > 0
    conv = s.convergence
    assert len(conv) == int(t["iterations"]) and conv[-1][1] == t["residue"]
    assert all(b[0] > a[0] for a, b in zip(conv, conv[1:]))  # elapsed ms grows
    f = tmp_path / "solver.csv"
    s.write_solver_timing("poisson", str(f), True)
    lines = f.read_text().splitlines()
    assert lines[0] == "experiment,coarsest_solve,cycles,iterations,reduction,residue,solver_total"
    assert lines[1].split(",")[0] == "poisson" and len(lines[1].split(",")) == 7
    g = tmp_path / "conv.csv"
    s.write_convergence(str(g))
    lines = g.read_text().splitlines()
    assert lines[0] == "time,residue" and len(lines) == 1 + len(conv)
    assert s.solver.last_launch_count() > 0


def test_error_behaviour(ico_small):
    p = ico_small
    s = p.new_solver()
    with pytest.raises(ValueError):
        s.solve(p.lhs, np.zeros((p.lhs.shape[0] + 1, 1)))
    bad = p.lhs.copy().tolil()
    bad[5, 5] = -1.0
    with pytest.raises(RuntimeError, match="diagonal"):
        s.solve(bad.tocsr(), p.rhs)
    other = sp.identity(p.lhs.shape[0] - 1, format="csr")
    with pytest.raises(ValueError):
        s.solver.solve(other, p.rhs)
    # an all-zero right-hand side: 0/0 residual, the loop simply ends (as upstream)
    x = s.solve(p.lhs, np.zeros((p.lhs.shape[0], 1)))
    assert (x == 0).all() and s.solver_timing["iterations"] == 1 and np.isnan(s.solver_timing["residue"])
    # the solver stays usable after an error
    x = s.solve(p.lhs, p.rhs)
    assert s.solver_timing["residue"] <= 1e-4
    # a diverging smoother is reported, not returned
    d = p.new_solver(smoother="jacobi", omega=2.5, max_iter=100)
    with pytest.raises(RuntimeError, match="diverged|non-finite"):
        d.solve(p.lhs, p.rhs)


def test_unsorted_and_duplicate_csr_input(ico_small):
    """scipy hands over whatever the caller built: unsorted column indices must work."""
    p = ico_small
    lhs = p.lhs.copy()
    rng = np.random.default_rng(2)
    for r in range(lhs.shape[0]):  # shuffle entries inside every row
        a, b = lhs.indptr[r], lhs.indptr[r + 1]
        perm = rng.permutation(b - a)
        lhs.indices[a:b] = lhs.indices[a:b][perm]
        lhs.data[a:b] = lhs.data[a:b][perm]
    lhs.has_sorted_indices = False
    s = p.new_solver(tolerance=1e-6)  # 1e-8 is below the rounding floor of this Poisson system
    x = s.solve(lhs, p.rhs)
    assert oracle.residual_check(p.lhs, p.rhs, x, 2, p.m) <= 1e-6
