"""Operator assembly on the device and device-resident solves (f1), through the C ABI.

The reference's callers build S = -igl.cotmatrix, M = igl.massmatrix, lhs = a M + b S, rhs = M V and
normalize_area on the host around every solve (demos/smoothing.py:28-47, demos/conformal_flow.py:22-59).
The device kernels are compared with the numpy restatements of those operators (gravo_mg_b200/synth.py,
gravo_mg_b200/util.py — the latter pinned against the reference's own util.py by tests/golden):
floating point, tolerance 1e-13 relative (different but fixed summation order), S bitwise symmetric.
"""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def mesh():
    import gravomg
    from gravo_mg_b200 import synth

    V, F = synth.icosphere(5)
    # an ellipsoid with ripples: half of the faces are obtuse (the mixed-Voronoi branch), the surface stays smooth
    V = V * np.array([1.0, 0.6, 1.9])
    V = V * (1.0 + 0.05 * np.sin(5 * V[:, :1]) * np.cos(4 * V[:, 1:2]))
    V, S, M, neigh = synth.mesh_operators(V, F)
    s = gravomg.MultigridSolver(V, neigh, M, tolerance=1e-8)
    s.attach_mesh(F, V)
    return V, F, S, M, s


def test_stiffness_and_mass_match_the_host_operators(mesh):
    from gravo_mg_b200 import synth

    V, F, S, M, s = mesh
    b = s.solver
    b.mesh_stiffness()
    S_host = sp.csr_matrix(S)
    S_host.sort_indices()
    pat = (sp.identity(V.shape[0], format="csr") + abs(S_host)).tocsr()  # adjacency + diagonal
    pat.sort_indices()
    got = sp.csr_matrix((b.mesh_get("stiffness"), pat.indices, pat.indptr), shape=S_host.shape)
    assert abs(got - S_host).max() <= 1e-13 * abs(S_host).max()
    assert abs(got - got.T).max() == 0.0  # bitwise symmetric
    assert np.abs(got @ np.ones(V.shape[0])).max() <= 1e-13 * abs(S_host).max()
    b.mesh_mass("voronoi")
    cots, _ = synth._face_cots(V, np.asarray(F, dtype=np.int64))
    assert (cots < 0).any()  # the obtuse branch is exercised
    assert _rel(b.mesh_get("mass"), synth.mass_voronoi(V, F).diagonal()) <= 1e-13
    b.mesh_mass("barycentric")
    assert _rel(b.mesh_get("mass"), synth.mass_barycentric(V, F).diagonal()) <= 1e-13


@pytest.mark.parametrize("kind", ["smoothing", "poisson"])
def test_device_system_solves_like_the_host_system(mesh, kind):
    V, F, S, M, s = mesh
    b = s.solver
    b.set_positions(V)
    b.mesh_stiffness()
    b.mesh_mass("voronoi")
    rng = np.random.default_rng(7)
    if kind == "smoothing":  # demos/smoothing.py:43-47
        alpha, beta, y = 1.0, 1e-3, None
        lhs, rhs = (M + 1e-3 * S).tocsr(), M @ V
    else:  # experiments/python/comparisons.py:75-96
        y = rng.standard_normal((V.shape[0], 1))
        alpha, beta = 1e-6, 1.0
        lhs, rhs = (1e-6 * M + S).tocsr(), M @ y
        b.set_option("tolerance", 1e-6)
    lhs.sort_indices()
    b.mesh_system(alpha, beta, y)
    assert _rel(b.mesh_get("lhs"), lhs.data) <= 1e-13
    assert _rel(b.mesh_get("rhs"), rhs) <= 1e-13
    b.solve_staged()
    x_dev = b.fetch()
    it_dev = int(b.solver_timing()["iterations"])
    assert b.solver_timing()["residue"] <= b.get_option("tolerance")
    assert b.transfer_timing()["h2d_bytes"] == 0.0
    x_host = s.solve(lhs, rhs)  # same pattern: only values travel
    assert b.transfer_timing()["pattern_reused"] == 1.0
    assert int(s.solver_timing["iterations"]) == it_dev
    m = M.diagonal()[:, None]
    if kind == "smoothing":
        assert np.sqrt((m * (x_dev - x_host) ** 2).sum()) <= 1e-10 * np.sqrt((m * x_host ** 2).sum())
    else:  # nearly singular: compare through the operator
        assert np.abs(lhs @ (x_dev - x_host)).max() <= 1e-6 * np.abs(rhs).max()
    b.set_option("tolerance", 1e-8)


def test_normalize_area_and_conformal_flow_match_the_host_loop(mesh):
    from gravo_mg_b200 import synth, util

    V, F, S, M, s = mesh
    b = s.solver
    b.set_option("tolerance", 1e-8)
    b.set_positions(V)
    b.mesh_stiffness()
    S_fixed = sp.csr_matrix(S)
    steps, tau = 3, 0.01
    Vt = V.copy()
    for _ in range(steps):  # demos/conformal_flow.py:54-59 on the host
        Mt = synth.mass_barycentric(Vt, F)
        lhs = (Mt + tau * S_fixed).tocsr()
        lhs.sort_indices()
        Vt = util.normalize_area(s.solve(lhs, Mt @ Vt), F)
    b.set_positions(V)
    got = s.conformal_flow(steps, tau=tau, mass="barycentric")
    assert _rel(got, Vt) <= 1e-9
    t = b.transfer_timing()
    assert t["flow_steps"] == steps and t["flow_iterations"] >= steps and t["flow_cycles_ms"] > 0
    # normalize_area alone: unit total area, centred
    assert abs(util.face_area(got, F).sum() - 1.0) <= 1e-12 and np.abs(got.mean(0)).max() <= 1e-12


def test_solve_device_takes_and_returns_cuda_tensors(mesh):
    import torch

    V, F, S, M, s = mesh
    lhs = (M + 1e-3 * S).tocsr()
    lhs.sort_indices()
    rhs = M @ V
    x_host = s.solve(lhs, rhs)
    vals = torch.from_numpy(lhs.data).cuda()
    b_dev = torch.from_numpy(np.ascontiguousarray(rhs)).cuda()
    x_dev = s.solve_device(vals, b_dev)
    assert x_dev.is_cuda and x_dev.shape == (V.shape[0], 3)
    np.testing.assert_array_equal(x_dev.cpu().numpy(), x_host)  # same values, same kernels: same bits
    with pytest.raises(ValueError):
        s.solve_device(vals[:-1], b_dev)
    with pytest.raises(TypeError):
        s.solve_device(vals.cpu(), b_dev)
