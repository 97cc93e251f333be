"""Host-side hierarchy builder (prerequisite of the V-cycle path): structural invariants the
reference's construction implies (multigrid_solver.cpp:62-469)."""
import numpy as np
import pytest
import scipy.sparse as sp

import gravomg
from gravo_mg_b200 import synth


def _check_prolongations(U, n0):
    rows = n0
    for u in U:
        assert sp.issparse(u) and u.format == "csc"  # Eigen hands back column-major matrices
        assert u.shape[0] == rows and u.shape[1] < rows
        r = u.tocsr()
        per_row = np.diff(r.indptr)
        assert per_row.min() >= 1 and per_row.max() <= 3
        np.testing.assert_allclose(np.asarray(r.sum(1)).ravel(), 1.0, rtol=0, atol=1e-12)
        assert np.diff(u.indptr).min() >= 1  # every coarse point prolongs to someone
        rows = u.shape[1]


def test_icosphere_default_levels(ico10k):
    p = ico10k
    _check_prolongations(p.U, 10242)
    assert len(p.U) == 1  # 10 242 -> ~1.6 k; the next level would drop below lower_bound = 1000
    assert 1000 <= p.U[0].shape[1] < 6000
    t = p.solver.hierarchy_timing
    for key in ["n_vertices", "hierarchy", "sampling", "cluster", "next_neighborhood", "next_positions",
                "triangle_finding", "triangle_selection", "levels", "PDS"]:
        assert key in t
    assert t["n_vertices"] == 10242 and t["levels"] == len(p.U)


def test_torus_levels_and_coarsening_rate(torus_mid):
    p = torus_mid
    _check_prolongations(p.U, 90000)
    sizes = [90000] + [u.shape[1] for u in p.U]
    assert len(p.U) == 2 and sizes[-1] >= 1000
    for a, b in zip(sizes, sizes[1:]):
        assert 4.0 < a / b < 8.0  # ratio 8 realises ~5.5-6x on a regular grid


def test_samples_and_clusters(ico_small):
    p = ico_small
    s = p.solver
    samples = s.sampling_indices
    nearest = s.nearest_source
    assert len(samples) == len(p.U) == len(nearest)
    n = p.V.shape[0]
    for k, u in enumerate(p.U):
        smp = np.asarray(samples[k])
        near = np.asarray(nearest[k])
        assert smp.shape[0] == u.shape[1] and near.shape[0] == u.shape[0]
        assert (np.diff(smp) > 0).all() and smp[0] == 0  # greedy sweep in index order
        assert (near[smp] == np.arange(smp.shape[0])).all()  # a sample belongs to its own cluster
        assert near.min() == 0 and near.max() == u.shape[1] - 1
        n = u.shape[1]


def test_barycentric_weights_are_convex(ico_small):
    for u in ico_small.U:
        assert u.data.min() >= -1e-12 and u.data.max() <= 1.0 + 1e-12


def test_lower_bound_stops_coarsening():
    V, F = synth.icosphere(3)  # 642 vertices
    V, S, M, neigh = synth.mesh_operators(V, F)
    assert gravomg.MultigridSolver(V, neigh, M).prolongation_matrices == []  # N <= lower_bound
    U = gravomg.MultigridSolver(V, neigh, M, lower_bound=50).prolongation_matrices
    assert len(U) >= 1 and U[-1].shape[1] >= 50
    _check_prolongations(U, 642)


def test_debug_arrays(ico_small):
    p = ico_small
    s = p.new_solver(debug=True)
    pts = s.level_points
    tris = s.all_triangles
    miss = s.notrimap
    assert len(pts) == len(tris) == len(miss) == len(p.U)
    for k, u in enumerate(p.U):
        assert pts[k].shape == (u.shape[1], 3)
        t = np.asarray(tris[k])
        assert t.shape[1] == 3 and t.min() >= 0 and t.max() < u.shape[1]
        assert (t[:, 0] < t[:, 1]).all() and (t[:, 1] < t[:, 2]).all()
        assert len(miss[k]) == u.shape[0]
    assert p.solver.level_points == []  # only filled with debug=True, as upstream


def test_weighting_variants(ico_small):
    p = ico_small
    for w in (gravomg.Weighting.UNIFORM, gravomg.Weighting.INVDIST):
        U = p.new_solver(weighting=w).prolongation_matrices
        _check_prolongations(U, p.V.shape[0])
    U = p.new_solver(weighting=gravomg.Weighting.UNIFORM).prolongation_matrices[0].tocsr()
    counts = np.diff(U.indptr)
    # uniform everywhere except the "closest three" fallback rows, which are inverse-distance
    # weighted under every scheme (multigrid_solver.cpp:430-447)
    uniform = np.isclose(U.data, np.repeat(1.0 / counts, counts), atol=1e-12)
    assert uniform.mean() > 0.99


def test_nested_hierarchy_keeps_samples_fixed(ico_small):
    p = ico_small
    s = p.new_solver(nested=True)
    U0 = s.prolongation_matrices[0].tocsr()
    smp = np.asarray(s.sampling_indices[0])
    rows = U0[smp]
    assert (np.diff(rows.indptr) == 1).all() and (rows.indices == np.arange(len(smp))).all()
    np.testing.assert_array_equal(rows.data, 1.0)


def test_set_and_get_prolongations_round_trip(ico_small):
    p = ico_small
    s = p.new_solver(build_hierarchy=False)
    assert s.prolongation_matrices == []
    s.set_prolongation_matrices(p.U)
    back = s.prolongation_matrices
    assert len(back) == len(p.U)
    for a, b in zip(back, p.U):
        assert abs(a - b).max() == 0
    with pytest.raises(RuntimeError):
        s.set_prolongation_matrices([p.U[1]])  # wrong row count for level 0


def test_point_cloud_knn_input():
    P = synth.torus_cloud(60, seed=1)  # 3 600 points
    nbr = synth.knn_grid(P, 60, k=8)
    L, M = synth.knn_graph_laplacian(nbr)
    neigh = gravomg.neighbors_from_stiffness(L)
    U = gravomg.MultigridSolver(P, neigh, M, lower_bound=50).prolongation_matrices
    assert len(U) >= 2
    _check_prolongations(U, 3600)


def test_rejected_inputs(ico_small):
    p = ico_small
    with pytest.raises(RuntimeError):
        p.new_solver(sampling_strategy=gravomg.Sampling.FPS)
    with pytest.raises(RuntimeError):
        p.new_solver(sig06=True)
    bad = p.neigh.copy()
    bad[0, 0] = p.V.shape[0] + 5
    with pytest.raises(RuntimeError):
        gravomg.MultigridSolver(p.V, bad, p.M)
    with pytest.raises(RuntimeError):
        gravomg.MultigridSolver(p.V, p.neigh, p.S.tocsr())  # mass must be diagonal
