"""gravo_mg_b200.util against golden vectors produced by the REFERENCE's util.py
(tests/golden/make_golden.py imports /root/reference/gravomg_bindings/src/gravomg/util.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from gravo_mg_b200 import synth, util

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "util_reference.npz"))


@pytest.mark.parametrize("name", ["ico", "torus"])
def test_mesh_helpers_match_reference(name):
    V, F = GOLD[f"{name}_V"], GOLD[f"{name}_F"]
    S = synth.cotangent_stiffness(V, F, fmt="csc")
    got = util.neighbors_from_stiffness(S)
    assert got.dtype == np.int32
    np.testing.assert_array_equal(got, GOLD[f"{name}_neigh_stiffness"])
    # CSR input: the reference scrambles it; here it gives the CSC answer
    np.testing.assert_array_equal(util.neighbors_from_stiffness(sp.csr_matrix(S)), GOLD[f"{name}_neigh_stiffness"])
    np.testing.assert_array_equal(util.neighbors_from_faces(F), GOLD[f"{name}_neigh_faces"])
    np.testing.assert_allclose(util.face_area(V, F), GOLD[f"{name}_face_area"], rtol=0, atol=0)
    np.testing.assert_allclose(util.normalize_area(V, F), GOLD[f"{name}_normalize_area"], rtol=0, atol=0)


def test_cloud_helpers_match_reference():
    P = GOLD["cloud_P"]
    np.testing.assert_array_equal(util.knn(P, 6), GOLD["cloud_knn6"])
    np.testing.assert_array_equal(util.knn_undirected(P, 6), GOLD["cloud_knn_undirected6"])
    np.testing.assert_allclose(util.normalize_bounding_box(P), GOLD["cloud_normalize_bbox"], rtol=0, atol=0)
    np.testing.assert_allclose(util.normalize_axes(P), GOLD["cloud_normalize_axes"], rtol=0, atol=0)


def test_edge_helpers_match_reference():
    ci, cj = util.coalesce_edges(GOLD["edges_i"], GOLD["edges_j"])
    np.testing.assert_array_equal(ci, GOLD["coalesce_i"])
    np.testing.assert_array_equal(cj, GOLD["coalesce_j"])
    np.testing.assert_array_equal(util.homogenize_edges(ci, cj), GOLD["homogenize"])


def test_neighbor_array_layout():
    V, F = synth.icosphere(2)
    S = synth.cotangent_stiffness(V, F)
    neigh = util.neighbors_from_stiffness(S)
    assert neigh.flags["C_CONTIGUOUS"] and neigh.shape[0] == V.shape[0]
    for i in range(V.shape[0]):
        row = neigh[i]
        valid = row[row >= 0]
        assert (row[len(valid):] == -1).all()  # padding only at the tail
        assert i in valid                       # the stiffness diagonal keeps the vertex itself
        assert (np.diff(valid) > 0).all()
