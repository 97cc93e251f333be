"""Host side of the device operator assembly (gmg_mesh_pattern, no GPU): the sparsity pattern staged for a
mesh is the pattern scipy gives alpha M + beta S (vertex adjacency + diagonal, sorted columns)."""
import ctypes as C

import numpy as np
import pytest

from gravo_mg_b200 import synth
from gravo_mg_b200._lib import lib, i32


def _pattern(n, F):
    F = np.ascontiguousarray(F, dtype=np.int32)
    indptr = np.empty(n + 1, dtype=np.int32)
    nnz = C.c_int64()
    assert lib.gmg_mesh_pattern(n, F.shape[0], i32(F), i32(indptr), None, C.byref(nnz)) == 0
    indices = np.empty(nnz.value, dtype=np.int32)
    assert lib.gmg_mesh_pattern(n, F.shape[0], i32(F), i32(indptr), i32(indices), C.byref(nnz)) == 0
    return indptr, indices


@pytest.mark.parametrize("mesh", ["ico", "torus"])
def test_mesh_pattern_is_the_scipy_pattern(mesh):
    V, F = synth.icosphere(3) if mesh == "ico" else synth.torus_grid(40, 30)
    V, S, M, _ = synth.mesh_operators(V, F)
    lhs = (M + 0.01 * S).tocsr()
    lhs.sort_indices()
    n = V.shape[0]
    indptr, indices = _pattern(n, F)
    import scipy.sparse as sp

    got = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(n, n))
    assert got.has_sorted_indices and (got.diagonal() == 1).all() and abs(got - got.T).nnz == 0
    # every mesh edge is stored; scipy drops the edges whose two cotangents cancel to exactly 0
    E = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    adj = sp.csr_matrix((np.ones(len(E)), (E[:, 0], E[:, 1])), shape=(n, n))
    adj = ((adj + adj.T + sp.identity(n)) > 0).astype(float).tocsr()
    assert abs(got - adj).nnz == 0
    have = (abs(lhs) > 0).astype(float)
    extra = got - have
    assert extra.min() >= 0 and extra.sum() == got.nnz - lhs.nnz
    if mesh == "ico":
        np.testing.assert_array_equal(indptr, lhs.indptr)
        np.testing.assert_array_equal(indices, lhs.indices)


def test_mesh_pattern_rejects_bad_faces():
    F = np.array([[0, 1, 7]], dtype=np.int32)
    indptr = np.empty(4, dtype=np.int32)
    nnz = C.c_int64()
    assert lib.gmg_mesh_pattern(3, 1, i32(F), i32(indptr), None, C.byref(nnz)) != 0
    assert b"out of range" in lib.gmg_last_error(None)
