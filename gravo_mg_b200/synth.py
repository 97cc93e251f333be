"""Synthetic inputs for the V-cycle path: meshes, point clouds and their operators.

The reference obtains its inputs from mesh files through libigl / robust_laplacian
(experiments/python/comparisons.py:30-55, demos/smoothing.py:20-31); neither the meshes
nor those packages exist offline, so the same *kinds* of systems are generated here with
vectorised numpy:

* ``icosphere`` / ``torus_grid``      closed triangle meshes (BASELINE configs 1, 2, 3, 5)
* ``cotangent_stiffness``             S = -igl.cotmatrix(V, F)  (positive semi-definite sign)
* ``mass_voronoi`` / ``mass_barycentric``  lumped mass, igl MASSMATRIX_TYPE_VORONOI / _BARYCENTRIC
* ``torus_cloud`` + ``knn_graph_laplacian``  point cloud + symmetrised kNN graph (config 4)
* ``poisson_system`` / ``smoothing_system``  lhs, rhs as in comparisons.py:75-96, smoothing.py:43-47

Everything returns plain numpy / scipy.sparse objects.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .util import normalize_area, neighbors_from_stiffness, homogenize_edges, coalesce_edges


# --------------------------------------------------------------------------- meshes
def icosphere(subdivisions: int):
    """Unit icosphere; ``subdivisions=5`` gives 10 242 vertices / 20 480 faces."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    V = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
         [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
         [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
         [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        n = V.shape[0]
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
        e_sorted = np.sort(e, axis=1)
        key = e_sorted[:, 0] * n + e_sorted[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = V[uniq // n] + V[uniq % n]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        V = np.concatenate([V, mid], axis=0)
        m = n + inv.reshape(3, -1).T  # midpoints of edges (01, 12, 20) per face
        a, b, c = F[:, 0], F[:, 1], F[:, 2]
        ab, bc, ca = m[:, 0], m[:, 1], m[:, 2]
        F = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1),
                            np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], axis=0)
    return np.ascontiguousarray(V), np.ascontiguousarray(F.astype(np.int32))


def torus_grid(nu: int, nv: int, R: float = 1.0, r: float = 0.4):
    """Periodic nu x nv grid on a torus, two triangles per quad (valence 6 everywhere).

    ``torus_grid(1000, 1000)`` is BASELINE config 2 (1 000 000 vertices, 2 000 000 faces).
    Vertex (i, j) has index ``i * nv + j``.
    """
    u = 2.0 * np.pi * np.arange(nu) / nu
    v = 2.0 * np.pi * np.arange(nv) / nv
    uu, vv = np.meshgrid(u, v, indexing="ij")
    V = np.stack([(R + r * np.cos(vv)) * np.cos(uu), (R + r * np.cos(vv)) * np.sin(uu), r * np.sin(vv)], axis=-1)
    V = V.reshape(-1, 3)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    i1 = (i + 1) % nu
    j1 = (j + 1) % nv
    v00 = (i * nv + j).ravel()
    v10 = (i1 * nv + j).ravel()
    v01 = (i * nv + j1).ravel()
    v11 = (i1 * nv + j1).ravel()
    F = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=0)
    return np.ascontiguousarray(V), np.ascontiguousarray(F.astype(np.int32))


def torus_cloud(n_side: int, seed: int = 0, R: float = 1.0, r: float = 0.4, jitter: float = 0.35):
    """Jittered parametric sampling of a torus: ``n_side**2`` points, grid-ordered.

    Stand-in for BASELINE config 4's point cloud; the parametric jitter keeps points on the
    surface and bounds displacement so ``knn_grid`` can find neighbours in a local window.
    """
    rng = np.random.default_rng(seed)
    n = n_side
    du = rng.uniform(-jitter, jitter, size=(n, n))
    dv = rng.uniform(-jitter, jitter, size=(n, n))
    u = 2.0 * np.pi * (np.arange(n)[:, None] + du) / n
    v = 2.0 * np.pi * (np.arange(n)[None, :] + dv) / n
    V = np.stack([(R + r * np.cos(v)) * np.cos(u), (R + r * np.cos(v)) * np.sin(u), r * np.sin(v)], axis=-1)
    return np.ascontiguousarray(V.reshape(-1, 3))


# --------------------------------------------------------------------------- operators
def _face_cots(V, F):
    """Cotangent of the angle at each corner of each face, (nf, 3), and double areas."""
    p0, p1, p2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    e0, e1, e2 = p2 - p1, p0 - p2, p1 - p0  # edge opposite corner 0, 1, 2
    dbl_area = np.linalg.norm(np.cross(e1, e2), axis=1)
    cot0 = -(e1 * e2).sum(1) / dbl_area
    cot1 = -(e2 * e0).sum(1) / dbl_area
    cot2 = -(e0 * e1).sum(1) / dbl_area
    return np.stack([cot0, cot1, cot2], 1), dbl_area


def cotangent_stiffness(V, F, fmt: str = "csc"):
    """Positive semi-definite cotangent stiffness ``S = -igl.cotmatrix(V, F)``.

    ``S_ij = -(cot a_ij + cot b_ij) / 2`` for an edge, ``S_ii = -sum_j S_ij``.
    Returned as CSC by default because igl.cotmatrix returns CSC and
    ``neighbors_from_stiffness`` of the reference depends on that.
    """
    n = V.shape[0]
    F = np.asarray(F, dtype=np.int64)
    cots, _ = _face_cots(V, F)
    # corner c is opposite the edge (c+1, c+2)
    ii = np.concatenate([F[:, 1], F[:, 2], F[:, 0]])
    jj = np.concatenate([F[:, 2], F[:, 0], F[:, 1]])
    w = 0.5 * np.concatenate([cots[:, 0], cots[:, 1], cots[:, 2]])
    rows = np.concatenate([ii, jj, ii, jj])
    cols = np.concatenate([jj, ii, ii, jj])
    vals = np.concatenate([-w, -w, w, w])
    S = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    S.sum_duplicates()
    S = (S + S.T) * 0.5  # bitwise symmetric whatever order the duplicates were summed in
    S = S.tocsc() if fmt == "csc" else S.tocsr()
    S.sort_indices()
    return S


def mass_barycentric(V, F):
    """Lumped mass, a third of each incident triangle's area (igl MASSMATRIX_TYPE_BARYCENTRIC)."""
    n = V.shape[0]
    _, dbl = _face_cots(V, np.asarray(F, dtype=np.int64))
    m = np.zeros(n)
    for c in range(3):
        np.add.at(m, F[:, c], dbl / 6.0)
    return sp.diags(m, format="csr")


def mass_voronoi(V, F):
    """Lumped mixed-Voronoi mass (igl MASSMATRIX_TYPE_VORONOI; Meyer et al. 2003).

    Non-obtuse triangle: corner i receives ``(l_j^2 cot_j + l_k^2 cot_k) / 8``.
    Obtuse triangle: the obtuse corner receives area/2, the other two area/4.
    """
    n = V.shape[0]
    F = np.asarray(F, dtype=np.int64)
    cots, dbl = _face_cots(V, F)
    area = 0.5 * dbl
    p0, p1, p2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    l2 = np.stack([((p2 - p1) ** 2).sum(1), ((p0 - p2) ** 2).sum(1), ((p1 - p0) ** 2).sum(1)], 1)  # opposite corner c
    contrib = np.empty((F.shape[0], 3))
    for c in range(3):
        j, k = (c + 1) % 3, (c + 2) % 3
        contrib[:, c] = (l2[:, j] * cots[:, j] + l2[:, k] * cots[:, k]) / 8.0
    obtuse = cots < 0.0
    any_obtuse = obtuse.any(axis=1)
    if any_obtuse.any():
        ob = np.where(obtuse[any_obtuse], 0.5, 0.25) * area[any_obtuse, None]
        contrib[any_obtuse] = ob
    m = np.zeros(n)
    for c in range(3):
        np.add.at(m, F[:, c], contrib[:, c])
    return sp.diags(m, format="csr")


def knn_grid(V, n_side: int, k: int = 8, window: int = 2):
    """k nearest neighbours for a grid-ordered periodic cloud from ``torus_cloud``.

    Candidates are the (2*window+1)^2 - 1 parametric neighbours (periodic); O(N) memory and
    time, unlike a KD-tree at 20 M points. Returns (N, k) int64 neighbour indices.
    """
    n = n_side
    idx = np.arange(n * n).reshape(n, n)
    P = V.reshape(n, n, 3)
    offs = [(a, b) for a in range(-window, window + 1) for b in range(-window, window + 1) if (a, b) != (0, 0)]
    d = np.empty((n, n, len(offs)), dtype=np.float32)
    for o, (a, b) in enumerate(offs):
        Q = np.roll(P, shift=(-a, -b), axis=(0, 1))
        d[:, :, o] = ((P - Q) ** 2).sum(-1)
    sel = np.argpartition(d, k - 1, axis=2)[:, :, :k]
    out = np.empty((n, n, k), dtype=np.int64)
    offs_a = np.array([o[0] for o in offs])
    offs_b = np.array([o[1] for o in offs])
    ii = np.arange(n)[:, None, None]
    jj = np.arange(n)[None, :, None]
    out = idx[(ii + offs_a[sel]) % n, (jj + offs_b[sel]) % n]
    return out.reshape(n * n, k)


def knn_graph_laplacian(nbr, n: int | None = None):
    """Symmetrised unit-weight kNN graph Laplacian ``L = D - W`` (PSD) and ``M = I / N``.

    ``nbr`` is (N, k) neighbour indices (directed); the graph is made undirected the way
    gravomg.util.knn_undirected does (reference util.py:19-27).
    """
    N = nbr.shape[0] if n is None else n
    k = nbr.shape[1]
    src = np.repeat(np.arange(N, dtype=np.int64), k)
    dst = nbr.reshape(-1).astype(np.int64)
    W = sp.coo_matrix((np.ones(src.shape[0]), (src, dst)), shape=(N, N)).tocsr()
    W = W + W.T
    W.data[:] = 1.0
    W.setdiag(0.0)
    W.eliminate_zeros()
    deg = np.asarray(W.sum(axis=1)).ravel()
    L = (sp.diags(deg) - W).tocsr()
    L.sort_indices()
    M = sp.diags(np.full(N, 1.0 / N), format="csr")
    return L, M


# --------------------------------------------------------------------------- systems
def mesh_operators(V, F, mass: str = "voronoi", normalize: bool = True):
    """(V, S, M, neigh) the way comparisons.py:30-55 prepares a mesh."""
    if normalize:
        V = normalize_area(V, F)
    S = cotangent_stiffness(V, F)
    M = mass_voronoi(V, F) if mass == "voronoi" else mass_barycentric(V, F)
    neigh = neighbors_from_stiffness(S)
    return np.ascontiguousarray(V), S, M, neigh


def poisson_system(S, M, tau: float = 1e-6, seed: int = 42, k: int = 1):
    """``lhs = tau*M + S``, ``rhs = M @ N(0,1)`` (comparisons.py:76, 85-96)."""
    lhs = (M * tau + S).tocsr()
    lhs.sort_indices()
    rng = np.random.default_rng(seed=seed)
    y = rng.standard_normal((S.shape[0], k))
    return lhs, np.ascontiguousarray(M @ y)


def smoothing_system(V, S, M, tau: float = 1e-3):
    """``lhs = M + tau*S``, ``rhs = M @ V`` (K = 3) (smoothing.py:43-47)."""
    lhs = (M + tau * S).tocsr()
    lhs.sort_indices()
    return lhs, np.ascontiguousarray(M @ V)
