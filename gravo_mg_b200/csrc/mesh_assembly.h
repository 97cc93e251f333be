// Operator assembly on the device for triangle meshes: cotangent stiffness S,
// lumped mass M, lhs = alpha M + beta S, rhs = M Y, and normalize_area — the host work the
// reference's demos repeat around every solve (demos/conformal_flow.py:54-59,
// demos/smoothing.py:43-47, experiments/python/comparisons.py:39-55, 75-79; the operators are
// igl.cotmatrix / igl.massmatrix upstream, gravomg/util.py:46-55 for normalize_area).
//
// Everything is a gather: the host builds, once per mesh, for every stored entry (i, j) of the
// operator pattern the list of face corners opposite to the edge (i, j), and for every vertex
// the list of its face corners, both in face order. The kernels then sum per entry / per vertex
// in that fixed order: no atomics, deterministic, S bitwise symmetric.
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"
#include "host_sparse.h"

namespace gmg {

enum MeshMassType { MESH_MASS_BARYCENTRIC = 0, MESH_MASS_VORONOI = 1 };

// Host side, once per mesh.
struct MeshTopology {
    int64_t n = 0, nf = 0;
    HostCsr pattern;                      // vertex adjacency + diagonal, sorted columns (= pattern of alpha M + beta S)
    std::vector<int> entry_off;           // nnz + 1: corners opposite to the edge of entry e are entry_corner[entry_off[e] .. entry_off[e+1])
    std::vector<int> entry_corner;        // corner id = 3 * face + corner
    std::vector<int> vert_off;            // n + 1
    std::vector<int> vert_corner;         // corner ids whose vertex is v
};
MeshTopology build_mesh_topology(int64_t n, int64_t nf, const int* faces);

class MeshAssembler {
public:
    void attach(const MeshTopology& topo, const int* faces, cudaStream_t s);
    bool attached() const { return n_ > 0; }
    void detach() { n_ = 0, nf_ = 0; }
    int64_t n() const { return n_; }
    int64_t nnz() const { return nnz_; }

    // Per-face corner cotangents and lumped-mass shares of the vertex positions `pos` (n x 3, device).
    void face_geometry(const double* pos, int mass_type, cudaStream_t s);
    // S on the attached pattern from the cotangents of the last face_geometry call.
    void stiffness(const int* rowptr, const int* colidx, double* s_vals, cudaStream_t s);
    // Lumped mass per vertex from the shares of the last face_geometry call.
    void mass(double* m, cudaStream_t s);
    // a_vals = alpha M + beta S on the pattern, rhs = M Y (Y: n x K row-major, device).
    void system(const int* rowptr, const int* colidx, double alpha, double beta, const double* s_vals, const double* m,
                const double* y, int K, double* a_vals, double* rhs, cudaStream_t s);
    // pos_out = normalize_area(x) (gravomg/util.py:52-55): x / sqrt(total face area of x), then centred.
    void normalize_area(const double* x, double* pos_out, cudaStream_t s);

private:
    int64_t n_ = 0, nf_ = 0, nnz_ = 0;
    DeviceBuffer<int> faces_, entry_off_, entry_corner_, vert_off_, vert_corner_;
    DeviceBuffer<double> cot_, share_, partial_, scalars_;
};

}  // namespace gmg
