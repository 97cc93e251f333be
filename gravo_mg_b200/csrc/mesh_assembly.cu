// Device-side operator assembly for triangle meshes (see mesh_assembly.h).
#include "mesh_assembly.h"

#include <algorithm>
#include <stdexcept>

namespace gmg {

// ------------------------------------------------------------------ host: topology, once per mesh
MeshTopology build_mesh_topology(int64_t n, int64_t nf, const int* faces) {
    if (n <= 0 || nf <= 0 || !faces) throw std::invalid_argument("mesh needs vertices and faces");
    MeshTopology t;
    t.n = n, t.nf = nf;
    for (int64_t i = 0; i < 3 * nf; ++i)
        if (faces[i] < 0 || faces[i] >= n) throw std::invalid_argument("face index out of range");
    // adjacency + diagonal: count (with duplicates), fill, sort + unique per row
    std::vector<int> cnt((size_t)n + 1, 0);
    for (int64_t v = 0; v < n; ++v) cnt[v + 1] = 1;  // diagonal
    for (int64_t f = 0; f < nf; ++f)
        for (int c = 0; c < 3; ++c) cnt[(size_t)faces[3 * f + c] + 1] += 2;
    for (int64_t v = 0; v < n; ++v) cnt[v + 1] += cnt[v];
    if ((int64_t)cnt[n] < 0) throw std::invalid_argument("mesh too large for 32-bit entry offsets");
    std::vector<int> raw((size_t)cnt[n]);
    std::vector<int> at(cnt.begin(), cnt.end() - 1);
    for (int64_t v = 0; v < n; ++v) raw[at[v]++] = (int)v;
    for (int64_t f = 0; f < nf; ++f)
        for (int c = 0; c < 3; ++c) {
            const int v = faces[3 * f + c];
            raw[at[v]++] = faces[3 * f + (c + 1) % 3];
            raw[at[v]++] = faces[3 * f + (c + 2) % 3];
        }
    HostCsr& p = t.pattern;
    p.rows = p.cols = n;
    p.indptr.assign((size_t)n + 1, 0);
    p.indices.reserve(raw.size() / 2 + (size_t)n);
    for (int64_t v = 0; v < n; ++v) {
        int* b = raw.data() + cnt[v];
        int* e = raw.data() + cnt[v + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        p.indices.insert(p.indices.end(), b, e);
        p.indptr[v + 1] = (int)p.indices.size();
    }
    auto find = [&](int row, int col) {
        const int* b = p.indices.data() + p.indptr[row];
        const int* e = p.indices.data() + p.indptr[row + 1];
        return (int)(std::lower_bound(b, e, col) - p.indices.data());
    };
    // corners opposite to every off-diagonal entry, in face order
    const size_t nnz = p.indices.size();
    t.entry_off.assign(nnz + 1, 0);
    for (int64_t f = 0; f < nf; ++f)
        for (int c = 0; c < 3; ++c) {
            const int u = faces[3 * f + (c + 1) % 3], v = faces[3 * f + (c + 2) % 3];
            if (u == v) continue;  // degenerate face: no edge
            ++t.entry_off[(size_t)find(u, v) + 1];
            ++t.entry_off[(size_t)find(v, u) + 1];
        }
    for (size_t e = 0; e < nnz; ++e) t.entry_off[e + 1] += t.entry_off[e];
    t.entry_corner.resize((size_t)t.entry_off[nnz]);
    std::vector<int> eat(t.entry_off.begin(), t.entry_off.end() - 1);
    for (int64_t f = 0; f < nf; ++f)
        for (int c = 0; c < 3; ++c) {
            const int u = faces[3 * f + (c + 1) % 3], v = faces[3 * f + (c + 2) % 3];
            if (u == v) continue;
            t.entry_corner[eat[find(u, v)]++] = (int)(3 * f + c);
            t.entry_corner[eat[find(v, u)]++] = (int)(3 * f + c);
        }
    // corners of every vertex, in face order
    t.vert_off.assign((size_t)n + 1, 0);
    for (int64_t i = 0; i < 3 * nf; ++i) ++t.vert_off[(size_t)faces[i] + 1];
    for (int64_t v = 0; v < n; ++v) t.vert_off[v + 1] += t.vert_off[v];
    t.vert_corner.resize((size_t)3 * nf);
    std::vector<int> vat(t.vert_off.begin(), t.vert_off.end() - 1);
    for (int64_t i = 0; i < 3 * nf; ++i) t.vert_corner[vat[faces[i]]++] = (int)i;
    return t;
}

// ------------------------------------------------------------------ kernels
namespace {

constexpr int kThreads = 256;
constexpr int kReduceBlocks = 148 * 4;

struct Vec3 {
    double x, y, z;
};
__device__ __forceinline__ Vec3 load3(const double* p, int i) { return {p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]}; }
__device__ __forceinline__ Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Per face: cotangent of the angle at every corner and the corner's share of the lumped mass.
// Corner c is opposite the edge e_c = p_{c+2} - p_{c+1}; cot_c = -(e_{c+1} . e_{c+2}) / (2 area).
__global__ void __launch_bounds__(kThreads) face_geometry_kernel(int64_t nf, const int* __restrict__ faces,
                                                                const double* __restrict__ pos, int mass_type,
                                                                double* __restrict__ cot, double* __restrict__ share) {
    const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const Vec3 p0 = load3(pos, faces[3 * f]), p1 = load3(pos, faces[3 * f + 1]), p2 = load3(pos, faces[3 * f + 2]);
    const Vec3 e0 = sub(p2, p1), e1 = sub(p0, p2), e2 = sub(p1, p0);
    const Vec3 cr = cross(e1, e2);
    const double dbl = sqrt((cr.x * cr.x + cr.y * cr.y) + cr.z * cr.z);
    double c[3];
    c[0] = -dot(e1, e2) / dbl;
    c[1] = -dot(e2, e0) / dbl;
    c[2] = -dot(e0, e1) / dbl;
    double m[3];
    if (mass_type == MESH_MASS_BARYCENTRIC) {
        m[0] = m[1] = m[2] = dbl / 6.0;  // a third of the face area (igl MASSMATRIX_TYPE_BARYCENTRIC)
    } else {
        // mixed Voronoi (Meyer et al. 2003, igl MASSMATRIX_TYPE_VORONOI): non-obtuse face: corner i gets
        // (l_j^2 cot_j + l_k^2 cot_k) / 8; obtuse face: the obtuse corner area / 2, the others area / 4
        const double l2[3] = {dot(e0, e0), dot(e1, e1), dot(e2, e2)};
        const bool obtuse = c[0] < 0.0 || c[1] < 0.0 || c[2] < 0.0;
        const double area = 0.5 * dbl;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            m[i] = obtuse ? (c[i] < 0.0 ? 0.5 : 0.25) * area : (l2[j] * c[j] + l2[k] * c[k]) / 8.0;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) cot[3 * f + i] = c[i], share[3 * f + i] = m[i];
}

// S_ij = -sum over the corners opposite to edge (i, j) of cot / 2 (S = -igl.cotmatrix: positive
// semi-definite sign); S_ii = -sum_{j != i} S_ij. One thread per row.
__global__ void __launch_bounds__(kThreads) stiffness_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                            const int* __restrict__ entry_off, const int* __restrict__ entry_corner,
                                                            const double* __restrict__ cot, double* __restrict__ s_vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double off_sum = 0.0;
    int diag = -1;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        if (colidx[e] == i) {
            diag = e;
            continue;
        }
        double w = 0.0;
        for (int t = entry_off[e]; t < entry_off[e + 1]; ++t) w += 0.5 * cot[entry_corner[t]];
        s_vals[e] = -w;
        off_sum += -w;
    }
    if (diag >= 0) s_vals[diag] = -off_sum;
}

__global__ void __launch_bounds__(kThreads) mass_kernel(int n, const int* __restrict__ vert_off, const int* __restrict__ vert_corner,
                                                       const double* __restrict__ share, double* __restrict__ m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int t = vert_off[i]; t < vert_off[i + 1]; ++t) s += share[vert_corner[t]];
    m[i] = s;
}

// lhs = alpha M + beta S on the pattern of S (M diagonal), rhs = M Y.
__global__ void __launch_bounds__(kThreads) system_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                         double alpha, double beta, const double* __restrict__ s_vals,
                                                         const double* __restrict__ m, const double* __restrict__ y, int K,
                                                         double* __restrict__ a_vals, double* __restrict__ rhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mi = m[i];
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        const double bs = beta * s_vals[e];
        a_vals[e] = colidx[e] == i ? alpha * mi + bs : bs;
    }
    for (int k = 0; k < K; ++k) rhs[(size_t)i * K + k] = mi * y[(size_t)i * K + k];
}

// Deterministic two-stage sums: block b adds its grid-strided share in a fixed tree, one block adds the partials.
template <int NV>
__device__ __forceinline__ void block_tree_sum(double (&v)[NV], double* out) {
    __shared__ double sh[NV][kThreads];
#pragma unroll
    for (int j = 0; j < NV; ++j) sh[j][threadIdx.x] = v[j];
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
#pragma unroll
            for (int j = 0; j < NV; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = sh[j][0];
    }
}

// partial[b] = sum of the areas of this block's faces of the surface x (gravomg/util.py:46-50).
__global__ void __launch_bounds__(kThreads) area_partial_kernel(int64_t nf, const int* __restrict__ faces, const double* __restrict__ x,
                                                               double* __restrict__ partial) {
    double a[1] = {0.0};
    for (int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; f < nf; f += (int64_t)gridDim.x * blockDim.x) {
        const Vec3 v1 = load3(x, faces[3 * f]), v2 = load3(x, faces[3 * f + 1]), v3 = load3(x, faces[3 * f + 2]);
        const Vec3 cr = cross(sub(v2, v1), sub(v3, v1));
        a[0] += sqrt((cr.x * cr.x + cr.y * cr.y) + cr.z * cr.z) / 2.0;
    }
    block_tree_sum<1>(a, partial + blockIdx.x);
}

template <int NV>
__global__ void __launch_bounds__(kThreads) final_sum_kernel(const double* __restrict__ partial, int n_partial, double* __restrict__ out) {
    double v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = 0.0;
    for (int b = threadIdx.x; b < n_partial; b += blockDim.x) {
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] += partial[(size_t)b * NV + j];
    }
    block_tree_sum<NV>(v, out);
}

// pos = x / sqrt(area) and the per-block column sums of pos (for the mean).
__global__ void __launch_bounds__(kThreads) scale_partial_kernel(int n, const double* __restrict__ x, const double* __restrict__ area,
                                                                double* __restrict__ pos, double* __restrict__ partial) {
    const double scale = sqrt(*area);
    double s[3] = {0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = x[3 * (size_t)i + k] / scale;
            pos[3 * (size_t)i + k] = v;
            s[k] += v;
        }
    }
    block_tree_sum<3>(s, partial + (size_t)blockIdx.x * 3);
}

__global__ void __launch_bounds__(kThreads) centre_kernel(int n, const double* __restrict__ col_sums, double* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) pos[3 * (size_t)i + k] -= col_sums[k] / (double)n;
}

inline unsigned blocks_for(int64_t count) { return (unsigned)((count + kThreads - 1) / kThreads); }

}  // namespace

// ------------------------------------------------------------------ MeshAssembler
void MeshAssembler::attach(const MeshTopology& t, const int* faces, cudaStream_t s) {
    n_ = t.n, nf_ = t.nf, nnz_ = t.pattern.nnz();
    faces_.upload(faces, (size_t)3 * nf_, s);
    entry_off_.upload(t.entry_off, s);
    entry_corner_.upload(t.entry_corner, s);
    vert_off_.upload(t.vert_off, s);
    vert_corner_.upload(t.vert_corner, s);
    cot_.ensure((size_t)3 * nf_), share_.ensure((size_t)3 * nf_);
    partial_.ensure((size_t)kReduceBlocks * 3);
    scalars_.ensure(4);
    GMG_CUDA(cudaStreamSynchronize(s));  // the topology vectors are the caller's locals
}

void MeshAssembler::face_geometry(const double* pos, int mass_type, cudaStream_t s) {
    face_geometry_kernel<<<blocks_for(nf_), kThreads, 0, s>>>(nf_, faces_.ptr, pos, mass_type, cot_.ptr, share_.ptr);
    GMG_CUDA(cudaGetLastError());
}

void MeshAssembler::stiffness(const int* rowptr, const int* colidx, double* s_vals, cudaStream_t s) {
    stiffness_kernel<<<blocks_for(n_), kThreads, 0, s>>>((int)n_, rowptr, colidx, entry_off_.ptr, entry_corner_.ptr, cot_.ptr, s_vals);
    GMG_CUDA(cudaGetLastError());
}

void MeshAssembler::mass(double* m, cudaStream_t s) {
    mass_kernel<<<blocks_for(n_), kThreads, 0, s>>>((int)n_, vert_off_.ptr, vert_corner_.ptr, share_.ptr, m);
    GMG_CUDA(cudaGetLastError());
}

void MeshAssembler::system(const int* rowptr, const int* colidx, double alpha, double beta, const double* s_vals, const double* m,
                           const double* y, int K, double* a_vals, double* rhs, cudaStream_t s) {
    system_kernel<<<blocks_for(n_), kThreads, 0, s>>>((int)n_, rowptr, colidx, alpha, beta, s_vals, m, y, K, a_vals, rhs);
    GMG_CUDA(cudaGetLastError());
}

void MeshAssembler::normalize_area(const double* x, double* pos_out, cudaStream_t s) {
    const int nb_f = (int)std::min<int64_t>(kReduceBlocks, blocks_for(nf_));
    area_partial_kernel<<<nb_f, kThreads, 0, s>>>(nf_, faces_.ptr, x, partial_.ptr);
    final_sum_kernel<1><<<1, kThreads, 0, s>>>(partial_.ptr, nb_f, scalars_.ptr);
    const int nb_v = (int)std::min<int64_t>(kReduceBlocks, blocks_for(n_));
    scale_partial_kernel<<<nb_v, kThreads, 0, s>>>((int)n_, x, scalars_.ptr, pos_out, partial_.ptr);
    final_sum_kernel<3><<<1, kThreads, 0, s>>>(partial_.ptr, nb_v, scalars_.ptr + 1);
    centre_kernel<<<blocks_for(n_), kThreads, 0, s>>>((int)n_, scalars_.ptr + 1, pos_out);
    GMG_CUDA(cudaGetLastError());
}

}  // namespace gmg
