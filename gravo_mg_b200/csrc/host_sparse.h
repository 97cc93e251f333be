// Host-side sparse containers and setup-time algorithms (no CUDA in this header).
//
// Everything the device path needs to be told once per hierarchy / sparsity pattern is
// computed here: CSR transposes (R = U^T), the symbolic Galerkin product (patterns of
// T = A*U and Abar = R*T, reference multigrid_solver.cpp:1387-1392) and the row-tile plans
// the staged SpMV kernels iterate over.
#pragma once
#include <cstdint>
#include <vector>

namespace gmg {

struct HostCsr {
    int64_t rows = 0, cols = 0;
    std::vector<int> indptr;    // rows + 1
    std::vector<int> indices;   // nnz
    std::vector<double> data;   // nnz (may be empty for a pattern-only matrix)
    int64_t nnz() const { return indptr.empty() ? 0 : indptr.back(); }
};

// CSR -> CSR of the transpose. Column indices of every output row come out ascending.
// If perm != nullptr it receives, for every output entry, the index of the source entry
// (so values can be refreshed without redoing the structure).
HostCsr transpose(const HostCsr& a, std::vector<int>* perm = nullptr);

// Sort column indices inside every row (values follow) and sum duplicate entries.
void sort_rows_sum_duplicates(HostCsr& a);

// Pattern of C = A * B (Gustavson, sorted columns). Values are not computed.
HostCsr spgemm_symbolic(const HostCsr& a, const HostCsr& b);

// Numeric C = A * B on the host into an existing pattern (used by host-side tests only;
// the solve path does this on the device).
void spgemm_numeric(const HostCsr& a, const HostCsr& b, HostCsr& c);

// Row tiles for the staged kernels: consecutive row ranges with at most `max_rows` rows and
// at most `max_nnz` stored entries each (a single row longer than max_nnz gets its own tile
// and *max_tile_nnz reports it so the caller can fall back to the direct kernel).
// Only rows [row_begin, row_end) are tiled (a rank's row range; the whole matrix by default).
std::vector<int> plan_row_tiles(const std::vector<int>& indptr, int max_rows, int max_nnz,
                                int* max_tile_nnz, int row_begin = 0, int row_end = -1);

}  // namespace gmg
