#include "host_sparse.h"

#include <algorithm>
#include <numeric>

namespace gmg {

HostCsr transpose(const HostCsr& a, std::vector<int>* perm) {
    HostCsr t;
    t.rows = a.cols;
    t.cols = a.rows;
    const int64_t nnz = a.nnz();
    t.indptr.assign(t.rows + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) t.indptr[a.indices[p] + 1]++;
    for (int64_t r = 0; r < t.rows; ++r) t.indptr[r + 1] += t.indptr[r];
    t.indices.resize(nnz);
    const bool has_vals = !a.data.empty();
    if (has_vals) t.data.resize(nnz);
    if (perm) perm->resize(nnz);
    std::vector<int> cursor(t.indptr.begin(), t.indptr.end() - 1);
    for (int64_t r = 0; r < a.rows; ++r) {
        for (int p = a.indptr[r]; p < a.indptr[r + 1]; ++p) {
            const int q = cursor[a.indices[p]]++;
            t.indices[q] = (int)r;
            if (has_vals) t.data[q] = a.data[p];
            if (perm) (*perm)[q] = p;
        }
    }
    return t;
}

void sort_rows_sum_duplicates(HostCsr& a) {
    const bool has_vals = !a.data.empty();
    std::vector<std::pair<int, double>> row;
    std::vector<int> new_ptr(a.rows + 1, 0);
    int64_t out = 0;
    for (int64_t r = 0; r < a.rows; ++r) {
        row.clear();
        for (int p = a.indptr[r]; p < a.indptr[r + 1]; ++p)
            row.emplace_back(a.indices[p], has_vals ? a.data[p] : 0.0);
        std::stable_sort(row.begin(), row.end(),
                         [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
        for (size_t i = 0; i < row.size();) {
            size_t j = i;
            double s = 0.0;
            while (j < row.size() && row[j].first == row[i].first) s += row[j++].second;
            a.indices[out] = row[i].first;
            if (has_vals) a.data[out] = s;
            ++out;
            i = j;
        }
        new_ptr[r + 1] = (int)out;
    }
    a.indptr.swap(new_ptr);
    a.indices.resize(out);
    if (has_vals) a.data.resize(out);
}

HostCsr spgemm_symbolic(const HostCsr& a, const HostCsr& b) {
    HostCsr c;
    c.rows = a.rows;
    c.cols = b.cols;
    c.indptr.assign(c.rows + 1, 0);
    std::vector<int> marker(b.cols, -1);
    std::vector<int> rowbuf;
    c.indices.reserve((size_t)a.nnz() * 2);
    for (int64_t i = 0; i < a.rows; ++i) {
        rowbuf.clear();
        for (int p = a.indptr[i]; p < a.indptr[i + 1]; ++p) {
            const int k = a.indices[p];
            for (int q = b.indptr[k]; q < b.indptr[k + 1]; ++q) {
                const int j = b.indices[q];
                if (marker[j] != (int)i) {
                    marker[j] = (int)i;
                    rowbuf.push_back(j);
                }
            }
        }
        std::sort(rowbuf.begin(), rowbuf.end());
        c.indices.insert(c.indices.end(), rowbuf.begin(), rowbuf.end());
        c.indptr[i + 1] = (int)c.indices.size();
    }
    return c;
}

void spgemm_numeric(const HostCsr& a, const HostCsr& b, HostCsr& c) {
    c.data.assign(c.indices.size(), 0.0);
    std::vector<int> pos(b.cols, -1);
    for (int64_t i = 0; i < a.rows; ++i) {
        for (int p = c.indptr[i]; p < c.indptr[i + 1]; ++p) pos[c.indices[p]] = p;
        for (int p = a.indptr[i]; p < a.indptr[i + 1]; ++p) {
            const int k = a.indices[p];
            const double v = a.data[p];
            for (int q = b.indptr[k]; q < b.indptr[k + 1]; ++q) c.data[pos[b.indices[q]]] += v * b.data[q];
        }
    }
}

std::vector<int> plan_row_tiles(const std::vector<int>& indptr, int max_rows, int max_nnz, int* max_tile_nnz,
                                int row_begin, int row_end) {
    const int n = row_end < 0 ? (int)indptr.size() - 1 : row_end;
    std::vector<int> tiles;
    tiles.push_back(row_begin);
    int worst = 0;
    int r = row_begin;
    while (r < n) {
        const int base = indptr[r] & ~3;  // slabs are fetched from a 16-byte aligned entry
        int e = r;
        while (e < n && e - r < max_rows && ((indptr[e + 1] + 3) & ~3) - base <= max_nnz) ++e;
        if (e == r) e = r + 1;  // one over-long row: own tile, caller checks max_tile_nnz
        worst = std::max(worst, ((indptr[e] + 3) & ~3) - base);
        tiles.push_back(e);
        r = e;
    }
    if (max_tile_nnz) *max_tile_nnz = worst;
    return tiles;
}

}  // namespace gmg
