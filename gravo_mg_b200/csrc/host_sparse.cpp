#include "host_sparse.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

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

// Worker threads of the symbolic products: GMG_HOST_THREADS, else min(8, hardware concurrency).
static int host_threads() {
    if (const char* e = std::getenv("GMG_HOST_THREADS")) return std::max(1, std::atoi(e));
    return (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
}

// Rows are independent (Gustavson with a per-thread marker array): contiguous row chunks on worker threads, the
// per-chunk column lists concatenated in row order afterwards. Same output as the sequential loop for any thread count.
HostCsr spgemm_symbolic(const HostCsr& a, const HostCsr& b) {
    HostCsr c;
    c.rows = a.rows;
    c.cols = b.cols;
    c.indptr.assign(c.rows + 1, 0);
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), a.rows / 8192 + 1));
    std::vector<std::vector<int>> chunk_cols(n_threads);
    auto work = [&](int t) {
        const int64_t lo = a.rows * t / n_threads, hi = a.rows * (t + 1) / n_threads;
        std::vector<int> marker(b.cols, -1);
        std::vector<int> rowbuf;
        std::vector<int>& out = chunk_cols[t];
        out.reserve((size_t)(a.indptr[hi] - a.indptr[lo]) * 2);
        for (int64_t i = lo; i < hi; ++i) {
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
            out.insert(out.end(), rowbuf.begin(), rowbuf.end());
            c.indptr[i + 1] = (int)rowbuf.size();  // row length for now
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int64_t i = 0; i < c.rows; ++i) c.indptr[i + 1] += c.indptr[i];
    c.indices.resize((size_t)c.indptr[c.rows]);
    for (int t = 0; t < n_threads; ++t) {
        const int64_t lo = a.rows * t / n_threads;
        if (!chunk_cols[t].empty()) std::memcpy(c.indices.data() + c.indptr[lo], chunk_cols[t].data(), chunk_cols[t].size() * sizeof(int));
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
