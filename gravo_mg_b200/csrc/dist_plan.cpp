#include "dist_plan.h"

#include <algorithm>
#include <stdexcept>

namespace gmg {

void build_ranges(const std::vector<int64_t>& level_rows, const std::vector<std::vector<int>>& samples, int world,
                  int64_t replicate_rows, DistLayout& out) {
    const int n_levels = (int)level_rows.size();
    out.world = world;
    out.ranges.assign(n_levels, std::vector<int64_t>(world + 1, 0));
    for (int p = 0; p <= world; ++p) out.ranges[0][p] = level_rows[0] * p / world;
    for (int k = 0; k + 1 < n_levels; ++k) {
        if ((int)samples.size() > k && (int64_t)samples[k].size() == level_rows[k + 1]) {
            const std::vector<int>& s = samples[k];
            for (int p = 0; p <= world; ++p)
                out.ranges[k + 1][p] = std::lower_bound(s.begin(), s.end(), (int)out.ranges[k][p]) - s.begin();
            out.ranges[k + 1][world] = level_rows[k + 1];
        } else {  // injected hierarchy without sample information: equal row counts
            for (int p = 0; p <= world; ++p) out.ranges[k + 1][p] = level_rows[k + 1] * p / world;
        }
    }
    out.first_replicated = n_levels;
    for (int k = 0; k < n_levels; ++k)
        if (level_rows[k] <= replicate_rows) {
            out.first_replicated = k;
            break;
        }
    // the coarsest level holds the replicated direct solve
    out.first_replicated = std::min(out.first_replicated, n_levels - 1);
    if (world <= 1) out.first_replicated = 0;
}

RowRanges merge_marked_rows(const std::vector<char>& mark, int64_t max_gap) {
    RowRanges out;
    const int64_t n_rows = (int64_t)mark.size();
    for (int64_t r = 0; r < n_rows;) {
        if (!mark[r]) {
            ++r;
            continue;
        }
        int64_t e = r + 1;
        while (e < n_rows && mark[e]) ++e;
        if (!out.empty() && r - out.back().second <= max_gap)
            out.back().second = e;
        else
            out.emplace_back(r, e);
        r = e;
    }
    return out;
}

HaloLists build_halo(const HostCsr& m, const std::vector<int64_t>& row_ranges, const std::vector<int64_t>& col_ranges,
                     int rank) {
    const int world = (int)row_ranges.size() - 1;
    HaloLists h;
    h.send.assign(world, {});
    h.recv.assign(world, {});
    std::vector<char> mark(m.cols, 0);
    auto owner = [&](int c) { return (int)(std::upper_bound(col_ranges.begin(), col_ranges.end(), (int64_t)c) - col_ranges.begin()) - 1; };
    // what peer q needs from this rank: columns inside my column range referenced by q's rows
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        std::vector<int>& list = h.send[q];
        for (int64_t r = row_ranges[q]; r < row_ranges[q + 1]; ++r)
            for (int p = m.indptr[r]; p < m.indptr[r + 1]; ++p) {
                const int c = m.indices[p];
                if (c >= col_ranges[rank] && c < col_ranges[rank + 1] && !mark[c]) {
                    mark[c] = 1;
                    list.push_back(c);
                }
            }
        std::sort(list.begin(), list.end());
        for (int c : list) mark[c] = 0;
    }
    // what this rank needs: columns outside my column range referenced by my rows, by owner
    for (int64_t r = row_ranges[rank]; r < row_ranges[rank + 1]; ++r)
        for (int p = m.indptr[r]; p < m.indptr[r + 1]; ++p) {
            const int c = m.indices[p];
            if ((c < col_ranges[rank] || c >= col_ranges[rank + 1]) && !mark[c]) {
                mark[c] = 1;
                h.recv[owner(c)].push_back(c);
            }
        }
    for (auto& list : h.recv) std::sort(list.begin(), list.end());
    return h;
}

}  // namespace gmg
