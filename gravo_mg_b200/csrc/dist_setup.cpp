// Host-only symbolic setup: level patterns and the multi-GPU layout (see solver.h, dist_plan.h).
#include <stdexcept>

#include "solver.h"

namespace gmg {

void compute_level_patterns(SolverState& s, int64_t n, const int* indptr, const int* indices) {
    const auto& U = s.hier.U;
    const int L = (int)U.size();
    if ((int)s.r_host.size() != L) {
        s.r_host.clear();
        for (int k = 0; k < L; ++k) s.r_host.push_back(transpose(U[k]));
    }
    s.a_pat.assign(L + 1, HostCsr());
    s.ap_pat.assign(L, HostCsr());
    HostCsr& a0 = s.a_pat[0];
    a0.rows = a0.cols = n;
    a0.indptr.assign(indptr, indptr + n + 1);
    a0.indices.assign(indices, indices + indptr[n]);
    for (int k = 0; k < L; ++k) {
        if (U[k].rows != s.a_pat[k].rows) throw std::invalid_argument("prolongation matrix has the wrong number of rows for its level");
        s.ap_pat[k] = spgemm_symbolic(s.a_pat[k], U[k]);
        s.a_pat[k + 1] = spgemm_symbolic(s.r_host[k], s.ap_pat[k]);
    }
}

void compute_dist_layout(SolverState& s) {
    const int L = (int)s.hier.U.size();
    std::vector<int64_t> rows(L + 1);
    rows[0] = s.n;
    for (int k = 0; k < L; ++k) rows[k + 1] = s.hier.U[k].cols;
    const int rank = s.dist.rank, world = s.dist.world;
    build_ranges(rows, s.hier.samples, world, s.replicate_rows, s.dist);
    s.dist.rank = rank;
    for (auto& h : s.dist.halo) h.assign(L + 1, HaloLists());
    if (world <= 1) return;
    if ((int)s.a_pat.size() != L + 1) throw std::logic_error("compute_dist_layout needs the level patterns");
    for (int k = 0; k <= L; ++k) {
        if (!s.dist.sharded(k)) break;
        s.dist.halo[HALO_A][k] = build_halo(s.a_pat[k], s.dist.ranges[k], s.dist.ranges[k], rank);
        if (k < L) {
            // R_k: rows are coarse points (level k + 1 ranges), the gathered vector lives on level k
            s.dist.halo[HALO_R][k] = build_halo(s.r_host[k], s.dist.ranges[k + 1], s.dist.ranges[k], rank);
            // U_k: rows on level k, gathered vector on level k + 1 (only exchanged when that is sharded)
            if (s.dist.sharded(k + 1)) s.dist.halo[HALO_P][k] = build_halo(s.hier.U[k], s.dist.ranges[k], s.dist.ranges[k + 1], rank);
        }
    }
}

void compute_level0_windows(SolverState& s) {
    const DistLayout& d = s.dist;
    Level0Windows& w = s.win0;
    w = Level0Windows();
    w.rhs_rows.assign(1, std::make_pair((int64_t)0, (int64_t)s.n));
    const bool shard_setup = d.world > 1 && s.dist_shard_setup != 0;
    w.on = shard_setup && s.dist_window && !s.hier.U.empty() && d.sharded(0);
    if (!w.on) return;
    const HostCsr& a0 = s.a_pat[0];
    const HostCsr& r0 = s.r_host[0];
    std::vector<char> ma((size_t)s.n, 0), mp((size_t)s.n, 0), mr((size_t)s.n, 0);
    for (int64_t r = d.begin(0); r < d.end(0); ++r) ma[r] = mp[r] = mr[r] = 1;
    for (int64_t I = d.begin(1); I < d.end(1); ++I)
        for (int q = r0.indptr[I]; q < r0.indptr[I + 1]; ++q) ma[r0.indices[q]] = 1;
    w.a_rows = merge_marked_rows(ma, 256);
    for (const auto& rr : w.a_rows)
        for (int q = a0.indptr[rr.first]; q < a0.indptr[rr.second]; ++q) mp[a0.indices[q]] = 1;
    w.p_rows = merge_marked_rows(mp, 256);
    w.c_rows.assign(1, std::make_pair(d.begin(1), d.end(1)));
    // x0 = rhs: this rank reads its own rows of b and x plus the entries of x its rows gather
    for (const auto& list : d.halo[HALO_A][0].recv)
        for (int c : list) mr[c] = 1;
    w.rhs_rows = merge_marked_rows(mr, 256);
}

}  // namespace gmg
