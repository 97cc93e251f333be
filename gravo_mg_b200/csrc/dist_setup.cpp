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

}  // namespace gmg
