// Row-range domain decomposition of the V-cycle across the GPUs of one box (host-only part).
//
// The reference is single-process; this layout is new design. One process per GPU;
// every rank holds the global hierarchy (built redundantly on the host) and the global operators
// on its device (the Galerkin setup is replicated), but runs the V-cycle only on its contiguous
// row range of every *sharded* level. Vectors stay global-length and globally indexed, so a halo
// exchange writes a peer's entries at their global positions and the kernels need no local
// renumbering. Levels at or below `replicate_rows` rows are replicated: every rank computes all
// rows (no halos where the cycle is latency bound); the restriction into the first replicated
// level is followed by an all-gather.
//
//   level 0 ranges     equal row counts
//   level k+1 ranges   coarse point c lives with the rank that owns its sample vertex; samples
//                      ascend with the fine index (multigrid_solver.cpp:979-1011), so coarse
//                      ranges are contiguous too
//   halo of (M, rank)  the columns of M's local rows that fall outside the rank's own range of
//                      the gathered vector, grouped by owning peer, ascending
#pragma once
#include <cstdint>
#include <vector>

#include "host_sparse.h"

namespace gmg {

enum HaloOp { HALO_A = 0, HALO_R = 1, HALO_P = 2 };

struct HaloLists {
    // per peer: global indices this rank sends (owned here, needed there) / receives
    std::vector<std::vector<int>> send, recv;
};

struct DistLayout {
    int rank = 0, world = 1;
    int first_replicated = 0;                      // levels >= this are replicated on every rank
    std::vector<std::vector<int64_t>> ranges;      // [level][world + 1] row offsets
    std::vector<HaloLists> halo[3];                // [op][level]; empty for replicated levels
    bool sharded(int level) const { return world > 1 && level < first_replicated; }
    int64_t begin(int level) const { return ranges[level][rank]; }
    int64_t end(int level) const { return ranges[level][rank + 1]; }
};

// Row segments of the finest level a rank stores and uploads (engine.cu, DevMat::segs): ascending, disjoint ranges.
typedef std::vector<std::pair<int64_t, int64_t>> RowRanges;
struct Level0Windows {
    bool on = false;      // decided from the layout and the options only: identical on every rank
    RowRanges a_rows;     // A_0 and A_0 U_0: own rows + the rows this rank's restriction gathers
    RowRanges p_rows;     // U_0: own rows + every column of the rows above
    RowRanges c_rows;     // U_0^T: this rank's coarse rows
    RowRanges rhs_rows;   // right-hand side / x0: own rows + the entries of x this rank's rows gather (all rows when off)
};
// Marked rows as ascending ranges; gaps of at most max_gap unmarked rows are swallowed (fewer, larger segments).
RowRanges merge_marked_rows(const std::vector<char>& mark, int64_t max_gap);

// samples[k][c] = fine index (level k) of coarse point c (level k + 1); level_rows[k] = n_k.
void build_ranges(const std::vector<int64_t>& level_rows, const std::vector<std::vector<int>>& samples, int world,
                  int64_t replicate_rows, DistLayout& out);

// Halo lists of one matrix whose rows are split by `row_ranges` and whose columns (the gathered
// vector) are split by `col_ranges`.
HaloLists build_halo(const HostCsr& m, const std::vector<int64_t>& row_ranges, const std::vector<int64_t>& col_ranges,
                     int rank);

}  // namespace gmg
