// Host-side Gravo MG hierarchy construction (prerequisite of the V-cycle path, not the path
// being accelerated): graph-Voronoi coarsening with barycentric prolongation, the default
// branch (FASTDISK sampling, non-SIG06, non-ablation) of
// reference gravomg/src/multigrid_solver.cpp:62-469, 471-526, 695-711, 975-1056.
// Plain arrays and std containers; no Eigen.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "host_sparse.h"

namespace gmg {

enum Weighting { BARYCENTRIC = 0, UNIFORM = 1, INVDIST = 2 };

struct HierarchyOptions {
    double ratio = 8.0;
    int low_bound = 1000;
    bool check_voronoi = true;
    bool nested = false;
    int weighting = BARYCENTRIC;
    bool debug = false;
    bool verbose = false;
};

struct Hierarchy {
    std::vector<HostCsr> U;                         // U[k]: n_k x n_{k+1}, rows sorted by column
    std::vector<int64_t> dof;                       // level sizes n_0 .. n_L
    std::vector<std::vector<int>> samples;          // fine index of every coarse point, per level
    std::vector<std::vector<int>> nearest_source;   // cluster id of every fine point, per level
    std::vector<std::vector<double>> level_points;  // coarse positions (n_{k+1} x 3), debug only
    std::vector<std::vector<int>> all_triangles;    // candidate triangles (nt x 3), debug only
    std::vector<std::vector<int>> no_tri_found;     // per fine point: 1 if no triangle contained it
    std::map<std::string, double> timing;           // reference hierarchyTiming keys
};

// pos: n x 3 row-major; neigh: n x kn row-major, -1 padded.
void build_hierarchy(const double* pos, int64_t n, const int* neigh, int kn,
                     const HierarchyOptions& opt, Hierarchy& out);

}  // namespace gmg
