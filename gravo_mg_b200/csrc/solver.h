// Solver handle behind the C ABI: hierarchy (host), device-resident levels, the V-cycle as a
// list of kernel launches replayed through a CUDA graph, and the reference's timing maps.
// Mirrors the state of MGBS::MultigridSolver that solve() touches
// (reference gravomg/include/gravomg/multigrid_solver.h:105-108, 131-134, 142-146, 157-159).
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gravomg_b200.h"
#include "dist_plan.h"
#include "hierarchy.h"
#include "host_sparse.h"

struct gmg_solver;

namespace gmg {

class EngineBase {
public:
    virtual ~EngineBase() = default;
    // wait = false: return once the copies are enqueued (the caller's buffers have been read)
    virtual void stage_system(int64_t n, const int* indptr, const int* indices, const double* data, const double* rhs,
                              int K, bool wait) = 0;
    virtual void solve_staged() = 0;
    virtual void fetch_solution(double* x_out) = 0;
    virtual double residual(int64_t n, const int* indptr, const int* indices, const double* data, const double* rhs,
                            const double* x, int K, int type) = 0;
    virtual void level_op(int kind, int level, const double* a, const double* b, double* out, int sweeps) = 0;
    virtual double time_op(int kind, int level, int reps) = 0;
    virtual void smoother_weights(int level, double* rho, double* pre, double* post) = 0;
    virtual void get_level_matrix(int level, int* indptr, int* indices, double* data) = 0;
    virtual void dist_init(const void* nccl_unique_id) = 0;
    virtual void invalidate_hierarchy() = 0;
    virtual void invalidate_cycle() = 0;
    virtual bool level_info(int level, int64_t* rows, int64_t* nnz_a, int64_t* nnz_u) = 0;
    virtual void kernel_profile(int kind, int level, double* total_ms, int64_t* launches) = 0;
    virtual void reset_kernel_profile() = 0;
    // ---- device-resident systems: values and right-hand side already in HBM
    virtual void update_values_device(const double* d_vals, const double* d_rhs, int K) = 0;
    virtual void fetch_solution_device(double* d_x) = 0;
    // ---- operator assembly on the device for triangle meshes (mesh_assembly.h)
    virtual void mesh_attach(int64_t nf, const int* faces) = 0;
    virtual void mesh_set_positions(const double* pos) = 0;
    virtual void mesh_stiffness() = 0;
    virtual void mesh_mass(int type) = 0;
    virtual void mesh_system(double alpha, double beta, const double* y, int K) = 0;
    virtual void mesh_flow(double tau, int mass_type, int steps) = 0;
    virtual void mesh_get(int which, double* out) = 0;  // 0 positions n x 3, 1 S values, 2 mass, 3 lhs values, 4 rhs n x K
};

struct SolverState {
    gmg_params params;
    int64_t n = 0;
    std::vector<double> mass_diag;   // lumped mass (constructor argument)
    Hierarchy hier;                  // U[k] and the debug arrays
    bool use_graph = true;
    int loop_mode = 1;               // 1 device while-graph (the whole cycle loop is one launch), 0 host loop (one sync per cycle)
    int kernel_path = 0;             // 0 staged (TMA) where it fits, 1 direct everywhere
    int restrict_path = -1;          // kernel path of the restriction operators only (-1 = follow kernel_path)
    int staged_lanes_r = 0;          // the same for the restriction operators U^T only (0 = follow staged_lanes)
    int staged_lanes = 0;            // staged kernels: threads per row; 0 = from the mean row length
    bool profile = false;
    bool trace = false;              // device timeline: (globaltimer, tag) per kernel of the cycles (gmg_get_trace)
    std::vector<unsigned long long> trace_log;
    int tail_rows = 0;               // levels with at most this many rows run inside the fused tail kernel; off by
                                     // default: measured slower than PDL-chained kernels ("Coarse tail")
    int cluster_tail_rows = 0;       // > 0: levels from the first one with at most this many rows down run as ONE thread-block cluster
                                     // (cluster_tail.cuh): operators staged in shared memory, hardware cluster barriers. Off by
                                     // default: measured 49 us against 32 us for the PDL-chained kernels it replaces (config 2)
    bool dist_graph = true;          // multi-GPU: capture the cycle (kernels + NCCL exchanges) into a CUDA graph
    bool p2p = true;                 // multi-GPU: halos and norms through NVLink peer memory (peer_exchange.h);
                                     // false: pack / ncclSend / ncclRecv / unpack and ncclAllReduce
    int dist_shard_setup = -1;       // multi-GPU: sharded levels compute only their share of the Galerkin product and the
                                     // values are all-gathered (NCCL). -1 / 1: on, 0: every rank computes whole products. Measured
                                     // reduction per solve, replicated -> sharded: 0.66 -> 0.76 ms at 2 ranks, 1.22 -> 1.43 ms at 4
                                     // (the all-gathers cost more than the products save) — but only the sharded product lets a
                                     // rank store and upload just its row window of the finest level (dist_window), which saves more
    bool dist_window = true;         // multi-GPU: finest-level operators stored (and uploaded per solve) by row windows, ~1/world per rank
    bool dist_skip_exchange = false; // measurement only: drop the halo exchanges of the cycle (results are wrong)
    bool p2p_fuse = true;            // p2p: pushes fused into the producing kernels, waits into the consuming ones
    bool l2_hints = false;           // L2 eviction-priority hints on the operator slabs of the finest level (evict_last for
                                     // A_0, evict_first for U_0 / U_0^T); measured on config 2: no effect (263.7 us per cycle either way)
    bool fp32_refine = true;         // dtype float32: the fp32 cycle corrects an fp64 iterate (fp64 defect and stopping norm)
    bool fuse_stop = true;           // single GPU, K <= 4: the norm kernel's last CTA applies the stopping rule (no finalize kernel)
    bool fuse_norm = true;           // stopping test fused with the next cycle's first sweep
    bool use_pdl = true;             // programmatic dependent launch of the row-product kernels
    bool coarse_dataflow = true;     // coarse factor as one dataflow kernel (dense_factor.cuh) / one kernel per phase
    int krylov_patience = 0;         // krylov: stop when the residual has not improved for this many iterations (0 = off);
                                     // used by gmg_direct_solve to run to the fp64 rounding floor
    std::shared_ptr<gmg_solver> direct_helper;  // hierarchy-free twin used by gmg_direct_solve for small systems
    int krylov = 0;                  // 0: the reference's loop of cycles; 1: conjugate gradients preconditioned with one cycle;
                                     // 2: plain conjugate gradients (the reference's solverType 4)
    bool diff_form = true;           // finest level: cancellation-free row product sum_{j != i} A_ij (x_j - x_i) + s_i x_i
                                     // (sparse_kernels.cuh, SpmvArgs::diff); false: plain sum_j A_ij x_j
    bool spgemm_plan = true;         // Galerkin products from index-pair lists built once per pattern
    long long spgemm_plan_max_pairs = 1500000000ll;  // 12 GB of pairs; beyond it levels fall back to the searching kernel
    int xfer_threads = -1;           // host threads staging caller buffers through pinned chunks (host_xfer.h);
                                     // -1 = cores / ranks (2..16), 0 = plain pageable cudaMemcpyAsync
    // ---- symbolic phase (host): patterns of every level operator for the staged lhs pattern
    std::vector<HostCsr> r_host;     // R[k] = U[k]^T
    std::vector<HostCsr> a_pat;      // pattern of A_k, k = 0..L (values unused)
    std::vector<HostCsr> ap_pat;     // pattern of A_k U_k, k = 0..L-1
    // ---- multi-GPU layout (world == 1: single GPU)
    DistLayout dist;
    Level0Windows win0;              // multi-GPU: row segments of the finest level this rank stores (compute_level0_windows)
    int64_t replicate_rows = 300000; // levels with at most this many rows are replicated on every rank
    std::map<std::string, double> solver_timing;           // reference solverTiming keys
    std::map<std::string, double> transfer_timing;         // host side of the last stage / fetch (not a reference map)
    std::vector<std::pair<double, double>> convergence;    // (elapsed ms, residue) per cycle
    int64_t last_launches = 0;
    std::string error;
    std::unique_ptr<EngineBase> engine;                    // created on first device use
};

std::unique_ptr<EngineBase> make_engine(SolverState* state);

// Host-only symbolic setup shared by the device engine and the layout queries of the C ABI.
// transposes U (once per hierarchy) and computes the patterns of A_k U_k and of the Galerkin
// operators for the given lhs pattern (multigrid_solver.cpp:1387-1392, structure only).
void compute_level_patterns(SolverState& s, int64_t n, const int* indptr, const int* indices);
// Row ranges of every level and the halo lists of every sharded operator for s.dist.rank / world.
void compute_dist_layout(SolverState& s);
// Multi-GPU data decomposition of the finest level (after compute_dist_layout): which rows of A_0, A_0 U_0, U_0, U_0^T
// and of the right-hand side this rank stores and uploads. Host only.
void compute_level0_windows(SolverState& s);

}  // namespace gmg

struct gmg_solver {
    gmg::SolverState s;
};
