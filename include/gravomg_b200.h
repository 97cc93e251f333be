/* gravomg_b200 — C ABI of the B200-native Gravo MG V-cycle path.
 *
 * This is the drop-in boundary for ONE path of rubenwiersma/gravo_mg: what the pybind11
 * module `gravomg_bindings` (reference gravomg_bindings/src/cpp/core.cpp:13-139) exposes
 * around MGBS::MultigridSolver::solve (reference gravomg/src/multigrid_solver.cpp:1279-1485,
 * solverType == 2). Host pointers in, host pointers out, plain sizes; device residency
 * (operators, prolongations, Galerkin levels, CUDA graphs) is private to the handle.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; gmg_last_error(h) then
 *     holds a message (h may be NULL for failures of gmg_create).
 *   - sparse matrices are CSR with int32 indices and fp64 values.
 *   - dense (N x K) blocks are row-major (numpy C order), fp64.
 *   - there is no CPU fallback: anything that computes needs a CUDA device and fails
 *     loudly without one. Hierarchy construction and the getters are host-only.
 */
#ifndef GRAVOMG_B200_H
#define GRAVOMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmg_solver* gmg_handle;

/* Enumerations of the reference (gravomg/include/gravomg/multigrid_solver.h:35-52). */
enum { GMG_SAMPLING_FASTDISK = 0, GMG_SAMPLING_POISSONDISK = 1, GMG_SAMPLING_FPS = 2, GMG_SAMPLING_RANDOM = 3, GMG_SAMPLING_MIS = 4 };
enum { GMG_WEIGHTING_BARYCENTRIC = 0, GMG_WEIGHTING_UNIFORM = 1, GMG_WEIGHTING_INVDIST = 2 };
enum { GMG_SMOOTHER_JACOBI = 0, GMG_SMOOTHER_CHEBYSHEV = 1 };
enum { GMG_DTYPE_F64 = 0, GMG_DTYPE_F32 = 1 };

/* Constructor arguments of the reference binding (core.cpp:20-58; Python defaults core.py:8-13)
 * followed by the options this implementation adds. */
typedef struct gmg_params {
    double ratio;               /* 8.0   */
    int32_t low_bound;          /* 1000  */
    int32_t cycle_type;         /* 0 = V-cycle, 1 = F-cycle, 2 = W-cycle (multigrid_solver.cpp:1059-1192) */
    double tolerance;           /* 1e-4  */
    int32_t stopping_criteria;  /* 2 = M-norm relative residual (multigrid_solver.cpp:1228-1277) */
    int32_t pre_iters;          /* 2 */
    int32_t post_iters;         /* 2 */
    int32_t max_iter;           /* 100 */
    int32_t check_voronoi;      /* 1 */
    int32_t nested;             /* 0 */
    int32_t sampling_strategy;  /* GMG_SAMPLING_FASTDISK (only one implemented) */
    int32_t weighting;          /* GMG_WEIGHTING_BARYCENTRIC */
    int32_t sig06;              /* must be 0 */
    int32_t verbose;
    int32_t debug;
    int32_t ablation;           /* must be 0 */
    int32_t ablation_num_points;
    int32_t ablation_random;
    /* ---- additions ---- */
    int32_t smoother;           /* GMG_SMOOTHER_CHEBYSHEV (default): Jacobi sweeps x += w_j D^-1 (b - A x) whose
                                   dampings w_j are the inverse roots of the Chebyshev polynomial of degree
                                   pre_iters (post_iters) on [rho/cheb_alpha, rho], rho = Gershgorin bound of
                                   D^-1 A per level; GMG_SMOOTHER_JACOBI: every sweep uses omega */
    double omega;               /* fixed damping of GMG_SMOOTHER_JACOBI, default 2/3 */
    int32_t dtype;              /* GMG_DTYPE_F64 (default) | GMG_DTYPE_F32 smoother levels */
    int32_t device;             /* CUDA device ordinal, default 0 */
    int32_t build_hierarchy;    /* 1: build U at create (reference behaviour); 0: caller injects U */
    double cheb_alpha;          /* width of the smoothed band, default 10 */
} gmg_params;

/* Fill *p with the reference's Python defaults. */
int gmg_default_params(gmg_params* p);

/* Replaces the binding constructor core.cpp:20-58 (hierarchy built eagerly, core.cpp:49).
 * pos: n x 3 row-major; neigh: n x kn row-major int32 padded with -1; mass: n x n CSR
 * (must be diagonal: the lumped mass every reference caller passes). Host-only: no CUDA call. */
int gmg_create(const gmg_params* p, int64_t n, const double* pos, const int32_t* neigh, int32_t kn,
               const int32_t* m_indptr, const int32_t* m_indices, const double* m_data, gmg_handle* out);

void gmg_destroy(gmg_handle h);
const char* gmg_last_error(gmg_handle h);

/* Change a solve-time setting after construction (the reference exposes them as public
 * members: accuracy, stoppingCriteria, preIters, postIters, maxIter; multigrid_solver.h:131-144).
 * Keys: "tolerance", "stopping_criteria", "pre_iters", "post_iters", "max_iter", "omega",
 * "smoother", "cheb_alpha", and implementation knobs "use_graph" (0/1), "loop_mode" (0 host loop,
 * 1 device while-graph), "kernel_path" (0 staged TMA, 1 direct), "lanes" (staged kernels: threads
 * per row, 0 = chosen from the row length, 1 = one thread per row, sums in CSR order),
 * "pdl" (0/1 programmatic dependent launch of consecutive kernels), "fuse_norm" (0/1 stopping
 * test and next cycle's first sweep in one kernel), "tail_rows" (levels with at most this many
 * rows run inside one persistent kernel with grid barriers; 0 = one kernel per operator), "profile" (0/1 per-kernel
 * event timing), "xfer_threads" (host threads that stage caller-owned buffers through pinned chunks during
 * gmg_solve / gmg_stage_system / gmg_fetch_solution; -1 = auto, 0 = plain pageable copies),
 * "fuse_stop" (0/1 the norm kernel's last CTA applies the stopping rule), "spgemm_plan" (0/1 Galerkin
 * products from per-pattern index-pair lists), "coarse_dataflow" (0/1 coarse factor as one tile-task kernel),
 * "fp32_refine" (0/1 float32 levels correct an fp64 iterate with an fp64 defect), "lanes_r" (threads per row
 * of the restriction operators), "l2_hints" (0/1 L2 eviction-priority hints on the finest operators),
 * "trace" (0/1 device timeline, gmg_get_trace); multi-GPU: "p2p" (1 halo rows through NVLink peer memory,
 * 0 NCCL send/recv), "p2p_fuse" (0/1 pushes and waits fused into the row-product kernels), "dist_graph",
 * "dist_shard_setup" (-1 auto, 0 replicated, 1 sharded Galerkin reduction), "dist_window" (0/1 finest level
 * stored and uploaded by per-rank row windows), "dist_skip_exchange" (measurement only), "diff_form" (0/1
 * cancellation-free row product on the finest level), "krylov" (0 the reference's loop of cycles, 1 conjugate
 * gradients preconditioned with one cycle, 2 plain conjugate gradients = the reference's solverType 4),
 * "krylov_patience" (stop after this many iterations without improvement; 0 off). Unknown keys fail. */
int gmg_set_option(gmg_handle h, const char* key, double value);
int gmg_get_option(gmg_handle h, const char* key, double* value);

/* ---- hierarchy access: prolongation_matrices / set_prolongation_matrices (core.cpp:82-88) ---- */
int gmg_num_levels(gmg_handle h, int32_t* n_prolongations);
int gmg_prolongation_shape(gmg_handle h, int32_t level, int64_t* rows, int64_t* cols, int64_t* nnz);
int gmg_get_prolongation(gmg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data);
/* Drop all prolongations, then set them one level at a time (level must equal the current count). */
int gmg_clear_prolongations(gmg_handle h);
int gmg_set_prolongation(gmg_handle h, int32_t level, int64_t rows, int64_t cols, const int32_t* indptr,
                         const int32_t* indices, const double* data);

/* sampling_indices / nearest_source / level_points (core.cpp:90-116). Query sizes with out == NULL. */
int gmg_get_samples(gmg_handle h, int32_t level, int32_t* out, int64_t* count);
int gmg_get_nearest_source(gmg_handle h, int32_t level, int32_t* out, int64_t* count);
int gmg_get_level_points(gmg_handle h, int32_t level, double* out, int64_t* count);   /* debug=1 only */
int gmg_get_all_triangles(gmg_handle h, int32_t level, int32_t* out, int64_t* count); /* debug=1 only */
int gmg_get_notrimap(gmg_handle h, int32_t level, int32_t* out, int64_t* count);      /* debug=1 only */

/* ---- the hot path: solve (core.cpp:68-72 -> multigrid_solver.cpp:1367-1449) ----
 * x0 = rhs (core.cpp:69); Galerkin operators and the coarse factor are recomputed from the
 * values of A on every call (multigrid_solver.cpp:1387-1401); at least one V-cycle runs; stops
 * when residual(stopping_criteria) <= tolerance or after max_iter cycles. rhs, x_out: n x K. */
int gmg_solve(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_data,
              const double* rhs, double* x_out, int32_t K);

/* The same call split in three so a caller can keep the system resident in HBM:
 * stage = host->device copies (+ symbolic setup when the sparsity pattern changed);
 * solve_staged = reduction + coarse factor + cycles, device only; fetch = device->host. */
int gmg_stage_system(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices,
                     const double* a_data, const double* rhs, int32_t K);
int gmg_solve_staged(gmg_handle h);
int gmg_fetch_solution(gmg_handle h, double* x_out);

/* direct_solve (core.cpp:74-78 -> multigrid_solver.cpp:1287-1321, Eigen::SimplicialLLT of the whole system;
 * writes solverTiming keys direct_factor, direct_solve, direct_residual). n <= 16384: dense Cholesky on the
 * device (the coarsest-level solver applied to the whole system). Larger n: no sparse factorisation exists on
 * the device; the system is solved to the fp64 rounding floor by conjugate gradients preconditioned with the
 * V-cycle until the residual stops decreasing. */
int gmg_direct_solve(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_data,
                     const double* rhs, double* x_out, int32_t K);

/* residualCheck (core.cpp:132-134 -> multigrid_solver.cpp:1228-1277). type 0..3. */
int gmg_residual(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_data,
                 const double* rhs, const double* x, int32_t K, int32_t type, double* out);

/* ---- device-resident systems: the values of lhs (same sparsity pattern as the staged
 * system, CSR order) and the right-hand side (n x K) are already in HBM — assembled by the caller's own
 * kernels — and the solution is wanted there too. Replaces the host->device staging of gmg_solve for
 * repeated-solve loops (demos/conformal_flow.py:54-59); gmg_solve_staged in between. d_* are device
 * pointers of the handle's device; the copies are device-to-device on the solver's stream. */
int gmg_update_values_device(gmg_handle h, const double* d_a_data, const double* d_rhs, int32_t K);
int gmg_fetch_solution_device(gmg_handle h, double* d_x_out);

/* ---- operator assembly on the device for triangle meshes ----
 * What the reference's callers compute on the host around every solve — S = -igl.cotmatrix(V, F),
 * M = igl.massmatrix(V, F, type), lhs = alpha M + beta S, rhs = M Y, V = normalize_area(x)
 * (demos/smoothing.py:28-30, 43-47; demos/conformal_flow.py:22-24, 54-59;
 * experiments/python/comparisons.py:39-55, 75-79; gravomg/util.py:46-55) — as kernels on the resident mesh.
 *   gmg_mesh_attach         faces (nf x 3, int32) -> sparsity pattern (vertex adjacency + diagonal, sorted)
 *                           staged on the device with its symbolic Galerkin setup, gather lists for assembly
 *   gmg_mesh_set_positions  vertex positions n x 3 (host) -> device
 *   gmg_mesh_stiffness      S from the resident positions (kept until recomputed)
 *   gmg_mesh_mass           lumped mass from the resident positions; type 0 barycentric, 1 mixed Voronoi
 *   gmg_mesh_system         lhs = alpha M + beta S, rhs = M Y staged for gmg_solve_staged; Y = y (host, n x K)
 *                           or, with y == NULL, the resident positions (K = 3)
 *   gmg_mesh_flow           `steps` conformal-flow steps without leaving the GPU: M_t = mass(V_t),
 *                           lhs = M_t + tau S, rhs = M_t V_t, V_{t+1} = normalize_area(solve(lhs, rhs));
 *                           totals in timing map 2: flow_assemble_ms, flow_reduction_ms, flow_factor_ms,
 *                           flow_cycles_ms, flow_normalize_ms, flow_iterations
 *   gmg_mesh_get            which: 0 positions (n x 3), 1 S values (nnz, CSR order of the pattern), 2 mass (n),
 *                           3 lhs values (nnz), 4 rhs (n x K) -> host
 *   gmg_mesh_pattern        host only, no handle: the pattern gmg_mesh_attach stages. Query nnz with
 *                           indices == NULL (indptr needs n + 1 slots). */
int gmg_mesh_attach(gmg_handle h, int64_t nf, const int32_t* faces);
int gmg_mesh_set_positions(gmg_handle h, const double* pos);
int gmg_mesh_stiffness(gmg_handle h);
int gmg_mesh_mass(gmg_handle h, int32_t type);
int gmg_mesh_system(gmg_handle h, double alpha, double beta, const double* y, int32_t K);
int gmg_mesh_flow(gmg_handle h, double tau, int32_t mass_type, int32_t steps);
int gmg_mesh_get(gmg_handle h, int32_t which, double* out);
int gmg_mesh_pattern(int64_t n, int64_t nf, const int32_t* faces, int32_t* indptr, int32_t* indices, int64_t* nnz);

/* ---- multi-GPU: row-range domain decomposition over the GPUs of one box (new design: the
 * reference is single-process). One process per GPU, SPMD: every rank creates the
 * same solver, configures (rank, world), joins the NCCL communicator and then makes the SAME
 * gmg_solve / gmg_stage_system / gmg_solve_staged / gmg_fetch_solution calls with the same global
 * inputs; every rank gets the full solution back. Levels with more than `replicate_rows` rows are
 * sharded by contiguous row ranges with NCCL halo exchanges before each operator; smaller
 * levels, the Galerkin setup and the coarse direct solve are replicated (-1 keeps the default). */
int gmg_dist_configure(gmg_handle h, int32_t rank, int32_t world, int64_t replicate_rows);
/* ncclGetUniqueId on the calling rank (usually 0); the caller broadcasts the bytes to all ranks. */
int gmg_dist_unique_id(void* id_out, int64_t capacity, int64_t* size);
/* ncclCommInitRank on the handle's device (collective over all ranks). */
int gmg_dist_init(gmg_handle h, const void* id, int64_t size);
/* Host-only: row ranges and halo lists for the given lhs pattern (what staging computes), so the
 * layout can be inspected and tested without a device. */
int gmg_dist_layout(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices);
/* Host only: sparsity pattern of the operator of a level as the symbolic Galerkin phase computed it (level 0: the lhs
 * pattern; k >= 1: pattern of U^T A U, sorted columns). Query sizes with indptr == indices == NULL. */
int gmg_level_pattern(gmg_handle h, int32_t level, int32_t* indptr, int32_t* indices, int64_t* rows, int64_t* nnz);
/* Data decomposition of the finest level (after gmg_dist_layout or staging): the row segments this rank stores and
 * uploads per solve. which 0: rows of A_0 and A_0 U_0, 1: rows of U_0, 2: rows of U_0^T, 3: rows of the right-hand side.
 * ranges receives (begin, end) pairs; query the number of pairs with ranges == NULL. *enabled = 0: whole operators are
 * stored (single GPU, replicated finest level, options dist_window / dist_shard_setup off). Host only. */
int gmg_dist_windows(gmg_handle h, int32_t which, int64_t* ranges, int64_t* count, int32_t* enabled);
/* ranges[world + 1] of a level; *replicated = 1 when every rank computes all rows of it. */
int gmg_dist_ranges(gmg_handle h, int32_t level, int64_t* ranges, int32_t* replicated);
/* Halo lists of operator op (0: A_k, 1: R_k = U_k^T, 2: U_k) on a level towards one peer: global
 * indices this rank sends / receives, ascending. Query sizes with send == recv == NULL. */
int gmg_dist_halo(gmg_handle h, int32_t op, int32_t level, int32_t peer, int32_t* send, int64_t* n_send, int32_t* recv,
                  int64_t* n_recv);

/* ---- timing maps and convergence trace (multigrid_solver.h:157-159; core.cpp:118-128) ----
 * which: 0 = hierarchyTiming, 1 = solverTiming, 2 = host side of the last transfer (stage_host_ms,
 * fetch_host_ms, h2d_bytes, d2h_bytes, pattern_reused, transfer_threads; not a reference map). Keys come back comma separated, in the
 * alphabetical order std::map gives the reference's CSV writer (utility.cpp:106-131). */
int gmg_timing_keys(gmg_handle h, int32_t which, char* buf, int64_t buflen);
int gmg_get_timing(gmg_handle h, int32_t which, const char* key, double* out);
/* capacity in *count on entry, number of cycles recorded on exit. */
int gmg_get_convergence(gmg_handle h, double* t_ms, double* residue, int32_t* count);

/* ---- measurement support (not part of the reference surface) ----
 * Level sizes of the operators staged on the device: rows and stored entries of A_k, and of U_k. */
int gmg_level_info(gmg_handle h, int32_t level, int64_t* rows, int64_t* nnz_a, int64_t* nnz_u);
/* Jacobi dampings the smoother uses on a level for the system last reduced on the device:
 * pre[pre_iters], post[post_iters] and the Gershgorin bound rho of D^-1 A_k they derive from. */
int gmg_get_smoother_weights(gmg_handle h, int32_t level, double* rho, double* pre, double* post);
/* CSR of the operator of a level as the device holds it (level 0: the staged lhs; k >= 1: the
 * Galerkin operator Abar[k], multigrid_solver.cpp:1389-1391). Sizes from gmg_level_info. */
int gmg_get_level_matrix(gmg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data);
/* One operator of the V-cycle (multigrid_solver.cpp:1059-1088) on caller-supplied host vectors,
 * for op-level parity tests and per-kernel measurement. Needs gmg_stage_system first (K of the
 * staged right-hand side is the K of every vector here); computes the Galerkin chain and the
 * coarse factor if the staged values have not been reduced yet. n_k x K row-major, fp64.
 *   kind 0 jacobi      a = x, b = rhs       out = x after `sweeps` sweeps              (level < L)
 *   kind 1 residual    a = x, b = rhs       out = rhs - A_k x
 *   kind 2 restrict    a = r (n_k)          out = U_k^T r (n_{k+1})                    (level < L)
 *   kind 3 prolong_add a = eps (n_{k+1}), b = x (n_k)   out = x + U_k eps              (level < L)
 *   kind 5 coarse      a = rhs (n_L)        out = Abar[L]^-1 rhs                       (level = L) */
int gmg_level_op(gmg_handle h, int32_t kind, int32_t level, const double* a, const double* b, double* out,
                 int32_t sweeps);
/* Mean device time (microseconds, CUDA events on the launch stream) of `reps` back-to-back
 * launches of one operator kind (0 jacobi, 1 residual, 2 restrict, 3 prolong_add) on a level's
 * resident buffers, after 3 warm-up launches; Jacobi alternates its two vectors as in the cycle. */
int gmg_time_op(gmg_handle h, int32_t kind, int32_t level, int32_t reps, double* us_per_launch);
/* Per-kernel device time accumulated while option "profile" = 1. Kernel kinds:
 * 0 jacobi, 1 residual, 2 restrict, 3 prolong_add, 4 norm, 5 coarse_solve, 7 fused coarse tail
 * (filed under its first level). Level -1 sums levels. */
int gmg_kernel_profile(gmg_handle h, int32_t kind, int32_t level, double* total_ms, int64_t* launches);
int gmg_reset_kernel_profile(gmg_handle h);
/* Number of kernel launches issued (or replayed through graphs) by the last solve. */
/* Device timeline of the last solve (option "trace" = 1): for every kernel of the cycles the
 * globaltimer value at which its dependencies were met, and a tag (rows << 8 | epilogue for the
 * row-product kernels; 100 stopping test, 101 / 102 coarse solve, 103 / 104 multi-GPU push / norm).
 * Query the count with t_ns == NULL. Measurement aid, no reference counterpart. */
int gmg_get_trace(gmg_handle h, uint64_t* t_ns, uint64_t* tags, int64_t* count);
int gmg_last_launch_count(gmg_handle h, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GRAVOMG_B200_H */
