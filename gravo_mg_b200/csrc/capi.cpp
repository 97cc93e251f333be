// extern "C" surface declared in include/gravomg_b200.h. Every entry point catches C++
// exceptions and turns them into a status code plus gmg_last_error().
#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <memory>
#include <stdexcept>
#include <string>

#include "mesh_assembly.h"
#include "nccl_dl.h"
#include "solver.h"

using gmg::SolverState;

namespace {

thread_local std::string g_create_error;

template <typename F>
int guarded(gmg_handle h, F&& body) {
    try {
        body();
        return 0;
    } catch (const std::exception& e) {
        if (h)
            h->s.error = e.what();
        else
            g_create_error = e.what();
        return 1;
    } catch (...) {
        if (h)
            h->s.error = "unknown error";
        else
            g_create_error = "unknown error";
        return 1;
    }
}

void require(bool ok, const char* msg) {
    if (!ok) throw std::invalid_argument(msg);
}

gmg::EngineBase& engine(gmg_handle h) {
    if (!h->s.engine) h->s.engine = gmg::make_engine(&h->s);
    return *h->s.engine;
}

template <typename V>
void copy_out(const V& src, typename V::value_type* out, int64_t* count) {
    if (out) {
        require(*count >= (int64_t)src.size(), "output buffer too small");
        std::copy(src.begin(), src.end(), out);
    }
    *count = (int64_t)src.size();
}

}  // namespace

extern "C" {

int gmg_default_params(gmg_params* p) {
    if (!p) return 1;
    std::memset(p, 0, sizeof *p);
    p->ratio = 8.0;
    p->low_bound = 1000;
    p->cycle_type = 0;
    p->tolerance = 1e-4;
    p->stopping_criteria = 2;
    p->pre_iters = 2;
    p->post_iters = 2;
    p->max_iter = 100;
    p->check_voronoi = 1;
    p->nested = 0;
    p->sampling_strategy = GMG_SAMPLING_FASTDISK;
    p->weighting = GMG_WEIGHTING_BARYCENTRIC;
    p->ablation_num_points = 3;
    p->smoother = GMG_SMOOTHER_CHEBYSHEV;
    p->cheb_alpha = 10.0;
    p->omega = 2.0 / 3.0;
    p->dtype = GMG_DTYPE_F64;
    p->device = 0;
    p->build_hierarchy = 1;
    return 0;
}

int gmg_create(const gmg_params* p, int64_t n, const double* pos, const int32_t* neigh, int32_t kn,
               const int32_t* m_indptr, const int32_t* m_indices, const double* m_data, gmg_handle* out) {
    if (out) *out = nullptr;
    return guarded(nullptr, [&] {
        require(p && out, "null argument");
        require(n > 0 && n < (int64_t)1 << 31, "number of points must be in (0, 2^31)");
        require(pos && neigh && kn > 0, "positions and a padded neighbour array are required");
        require(m_indptr && m_indices && m_data, "mass matrix is required");
        require(p->sampling_strategy == GMG_SAMPLING_FASTDISK, "only Sampling.FASTDISK is implemented (the others are paper ablations)");
        require(!p->sig06 && !p->ablation, "the SIG06 and ablation hierarchies are out of scope");
        require(p->weighting >= 0 && p->weighting <= 2, "unknown weighting scheme");
        require(p->smoother == GMG_SMOOTHER_JACOBI || p->smoother == GMG_SMOOTHER_CHEBYSHEV, "unknown smoother");
        require(p->cheb_alpha > 1.0, "cheb_alpha must be > 1");
        require(p->pre_iters >= 0 && p->pre_iters <= 16 && p->post_iters >= 0 && p->post_iters <= 16, "sweep counts must be 0..16");
        require(p->dtype == GMG_DTYPE_F64 || p->dtype == GMG_DTYPE_F32, "unknown dtype");
        std::unique_ptr<gmg_solver> h(new gmg_solver());
        SolverState& s = h->s;
        s.params = *p;
        s.n = n;
        s.mass_diag.assign(n, 0.0);
        for (int64_t i = 0; i < n; ++i) {
            for (int q = m_indptr[i]; q < m_indptr[i + 1]; ++q) {
                if (m_indices[q] == i)
                    s.mass_diag[i] += m_data[q];
                else
                    require(m_data[q] == 0.0, "mass matrix must be diagonal (lumped), as every caller of the reference passes");
            }
        }
        for (int64_t i = 0; i < n; ++i)
            for (int j = 0; j < kn; ++j) require(neigh[i * kn + j] >= -1 && neigh[i * kn + j] < n, "neighbour index out of range");
        if (p->build_hierarchy) {
            gmg::HierarchyOptions o;
            o.ratio = p->ratio, o.low_bound = p->low_bound, o.check_voronoi = p->check_voronoi != 0, o.nested = p->nested != 0;
            o.weighting = p->weighting, o.debug = p->debug != 0, o.verbose = p->verbose != 0;
            gmg::build_hierarchy(pos, n, neigh, kn, o, s.hier);
        } else {
            s.hier.dof.push_back(n);
            s.hier.timing["n_vertices"] = (double)n;
        }
        *out = h.release();
    });
}

void gmg_destroy(gmg_handle h) { delete h; }

const char* gmg_last_error(gmg_handle h) { return h ? h->s.error.c_str() : g_create_error.c_str(); }

int gmg_set_option(gmg_handle h, const char* key, double value) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(key != nullptr, "null key");
        SolverState& s = h->s;
        const std::string k(key);
        bool cycle = false, hierarchy = false;
        // a rejected value must not stay behind (the cached launch list would run with it): keep the
        // validated fields and put them back if a check below throws
        const gmg_params saved_params = s.params;
        const int saved_xfer = s.xfer_threads, saved_lanes = s.staged_lanes, saved_lanes_r = s.staged_lanes_r;
        try {
        if (k == "tolerance") s.params.tolerance = value;
        else if (k == "max_iter") s.params.max_iter = (int)value;
        else if (k == "stopping_criteria") s.params.stopping_criteria = (int)value, cycle = true;
        else if (k == "pre_iters") s.params.pre_iters = (int)value, cycle = true;
        else if (k == "post_iters") s.params.post_iters = (int)value, cycle = true;
        else if (k == "omega") s.params.omega = value, cycle = true;
        else if (k == "smoother") s.params.smoother = (int)value, cycle = true;
        else if (k == "cheb_alpha") s.params.cheb_alpha = value, cycle = true;
        else if (k == "lanes") s.staged_lanes = (int)value, hierarchy = true;
        else if (k == "lanes_r") s.staged_lanes_r = (int)value, hierarchy = true;
        else if (k == "restrict_path") s.restrict_path = (int)value, hierarchy = true;
        else if (k == "cycle_type") s.params.cycle_type = (int)value, cycle = true;
        else if (k == "use_graph") s.use_graph = value != 0.0;
        else if (k == "loop_mode") s.loop_mode = (int)value;
        else if (k == "profile") s.profile = value != 0.0;
        else if (k == "trace") s.trace = value != 0.0;
        else if (k == "pdl") s.use_pdl = value != 0.0, cycle = true;
        else if (k == "fuse_norm") s.fuse_norm = value != 0.0, cycle = true;
        else if (k == "fuse_stop") s.fuse_stop = value != 0.0, cycle = true;
        else if (k == "fp32_refine") s.fp32_refine = value != 0.0, cycle = true;
        else if (k == "l2_hints") s.l2_hints = value != 0.0, cycle = true;
        else if (k == "tail_rows") s.tail_rows = (int)value, cycle = true;
        else if (k == "cluster_tail_rows") s.cluster_tail_rows = (int)value, cycle = true;
        else if (k == "dist_graph") s.dist_graph = value != 0.0, cycle = true;
        else if (k == "kernel_path") s.kernel_path = (int)value, hierarchy = true;
        else if (k == "xfer_threads") s.xfer_threads = (int)value;
        else if (k == "p2p") s.p2p = value != 0.0, hierarchy = true;
        else if (k == "p2p_fuse") s.p2p_fuse = value != 0.0, cycle = true;
        else if (k == "dist_shard_setup") s.dist_shard_setup = (int)value, hierarchy = true;
        else if (k == "dist_window") s.dist_window = value != 0.0, hierarchy = true;
        else if (k == "dist_skip_exchange") s.dist_skip_exchange = value != 0.0, cycle = true;
        else if (k == "spgemm_plan") s.spgemm_plan = value != 0.0, hierarchy = true;
        else if (k == "coarse_dataflow") s.coarse_dataflow = value != 0.0, cycle = true;
        else if (k == "diff_form") s.diff_form = value != 0.0, hierarchy = true;
        else if (k == "krylov_patience") s.krylov_patience = (int)value;
        else if (k == "krylov") {
            require(value == 0.0 || value == 1.0 || value == 2.0, "krylov must be 0 (cycle loop), 1 (CG preconditioned with the cycle) or 2 (plain CG)");
            s.krylov = (int)value;
        }
        else throw std::invalid_argument("unknown option: " + k);
        require(s.params.pre_iters >= 0 && s.params.post_iters >= 0 && s.params.pre_iters <= 16 && s.params.post_iters <= 16, "sweep counts must be 0..16");
        require(s.params.smoother == GMG_SMOOTHER_JACOBI || s.params.smoother == GMG_SMOOTHER_CHEBYSHEV, "unknown smoother");
        require(s.params.cheb_alpha > 1.0, "cheb_alpha must be > 1");
        require(s.xfer_threads >= -1 && s.xfer_threads <= 64, "xfer_threads must be -1 (auto), 0 (off) or 1..64");
        require(s.staged_lanes == 0 || s.staged_lanes == 1 || s.staged_lanes == 2 || s.staged_lanes == 4 || s.staged_lanes == 8, "lanes must be 0, 1, 2, 4 or 8");
        require(s.staged_lanes_r == 0 || s.staged_lanes_r == 1 || s.staged_lanes_r == 2 || s.staged_lanes_r == 4 || s.staged_lanes_r == 8, "lanes_r must be 0, 1, 2, 4 or 8");
        require(s.params.max_iter >= 1, "max_iter must be >= 1");
        require(s.params.cycle_type >= 0 && s.params.cycle_type <= 2, "cycle_type must be 0 (V), 1 (F) or 2 (W)");
        require(s.params.stopping_criteria >= 0 && s.params.stopping_criteria <= 3, "stopping_criteria must be 0..3");
        } catch (...) {
            s.params = saved_params;
            s.xfer_threads = saved_xfer, s.staged_lanes = saved_lanes, s.staged_lanes_r = saved_lanes_r;
            throw;
        }
        if (s.engine && hierarchy) s.engine->invalidate_hierarchy();
        if (s.engine && cycle) s.engine->invalidate_cycle();
    });
}

int gmg_get_option(gmg_handle h, const char* key, double* value) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(key && value, "null argument");
        const SolverState& s = h->s;
        const std::string k(key);
        if (k == "tolerance") *value = s.params.tolerance;
        else if (k == "max_iter") *value = s.params.max_iter;
        else if (k == "stopping_criteria") *value = s.params.stopping_criteria;
        else if (k == "pre_iters") *value = s.params.pre_iters;
        else if (k == "post_iters") *value = s.params.post_iters;
        else if (k == "omega") *value = s.params.omega;
        else if (k == "smoother") *value = s.params.smoother;
        else if (k == "cheb_alpha") *value = s.params.cheb_alpha;
        else if (k == "lanes") *value = s.staged_lanes;
        else if (k == "lanes_r") *value = s.staged_lanes_r;
        else if (k == "restrict_path") *value = s.restrict_path;
        else if (k == "cycle_type") *value = s.params.cycle_type;
        else if (k == "use_graph") *value = s.use_graph;
        else if (k == "loop_mode") *value = s.loop_mode;
        else if (k == "profile") *value = s.profile;
        else if (k == "trace") *value = s.trace;
        else if (k == "pdl") *value = s.use_pdl;
        else if (k == "fuse_norm") *value = s.fuse_norm;
        else if (k == "fuse_stop") *value = s.fuse_stop;
        else if (k == "fp32_refine") *value = s.fp32_refine;
        else if (k == "l2_hints") *value = s.l2_hints;
        else if (k == "tail_rows") *value = s.tail_rows;
        else if (k == "cluster_tail_rows") *value = s.cluster_tail_rows;
        else if (k == "dist_graph") *value = s.dist_graph;
        else if (k == "kernel_path") *value = s.kernel_path;
        else if (k == "xfer_threads") *value = s.xfer_threads;
        else if (k == "p2p") *value = s.p2p;
        else if (k == "p2p_fuse") *value = s.p2p_fuse;
        else if (k == "dist_shard_setup") *value = s.dist_shard_setup;
        else if (k == "dist_window") *value = s.dist_window;
        else if (k == "dist_skip_exchange") *value = s.dist_skip_exchange;
        else if (k == "spgemm_plan") *value = s.spgemm_plan;
        else if (k == "coarse_dataflow") *value = s.coarse_dataflow;
        else if (k == "diff_form") *value = s.diff_form;
        else if (k == "krylov") *value = s.krylov;
        else if (k == "krylov_patience") *value = s.krylov_patience;
        else throw std::invalid_argument("unknown option: " + k);
    });
}

// ---- hierarchy ------------------------------------------------------------------------
int gmg_num_levels(gmg_handle h, int32_t* n_prolongations) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(n_prolongations != nullptr, "null argument");
        *n_prolongations = (int32_t)h->s.hier.U.size();
    });
}

int gmg_prolongation_shape(gmg_handle h, int32_t level, int64_t* rows, int64_t* cols, int64_t* nnz) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(level >= 0 && level < (int)h->s.hier.U.size(), "level out of range");
        const gmg::HostCsr& u = h->s.hier.U[level];
        *rows = u.rows, *cols = u.cols, *nnz = u.nnz();
    });
}

int gmg_get_prolongation(gmg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(level >= 0 && level < (int)h->s.hier.U.size(), "level out of range");
        const gmg::HostCsr& u = h->s.hier.U[level];
        std::copy(u.indptr.begin(), u.indptr.end(), indptr);
        std::copy(u.indices.begin(), u.indices.end(), indices);
        std::copy(u.data.begin(), u.data.end(), data);
    });
}

int gmg_clear_prolongations(gmg_handle h) {
    if (!h) return 1;
    return guarded(h, [&] {
        SolverState& s = h->s;
        s.hier.U.clear();
        s.hier.dof.assign(1, s.n);
        if (s.engine) s.engine->invalidate_hierarchy();
    });
}

int gmg_set_prolongation(gmg_handle h, int32_t level, int64_t rows, int64_t cols, const int32_t* indptr,
                         const int32_t* indices, const double* data) {
    if (!h) return 1;
    return guarded(h, [&] {
        SolverState& s = h->s;
        require(level == (int)s.hier.U.size(), "prolongations must be set in order, after gmg_clear_prolongations");
        require(rows == s.hier.dof.back(), "prolongation has the wrong number of rows for this level");
        require(cols > 0 && indptr && indices && data, "invalid prolongation");
        require(indptr[0] == 0, "indptr must start at 0");
        gmg::HostCsr u;
        u.rows = rows, u.cols = cols;
        u.indptr.assign(indptr, indptr + rows + 1);
        const int64_t nnz = indptr[rows];
        for (int64_t r = 0; r < rows; ++r) require(indptr[r + 1] >= indptr[r], "indptr must be non-decreasing");
        for (int64_t q = 0; q < nnz; ++q) require(indices[q] >= 0 && indices[q] < cols, "prolongation column index out of range");
        u.indices.assign(indices, indices + nnz);
        u.data.assign(data, data + nnz);
        gmg::sort_rows_sum_duplicates(u);
        s.hier.U.push_back(std::move(u));
        s.hier.dof.push_back(cols);
        if (s.engine) s.engine->invalidate_hierarchy();
    });
}

#define GMG_LEVEL_GETTER(NAME, FIELD, TYPE)                                                        \
    int NAME(gmg_handle h, int32_t level, TYPE* out, int64_t* count) {                             \
        if (!h) return 1;                                                                          \
        return guarded(h, [&] {                                                                    \
            require(count != nullptr, "null argument");                                            \
            require(level >= 0 && level < (int)h->s.hier.FIELD.size(), "level out of range (some arrays exist only with debug=True)"); \
            copy_out(h->s.hier.FIELD[level], out, count);                                          \
        });                                                                                        \
    }
GMG_LEVEL_GETTER(gmg_get_samples, samples, int32_t)
GMG_LEVEL_GETTER(gmg_get_nearest_source, nearest_source, int32_t)
GMG_LEVEL_GETTER(gmg_get_level_points, level_points, double)
GMG_LEVEL_GETTER(gmg_get_all_triangles, all_triangles, int32_t)
GMG_LEVEL_GETTER(gmg_get_notrimap, no_tri_found, int32_t)
#undef GMG_LEVEL_GETTER

// ---- solve ------------------------------------------------------------------------------
int gmg_stage_system(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices,
                     const double* a_data, const double* rhs, int32_t K) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(a_indptr && a_indices && a_data && rhs, "null argument");
        engine(h).stage_system(n, a_indptr, a_indices, a_data, rhs, K, true);
    });
}

int gmg_solve_staged(gmg_handle h) {
    if (!h) return 1;
    return guarded(h, [&] { engine(h).solve_staged(); });
}

int gmg_fetch_solution(gmg_handle h, double* x_out) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(x_out != nullptr, "null argument");
        engine(h).fetch_solution(x_out);
    });
}

int gmg_solve(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_data,
              const double* rhs, double* x_out, int32_t K) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(a_indptr && a_indices && a_data && rhs && x_out, "null argument");
        gmg::EngineBase& e = engine(h);
        e.stage_system(n, a_indptr, a_indices, a_data, rhs, K, false);
        e.solve_staged();
        e.fetch_solution(x_out);
    });
}

// direct_solve (core.cpp:74-78 -> multigrid_solver.cpp:1287-1321: Eigen::SimplicialLLT of the whole system).
// Up to 16384 rows the system is factorised on the device by the dense Cholesky of the coarsest level (a twin
// handle without hierarchy). Beyond that there is no sparse factorisation on the device: the system is solved
// to the fp64 rounding floor by conjugate gradients preconditioned with the V-cycle, until the residual stops
// decreasing (at most 100 iterations) — the accuracy of a backward-stable factorisation, not its algorithm.
int gmg_direct_solve(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_data,
                     const double* rhs, double* x_out, int32_t K) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(a_indptr && a_indices && a_data && rhs && x_out, "null argument");
        SolverState& s = h->s;
        require(n == s.n, "lhs has a different number of rows than the point set of the constructor");
        auto& tm = s.solver_timing;
        if (n <= 16384) {
            if (!s.direct_helper) {
                s.direct_helper.reset(new gmg_solver());
                SolverState& d = s.direct_helper->s;
                d.params = s.params;
                d.params.dtype = GMG_DTYPE_F64, d.params.build_hierarchy = 0, d.params.max_iter = 1;
                d.n = n;
                d.mass_diag = s.mass_diag;
                d.hier.dof.push_back(n);
                d.loop_mode = 0;  // one "cycle" (the dense solve) per call: no device-side loop needed
            }
            SolverState& d = s.direct_helper->s;
            d.params.stopping_criteria = s.params.stopping_criteria;
            gmg::EngineBase& e = engine(s.direct_helper.get());
            try {
                e.stage_system(n, a_indptr, a_indices, a_data, rhs, K, false);
                e.solve_staged();
                e.fetch_solution(x_out);
            } catch (const std::exception& ex) {
                throw std::runtime_error(std::string("direct_solve: ") + ex.what());
            }
            tm["direct_factor"] = d.solver_timing["reduction"] + d.solver_timing["coarsest_solve"];
            tm["direct_solve"] = d.solver_timing["cycles"];
            tm["direct_residual"] = d.solver_timing["residue"];
        } else {
            require(!s.hier.U.empty(), "direct_solve of more than 16384 rows needs the hierarchy (it is solved iteratively to the rounding floor)");
            const gmg_params saved = s.params;
            const int saved_krylov = s.krylov, saved_patience = s.krylov_patience;
            s.params.tolerance = 1e-14, s.params.max_iter = 100;
            s.krylov = 1, s.krylov_patience = 3;
            gmg::EngineBase& e = engine(h);
            auto put_back = [&] { s.params = saved, s.krylov = saved_krylov, s.krylov_patience = saved_patience; };
            try {
                for (int k0 = 0; k0 < K; k0 += 4) {  // the conjugate-gradient wrapper takes up to 4 columns at a time
                    const int kt = std::min(4, K - k0);
                    std::vector<double> b((size_t)n * kt), x((size_t)n * kt);
                    for (int64_t i = 0; i < n; ++i)
                        for (int k = 0; k < kt; ++k) b[(size_t)i * kt + k] = rhs[(size_t)i * K + k0 + k];
                    e.stage_system(n, a_indptr, a_indices, a_data, b.data(), kt, false);
                    e.solve_staged();
                    e.fetch_solution(x.data());
                    for (int64_t i = 0; i < n; ++i)
                        for (int k = 0; k < kt; ++k) x_out[(size_t)i * K + k0 + k] = x[(size_t)i * kt + k];
                    tm["direct_residual"] = k0 == 0 ? tm["residue"] : std::max(tm["direct_residual"], tm["residue"]);
                }
            } catch (...) {
                put_back();
                throw;
            }
            put_back();
            tm["direct_factor"] = tm["reduction"] + tm["coarsest_solve"];
            tm["direct_solve"] = tm["cycles"];
        }
    });
}

int gmg_residual(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_data,
                 const double* rhs, const double* x, int32_t K, int32_t type, double* out) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(a_indptr && a_indices && a_data && rhs && x && out, "null argument");
        *out = engine(h).residual(n, a_indptr, a_indices, a_data, rhs, x, K, type);
    });
}

// ---- device-resident systems and mesh assembly ------------------------------------------------
int gmg_update_values_device(gmg_handle h, const double* d_a_data, const double* d_rhs, int32_t K) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(d_a_data && d_rhs, "null argument");
        engine(h).update_values_device(d_a_data, d_rhs, K);
    });
}

int gmg_fetch_solution_device(gmg_handle h, double* d_x_out) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(d_x_out != nullptr, "null argument");
        engine(h).fetch_solution_device(d_x_out);
    });
}

int gmg_mesh_attach(gmg_handle h, int64_t nf, const int32_t* faces) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(faces != nullptr && nf > 0, "faces are required");
        engine(h).mesh_attach(nf, faces);
    });
}

int gmg_mesh_set_positions(gmg_handle h, const double* pos) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(pos != nullptr, "null argument");
        engine(h).mesh_set_positions(pos);
    });
}

int gmg_mesh_stiffness(gmg_handle h) {
    if (!h) return 1;
    return guarded(h, [&] { engine(h).mesh_stiffness(); });
}

int gmg_mesh_mass(gmg_handle h, int32_t type) {
    if (!h) return 1;
    return guarded(h, [&] { engine(h).mesh_mass(type); });
}

int gmg_mesh_system(gmg_handle h, double alpha, double beta, const double* y, int32_t K) {
    if (!h) return 1;
    return guarded(h, [&] { engine(h).mesh_system(alpha, beta, y, K); });
}

int gmg_mesh_flow(gmg_handle h, double tau, int32_t mass_type, int32_t steps) {
    if (!h) return 1;
    return guarded(h, [&] { engine(h).mesh_flow(tau, mass_type, steps); });
}

int gmg_mesh_get(gmg_handle h, int32_t which, double* out) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(out != nullptr, "null argument");
        engine(h).mesh_get(which, out);
    });
}

int gmg_mesh_pattern(int64_t n, int64_t nf, const int32_t* faces, int32_t* indptr, int32_t* indices, int64_t* nnz) {
    return guarded(nullptr, [&] {
        require(faces && indptr && nnz, "null argument");
        const gmg::MeshTopology t = gmg::build_mesh_topology(n, nf, faces);
        *nnz = t.pattern.nnz();
        std::copy(t.pattern.indptr.begin(), t.pattern.indptr.end(), indptr);
        if (indices) std::copy(t.pattern.indices.begin(), t.pattern.indices.end(), indices);
    });
}

// ---- multi-GPU ------------------------------------------------------------------------------
int gmg_dist_configure(gmg_handle h, int32_t rank, int32_t world, int64_t replicate_rows) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(world >= 1 && rank >= 0 && rank < world, "rank must be in [0, world)");
        SolverState& s = h->s;
        s.dist = gmg::DistLayout();
        s.dist.rank = rank, s.dist.world = world;
        if (replicate_rows >= 0) s.replicate_rows = replicate_rows;
        if (s.engine) s.engine->invalidate_hierarchy();
    });
}

int gmg_dist_unique_id(void* id_out, int64_t capacity, int64_t* size) {
    return guarded(nullptr, [&] {
        require(id_out && size, "null argument");
        require(capacity >= (int64_t)sizeof(ncclUniqueId), "buffer too small for an NCCL unique id");
        ncclUniqueId id;
        GMG_NCCL(gmg::nccl().GetUniqueId(&id));
        std::memcpy(id_out, &id, sizeof id);
        *size = (int64_t)sizeof id;
    });
}

int gmg_dist_init(gmg_handle h, const void* id, int64_t size) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(id && size == (int64_t)sizeof(ncclUniqueId), "expected the bytes of an NCCL unique id");
        engine(h).dist_init(id);
    });
}

int gmg_dist_layout(gmg_handle h, int64_t n, const int32_t* a_indptr, const int32_t* a_indices) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(a_indptr && a_indices, "null argument");
        require(n == h->s.n, "lhs has a different number of rows than the point set of the constructor");
        gmg::compute_level_patterns(h->s, n, a_indptr, a_indices);
        gmg::compute_dist_layout(h->s);
        gmg::compute_level0_windows(h->s);
    });
}

int gmg_level_pattern(gmg_handle h, int32_t level, int32_t* indptr, int32_t* indices, int64_t* rows, int64_t* nnz) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(rows && nnz, "null argument");
        require(level >= 0 && level < (int)h->s.a_pat.size(), "level out of range (call gmg_dist_layout or stage a system first)");
        const gmg::HostCsr& m = h->s.a_pat[level];
        *rows = m.rows, *nnz = m.nnz();
        if (indptr) std::copy(m.indptr.begin(), m.indptr.end(), indptr);
        if (indices) std::copy(m.indices.begin(), m.indices.end(), indices);
    });
}

int gmg_dist_windows(gmg_handle h, int32_t which, int64_t* ranges, int64_t* count, int32_t* enabled) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(which >= 0 && which <= 3 && count && enabled, "bad argument");
        const gmg::Level0Windows& w = h->s.win0;
        const gmg::RowRanges& r = which == 0 ? w.a_rows : which == 1 ? w.p_rows : which == 2 ? w.c_rows : w.rhs_rows;
        *enabled = w.on ? 1 : 0;
        if (ranges) {
            require(*count >= (int64_t)r.size(), "output buffer too small");
            for (size_t i = 0; i < r.size(); ++i) ranges[2 * i] = r[i].first, ranges[2 * i + 1] = r[i].second;
        }
        *count = (int64_t)r.size();
    });
}

int gmg_dist_ranges(gmg_handle h, int32_t level, int64_t* ranges, int32_t* replicated) {
    if (!h) return 1;
    return guarded(h, [&] {
        const gmg::DistLayout& d = h->s.dist;
        require(level >= 0 && level < (int)d.ranges.size(), "level out of range (call gmg_dist_layout or stage a system first)");
        require(ranges && replicated, "null argument");
        std::copy(d.ranges[level].begin(), d.ranges[level].end(), ranges);
        *replicated = d.sharded(level) ? 0 : 1;
    });
}

int gmg_dist_halo(gmg_handle h, int32_t op, int32_t level, int32_t peer, int32_t* send, int64_t* n_send, int32_t* recv,
                  int64_t* n_recv) {
    if (!h) return 1;
    return guarded(h, [&] {
        const gmg::DistLayout& d = h->s.dist;
        require(op >= 0 && op < 3 && n_send && n_recv, "bad argument");
        require(level >= 0 && level < (int)d.halo[op].size() && peer >= 0 && peer < d.world, "level or peer out of range");
        const gmg::HaloLists& hl = d.halo[op][level];
        static const std::vector<int> none;
        const std::vector<int>& s = hl.send.empty() ? none : hl.send[peer];
        const std::vector<int>& r = hl.recv.empty() ? none : hl.recv[peer];
        if (send) {
            require(*n_send >= (int64_t)s.size(), "send buffer too small");
            std::copy(s.begin(), s.end(), send);
        }
        if (recv) {
            require(*n_recv >= (int64_t)r.size(), "recv buffer too small");
            std::copy(r.begin(), r.end(), recv);
        }
        *n_send = (int64_t)s.size(), *n_recv = (int64_t)r.size();
    });
}

// ---- timing -----------------------------------------------------------------------------
int gmg_timing_keys(gmg_handle h, int32_t which, char* buf, int64_t buflen) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(buf && buflen > 0, "null argument");
        const auto& m = which == 0 ? h->s.hier.timing : which == 2 ? h->s.transfer_timing : h->s.solver_timing;
        std::string joined;
        for (const auto& kv : m) {
            if (!joined.empty()) joined += ',';
            joined += kv.first;
        }
        require((int64_t)joined.size() + 1 <= buflen, "buffer too small");
        std::memcpy(buf, joined.c_str(), joined.size() + 1);
    });
}

int gmg_get_timing(gmg_handle h, int32_t which, const char* key, double* out) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(key && out, "null argument");
        const auto& m = which == 0 ? h->s.hier.timing : which == 2 ? h->s.transfer_timing : h->s.solver_timing;
        auto it = m.find(key);
        require(it != m.end(), "unknown timing key");
        *out = it->second;
    });
}

int gmg_get_convergence(gmg_handle h, double* t_ms, double* residue, int32_t* count) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(count != nullptr, "null argument");
        const auto& c = h->s.convergence;
        if (t_ms && residue) {
            require(*count >= (int)c.size(), "output buffer too small");
            for (size_t i = 0; i < c.size(); ++i) t_ms[i] = c[i].first, residue[i] = c[i].second;
        }
        *count = (int32_t)c.size();
    });
}

// ---- measurement ------------------------------------------------------------------------
int gmg_level_info(gmg_handle h, int32_t level, int64_t* rows, int64_t* nnz_a, int64_t* nnz_u) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(rows && nnz_a && nnz_u, "null argument");
        require(h->s.engine && h->s.engine->level_info(level, rows, nnz_a, nnz_u), "no such level staged on the device");
    });
}

int gmg_level_op(gmg_handle h, int32_t kind, int32_t level, const double* a, const double* b, double* out, int32_t sweeps) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(out != nullptr, "null argument");
        require(h->s.engine != nullptr, "gmg_level_op needs a staged system (gmg_stage_system)");
        h->s.engine->level_op(kind, level, a, b, out, sweeps);
    });
}

int gmg_time_op(gmg_handle h, int32_t kind, int32_t level, int32_t reps, double* us_per_launch) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(us_per_launch != nullptr, "null argument");
        require(h->s.engine != nullptr, "gmg_time_op needs a staged system (gmg_stage_system)");
        *us_per_launch = h->s.engine->time_op(kind, level, reps);
    });
}

int gmg_get_smoother_weights(gmg_handle h, int32_t level, double* rho, double* pre, double* post) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(rho && pre && post, "null argument");
        require(h->s.engine != nullptr, "no system staged on the device");
        h->s.engine->smoother_weights(level, rho, pre, post);
    });
}

int gmg_get_level_matrix(gmg_handle h, int32_t level, int32_t* indptr, int32_t* indices, double* data) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(indptr && indices && data, "null argument");
        require(h->s.engine != nullptr, "no system staged on the device");
        h->s.engine->get_level_matrix(level, indptr, indices, data);
    });
}

int gmg_kernel_profile(gmg_handle h, int32_t kind, int32_t level, double* total_ms, int64_t* launches) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(total_ms && launches, "null argument");
        *total_ms = 0.0, *launches = 0;
        if (h->s.engine) h->s.engine->kernel_profile(kind, level, total_ms, launches);
    });
}

int gmg_reset_kernel_profile(gmg_handle h) {
    if (!h) return 1;
    return guarded(h, [&] {
        if (h->s.engine) h->s.engine->reset_kernel_profile();
    });
}

int gmg_get_trace(gmg_handle h, uint64_t* t_ns, uint64_t* tags, int64_t* count) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(count != nullptr, "null argument");
        const auto& log = h->s.trace_log;
        const int64_t n = (int64_t)log.size() / 2;
        if (t_ns && tags) {
            require(*count >= n, "output buffer too small");
            for (int64_t i = 0; i < n; ++i) t_ns[i] = log[2 * i], tags[i] = log[2 * i + 1];
        }
        *count = n;
    });
}

int gmg_last_launch_count(gmg_handle h, int64_t* launches) {
    if (!h) return 1;
    return guarded(h, [&] {
        require(launches != nullptr, "null argument");
        *launches = h->s.last_launches;
    });
}

}  // extern "C"
