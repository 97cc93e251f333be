// Device engine: stages the operators once, recomputes the Galerkin chain and the coarse
// factor per solve, and runs the V-cycle loop
//     do { V-cycle; residualCheck } while (residue > tol && iter < max_iter)
// (reference multigrid_solver.cpp:1367-1449, 1059-1088) as a fixed list of kernel launches
// that is captured once into a CUDA graph and replayed per cycle — or wrapped in a device-side
// while-node so the whole loop is one launch.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "dense_coarse.h"
#include "host_xfer.h"
#include "mesh_assembly.h"
#include "nccl_dl.h"
#include "pcg_kernels.h"
#include "peer_exchange.h"
#include "solver.h"
#include "sparse_kernels.h"
#include "cluster_tail.cuh"
#include "tail_kernel.cuh"

namespace gmg {
namespace {

enum OpKind { OP_JACOBI = 0, OP_RESIDUAL = 1, OP_RESTRICT = 2, OP_PROLONG = 3, OP_NORM = 4, OP_COARSE = 5, OP_ZERO = 6, OP_TAIL = 7, OP_HALO = 8, OP_ALLGATHER = 9, OP_REFINE = 10, OP_KINDS = 11 };
constexpr int kMaxLevels = 16;

// CSR matrix on the device. Setup arithmetic (Galerkin products, factorisation) is always
// fp64 (`v64`); when the smoother runs in fp32 a cast copy `v32` is what the cycle reads.
template <typename T>
struct DevMat {
    int rows = 0, cols = 0;
    int64_t nnz = 0;
    DeviceBuffer<int> indptr, indices, rowidx;
    DeviceBuffer<long long> pair_off;   // product plan of the Galerkin step that fills this matrix
    DeviceBuffer<int2> pairs;           //   (sparse_kernels.h, build_spgemm_plan); empty: search per solve
    bool planned = false;
    int l2_hint = 0;  // L2 eviction priority of this operator's slabs inside the cycle (sparse_kernels.cuh)
    // multi-GPU: the stored entries this rank computes in the Galerkin product ([e0, e1); all by default)
    // and, for the level operators, every rank's share (entry offsets, world + 1) for the all-gather
    int64_t e0 = 0, e1 = -1;
    std::vector<int64_t> share;
    DeviceBuffer<int4> tiles;
    DeviceBuffer<double> v64;
    DeviceBuffer<float> v32;
    DeviceBuffer<double> vd;  // finest level, option diff_form: v64 with the row sum in place of the diagonal (SpmvArgs::diff)
    bool diff = false;
    SpmvPlan plan;

    // Multi-GPU: the finest level's operators are stored by ROW SEGMENTS — a rank holds only the rows it works on
    // (its own range plus what its share of the Galerkin product reads; on a periodic mesh that is two or three
    // disjoint row ranges). The entry arrays (indices, values, rowidx) hold the entries of the stored rows, segment
    // after segment, and the device row pointer is REBASED onto that storage: indptr[r] is the offset of row r's
    // first entry in the local arrays for a stored row; rows that are not stored are empty. Kernels, tile
    // descriptors and product plans work from the row pointer, so they are the same as for a fully stored
    // matrix; row indices (and so every vector) stay global. Segments start at multiples of 4 entries (the
    // 16-byte alignment of the bulk copies); the padding is marked with column -1.
    struct Segment {
        int r0, r1;        // rows [r0, r1)
        int64_t g0, l0;    // first entry in the caller's (global) arrays / in the local arrays
        int64_t len;
    };
    std::vector<Segment> segs;
    std::vector<int> indptr_local;   // host copy of the rebased row pointer (empty: the global one is used)
    int64_t stored = 0;
    int* colp() const { return indices.ptr; }
    int* rowidxp() const { return rowidx.ptr; }
    double* v64p() const { return v64.ptr; }
    double* vdp() const { return vd.ptr; }
    float* v32p() const { return v32.ptr; }
    bool windowed() const { return !indptr_local.empty(); }
    // row pointer the tile planner and the entry ranges of this matrix are taken from
    const std::vector<int>& indptr_host(const HostCsr& m) const { return windowed() ? indptr_local : m.indptr; }

    const T* vals() const;
    // rows: ascending, disjoint, non-adjacent row ranges to store (nullptr: every row)
    void upload_pattern(const HostCsr& m, cudaStream_t s, const RowRanges* rows_kept = nullptr) {
        rows = (int)m.rows, cols = (int)m.cols, nnz = m.nnz();
        segs.clear(), indptr_local.clear();
        if (rows_kept && !(rows_kept->size() == 1 && (*rows_kept)[0].first == 0 && (*rows_kept)[0].second == m.rows)) {
            indptr_local.assign((size_t)rows + 1, 0);
            int64_t off = 0;
            size_t next = 0;
            for (int r = 0; r < rows; ++r) {
                const bool starts = next < rows_kept->size() && (*rows_kept)[next].first == r;
                if (starts) {
                    off = (off + 3) & ~(int64_t)3;
                    const int r1 = (int)(*rows_kept)[next].second;
                    segs.push_back({r, r1, (int64_t)m.indptr[r], off, (int64_t)m.indptr[r1] - m.indptr[r]});
                    ++next;
                }
                indptr_local[r] = (int)off;
                const bool kept = !segs.empty() && r < segs.back().r1;
                if (kept) off += m.indptr[r + 1] - m.indptr[r];
            }
            indptr_local[rows] = (int)off;
            stored = off;
            indptr.upload(indptr_local, s, 8);
            indices.ensure(stored, 8);
            GMG_CUDA(cudaMemsetAsync(indices.ptr, 0xFF, (stored + 8) * sizeof(int), s));  // padding: column -1
            for (const Segment& g : segs)
                if (g.len) GMG_CUDA(cudaMemcpyAsync(indices.ptr + g.l0, m.indices.data() + g.g0, g.len * sizeof(int), cudaMemcpyHostToDevice, s));
        } else {
            segs.push_back({0, rows, 0, 0, nnz});
            stored = nnz;
            indptr.upload(m.indptr, s, 8);
            indices.upload(m.indices.data(), m.indices.size(), s, 8);
        }
        v64.ensure(stored, 8);
        GMG_CUDA(cudaMemsetAsync(v64.ptr, 0, (stored + 8) * sizeof(double), s));
        if (sizeof(T) == 4) {
            v32.ensure(stored, 8);
            GMG_CUDA(cudaMemsetAsync(v32.ptr, 0, (stored + 8) * sizeof(float), s));
        }
    }
    void upload_values(const double* host, cudaStream_t s) {  // host: the caller's (global) value array
        for (const Segment& g : segs)
            if (g.len) GMG_CUDA(cudaMemcpyAsync(v64.ptr + g.l0, host + g.g0, g.len * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    void refresh_cast(cudaStream_t s) {
        if (sizeof(T) == 4) launch_cast_f64_f32(v64.ptr, v32.ptr, (size_t)stored, s);
    }
    void make_rowidx(cudaStream_t s) {
        rowidx.ensure(std::max<int64_t>(stored, 1));
        GMG_CUDA(cudaMemsetAsync(rowidx.ptr, 0, std::max<int64_t>(stored, 1) * sizeof(int), s));  // padding entries: row 0
        for (const Segment& g : segs) launch_expand_rows(g.r1, indptr.ptr, rowidx.ptr, s, g.r0);
    }
    // Choose the kernel path and build the row tiles for the staged one.
    // `early_rows` (multi-GPU, may be null): rows that are pushed to peers or gather halo entries;
    // their tiles are moved to the front of the walk so the exchange can be signalled early.
    void make_plan(const std::vector<int>& indptr_h, int prefer_path, int staged_lanes, cudaStream_t s, int row_begin = 0,
                   int row_end = -1, const std::vector<char>* early_rows = nullptr) {
        if (row_end < 0) row_end = rows;
        const double avg = rows ? (double)nnz / rows : 0.0;
        int lanes = 1;
        while (lanes < 32 && lanes < avg) lanes *= 2;
        plan = SpmvPlan();
        plan.lanes = lanes;
        plan.path = 1;
        plan.row_begin = row_begin, plan.row_end = row_end;
        if (prefer_path == 0 && row_end > row_begin) {
            // threads per row of the staged kernel: enough rows per tile to keep the CTA busy, few
            // enough entries per tile that many CTAs fit one SM's shared memory
            int sl = staged_lanes;
            if (sl == 0) sl = avg <= 10.0 ? 1 : avg <= 26.0 ? 2 : avg <= 60.0 ? 4 : 8;  // measured: tools/spmv_lab.cu
            // stage budget: kStagedStages stages, at least two resident CTAs per SM
            const int stage_rows = kStagedThreads / sl;
            const size_t per_stage = (staged_smem_limit() / 2 - 1280) / kStagedStages - 16 - (size_t)(stage_rows + 8) * sizeof(int);
            const int cap = (int)(per_stage / (sizeof(T) + sizeof(int))) & ~3;
            int worst = 0;
            std::vector<int> t = plan_row_tiles(indptr_h, stage_rows, cap, &worst, row_begin, row_end);
            if (worst <= cap) {
                std::vector<int4> desc, late;
                desc.reserve(t.size());
                for (size_t i = 0; i + 1 < t.size(); ++i) {
                    const int4 d4 = make_int4(t[i], t[i + 1], indptr_h[t[i]] & ~3, (indptr_h[t[i + 1]] + 3) & ~3);
                    bool early = false;
                    if (early_rows)
                        for (int r = t[i]; r < t[i + 1] && !early; ++r) early = (*early_rows)[r] != 0;
                    (early_rows && !early ? late : desc).push_back(d4);
                }
                plan.n_early = early_rows ? (int)desc.size() : 0;
                desc.insert(desc.end(), late.begin(), late.end());
                tiles.upload(desc, s);
                plan.path = 0;
                plan.staged_lanes = sl;
                plan.n_tiles = (int)desc.size();
                plan.stage_elems = std::max((worst + 3) & ~3, 4);
                plan.stage_rows = stage_rows;
                plan.tile_desc = tiles.ptr;
                GMG_CUDA(cudaStreamSynchronize(s));  // `desc` is a local
            }
        }
    }
};
template <> const double* DevMat<double>::vals() const { return v64p(); }
template <> const float* DevMat<float>::vals() const { return v32p(); }

struct ProfileSlot {
    double ms = 0.0;
    int64_t launches = 0;
};

template <typename T>
class Engine : public EngineBase {
public:
    explicit Engine(SolverState* st) : st_(st) {
        int count = 0;
        GMG_CUDA(cudaGetDeviceCount(&count));
        if (count <= 0) throw CudaError("no CUDA device available (this library has no CPU fallback)");
        if (st->params.device < 0 || st->params.device >= count) throw std::invalid_argument("invalid CUDA device ordinal");
        GMG_CUDA(cudaSetDevice(st->params.device));
        GMG_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        GMG_CUDA(cudaStreamCreateWithFlags(&stream2_, cudaStreamNonBlocking));
        GMG_CUDA(cudaEventCreateWithFlags(&rhs_ready_, cudaEventDisableTiming));
        for (auto& e : ev_) GMG_CUDA(cudaEventCreate(&e));
        ctl_.ensure(3);
        GMG_CUDA(cudaMemsetAsync(ctl_.ptr, 0, 3 * sizeof(CycleControl), stream_));
        partials_.ensure(kNormChunkStride * kMaxNormChunks);
        rho_.ensure(kMaxLevels);
        tail_bar_.ensure(8);
        GMG_CUDA(cudaMemsetAsync(tail_bar_.ptr, 0, 8 * sizeof(unsigned), stream_));
        weights_.ensure((size_t)kMaxLevels * 2 * kMaxSweeps);
        weights64_.ensure((size_t)kMaxLevels * 2 * kMaxSweeps);
        GMG_CUDA(cudaMallocHost((void**)&ctl_host_, sizeof(CycleControl)));
        std::vector<double> minv(st->mass_diag.size());
        for (size_t i = 0; i < minv.size(); ++i) minv[i] = st->mass_diag[i] != 0.0 ? 1.0 / st->mass_diag[i] : 0.0;  // igl::invert_diag
        mass_.upload(st->mass_diag, stream_);
        minv_.upload(minv, stream_);
        GMG_CUDA(cudaStreamSynchronize(stream_));
    }

    ~Engine() override {
        cudaSetDevice(st_->params.device);
        xfer_.reset();
        drop_graphs();
        try {
            release_peer_arena(false);
        } catch (...) {
        }
        if (comm_) nccl().CommDestroy(comm_);
        for (auto& e : ev_) cudaEventDestroy(e);
        for (auto& e : prof_events_) cudaEventDestroy(e);
        if (ctl_host_) cudaFreeHost(ctl_host_);
        if (rhs_ready_) cudaEventDestroy(rhs_ready_);
        if (stream2_) cudaStreamDestroy(stream2_);
        if (stream_) cudaStreamDestroy(stream_);
    }

    void invalidate_hierarchy() override {
        hierarchy_ready_ = false;
        pattern_ready_ = false;
        invalidate_cycle();
    }
    void invalidate_cycle() override {
        cycle_dirty_ = true;
        kry_ops_dirty_ = true;
        numeric_ready_ = false;  // the smoother dampings are part of the numeric setup
    }

    // ------------------------------------------------------------------ staging
    void stage_system(int64_t n, const int* indptr, const int* indices, const double* data, const double* rhs,
                      int K, bool wait) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (n != st_->n) throw std::invalid_argument("lhs has a different number of rows than the point set of the constructor");
        if (K < 1 || K > kMaxRhsTile * kMaxNormChunks) throw std::invalid_argument("number of right-hand sides must be 1..32");
        if (indptr[0] != 0) throw std::invalid_argument("lhs indptr must start at 0");
        const int64_t nnz = indptr[n];
        const auto t0 = std::chrono::steady_clock::now();
        if (!hierarchy_ready_) pattern_ready_ = false;
        if (K != K_) {
            K_ = K;
            if (pattern_ready_) allocate_vectors();
            invalidate_cycle();
        }
        const bool same_shape = pattern_ready_ && !st_->a_pat.empty() && st_->a_pat[0].rows == n &&
                                (int64_t)st_->a_pat[0].indices.size() == nnz;
        HostTransfer* xf = transfer();
        bool same = false;
        if (same_shape && xf) {
            // the usual repeated solve: same pattern, new values. Values and rhs stream to HBM through
            // pinned chunks while the same worker threads compare the pattern with the staged one.
            std::vector<HostTransfer::Copy> copies = value_copies(data, rhs, nnz, n, K);
            // (multi-GPU, windowed storage: every rank checks the column indices of its own window; the ranks
            // then agree, so a change anywhere re-stages everywhere)
            const DevMat<T>& a0 = lv_[0].A;
            std::vector<HostTransfer::Compare> compares(1);
            compares[0].a = st_->a_pat[0].indptr.data(), compares[0].b = indptr, compares[0].bytes = (size_t)(n + 1) * sizeof(int);
            for (const auto& g : a0.segs) {
                HostTransfer::Compare c;
                c.a = st_->a_pat[0].indices.data() + g.g0, c.b = indices + g.g0, c.bytes = (size_t)g.len * sizeof(int);
                compares.push_back(c);
            }
            same = xf->upload_and_compare(copies, compares, stream_);
            if (window0_) same = all_ranks_agree(same);  // layout property: every rank takes this branch
        } else if (same_shape) {
            const DevMat<T>& a0 = lv_[0].A;
            same = std::memcmp(st_->a_pat[0].indptr.data(), indptr, (n + 1) * sizeof(int)) == 0;
            for (const auto& g : a0.segs)
                same = same && std::memcmp(st_->a_pat[0].indices.data() + g.g0, indices + g.g0, g.len * sizeof(int)) == 0;
            if (window0_) same = all_ranks_agree(same);  // layout property: every rank takes this branch
        }
        const bool uploaded = same && xf;
        if (!same) {
            GMG_CUDA(cudaStreamSynchronize(stream_));  // a speculative upload may still be in flight
            // a new pattern is checked once (repeated solves compare against the checked copy): the symbolic
            // products and the kernels index with these arrays
            for (int64_t r = 0; r < n; ++r)
                if (indptr[r + 1] < indptr[r]) throw std::invalid_argument("lhs indptr must be non-decreasing");
            for (int64_t q = 0; q < nnz; ++q)
                if (indices[q] < 0 || indices[q] >= n) throw std::invalid_argument("lhs column index out of range");
            setup_pattern(n, indptr, indices);
        }
        if (!uploaded) {
            if (xf) {
                xf->upload_and_compare(value_copies(data, rhs, nnz, n, K), {}, stream_);
            } else {
                lv_[0].A.upload_values(data, stream_);
                for (const auto& rr : rhs_rows_)
                    GMG_CUDA(cudaMemcpyAsync(rhs64_.ptr + (size_t)rr.first * K, rhs + (size_t)rr.first * K,
                                             (size_t)(rr.second - rr.first) * K * sizeof(double), cudaMemcpyHostToDevice, stream_));
            }
        }
        numeric_ready_ = false;
        if (xf) {
            GMG_CUDA(cudaEventRecord(rhs_ready_, stream2_));
            rhs_pending_ = true;
        }
        if (wait || !xf) {
            GMG_CUDA(cudaStreamSynchronize(stream_));
            GMG_CUDA(cudaStreamSynchronize(stream2_));
            rhs_pending_ = false;
        }
        staged_ = true;
        auto& tt = st_->transfer_timing;
        tt["stage_host_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        tt["pattern_reused"] = same ? 1.0 : 0.0;
        size_t value_entries = 0, rhs_rows = 0;
        for (const auto& g : lv_[0].A.segs) value_entries += (size_t)g.len;
        for (const auto& rr : rhs_rows_) rhs_rows += (size_t)(rr.second - rr.first);
        tt["h2d_bytes"] = (double)(value_entries * sizeof(double) + rhs_rows * K * sizeof(double) +
                                   (same ? 0 : (size_t)(n + 1 + value_entries) * sizeof(int)));
        tt["transfer_threads"] = xf ? (double)xf->threads() : 0.0;
    }

    // Multi-GPU: true only when `mine` is true on every rank (NCCL min over one int; one host synchronisation).
    bool all_ranks_agree(bool mine) {
        if (st_->dist.world <= 1 || !comm_) return mine;
        flag_dev_.ensure(1);
        int v = mine ? 1 : 0;
        GMG_CUDA(cudaMemcpyAsync(flag_dev_.ptr, &v, sizeof v, cudaMemcpyHostToDevice, stream_));
        GMG_NCCL(nccl().AllReduce(flag_dev_.ptr, flag_dev_.ptr, 1, ncclInt, ncclMin, comm_, stream_));
        GMG_CUDA(cudaMemcpyAsync(&v, flag_dev_.ptr, sizeof v, cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        return v != 0;
    }

    std::vector<HostTransfer::Copy> value_copies(const double* data, const double* rhs, int64_t nnz, int64_t n, int K) {
        std::vector<HostTransfer::Copy> c;
        // the matrix values this rank stores (multi-GPU: the row segments of the finest level it works on)
        for (const auto& g : lv_[0].A.segs) {
            HostTransfer::Copy cp;
            cp.dev = lv_[0].A.v64.ptr + g.l0, cp.host = data + g.g0, cp.bytes = (size_t)g.len * sizeof(double);
            c.push_back(cp);
        }
        // the right-hand side is not needed before the cycles start: it travels on a second stream so
        // the Galerkin reduction and the coarse factor begin as soon as the matrix values have arrived
        for (const auto& rr : rhs_rows_) {
            HostTransfer::Copy cp;
            cp.dev = rhs64_.ptr + (size_t)rr.first * K, cp.host = rhs + (size_t)rr.first * K;
            cp.bytes = (size_t)(rr.second - rr.first) * K * sizeof(double), cp.stream = stream2_;
            c.push_back(cp);
        }
        return c;
    }

    void fetch_solution(double* x_out) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!solved_) throw std::logic_error("fetch_solution before solve_staged");
        const auto t0 = std::chrono::steady_clock::now();
        const size_t count = (size_t)st_->n * K_;
        if (st_->dist.sharded(0)) allgather_rows(0, x_final_, stream_);
        if (sizeof(T) == 4 && !refine()) launch_cast_f32_f64(reinterpret_cast<const float*>(x_final_), x64_.ptr, count, stream_);
        const double* src = sizeof(T) == 4 ? x64_.ptr : reinterpret_cast<const double*>(x_final_);
        if (HostTransfer* xf = transfer()) {
            GMG_CUDA(cudaStreamSynchronize(stream_));
            xf->download(x_out, src, count * sizeof(double));
        } else {
            GMG_CUDA(cudaMemcpyAsync(x_out, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            GMG_CUDA(cudaStreamSynchronize(stream_));
        }
        st_->transfer_timing["fetch_host_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        st_->transfer_timing["d2h_bytes"] = (double)(count * sizeof(double));
    }

    // ------------------------------------------------------------------ device-resident systems
    // The staged pattern stays; only values and right-hand side change, and they are already in HBM
    // (assembled by the caller's own kernels or by the mesh functions below): device-to-device copies
    // on the solver stream, no host staging.
    void update_values_device(const double* d_vals, const double* d_rhs, int K) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!pattern_ready_ || !hierarchy_ready_) throw std::logic_error("update_values_device needs a staged sparsity pattern (stage a system or attach a mesh first)");
        if (st_->dist.world > 1) throw std::logic_error("update_values_device is single-GPU");
        set_columns(K);
        cudaPointerAttributes pa;
        for (const void* ptr : {(const void*)d_vals, (const void*)d_rhs}) {
            GMG_CUDA(cudaPointerGetAttributes(&pa, ptr));
            if (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged) throw std::invalid_argument("update_values_device expects device pointers");
        }
        GMG_CUDA(cudaMemcpyAsync(lv_[0].A.v64.ptr, d_vals, (size_t)lv_[0].A.nnz * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        GMG_CUDA(cudaMemcpyAsync(rhs64_.ptr, d_rhs, (size_t)st_->n * K * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        mark_staged_on_device();
    }

    void fetch_solution_device(double* d_x) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!solved_) throw std::logic_error("fetch_solution before solve_staged");
        if (st_->dist.world > 1) throw std::logic_error("fetch_solution_device is single-GPU");
        GMG_CUDA(cudaMemcpyAsync(d_x, solution_device(), (size_t)st_->n * K_ * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
    }

    // ------------------------------------------------------------------ mesh operators on the device
    void mesh_attach(int64_t nf, const int* faces) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (st_->dist.world > 1) throw std::logic_error("mesh assembly is single-GPU");
        const MeshTopology topo = build_mesh_topology(st_->n, nf, faces);
        if (indptr_changed(topo.pattern.indptr.data(), topo.pattern.indices.data(), st_->n)) {
            GMG_CUDA(cudaStreamSynchronize(stream_));
            setup_pattern(st_->n, topo.pattern.indptr.data(), topo.pattern.indices.data());
        }
        mesh_.attach(topo, faces, stream_);
        mesh_pos_.ensure((size_t)st_->n * 3), mesh_m_.ensure((size_t)st_->n), mesh_s_.ensure((size_t)lv_[0].A.nnz);
        mesh_has_pos_ = mesh_has_s_ = mesh_has_m_ = false;
    }

    void mesh_set_positions(const double* pos) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        require_mesh();
        GMG_CUDA(cudaMemcpyAsync(mesh_pos_.ptr, pos, (size_t)st_->n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        mesh_has_pos_ = true;
    }

    void mesh_stiffness() override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        require_mesh(true);
        mesh_.face_geometry(mesh_pos_.ptr, MESH_MASS_BARYCENTRIC, stream_);
        mesh_.stiffness(lv_[0].A.indptr.ptr, lv_[0].A.colp(), mesh_s_.ptr, stream_);
        mesh_has_s_ = true;
    }

    void mesh_mass(int type) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        require_mesh(true);
        if (type != MESH_MASS_BARYCENTRIC && type != MESH_MASS_VORONOI) throw std::invalid_argument("mass type must be 0 (barycentric) or 1 (Voronoi)");
        mesh_.face_geometry(mesh_pos_.ptr, type, stream_);
        mesh_.mass(mesh_m_.ptr, stream_);
        mesh_has_m_ = true;
    }

    // lhs = alpha M + beta S, rhs = M Y staged for solve_staged(); Y = the host array y (n x K) or, when y is
    // null, the resident vertex positions (K = 3).
    void mesh_system(double alpha, double beta, const double* y, int K) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        require_mesh(true);
        if (!mesh_has_s_ || !mesh_has_m_) throw std::logic_error("mesh_system needs mesh_stiffness and mesh_mass first");
        if (!y) K = 3;
        set_columns(K);
        const double* yd = mesh_pos_.ptr;
        if (y) {
            mesh_y_.upload(y, (size_t)st_->n * K, stream_);
            yd = mesh_y_.ptr;
        }
        mesh_.system(lv_[0].A.indptr.ptr, lv_[0].A.colp(), alpha, beta, mesh_s_.ptr, mesh_m_.ptr, yd, K, lv_[0].A.v64p(),
                     rhs64_.ptr, stream_);
        if (y) GMG_CUDA(cudaStreamSynchronize(stream_));  // the caller's buffer has been read
        mark_staged_on_device();
    }

    // demos/conformal_flow.py:54-59 without leaving the GPU: per step M_t = mass(V_t), lhs = M_t + tau S,
    // rhs = M_t V_t, V_{t+1} = normalize_area(solve(lhs, rhs)); S is the resident stiffness (fixed, as upstream).
    void mesh_flow(double tau, int mass_type, int steps) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        require_mesh(true);
        if (!mesh_has_s_) throw std::logic_error("mesh_flow needs mesh_stiffness first");
        if (steps < 0) throw std::invalid_argument("steps must be >= 0");
        cudaEvent_t e[4];
        for (auto& ev : e) GMG_CUDA(cudaEventCreate(&ev));
        double t_asm = 0, t_norm = 0, t_red = 0, t_fac = 0, t_cyc = 0, iters = 0;
        try {
            for (int it = 0; it < steps; ++it) {
                GMG_CUDA(cudaEventRecord(e[0], stream_));
                mesh_mass(mass_type);
                mesh_system(1.0, tau, nullptr, 3);
                GMG_CUDA(cudaEventRecord(e[1], stream_));
                solve_staged();
                GMG_CUDA(cudaEventRecord(e[2], stream_));
                mesh_.normalize_area(solution_device(), mesh_pos_.ptr, stream_);
                GMG_CUDA(cudaEventRecord(e[3], stream_));
                GMG_CUDA(cudaStreamSynchronize(stream_));
                float a = 0, b = 0;
                GMG_CUDA(cudaEventElapsedTime(&a, e[0], e[1]));
                GMG_CUDA(cudaEventElapsedTime(&b, e[2], e[3]));
                t_asm += a, t_norm += b;
                const auto& tm = st_->solver_timing;
                t_red += tm.at("reduction"), t_fac += tm.at("coarsest_solve"), t_cyc += tm.at("cycles"), iters += tm.at("iterations");
            }
        } catch (...) {
            for (auto& ev : e) cudaEventDestroy(ev);
            throw;
        }
        for (auto& ev : e) cudaEventDestroy(ev);
        auto& tt = st_->transfer_timing;
        tt["flow_steps"] = steps, tt["flow_assemble_ms"] = t_asm, tt["flow_normalize_ms"] = t_norm, tt["flow_reduction_ms"] = t_red;
        tt["flow_factor_ms"] = t_fac, tt["flow_cycles_ms"] = t_cyc, tt["flow_iterations"] = iters;
    }

    void mesh_get(int which, double* out) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        require_mesh();
        const double* src = nullptr;
        size_t count = 0;
        switch (which) {
            case 0: src = mesh_pos_.ptr, count = (size_t)st_->n * 3; if (!mesh_has_pos_) src = nullptr; break;
            case 1: src = mesh_s_.ptr, count = (size_t)lv_[0].A.nnz; if (!mesh_has_s_) src = nullptr; break;
            case 2: src = mesh_m_.ptr, count = (size_t)st_->n; if (!mesh_has_m_) src = nullptr; break;
            case 3: src = lv_[0].A.v64.ptr, count = (size_t)lv_[0].A.nnz; if (!staged_) src = nullptr; break;
            case 4: src = rhs64_.ptr, count = (size_t)st_->n * K_; if (!staged_) src = nullptr; break;
            default: throw std::invalid_argument("mesh_get: which must be 0..4");
        }
        if (!src) throw std::logic_error("mesh_get: that array has not been computed yet");
        GMG_CUDA(cudaMemcpyAsync(out, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
    }

    // ------------------------------------------------------------------ solve
    void solve_staged() override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!staged_) throw std::logic_error("solve_staged before stage_system");
        set_launch_pdl(st_->use_pdl);
        const gmg_params& p = st_->params;
        if (p.cycle_type < 0 || p.cycle_type > 2) throw std::invalid_argument("cycle_type must be 0 (V), 1 (F) or 2 (W)");
        if (p.cycle_type != 0 && st_->dist.world > 1) throw std::invalid_argument("F- and W-cycles are single-GPU for now");
        if (sizeof(T) == 4 && st_->dist.world > 1)
            throw std::invalid_argument("dtype float32 is single-GPU for now: the fp64 defect correction that lets fp32 levels reach a 1e-4 tolerance is not sharded");
        if (p.max_iter < 1) throw std::invalid_argument("max_iter must be >= 1");
        if (p.stopping_criteria < 0 || p.stopping_criteria > 3) throw std::invalid_argument("stopping_criteria must be 0..3");
        if (hist_res_.count < (size_t)p.max_iter) {
            drop_graphs();  // captured launches hold the old history pointers
            hist_res_.ensure(p.max_iter);
            hist_ms_.ensure(p.max_iter);
        }
        if (cycle_dirty_) build_cycle();
        if (st_->krylov) {
            solve_staged_krylov();
            return;
        }
        x_final_ = x_final_cycle_;
        int64_t launches = 0;

        GMG_CUDA(cudaMemsetAsync(ctl_.ptr, 0, sizeof(CycleControl), stream_));
        GMG_CUDA(cudaEventRecord(ev_[0], stream_));
        launches += setup_numeric(ev_[1]);
        // x0 = rhs (core.cpp:69), b = rhs
        if (rhs_pending_) {
            GMG_CUDA(cudaStreamWaitEvent(stream_, rhs_ready_, 0));
            rhs_pending_ = false;
        }
        set_initial_guess();
        launches += sizeof(T) == 4 ? 2 : 0;
        GMG_CUDA(cudaEventRecord(ev_[2], stream_));

        // ---- "cycles" (multigrid_solver.cpp:1411-1417)
        constexpr int kTraceCap = 1 << 16;
        if (st_->trace) trace_buf_.ensure(2 * (size_t)kTraceCap);
        launch_cycle_begin(ctl_.ptr, p.max_iter, p.stopping_criteria, p.tolerance, K_, stream_,
                           st_->trace ? trace_buf_.ptr : nullptr, st_->trace ? kTraceCap : 0);
        ++launches;
        for (const Op& op : prologue_) launches += run_op(op, stream_, 0);
        // multi-GPU: the NCCL exchanges are captured into the cycle graph too (option dist_graph)
        const bool graph = st_->use_graph && !st_->profile && (st_->dist.world <= 1 || st_->dist_graph);
        if (graph && st_->loop_mode == 1 && (st_->dist.world <= 1 || use_p2p())) {
            if (!while_exec_) build_while_graph();
            GMG_CUDA(cudaGraphLaunch(while_exec_, stream_));
        } else {
            if (graph && !cycle_exec_) build_cycle_graph();
            for (int it = 0; it < p.max_iter; ++it) {
                if (graph)
                    GMG_CUDA(cudaGraphLaunch(cycle_exec_, stream_));
                else
                    run_cycle(stream_, 0, st_->profile);
                GMG_CUDA(cudaMemcpyAsync(ctl_host_, ctl_.ptr, sizeof(CycleControl), cudaMemcpyDeviceToHost, stream_));
                GMG_CUDA(cudaStreamSynchronize(stream_));
                if (ctl_host_->done || ctl_host_->error) break;
            }
        }
        GMG_CUDA(cudaEventRecord(ev_[3], stream_));
        GMG_CUDA(cudaMemcpyAsync(ctl_host_, ctl_.ptr, sizeof(CycleControl), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        if (st_->profile) collect_profile();
        st_->trace_log.clear();
        if (st_->trace) {
            const int nt = std::min(ctl_host_->trace_n, ctl_host_->trace_cap);
            st_->trace_log.resize(2 * (size_t)nt);
            if (nt) GMG_CUDA(cudaMemcpy(st_->trace_log.data(), trace_buf_.ptr, 2 * (size_t)nt * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        }

        const int iters = ctl_host_->iter;
        launches += (int64_t)iters * launches_per_cycle_;
        st_->last_launches = launches;
        std::vector<double> res(std::max(iters, 1)), ms(std::max(iters, 1));
        if (iters > 0) {
            GMG_CUDA(cudaMemcpy(res.data(), hist_res_.ptr, iters * sizeof(double), cudaMemcpyDeviceToHost));
            GMG_CUDA(cudaMemcpy(ms.data(), hist_ms_.ptr, iters * sizeof(double), cudaMemcpyDeviceToHost));
        }
        st_->convergence.clear();
        for (int i = 0; i < iters; ++i) st_->convergence.emplace_back(ms[i], res[i]);
        float t01 = 0, t12 = 0, t23 = 0, t03 = 0;
        GMG_CUDA(cudaEventElapsedTime(&t01, ev_[0], ev_[1]));
        GMG_CUDA(cudaEventElapsedTime(&t12, ev_[1], ev_[2]));
        GMG_CUDA(cudaEventElapsedTime(&t23, ev_[2], ev_[3]));
        GMG_CUDA(cudaEventElapsedTime(&t03, ev_[0], ev_[3]));
        auto& tm = st_->solver_timing;
        tm["reduction"] = t01;
        tm["coarsest_solve"] = t12;
        tm["cycles"] = t23;
        tm["solver_total"] = t03;
        tm["iterations"] = (double)iters;
        tm["residue"] = ctl_host_->residue;
        solved_ = true;
        if (ctl_host_->error & 1) throw std::runtime_error("an operator has a missing, non-positive or non-finite diagonal entry (Jacobi smoother needs A_ii > 0)");
        if (ctl_host_->error & 4) throw std::runtime_error("coarsest-level Cholesky broke down: the Galerkin operator is not positive definite");
        if (ctl_host_->error & 16) throw std::runtime_error("coarsest-level factorisation stalled waiting for a tile (internal error)");
        if (ctl_host_->error & 8) throw std::runtime_error("multi-GPU halo exchange timed out waiting for a peer rank (ranks out of step, or a peer failed)");
        if (ctl_host_->error & 2) throw std::runtime_error("residual became non-finite (diverged); try a smaller omega");
    }

    // ------------------------------------------------------------------ conjugate gradients around the cycle
    // Option krylov = 1: preconditioned CG with ONE cycle (V, F or W; zero initial guess; the post-smoothing runs
    // the pre-smoothing dampings in reverse, so the cycle is a symmetric operator) as the preconditioner;
    // krylov = 2: plain CG (the reference's solverType 4, multigrid_solver.cpp:1453-1477). Same stopping rule,
    // timing keys and convergence trace as the cycle loop: one iteration = one cycle + one product with A.
    // The residual is carried by the recurrence r -= alpha A p; when it meets the tolerance the true residual
    // b - A x is evaluated, and the iteration restarts from it if rounding has let the two drift apart.
    void solve_staged_krylov() {
        const gmg_params& p = st_->params;
        if (sizeof(T) != 8) throw std::invalid_argument("the conjugate-gradient wrapper needs dtype float64");
        if (st_->dist.world > 1) throw std::invalid_argument("the conjugate-gradient wrapper is single-GPU for now");
        if (K_ > kPcgMaxK) throw std::invalid_argument("the conjugate-gradient wrapper handles 1..4 right-hand sides");
        if (n_levels_ == 0 && st_->krylov == 1) throw std::invalid_argument("no hierarchy to precondition with (N <= lower_bound): use the plain solve");
        const bool precond = st_->krylov == 1;
        const int n = lv_[0].n;
        const size_t count = (size_t)n * K_;
        double* xk = reinterpret_cast<double*>(kry_x_.ptr);
        if (kry_x_.count < count) kry_x_.ensure(count), kry_p_.ensure(count), kry_q_.ensure(count), xk = kry_x_.ptr;
        kry_sc_.ensure(1), kry_part_.ensure((size_t)kPcgBlocks * 2 * kPcgMaxK);
        if (kry_ops_dirty_ || kry_K_ != K_) build_krylov_cycle();
        double* r = reinterpret_cast<double*>(lv_[0].b.ptr);   // the residual is the right-hand side of the preconditioning cycle
        double* z = precond ? reinterpret_cast<double*>(kry_z_) : r;
        const double* w = p.stopping_criteria == 2 ? mass_.ptr : p.stopping_criteria == 1 ? minv_.ptr : nullptr;
        int64_t launches = 0;

        GMG_CUDA(cudaMemsetAsync(ctl_.ptr, 0, sizeof(CycleControl), stream_));
        GMG_CUDA(cudaEventRecord(ev_[0], stream_));
        launches += setup_numeric(ev_[1]);
        if (rhs_pending_) {
            GMG_CUDA(cudaStreamWaitEvent(stream_, rhs_ready_, 0));
            rhs_pending_ = false;
        }
        GMG_CUDA(cudaEventRecord(ev_[2], stream_));
        launch_cycle_begin(ctl_.ptr, p.max_iter, p.stopping_criteria, p.tolerance, K_, stream_);
        // x0 = rhs (core.cpp:69)
        GMG_CUDA(cudaMemcpyAsync(xk, rhs64_.ptr, count * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        NormChunks chunks;
        chunks.n_chunks = 1, chunks.kt[0] = K_;
        auto true_residual = [&](bool also_r) {  // r = b - A x (also_r) and its norm partials
            SpmvArgs<T> a = base_args(lv_[0].A);
            a.x = reinterpret_cast<const T*>(xk), a.b = reinterpret_cast<const T*>(rhs64_.ptr), a.out = reinterpret_cast<T*>(r);
            a.weight = w, a.partials = partials_.ptr;
            chunks.n_blocks[0] = launch_spmv<T>(also_r ? EPI_RESNORM : EPI_NORM, K_, a, lv_[0].A.plan, stream_);
            ++launches;
        };
        auto restart = [&] {  // direction state cleared: the next iteration is a steepest-descent step from r
            GMG_CUDA(cudaMemsetAsync(kry_sc_.ptr, 0, sizeof(PcgScalars), stream_));
            GMG_CUDA(cudaMemsetAsync(kry_p_.ptr, 0, count * sizeof(double), stream_));
            GMG_CUDA(cudaMemsetAsync(kry_q_.ptr, 0, count * sizeof(double), stream_));
            true_residual(true);
            if (precond)  // first smoothing sweep of the cycle from a zero guess: z0 = omega D^-1 r (alpha = 0: x, r unchanged)
                launch_pcg_update(n, K_, xk, r, kry_p_.ptr, kry_q_.ptr, rhs64_.ptr, w, reinterpret_cast<const double*>(lv_[0].dinv.ptr),
                                  weights64_.ptr, reinterpret_cast<double*>(kry_z0_), kry_sc_.ptr, kry_part_.ptr, stream_), ++launches;
        };
        restart();
        int restarts = 0, stale = 0;
        double best = 1.7976931348623157e308;
        bool done = false;
        while (!done) {
            // ---- one iteration: z = M^-1 r, beta, p, q = A p, alpha, x / r update, stopping test
            if (precond)
                for (const Op& op : kry_ops_) launches += run_op(op, stream_, 0);
            launch_pcg_dot(0, n, K_, r, z, kry_part_.ptr, tail_bar_.ptr + 4, kry_sc_.ptr, stream_);
            launch_pcg_direction(n, K_, z, kry_p_.ptr, kry_sc_.ptr, stream_);
            {
                SpmvArgs<T> a = base_args(lv_[0].A);  // q = A p, cancellation-free form where the level has it
                a.x = reinterpret_cast<const T*>(kry_p_.ptr), a.out = reinterpret_cast<T*>(kry_q_.ptr);
                launch_spmv<T>(EPI_SPMV, K_, a, lv_[0].A.plan, stream_);
            }
            launch_pcg_dot(1, n, K_, kry_p_.ptr, kry_q_.ptr, kry_part_.ptr, tail_bar_.ptr + 4, kry_sc_.ptr, stream_);
            chunks.n_blocks[0] = launch_pcg_update(n, K_, xk, r, kry_p_.ptr, kry_q_.ptr, rhs64_.ptr, w, reinterpret_cast<const double*>(lv_[0].dinv.ptr),
                                                   weights64_.ptr, precond ? reinterpret_cast<double*>(kry_z0_) : nullptr, kry_sc_.ptr, partials_.ptr, stream_);
            launch_norm_finalize(partials_.ptr, chunks, ctl_.ptr, hist_res_.ptr, hist_ms_.ptr, 1, 0, stream_);
            launches += 6;
            GMG_CUDA(cudaMemcpyAsync(ctl_host_, ctl_.ptr, sizeof(CycleControl), cudaMemcpyDeviceToHost, stream_));
            GMG_CUDA(cudaStreamSynchronize(stream_));
            if (ctl_host_->error) break;
            if (st_->krylov_patience > 0 && !ctl_host_->done) {  // run to the rounding floor: stop once the residual stagnates
                stale = ctl_host_->residue < 0.98 * best ? 0 : stale + 1;
                best = std::min(best, ctl_host_->residue);
                if (stale >= st_->krylov_patience) break;
            }
            if (!ctl_host_->done) continue;
            done = true;
            if (ctl_host_->residue <= p.tolerance) {
                // met by the recurrence: judge the true residual, carry on from it if it is not there yet
                true_residual(false);
                launch_norm_finalize(partials_.ptr, chunks, ctl_.ptr, nullptr, nullptr, 0, 0, stream_);
                ++launches;
                GMG_CUDA(cudaMemcpyAsync(ctl_host_, ctl_.ptr, sizeof(CycleControl), cudaMemcpyDeviceToHost, stream_));
                GMG_CUDA(cudaStreamSynchronize(stream_));
                if (ctl_host_->residue > p.tolerance && ctl_host_->iter < p.max_iter && restarts < 3) {
                    ++restarts;
                    GMG_CUDA(cudaMemsetAsync(&ctl_.ptr->done, 0, sizeof(int), stream_));
                    restart();
                    done = false;
                }
            }
        }
        GMG_CUDA(cudaEventRecord(ev_[3], stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        x_final_ = reinterpret_cast<T*>(xk);
        const int iters = ctl_host_->iter;
        st_->last_launches = launches;
        std::vector<double> res(std::max(iters, 1)), ms(std::max(iters, 1));
        if (iters > 0) {
            GMG_CUDA(cudaMemcpy(res.data(), hist_res_.ptr, iters * sizeof(double), cudaMemcpyDeviceToHost));
            GMG_CUDA(cudaMemcpy(ms.data(), hist_ms_.ptr, iters * sizeof(double), cudaMemcpyDeviceToHost));
        }
        st_->convergence.clear();
        for (int i = 0; i < iters; ++i) st_->convergence.emplace_back(ms[i], res[i]);
        float t01 = 0, t12 = 0, t23 = 0, t03 = 0;
        GMG_CUDA(cudaEventElapsedTime(&t01, ev_[0], ev_[1]));
        GMG_CUDA(cudaEventElapsedTime(&t12, ev_[1], ev_[2]));
        GMG_CUDA(cudaEventElapsedTime(&t23, ev_[2], ev_[3]));
        GMG_CUDA(cudaEventElapsedTime(&t03, ev_[0], ev_[3]));
        auto& tm = st_->solver_timing;
        tm["reduction"] = t01, tm["coarsest_solve"] = t12, tm["cycles"] = t23, tm["solver_total"] = t03;
        tm["iterations"] = (double)iters, tm["residue"] = ctl_host_->residue;
        st_->transfer_timing["krylov_restarts"] = restarts;
        solved_ = true;
        if (ctl_host_->error & 1) throw std::runtime_error("an operator has a missing, non-positive or non-finite diagonal entry (Jacobi smoother needs A_ii > 0)");
        if (ctl_host_->error & 4) throw std::runtime_error("coarsest-level Cholesky broke down: the Galerkin operator is not positive definite");
        if (ctl_host_->error & 2) throw std::runtime_error("residual became non-finite (diverged)");
    }

    // The preconditioning cycle: eps = cycle(A, r) from a zero guess, the residual in lv_[0].b; its first
    // pre-smoothing sweep is written by the update kernel of the CG iteration (kry_z0_), its result is kry_z_.
    void build_krylov_cycle() {
        const gmg_params& p = st_->params;
        std::vector<Op> saved;
        saved.swap(ops_);
        const int tail_level = tail_level_;
        tail_level_ = -1;
        T* cur = lv_[0].x.ptr;
        T* alt = lv_[0].t.ptr;
        kry_z0_ = cur;
        if (n_levels_ > 0) {
            const int pre_done = p.pre_iters >= 1 ? 1 : 0;
            if (!pre_done) {
                Op z;
                z.kind = OP_ZERO, z.level = 0, z.zero_ptr = cur, z.zero_bytes = (size_t)lv_[0].n * K_ * sizeof(T);
                ops_.push_back(z);
            }
            push_vcycle(0, cur, alt, pre_done, false, p.cycle_type);
        }
        kry_z_ = cur;
        kry_ops_.swap(ops_);
        ops_.swap(saved);
        tail_level_ = tail_level;
        kry_ops_dirty_ = false;
        kry_K_ = K_;
        set_launch_dry_run(true);
        try {
            for (const Op& op : kry_ops_)
                if (op.plan) run_op(op, stream_, 0);
        } catch (...) {
            set_launch_dry_run(false);
            throw;
        }
        set_launch_dry_run(false);
    }

    // ------------------------------------------------------------------ residualCheck
    double residual(int64_t n, const int* indptr, const int* indices, const double* data, const double* rhs,
                    const double* x, int K, int type) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (n != st_->n) throw std::invalid_argument("lhs has a different number of rows than the point set of the constructor");
        if (type < 0 || type > 3) throw std::invalid_argument("residual type must be 0..3");
        set_launch_pdl(st_->use_pdl);
        if (K < 1 || K > kMaxRhsTile * kMaxNormChunks) throw std::invalid_argument("number of right-hand sides must be 1..32");
        const int64_t nnz = indptr[n];
        q_indptr_.upload(indptr, n + 1, stream_);
        q_indices_.upload(indices, nnz, stream_, 8);
        q_vals_.upload(data, nnz, stream_, 8);
        q_b_.upload(rhs, (size_t)n * K, stream_);
        q_x_.upload(x, (size_t)n * K, stream_);
        CycleControl* aux = ctl_.ptr + 1;
        launch_cycle_begin(aux, 1, type, 0.0, K, stream_);
        SpmvPlan plan;
        plan.path = 1;
        plan.lanes = 1;
        SpmvArgs<double> a;
        a.n_rows = (int)n, a.ld = K;
        a.rowptr = q_indptr_.ptr, a.colidx = q_indices_.ptr, a.vals = q_vals_.ptr;
        if (st_->diff_form) {
            // the same cancellation-free row product the cycle's stopping test uses (scratch: loop-state slot 2)
            q_vd_.ensure((size_t)nnz, 8), q_dinv_.ensure((size_t)n), q_rho_.ensure(1);
            launch_extract_dinv<double>((int)n, q_indptr_.ptr, q_indices_.ptr, q_vals_.ptr, q_dinv_.ptr, q_rho_.ptr, ctl_.ptr + 2,
                                        stream_, q_vd_.ptr);
            a.vals = q_vd_.ptr, a.diff = 1;
        }
        a.weight = type == 2 ? mass_.ptr : type == 1 ? minv_.ptr : nullptr;
        NormChunks chunks;
        for (int k0 = 0; k0 < K; k0 += kMaxRhsTile) {
            const int kt = std::min(kMaxRhsTile, K - k0);
            a.x = q_x_.ptr + k0, a.b = q_b_.ptr + k0;
            a.partials = partials_.ptr + (size_t)chunks.n_chunks * kNormChunkStride;
            chunks.kt[chunks.n_chunks] = kt;
            chunks.n_blocks[chunks.n_chunks] = launch_spmv<double>(EPI_NORM, kt, a, plan, stream_);
            ++chunks.n_chunks;
        }
        launch_norm_finalize(partials_.ptr, chunks, aux, nullptr, nullptr, 0, 0, stream_);
        GMG_CUDA(cudaMemcpyAsync(ctl_host_, aux, sizeof(CycleControl), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        return ctl_host_->residue;
    }


    // "reduction" (multigrid_solver.cpp:1387-1392) and "coarsest_solve" factorisation (:1401) from
    // the values of A staged on the device. Returns the number of kernel launches.
    int64_t setup_numeric(cudaEvent_t after_reduction) {
        const int L = n_levels_;
        const gmg_params& p = st_->params;
        int64_t launches = 0;
        GMG_CUDA(cudaMemsetAsync(rho_.ptr, 0, kMaxLevels * sizeof(double), stream_));
        if (L > 0) {
            // windowed storage (multi-GPU): every rank sees its own rows only; the Gershgorin bound is their maximum
            const DevMat<T>& a0 = lv_[0].A;
            const bool win = window0_;  // a property of the layout, the same on every rank (it decides a collective)
            launch_extract_dinv<T>(win ? (int)st_->dist.end(0) : lv_[0].n, a0.indptr.ptr, a0.colp(), a0.v64p(), lv_[0].dinv.ptr, rho_.ptr,
                                   ctl_.ptr, stream_, a0.diff ? a0.vdp() : nullptr, win ? (int)st_->dist.begin(0) : 0);
            if (win) GMG_NCCL(nccl().AllReduce(rho_.ptr, rho_.ptr, 1, ncclDouble, ncclMax, comm_, stream_));
            ++launches;
        }
        lv_[0].A.refresh_cast(stream_);
        launches += sizeof(T) == 4 ? 1 : 0;
        for (int k = 0; k < L; ++k) {
            Level& f = lv_[k];
            Level& c = lv_[k + 1];
            const int64_t na = f.AP.e1 - f.AP.e0, nc = c.A.e1 - c.A.e0;
            if (f.AP.planned)
                launch_spgemm_planned(na, f.AP.pair_off.ptr, f.AP.pairs.ptr, f.A.v64p(), f.P.v64p(), f.AP.v64p() + f.AP.e0, stream_);
            else
                launch_spgemm_numeric(na, f.AP.rowidxp() + f.AP.e0, f.AP.colp() + f.AP.e0, f.AP.v64p() + f.AP.e0, f.A.indptr.ptr,
                                      f.A.colp(), f.A.v64p(), f.P.indptr.ptr, f.P.colp(), f.P.v64p(), stream_);
            if (c.A.planned)
                launch_spgemm_planned(nc, c.A.pair_off.ptr, c.A.pairs.ptr, f.R.v64p(), f.AP.v64p(), c.A.v64p() + c.A.e0, stream_);
            else
                launch_spgemm_numeric(nc, c.A.rowidxp() + c.A.e0, c.A.colp() + c.A.e0, c.A.v64p() + c.A.e0, f.R.indptr.ptr,
                                      f.R.colp(), f.R.v64p(), f.AP.indptr.ptr, f.AP.colp(), f.AP.v64p(), stream_);
            if (!c.A.share.empty()) {
                // every rank computed the rows of its coarse range: all-gather the values of A_{k+1}
                GMG_NCCL(nccl().GroupStart());
                for (int q = 0; q < st_->dist.world; ++q) {
                    const size_t cnt = (size_t)(c.A.share[q + 1] - c.A.share[q]);
                    double* at = c.A.v64p() + c.A.share[q];
                    if (cnt) GMG_NCCL(nccl().Broadcast(at, at, cnt, ncclDouble, q, comm_, stream_));
                }
                GMG_NCCL(nccl().GroupEnd());
            }
            launches += 2;
            if (k + 1 < L) {
                launch_extract_dinv<T>(c.n, c.A.indptr.ptr, c.A.colp(), c.A.v64p(), c.dinv.ptr, rho_.ptr + k + 1, ctl_.ptr,
                                       stream_);
                c.A.refresh_cast(stream_);
                launches += sizeof(T) == 4 ? 2 : 1;
            }
        }
        if (L > 0) {
            launch_smoother_weights<T>(rho_.ptr, L, p.pre_iters, p.post_iters, p.smoother, p.omega, p.cheb_alpha, weights_.ptr,
                                       weights64_.ptr, stream_);
            ++launches;
        }
        if (after_reduction) GMG_CUDA(cudaEventRecord(after_reduction, stream_));
        coarse_.set_dataflow(st_->coarse_dataflow);
        coarse_.factor(lv_[L].A.indptr.ptr, lv_[L].A.colp(), lv_[L].A.v64p(), ctl_.ptr, stream_, st_->profile);
        launches += coarse_.launches_per_factor();
        numeric_ready_ = true;
        return launches;
    }

    void smoother_weights(int level, double* rho, double* pre, double* post) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!numeric_ready_) throw std::logic_error("smoother weights exist after a solve or a level_op");
        if (level < 0 || level >= n_levels_) throw std::invalid_argument("level has no smoother");
        double w[2 * kMaxSweeps];
        GMG_CUDA(cudaMemcpyAsync(rho, rho_.ptr + level, sizeof(double), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaMemcpyAsync(w, weights64_.ptr + (size_t)level * 2 * kMaxSweeps, sizeof w, cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        for (int j = 0; j < st_->params.pre_iters; ++j) pre[j] = w[j];
        for (int j = 0; j < st_->params.post_iters; ++j) post[j] = w[kMaxSweeps + j];
    }

    void check_setup_errors() {
        GMG_CUDA(cudaMemcpyAsync(ctl_host_, ctl_.ptr, sizeof(CycleControl), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        if (ctl_host_->error & 1) throw std::runtime_error("an operator has a missing, non-positive or non-finite diagonal entry (Jacobi smoother needs A_ii > 0)");
        if (ctl_host_->error & 4) throw std::runtime_error("coarsest-level Cholesky broke down: the Galerkin operator is not positive definite");
        if (ctl_host_->error & 16) throw std::runtime_error("coarsest-level factorisation stalled waiting for a tile (internal error)");
    }

    // ------------------------------------------------------------------ single operators
    // One operator of the V-cycle on caller-supplied vectors (parity tests and per-kernel
    // measurement); uses the level's own device buffers, so it discards a previous solution.
    void level_op(int kind, int level, const double* a, const double* b, double* out, int sweeps) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!staged_) throw std::logic_error("level_op before stage_system");
        if (st_->dist.world > 1) throw std::logic_error("level_op is a single-GPU test entry point");
        set_launch_pdl(st_->use_pdl);
        const int L = n_levels_;
        if (level < 0 || level > L) throw std::invalid_argument("level out of range");
        if (cycle_dirty_) build_cycle();
        if (!numeric_ready_) {
            GMG_CUDA(cudaMemsetAsync(ctl_.ptr, 0, sizeof(CycleControl), stream_));
            setup_numeric(nullptr);
            check_setup_errors();
        }
        solved_ = false;
        GMG_CUDA(cudaMemsetAsync(ctl_.ptr, 0, sizeof(CycleControl), stream_));  // done = 0
        io64_.ensure((size_t)st_->n * K_);
        auto up = [&](T* dst, const double* src, size_t count) {
            if (sizeof(T) == 8) {
                GMG_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, stream_));
            } else {
                GMG_CUDA(cudaMemcpyAsync(io64_.ptr, src, count * sizeof(double), cudaMemcpyHostToDevice, stream_));
                launch_cast_f64_f32(io64_.ptr, reinterpret_cast<float*>(dst), count, stream_);
                GMG_CUDA(cudaStreamSynchronize(stream_));
            }
        };
        auto down = [&](const T* src, size_t count) {
            if (sizeof(T) == 8) {
                GMG_CUDA(cudaMemcpyAsync(out, src, count * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            } else {
                launch_cast_f32_f64(reinterpret_cast<const float*>(src), io64_.ptr, count, stream_);
                GMG_CUDA(cudaMemcpyAsync(out, io64_.ptr, count * sizeof(double), cudaMemcpyDeviceToHost, stream_));
            }
            GMG_CUDA(cudaStreamSynchronize(stream_));
        };
        Level& f = lv_[level];
        const size_t nf = (size_t)f.n * K_;
        Op op;
        op.level = level;
        switch (kind) {
            case OP_JACOBI: {
                if (level >= L) throw std::invalid_argument("the coarsest level has no smoother");
                if (!a || !b) throw std::invalid_argument("jacobi needs x and b");
                up(f.x.ptr, a, nf), up(f.b.ptr, b, nf);
                T* cur = f.x.ptr;
                T* alt = f.t.ptr;
                op.kind = OP_JACOBI, op.epi = EPI_JACOBI, op.plan = &f.A.plan, op.args = base_args(f.A);
                if (st_->params.pre_iters < 1) throw std::invalid_argument("level_op jacobi uses the pre-smoothing dampings: pre_iters must be >= 1");
                for (int i = 0; i < sweeps; ++i) {
                    op.args.x = cur, op.args.b = f.b.ptr, op.args.dinv = f.dinv.ptr, op.args.out = alt;
                    op.args.omega_ptr = weight_ptr(level, false, i % st_->params.pre_iters);
                    run_op(op, stream_, 0);
                    std::swap(cur, alt);
                }
                down(cur, nf);
                break;
            }
            case OP_RESIDUAL: {
                if (!a || !b) throw std::invalid_argument("residual needs x and b");
                up(f.x.ptr, a, nf), up(f.b.ptr, b, nf);
                op.kind = OP_RESIDUAL, op.epi = EPI_RESIDUAL, op.plan = &f.A.plan, op.args = base_args(f.A);
                op.args.x = f.x.ptr, op.args.b = f.b.ptr, op.args.out = f.r.ptr;
                run_op(op, stream_, 0);
                down(f.r.ptr, nf);
                break;
            }
            case OP_RESTRICT: {
                if (level >= L) throw std::invalid_argument("no prolongation below the coarsest level");
                if (!a) throw std::invalid_argument("restrict needs r");
                up(f.r.ptr, a, nf);
                op.kind = OP_RESTRICT, op.epi = EPI_SPMV, op.plan = &f.R.plan, op.args = base_args(f.R);
                op.args.x = f.r.ptr, op.args.out = lv_[level + 1].b.ptr;
                run_op(op, stream_, 0);
                down(lv_[level + 1].b.ptr, (size_t)lv_[level + 1].n * K_);
                break;
            }
            case OP_PROLONG: {
                if (level >= L) throw std::invalid_argument("no prolongation below the coarsest level");
                if (!a || !b) throw std::invalid_argument("prolong_add needs eps and x");
                up(lv_[level + 1].x.ptr, a, (size_t)lv_[level + 1].n * K_), up(f.x.ptr, b, nf);
                op.kind = OP_PROLONG, op.epi = EPI_ADD, op.plan = &f.P.plan, op.args = base_args(f.P);
                op.args.x = lv_[level + 1].x.ptr, op.args.xin = f.x.ptr, op.args.out = f.t.ptr;
                run_op(op, stream_, 0);
                down(f.t.ptr, nf);
                break;
            }
            case OP_COARSE: {
                if (level != L) throw std::invalid_argument("the direct solve lives on the coarsest level");
                if (!a) throw std::invalid_argument("coarse_solve needs b");
                up(f.b.ptr, a, nf);
                op.kind = OP_COARSE;
                run_op(op, stream_, 0);
                down(f.x.ptr, nf);
                break;
            }
            default:
                throw std::invalid_argument("unknown operator kind");
        }
    }

    // Average device time of `reps` back-to-back launches of one operator on the level's resident
    // buffers (CUDA events on the launch stream; Jacobi ping-pongs x <-> t as in the cycle).
    double time_op(int kind, int level, int reps) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!staged_) throw std::logic_error("time_op before stage_system");
        if (st_->dist.world > 1) throw std::logic_error("time_op is a single-GPU measurement entry point");
        set_launch_pdl(st_->use_pdl);
        const int L = n_levels_;
        if (level < 0 || level > L || reps < 1) throw std::invalid_argument("level or repetition count out of range");
        if (cycle_dirty_) build_cycle();
        if (!numeric_ready_) {
            GMG_CUDA(cudaMemsetAsync(ctl_.ptr, 0, sizeof(CycleControl), stream_));
            setup_numeric(nullptr);
            check_setup_errors();
        }
        solved_ = false;
        Level& f = lv_[level];
        Op op;
        op.level = level;
        T* cur = f.x.ptr;
        T* alt = f.t.ptr;
        switch (kind) {
            case OP_JACOBI:
                if (level >= L || st_->params.pre_iters < 1) throw std::invalid_argument("no smoother on this level");
                op.kind = OP_JACOBI, op.epi = EPI_JACOBI, op.plan = &f.A.plan, op.args = base_args(f.A);
                op.args.b = f.b.ptr, op.args.dinv = f.dinv.ptr, op.args.omega_ptr = weight_ptr(level, false, 0);
                break;
            case OP_RESIDUAL:
                op.kind = OP_RESIDUAL, op.epi = EPI_RESIDUAL, op.plan = &f.A.plan, op.args = base_args(f.A);
                op.args.b = f.b.ptr;
                break;
            case OP_RESTRICT:
                if (level >= L) throw std::invalid_argument("no prolongation below the coarsest level");
                op.kind = OP_RESTRICT, op.epi = EPI_SPMV, op.plan = &f.R.plan, op.args = base_args(f.R);
                break;
            case OP_PROLONG:
                if (level >= L) throw std::invalid_argument("no prolongation below the coarsest level");
                op.kind = OP_PROLONG, op.epi = EPI_ADD, op.plan = &f.P.plan, op.args = base_args(f.P);
                break;
            default:
                throw std::invalid_argument("time_op: kind must be jacobi, residual, restrict or prolong_add");
        }
        cudaEvent_t e0, e1;
        GMG_CUDA(cudaEventCreate(&e0));
        GMG_CUDA(cudaEventCreate(&e1));
        for (int i = -3; i < reps; ++i) {
            if (i == 0) GMG_CUDA(cudaEventRecord(e0, stream_));
            switch (kind) {
                case OP_JACOBI: op.args.x = cur, op.args.out = alt; std::swap(cur, alt); break;
                case OP_RESIDUAL: op.args.x = f.x.ptr, op.args.out = f.r.ptr; break;
                case OP_RESTRICT: op.args.x = f.r.ptr, op.args.out = lv_[level + 1].b.ptr; break;
                default: op.args.x = lv_[level + 1].x.ptr, op.args.xin = f.x.ptr, op.args.out = f.t.ptr; break;
            }
            run_op(op, stream_, 0);
        }
        GMG_CUDA(cudaEventRecord(e1, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        float ms = 0;
        GMG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0), cudaEventDestroy(e1);
        return 1e3 * ms / reps;
    }

    void get_level_matrix(int level, int* indptr, int* indices, double* data) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (!pattern_ready_ || level < 0 || level > n_levels_) throw std::invalid_argument("no such level staged on the device");
        if (level > 0 && !numeric_ready_) throw std::logic_error("Galerkin operators exist after a solve or a level_op");
        const DevMat<T>& m = lv_[level].A;
        if (m.windowed()) throw std::logic_error("this level is stored by row windows on a multi-GPU layout: no rank holds the whole operator");
        GMG_CUDA(cudaMemcpyAsync(indptr, m.indptr.ptr, ((size_t)m.rows + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaMemcpyAsync(indices, m.indices.ptr, (size_t)m.nnz * sizeof(int), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaMemcpyAsync(data, m.v64.ptr, (size_t)m.nnz * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
    }

    bool level_info(int level, int64_t* rows, int64_t* nnz_a, int64_t* nnz_u) override {
        if (!pattern_ready_ || level < 0 || level > n_levels_) return false;
        *rows = lv_[level].n;
        *nnz_a = lv_[level].A.nnz;
        *nnz_u = level < n_levels_ ? lv_[level].P.nnz : 0;
        return true;
    }
    void kernel_profile(int kind, int level, double* total_ms, int64_t* launches) override {
        *total_ms = 0.0, *launches = 0;
        if (kind < 0 || kind >= OP_KINDS) return;
        for (int l = 0; l < kMaxLevels; ++l) {
            if (level >= 0 && l != level) continue;
            *total_ms += prof_[kind][l].ms;
            *launches += prof_[kind][l].launches;
        }
    }
    void reset_kernel_profile() override {
        for (auto& row : prof_)
            for (auto& s : row) s = ProfileSlot();
    }

private:
    struct Level {
        int n = 0;
        DevMat<T> A;    // operator of this level (level 0: the caller's lhs; k >= 1: Galerkin)
        DevMat<T> P;    // U[k]    : n_k x n_{k+1}
        DevMat<T> R;    // U[k]^T  : n_{k+1} x n_k, explicit so restriction is a gather
        DevMat<T> AP;   // A_k U_k : intermediate of the Galerkin product (fp64 only)
        DeviceBuffer<T> dinv, x, t, b, r;
    };
    struct Op {
        int kind = 0, level = 0, epi = 0;
        SpmvArgs<T> args;
        const SpmvPlan* plan = nullptr;
        void* zero_ptr = nullptr;
        size_t zero_bytes = 0;
        int halo_op = 0;          // OP_HALO: HALO_A / HALO_R / HALO_P of `level`
        T* vec = nullptr;         // OP_HALO / OP_ALLGATHER: the vector
        T* vec2 = nullptr;        // OP_ALLGATHER: optional second vector of the same level
    };

    // ---- setup -------------------------------------------------------------------------
    // Symbolic phase, once per hierarchy and sparsity pattern of the lhs (host work in
    // compute_level_patterns / compute_dist_layout, solver.h): upload U, R = U^T and the patterns
    // of A_k U_k and of every Galerkin operator, build the row-tile plans for this rank's row
    // ranges, the halo index lists and the coarse workspace.
    void setup_pattern(int64_t n, const int* indptr, const int* indices) {
        mesh_.detach();  // gather lists of an attached mesh belong to the pattern being replaced
        const auto& U = st_->hier.U;
        n_levels_ = (int)U.size();
        if (n_levels_ + 1 > kMaxLevels) throw std::invalid_argument("too many levels");
        if (!hierarchy_ready_) st_->r_host.clear();
        compute_level_patterns(*st_, n, indptr, indices);
        compute_dist_layout(*st_);
        const DistLayout& d = st_->dist;
        if (d.world > 1 && !comm_) throw std::logic_error("multi-GPU layout configured but gmg_dist_init has not been called");
        auto range = [&](int level, bool sharded, int& b, int& e) {
            b = sharded ? (int)d.begin(level) : 0;
            e = sharded ? (int)d.end(level) : -1;
        };
        release_peer_arena();
        lv_.clear();
        lv_.resize(n_levels_ + 1);
        lv_[0].n = (int)st_->n;
        for (int k = 0; k < n_levels_; ++k) lv_[k + 1].n = (int)U[k].cols;
        // multi-GPU: rows of a sharded operator that take part in an exchange, either as producer
        // (rows peers gather, from the send lists of every halo kind whose vector has these rows) or
        // as consumer (rows with a column outside this rank's range of the gathered vector)
        auto early_rows = [&](const HostCsr& m, int row_level, int col_level, std::initializer_list<std::pair<int, int>> sends) {
            std::vector<char> mark((size_t)m.rows, 0);
            const int64_t rb = d.begin(row_level), re = d.end(row_level);
            const bool cols_sharded = d.sharded(col_level);
            const int64_t cb = cols_sharded ? d.begin(col_level) : 0, ce = cols_sharded ? d.end(col_level) : m.cols;
            for (int64_t r = rb; r < re; ++r)
                for (int q = m.indptr[r]; q < m.indptr[r + 1] && !mark[r]; ++q)
                    if (m.indices[q] < cb || m.indices[q] >= ce) mark[r] = 1;
            for (auto hk : sends) {
                if (hk.second < 0 || hk.second >= (int)d.halo[hk.first].size()) continue;
                for (const auto& list : d.halo[hk.first][hk.second].send)
                    for (int r : list) mark[r] = 1;
            }
            return mark;
        };
        // multi-GPU: a sharded fine level computes only its share of the Galerkin product — the rows of A_{k+1} in
        // this rank's coarse range and the rows [lo, hi) of A_k U_k those need (own fine range plus the rows its
        // restriction gathers); the level operator is all-gathered per solve. On by default (option
        // dist_shard_setup = 0: every rank computes whole products from whole operators).
        const bool shard_setup = d.world > 1 && st_->dist_shard_setup != 0;
        auto product_rows = [&](int k, int64_t& lo, int64_t& hi) {
            const HostCsr& r = st_->r_host[k];
            const int64_t cb = d.begin(k + 1), ce = d.end(k + 1);
            lo = d.begin(k), hi = d.end(k);
            for (int64_t I = cb; I < ce; ++I)
                for (int q = r.indptr[I]; q < r.indptr[I + 1]; ++q) {
                    lo = std::min<int64_t>(lo, r.indices[q]);
                    hi = std::max<int64_t>(hi, (int64_t)r.indices[q] + 1);
                }
            if (ce <= cb) lo = hi = d.begin(k);
        };
        // ... and then the finest level is STORED by row segments too (DevMat): A_0 and A_0 U_0 the rows this rank's
        // products read (own range + the rows its restriction gathers), U_0 the rows those products and the
        // prolongation read, U_0^T this rank's coarse rows. HBM footprint and the per-solve upload of a rank are
        // ~1/world of the system (option dist_window = 0: whole operators everywhere).
        compute_level0_windows(*st_);
        const bool window0 = st_->win0.on;
        window0_ = window0;
        const RowRanges &a_rows = st_->win0.a_rows, &p_rows = st_->win0.p_rows, &c_rows = st_->win0.c_rows;
        rhs_rows_ = st_->win0.rhs_rows;
        int b = 0, e = -1;
        for (int k = 0; k < n_levels_; ++k) {
            const HostCsr& u = U[k];
            const HostCsr& r = st_->r_host[k];
            const bool win = window0 && k == 0;
            lv_[k].P.upload_pattern(u, stream_, win ? &p_rows : nullptr);
            lv_[k].P.upload_values(u.data.data(), stream_);
            lv_[k].P.refresh_cast(stream_);
            range(k, d.sharded(k), b, e);
            if (d.sharded(k)) {  // prolongation: rows of level k, gathers level k + 1, writes x_k (gathered through A_k)
                const std::vector<char> early = early_rows(u, k, k + 1, {{HALO_A, k}});
                lv_[k].P.make_plan(lv_[k].P.indptr_host(u), st_->kernel_path, st_->staged_lanes, stream_, b, e, &early);
            } else
                lv_[k].P.make_plan(u.indptr, st_->kernel_path, st_->staged_lanes, stream_, b, e);
            lv_[k].R.upload_pattern(r, stream_, win ? &c_rows : nullptr);
            lv_[k].R.upload_values(r.data.data(), stream_);
            lv_[k].R.refresh_cast(stream_);
            range(k + 1, d.sharded(k), b, e);  // rows of R are coarse points; sharded with the fine level
            const int lanes_r = st_->staged_lanes_r ? st_->staged_lanes_r : st_->staged_lanes;
            const int path_r = st_->restrict_path >= 0 ? st_->restrict_path : st_->kernel_path;
            if (d.sharded(k)) {  // restriction: rows of level k + 1, gathers r_k, writes b_{k+1} and the first x_{k+1}
                const std::vector<char> early = early_rows(r, k + 1, k, {{HALO_A, k + 1}});
                lv_[k].R.make_plan(lv_[k].R.indptr_host(r), path_r, lanes_r, stream_, b, e, &early);
            } else
                lv_[k].R.make_plan(r.indptr, path_r, lanes_r, stream_, b, e);
            if (lv_[k].R.plan.path == 1 && st_->staged_lanes_r) lv_[k].R.plan.lanes = st_->staged_lanes_r;
            lv_[k].AP.upload_pattern(st_->ap_pat[k], stream_, win ? &a_rows : nullptr);
            lv_[k].AP.make_rowidx(stream_);
        }
        for (int k = 0; k <= n_levels_; ++k) {
            lv_[k].A.upload_pattern(st_->a_pat[k], stream_, window0 && k == 0 ? &a_rows : nullptr);
            if (k == 0 && n_levels_ > 0 && st_->diff_form) {
                lv_[0].A.vd.ensure(lv_[0].A.stored, 8);
                GMG_CUDA(cudaMemsetAsync(lv_[0].A.vd.ptr, 0, (lv_[0].A.stored + 8) * sizeof(double), stream_));
                lv_[0].A.diff = true;
            }
            if (k > 0) lv_[k].A.make_rowidx(stream_);
            range(k, d.sharded(k), b, e);
            if (d.sharded(k)) {  // sweeps / residual / norm: write x_k (gathered through A_k, or U_{k-1}) or r_k (through R_k)
                const std::vector<char> early = early_rows(st_->a_pat[k], k, k, {{HALO_A, k}, {HALO_R, k}, {HALO_P, k - 1}});
                lv_[k].A.make_plan(lv_[k].A.indptr_host(st_->a_pat[k]), st_->kernel_path, st_->staged_lanes, stream_, b, e, &early);
            } else
                lv_[k].A.make_plan(st_->a_pat[k].indptr, st_->kernel_path, st_->staged_lanes, stream_, b, e);
            lv_[k].dinv.ensure(std::max(lv_[k].n, 1));
        }
        // Galerkin product plans (once per pattern): which value pairs make up every entry of
        // A_k U_k and of U_k^T (A_k U_k). Skipped above a memory budget (the search kernel remains).
        long long plan_pairs = 0;
        for (int k = 0; k < n_levels_; ++k) {
            Level& f = lv_[k];
            Level& c = lv_[k + 1];
            f.AP.planned = c.A.planned = false;
            f.AP.e0 = 0, f.AP.e1 = f.AP.nnz, c.A.e0 = 0, c.A.e1 = c.A.nnz, c.A.share.clear();
            if (d.sharded(k) && shard_setup) {
                const HostCsr& ap = st_->ap_pat[k];
                const HostCsr& ac = st_->a_pat[k + 1];
                const int64_t cb = d.begin(k + 1), ce = d.end(k + 1);
                int64_t lo = 0, hi = 0;
                product_rows(k, lo, hi);
                f.AP.e0 = ap.indptr[lo], f.AP.e1 = ap.indptr[hi];
                if (f.AP.windowed()) f.AP.e0 = 0, f.AP.e1 = f.AP.stored;  // exactly the rows this rank needs are stored
                c.A.e0 = ac.indptr[cb], c.A.e1 = ac.indptr[ce];
                c.A.share.resize(d.world + 1);
                for (int q = 0; q <= d.world; ++q) c.A.share[q] = ac.indptr[d.ranges[k + 1][q]];
            }
            if (!st_->spgemm_plan || plan_pairs > st_->spgemm_plan_max_pairs) continue;
            plan_pairs += build_spgemm_plan(f.AP.e1 - f.AP.e0, f.AP.rowidxp() + f.AP.e0, f.AP.colp() + f.AP.e0, f.A.indptr.ptr,
                                            f.A.colp(), f.P.indptr.ptr, f.P.colp(), f.AP.pair_off, f.AP.pairs, stream_);
            f.AP.planned = true;
            plan_pairs += build_spgemm_plan(c.A.e1 - c.A.e0, c.A.rowidxp() + c.A.e0, c.A.colp() + c.A.e0, f.R.indptr.ptr,
                                            f.R.colp(), f.AP.indptr.ptr, f.AP.colp(), c.A.pair_off, c.A.pairs, stream_);
            c.A.planned = true;
        }
        st_->transfer_timing["galerkin_plan_pairs"] = (double)plan_pairs;
        // L2 residency of the cycle: the finest operator is read by six kernels per cycle (sweeps, residual,
        // norm) and, with its vectors, fits the 126 MB L2 of a B200 when nothing else displaces it. Its slabs are
        // fetched with evict_last, the once-per-cycle transfer operators of the finest level with evict_first.
        {
            int l2_bytes = 0, dev = 0;
            GMG_CUDA(cudaGetDevice(&dev));
            GMG_CUDA(cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, dev));
            for (auto& l : lv_) l.A.l2_hint = l.P.l2_hint = l.R.l2_hint = 0;
            if (n_levels_ > 0) {
                const DevMat<T>& a0 = lv_[0].A;
                const double rows = a0.plan.row_end >= 0 ? (double)(a0.plan.row_end - a0.plan.row_begin) / std::max(a0.rows, 1) : 1.0;
                const double a_bytes = rows * ((double)a0.nnz * (sizeof(T) + 4) + (double)a0.rows * 4);
                if (a_bytes <= 0.75 * l2_bytes) lv_[0].A.l2_hint = 1;
                lv_[0].P.l2_hint = lv_[0].R.l2_hint = 2;
            }
        }
        upload_halos();
        coarse_.setup(lv_[n_levels_].n, stream_);
        GMG_CUDA(cudaStreamSynchronize(stream_));
        hierarchy_ready_ = true;
        pattern_ready_ = true;
        if (K_) allocate_vectors();
        invalidate_cycle();
    }

    // ---- multi-GPU: halo index lists on the device and the staging buffers of the exchanges
    struct DevHalo {
        std::vector<DeviceBuffer<int>> send_idx, recv_idx;  // per peer
        std::vector<int> n_send, n_recv;
        DeviceBuffer<unsigned char> mask;  // per row of the exchanged vector: bit q = peer q gathers it (fused push)
    };

    void upload_halos() {
        const DistLayout& d = st_->dist;
        for (auto& per_op : halo_) per_op.clear();
        max_halo_ = 0;
        if (d.world <= 1) return;
        for (int hop = 0; hop < 3; ++hop) {
            halo_[hop].resize(n_levels_ + 1);
            for (int k = 0; k <= n_levels_; ++k) {
                const HaloLists& h = d.halo[hop][k];
                DevHalo& dh = halo_[hop][k];
                if (h.send.empty()) continue;
                dh.send_idx.resize(d.world), dh.recv_idx.resize(d.world);
                dh.n_send.assign(d.world, 0), dh.n_recv.assign(d.world, 0);
                size_t tot_s = 0, tot_r = 0;
                for (int q = 0; q < d.world; ++q) {
                    dh.n_send[q] = (int)h.send[q].size(), dh.n_recv[q] = (int)h.recv[q].size();
                    if (dh.n_send[q]) dh.send_idx[q].upload(h.send[q], stream_);
                    if (dh.n_recv[q]) dh.recv_idx[q].upload(h.recv[q], stream_);
                    tot_s += h.send[q].size(), tot_r += h.recv[q].size();
                }
                max_halo_ = std::max(max_halo_, std::max(tot_s, tot_r));
                // rows of the exchanged vector: level k for A and R, level k + 1 for P
                const int vec_level = hop == HALO_P ? k + 1 : k;
                std::vector<unsigned char> m((size_t)std::max(lv_[vec_level].n, 1), 0);
                for (int q = 0; q < d.world; ++q)
                    for (int row : h.send[q]) m[row] |= (unsigned char)(1u << q);
                dh.mask.upload(m, stream_);
            }
        }
        GMG_CUDA(cudaStreamSynchronize(stream_));  // the host lists may be rebuilt
    }

    ncclDataType_t nccl_type() const { return sizeof(T) == 8 ? ncclDouble : ncclFloat; }

    // Exchange the entries of the global-length vector v (K_ columns) that peers need / own.
    void exchange_halo(const DevHalo& dh, T* v, cudaStream_t s) {
        const int world = st_->dist.world;
        size_t off = 0;
        std::vector<size_t> soff(world), roff(world);
        for (int q = 0; q < world; ++q) {
            soff[q] = off;
            if (dh.n_send[q]) launch_pack<T>(v, dh.send_idx[q].ptr, dh.n_send[q], K_, halo_send_.ptr + off * K_, s);
            off += dh.n_send[q];
        }
        off = 0;
        for (int q = 0; q < world; ++q) roff[q] = off, off += dh.n_recv[q];
        GMG_NCCL(nccl().GroupStart());
        for (int q = 0; q < world; ++q) {
            if (dh.n_send[q]) GMG_NCCL(nccl().Send(halo_send_.ptr + soff[q] * K_, (size_t)dh.n_send[q] * K_, nccl_type(), q, comm_, s));
            if (dh.n_recv[q]) GMG_NCCL(nccl().Recv(halo_recv_.ptr + roff[q] * K_, (size_t)dh.n_recv[q] * K_, nccl_type(), q, comm_, s));
        }
        GMG_NCCL(nccl().GroupEnd());
        for (int q = 0; q < world; ++q)
            if (dh.n_recv[q]) launch_unpack<T>(v, dh.recv_idx[q].ptr, dh.n_recv[q], K_, halo_recv_.ptr + roff[q] * K_, s);
    }

    // Every rank contributes its row range of a global-length vector (in place).
    void allgather_rows(int level, T* v, cudaStream_t s) {
        const DistLayout& d = st_->dist;
        GMG_NCCL(nccl().GroupStart());
        for (int q = 0; q < d.world; ++q) {
            const size_t b = (size_t)d.ranges[level][q] * K_, cnt = (size_t)(d.ranges[level][q + 1] - d.ranges[level][q]) * K_;
            if (cnt) GMG_NCCL(nccl().Broadcast(v + b, v + b, cnt, nccl_type(), q, comm_, s));
        }
        GMG_NCCL(nccl().GroupEnd());
    }

    void dist_init(const void* unique_id) override {
        GMG_CUDA(cudaSetDevice(st_->params.device));
        if (comm_) {
            nccl().CommDestroy(comm_);
            comm_ = nullptr;
        }
        if (st_->dist.world <= 1) return;
        ncclUniqueId id;
        std::memcpy(&id, unique_id, sizeof id);
        GMG_NCCL(nccl().CommInitRank(&comm_, st_->dist.world, id, st_->dist.rank));
        invalidate_hierarchy();
    }

    // Worker threads + pinned slots for caller-owned buffers (host_xfer.h); nullptr when the
    // option xfer_threads is 0 (plain cudaMemcpyAsync from pageable memory).
    HostTransfer* transfer() {
        int want = st_->xfer_threads;
        // auto: the host cores of this rank's share of the box (one process per GPU), at most 16
        if (want < 0) want = std::min(16, std::max(2, (int)std::thread::hardware_concurrency() / std::max(1, st_->dist.world)));
        if (want == 0) {
            xfer_.reset();
            return nullptr;
        }
        if (!xfer_ || xfer_->threads() != want) {
            xfer_.reset();
            xfer_.reset(new HostTransfer(st_->params.device, want));
        }
        return xfer_.get();
    }

    bool use_p2p() const { return st_->dist.world > 1 && st_->p2p; }

    void allocate_vectors() {
        norm_sums_.ensure(2 * kMaxRhsTile * kMaxNormChunks);  // also the scratch word of box_barrier()
        release_peer_arena();
        if (use_p2p()) {
            build_peer_arena();
        } else {
            for (int k = 0; k <= n_levels_; ++k) {
                const size_t count = (size_t)std::max(lv_[k].n, 1) * K_;
                lv_[k].x.ensure(count), lv_[k].t.ensure(count), lv_[k].b.ensure(count), lv_[k].r.ensure(count);
            }
        }
        rhs64_.ensure((size_t)st_->n * K_);
        rhs64_.zero(stream_);  // multi-GPU: a rank uploads only the rows it reads
        if (max_halo_ && !use_p2p()) halo_send_.ensure(max_halo_ * K_), halo_recv_.ensure(max_halo_ * K_);
        if (sizeof(T) == 4) {
            r64_.ensure((size_t)st_->n * K_);
            x64_.ensure((size_t)st_->n * K_);
            const size_t nc = (size_t)lv_[n_levels_].n * K_;
            coarse_b64_.ensure(nc), coarse_x64_.ensure(nc);
        }
    }

    // ---- multi-GPU peer arena (peer_exchange.h): mailbox + the vectors of every level in one
    // allocation with the same layout on every rank, exported / imported with CUDA IPC.
    void build_peer_arena() {
        const DistLayout& d = st_->dist;
        if (d.world > kMaxPeers) throw std::invalid_argument("peer-memory halo exchange supports at most 8 ranks (one NVSwitch box)");
        if (!comm_) throw std::logic_error("multi-GPU layout configured but gmg_dist_init has not been called");
        auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
        size_t bytes = align(sizeof(PeerMailbox));
        std::vector<size_t> off;
        for (int k = 0; k <= n_levels_; ++k) {
            const size_t vb = align((size_t)std::max(lv_[k].n, 1) * K_ * sizeof(T));
            for (int j = 0; j < 4; ++j) off.push_back(bytes), bytes += vb;
        }
        GMG_CUDA(cudaMalloc(&arena_, bytes));
        arena_bytes_ = bytes;
        GMG_CUDA(cudaMemsetAsync(arena_, 0, bytes, stream_));
        peer_local_.ensure(4);
        GMG_CUDA(cudaMemsetAsync(peer_local_.ptr, 0, 4 * sizeof(unsigned long long), stream_));
        for (int k = 0; k <= n_levels_; ++k) {
            const size_t count = (size_t)std::max(lv_[k].n, 1) * K_;
            char* base = static_cast<char*>(arena_);
            lv_[k].x.view(reinterpret_cast<T*>(base + off[4 * k + 0]), count);
            lv_[k].t.view(reinterpret_cast<T*>(base + off[4 * k + 1]), count);
            lv_[k].b.view(reinterpret_cast<T*>(base + off[4 * k + 2]), count);
            lv_[k].r.view(reinterpret_cast<T*>(base + off[4 * k + 3]), count);
        }
        // every rank's mailbox must be zero before any peer can reach it
        GMG_CUDA(cudaStreamSynchronize(stream_));
        cudaIpcMemHandle_t mine;
        GMG_CUDA(cudaIpcGetMemHandle(&mine, arena_));
        std::vector<cudaIpcMemHandle_t> all(d.world);
        ipc_buf_.ensure((size_t)d.world * sizeof(cudaIpcMemHandle_t));
        GMG_CUDA(cudaMemcpyAsync(ipc_buf_.ptr + (size_t)d.rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, stream_));
        GMG_NCCL(nccl().AllGather(ipc_buf_.ptr + (size_t)d.rank * sizeof mine, ipc_buf_.ptr, sizeof mine, ncclChar, comm_, stream_));
        GMG_CUDA(cudaMemcpyAsync(all.data(), ipc_buf_.ptr, all.size() * sizeof mine, cudaMemcpyDeviceToHost, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
        fabric_ = PeerFabric();
        fabric_.rank = d.rank, fabric_.world = d.world;
        fabric_.box = static_cast<PeerMailbox*>(arena_);
        fabric_.epoch = peer_local_.ptr;
        fabric_.ticket = reinterpret_cast<unsigned int*>(peer_local_.ptr + 1);
        for (int q = 0; q < d.world; ++q) {
            if (q == d.rank) continue;
            GMG_CUDA(cudaIpcOpenMemHandle(&peer_base_[q], all[q], cudaIpcMemLazyEnablePeerAccess));
            fabric_.peer_delta[q] = static_cast<char*>(peer_base_[q]) - static_cast<char*>(arena_);
        }
        fabric_dev_.ensure(1);
        GMG_CUDA(cudaMemcpyAsync(fabric_dev_.ptr, &fabric_, sizeof fabric_, cudaMemcpyHostToDevice, stream_));
        box_barrier();  // nobody pushes before everybody has mapped everybody
    }

    void box_barrier() {
        GMG_CUDA(cudaMemsetAsync(norm_sums_.ptr, 0, sizeof(double), stream_));
        GMG_NCCL(nccl().AllReduce(norm_sums_.ptr, norm_sums_.ptr, 1, ncclDouble, ncclSum, comm_, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));
    }

    void release_peer_arena(bool collective = true) {
        if (!arena_) return;
        GMG_CUDA(cudaStreamSynchronize(stream_));
        for (auto& pb : peer_base_) {
            if (pb) cudaIpcCloseMemHandle(pb);
            pb = nullptr;
        }
        // peers unmap before the owner frees (every rank re-allocates at the same point of the program)
        if (collective && comm_) box_barrier();
        for (auto& l : lv_) l.x.release(), l.t.release(), l.b.release(), l.r.release();
        cudaFree(arena_);
        arena_ = nullptr, arena_bytes_ = 0;
        drop_graphs();
    }

    void require_mesh(bool need_positions = false) const {
        if (!mesh_.attached() || !pattern_ready_) throw std::logic_error("no mesh attached (gmg_mesh_attach)");
        if (need_positions && !mesh_has_pos_) throw std::logic_error("no vertex positions on the device (gmg_mesh_set_positions)");
    }

    // K right-hand sides from now on (vectors re-allocated, cycle rebuilt when it changes).
    void set_columns(int K) {
        if (K < 1 || K > kMaxRhsTile * kMaxNormChunks) throw std::invalid_argument("number of right-hand sides must be 1..32");
        if (K != K_) {
            K_ = K;
            if (pattern_ready_) allocate_vectors();
            invalidate_cycle();
        }
    }

    // Is (indptr, indices) different from the staged level-0 pattern (or is nothing staged)?
    bool indptr_changed(const int* indptr, const int* indices, int64_t n) const {
        if (!pattern_ready_ || !hierarchy_ready_ || st_->a_pat.empty() || st_->a_pat[0].rows != n) return true;
        const int64_t nnz = indptr[n];
        if ((int64_t)st_->a_pat[0].indices.size() != nnz) return true;
        return std::memcmp(st_->a_pat[0].indptr.data(), indptr, (n + 1) * sizeof(int)) != 0 ||
               std::memcmp(st_->a_pat[0].indices.data(), indices, nnz * sizeof(int)) != 0;
    }

    // Values and rhs of the system were written by device work queued on stream_.
    void mark_staged_on_device() {
        numeric_ready_ = false;
        rhs_pending_ = false;
        staged_ = true;
        solved_ = false;
        st_->transfer_timing["h2d_bytes"] = 0.0;
        st_->transfer_timing["pattern_reused"] = 1.0;
    }

    // The fp64 solution of the last solve, in HBM.
    const double* solution_device() {
        if (sizeof(T) == 4 && !refine()) launch_cast_f32_f64(reinterpret_cast<const float*>(x_final_), x64_.ptr, (size_t)st_->n * K_, stream_);
        return sizeof(T) == 4 ? x64_.ptr : reinterpret_cast<const double*>(x_final_);
    }

    void set_initial_guess() {
        const size_t count = (size_t)st_->n * K_;
        if (sizeof(T) == 8) {
            GMG_CUDA(cudaMemcpyAsync(lv_[0].b.ptr, rhs64_.ptr, count * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
            GMG_CUDA(cudaMemcpyAsync(lv_[0].x.ptr, rhs64_.ptr, count * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        } else if (refine()) {
            // mixed precision: the iterate lives in fp64 (x0 = rhs); the prologue forms its defect for the fp32 levels
            GMG_CUDA(cudaMemcpyAsync(x64_.ptr, rhs64_.ptr, count * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
        } else {
            launch_cast_f64_f32(rhs64_.ptr, reinterpret_cast<float*>(lv_[0].b.ptr), count, stream_);
            launch_cast_f64_f32(rhs64_.ptr, reinterpret_cast<float*>(lv_[0].x.ptr), count, stream_);
        }
    }

    // fp32 levels as the correction scheme of an fp64 iterate (defect correction): per cycle
    //     e = V-cycle_fp32(A, r) from a zero guess;  x += e (fp64);  r = b - A x (fp64 values, fp64 x)
    // and the stopping norm is the norm of that fp64 defect. Without it a cycle that carries x in fp32
    // stalls at the fp32 rounding floor of A x (1.7e-3 relative on BASELINE config 3, 5 M vertices).
    bool refine() const { return sizeof(T) == 4 && n_levels_ > 0 && st_->fp32_refine && st_->dist.world <= 1; }

    // ---- the V-cycle as a launch list ----------------------------------------------------
    SpmvArgs<T> base_args(const DevMat<T>& m) const {
        SpmvArgs<T> a;
        a.n_rows = m.rows;
        a.ld = K_;
        a.rowptr = m.indptr.ptr, a.colidx = m.colp(), a.vals = m.vals();
        if (m.diff && sizeof(T) == 8) a.vals = reinterpret_cast<const T*>(m.vdp()), a.diff = 1;
        a.l2_hint = st_->l2_hints ? m.l2_hint : 0;
        a.omega = (T)st_->params.omega;
        a.ctl = ctl_.ptr;
        return a;
    }

    const T* weight_ptr(int level, bool post, int sweep) const {
        return weights_.ptr + ((size_t)level * 2 + (post ? 1 : 0)) * kMaxSweeps + sweep;
    }

    // Multi-GPU: before an operator gathers `v` through matrix `hop` of a sharded level, fetch the
    // entries peers own (no-op on a single GPU and on replicated levels).
    void push_halo(int level, int hop, T* v) {
        const DistLayout& d = st_->dist;
        if (!d.sharded(level)) return;
        if (hop == HALO_P && !d.sharded(level + 1)) return;  // the coarse vector is replicated
        if (st_->dist_skip_exchange) return;  // measurement only
        Op op;
        op.kind = OP_HALO, op.level = level, op.halo_op = hop, op.vec = v;
        ops_.push_back(op);
    }

    // Sweeps [first, last) of the pre- or post-smoothing sequence of level k.
    void push_sweeps(int k, bool post, int first, int last, T*& cur, T*& alt) {
        for (int i = first; i < last; ++i) {
            push_halo(k, HALO_A, cur);
            Op op;
            op.kind = OP_JACOBI, op.level = k, op.epi = EPI_JACOBI, op.plan = &lv_[k].A.plan;
            op.args = base_args(lv_[k].A);
            op.args.x = cur, op.args.b = lv_[k].b.ptr, op.args.dinv = lv_[k].dinv.ptr, op.args.out = alt;
            op.args.omega_ptr = weight_ptr(k, post, i);
            ops_.push_back(op);
            std::swap(cur, alt);
        }
    }

    // One cycle at level k: multiGridVCycleGS (multigrid_solver.cpp:1059-1088; type 0), multiGridFCycleGS
    // (:1091-1140; type 1) or multiGridWCycleGS (:1143-1192; type 2). `cur` holds x_k on entry and on exit;
    // `pre_done` sweeps were already applied by the caller (zero-guess shortcut). F and W repeat
    // residual / restriction / recursion / prolongation / post-smoothing; the second recursion (V inside F,
    // W inside W) starts from the eps of the first one, as upstream (eps is not reset at :1126 / :1178), and
    // uses the first recursion's test for the coarsest level (upstream's `k == DoF.size() - 2` is off by one).
    void push_vcycle(int k, T*& cur, T*& alt, int pre_done, bool fused_top, int type = 0) {
        const gmg_params& p = st_->params;
        const int L = n_levels_;
        Level& f = lv_[k];
        Level& c = lv_[k + 1];
        push_sweeps(k, false, pre_done, p.pre_iters, cur, alt);
        const bool next_is_coarsest = (k + 1 == L);
        const int halves = type == 0 ? 1 : 2;
        T* ccur = c.x.ptr;
        T* calt = c.t.ptr;
        for (int half = 0; half < halves; ++half) {
            push_halo(k, HALO_A, cur);
            {   // res = b - A x
                Op op;
                op.kind = OP_RESIDUAL, op.level = k, op.epi = EPI_RESIDUAL, op.plan = &f.A.plan;
                op.args = base_args(f.A);
                op.args.x = cur, op.args.b = f.b.ptr, op.args.out = f.r.ptr;
                ops_.push_back(op);
            }
            // resRest = U^T res ; eps = 0 (first recursion). With a zero guess the first Jacobi sweep on the
            // next level is eps = omega D^-1 resRest, which the restriction writes as a by-product.
            const bool zero_guess = half == 0;
            const int next_pre_done = (zero_guess && !next_is_coarsest && p.pre_iters >= 1) ? 1 : 0;
            push_halo(k, HALO_R, f.r.ptr);
            {
                Op op;
                op.kind = OP_RESTRICT, op.level = k, op.epi = EPI_SPMV, op.plan = &f.R.plan;
                op.args = base_args(f.R);
                op.args.x = f.r.ptr, op.args.out = c.b.ptr;
                if (next_pre_done) op.args.out2 = c.x.ptr, op.args.dinv = c.dinv.ptr, op.args.omega_ptr = weight_ptr(k + 1, false, 0);
                ops_.push_back(op);
            }
            if (st_->dist.sharded(k) && !st_->dist.sharded(k + 1)) {
                // every rank restricted its own coarse rows; the next level is replicated
                Op op;
                op.kind = OP_ALLGATHER, op.level = k + 1, op.vec = c.b.ptr, op.vec2 = next_pre_done ? c.x.ptr : nullptr;
                ops_.push_back(op);
            }
            if (k + 1 == tail_level_) tail_begin_ = ops_.size();
            if (next_is_coarsest) {
                Op op;
                op.kind = OP_COARSE, op.level = k + 1;
                ops_.push_back(op);
            } else {
                if (zero_guess && !next_pre_done) {
                    Op op;
                    op.kind = OP_ZERO, op.level = k + 1, op.zero_ptr = ccur, op.zero_bytes = (size_t)c.n * K_ * sizeof(T);
                    ops_.push_back(op);
                }
                push_vcycle(k + 1, ccur, calt, next_pre_done, false, half == 0 ? type : (type == 1 ? 0 : 2));
            }
            if (k + 1 == tail_level_) tail_end_ = ops_.size();
            push_halo(k, HALO_P, ccur);
            {   // x = x + U eps. On level 0 an odd sweep count is evened out by writing the last prolongation
                // to the other buffer, so a cycle always ends in the buffer it started from (graph replay).
                Op op;
                op.kind = OP_PROLONG, op.level = k, op.epi = EPI_ADD, op.plan = &f.P.plan;
                op.args = base_args(f.P);
                op.args.x = ccur, op.args.xin = cur;
                // swaps of this level: (pre - pre_done) + halves * post sweeps [+ 1 if the prolongation flips].
                // plain cycle: end where it started (even); fused top level: start in t, end in x (odd)
                const int sweeps_here = p.pre_iters - pre_done + halves * p.post_iters;
                const bool flip = (k == 0) && half == halves - 1 && (fused_top ? sweeps_here % 2 == 0 : sweeps_here % 2 != 0);
                op.args.out = flip ? alt : cur;
                ops_.push_back(op);
                if (flip) std::swap(cur, alt);
            }
            push_sweeps(k, true, 0, p.post_iters, cur, alt);
        }
    }

    void build_cycle() {
        drop_graphs();
        ops_.clear();
        prologue_.clear();
        const gmg_params& p = st_->params;
        T* cur = lv_[0].x.ptr;
        T* alt = lv_[0].t.ptr;
        // The stopping test of cycle i and the first pre-smoothing sweep of cycle i + 1 both need
        // b - A x of the same iterate: one kernel does both (EPI_NORMJAC). The sweep is speculative
        // (written to the other buffer), so the final iterate survives when the loop stops.
        const bool fused = n_levels_ > 0 && p.pre_iters >= 1 && st_->fuse_norm;
        // levels small enough to live in L2 and be launch-latency bound run as one persistent
        // kernel (tail_kernel.cuh); never the finest level, whose streaming kernels are better
        tail_level_ = -1, tail_begin_ = tail_end_ = 0;
        tail_cluster_ = false;
        for (int k = 1; k <= n_levels_ && st_->tail_rows > 0 && st_->dist.world <= 1 && p.cycle_type == 0; ++k)
            if (lv_[k].n <= st_->tail_rows) {
                tail_level_ = k;
                break;
            }
        // ... or, the default: the levels of a few thousand rows as one thread-block cluster (cluster_tail.cuh)
        for (int k = 1; k < n_levels_ && tail_level_ < 0 && st_->cluster_tail_rows > 0 && st_->dist.world <= 1 && p.cycle_type == 0; ++k)
            if (lv_[k].n <= st_->cluster_tail_rows) {
                tail_level_ = k;
                tail_cluster_ = true;
                break;
            }
        if (n_levels_ == 0) {
            // no hierarchy could be built (N <= lower_bound): the reference is undefined here
            // (SURVEY Appendix A.13); the whole system goes to the direct coarse solve.
            Op op;
            op.kind = OP_COARSE, op.level = 0;
            ops_.push_back(op);
        } else if (refine()) {
            Op pr;  // prologue: defect of the initial guess, no correction to add, no stopping test
            pr.kind = OP_REFINE, pr.level = -1;
            prologue_.push_back(pr);
            Op z;
            z.kind = OP_ZERO, z.level = 0, z.zero_ptr = cur, z.zero_bytes = (size_t)lv_[0].n * K_ * sizeof(T);
            ops_.push_back(z);
            push_vcycle(0, cur, alt, 0, false, p.cycle_type);
            Op rf;  // x += e, new defect, stopping test
            rf.kind = OP_REFINE, rf.level = 0, rf.vec = cur;
            ops_.push_back(rf);
        } else if (!fused) {
            push_vcycle(0, cur, alt, 0, false, p.cycle_type);
        } else {
            // prologue (once per solve): sweep 0 of the first cycle, x (cur) -> alt
            push_sweeps(0, false, 0, 1, cur, alt);
            prologue_.push_back(ops_.back());
            ops_.clear();
            // every cycle starts from the swept iterate in `cur` (= lv_[0].t) and must end with the
            // final iterate in lv_[0].x so that the speculative sweep lands in lv_[0].t again
            push_vcycle(0, cur, alt, 1, true, p.cycle_type);
        }
        if (tail_level_ > 0 && tail_end_ > tail_begin_) collapse_tail();
        x_final_ = x_final_cycle_ = cur;
        if (n_levels_ > 0 && !refine()) push_halo(0, HALO_A, cur);
        if (!refine())
        {   // residualCheck(LHS, b, x, stoppingCriteria) (multigrid_solver.cpp:1413)
            Op op;
            op.kind = OP_NORM, op.level = 0, op.epi = fused ? EPI_NORMJAC : EPI_NORM, op.plan = &lv_[0].A.plan;
            op.args = base_args(lv_[0].A);
            op.args.x = cur, op.args.b = lv_[0].b.ptr;
            if (fused) op.args.dinv = lv_[0].dinv.ptr, op.args.omega_ptr = weight_ptr(0, false, 0), op.args.out = alt;
            ops_.push_back(op);
        }
        if (use_p2p() && (st_->p2p_fuse || st_->dist_skip_exchange)) fuse_exchanges();
        cycle_dirty_ = false;
        // one-time per-kernel attribute/occupancy calls must not land inside a stream capture
        set_launch_dry_run(true);
        try {
            for (const Op& op : prologue_)
                if (op.plan) run_op(op, stream_, 0);
            for (const Op& op : ops_)
                if (op.plan) run_op(op, stream_, 0);
        } catch (...) {
            set_launch_dry_run(false);
            throw;
        }
        set_launch_dry_run(false);
    }

    // Multi-GPU: fold every halo push into the kernel that produces the vector (its epilogue also
    // stores the rows peers gather into the peers' HBM and the last CTA signals) and the wait into
    // the kernel that gathers it. A push stays a kernel of its own when producer or consumer is
    // not a row-product kernel (all-gather into the first replicated level, memset, coarse solve).
    void fuse_exchanges() {
        auto writes = [](const Op& op, const T* v) {
            if (op.kind == OP_ZERO) return op.zero_ptr == (const void*)v;
            if (op.kind == OP_ALLGATHER) return op.vec == v || op.vec2 == v;
            if (op.kind == OP_HALO || op.kind == OP_TAIL) return false;
            if (op.kind == OP_COARSE) return false;  // writes the coarsest x, which is never exchanged
            return op.args.out == v || op.args.out2 == v;
        };
        auto row_product = [](const Op& op) {
            return op.plan != nullptr && (op.kind == OP_JACOBI || op.kind == OP_RESIDUAL || op.kind == OP_RESTRICT ||
                                          op.kind == OP_PROLONG || op.kind == OP_NORM);
        };
        auto mark_producer = [&](Op& w, const T* v, const unsigned char* mask) {
            if (w.args.send_mask && w.args.send_mask != mask) return false;
            w.args.fabric = fabric_dev_.ptr, w.args.send_mask = mask, w.args.push_out2 = (w.args.out2 == v) ? 1 : 0;
            return true;
        };
        for (size_t ih = 0; ih < ops_.size();) {
            Op& h = ops_[ih];
            if (h.kind != OP_HALO || ih + 1 >= ops_.size()) {
                ++ih;
                continue;
            }
            Op& c = ops_[ih + 1];
            const T* v = h.vec;
            const DevHalo& dh = halo_[h.halo_op][h.level];
            bool ok = row_product(c) && c.args.x == v && dh.mask.ptr != nullptr;
            // most recent writer of v: earlier in the cycle, else the end of the previous cycle (wrap)
            int writer = -1;
            bool wrapped = false;
            for (int j = (int)ih - 1; ok && j >= 0 && writer < 0; --j)
                if (writes(ops_[j], v)) writer = j;
            for (int j = (int)ops_.size() - 1; ok && j > (int)ih && writer < 0; --j)
                if (writes(ops_[j], v)) writer = j, wrapped = true;
            if (ok && writer >= 0 && !(row_product(ops_[writer]) && ops_[writer].kind != OP_NORM) &&
                !(ops_[writer].kind == OP_NORM && ops_[writer].epi == EPI_NORMJAC))
                ok = false;
            if (ok && writer >= 0) ok = mark_producer(ops_[writer], v, dh.mask.ptr);
            if (ok && (writer < 0 || wrapped)) {
                // first cycle of a solve: the prologue sweep (if it writes v) pushes; otherwise v is the
                // initial guess, identical on every rank
                for (Op& pr : prologue_)
                    if (writes(pr, v) && !mark_producer(pr, v, dh.mask.ptr)) ok = false;
            }
            if (st_->dist_skip_exchange) {  // measurement only: the cost of the partition without any exchange (wrong results)
                ops_.erase(ops_.begin() + ih);
                continue;
            }
            if (!ok) {
                ++ih;
                continue;
            }
            c.args.fabric = fabric_dev_.ptr, c.args.wait_peers = 1;
            ops_.erase(ops_.begin() + ih);
        }
    }

    int tail_grid() {
        if (!tail_grid_) {
            int dev = 0;
            GMG_CUDA(cudaGetDevice(&dev));
            GMG_CUDA(cudaDeviceGetAttribute(&tail_grid_, cudaDevAttrMultiProcessorCount, dev));
        }
        return tail_grid_;
    }

    // Replace ops_[tail_begin_, tail_end_) — everything between the restriction into the tail
    // level and the prolongation out of it — by one OP_TAIL whose operator table is on the device.
    void collapse_tail() {
        std::vector<TailOp<T>> table;
        for (size_t i = tail_begin_; i < tail_end_; ++i) {
            const Op& op = ops_[i];
            if (op.kind == OP_ZERO) {
                TailOp<T> z;
                z.kind = TAIL_ZERO, z.dst = op.zero_ptr, z.count = op.zero_bytes;
                table.push_back(z);
            } else if (op.kind == OP_COARSE) {
                Level& c = lv_[op.level];
                const size_t count = (size_t)c.n * K_;
                const double* b64 = reinterpret_cast<const double*>(c.b.ptr);
                double* x64 = reinterpret_cast<double*>(c.x.ptr);
                if (sizeof(T) == 4) {
                    TailOp<T> cast;
                    cast.kind = TAIL_TO_F64, cast.src = c.b.ptr, cast.dst = coarse_b64_.ptr, cast.count = count;
                    table.push_back(cast);
                    b64 = coarse_b64_.ptr, x64 = coarse_x64_.ptr;
                }
                for (int k0 = 0; k0 < K_; k0 += kMaxRhsTile) {
                    const int kt = std::min(kMaxRhsTile, K_ - k0);
                    TailOp<T> u;  // y = W b: row i of W is column i of Wt
                    u.kind = TAIL_COLDOT, u.kcols = kt, u.M = coarse_.wt(), u.ldm = coarse_.ld(), u.n = c.n, u.upper = 1;
                    u.v = b64 + k0, u.v_ld = K_, u.out = coarse_.y(), u.out_ld = kt;
                    table.push_back(u);
                    TailOp<T> l;  // x = W^T y
                    l.kind = TAIL_COLDOT, l.kcols = kt, l.M = coarse_.w(), l.ldm = coarse_.ld(), l.n = c.n, l.upper = 0;
                    l.v = coarse_.y(), l.v_ld = kt, l.out = x64 + k0, l.out_ld = K_;
                    table.push_back(l);
                }
                if (sizeof(T) == 4) {
                    TailOp<T> cast;
                    cast.kind = TAIL_FROM_F64, cast.src = coarse_x64_.ptr, cast.dst = c.x.ptr, cast.count = count;
                    table.push_back(cast);
                }
            } else {
                for (int k0 = 0; k0 < K_; k0 += kMaxRhsTile) {
                    TailOp<T> r;
                    r.kind = TAIL_ROWS, r.epi = op.epi, r.kcols = std::min(kMaxRhsTile, K_ - k0);
                    {   // threads per row: as many as fill the grid in one pass, at most the row length
                        const int threads = tail_grid() * kTailThreads;
                        int lanes = 1;
                        while (lanes < 8 && lanes * 2 <= op.plan->lanes && (int64_t)op.args.n_rows * lanes * 2 <= threads) lanes *= 2;
                        r.lanes = lanes;
                    }
                    r.a = op.args;
                    r.a.x = op.args.x + k0;
                    if (r.a.b) r.a.b = op.args.b + k0;
                    if (r.a.xin) r.a.xin = op.args.xin + k0;
                    r.a.out = op.args.out + k0;
                    if (r.a.out2) r.a.out2 = op.args.out2 + k0;
                    table.push_back(r);
                }
            }
        }
        if (tail_cluster_ && !plan_cluster_tail(table)) {
            tail_level_ = -1;  // does not fit one cluster's shared memory: the per-operator kernels stay
            return;
        }
        tail_table_.ensure(table.size() * sizeof(TailOp<T>));
        GMG_CUDA(cudaMemcpyAsync(tail_table_.ptr, table.data(), table.size() * sizeof(TailOp<T>), cudaMemcpyHostToDevice, stream_));
        GMG_CUDA(cudaStreamSynchronize(stream_));  // `table` is a local
        n_tail_ops_ = (int)table.size();
        cluster_args_.ops = tail_table_.ptr, cluster_args_.n_ops = n_tail_ops_, cluster_args_.ctl = ctl_.ptr;
        Op tail;
        tail.kind = OP_TAIL, tail.level = tail_level_;
        ops_.erase(ops_.begin() + tail_begin_, ops_.begin() + tail_end_);
        ops_.insert(ops_.begin() + tail_begin_, tail);
    }

    // Cluster tail: number the sparse operators of the table, pick the cluster size and the shared memory one CTA
    // needs for its slabs of all of them (worst CTA), choose the threads per row. False: does not fit.
    bool plan_cluster_tail(std::vector<TailOp<T>>& table) {
        cluster_args_ = ClusterTailArgs();
        std::vector<const std::vector<int>*> host_rp;
        for (TailOp<T>& op : table) {
            if (op.kind != TAIL_ROWS) continue;
            int m = -1;
            for (int j = 0; j < cluster_args_.n_mats; ++j)
                if (cluster_args_.mats[j].rowptr == op.a.rowptr && cluster_args_.mats[j].vals == (const void*)op.a.vals) m = j;
            if (m < 0) {
                if (cluster_args_.n_mats == kClusterTailMaxMats) return false;
                const std::vector<int>* rp = nullptr;
                for (int k = 0; k <= n_levels_ && !rp; ++k) {
                    if (op.a.rowptr == lv_[k].A.indptr.ptr) rp = &st_->a_pat[k].indptr;
                    if (k < n_levels_ && op.a.rowptr == lv_[k].P.indptr.ptr) rp = &st_->hier.U[k].indptr;
                    if (k < n_levels_ && op.a.rowptr == lv_[k].R.indptr.ptr) rp = &st_->r_host[k].indptr;
                }
                if (!rp) return false;
                m = cluster_args_.n_mats++;
                cluster_args_.mats[m].rowptr = op.a.rowptr, cluster_args_.mats[m].colidx = op.a.colidx;
                cluster_args_.mats[m].vals = op.a.vals, cluster_args_.mats[m].n_rows = op.a.n_rows;
                host_rp.push_back(rp);
            }
            op.mat = m;
        }
        auto kernel = cluster_tail_kernel<T>;
        static bool opted_in = false;
        if (!opted_in) {
            cudaFuncAttributes attr;
            GMG_CUDA(cudaFuncGetAttributes(&attr, kernel));
            GMG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(staged_smem_limit() - attr.sharedSizeBytes)));
            cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaGetLastError();
            opted_in = true;
        }
        cudaFuncAttributes attr;
        GMG_CUDA(cudaFuncGetAttributes(&attr, kernel));
        const char* forced = std::getenv("GMG_CLUSTER_SIZE");  // measurement / debugging aid: 8 = portable cluster size only
        for (int C : {16, 8}) {
            if (forced && std::atoi(forced) != C) continue;
            size_t worst = 0;
            for (int r = 0; r < C; ++r) {
                size_t bytes = 128;
                for (int m = 0; m < cluster_args_.n_mats; ++m) {
                    const std::vector<int>& rp = *host_rp[m];
                    const int n = cluster_args_.mats[m].n_rows;
                    const int r0 = (int)((long long)n * r / C), r1 = (int)((long long)n * (r + 1) / C);
                    const size_t n_rp = (size_t)(((r1 + 1 + 3) & ~3) - (r0 & ~3));
                    const size_t cnt = (size_t)(((rp[r1] + 3) & ~3) - (rp[r0] & ~3));
                    bytes += n_rp * 4 + ((cnt * sizeof(T) + 15) & ~(size_t)15) + ((cnt * 4 + 15) & ~(size_t)15);
                }
                worst = std::max(worst, bytes);
            }
            if (worst + attr.sharedSizeBytes > staged_smem_limit()) continue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)C, 1, 1), cfg.blockDim = dim3(kClusterTailThreads, 1, 1), cfg.dynamicSmemBytes = worst;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)C, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
            cfg.attrs = at, cfg.numAttrs = 1;
            int n_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&n_clusters, kernel, &cfg) != cudaSuccess || n_clusters < 1) {
                cudaGetLastError();
                continue;
            }
            cluster_size_ = C, cluster_smem_ = worst;
            // threads per row: as many as fill the cluster in one pass, at most 8 and at most the row length
            for (TailOp<T>& op : table) {
                if (op.kind != TAIL_ROWS) continue;
                const int64_t threads = (int64_t)C * kClusterTailThreads;
                const double avg = op.a.n_rows ? (double)((*host_rp[op.mat])[op.a.n_rows]) / op.a.n_rows : 1.0;
                int lanes = 1;
                while (lanes < 8 && lanes * 2 <= avg && (int64_t)op.a.n_rows * lanes * 2 <= threads) lanes *= 2;
                op.lanes = lanes;
            }
            return true;
        }
        return false;
    }

    int run_op(const Op& op, cudaStream_t s, unsigned long long cond) {
        int launches = 0;
        const gmg_params& p = st_->params;
        switch (op.kind) {
            case OP_ZERO:
                GMG_CUDA(cudaMemsetAsync(op.zero_ptr, 0, op.zero_bytes, s));
                break;
            case OP_HALO:
                if (use_p2p()) {
                    const DevHalo& dh = halo_[op.halo_op][op.level];
                    PeerPushArgs<T> a;
                    a.v = op.vec, a.K = K_;
                    for (int q = 0; q < st_->dist.world; ++q)
                        if (!dh.n_send.empty() && dh.n_send[q]) a.idx[q] = dh.send_idx[q].ptr, a.count[q] = dh.n_send[q];
                    launch_peer_push<T>(a, fabric_, ctl_.ptr, s);
                    launches += 1;
                } else {
                    exchange_halo(halo_[op.halo_op][op.level], op.vec, s);
                    launches += 2;
                }
                break;
            case OP_ALLGATHER:
                if (use_p2p()) {
                    const DistLayout& d = st_->dist;
                    PeerPushArgs<T> a;
                    a.v = op.vec, a.v2 = op.vec2, a.K = K_;
                    for (int q = 0; q < d.world; ++q)
                        if (q != d.rank) a.first[q] = (int)d.begin(op.level), a.count[q] = (int)(d.end(op.level) - d.begin(op.level));
                    launch_peer_push<T>(a, fabric_, ctl_.ptr, s);
                    launches += 1;
                } else {
                    allgather_rows(op.level, op.vec, s);
                    if (op.vec2) allgather_rows(op.level, op.vec2, s);
                }
                break;
            case OP_REFINE: {
                const size_t count = (size_t)st_->n * K_;
                if (op.level >= 0) launch_add_f32_to_f64(reinterpret_cast<const float*>(op.vec), x64_.ptr, count, s), ++launches;
                SpmvPlan plan;  // fp64 defect straight from global memory (the row tiles are sized for fp32 slabs)
                plan.path = 1, plan.lanes = lv_[0].A.plan.lanes;
                SpmvArgs<double> a;
                a.n_rows = lv_[0].n, a.ld = K_;
                a.rowptr = lv_[0].A.indptr.ptr, a.colidx = lv_[0].A.colp(), a.vals = lv_[0].A.v64p();
                if (lv_[0].A.diff) a.vals = lv_[0].A.vdp(), a.diff = 1;
                a.weight = p.stopping_criteria == 2 ? mass_.ptr : p.stopping_criteria == 1 ? minv_.ptr : nullptr;
                a.ctl = ctl_.ptr;
                const bool test = op.level >= 0;
                const bool fused_test = test && K_ <= kMaxRhsTile && st_->fuse_stop;
                if (fused_test) {
                    a.fin_ticket = tail_bar_.ptr + 2;
                    a.hist_res = hist_res_.ptr, a.hist_ms = hist_ms_.ptr, a.cond_handle = cond;
                }
                NormChunks chunks;
                for (int k0 = 0; k0 < K_; k0 += kMaxRhsTile) {
                    const int kt = std::min(kMaxRhsTile, K_ - k0);
                    a.x = x64_.ptr + k0, a.b = rhs64_.ptr + k0, a.out = r64_.ptr + k0;
                    a.partials = partials_.ptr + (size_t)chunks.n_chunks * kNormChunkStride;
                    chunks.kt[chunks.n_chunks] = kt;
                    chunks.n_blocks[chunks.n_chunks] = launch_spmv<double>(EPI_RESNORM, kt, a, plan, s);
                    ++chunks.n_chunks;
                    ++launches;
                }
                launch_cast_f64_f32(r64_.ptr, reinterpret_cast<float*>(lv_[0].b.ptr), count, s);
                ++launches;
                if (test && !fused_test) {
                    launch_norm_finalize(partials_.ptr, chunks, ctl_.ptr, hist_res_.ptr, hist_ms_.ptr, 1, cond, s);
                    ++launches;
                }
                break;
            }
            case OP_TAIL: {
                if (tail_cluster_) {
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3((unsigned)cluster_size_, 1, 1), cfg.blockDim = dim3(kClusterTailThreads, 1, 1);
                    cfg.dynamicSmemBytes = cluster_smem_, cfg.stream = s;
                    cudaLaunchAttribute at[2];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = (unsigned)cluster_size_, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
                    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[1].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at, cfg.numAttrs = st_->use_pdl ? 2 : 1;
                    GMG_CUDA(cudaLaunchKernelEx(&cfg, cluster_tail_kernel<T>, cluster_args_));
                    launches += 1;
                    break;
                }
                const int sms = tail_grid();
                tail_kernel<T><<<sms, kTailThreads, 0, s>>>(reinterpret_cast<const TailOp<T>*>(tail_table_.ptr), n_tail_ops_, tail_bar_.ptr);
                GMG_CUDA(cudaGetLastError());
                launches += 1;
                break;
            }
            case OP_COARSE: {
                Level& c = lv_[op.level];
                if (sizeof(T) == 8) {
                    coarse_.solve(reinterpret_cast<const double*>(c.b.ptr), reinterpret_cast<double*>(c.x.ptr), K_, K_, ctl_.ptr, s);
                } else {
                    const size_t count = (size_t)c.n * K_;
                    launch_cast_f32_f64(reinterpret_cast<const float*>(c.b.ptr), coarse_b64_.ptr, count, s);
                    coarse_.solve(coarse_b64_.ptr, coarse_x64_.ptr, K_, K_, ctl_.ptr, s);
                    launch_cast_f64_f32(coarse_x64_.ptr, reinterpret_cast<float*>(c.x.ptr), count, s);
                    launches += 2;
                }
                launches += 2 * ((K_ + kMaxRhsTile - 1) / kMaxRhsTile);
                break;
            }
            case OP_NORM: {
                NormChunks chunks;
                SpmvArgs<T> a = op.args;
                a.weight = p.stopping_criteria == 2 ? mass_.ptr : p.stopping_criteria == 1 ? minv_.ptr : nullptr;
                // single GPU, one column pass: the kernel's last CTA is the stopping test
                const bool fused_test = st_->dist.world <= 1 && K_ <= kMaxRhsTile && st_->fuse_stop;
                if (fused_test) {
                    a.fin_ticket = tail_bar_.ptr + 2;
                    a.hist_res = hist_res_.ptr, a.hist_ms = hist_ms_.ptr, a.cond_handle = cond;
                }
                for (int k0 = 0; k0 < K_; k0 += kMaxRhsTile) {
                    const int kt = std::min(kMaxRhsTile, K_ - k0);
                    a.x = op.args.x + k0, a.b = op.args.b + k0;
                    a.partials = partials_.ptr + (size_t)chunks.n_chunks * kNormChunkStride;
                    chunks.kt[chunks.n_chunks] = kt;
                    if (op.args.out) a.out = op.args.out + k0;
                    chunks.n_blocks[chunks.n_chunks] = launch_spmv<T>(op.epi, kt, a, *op.plan, s);
                    ++chunks.n_chunks;
                    ++launches;
                }
                if (fused_test) break;
                if (use_p2p()) {
                    // rows are split across ranks: partial sums travel through the peers' mailboxes
                    launch_peer_norm(partials_.ptr, chunks, K_, fabric_, ctl_.ptr, hist_res_.ptr, hist_ms_.ptr, cond, s);
                    ++launches;
                } else if (st_->dist.world > 1) {
                    // rows are split across ranks: sum the partial sums over the box first
                    launch_norm_partial_sums(partials_.ptr, chunks, norm_sums_.ptr, s);
                    GMG_NCCL(nccl().AllReduce(norm_sums_.ptr, norm_sums_.ptr, (size_t)2 * K_, ncclDouble, ncclSum, comm_, s));
                    launch_norm_finalize_sums(norm_sums_.ptr, K_, ctl_.ptr, hist_res_.ptr, hist_ms_.ptr, s);
                    launches += 2;
                } else {
                    launch_norm_finalize(partials_.ptr, chunks, ctl_.ptr, hist_res_.ptr, hist_ms_.ptr, 1, cond, s);
                    ++launches;
                }
                break;
            }
            default: {
                for (int k0 = 0; k0 < K_; k0 += kMaxRhsTile) {
                    const int kt = std::min(kMaxRhsTile, K_ - k0);
                    SpmvArgs<T> a = op.args;
                    a.x = op.args.x + k0;
                    if (a.b) a.b = op.args.b + k0;
                    if (a.xin) a.xin = op.args.xin + k0;
                    a.out = op.args.out + k0;
                    if (a.out2) a.out2 = op.args.out2 + k0;
                    launch_spmv<T>(op.epi, kt, a, *op.plan, s);
                    ++launches;
                }
            }
        }
        return launches;
    }

    void run_cycle(cudaStream_t s, unsigned long long cond, bool profile) {
        int launches = 0;
        for (const Op& op : ops_) {
            cudaEvent_t a = nullptr, b = nullptr;
            if (profile) {
                a = next_prof_event(), b = next_prof_event();
                GMG_CUDA(cudaEventRecord(a, s));
            }
            launches += run_op(op, s, cond);
            if (profile) {
                GMG_CUDA(cudaEventRecord(b, s));
                prof_pending_.push_back({op.kind, op.level, a, b});
            }
        }
        launches_per_cycle_ = launches;
    }

    void build_cycle_graph() {
        cudaGraph_t g = nullptr;
        GMG_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
        try {
            run_cycle(stream_, 0, false);
        } catch (...) {
            cudaStreamEndCapture(stream_, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        GMG_CUDA(cudaStreamEndCapture(stream_, &g));
        GMG_CUDA(cudaGraphInstantiate(&cycle_exec_, g, 0));
        GMG_CUDA(cudaGraphDestroy(g));
    }

    // One graph = the whole do { V-cycle; residualCheck } while (...) loop: a conditional WHILE
    // node whose body is the captured cycle; the finalize kernel of the stopping test sets the
    // loop condition on the device, so the host launches once per solve.
    void build_while_graph() {
        cudaGraph_t g = nullptr;
        GMG_CUDA(cudaGraphCreate(&g, 0));
        cudaGraphConditionalHandle handle;
        GMG_CUDA(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        GMG_CUDA(cudaGraphAddNode(&node, g, nullptr, 0, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        GMG_CUDA(cudaStreamBeginCaptureToGraph(stream_, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        cudaGraph_t captured = nullptr;
        try {
            run_cycle(stream_, (unsigned long long)handle, false);
        } catch (...) {
            cudaStreamEndCapture(stream_, &captured);
            cudaGraphDestroy(g);
            throw;
        }
        GMG_CUDA(cudaStreamEndCapture(stream_, &captured));
        GMG_CUDA(cudaGraphInstantiate(&while_exec_, g, 0));
        GMG_CUDA(cudaGraphDestroy(g));
    }

    void drop_graphs() {
        if (cycle_exec_) cudaGraphExecDestroy(cycle_exec_);
        if (while_exec_) cudaGraphExecDestroy(while_exec_);
        cycle_exec_ = nullptr, while_exec_ = nullptr;
    }

    cudaEvent_t next_prof_event() {
        if (prof_next_ == prof_events_.size()) {
            cudaEvent_t e;
            GMG_CUDA(cudaEventCreate(&e));
            prof_events_.push_back(e);
        }
        return prof_events_[prof_next_++];
    }
    void collect_profile() {
        for (const auto& pe : prof_pending_) {
            float ms = 0;
            GMG_CUDA(cudaEventElapsedTime(&ms, pe.a, pe.b));
            ProfileSlot& slot = prof_[pe.kind][std::min(pe.level, kMaxLevels - 1)];
            slot.ms += ms;
            slot.launches += 1;
        }
        prof_pending_.clear();
        prof_next_ = 0;
    }

    struct PendingEvent {
        int kind, level;
        cudaEvent_t a, b;
    };

    SolverState* st_;
    std::unique_ptr<HostTransfer> xfer_;
    cudaStream_t stream_ = nullptr, stream2_ = nullptr;
    cudaEvent_t rhs_ready_ = nullptr;
    bool rhs_pending_ = false;
    cudaEvent_t ev_[4] = {};
    std::vector<Level> lv_;
    int n_levels_ = 0;
    int K_ = 0;
    DenseCoarseSolver coarse_;
    DeviceBuffer<CycleControl> ctl_;
    CycleControl* ctl_host_ = nullptr;
    DeviceBuffer<double> hist_res_, hist_ms_, partials_, mass_, minv_, rhs64_, x64_, r64_, coarse_b64_, coarse_x64_, io64_;
    DeviceBuffer<double> rho_, weights64_;
    DeviceBuffer<unsigned long long> trace_buf_;
    DeviceBuffer<T> weights_;
    DeviceBuffer<int> q_indptr_, q_indices_;
    DeviceBuffer<double> q_vals_, q_b_, q_x_, q_vd_, q_dinv_, q_rho_;
    DeviceBuffer<double> kry_x_, kry_p_, kry_q_, kry_part_;   // conjugate-gradient wrapper (option krylov)
    DeviceBuffer<PcgScalars> kry_sc_;
    std::vector<Op> kry_ops_;
    T* kry_z_ = nullptr;
    T* kry_z0_ = nullptr;
    bool kry_ops_dirty_ = true;
    int kry_K_ = 0;
    bool window0_ = false;   // multi-GPU: finest level stored by row segments (decided from the layout: identical on all ranks)
    std::vector<std::pair<int64_t, int64_t>> rhs_rows_;   // row ranges of the right-hand side this rank uploads (all rows on a single GPU)
    DeviceBuffer<int> flag_dev_;
    MeshAssembler mesh_;               // device-side operator assembly (mesh_assembly.h)
    DeviceBuffer<double> mesh_pos_, mesh_s_, mesh_m_, mesh_y_;
    bool mesh_has_pos_ = false, mesh_has_s_ = false, mesh_has_m_ = false;
    std::vector<DevHalo> halo_[3];
    size_t max_halo_ = 0;
    DeviceBuffer<T> halo_send_, halo_recv_;
    DeviceBuffer<double> norm_sums_;
    ncclComm_t comm_ = nullptr;
    void* arena_ = nullptr;            // peer arena (multi-GPU, option p2p)
    size_t arena_bytes_ = 0;
    void* peer_base_[kMaxPeers] = {};  // peers' arenas mapped into this process
    PeerFabric fabric_;
    DeviceBuffer<PeerFabric> fabric_dev_;
    DeviceBuffer<unsigned long long> peer_local_;
    DeviceBuffer<unsigned char> ipc_buf_;
    bool hierarchy_ready_ = false, pattern_ready_ = false, staged_ = false, solved_ = false, cycle_dirty_ = true;
    bool numeric_ready_ = false;
    std::vector<Op> ops_, prologue_;
    int tail_level_ = -1, n_tail_ops_ = 0, tail_grid_ = 0;
    bool tail_cluster_ = false;        // the tail runs as one thread-block cluster (cluster_tail.cuh), not as a grid with L2 barriers
    ClusterTailArgs cluster_args_;
    int cluster_size_ = 0;
    size_t cluster_smem_ = 0;
    size_t tail_begin_ = 0, tail_end_ = 0;
    DeviceBuffer<unsigned char> tail_table_;
    DeviceBuffer<unsigned> tail_bar_;
    T* x_final_ = nullptr;
    T* x_final_cycle_ = nullptr;       // where the plain cycle loop leaves the iterate (x_final_ may point at the Krylov iterate)
    int launches_per_cycle_ = 0;
    cudaGraphExec_t cycle_exec_ = nullptr, while_exec_ = nullptr;
    std::vector<cudaEvent_t> prof_events_;
    size_t prof_next_ = 0;
    std::vector<PendingEvent> prof_pending_;
    ProfileSlot prof_[OP_KINDS][kMaxLevels];
};

}  // namespace

std::unique_ptr<EngineBase> make_engine(SolverState* state) {
    if (state->params.dtype == GMG_DTYPE_F32) return std::unique_ptr<EngineBase>(new Engine<float>(state));
    return std::unique_ptr<EngineBase>(new Engine<double>(state));
}

}  // namespace gmg
