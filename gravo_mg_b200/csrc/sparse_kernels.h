// Host-callable launchers for the sparse kernels (definitions in sparse_kernels.cu).
#pragma once
#include "sparse_kernels.cuh"

namespace gmg {

// How one CSR matrix is walked by the staged kernel; built once per sparsity pattern.
struct SpmvPlan {
    int path = 1;            // 0 staged (TMA), 1 direct
    int lanes = 8;           // direct: threads per row (1, 2, 4, 8, 16, 32)
    int staged_lanes = 1;    // staged: threads per row (1, 2, 4, 8); 1 sums in CSR order
    int row_begin = 0, row_end = -1;  // rows this plan covers (a rank's range; -1 = all)
    int n_tiles = 0;         // staged
    int n_early = 0;         // staged, multi-GPU: leading tiles that hold the pushed / halo-gathering rows
    int stage_elems = 0;     // staged: entries per stage (multiple of 4)
    int stage_rows = 0;      // staged: rows per tile = kStagedThreads / staged_lanes
    const int4* tile_desc = nullptr;  // device array n_tiles: {first row, end row, first entry & ~3, (end entry + 3) & ~3}
};

constexpr int kMaxRhsTile = 4;        // kernels are instantiated for K = 1..4 columns per pass
constexpr int kMaxNormBlocks = 148 * 16;
constexpr int kMaxNormChunks = 8;     // => at most 32 right-hand sides per solve
constexpr size_t kNormChunkStride = (size_t)kMaxNormBlocks * 2 * kMaxRhsTile;

// NORM partial sums of a multi-pass (K > 4) residual: one (grid x 2*kt) block per pass.
struct NormChunks {
    int n_chunks = 0;
    int kt[kMaxNormChunks] = {0};
    int n_blocks[kMaxNormChunks] = {0};
};

size_t staged_smem_bytes(int stage_rows, int stage_elems, size_t value_size);
size_t staged_smem_limit();  // usable dynamic shared memory per CTA on this device

// Launch acc = A x with epilogue `epi` over K (1..4) columns. Returns the grid size used
// (NORM callers need it to finalize the partial sums).
template <typename T>
int launch_spmv(int epi, int K, SpmvArgs<T> args, const SpmvPlan& plan, cudaStream_t stream);

// While on, launch_spmv does everything except launch (occupancy query, opt-in to large shared
// memory): lets the one-time attribute calls happen outside of stream capture.
void set_launch_dry_run(bool on);
// Programmatic dependent launch of the row-product kernels (on by default; thread-local).
void set_launch_pdl(bool on);

// Reduce NORM partials and update the loop state: residue, history, iter, done.
// `cond_handle` != 0 additionally drives a CUDA graph while-node (cudaGraphSetConditional).
void launch_norm_finalize(const double* partials, const NormChunks& chunks, CycleControl* ctl, double* hist_res,
                          double* hist_ms, int record, unsigned long long cond_handle, cudaStream_t stream);
// Multi-GPU stopping test: local NORM partials -> sums[2K] (then all-reduced) -> loop state.
void launch_norm_partial_sums(const double* partials, const NormChunks& chunks, double* sums, cudaStream_t stream);
void launch_norm_finalize_sums(const double* sums, int K, CycleControl* ctl, double* hist_res, double* hist_ms,
                               cudaStream_t stream);
// Halo exchange staging: buf[i, :] = v[idx[i], :] and back (K columns, row-major).
template <typename T>
void launch_pack(const T* v, const int* idx, int n, int K, T* buf, cudaStream_t stream);
template <typename T>
void launch_unpack(T* v, const int* idx, int n, int K, const T* buf, cudaStream_t stream);
void launch_cycle_begin(CycleControl* ctl, int max_iter, int criterion, double tol, int n_cols, cudaStream_t stream,
                        unsigned long long* trace = nullptr, int trace_cap = 0);

constexpr int kMaxSweeps = 16;        // pre / post sweeps per level the weight table holds

// dinv = 1 / diag(A) and *rho = max(*rho, Gershgorin bound of D^-1 A) (zero *rho first).
// vals_diff != nullptr: also the value array of the cancellation-free row product (SpmvArgs::diff):
// a copy of vals with the row sum (error-free TwoSum accumulation) in place of the diagonal entry.
template <typename T>
void launch_extract_dinv(int n, const int* rowptr, const int* colidx, const double* vals, T* dinv, double* rho,
                         CycleControl* ctl, cudaStream_t stream, double* vals_diff = nullptr, int row_begin = 0);  // rows [row_begin, n)
// Jacobi dampings per level and sweep from rho[level]: weights[(level * 2 + post) * kMaxSweeps + sweep].
template <typename T>
void launch_smoother_weights(const double* rho, int n_levels, int pre, int post, int smoother, double omega, double alpha,
                             T* weights, double* weights64, cudaStream_t stream);
void launch_cast_f64_f32(const double* src, float* dst, size_t n, cudaStream_t stream);
void launch_cast_f32_f64(const float* src, double* dst, size_t n, cudaStream_t stream);
void launch_add_f32_to_f64(const float* e, double* x, size_t n, cudaStream_t stream);  // x += e
void launch_expand_rows(int n_rows, const int* rowptr, int* rowidx, cudaStream_t stream, int row_begin = 0);  // rows [row_begin, n_rows)
// C = A * B on an existing sorted pattern of C, one thread per stored entry of C, products
// added in the order of A's row (the order a row-wise CPU Gustavson pass uses).
void launch_spgemm_numeric(int64_t nnz_c, const int* c_rowidx, const int* c_col, double* c_val, const int* a_ptr,
                           const int* a_col, const double* a_val, const int* b_ptr, const int* b_col,
                           const double* b_val, cudaStream_t stream);
// The same product with the search done once per sparsity pattern: build_spgemm_plan records, for
// every stored entry of C, the (A value, B value) index pairs whose products it sums (in the order
// of A's row); launch_spgemm_planned is then a plain gather-multiply-add, bit-identical to
// launch_spgemm_numeric. Returns the number of pairs.
long long build_spgemm_plan(int64_t nnz_c, const int* c_rowidx, const int* c_col, const int* a_ptr, const int* a_col,
                            const int* b_ptr, const int* b_col, DeviceBuffer<long long>& offsets, DeviceBuffer<int2>& pairs,
                            cudaStream_t stream);
void launch_spgemm_planned(int64_t nnz_c, const long long* offsets, const int2* pairs, const double* a_val,
                           const double* b_val, double* c_val, cudaStream_t stream);
void launch_csr_to_dense(int n, const int* rowptr, const int* colidx, const double* vals, double* dense, int lda,
                         cudaStream_t stream);

}  // namespace gmg
