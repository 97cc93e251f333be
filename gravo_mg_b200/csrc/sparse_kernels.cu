#include "sparse_kernels.h"

#include <cub/device/device_scan.cuh>
#include <map>
#include <mutex>
#include <tuple>

namespace gmg {

// ------------------------------------------------------------------ small device kernels
__global__ void cycle_begin_kernel(CycleControl* ctl, int max_iter, int criterion, double tol, int n_cols,
                                   unsigned long long* trace, int trace_cap) {
    ctl->trace = trace;
    ctl->trace_n = 0;
    ctl->trace_cap = trace_cap;
    ctl->iter = 0;
    ctl->done = 0;
    ctl->max_iter = max_iter;
    ctl->criterion = criterion;
    ctl->tol = tol;
    ctl->residue = 1.7976931348623157e308;
    ctl->t_start_ns = global_timer_ns();
    ctl->n_cols = n_cols;
    // ctl->error is sticky across the phases of one solve; the host clears it when staging
}

__global__ void norm_finalize_kernel(const double* __restrict__ partials, NormChunks chunks, CycleControl* ctl,
                                     double* hist_res, double* hist_ms, int record, unsigned long long cond_handle) {
    if (record && ctl->done) {
        if (cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, 0);
        return;
    }
    if (threadIdx.x == 0) trace_mark(ctl, 100);
    __shared__ double sums[2 * kMaxRhsTile * kMaxNormChunks];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n_sums = 0;
    for (int c = 0; c < chunks.n_chunks; ++c) {
        const int nv = 2 * chunks.kt[c];
        const double* part = partials + (size_t)c * kNormChunkStride;
        for (int j = warp; j < nv; j += blockDim.x / 32) {
            double s = 0.0;
            for (int blk = lane; blk < chunks.n_blocks[c]; blk += 32) s += part[(size_t)blk * nv + j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sums[n_sums + j] = s;
        }
        n_sums += nv;
    }
    __syncthreads();
    if (threadIdx.x == 0) apply_stopping_rule(sums, n_sums / 2, ctl, hist_res, hist_ms, record, cond_handle);
}

// dinv_i = 1 / A_ii, and rho = max_i sum_j |A_ij| / A_ii (Gershgorin bound of the spectral radius
// of D^-1 A, the upper end of the band the Chebyshev-weighted Jacobi sweeps damp).
__global__ void norm_partial_sums_kernel(const double* __restrict__ partials, NormChunks chunks, double* __restrict__ sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int n_sums = 0;
    for (int c = 0; c < chunks.n_chunks; ++c) {
        const int nv = 2 * chunks.kt[c];
        const double* part = partials + (size_t)c * kNormChunkStride;
        for (int j = warp; j < nv; j += blockDim.x / 32) {
            double s = 0.0;
            for (int blk = lane; blk < chunks.n_blocks[c]; blk += 32) s += part[(size_t)blk * nv + j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sums[n_sums + j] = s;
        }
        n_sums += nv;
    }
}

__global__ void norm_finalize_sums_kernel(const double* __restrict__ sums, int K, CycleControl* ctl, double* hist_res,
                                          double* hist_ms) {
    if (threadIdx.x != 0 || ctl->done) return;
    apply_stopping_rule(sums, K, ctl, hist_res, hist_ms, 1, 0);
}

template <typename T>
__global__ void pack_kernel(const T* __restrict__ v, const int* __restrict__ idx, int n, int K, T* __restrict__ buf) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * K) return;
    const int i = e / K, k = e - i * K;
    buf[e] = v[(size_t)idx[i] * K + k];
}
template <typename T>
__global__ void unpack_kernel(T* __restrict__ v, const int* __restrict__ idx, int n, int K, const T* __restrict__ buf) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * K) return;
    const int i = e / K, k = e - i * K;
    v[(size_t)idx[i] * K + k] = buf[e];
}

template <typename T>
__global__ void __launch_bounds__(256) extract_dinv_kernel(int row_begin, int n, const int* __restrict__ rowptr,
                                                          const int* __restrict__ colidx,
                                                          const double* __restrict__ vals, T* __restrict__ dinv,
                                                          double* __restrict__ rho, CycleControl* ctl,
                                                          double* __restrict__ vals_diff) {
    const int i = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    double bound = 0.0;
    if (i < n) {
        double d = 0.0, absum = 0.0;
        double s = 0.0, e = 0.0;  // row sum by an error-free TwoSum chain: s the running sum, e its rounding errors
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const double v = vals[p];
            if (colidx[p] == i) d += v;
            absum += fabs(v);
            const double t = s + v;
            const double bp = t - s;
            e += (s - (t - bp)) + (v - bp);
            s = t;
        }
        if (vals_diff) {
            // operator of the cancellation-free row product: off-diagonal entries as they are, the row
            // sum in place of the (first stored) diagonal entry
            const double rowsum = s + e;
            bool seen = false;
            for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
                const bool diag = colidx[p] == i;
                vals_diff[p] = diag ? (seen ? 0.0 : rowsum) : vals[p];
                seen = seen || diag;
            }
        }
        if (!(d > 0.0) || d > 1.7976931348623157e308) {
            atomicOr(&ctl->error, 1);
            dinv[i] = T(0);
        } else {
            const double inv = 1.0 / d;
            dinv[i] = (T)inv;
            bound = absum * inv;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bound = fmax(bound, __shfl_xor_sync(0xffffffffu, bound, o));
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = bound;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) bound = fmax(bound, sh[w]);
        // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(rho), (unsigned long long)__double_as_longlong(bound));
    }
}

// Per-level, per-sweep Jacobi dampings. Layout: weights[(level * 2 + post) * kMaxSweeps + sweep].
template <typename T>
__global__ void smoother_weights_kernel(const double* __restrict__ rho, int n_levels, int pre, int post, int smoother,
                                        double omega, double alpha, T* __restrict__ weights, double* __restrict__ weights64) {
    const int level = threadIdx.x;
    if (level >= n_levels) return;
    const double hi = rho[level], lo = hi / alpha;
    for (int side = 0; side < 2; ++side) {
        const int deg = side ? post : pre;
        for (int j = 0; j < deg; ++j) {
            double w = omega;
            if (smoother == 1) {
                const int jj = side ? deg - 1 - j : j;  // post-smoothing runs the roots in reverse
                const double root = 0.5 * (hi + lo) + 0.5 * (hi - lo) * cospi((2.0 * jj + 1.0) / (2.0 * deg));
                w = 1.0 / root;
            }
            weights[(level * 2 + side) * kMaxSweeps + j] = (T)w;
            weights64[(level * 2 + side) * kMaxSweeps + j] = w;
        }
    }
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ s, float* __restrict__ d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = (float)s[i];
}
__global__ void add_f32_to_f64_kernel(const float* __restrict__ e, double* __restrict__ x, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] += (double)e[i];
}
__global__ void cast_f32_f64_kernel(const float* __restrict__ s, double* __restrict__ d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = (double)s[i];
}

__global__ void expand_rows_kernel(int row_begin, int n_rows, const int* __restrict__ rowptr, int* __restrict__ rowidx) {
    const int r = row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) rowidx[p] = r;
}

__global__ void spgemm_numeric_kernel(int64_t nnz_c, const int* __restrict__ c_rowidx, const int* __restrict__ c_col,
                                      double* __restrict__ c_val, const int* __restrict__ a_ptr,
                                      const int* __restrict__ a_col, const double* __restrict__ a_val,
                                      const int* __restrict__ b_ptr, const int* __restrict__ b_col,
                                      const double* __restrict__ b_val) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz_c) return;
    const int i = c_rowidx[e];
    const int j = c_col[e];
    double s = 0.0;
    for (int p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
        const int k = a_col[p];
        int lo = b_ptr[k];
        const int end = b_ptr[k + 1];
        int hi = end;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (b_col[mid] < j)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < end && b_col[lo] == j) s += a_val[p] * b_val[lo];
    }
    c_val[e] = s;
}

// Product plan of C = A * B on the fixed pattern of C: for every stored entry e = (i, j) of C the
// list of (index into A's values, index into B's values) whose products make it up, in the order
// of A's row i. Built once per sparsity pattern with the same search the per-solve kernel used to
// repeat; the per-solve numeric product is then a gather-multiply-add without any search.
// MODE 0: count the pairs of every entry; MODE 1: write them at offsets[e].
template <int MODE>
__global__ void spgemm_pairs_kernel(int64_t nnz_c, const int* __restrict__ c_rowidx, const int* __restrict__ c_col,
                                    const int* __restrict__ a_ptr, const int* __restrict__ a_col,
                                    const int* __restrict__ b_ptr, const int* __restrict__ b_col,
                                    long long* __restrict__ offsets, int2* __restrict__ pairs) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz_c) return;
    const int i = c_rowidx[e];
    const int j = c_col[e];
    long long at = MODE ? offsets[e] : 0;
    for (int p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
        const int k = a_col[p];
        int lo = b_ptr[k];
        const int end = b_ptr[k + 1];
        int hi = end;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (b_col[mid] < j)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < end && b_col[lo] == j) {
            if (MODE) pairs[at] = make_int2(p, lo);
            ++at;
        }
    }
    if (!MODE) offsets[e] = at;
}

__global__ void spgemm_planned_kernel(int64_t nnz_c, const long long* __restrict__ offsets, const int2* __restrict__ pairs,
                                      const double* __restrict__ a_val, const double* __restrict__ b_val,
                                      double* __restrict__ c_val) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz_c) return;
    const long long t0 = offsets[e], t1 = offsets[e + 1];
    double s = 0.0;
    for (long long t = t0; t < t1; ++t) {
        const int2 pr = __ldg(pairs + t);
        s += a_val[pr.x] * b_val[pr.y];  // same order as the searching kernel: bit-identical sums
    }
    c_val[e] = s;
}

__global__ void csr_to_dense_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                    const double* __restrict__ vals, double* __restrict__ dense, int lda) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) dense[(size_t)colidx[p] * lda + r] = vals[p];  // column-major
}

// ------------------------------------------------------------------ launchers
size_t staged_smem_bytes(int stage_rows, int stage_elems, size_t value_size) {
    return 128 + (size_t)kStagedStages * staged_stage_bytes(stage_rows, stage_elems, value_size);
}

size_t staged_smem_limit() {
    static size_t limit = 0;
    if (!limit) {
        int dev = 0, v = 0;
        GMG_CUDA(cudaGetDevice(&dev));
        GMG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        limit = (size_t)v;
    }
    return limit;
}

namespace {

thread_local bool g_dry_run = false;
thread_local bool g_use_pdl = true;

// Launch with the programmatic-stream-serialization attribute: the kernel may begin (up to its
// grid_dependency_wait()) while the previous kernel in the stream is still draining. Captured
// into CUDA graphs as programmatic dependency edges.
template <typename Kernel, typename Args>
void launch_pdl(Kernel kernel, int grid, int block, size_t smem, cudaStream_t stream, const Args& args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    GMG_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        GMG_CUDA(cudaGetDevice(&dev));
        GMG_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

// CTAs per SM for (kernel, smem): queried once, also opts the kernel into large dynamic smem.
int resident_blocks(const void* kernel, int threads, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<const void*, size_t>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, smem);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    static std::map<const void*, bool> opted_in;
    if (smem > 48 * 1024 && !opted_in[kernel]) {
        // opt in once per kernel to everything the device allows beyond its static shared memory
        cudaFuncAttributes attr;
        GMG_CUDA(cudaFuncGetAttributes(&attr, kernel));
        GMG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(staged_smem_limit() - attr.sharedSizeBytes)));
        opted_in[kernel] = true;
    }
    int nb = 0;
    GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem));
    if (nb < 1) nb = 1;
    cache[key] = nb;
    return nb;
}

template <typename T, int K, int EPI, int LANES>
int launch_staged(SpmvArgs<T>& a, const SpmvPlan& plan, cudaStream_t stream) {
    const size_t smem = staged_smem_bytes(plan.stage_rows, plan.stage_elems, sizeof(T));
    auto kernel = spmv_staged_kernel<T, K, EPI, LANES>;
    const int per_sm = resident_blocks((const void*)kernel, kStagedThreads + 32, smem);
    int grid = std::min(std::max(plan.n_tiles, 1), per_sm * num_sms());
    if (EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) grid = std::min(grid, kMaxNormBlocks);
    if (!g_dry_run) launch_pdl(kernel, grid, kStagedThreads + 32, smem, stream, a);
    return grid;
}

template <typename T, int K, int EPI>
int launch_one(SpmvArgs<T>& a, const SpmvPlan& plan, cudaStream_t stream) {
    int grid = 0;
    if (plan.path == 0) {
        a.tile_desc = plan.tile_desc;
        a.n_tiles = plan.n_tiles;
        a.n_early = plan.n_early;
        a.stage_elems = plan.stage_elems;
        a.stage_rows = plan.stage_rows;
        switch (plan.staged_lanes) {
            case 1: grid = launch_staged<T, K, EPI, 1>(a, plan, stream); break;
            case 2: grid = launch_staged<T, K, EPI, 2>(a, plan, stream); break;
            case 4: grid = launch_staged<T, K, EPI, 4>(a, plan, stream); break;
            default: grid = launch_staged<T, K, EPI, 8>(a, plan, stream); break;
        }
    } else {
        if (plan.row_end >= 0) a.row_begin = plan.row_begin, a.n_rows = plan.row_end;
        const int rows_per_block = kDirectThreads / plan.lanes;
        const int64_t want = ((int64_t)(a.n_rows - a.row_begin) + rows_per_block - 1) / rows_per_block;
        grid = (int)std::min<int64_t>(std::max<int64_t>(want, 1), (int64_t)num_sms() * 32);
        if (EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) grid = std::min(grid, kMaxNormBlocks);
        if (g_dry_run) return grid;
        switch (plan.lanes) {
            case 1: launch_pdl(spmv_direct_kernel<T, K, EPI, 1>, grid, kDirectThreads, 0, stream, a); break;
            case 2: launch_pdl(spmv_direct_kernel<T, K, EPI, 2>, grid, kDirectThreads, 0, stream, a); break;
            case 4: launch_pdl(spmv_direct_kernel<T, K, EPI, 4>, grid, kDirectThreads, 0, stream, a); break;
            case 8: launch_pdl(spmv_direct_kernel<T, K, EPI, 8>, grid, kDirectThreads, 0, stream, a); break;
            case 16: launch_pdl(spmv_direct_kernel<T, K, EPI, 16>, grid, kDirectThreads, 0, stream, a); break;
            default: launch_pdl(spmv_direct_kernel<T, K, EPI, 32>, grid, kDirectThreads, 0, stream, a); break;
        }
    }
    GMG_CUDA(cudaGetLastError());
    return grid;
}

template <typename T, int K>
int launch_k(int epi, SpmvArgs<T>& a, const SpmvPlan& plan, cudaStream_t stream) {
    switch (epi) {
        case EPI_SPMV: return launch_one<T, K, EPI_SPMV>(a, plan, stream);
        case EPI_JACOBI: return launch_one<T, K, EPI_JACOBI>(a, plan, stream);
        case EPI_RESIDUAL: return launch_one<T, K, EPI_RESIDUAL>(a, plan, stream);
        case EPI_ADD: return launch_one<T, K, EPI_ADD>(a, plan, stream);
        case EPI_NORM: return launch_one<T, K, EPI_NORM>(a, plan, stream);
        case EPI_NORMJAC: return launch_one<T, K, EPI_NORMJAC>(a, plan, stream);
        case EPI_RESNORM: return launch_one<T, K, EPI_RESNORM>(a, plan, stream);
    }
    throw std::invalid_argument("unknown epilogue");
}

}  // namespace

void set_launch_dry_run(bool on) { g_dry_run = on; }
void set_launch_pdl(bool on) { g_use_pdl = on; }

template <typename T>
int launch_spmv(int epi, int K, SpmvArgs<T> args, const SpmvPlan& plan, cudaStream_t stream) {
    switch (K) {
        case 1: return launch_k<T, 1>(epi, args, plan, stream);
        case 2: return launch_k<T, 2>(epi, args, plan, stream);
        case 3: return launch_k<T, 3>(epi, args, plan, stream);
        case 4: return launch_k<T, 4>(epi, args, plan, stream);
    }
    throw std::invalid_argument("launch_spmv: K must be 1..4");
}
template int launch_spmv<double>(int, int, SpmvArgs<double>, const SpmvPlan&, cudaStream_t);
template int launch_spmv<float>(int, int, SpmvArgs<float>, const SpmvPlan&, cudaStream_t);

void launch_norm_finalize(const double* partials, const NormChunks& chunks, CycleControl* ctl, double* hist_res,
                          double* hist_ms, int record, unsigned long long cond_handle, cudaStream_t stream) {
    if (g_dry_run) return;
    norm_finalize_kernel<<<1, 256, 0, stream>>>(partials, chunks, ctl, hist_res, hist_ms, record, cond_handle);
    GMG_CUDA(cudaGetLastError());
}

void launch_norm_partial_sums(const double* partials, const NormChunks& chunks, double* sums, cudaStream_t stream) {
    norm_partial_sums_kernel<<<1, 256, 0, stream>>>(partials, chunks, sums);
    GMG_CUDA(cudaGetLastError());
}
void launch_norm_finalize_sums(const double* sums, int K, CycleControl* ctl, double* hist_res, double* hist_ms,
                               cudaStream_t stream) {
    norm_finalize_sums_kernel<<<1, 32, 0, stream>>>(sums, K, ctl, hist_res, hist_ms);
    GMG_CUDA(cudaGetLastError());
}
template <typename T>
void launch_pack(const T* v, const int* idx, int n, int K, T* buf, cudaStream_t stream) {
    if (n <= 0) return;
    pack_kernel<T><<<(n * K + 255) / 256, 256, 0, stream>>>(v, idx, n, K, buf);
    GMG_CUDA(cudaGetLastError());
}
template <typename T>
void launch_unpack(T* v, const int* idx, int n, int K, const T* buf, cudaStream_t stream) {
    if (n <= 0) return;
    unpack_kernel<T><<<(n * K + 255) / 256, 256, 0, stream>>>(v, idx, n, K, buf);
    GMG_CUDA(cudaGetLastError());
}
template void launch_pack<double>(const double*, const int*, int, int, double*, cudaStream_t);
template void launch_pack<float>(const float*, const int*, int, int, float*, cudaStream_t);
template void launch_unpack<double>(double*, const int*, int, int, const double*, cudaStream_t);
template void launch_unpack<float>(float*, const int*, int, int, const float*, cudaStream_t);

void launch_cycle_begin(CycleControl* ctl, int max_iter, int criterion, double tol, int n_cols, cudaStream_t stream,
                        unsigned long long* trace, int trace_cap) {
    cycle_begin_kernel<<<1, 1, 0, stream>>>(ctl, max_iter, criterion, tol, n_cols, trace, trace_cap);
    GMG_CUDA(cudaGetLastError());
}

template <typename T>
void launch_extract_dinv(int n, const int* rowptr, const int* colidx, const double* vals, T* dinv, double* rho,
                         CycleControl* ctl, cudaStream_t stream, double* vals_diff, int row_begin) {
    if (n <= row_begin) return;
    extract_dinv_kernel<T><<<(n - row_begin + 255) / 256, 256, 0, stream>>>(row_begin, n, rowptr, colidx, vals, dinv, rho, ctl, vals_diff);
    GMG_CUDA(cudaGetLastError());
}
template void launch_extract_dinv<double>(int, const int*, const int*, const double*, double*, double*, CycleControl*, cudaStream_t, double*, int);
template void launch_extract_dinv<float>(int, const int*, const int*, const double*, float*, double*, CycleControl*, cudaStream_t, double*, int);

template <typename T>
void launch_smoother_weights(const double* rho, int n_levels, int pre, int post, int smoother, double omega, double alpha,
                             T* weights, double* weights64, cudaStream_t stream) {
    if (n_levels <= 0) return;
    smoother_weights_kernel<T><<<1, 32, 0, stream>>>(rho, n_levels, pre, post, smoother, omega, alpha, weights, weights64);
    GMG_CUDA(cudaGetLastError());
}
template void launch_smoother_weights<double>(const double*, int, int, int, int, double, double, double*, double*, cudaStream_t);
template void launch_smoother_weights<float>(const double*, int, int, int, int, double, double, float*, double*, cudaStream_t);

static int stream_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)148 * 16); }

void launch_cast_f64_f32(const double* src, float* dst, size_t n, cudaStream_t stream) {
    if (!n) return;
    cast_f64_f32_kernel<<<stream_grid(n), 256, 0, stream>>>(src, dst, n);
    GMG_CUDA(cudaGetLastError());
}
void launch_cast_f32_f64(const float* src, double* dst, size_t n, cudaStream_t stream) {
    if (!n) return;
    cast_f32_f64_kernel<<<stream_grid(n), 256, 0, stream>>>(src, dst, n);
    GMG_CUDA(cudaGetLastError());
}
void launch_add_f32_to_f64(const float* e, double* x, size_t n, cudaStream_t stream) {
    if (!n) return;
    add_f32_to_f64_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, stream>>>(e, x, n);
    GMG_CUDA(cudaGetLastError());
}
void launch_expand_rows(int n_rows, const int* rowptr, int* rowidx, cudaStream_t stream, int row_begin) {
    if (n_rows <= row_begin) return;
    expand_rows_kernel<<<(n_rows - row_begin + 255) / 256, 256, 0, stream>>>(row_begin, n_rows, rowptr, rowidx);
    GMG_CUDA(cudaGetLastError());
}
void launch_spgemm_numeric(int64_t nnz_c, const int* c_rowidx, const int* c_col, double* c_val, const int* a_ptr,
                           const int* a_col, const double* a_val, const int* b_ptr, const int* b_col,
                           const double* b_val, cudaStream_t stream) {
    if (nnz_c <= 0) return;
    const int64_t blocks = (nnz_c + 255) / 256;
    spgemm_numeric_kernel<<<(unsigned)blocks, 256, 0, stream>>>(nnz_c, c_rowidx, c_col, c_val, a_ptr, a_col, a_val, b_ptr,
                                                                b_col, b_val);
    GMG_CUDA(cudaGetLastError());
}
long long build_spgemm_plan(int64_t nnz_c, const int* c_rowidx, const int* c_col, const int* a_ptr, const int* a_col,
                            const int* b_ptr, const int* b_col, DeviceBuffer<long long>& offsets, DeviceBuffer<int2>& pairs,
                            cudaStream_t stream) {
    offsets.ensure((size_t)nnz_c + 1);
    GMG_CUDA(cudaMemsetAsync(offsets.ptr, 0, ((size_t)nnz_c + 1) * sizeof(long long), stream));
    if (nnz_c <= 0) return 0;
    const unsigned blocks = (unsigned)((nnz_c + 255) / 256);
    spgemm_pairs_kernel<0><<<blocks, 256, 0, stream>>>(nnz_c, c_rowidx, c_col, a_ptr, a_col, b_ptr, b_col, offsets.ptr, nullptr);
    GMG_CUDA(cudaGetLastError());
    size_t tmp_bytes = 0;
    GMG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, offsets.ptr, offsets.ptr, (int64_t)nnz_c + 1, stream));
    DeviceBuffer<unsigned char> tmp;
    tmp.ensure(tmp_bytes);
    GMG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, tmp_bytes, offsets.ptr, offsets.ptr, (int64_t)nnz_c + 1, stream));
    long long total = 0;
    GMG_CUDA(cudaMemcpyAsync(&total, offsets.ptr + nnz_c, sizeof total, cudaMemcpyDeviceToHost, stream));
    GMG_CUDA(cudaStreamSynchronize(stream));
    pairs.ensure((size_t)std::max<long long>(total, 1));
    spgemm_pairs_kernel<1><<<blocks, 256, 0, stream>>>(nnz_c, c_rowidx, c_col, a_ptr, a_col, b_ptr, b_col, offsets.ptr, pairs.ptr);
    GMG_CUDA(cudaGetLastError());
    GMG_CUDA(cudaStreamSynchronize(stream));  // `tmp` is a local
    return total;
}

void launch_spgemm_planned(int64_t nnz_c, const long long* offsets, const int2* pairs, const double* a_val,
                           const double* b_val, double* c_val, cudaStream_t stream) {
    if (nnz_c <= 0) return;
    spgemm_planned_kernel<<<(unsigned)((nnz_c + 255) / 256), 256, 0, stream>>>(nnz_c, offsets, pairs, a_val, b_val, c_val);
    GMG_CUDA(cudaGetLastError());
}

void launch_csr_to_dense(int n, const int* rowptr, const int* colidx, const double* vals, double* dense, int lda,
                         cudaStream_t stream) {
    if (n <= 0) return;
    csr_to_dense_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, rowptr, colidx, vals, dense, lda);
    GMG_CUDA(cudaGetLastError());
}

}  // namespace gmg
