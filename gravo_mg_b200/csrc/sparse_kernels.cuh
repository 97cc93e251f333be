// Sparse kernels of the V-cycle (sm_100a). All of them are memory-bound CSR row products
//     acc_i = sum_j A_ij * x_j            (K right-hand sides, row-major N x K vectors)
// followed by a per-row epilogue, so there is ONE kernel template with five epilogues:
//
//   EPI_SPMV      out_i = acc_i  [, out2_i = omega * dinv_i * acc_i]   restriction r_c = U^T r
//                                                  (reference multigrid_solver.cpp:1069)
//   EPI_JACOBI    out_i = x_i + omega * dinv_i * (b_i - acc_i)          smoother sweep
//                                                  (replaces the GS sweep of :1194-1226)
//   EPI_RESIDUAL  out_i = b_i - acc_i                                   (:1066)
//   EPI_ADD       out_i = xin_i + acc_i                                 x += U e (:1082)
//   EPI_NORM      partial sums of w_i (acc_i - b_i)^2 and w_i b_i^2     residualCheck (:1228-1277)
//
// Two data paths:
//   staged  persistent CTAs walk row tiles; the tile's contiguous (colidx, vals) slab is
//           brought into shared memory by the TMA engine (cp.async.bulk + mbarrier, a
//           STAGES-deep ring), then LANES threads own one row and read its entries from
//           shared memory (consecutive lanes, consecutive entries: conflict free), gather x
//           through L1/L2 and combine with a shuffle butterfly. With LANES = 1 the products
//           are summed in CSR order, bit-identical to a sequential CPU loop (the library is
//           compiled with -fmad=false for that reason); LANES > 1 trades that for more rows
//           in flight per byte of shared memory (short tiles, many resident CTAs).
//   direct  LANES threads per row straight from global memory with a shuffle reduction;
//           used when a row does not fit a stage and as an independent cross-check.
#pragma once
#include "common.cuh"

namespace gmg {

enum Epilogue { EPI_SPMV = 0, EPI_JACOBI = 1, EPI_RESIDUAL = 2, EPI_ADD = 3, EPI_NORM = 4 };

// Device-resident loop state of one solve (multigrid_solver.cpp:1411-1417).
struct CycleControl {
    int iter;       // V-cycles completed
    int done;       // 1: every cycle kernel returns immediately
    int max_iter;
    int criterion;  // stoppingCriteria 0..3
    double tol;
    double residue;
    unsigned long long t_start_ns;
    int error;      // sticky: 1 bad diagonal, 2 non-finite residual, 4 coarse factor breakdown
    int n_cols;
};

template <typename T>
struct SpmvArgs {
    int n_rows = 0;
    int ld = 1;                        // leading dimension of every vector (= total K of the solve)
    const int* rowptr = nullptr;
    const int* colidx = nullptr;
    const T* vals = nullptr;
    const T* x = nullptr;              // gathered vector
    const T* b = nullptr;              // JACOBI / RESIDUAL / NORM
    const T* xin = nullptr;            // ADD: own-row input
    const T* dinv = nullptr;           // JACOBI; SPMV when out2 != nullptr
    const double* weight = nullptr;    // NORM: per-row weight (nullptr = 1)
    T* out = nullptr;
    T* out2 = nullptr;
    T omega = T(0);
    const T* omega_ptr = nullptr;      // device-resident damping of this sweep (overrides omega)
    double* partials = nullptr;        // NORM: [gridDim.x][2*K]
    const int* tile_rows = nullptr;    // staged: n_tiles + 1 row offsets
    int n_tiles = 0;
    int stage_elems = 0;               // staged: capacity of one stage in entries (multiple of 4)
    const CycleControl* ctl = nullptr; // optional early-out
};

constexpr int kStagedThreads = 256;
constexpr int kStagedStages = 3;
constexpr int kDirectThreads = 256;

template <int NV, int TPB>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double* out) {
    __shared__ double sh[NV][TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double s = v[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh[j][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < TPB / 32; ++w) s += sh[threadIdx.x][w];
        out[threadIdx.x] = s;
    }
}

// Own-row operands of the epilogue. They do not depend on the row product, so the staged kernel
// loads them before it waits for the tile's slab (their latency hides behind the bulk copy).
template <typename T, int K>
struct RowOperands {
    T b[K];
    T xo[K];
    T scale;     // omega * dinv
    double w;    // NORM weight
};

template <typename T, int K, int EPI>
__device__ __forceinline__ void load_row_operands(const SpmvArgs<T>& a, int row, RowOperands<T, K>& r) {
    const size_t o = (size_t)row * a.ld;
    if (EPI == EPI_JACOBI || (EPI == EPI_SPMV && a.out2)) {
        const T om = a.omega_ptr ? *a.omega_ptr : a.omega;
        r.scale = om * a.dinv[row];
    }
    if (EPI == EPI_JACOBI || EPI == EPI_RESIDUAL || EPI == EPI_NORM) {
#pragma unroll
        for (int k = 0; k < K; ++k) r.b[k] = a.b[o + k];
    }
    if (EPI == EPI_JACOBI) {
#pragma unroll
        for (int k = 0; k < K; ++k) r.xo[k] = a.x[o + k];
    }
    if (EPI == EPI_ADD) {
#pragma unroll
        for (int k = 0; k < K; ++k) r.xo[k] = a.xin[o + k];
    }
    if (EPI == EPI_NORM) r.w = a.weight ? a.weight[row] : 1.0;
}

template <typename T, int K, int EPI>
__device__ __forceinline__ void row_epilogue(const SpmvArgs<T>& a, int row, const T (&acc)[K], const RowOperands<T, K>& r,
                                             double (&nrm)[2 * K]) {
    const size_t o = (size_t)row * a.ld;
    if (EPI == EPI_SPMV) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = acc[k];
        if (a.out2) {
#pragma unroll
            for (int k = 0; k < K; ++k) a.out2[o + k] = r.scale * acc[k];
        }
    } else if (EPI == EPI_JACOBI) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = r.xo[k] + r.scale * (r.b[k] - acc[k]);
    } else if (EPI == EPI_RESIDUAL) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = r.b[k] - acc[k];
    } else if (EPI == EPI_ADD) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = r.xo[k] + acc[k];
    } else {  // EPI_NORM
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double bk = (double)r.b[k];
            const double d = (double)acc[k] - bk;
            nrm[2 * k] += r.w * d * d;
            nrm[2 * k + 1] += r.w * bk * bk;
        }
    }
}

// ---------------------------------------------------------------------------- staged path
template <typename T, int K, int EPI, int LANES>
__global__ void __launch_bounds__(kStagedThreads) spmv_staged_kernel(const SpmvArgs<T> a) {
    constexpr int TPB = kStagedThreads;
    constexpr int STAGES = kStagedStages;
    if (a.ctl && a.ctl->done) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // first 128 bytes
    const size_t stage_bytes = (size_t)a.stage_elems * (sizeof(T) + sizeof(int));
    unsigned char* stage0 = smem_raw + 128;

    const int tid = threadIdx.x;
    const int lane = tid % LANES;
    const int n_my = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
    }
    __syncthreads();

    auto issue = [&](int it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int s = it % STAGES;
        const int p0 = a.rowptr[a.tile_rows[tile]] & ~3;
        const int p1 = (a.rowptr[a.tile_rows[tile + 1]] + 3) & ~3;
        const uint32_t cnt = (uint32_t)(p1 - p0);
        unsigned char* sv = stage0 + s * stage_bytes;
        unsigned char* sc = sv + (size_t)a.stage_elems * sizeof(T);
        mbar_arrive_expect_tx(&bars[s], cnt * (uint32_t)(sizeof(T) + sizeof(int)));
        if (cnt) {
            bulk_copy_g2s(sv, a.vals + p0, cnt * (uint32_t)sizeof(T), &bars[s]);
            bulk_copy_g2s(sc, a.colidx + p0, cnt * (uint32_t)sizeof(int), &bars[s]);
        }
    };
    if (tid == 0)
        for (int it = 0; it < STAGES && it < n_my; ++it) issue(it);

    double nrm[2 * K];
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) nrm[j] = 0.0;

    for (int it = 0; it < n_my; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int s = it % STAGES;
        const int r0 = a.tile_rows[tile];
        const int r1 = a.tile_rows[tile + 1];
        const int base = a.rowptr[r0] & ~3;
        const int row = r0 + tid / LANES;
        const bool active = row < r1;
        int ps = 0, pe = 0;
        RowOperands<T, K> ops;
        if (active) {
            ps = a.rowptr[row] - base;
            pe = a.rowptr[row + 1] - base;
            if (lane == 0) load_row_operands<T, K, EPI>(a, row, ops);
        }
        const T* sv = reinterpret_cast<const T*>(stage0 + s * stage_bytes);
        const int* sc = reinterpret_cast<const int*>(stage0 + s * stage_bytes + (size_t)a.stage_elems * sizeof(T));

        mbar_wait(&bars[s], (uint32_t)((it / STAGES) & 1));

        T acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = T(0);
#pragma unroll 4
        for (int p = ps + lane; p < pe; p += LANES) {
            const int c = sc[p];
            const T v = sv[p];
            const T* xp = a.x + (size_t)c * a.ld;
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += v * __ldg(xp + k);
        }
        if (LANES > 1) {
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
        }
        if (active && lane == 0) row_epilogue<T, K, EPI>(a, row, acc, ops, nrm);
        __syncthreads();  // everyone is done reading stage s before it is refilled
        if (tid == 0 && it + STAGES < n_my) issue(it + STAGES);
    }
    if (EPI == EPI_NORM) block_sum_store<2 * K, TPB>(nrm, a.partials + (size_t)blockIdx.x * 2 * K);
}

// ---------------------------------------------------------------------------- direct path
template <typename T, int K, int EPI, int LANES>
__global__ void __launch_bounds__(kDirectThreads) spmv_direct_kernel(const SpmvArgs<T> a) {
    constexpr int TPB = kDirectThreads;
    if (a.ctl && a.ctl->done) return;
    const int lane = threadIdx.x % LANES;
    const int rows_per_block = TPB / LANES;
    double nrm[2 * K];
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) nrm[j] = 0.0;

    // block-uniform trip count so the shuffles below always see full warps
    for (int first = blockIdx.x * rows_per_block; first < a.n_rows; first += gridDim.x * rows_per_block) {
        const int row = first + threadIdx.x / LANES;
        const bool active = row < a.n_rows;
        T acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = T(0);
        if (active) {
            const int pe = a.rowptr[row + 1];
            for (int p = a.rowptr[row] + lane; p < pe; p += LANES) {
                const int c = __ldg(a.colidx + p);
                const T v = __ldg(a.vals + p);
                const T* xp = a.x + (size_t)c * a.ld;
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] += v * __ldg(xp + k);
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        if (active && lane == 0) {
            RowOperands<T, K> ops;
            load_row_operands<T, K, EPI>(a, row, ops);
            row_epilogue<T, K, EPI>(a, row, acc, ops, nrm);
        }
    }
    if (EPI == EPI_NORM) block_sum_store<2 * K, TPB>(nrm, a.partials + (size_t)blockIdx.x * 2 * K);
}

}  // namespace gmg
