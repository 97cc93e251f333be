// Coarse tail of the V-cycle in ONE persistent kernel.
//
// Below a few ten-thousand rows every operator of the cycle (multigrid_solver.cpp:1059-1088) is
// a launch-latency-sized kernel: the levels hold < 2 % of the bytes of a cycle but the same ~8
// dependent operators each. This kernel runs all of them — sweeps, residual, restriction, the
// dense coarse solve, prolongation, post-sweeps of every level from `tail level` down and back
// up — as a table of operators separated by a grid-wide barrier instead of kernel boundaries.
// All its data is L2 resident, so rows are read straight from global memory (LANES threads per
// row, shuffle reduction); vectors written inside the kernel are re-read with ld.global.cg
// because another SM's L1 may hold a stale line (L1 is only invalidated at kernel boundaries).
#pragma once
#include "sparse_kernels.cuh"

namespace gmg {

enum TailKind { TAIL_ROWS = 0, TAIL_COLDOT = 1, TAIL_TO_F64 = 2, TAIL_FROM_F64 = 3, TAIL_ZERO = 4 };

constexpr int kTailThreads = 1024;

template <typename T>
struct TailOp {
    int kind = TAIL_ROWS;
    int epi = EPI_SPMV;     // TAIL_ROWS
    int lanes = 8;          // TAIL_ROWS: threads per row (1, 2, 4 or 8), chosen so one pass covers the level
    int kcols = 1;          // columns handled by this operator (1..4)
    int mat = -1;           // TAIL_ROWS inside the cluster kernel (cluster_tail.cuh): index of the staged operator
    SpmvArgs<T> a;          // TAIL_ROWS
    // TAIL_COLDOT: out[c, k] = sum over the stored triangle of column c of M of M[r, c] * v[r, k]
    const double* M = nullptr;
    int ldm = 0, n = 0, upper = 0;
    const double* v = nullptr;
    int v_ld = 1;
    double* out = nullptr;
    int out_ld = 1;
    // TAIL_TO_F64 / TAIL_FROM_F64 / TAIL_ZERO
    const void* src = nullptr;
    void* dst = nullptr;
    size_t count = 0;  // elements (casts) or bytes (zero)
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Self-resetting grid barrier: bar[0] counts arrivals, bar[1] is the generation. Every CTA of
// the grid must be resident (the launcher sizes the grid to one CTA per SM).
__device__ __forceinline__ void grid_barrier(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = ld_acquire_gpu(&bar[1]);
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(&bar[1], 1u);
        } else {
            while (ld_acquire_gpu(&bar[1]) == gen) {
            }
        }
        __threadfence();
    }
    __syncthreads();
}

template <typename U>
__device__ __forceinline__ U ld_cg(const U* p) {
    return __ldcg(p);
}

template <typename T, int K, int EPI>
__device__ __forceinline__ void tail_row_finish(const SpmvArgs<T>& a, int row, const T (&acc)[K]) {
    const size_t o = (size_t)row * a.ld;
    if (EPI == EPI_SPMV) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = acc[k];
        if (a.out2) {
            const T s = (a.omega_ptr ? *a.omega_ptr : a.omega) * a.dinv[row];
#pragma unroll
            for (int k = 0; k < K; ++k) a.out2[o + k] = s * acc[k];
        }
    } else if (EPI == EPI_JACOBI) {
        const T s = (a.omega_ptr ? *a.omega_ptr : a.omega) * a.dinv[row];
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = ld_cg(a.x + o + k) + s * (ld_cg(a.b + o + k) - acc[k]);
    } else if (EPI == EPI_RESIDUAL) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = ld_cg(a.b + o + k) - acc[k];
    } else {  // EPI_ADD
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = ld_cg(a.xin + o + k) + acc[k];
    }
}

template <typename T, int K, int LANES>
__device__ __forceinline__ void tail_rows(const TailOp<T>& op) {
    const SpmvArgs<T>& a = op.a;
    constexpr int ROWS_PER_WARP = 32 / LANES;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x % LANES;
    const int sub = (threadIdx.x & 31) / LANES;
    for (int base = warp * ROWS_PER_WARP; base < a.n_rows; base += n_warps * ROWS_PER_WARP) {
        const int row = base + sub;
        const bool active = row < a.n_rows;
        T acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = T(0);
        if (active) {
            const int ps = a.rowptr[row], pe = a.rowptr[row + 1];
#pragma unroll 4
            for (int p = ps + lane; p < pe; p += LANES) {
                const int c = a.colidx[p];
                const T v = a.vals[p];
                const T* xp = a.x + (size_t)c * a.ld;
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] += v * ld_cg(xp + k);
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        if (active && lane == 0) {
            switch (op.epi) {
                case EPI_SPMV: tail_row_finish<T, K, EPI_SPMV>(a, row, acc); break;
                case EPI_JACOBI: tail_row_finish<T, K, EPI_JACOBI>(a, row, acc); break;
                case EPI_RESIDUAL: tail_row_finish<T, K, EPI_RESIDUAL>(a, row, acc); break;
                default: tail_row_finish<T, K, EPI_ADD>(a, row, acc); break;
            }
        }
    }
}

template <typename T, int K>
__device__ __forceinline__ void tail_coldot(const TailOp<T>& op) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int c = warp; c < op.n; c += n_warps) {
        const int lo = op.upper ? 0 : c, hi = op.upper ? c + 1 : op.n;
        const double* col = op.M + (size_t)c * op.ldm;
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int r = lo + lane; r < hi; r += 32) {
            const double m = col[r];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fma(m, ld_cg(op.v + (size_t)r * op.v_ld + k), acc[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < K; ++k) op.out[(size_t)c * op.out_ld + k] = acc[k];
    }
}

template <typename T, int K>
__device__ __forceinline__ void tail_dispatch(const TailOp<T>& op) {
    switch (op.kind) {
        case TAIL_ROWS:
            switch (op.lanes) {
                case 1: tail_rows<T, K, 1>(op); break;
                case 2: tail_rows<T, K, 2>(op); break;
                case 4: tail_rows<T, K, 4>(op); break;
                default: tail_rows<T, K, 8>(op); break;
            }
            break;
        case TAIL_COLDOT: tail_coldot<T, K>(op); break;
        default: break;
    }
}

template <typename T>
__global__ void __launch_bounds__(kTailThreads, 1) tail_kernel(const TailOp<T>* __restrict__ ops, int n_ops, unsigned* bar) {
    const size_t gtid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
    for (int i = 0; i < n_ops; ++i) {
        const TailOp<T>& op = ops[i];
        if (op.kind == TAIL_TO_F64) {
            const T* s = static_cast<const T*>(op.src);
            double* d = static_cast<double*>(op.dst);
            for (size_t e = gtid; e < op.count; e += gthreads) d[e] = (double)ld_cg(s + e);
        } else if (op.kind == TAIL_FROM_F64) {
            const double* s = static_cast<const double*>(op.src);
            T* d = static_cast<T*>(op.dst);
            for (size_t e = gtid; e < op.count; e += gthreads) d[e] = (T)ld_cg(s + e);
        } else if (op.kind == TAIL_ZERO) {
            unsigned* d = static_cast<unsigned*>(op.dst);
            for (size_t e = gtid; e < op.count / 4; e += gthreads) d[e] = 0u;
        } else {
            switch (op.kcols) {
                case 1: tail_dispatch<T, 1>(op); break;
                case 2: tail_dispatch<T, 2>(op); break;
                case 3: tail_dispatch<T, 3>(op); break;
                default: tail_dispatch<T, 4>(op); break;
            }
        }
        if (i + 1 < n_ops) grid_barrier(bar);
    }
}

}  // namespace gmg
