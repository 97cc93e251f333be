// Coarse tail of the V-cycle as ONE thread-block cluster (sm_100a).
//
// Levels of a few thousand rows hold < 1 % of the bytes of a cycle, but every operator on them is a kernel
// whose duration is launch + dependency latency (~3.5 us each, PDL-chained; 9 of them below level 2 of the
// 1 M-vertex system: 37 us of a 284 us cycle). This kernel runs all operators of those levels — sweeps,
// residual, restriction, the dense coarse solve, prolongation, post-sweeps — inside one cluster of up to 16 CTAs:
//
//   * the operators are constant during a solve, so every CTA brings ITS row slab of every sparse operator of
//     the tail (row pointer slice, column indices, values) into shared memory once per launch with bulk
//     asynchronous copies (TMA engine, one mbarrier), issued BEFORE the programmatic-dependency wait: the
//     copies overlap the tail of the previous kernel of the cycle;
//   * an operator is then: LANES threads per row read the slab from shared memory, gather the vector through
//     L2 (ld.global.cg: the vectors are rewritten inside the kernel by other SMs), shuffle-reduce, store;
//   * operators are separated by the hardware cluster barrier (barrier.cluster arrive.release / wait.acquire,
//     ~0.3 us) instead of a kernel boundary or a grid barrier through L2 atomics (tail_kernel.cuh: ~6 us).
//
// The dense coarse solve x = W^T (W b) streams W from L2 with one warp per column, as in tail_kernel.cuh.
// Row products are summed by LANES threads with a shuffle butterfly: same arithmetic as the per-operator
// kernels with more than one lane (not the CSR-order sum of LANES = 1); parity bar 1e-13 relative per operator.
#pragma once
#include "tail_kernel.cuh"

namespace gmg {

constexpr int kClusterTailThreads = 1024;
constexpr int kClusterTailMaxMats = 8;

// One sparse operator of the tail as the kernel stages it (device table, built once per cycle list).
struct ClusterMat {
    const int* rowptr = nullptr;
    const int* colidx = nullptr;
    const void* vals = nullptr;
    int n_rows = 0;
};

struct ClusterTailArgs {
    const void* ops = nullptr;   // TailOp<T>[n_ops]; TAIL_ROWS ops carry the index of their operator in `mats` (TailOp::mat)
    int n_ops = 0;
    int n_mats = 0;
    ClusterMat mats[kClusterTailMaxMats];
    const CycleControl* ctl = nullptr;
};

__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_n_ctas() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// All threads of all CTAs of the cluster. The warps arrive from divergent code (row loops with different trip counts,
// single-thread trace marks), so the barrier must not be the .aligned form (synccheck: "divergent thread(s) in warp").
__device__ __forceinline__ void cluster_barrier() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}

// Where one CTA keeps its slab of one operator in shared memory.
struct ClusterSlab {
    int r0, r1;        // rows of this CTA
    int rp0;           // first row of the staged row-pointer slice (r0 & ~3)
    int e0;            // first staged entry (rowptr[r0] & ~3)
    const int* rp;     // shared: row pointer slice, rp[r - rp0]
    const void* vals;  // shared: values from entry e0
    const int* cols;   // shared: column indices from entry e0
};

// Own-row operands of the epilogue, requested BEFORE the gathers of the row so that everything a row needs is in
// flight at once (one L2 round trip per operator).
template <typename T, int K>
struct ClusterOwn {
    T b[K], x[K], scale;
};

template <typename T, int K, int LANES>
__device__ __forceinline__ void cluster_rows(const TailOp<T>& op, const ClusterSlab& sl) {
    const SpmvArgs<T>& a = op.a;
    constexpr int ROWS_PER_WARP = 32 / LANES;
    constexpr int U = 8;  // gathers in flight per thread
    const int warp = threadIdx.x >> 5, n_warps = kClusterTailThreads >> 5;
    const int lane = threadIdx.x % LANES;
    const int sub = (threadIdx.x & 31) / LANES;
    const T* sv = static_cast<const T*>(sl.vals);
    const int epi = op.epi;
    for (int base = sl.r0 + warp * ROWS_PER_WARP; base < sl.r1; base += n_warps * ROWS_PER_WARP) {
        const int row = base + sub;
        const bool active = row < sl.r1;
        T acc[K];
        ClusterOwn<T, K> own;
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = T(0), own.b[k] = T(0), own.x[k] = T(0);
        own.scale = T(0);
        if (active) {
            if (lane == 0) {
                const size_t o = (size_t)row * a.ld;
                if (epi == EPI_JACOBI || (epi == EPI_SPMV && a.out2)) own.scale = (a.omega_ptr ? *a.omega_ptr : a.omega) * a.dinv[row];
                if (epi == EPI_JACOBI || epi == EPI_RESIDUAL) {
#pragma unroll
                    for (int k = 0; k < K; ++k) own.b[k] = ld_cg(a.b + o + k);
                }
                if (epi == EPI_JACOBI) {
#pragma unroll
                    for (int k = 0; k < K; ++k) own.x[k] = ld_cg(a.x + o + k);
                }
                if (epi == EPI_ADD) {
#pragma unroll
                    for (int k = 0; k < K; ++k) own.x[k] = ld_cg(a.xin + o + k);
                }
            }
            const int ps = sl.rp[row - sl.rp0] - sl.e0, pe = sl.rp[row + 1 - sl.rp0] - sl.e0;
            for (int p0 = ps + lane; p0 < pe; p0 += LANES * U) {
                T v[U], g[U][K];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int p = p0 + u * LANES;
                    const bool in = p < pe;
                    v[u] = in ? sv[p] : T(0);
                    const T* xp = a.x + (size_t)(in ? sl.cols[p] : 0) * a.ld;
#pragma unroll
                    for (int k = 0; k < K; ++k) g[u][k] = in ? ld_cg(xp + k) : T(0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (p0 + u * LANES < pe) {
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] += v[u] * g[u][k];
                    }
                }
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        if (active && lane == 0) {
            const size_t o = (size_t)row * a.ld;
            if (epi == EPI_SPMV) {
#pragma unroll
                for (int k = 0; k < K; ++k) a.out[o + k] = acc[k];
                if (a.out2) {
#pragma unroll
                    for (int k = 0; k < K; ++k) a.out2[o + k] = own.scale * acc[k];
                }
            } else if (epi == EPI_JACOBI) {
#pragma unroll
                for (int k = 0; k < K; ++k) a.out[o + k] = own.x[k] + own.scale * (own.b[k] - acc[k]);
            } else if (epi == EPI_RESIDUAL) {
#pragma unroll
                for (int k = 0; k < K; ++k) a.out[o + k] = own.b[k] - acc[k];
            } else {  // EPI_ADD
#pragma unroll
                for (int k = 0; k < K; ++k) a.out[o + k] = own.x[k] + acc[k];
            }
        }
    }
}

template <typename T, int K>
__device__ __forceinline__ void cluster_rows_lanes(const TailOp<T>& op, const ClusterSlab& sl) {
    switch (op.lanes) {
        case 1: cluster_rows<T, K, 1>(op, sl); break;
        case 2: cluster_rows<T, K, 2>(op, sl); break;
        case 4: cluster_rows<T, K, 4>(op, sl); break;
        default: cluster_rows<T, K, 8>(op, sl); break;
    }
}

// out[c, k] = sum over the stored triangle of column c of M of M[r, c] * v[r, k]; columns over all warps of the cluster
template <typename T, int K>
__device__ __forceinline__ void cluster_coldot(const TailOp<T>& op, unsigned rank, unsigned n_ctas) {
    const int warp = (int)rank * (kClusterTailThreads >> 5) + (threadIdx.x >> 5);
    const int n_warps = (int)n_ctas * (kClusterTailThreads >> 5);
    const int lane = threadIdx.x & 31;
    for (int c = warp; c < op.n; c += n_warps) {
        const int lo = op.upper ? 0 : c, hi = op.upper ? c + 1 : op.n;
        const double* col = op.M + (size_t)c * op.ldm;
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        constexpr int U = 8;  // 32-row strips in flight per lane (the products are added in row order, as a plain loop would)
        for (int r0 = lo + lane; r0 < hi; r0 += 32 * U) {
            double m[U], v[U][K];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = r0 + 32 * u;
                const bool in = r < hi;
                m[u] = in ? col[r] : 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) v[u][k] = in ? ld_cg(op.v + (size_t)r * op.v_ld + k) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r0 + 32 * u < hi) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] = fma(m[u], v[u][k], acc[k]);
                }
            }
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

template <typename T>
__global__ void __launch_bounds__(kClusterTailThreads, 1) cluster_tail_kernel(const ClusterTailArgs args) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ ClusterSlab slabs[kClusterTailMaxMats];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const unsigned rank = cluster_cta_rank(), n_ctas = cluster_n_ctas();
    const TailOp<T>* ops = static_cast<const TailOp<T>*>(args.ops);

    // ---- stage this CTA's slab of every sparse operator (constant data: before the dependency wait); one lane of
    // warp 0 issues, the warp stays together
    if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
        size_t off = 128;
        uint32_t total = 0;
        for (int m = 0; m < args.n_mats; ++m) {
            const ClusterMat& cm = args.mats[m];
            ClusterSlab s;
            s.r0 = (int)((long long)cm.n_rows * rank / n_ctas);
            s.r1 = (int)((long long)cm.n_rows * (rank + 1) / n_ctas);
            s.rp0 = s.r0 & ~3;
            const int e_begin = cm.rowptr[s.r0], e_end = cm.rowptr[s.r1];
            s.e0 = e_begin & ~3;
            const uint32_t n_rp = (uint32_t)(((s.r1 + 1 + 3) & ~3) - s.rp0);
            const uint32_t cnt = (uint32_t)(((e_end + 3) & ~3) - s.e0);
            int* rp = reinterpret_cast<int*>(smem_raw + off);
            off += (size_t)n_rp * sizeof(int);
            unsigned char* sv = smem_raw + off;
            off += (size_t)cnt * sizeof(T);
            off = (off + 15) & ~(size_t)15;
            int* sc = reinterpret_cast<int*>(smem_raw + off);
            off += (size_t)cnt * sizeof(int);
            off = (off + 15) & ~(size_t)15;
            s.rp = rp, s.vals = sv, s.cols = sc;
            slabs[m] = s;
            total += n_rp * 4u + cnt * (uint32_t)(sizeof(T) + sizeof(int));
        }
        mbar_arrive_expect_tx(bar, total);
        for (int m = 0; m < args.n_mats; ++m) {
            const ClusterMat& cm = args.mats[m];
            const ClusterSlab& s = slabs[m];
            const int e_end = cm.rowptr[s.r1];
            const uint32_t n_rp = (uint32_t)(((s.r1 + 1 + 3) & ~3) - s.rp0);
            const uint32_t cnt = (uint32_t)(((e_end + 3) & ~3) - s.e0);
            bulk_copy_g2s(const_cast<int*>(s.rp), cm.rowptr + s.rp0, n_rp * 4u, bar);
            if (cnt) {
                bulk_copy_g2s(const_cast<void*>(s.vals), static_cast<const T*>(cm.vals) + s.e0, cnt * (uint32_t)sizeof(T), bar);
                bulk_copy_g2s(const_cast<int*>(s.cols), cm.colidx + s.e0, cnt * (uint32_t)sizeof(int), bar);
            }
        }
    }
    __syncwarp();
    }
    __syncthreads();           // slabs[] visible
    grid_dependency_wait();    // vectors of the previous kernel of the cycle
    grid_launch_dependents();
    if (rank == 0 && threadIdx.x == 0) trace_mark(args.ctl, 105);
    mbar_wait(bar, 0);
    if (rank == 0 && threadIdx.x == 0) trace_mark(args.ctl, 106);  // slabs in shared memory

    const size_t ctid = (size_t)rank * kClusterTailThreads + threadIdx.x, cthreads = (size_t)n_ctas * kClusterTailThreads;
    for (int i = 0; i < args.n_ops; ++i) {
        const TailOp<T>& op = ops[i];
        if (op.kind == TAIL_TO_F64) {
            const T* s = static_cast<const T*>(op.src);
            double* d = static_cast<double*>(op.dst);
            for (size_t e = ctid; e < op.count; e += cthreads) d[e] = (double)ld_cg(s + e);
        } else if (op.kind == TAIL_FROM_F64) {
            const double* s = static_cast<const double*>(op.src);
            T* d = static_cast<T*>(op.dst);
            for (size_t e = ctid; e < op.count; e += cthreads) d[e] = (T)ld_cg(s + e);
        } else if (op.kind == TAIL_ZERO) {
            unsigned* d = static_cast<unsigned*>(op.dst);
            for (size_t e = ctid; e < op.count / 4; e += cthreads) d[e] = 0u;
        } else if (op.kind == TAIL_ROWS) {
            const ClusterSlab& sl = slabs[op.mat];
            switch (op.kcols) {
                case 1: cluster_rows_lanes<T, 1>(op, sl); break;
                case 2: cluster_rows_lanes<T, 2>(op, sl); break;
                case 3: cluster_rows_lanes<T, 3>(op, sl); break;
                default: cluster_rows_lanes<T, 4>(op, sl); break;
            }
        } else {
            switch (op.kcols) {
                case 1: cluster_coldot<T, 1>(op, rank, n_ctas); break;
                case 2: cluster_coldot<T, 2>(op, rank, n_ctas); break;
                case 3: cluster_coldot<T, 3>(op, rank, n_ctas); break;
                default: cluster_coldot<T, 4>(op, rank, n_ctas); break;
            }
        }
        if (i + 1 < args.n_ops) cluster_barrier();
        if (rank == 0 && threadIdx.x == 0) trace_mark(args.ctl, 110 + (unsigned)op.kind * 10 + (op.kind == TAIL_ROWS ? (unsigned)op.epi : 0u));
    }
}

}  // namespace gmg
