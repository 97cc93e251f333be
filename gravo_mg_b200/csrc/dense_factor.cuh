// Coarsest-level factorisation as ONE dataflow kernel (replaces Eigen::SimplicialLDLT::compute at
// reference multigrid_solver.cpp:1401 together with dense_coarse.cu's kernel-per-phase version).
//
// A = L L^T and W = L^-1 of the dense SPD coarse operator (n = low_bound .. ~6 low_bound rows) are
// cut into 64x64 tiles. Every tile is a task; CTAs claim tasks in dependency order from an atomic
// counter and wait on per-tile flags (release/acquire at gpu scope), so there is no grid barrier
// and no kernel boundary on the critical path
//     factor(diag j) -> L(j+1, j) -> factor(diag j+1) -> ...
// and everything else (left-looking GEMM accumulation of the other tiles, the triangular inverse)
// runs beside it on the other SMs.
//
//   Cholesky tile (i, j), i >= j   C = A_ij - sum_{k<j} L_ik L_jk^T               (fp64 register-tiled GEMM)
//        i == j                    L_jj = chol(C): 8-column panels; the 8x8 pivot block is factored
//                                  redundantly in registers by every thread (no broadcast), rows below
//                                  by one thread each, rank-8 update of the register tile;
//                                  afterwards W_jj = L_jj^-1 (off the critical path)
//        i >  j                    L_ij = C L_jj^-T by the same panel scheme (triangular solve, no
//                                  dependence on W_jj)
//   inverse tile (i, j), i > j     X_ij = -W_ii sum_{k=j}^{i-1} L_ik X_kj,  X_jj = W_jj   (W = L^-1, also stored transposed)
//
// Tasks are ordered so that a task only depends on tasks with a smaller index; a CTA that waits
// therefore always waits for a task that is running or done (no deadlock for any grid size).
#pragma once
#include "sparse_kernels.cuh"

namespace gmg {
namespace {

constexpr int FB = 64;       // tile
constexpr int FP = 8;        // panel width inside a tile
constexpr int FS = FB + 1;   // padded row stride of the 64x64 shared-memory tiles
constexpr int FPS = FP + 1;  // padded row stride of the panel buffers

struct FactorTask {
    int kind, i, j;  // kind 0: Cholesky tile L(i, j); 1: inverse tile X(i, j), i > j
};

struct FactorArgs {
    double* L;       // in: A (lower tiles read), out: L
    double* W;       // out: L^-1 (lower)
    double* Wt;      // out: (L^-1)^T (upper)
    double* rdiag;   // out: 1 / L_kk
    int ld, nb;
    const FactorTask* tasks;
    int n_tasks;
    unsigned* next;    // task counter (zeroed before the launch)
    unsigned* flag_l;  // [nb * nb] L(i, j) final        (== epoch)
    unsigned* flag_w;  // [nb]      W_jj final
    unsigned* flag_x;  // [nb * nb] X(i, j) final
    unsigned epoch;
    CycleControl* ctl;
};

// Optional per-task timeline for tools/factor_lab.cu (claim, inputs ready, tile posted, task end).
#ifdef GMG_FACTOR_TRACE
__device__ unsigned long long g_ftrace[8192][4];
#define FTRACE(task, slot) do { if (threadIdx.x == 0 && (task) < 8192) g_ftrace[task][slot] = global_timer_ns(); } while (0)
#else
#define FTRACE(task, slot) do { } while (0)
#endif

constexpr size_t kFactorSmem = (3 * (size_t)FB * FS + 2 * (size_t)FB * FPS + 2 * FB) * sizeof(double) + 16;

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Block until the tile behind `flag` is final (all threads; ends with a CTA barrier).
__device__ __forceinline__ void wait_tile(const unsigned* flag, unsigned epoch, CycleControl* ctl) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_gpu_u32(flag) != epoch) {
            __nanosleep(32);
            if ((++spins & 4095u) == 0) {  // a bug must not hang the device
                if (!t0) t0 = global_timer_ns();
                else if (global_timer_ns() - t0 > 2000000000ull) {
                    atomicOr(&ctl->error, 16);
                    break;
                }
            }
        }
    }
    __syncthreads();
}
// Publish a tile: every thread's global stores, then the flag.
__device__ __forceinline__ void post_tile(unsigned* flag, unsigned epoch) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu_u32(flag, epoch);
}

// dst[k][m] = G[m + k * ld]: a column-major 64x64 tile as the A operand (or as B of C = A B^T).
__device__ __forceinline__ void load_tile_cols(double (*dst)[FS], const double* G, int ld) {
#pragma unroll
    for (int it = 0; it < FB * FB / 256; ++it) {
        const int e = threadIdx.x + 256 * it;
        const int m = e & 63, k = e >> 6;
        dst[k][m] = __ldcg(G + m + (size_t)k * ld);
    }
}
// dst[r][c] = G[r + c * ld]: the tile as it is (B operand of C = A B, or a triangle to read by rows).
__device__ __forceinline__ void load_tile_rows(double (*dst)[FS], const double* G, int ld) {
#pragma unroll
    for (int it = 0; it < FB * FB / 256; ++it) {
        const int e = threadIdx.x + 256 * it;
        const int r = e & 63, c = e >> 6;
        dst[r][c] = __ldcg(G + r + (size_t)c * ld);
    }
}

// acc[i][j] += sum_k As[k][tx + 16 i] * Bs[k][ty + 16 j]   (acc[i][j] is C[tx + 16 i][ty + 16 j])
__device__ __forceinline__ void mma_tile(double (&acc)[4][4], const double (*As)[FS], const double (*Bs)[FS], int tx, int ty) {
#pragma unroll 8
    for (int k = 0; k < FB; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][tx + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][ty + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

// The 64x64 tile lives in registers: a[i][j] = C[tx + 16 i][ty + 16 j].
//   DIAG:  C (lower part) -> its Cholesky factor; rd[k] = 1 / L_kk is written to shared memory.
//   !DIAG: C -> C Ljj^-T, with Ljj[r][c] in shared memory (lower) and rd[] its reciprocal diagonal.
// Eight panels of eight columns; two CTA barriers per panel.
template <bool DIAG>
__device__ __forceinline__ void panel_factor(double (&a)[4][4], double (*Sp)[FPS], double (*Lp)[FPS],
                                             const double (*Ljj)[FS], double* rd, int tx, int ty, CycleControl* ctl) {
    const int tid = threadIdx.x;
    // The panel loop stays rolled: straight-line code that runs once is instruction-fetch bound
    // (measured: the unrolled version took 2.6 us per panel, tools/factor_lab.cu).
#pragma unroll 1
    for (int p = 0; p < FB / FP; ++p) {
        const int c0 = FP * p;
        // 1. the owners of columns c0 .. c0+7 publish their current values
        if ((ty >> 3) == (p & 1)) {
            const int jp = p >> 1;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                Sp[tx + 16 * i][ty & 7] = jp == 0 ? a[i][0] : jp == 1 ? a[i][1] : jp == 2 ? a[i][2] : a[i][3];
        }
        __syncthreads();
        // 2. the 8x8 pivot block and its reciprocal diagonal in registers, redundantly in every thread
        //    that owns a row in step 3 (no broadcast); the other warps leave the fp64 pipe to them
        double d[FP][FP], rs[FP];
        if (tid < FB) {
            if (DIAG) {
#pragma unroll
                for (int r = 0; r < FP; ++r)
#pragma unroll
                    for (int k = 0; k <= r; ++k) d[r][k] = Sp[c0 + r][k];
#pragma unroll
                for (int k = 0; k < FP; ++k) {
                    double piv = d[k][k];
                    if (!(piv > 0.0) || piv > 1.7976931348623157e308) {
                        if (tid == 0) atomicOr(&ctl->error, 4);
                        piv = 1.0;
                    }
                    rs[k] = rsqrt(piv);
                    d[k][k] = piv * rs[k];
#pragma unroll
                    for (int r = k + 1; r < FP; ++r) d[r][k] *= rs[k];
#pragma unroll
                    for (int j = k + 1; j < FP; ++j)
#pragma unroll
                        for (int r = j; r < FP; ++r) d[r][j] = fma(-d[r][k], d[j][k], d[r][j]);
                }
                if (tid == 0) {
#pragma unroll
                    for (int k = 0; k < FP; ++k) rd[c0 + k] = rs[k];
                }
            } else {
#pragma unroll
                for (int r = 0; r < FP; ++r)
#pragma unroll
                    for (int k = 0; k <= r; ++k) d[r][k] = Ljj[c0 + r][c0 + k];
#pragma unroll
                for (int k = 0; k < FP; ++k) rs[k] = rd[c0 + k];
            }
        }
        // 3. one thread per row: l = v D^-T (forward substitution against the pivot block)
        if (tid < FB) {
            const int row = tid;
            if (!DIAG || row >= c0 + FP) {
                double v[FP];
#pragma unroll
                for (int k = 0; k < FP; ++k) v[k] = Sp[row][k];
#pragma unroll
                for (int k = 0; k < FP; ++k) {  // right-looking inside the row: the updates of v[k+1..] are independent
                    const double l = v[k] * rs[k];
                    Lp[row][k] = l;
#pragma unroll
                    for (int j = k + 1; j < FP; ++j) v[j] = fma(-l, d[j][k], v[j]);
                }
            } else if (row >= c0) {
#pragma unroll
                for (int r = 0; r < FP; ++r)
                    if (row == c0 + r) {
#pragma unroll
                        for (int k = 0; k < FP; ++k) Lp[row][k] = k <= r ? d[r][k] : 0.0;
                    }
            } else {
#pragma unroll
                for (int k = 0; k < FP; ++k) Lp[row][k] = 0.0;
            }
        }
        __syncthreads();
        // 4. rank-8 update of the columns to the right; the panel's own columns become final.
        //    Row operands are loaded once per panel; DIAG skips the sub-blocks above the diagonal.
        double lr[4][FP];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < FP; ++k) lr[i][k] = Lp[tx + 16 * i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = ty + 16 * j;
            if (col >= c0 + FP) {
                double cv[FP];
#pragma unroll
                for (int k = 0; k < FP; ++k) cv[k] = DIAG ? Lp[col][k] : Ljj[col][c0 + k];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (DIAG && i < j) continue;  // rows tx + 16 i < 16 j <= col: above the diagonal
                    double s = a[i][j];
#pragma unroll
                    for (int k = 0; k < FP; ++k) s = fma(-lr[i][k], cv[k], s);
                    a[i][j] = s;
                }
            } else if (col >= c0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i][j] = Lp[tx + 16 * i][col - c0];
            }
        }
        // (the next panel's barrier after step 1 orders these reads of Lp before its rewrite)
    }
}

// X = S^-1 for the lower-triangular 64x64 S in shared memory (T is scratch, dd[64] too):
// 16x16 diagonal blocks by forward substitution, then two doubling steps W21 = -W22 (L21 W11).
__device__ __forceinline__ void tri_inverse_tile(const double (*S)[FS], double (*X)[FS], double (*T)[FS], double* dd) {
    const int tid = threadIdx.x;
    for (int e = tid; e < FB * FB; e += 256) X[e & 63][e >> 6] = 0.0;
    if (tid < FB) dd[tid] = 1.0 / S[tid][tid];
    __syncthreads();
    if (tid < FB) {
        const int b0 = tid & ~15, c = tid;
        X[c][c] = dd[c];
        for (int i = c + 1; i < b0 + 16; ++i) {
            double acc0 = 0.0, acc1 = 0.0;
            int m = c;
            for (; m + 1 < i; m += 2) {
                acc0 = fma(S[i][m], X[m][c], acc0);
                acc1 = fma(S[i][m + 1], X[m + 1][c], acc1);
            }
            if (m < i) acc0 = fma(S[i][m], X[m][c], acc0);
            X[i][c] = -(acc0 + acc1) * dd[i];
        }
    }
    __syncthreads();
    for (int b = 16; b < FB; b *= 2) {
        const int per_pair = b * b, pairs = FB / (2 * b);
        for (int e = tid; e < pairs * per_pair; e += 256) {
            const int p = e / per_pair, q = e % per_pair;
            const int i = q % b, j = q / b, o = p * 2 * b;
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            int m = j;
            for (; m + 3 < b; m += 4) {
                acc0 = fma(S[o + b + i][o + m], X[o + m][o + j], acc0);
                acc1 = fma(S[o + b + i][o + m + 1], X[o + m + 1][o + j], acc1);
                acc2 = fma(S[o + b + i][o + m + 2], X[o + m + 2][o + j], acc2);
                acc3 = fma(S[o + b + i][o + m + 3], X[o + m + 3][o + j], acc3);
            }
            for (; m < b; ++m) acc0 = fma(S[o + b + i][o + m], X[o + m][o + j], acc0);
            T[o + b + i][o + j] = (acc0 + acc1) + (acc2 + acc3);
        }
        __syncthreads();
        for (int e = tid; e < pairs * per_pair; e += 256) {
            const int p = e / per_pair, q = e % per_pair;
            const int i = q % b, j = q / b, o = p * 2 * b;
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            int m = 0;
            for (; m + 3 <= i; m += 4) {
                acc0 = fma(X[o + b + i][o + b + m], T[o + b + m][o + j], acc0);
                acc1 = fma(X[o + b + i][o + b + m + 1], T[o + b + m + 1][o + j], acc1);
                acc2 = fma(X[o + b + i][o + b + m + 2], T[o + b + m + 2][o + j], acc2);
                acc3 = fma(X[o + b + i][o + b + m + 3], T[o + b + m + 3][o + j], acc3);
            }
            for (; m <= i; ++m) acc0 = fma(X[o + b + i][o + b + m], T[o + b + m][o + j], acc0);
            X[o + b + i][o + j] = -((acc0 + acc1) + (acc2 + acc3));
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) dense_factor_kernel(const FactorArgs f) {
    extern __shared__ __align__(16) double fsm[];
    double(*As)[FS] = reinterpret_cast<double(*)[FS]>(fsm);
    double(*Bs)[FS] = reinterpret_cast<double(*)[FS]>(fsm + FB * FS);
    double(*Ts)[FS] = reinterpret_cast<double(*)[FS]>(fsm + 2 * FB * FS);
    double(*Sp)[FPS] = reinterpret_cast<double(*)[FPS]>(fsm + 3 * FB * FS);
    double(*Lp)[FPS] = reinterpret_cast<double(*)[FPS]>(fsm + 3 * FB * FS + FB * FPS);
    double* rd = fsm + 3 * FB * FS + 2 * FB * FPS;  // [64] reciprocal diagonal of the current diagonal tile
    double* dd = rd + FB;                           // [64] scratch of the triangular inverse
    int* sh_task = reinterpret_cast<int*>(dd + FB);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ld = f.ld, nb = f.nb;
    auto tile = [&](double* M, int i, int j) { return M + (size_t)FB * i + (size_t)FB * j * ld; };

    for (;;) {
        __syncthreads();
        if (tid == 0) *sh_task = (int)atomicAdd(f.next, 1u);
        __syncthreads();
        const int t = *sh_task;
        if (t >= f.n_tasks) break;
        const FactorTask task = f.tasks[t];
        const int i = task.i, j = task.j;
        FTRACE(t, 0);
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;

        if (task.kind == 0) {
            // ---------------------------------------------------------------- Cholesky tile (i, j)
            for (int k = 0; k < j; ++k) {
                wait_tile(&f.flag_l[i * nb + k], f.epoch, f.ctl);
                if (i != j) wait_tile(&f.flag_l[j * nb + k], f.epoch, f.ctl);
                load_tile_cols(As, tile(f.L, i, k), ld);
                load_tile_cols(Bs, tile(f.L, j, k), ld);
                __syncthreads();
                mma_tile(acc, As, Bs, tx, ty);
                __syncthreads();
            }
            FTRACE(t, 1);
            double* Aij = tile(f.L, i, j);
            double a[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) a[r][c] = Aij[(tx + 16 * r) + (size_t)(ty + 16 * c) * ld] - acc[r][c];
            if (i == j) {
                panel_factor<true>(a, Sp, Lp, Ts, rd, tx, ty, f.ctl);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int row = tx + 16 * r, col = ty + 16 * c;
                        const double v = row >= col ? a[r][c] : 0.0;
                        Aij[row + (size_t)col * ld] = v;
                        As[row][col] = v;  // S of the triangular inverse below
                    }
                __syncthreads();  // rd[] complete (written during the panels), As complete
                if (tid < FB) f.rdiag[FB * j + tid] = rd[tid];
                post_tile(&f.flag_l[j * nb + j], f.epoch);
                FTRACE(t, 2);
                // W_jj = L_jj^-1: needed by the inverse tiles only, so after L_jj is published
                tri_inverse_tile(As, Bs, Ts, dd);
                double* Wjj = tile(f.W, j, j);
                double* Wtjj = tile(f.Wt, j, j);
                for (int e = tid; e < FB * FB; e += 256) {
                    const int row = e & 63, col = e >> 6;
                    Wjj[row + (size_t)col * ld] = row >= col ? Bs[row][col] : 0.0;
                    Wtjj[row + (size_t)col * ld] = col >= row ? Bs[col][row] : 0.0;
                }
                post_tile(&f.flag_w[j], f.epoch);
                FTRACE(t, 3);
            } else {
                wait_tile(&f.flag_l[j * nb + j], f.epoch, f.ctl);
                FTRACE(t, 2);
                load_tile_rows(Ts, tile(f.L, j, j), ld);
                if (tid < FB) rd[tid] = __ldcg(f.rdiag + FB * j + tid);
                __syncthreads();
                panel_factor<false>(a, Sp, Lp, Ts, rd, tx, ty, f.ctl);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) Aij[(tx + 16 * r) + (size_t)(ty + 16 * c) * ld] = a[r][c];
                post_tile(&f.flag_l[i * nb + j], f.epoch);
                FTRACE(t, 3);
            }
        } else {
            // ---------------------------------------------------------------- inverse tile X(i, j), i > j
            for (int k = j; k < i; ++k) {
                wait_tile(&f.flag_l[i * nb + k], f.epoch, f.ctl);
                wait_tile(k == j ? &f.flag_w[j] : &f.flag_x[k * nb + j], f.epoch, f.ctl);
                load_tile_cols(As, tile(f.L, i, k), ld);
                load_tile_rows(Bs, tile(f.W, k, j), ld);
                __syncthreads();
                mma_tile(acc, As, Bs, tx, ty);
                __syncthreads();
            }
            FTRACE(t, 1);
            wait_tile(&f.flag_w[i], f.epoch, f.ctl);
            FTRACE(t, 2);
            load_tile_cols(As, tile(f.W, i, i), ld);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) Bs[tx + 16 * r][ty + 16 * c] = acc[r][c];
            __syncthreads();
            double x[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) x[r][c] = 0.0;
            mma_tile(x, As, Bs, tx, ty);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) Ts[tx + 16 * r][ty + 16 * c] = -x[r][c];
            __syncthreads();
            double* Wij = tile(f.W, i, j);
            double* Wtji = tile(f.Wt, j, i);
            for (int e = tid; e < FB * FB; e += 256) {
                const int row = e & 63, col = e >> 6;
                Wij[row + (size_t)col * ld] = Ts[row][col];
                Wtji[row + (size_t)col * ld] = Ts[col][row];
            }
            post_tile(&f.flag_x[i * nb + j], f.epoch);
            FTRACE(t, 3);
        }
    }
}

}  // namespace
}  // namespace gmg
