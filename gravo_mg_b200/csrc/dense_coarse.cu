#include "dense_coarse.h"

#include <algorithm>

#include "dense_factor.cuh"
#include "sparse_kernels.h"

namespace gmg {
namespace {

constexpr int NB = 64;  // panel / tile width

// One 64x64 output tile: C = alpha * A[:, k0:k1] * op(B)[k0:k1, :] + beta * C  (column-major).
struct GemmTask {
    const double* A;
    const double* B;
    double* C;
    int lda, ldb, ldc;
    int k0, k1;
    double alpha, beta;
};

constexpr int KC = 64;  // k-chunk staged in shared memory: one round of global loads per 64 k-steps
constexpr size_t kGemmSmem = 2 * (size_t)KC * (NB + 1) * sizeof(double);
constexpr size_t kPotrfSmem = (3 * (size_t)NB * (NB + 1) + NB + 2 * NB + 2 + NB) * sizeof(double);

template <bool TRANS_B>
__global__ void __launch_bounds__(256) gemm_tile_kernel(const GemmTask* __restrict__ tasks) {
    const GemmTask t = tasks[blockIdx.x];
    extern __shared__ double gsm[];
    double(*As)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(gsm);
    double(*Bs)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(gsm + KC * (NB + 1));
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // rows tx + 16 i, cols ty + 16 j (conflict-free shared reads)
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int kc = t.k0; kc < t.k1; kc += KC) {
        const int kn = min(KC, t.k1 - kc);
#pragma unroll
        for (int it = 0; it < KC * NB / 256; ++it) {
            const int e = tid + 256 * it;
            const int m = e & 63, k = e >> 6;
            if (k < kn) {
                As[k][m] = t.A[m + (size_t)(kc + k) * t.lda];
                if (TRANS_B) Bs[k][m] = t.B[m + (size_t)(kc + k) * t.ldb];  // B[n, k]
            }
            if (!TRANS_B) {
                const int kk = e & 63, n = e >> 6;
                if (kk < kn) Bs[kk][n] = t.B[(kc + kk) + (size_t)n * t.ldb];  // B[k, n]
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kn; ++k) {
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
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double* c = t.C + (size_t)(ty + 16 * j) * t.ldc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double v = t.alpha * acc[i][j];
            double* ce = c + tx + 16 * i;
            *ce = (t.beta == 0.0) ? v : fma(t.beta, *ce, v);
        }
    }
}

// Cholesky of one 64x64 diagonal block plus the inverse of its factor, one CTA of 256 threads.
// A (lower part read) -> L_jj written back to A (upper part zeroed); inv(L_jj) -> W block.
//   factor   right-looking on the unscaled columns: step k subtracts S[r][k] S[c][k] / S[k][k]
//            from the trailing lower triangle (one barrier per step, no serial sqrt/scale phase);
//            columns are scaled by 1/sqrt(pivot) once at the end.
//   inverse  the four 16x16 diagonal sub-blocks by forward substitution (one thread per column,
//            chains of <= 120 steps), then two doubling steps W21 = -W22 (L21 W11) as small
//            shared-memory products spread over all threads.
// Compact loops on purpose: a fully unrolled register version is instruction-fetch bound.
// 1/x to double precision without the long IEEE division sequence: single-precision seed and two
// Newton steps (relative error ~1e-16 for the well-scaled positive pivots this is used on).
__device__ __forceinline__ double fast_rcp(double x) {
    double r = (double)__frcp_rn((float)x);
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r;
}

#ifdef GMG_POTRF_CLK
__device__ long long g_potrf_clk[8];
#define POTRF_CLK(i) do { if (threadIdx.x == 0) g_potrf_clk[i] = clock64(); } while (0)
#else
#define POTRF_CLK(i) do { } while (0)
#endif

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* A, double* W, int ld, int j0, CycleControl* ctl) {
    extern __shared__ double psm[];
    double(*S)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(psm);
    double(*X)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(psm + NB * (NB + 1));
    double(*T)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(psm + 2 * NB * (NB + 1));
    double* dd = psm + 3 * NB * (NB + 1);
    const int tid = threadIdx.x;
    double* Ajj = A + j0 + (size_t)j0 * ld;
    double* Wjj = W + j0 + (size_t)j0 * ld;
    // ---- factor: thread (ty, tx) keeps the 4 x 4 entries (ty + 16 i, tx + 16 j) in registers; per
    // step only column k and the reciprocal pivot travel through shared memory (double-buffered,
    // one barrier per step), everything else is register arithmetic.
    double* colk = dd + NB;        // [2][NB]
    double* pinv = colk + 2 * NB;  // [2]
    double* pivots = pinv + 2;     // [NB]
    const int tx = tid & 15, ty = tid >> 4;
    double a[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = ty + 16 * i, cc = tx + 16 * j;
            a[i][j] = rr >= cc ? Ajj[rr + (size_t)cc * ld] : 0.0;
        }
    for (int e = tid; e < NB * NB; e += 256) X[e & 63][e >> 6] = 0.0;
    POTRF_CLK(0);
    // publish column 0 and 1 / pivot 0
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) colk[ty + 16 * i] = a[i][0];
        if (ty == 0) pinv[0] = fast_rcp(a[0][0]), pivots[0] = a[0][0];
    }
    __syncthreads();
    for (int k = 0; k < NB - 1; ++k) {
        const double* ck = colk + (k & 1) * NB;
        double* cn = colk + ((k + 1) & 1) * NB;
        const double pi = pinv[k & 1];
        double lr[4], lc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) lr[i] = ck[ty + 16 * i] * pi;
#pragma unroll
        for (int j = 0; j < 4; ++j) lc[j] = ck[tx + 16 * j];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int rr = ty + 16 * i, cc = tx + 16 * j;
                if (cc > k && rr >= cc) a[i][j] = fma(-lr[i], lc[j], a[i][j]);
            }
        // owners of column k + 1 publish it (it is final now) together with its reciprocal pivot
        const int kn = k + 1;
        if (tx == (kn & 15)) {
            const int jn = kn >> 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double v = jn == 0 ? a[i][0] : jn == 1 ? a[i][1] : jn == 2 ? a[i][2] : a[i][3];
                cn[ty + 16 * i] = v;
                if (ty + 16 * i == kn) pinv[kn & 1] = fast_rcp(v), pivots[kn] = v;
            }
        }
        __syncthreads();
    }
    POTRF_CLK(1);
    // ---- scale the columns by 1 / sqrt(pivot) and put L into shared memory for the inverse
    if (tid < NB) {
        double d = pivots[tid];
        if (!(d > 0.0) || d > 1.7976931348623157e308) {
            atomicOr(&ctl->error, 4);
            d = 1.0;
        }
        dd[tid] = sqrt(d);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = ty + 16 * i, cc = tx + 16 * j;
            S[rr][cc] = rr > cc ? a[i][j] / dd[cc] : (rr == cc ? dd[cc] : 0.0);
        }
    __syncthreads();
    POTRF_CLK(2);
    // ---- inverse of the lower-triangular S into X
    if (tid < NB) dd[tid] = 1.0 / S[tid][tid];  // reciprocal diagonal of L
    __syncthreads();
    if (tid < NB) {  // 16x16 diagonal sub-blocks: column c of block b
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
    POTRF_CLK(3);
    for (int b = 16; b < NB; b *= 2) {
        // for every pair of adjacent b x b diagonal blocks at offset o: T = L21 * W11, then W21 = -W22 * T
        const int per_pair = b * b, pairs = NB / (2 * b);
        for (int e = tid; e < pairs * per_pair; e += 256) {
            const int p = e / per_pair, q = e % per_pair;
            const int i = q % b, j = q / b, o = p * 2 * b;
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            int m = j;  // W11 lower: m >= j
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
            int m = 0;  // W22 lower: m <= i
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
    POTRF_CLK(4);
    for (int e = tid; e < NB * NB; e += 256) {
        const int rr = e & 63, cc = e >> 6;
        Ajj[rr + (size_t)cc * ld] = S[rr][cc];  // upper part is zero
        Wjj[rr + (size_t)cc * ld] = rr >= cc ? X[rr][cc] : 0.0;
    }
    POTRF_CLK(5);
}

__global__ void pad_identity_kernel(double* A, int ld, int n, int npad) {
    const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) A[i + (size_t)i * ld] = 1.0;
}

__global__ void transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int ld) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = bx + threadIdx.x, c = by + j;
        if (r < n && c < n) tile[j][threadIdx.x] = src[r + (size_t)c * ld];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = by + threadIdx.x, c = bx + j;  // dst[r, c] = src[c, r]
        if (r < n && c < n) dst[r + (size_t)c * ld] = tile[threadIdx.x][j];
    }
}

// out_c = sum over the stored triangle of column c of M: rows [0, c] (upper) or [c, n) (lower).
// One warp per column, rows read contiguously, eight independent 32-row strips in flight per lane
// (the loop is latency bound: a column is at most n / 32 strips long); the fixed strip / shuffle
// order keeps it deterministic. Launched with programmatic dependent launch like the row products.
template <int K>
__global__ void __launch_bounds__(256) tri_coldot_kernel(const double* __restrict__ M, int ld, int n,
                                                         const double* __restrict__ v, int v_ld,
                                                         double* __restrict__ out, int out_ld, int upper,
                                                         const CycleControl* ctl) {
    grid_dependency_wait();
    grid_launch_dependents();
    if (ctl && ctl->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(ctl, 101 + (upper ? 0 : 1));
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n) return;
    const int lo = upper ? 0 : c, hi = upper ? c + 1 : n;
    const double* col = M + (size_t)c * ld;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    constexpr int U = 8;
    for (int r0 = lo + lane; r0 < hi; r0 += 32 * U) {
        double m[U], w[U][K];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + 32 * u;
            const bool in = r < hi;
            m[u] = in ? col[r] : 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) w[u][k] = in ? v[(size_t)r * v_ld + k] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fma(m[u], w[u][k], acc[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) out[(size_t)c * out_ld + k] = acc[k];
}

template <int K>
void launch_coldot(const double* M, int ld, int n, const double* v, int v_ld, double* out, int out_ld, int upper,
                   const CycleControl* ctl, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n + 7) / 8), 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GMG_CUDA(cudaLaunchKernelEx(&cfg, tri_coldot_kernel<K>, M, ld, n, v, v_ld, out, out_ld, upper, ctl));
}

void coldot(int K, const double* M, int ld, int n, const double* v, int v_ld, double* out, int out_ld, int upper,
            const CycleControl* ctl, cudaStream_t s) {
    switch (K) {
        case 1: return launch_coldot<1>(M, ld, n, v, v_ld, out, out_ld, upper, ctl, s);
        case 2: return launch_coldot<2>(M, ld, n, v, v_ld, out, out_ld, upper, ctl, s);
        case 3: return launch_coldot<3>(M, ld, n, v, v_ld, out, out_ld, upper, ctl, s);
        case 4: return launch_coldot<4>(M, ld, n, v, v_ld, out, out_ld, upper, ctl, s);
    }
    throw std::invalid_argument("coarse solve: K must be 1..4 per pass");
}

}  // namespace

void DenseCoarseSolver::setup(int n, cudaStream_t stream) {
    if (n == n_ && L_.ptr) return;
    if (n > 16384) throw std::invalid_argument("coarsest level too large for the dense direct solve (> 16384 rows); lower `lower_bound`");
    n_ = n;
    npad_ = (n + NB - 1) / NB * NB;
    nb_ = npad_ / NB;
    const size_t elems = (size_t)npad_ * npad_;
    L_.ensure(elems);
    W_.ensure(elems);
    Wt_.ensure(elems);
    tmp_.ensure(elems);
    y_.ensure((size_t)npad_ * kMaxRhsTile);
    GMG_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem));
    GMG_CUDA(cudaFuncSetAttribute(gemm_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    GMG_CUDA(cudaFuncSetAttribute(gemm_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));

    const int ld = npad_;
    std::vector<GemmTask> tasks;
    chol_panel_.clear(), chol_update_.clear(), inv_first_.clear(), inv_second_.clear();
    auto Lp = [&](int r, int c) { return L_.ptr + r + (size_t)c * ld; };
    auto Wp = [&](int r, int c) { return W_.ptr + r + (size_t)c * ld; };
    auto Tp = [&](int r, int c) { return tmp_.ptr + r + (size_t)c * ld; };
    for (int j = 0; j < nb_; ++j) {
        Batch panel{(int)tasks.size(), 0, true};
        for (int i = j + 1; i < nb_; ++i)  // A_ij <- A_ij * inv(L_jj)^T
            tasks.push_back({Lp(NB * i, NB * j), Wp(NB * j, NB * j), Lp(NB * i, NB * j), ld, ld, ld, 0, NB, 1.0, 0.0});
        panel.count = (int)tasks.size() - panel.first;
        chol_panel_.push_back(panel);
        Batch upd{(int)tasks.size(), 0, true};
        for (int l = j + 1; l < nb_; ++l)
            for (int i = l; i < nb_; ++i)  // A_il -= A_ij * A_lj^T
                tasks.push_back({Lp(NB * i, NB * j), Lp(NB * l, NB * j), Lp(NB * i, NB * l), ld, ld, ld, 0, NB, -1.0, 1.0});
        upd.count = (int)tasks.size() - upd.first;
        chol_update_.push_back(upd);
    }
    // inv([L11 0; L21 L22]) = [W11 0; -W22 L21 W11, W22], block size doubling each level
    for (int b = NB; b < npad_; b *= 2) {
        Batch first{(int)tasks.size(), 0, false};
        for (int o = 0; o + b < npad_; o += 2 * b) {
            const int b2 = std::min(b, npad_ - o - b);
            for (int mi = 0; mi < b2 / NB; ++mi)
                for (int ni = 0; ni < b / NB; ++ni)  // T = L21 * W11, W11 lower triangular
                    tasks.push_back({Lp(o + b + NB * mi, o), Wp(o, o + NB * ni), Tp(o + b + NB * mi, o + NB * ni), ld, ld, ld,
                                     NB * ni, b, 1.0, 0.0});
        }
        first.count = (int)tasks.size() - first.first;
        inv_first_.push_back(first);
        Batch second{(int)tasks.size(), 0, false};
        for (int o = 0; o + b < npad_; o += 2 * b) {
            const int b2 = std::min(b, npad_ - o - b);
            for (int mi = 0; mi < b2 / NB; ++mi)
                for (int ni = 0; ni < b / NB; ++ni)  // W21 = -W22 * T, W22 lower triangular
                    tasks.push_back({Wp(o + b + NB * mi, o + b), Tp(o + b, o + NB * ni), Wp(o + b + NB * mi, o + NB * ni), ld, ld,
                                     ld, 0, NB * (mi + 1), -1.0, 0.0});
        }
        second.count = (int)tasks.size() - second.first;
        inv_second_.push_back(second);
    }
    // dataflow factorisation (dense_factor.cuh): tiles in dependency order — Cholesky column c, then
    // the inverse tiles of row c (they need L(c, .) and W_cc, all final once column c is done)
    std::vector<FactorTask> ftasks;
    for (int c = 0; c < nb_; ++c) {
        for (int i = c; i < nb_; ++i) ftasks.push_back({0, i, c});
        for (int j = 0; j < c; ++j) ftasks.push_back({1, c, j});
    }
    n_ftasks_ = (int)ftasks.size();
    ftasks_.ensure(ftasks.size() * sizeof(FactorTask));
    GMG_CUDA(cudaMemcpyAsync(ftasks_.ptr, ftasks.data(), ftasks.size() * sizeof(FactorTask), cudaMemcpyHostToDevice, stream));
    fflags_.ensure((size_t)2 * nb_ * nb_ + nb_ + 1);
    GMG_CUDA(cudaMemsetAsync(fflags_.ptr, 0, fflags_.count * sizeof(unsigned), stream));
    rdiag_.ensure(npad_);
    GMG_CUDA(cudaFuncSetAttribute(dense_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFactorSmem));
    int sms = 0, dev = 0;
    GMG_CUDA(cudaGetDevice(&dev));
    GMG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int per_sm = 0;
    GMG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_factor_kernel, 256, kFactorSmem));
    factor_grid_ = std::max(1, std::min(n_ftasks_, sms * std::max(per_sm, 1)));
    tasks_.ensure(std::max<size_t>(tasks.size(), 1) * sizeof(GemmTask));
    if (!tasks.empty())
        GMG_CUDA(cudaMemcpyAsync(tasks_.ptr, tasks.data(), tasks.size() * sizeof(GemmTask), cudaMemcpyHostToDevice, stream));
    GMG_CUDA(cudaStreamSynchronize(stream));  // `tasks` is a local
}

void DenseCoarseSolver::factor(const int* rowptr, const int* colidx, const double* vals, CycleControl* ctl,
                               cudaStream_t stream, bool profile) {
    const int ld = npad_;
    const size_t bytes = (size_t)npad_ * npad_ * sizeof(double);
    int launches = 0;
    // optional phase timing (debug aid, prints to stderr): 0 densify, 1 potrf, 2 panel, 3 update, 4 inverse, 5 transpose
    std::vector<cudaEvent_t> ev;
    std::vector<int> ev_phase;
    auto mark = [&](int phase) {
        if (!profile) return;
        cudaEvent_t e;
        GMG_CUDA(cudaEventCreate(&e));
        GMG_CUDA(cudaEventRecord(e, stream));
        ev.push_back(e);
        ev_phase.push_back(phase);
    };
    mark(0);
    GMG_CUDA(cudaMemsetAsync(L_.ptr, 0, bytes, stream));
    if (!dataflow_) GMG_CUDA(cudaMemsetAsync(W_.ptr, 0, bytes, stream));
    launch_csr_to_dense(n_, rowptr, colidx, vals, L_.ptr, ld, stream);
    ++launches;
    if (npad_ > n_) {
        pad_identity_kernel<<<(npad_ - n_ + 63) / 64, 64, 0, stream>>>(L_.ptr, ld, n_, npad_);
        ++launches;
    }
    if (dataflow_) {
        // one kernel: tiles of L, W = L^-1 and W^T as tasks ordered by per-tile flags (dense_factor.cuh)
        FactorArgs fa;
        fa.L = L_.ptr, fa.W = W_.ptr, fa.Wt = Wt_.ptr, fa.rdiag = rdiag_.ptr, fa.ld = ld, fa.nb = nb_;
        fa.tasks = reinterpret_cast<const FactorTask*>(ftasks_.ptr), fa.n_tasks = n_ftasks_;
        fa.next = fflags_.ptr;
        fa.flag_l = fflags_.ptr + 1, fa.flag_x = fa.flag_l + (size_t)nb_ * nb_, fa.flag_w = fa.flag_x + (size_t)nb_ * nb_;
        fa.epoch = ++epoch_;
        if (epoch_ == 0) fa.epoch = ++epoch_;  // flags start at 0
        fa.ctl = ctl;
        GMG_CUDA(cudaMemsetAsync(fflags_.ptr, 0, sizeof(unsigned), stream));
        mark(1);
        dense_factor_kernel<<<factor_grid_, 256, kFactorSmem, stream>>>(fa);
        GMG_CUDA(cudaGetLastError());
        ++launches;
        mark(6);
        factor_launches_ = launches;
        if (profile) {
            GMG_CUDA(cudaStreamSynchronize(stream));
            float ms0 = 0, ms1 = 0;
            GMG_CUDA(cudaEventElapsedTime(&ms0, ev[0], ev[1]));
            GMG_CUDA(cudaEventElapsedTime(&ms1, ev[1], ev[2]));
            std::fprintf(stderr, "[gravomg_b200] coarse factor n=%d (dataflow, %d tile tasks, grid %d): densify %.1f us, factor + inverse %.1f us\n",
                         n_, n_ftasks_, factor_grid_, 1e3 * ms0, 1e3 * ms1);
            for (auto e : ev) cudaEventDestroy(e);
        }
        return;
    }
    const GemmTask* tasks = reinterpret_cast<const GemmTask*>(tasks_.ptr);
    for (int j = 0; j < nb_; ++j) {
        mark(1);
        potrf_diag_kernel<<<1, 256, kPotrfSmem, stream>>>(L_.ptr, W_.ptr, ld, NB * j, ctl);
        ++launches;
        mark(2);
        if (chol_panel_[j].count) {
            gemm_tile_kernel<true><<<chol_panel_[j].count, 256, kGemmSmem, stream>>>(tasks + chol_panel_[j].first);
            ++launches;
        }
        mark(3);
        if (chol_update_[j].count) {
            gemm_tile_kernel<true><<<chol_update_[j].count, 256, kGemmSmem, stream>>>(tasks + chol_update_[j].first);
            ++launches;
        }
    }
    mark(4);
    for (size_t s = 0; s < inv_first_.size(); ++s) {
        gemm_tile_kernel<false><<<inv_first_[s].count, 256, kGemmSmem, stream>>>(tasks + inv_first_[s].first);
        gemm_tile_kernel<false><<<inv_second_[s].count, 256, kGemmSmem, stream>>>(tasks + inv_second_[s].first);
        launches += 2;
    }
    mark(5);
    dim3 tgrid((npad_ + 31) / 32, (npad_ + 31) / 32), tblock(32, 8);
    transpose_kernel<<<tgrid, tblock, 0, stream>>>(W_.ptr, Wt_.ptr, npad_, ld);
    ++launches;
    mark(6);
    GMG_CUDA(cudaGetLastError());
    factor_launches_ = launches;
    if (profile) {
        GMG_CUDA(cudaStreamSynchronize(stream));
        double tot[6] = {0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i + 1 < ev.size(); ++i) {
            float ms = 0;
            GMG_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            tot[ev_phase[i]] += ms;
        }
        std::fprintf(stderr, "[gravomg_b200] coarse factor n=%d: densify %.1f us, potrf %.1f, panel %.1f, update %.1f, inverse %.1f, transpose %.1f\n",
                     n_, 1e3 * tot[0], 1e3 * tot[1], 1e3 * tot[2], 1e3 * tot[3], 1e3 * tot[4], 1e3 * tot[5]);
        for (auto e : ev) cudaEventDestroy(e);
    }
}

void DenseCoarseSolver::solve(const double* b, double* x, int K, int ld, const CycleControl* ctl, cudaStream_t stream) {
    for (int k0 = 0; k0 < K; k0 += kMaxRhsTile) {
        const int kt = std::min(kMaxRhsTile, K - k0);
        // y = W b : row i of W is column i of Wt (upper triangle stored contiguously)
        coldot(kt, Wt_.ptr, npad_, n_, b + k0, ld, y_.ptr, kt, /*upper=*/1, ctl, stream);
        // x = W^T y : column c of W, rows c..n
        coldot(kt, W_.ptr, npad_, n_, y_.ptr, kt, x + k0, ld, /*upper=*/0, ctl, stream);
    }
}

}  // namespace gmg
