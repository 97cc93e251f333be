#include "dense_coarse.h"

#include <algorithm>

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

template <bool TRANS_B>
__global__ void __launch_bounds__(256) gemm_tile_kernel(const GemmTask* __restrict__ tasks) {
    const GemmTask t = tasks[blockIdx.x];
    __shared__ double As[16][NB + 2];
    __shared__ double Bs[16][NB + 2];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // rows 4*tx.., cols 4*ty..
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int kc = t.k0; kc < t.k1; kc += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + 256 * i;
            const int m = e & 63, k = e >> 6;
            As[k][m] = t.A[m + (size_t)(kc + k) * t.lda];
            if (TRANS_B) {
                Bs[k][m] = t.B[m + (size_t)(kc + k) * t.ldb];  // B[n, k]
            } else {
                const int kk = e & 15, n = e >> 4;
                Bs[kk][n] = t.B[(kc + kk) + (size_t)n * t.ldb];  // B[k, n]
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][4 * tx + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][4 * ty + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double* c = t.C + (size_t)(4 * ty + j) * t.ldc + 4 * tx;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double v = t.alpha * acc[i][j];
            c[i] = (t.beta == 0.0) ? v : fma(t.beta, c[i], v);
        }
    }
}

// Cholesky of one 64x64 diagonal block in shared memory, plus the inverse of its factor.
// A (lower part read) -> L_jj written back to A (upper part zeroed); inv(L_jj) -> W block.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* A, double* W, int ld, int j0, CycleControl* ctl) {
    extern __shared__ double sm[];
    double(*S)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm);
    double(*X)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm + NB * (NB + 1));
    const int tid = threadIdx.x;
    double* Ajj = A + j0 + (size_t)j0 * ld;
    double* Wjj = W + j0 + (size_t)j0 * ld;
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e & 63, c = e >> 6;
        S[r][c] = Ajj[r + (size_t)c * ld];
        X[r][c] = 0.0;
    }
    __syncthreads();
    for (int k = 0; k < NB; ++k) {
        if (tid == 0) {
            double d = S[k][k];
            if (!(d > 0.0) || d > 1.7976931348623157e308) {
                atomicOr(&ctl->error, 4);
                d = 1.0;
            }
            S[k][k] = sqrt(d);
        }
        __syncthreads();
        if (tid > k && tid < NB) S[tid][k] /= S[k][k];
        __syncthreads();
        for (int e = tid; e < NB * NB; e += 256) {
            const int r = e & 63, c = e >> 6;
            if (c > k && r >= c) S[r][c] = fma(-S[r][k], S[c][k], S[r][c]);
        }
        __syncthreads();
    }
    if (tid < NB) {  // column tid of inv(L): forward substitution against e_tid
        const int c = tid;
        X[c][c] = 1.0 / S[c][c];
        for (int i = c + 1; i < NB; ++i) {
            double s = 0.0;
            for (int m = c; m < i; ++m) s = fma(S[i][m], X[m][c], s);
            X[i][c] = -s / S[i][i];
        }
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += 256) {
        const int r = e & 63, c = e >> 6;
        Ajj[r + (size_t)c * ld] = r >= c ? S[r][c] : 0.0;
        Wjj[r + (size_t)c * ld] = r >= c ? X[r][c] : 0.0;
    }
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
// One warp per column, rows read contiguously; the fixed lane/shuffle order keeps it deterministic.
template <int K>
__global__ void __launch_bounds__(256) tri_coldot_kernel(const double* __restrict__ M, int ld, int n,
                                                         const double* __restrict__ v, int v_ld,
                                                         double* __restrict__ out, int out_ld, int upper,
                                                         const CycleControl* ctl) {
    if (ctl && ctl->done) return;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n) return;
    const int lo = upper ? 0 : c, hi = upper ? c + 1 : n;
    const double* col = M + (size_t)c * ld;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int r = lo + lane; r < hi; r += 32) {
        const double m = col[r];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fma(m, v[(size_t)r * v_ld + k], acc[k]);
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
    tri_coldot_kernel<K><<<(n + 7) / 8, 256, 0, s>>>(M, ld, n, v, v_ld, out, out_ld, upper, ctl);
    GMG_CUDA(cudaGetLastError());
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
    GMG_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * NB * (NB + 1) * sizeof(double))));

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
    tasks_.ensure(std::max<size_t>(tasks.size(), 1) * sizeof(GemmTask));
    if (!tasks.empty())
        GMG_CUDA(cudaMemcpyAsync(tasks_.ptr, tasks.data(), tasks.size() * sizeof(GemmTask), cudaMemcpyHostToDevice, stream));
    GMG_CUDA(cudaStreamSynchronize(stream));  // `tasks` is a local
}

void DenseCoarseSolver::factor(const int* rowptr, const int* colidx, const double* vals, CycleControl* ctl,
                               cudaStream_t stream) {
    const int ld = npad_;
    const size_t bytes = (size_t)npad_ * npad_ * sizeof(double);
    int launches = 0;
    GMG_CUDA(cudaMemsetAsync(L_.ptr, 0, bytes, stream));
    GMG_CUDA(cudaMemsetAsync(W_.ptr, 0, bytes, stream));
    launch_csr_to_dense(n_, rowptr, colidx, vals, L_.ptr, ld, stream);
    ++launches;
    if (npad_ > n_) {
        pad_identity_kernel<<<(npad_ - n_ + 63) / 64, 64, 0, stream>>>(L_.ptr, ld, n_, npad_);
        ++launches;
    }
    const GemmTask* tasks = reinterpret_cast<const GemmTask*>(tasks_.ptr);
    const size_t potrf_smem = 2 * NB * (NB + 1) * sizeof(double);
    for (int j = 0; j < nb_; ++j) {
        potrf_diag_kernel<<<1, 256, potrf_smem, stream>>>(L_.ptr, W_.ptr, ld, NB * j, ctl);
        ++launches;
        if (chol_panel_[j].count) {
            gemm_tile_kernel<true><<<chol_panel_[j].count, 256, 0, stream>>>(tasks + chol_panel_[j].first);
            ++launches;
        }
        if (chol_update_[j].count) {
            gemm_tile_kernel<true><<<chol_update_[j].count, 256, 0, stream>>>(tasks + chol_update_[j].first);
            ++launches;
        }
    }
    for (size_t s = 0; s < inv_first_.size(); ++s) {
        gemm_tile_kernel<false><<<inv_first_[s].count, 256, 0, stream>>>(tasks + inv_first_[s].first);
        gemm_tile_kernel<false><<<inv_second_[s].count, 256, 0, stream>>>(tasks + inv_second_[s].first);
        launches += 2;
    }
    dim3 tgrid((npad_ + 31) / 32, (npad_ + 31) / 32), tblock(32, 8);
    transpose_kernel<<<tgrid, tblock, 0, stream>>>(W_.ptr, Wt_.ptr, npad_, ld);
    ++launches;
    GMG_CUDA(cudaGetLastError());
    factor_launches_ = launches;
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
