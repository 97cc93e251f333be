#include "pcg_kernels.h"

namespace gmg {
namespace {

constexpr int kThreads = 256;

template <int NV>
__device__ __forceinline__ void block_tree_sum(double (&v)[NV], double* out) {
    __shared__ double sh[NV][kThreads];
#pragma unroll
    for (int j = 0; j < NV; ++j) sh[j][threadIdx.x] = v[j];
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
#pragma unroll
            for (int j = 0; j < NV; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j) out[j] = sh[j][0];
    }
}

template <int K>
__global__ void __launch_bounds__(kThreads) pcg_dot_kernel(int mode, int n, const double* __restrict__ a, const double* __restrict__ b,
                                                          double* __restrict__ partials, unsigned* ticket, PcgScalars* sc) {
    double s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) s[k] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] += a[(size_t)i * K + k] * b[(size_t)i * K + k];
    }
    block_tree_sum<K>(s, partials + (size_t)blockIdx.x * K);
    __shared__ int last;
    __threadfence();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < K) {
        double tot = 0.0;
        for (int blk = 0; blk < (int)gridDim.x; ++blk) tot += __ldcg(partials + (size_t)blk * K + threadIdx.x);
        const int k = threadIdx.x;
        if (mode == 0) {
            const double old = sc->rz[k];
            sc->beta[k] = old != 0.0 ? tot / old : 0.0;
            sc->rz[k] = tot;
        } else {
            sc->pq[k] = tot;
            sc->alpha[k] = tot != 0.0 ? sc->rz[k] / tot : 0.0;
        }
    }
    if (threadIdx.x == 0) *ticket = 0;
}

template <int K>
__global__ void __launch_bounds__(kThreads) pcg_direction_kernel(int n, const double* __restrict__ z, double* __restrict__ p,
                                                                const PcgScalars* __restrict__ sc) {
    double beta[K];
#pragma unroll
    for (int k = 0; k < K; ++k) beta[k] = sc->beta[k];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) p[(size_t)i * K + k] = z[(size_t)i * K + k] + beta[k] * p[(size_t)i * K + k];
    }
}

template <int K>
__global__ void __launch_bounds__(kThreads) pcg_update_kernel(int n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                                                             const double* __restrict__ q, const double* __restrict__ b,
                                                             const double* __restrict__ weight, const double* __restrict__ dinv,
                                                             const double* __restrict__ omega, double* __restrict__ z0,
                                                             const PcgScalars* __restrict__ sc, double* __restrict__ partials) {
    double alpha[K], nrm[2 * K];
#pragma unroll
    for (int k = 0; k < K; ++k) alpha[k] = sc->alpha[k], nrm[2 * k] = nrm[2 * k + 1] = 0.0;
    const double om = z0 ? *omega : 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double w = weight ? weight[i] : 1.0;
        const double scale = z0 ? om * dinv[i] : 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const size_t o = (size_t)i * K + k;
            x[o] = x[o] + alpha[k] * p[o];
            const double rr = r[o] - alpha[k] * q[o];
            r[o] = rr;
            if (z0) z0[o] = scale * rr;
            const double bk = b[o];
            nrm[2 * k] += w * rr * rr;
            nrm[2 * k + 1] += w * bk * bk;
        }
    }
    block_tree_sum<2 * K>(nrm, partials + (size_t)blockIdx.x * 2 * K);
}

int grid_for(int n) { return std::max(1, std::min(kPcgBlocks, (n + kThreads - 1) / kThreads)); }

}  // namespace

void launch_pcg_dot(int mode, int n, int K, const double* a, const double* b, double* partials, unsigned* ticket, PcgScalars* sc,
                    cudaStream_t s) {
    const int g = grid_for(n);
    switch (K) {
        case 1: pcg_dot_kernel<1><<<g, kThreads, 0, s>>>(mode, n, a, b, partials, ticket, sc); break;
        case 2: pcg_dot_kernel<2><<<g, kThreads, 0, s>>>(mode, n, a, b, partials, ticket, sc); break;
        case 3: pcg_dot_kernel<3><<<g, kThreads, 0, s>>>(mode, n, a, b, partials, ticket, sc); break;
        case 4: pcg_dot_kernel<4><<<g, kThreads, 0, s>>>(mode, n, a, b, partials, ticket, sc); break;
        default: throw std::invalid_argument("the conjugate-gradient wrapper handles 1..4 right-hand sides");
    }
    GMG_CUDA(cudaGetLastError());
}

void launch_pcg_direction(int n, int K, const double* z, double* p, const PcgScalars* sc, cudaStream_t s) {
    const int g = grid_for(n);
    switch (K) {
        case 1: pcg_direction_kernel<1><<<g, kThreads, 0, s>>>(n, z, p, sc); break;
        case 2: pcg_direction_kernel<2><<<g, kThreads, 0, s>>>(n, z, p, sc); break;
        case 3: pcg_direction_kernel<3><<<g, kThreads, 0, s>>>(n, z, p, sc); break;
        case 4: pcg_direction_kernel<4><<<g, kThreads, 0, s>>>(n, z, p, sc); break;
        default: throw std::invalid_argument("the conjugate-gradient wrapper handles 1..4 right-hand sides");
    }
    GMG_CUDA(cudaGetLastError());
}

int launch_pcg_update(int n, int K, double* x, double* r, const double* p, const double* q, const double* b, const double* weight,
                      const double* dinv, const double* omega, double* z0, const PcgScalars* sc, double* partials, cudaStream_t s) {
    const int g = grid_for(n);
    switch (K) {
        case 1: pcg_update_kernel<1><<<g, kThreads, 0, s>>>(n, x, r, p, q, b, weight, dinv, omega, z0, sc, partials); break;
        case 2: pcg_update_kernel<2><<<g, kThreads, 0, s>>>(n, x, r, p, q, b, weight, dinv, omega, z0, sc, partials); break;
        case 3: pcg_update_kernel<3><<<g, kThreads, 0, s>>>(n, x, r, p, q, b, weight, dinv, omega, z0, sc, partials); break;
        case 4: pcg_update_kernel<4><<<g, kThreads, 0, s>>>(n, x, r, p, q, b, weight, dinv, omega, z0, sc, partials); break;
        default: throw std::invalid_argument("the conjugate-gradient wrapper handles 1..4 right-hand sides");
    }
    GMG_CUDA(cudaGetLastError());
    return g;
}

}  // namespace gmg
