#include "peer_exchange.h"

namespace gmg {
namespace {

// All threads of the calling kernel have stored their data into peer memory. One thread per block
// counts the block in (peer_signal_from_cta: the last block publishes this rank's next epoch with
// a single system-scope fence); the last block then waits until every peer has published the same
// epoch, i.e. the peers' stores into OUR arena are complete and visible.
__device__ void peer_handshake(const PeerFabric& f, CycleControl* ctl) {
    __shared__ int sh_last;
    __syncthreads();
    if (threadIdx.x == 0) sh_last = peer_signal_from_cta(f, true) ? 1 : 0;
    __syncthreads();
    if (!sh_last) return;
    peer_wait_warp(f, &ctl->error);  // every warp of the last block (cheap), so all its threads may read peer data after
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) peer_push_kernel(const PeerPushArgs<T> a, const PeerFabric f, CycleControl* ctl) {
    grid_dependency_wait();    // the kernel that wrote v must be complete
    grid_launch_dependents();  // the consumer may prefetch its operator slabs; it waits for this grid before it gathers
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(ctl, 103);
    const int K = a.K;
    for (int q = 0; q < f.world; ++q) {
        const int cnt = a.count[q];
        if (q == f.rank || cnt == 0) continue;
        T* pv = on_peer(a.v, f, q);
        T* pv2 = a.v2 ? on_peer(a.v2, f, q) : nullptr;
        const int* idx = a.idx[q];
        const long long total = (long long)cnt * K;
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
            const int i = (int)(e / K), k = (int)(e - (long long)i * K);
            const int row = idx ? idx[i] : a.first[q] + i;
            const size_t o = (size_t)row * K + k;
            pv[o] = a.v[o];
            if (pv2) pv2[o] = a.v2[o];
        }
    }
    peer_handshake(f, ctl);
}

__global__ void __launch_bounds__(256) peer_norm_kernel(const double* __restrict__ partials, NormChunks chunks, int K,
                                                        const PeerFabric f, CycleControl* ctl, double* hist_res,
                                                        double* hist_ms, unsigned long long cond_handle) {
    grid_dependency_wait();
    grid_launch_dependents();
    if (threadIdx.x == 0) trace_mark(ctl, 104);
    __shared__ double sums[kPeerNormSlots];
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
    // slot parity alternates per cycle: a fast rank may already deliver the sums of cycle i + 1
    // while a slow one still adds up those of cycle i
    const int par = ctl->iter & 1;
    for (int t = threadIdx.x; t < f.world * n_sums; t += blockDim.x) {
        const int q = t / n_sums, j = t - q * n_sums;
        PeerMailbox* box = q == f.rank ? f.box : on_peer(f.box, f, q);
        box->norm[par][f.rank][j] = sums[j];
    }
    peer_handshake(f, ctl);  // one block: it is the last one
    if (threadIdx.x < n_sums) {
        double s = 0.0;
        for (int q = 0; q < f.world; ++q) s += *const_cast<volatile double*>(&f.box->norm[par][q][threadIdx.x]);
        sums[threadIdx.x] = s;  // rank order: the same bits on every rank
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (ctl->done || (ctl->error & 8)) {  // stopped already, or a peer never answered: end the loop
            ctl->done = 1;
            if (cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, 0);
            return;
        }
        apply_stopping_rule(sums, K, ctl, hist_res, hist_ms, 1, cond_handle);
    }
}

void launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, cudaStream_t stream) {
    cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

}  // namespace

template <typename T>
void launch_peer_push(const PeerPushArgs<T>& args, const PeerFabric& fabric, CycleControl* ctl, cudaStream_t stream) {
    long long most = 0;
    for (int q = 0; q < fabric.world; ++q)
        if (q != fabric.rank) most = std::max(most, (long long)args.count[q] * args.K);
    const int grid = (int)std::min<long long>(std::max<long long>((most + 1023) / 1024, 1), 148);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, grid, stream);
    GMG_CUDA(cudaLaunchKernelEx(&cfg, peer_push_kernel<T>, args, fabric, ctl));
}
template void launch_peer_push<double>(const PeerPushArgs<double>&, const PeerFabric&, CycleControl*, cudaStream_t);
template void launch_peer_push<float>(const PeerPushArgs<float>&, const PeerFabric&, CycleControl*, cudaStream_t);

void launch_peer_norm(const double* partials, const NormChunks& chunks, int K, const PeerFabric& fabric, CycleControl* ctl,
                      double* hist_res, double* hist_ms, unsigned long long cond_handle, cudaStream_t stream) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, 1, stream);
    GMG_CUDA(cudaLaunchKernelEx(&cfg, peer_norm_kernel, partials, chunks, K, fabric, ctl, hist_res, hist_ms, cond_handle));
}

}  // namespace gmg
