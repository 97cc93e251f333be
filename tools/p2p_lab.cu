// NVLink peer-memory flag latency between two GPUs of one box (single process, peer access). Not product code.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ void st_rel(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// mode 0: flag only; 1: 256 threads store 2 KB of data to the peer, fence.sys, then flag
__global__ void pingpong(int me, unsigned long long* my_flag, unsigned long long* peer_flag, double* peer_data, int iters, int mode, unsigned long long* out_ns) {
    __shared__ int dummy;
    unsigned long long t0 = gtime();
    for (int i = 1; i <= iters; ++i) {
        if (me == 1 || i > 1) {  // wait for the other side (rank 0 starts)
            if (threadIdx.x == 0) { while (ld_acq(my_flag) < (unsigned long long)(me == 0 ? i - 1 : i)) {} }
            __syncthreads();
        }
        if (mode == 1) { peer_data[threadIdx.x] = (double)i; __threadfence_system(); __syncthreads(); }
        if (threadIdx.x == 0) st_rel(peer_flag, (unsigned long long)i);
    }
    if (me == 0) { if (threadIdx.x == 0) { while (ld_acq(my_flag) < (unsigned long long)iters) {} } __syncthreads(); }
    if (threadIdx.x == 0) *out_ns = gtime() - t0;
    dummy = 0; (void)dummy;
}
__global__ void fence_cost(double* peer_data, double* local_data, unsigned long long* out) {
    unsigned long long t0 = gtime();
    for (int i = 0; i < 100; ++i) { local_data[threadIdx.x] = i; __threadfence_system(); }
    unsigned long long t1 = gtime();
    for (int i = 0; i < 100; ++i) { peer_data[threadIdx.x] = i; __threadfence_system(); }
    unsigned long long t2 = gtime();
    for (int i = 0; i < 100; ++i) { local_data[threadIdx.x] = i; __threadfence(); }
    unsigned long long t3 = gtime();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
}
int main() {
    int n = 0; CK(cudaGetDeviceCount(&n)); if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    unsigned long long *f[2], *o[2]; double* d[2];
    for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceEnablePeerAccess(1 - g, 0)); CK(cudaMalloc(&f[g], 64)); CK(cudaMemset(f[g], 0, 64)); CK(cudaMalloc(&o[g], 64)); CK(cudaMalloc(&d[g], 1 << 16)); }
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaMemset(f[g], 0, 64)); CK(cudaDeviceSynchronize()); }
        for (int g = 1; g >= 0; --g) { CK(cudaSetDevice(g)); pingpong<<<1, 256>>>(g, f[g], f[1 - g], d[1 - g], iters, mode, o[g]); }
        unsigned long long ns = 0;
        for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
        CK(cudaSetDevice(0)); CK(cudaMemcpy(&ns, o[0], 8, cudaMemcpyDeviceToHost));
        printf("mode %d (%s): %.2f us per round trip (2 one-way hops)\n", mode, mode ? "2 KB data + fence.sys + flag" : "flag only", ns * 1e-3 / iters);
    }
    CK(cudaSetDevice(0));
    fence_cost<<<1, 256>>>(d[1], d[0], o[0]);
    CK(cudaDeviceSynchronize());
    unsigned long long c[3]; CK(cudaMemcpy(c, o[0], 24, cudaMemcpyDeviceToHost));
    printf("store + __threadfence_system: local %.2f us, peer %.2f us; store + __threadfence (gpu): %.2f us\n", c[0] * 1e-5, c[1] * 1e-5, c[2] * 1e-5);
    return 0;
}
