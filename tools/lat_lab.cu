// Latency probes for the dense coarse factor design (dependent fp64 chains, shuffles, barriers). Not product code.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_kernel(double* out, long long* clk, double a, double b) {
    __shared__ double sh[512];
    const int tid = threadIdx.x;
    double x = tid + 1.0;
    long long t0, t1;
    // 1: dependent DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fma(x, a, b);
    }
    t1 = clock64();
    if (tid == 0) clk[0] = t1 - t0;
    // 2: dependent DMUL chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = x * a;
    }
    t1 = clock64();
    if (tid == 0) clk[1] = t1 - t0;
    // 3: dependent double shuffle (2 x SHFL) chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = __shfl_sync(0xffffffffu, x, (tid + 1) & 31);
    }
    t1 = clock64();
    if (tid == 0) clk[2] = t1 - t0;
    // 4: barrier chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) __syncthreads();
    t1 = clock64();
    if (tid == 0) clk[3] = t1 - t0;
    // 5: STS -> bar -> LDS round trip chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
        sh[tid] = x;
        __syncthreads();
        x = sh[(tid + 1) % blockDim.x];
        __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) clk[4] = t1 - t0;
    // 6: reciprocal: f32 seed + 2 Newton steps, dependent chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
        double r = (double)__frcp_rn((float)x);
        r = r * fma(-x, r, 2.0);
        r = r * fma(-x, r, 2.0);
        x = r + 1.5;
    }
    t1 = clock64();
    if (tid == 0) clk[5] = t1 - t0;
    // 7: IEEE division + sqrt chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) x = 1.0 / sqrt(x + 2.0) + 1.0;
    t1 = clock64();
    if (tid == 0) clk[6] = t1 - t0;
    // 8: rsqrt builtin chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) x = rsqrt(x + 2.0) + 1.0;
    t1 = clock64();
    if (tid == 0) clk[7] = t1 - t0;
    // 9: independent DFMA throughput per warp: 16 accumulators
    double y[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) y[j] = x + j;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = fma(y[j], a, b);
    }
    t1 = clock64();
    if (tid == 0) clk[8] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x += y[j];
    // 10: LDS dependent chain (pointer chase in smem)
    __shared__ int nxt[512];
    nxt[tid] = (tid + 33) % blockDim.x;
    __syncthreads();
    int p = tid;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) p = nxt[p];
    t1 = clock64();
    if (tid == 0) clk[9] = t1 - t0;
    out[blockIdx.x * blockDim.x + tid] = x + p;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 128);
    const char* names[10] = {"dep DFMA", "dep DMUL", "dep SHFL.f64", "bar.sync", "STS-bar-LDS-bar", "rcp f32seed+2 newton (+add)", "1/sqrt ieee (+2 add)", "rsqrt builtin (+2 add)", "16 indep DFMA (per group of 16)", "dep LDS"};
    const double per[10] = {1024, 1024, 1024, 1024, 1024, 1024, 1024, 1024, 256, 1024};
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 2; ++rep) lat_kernel<<<1, threads>>>(o, c, 1.0000001, 1e-9);
        cudaDeviceSynchronize();
        long long h[10]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
        printf("threads=%d (%s)\n", threads, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 10; ++i) printf("  %-34s %8.1f cycles each\n", names[i], h[i] / per[i]);
    }
    return 0;
}
