// FP64 pipe throughput probe (DFMA / DMUL+DADD) on the device. Not product code.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    double* o; cudaMalloc(&o, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); dfma_kernel<<<148 * 4, 512>>>(o, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * 4 * 512 * iters * 8 * 2;
        printf("DFMA: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.9e9);
        cudaEventRecord(e0); ffma_kernel<<<148 * 4, 512>>>((float*)o, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA: %.3f ms  %.2f TFLOP/s\n", ms, fl / ms / 1e9);
    }
    return 0;
}
