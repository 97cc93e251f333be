// Times the phases of potrf_diag_kernel and a 64x64x64 gemm tile in isolation. Not product code.
#define GMG_POTRF_CLK
#include "../gravo_mg_b200/csrc/dense_coarse.cu"
#include <cstdio>
#include <vector>
using namespace gmg;
int main() {
    const int n = 64, ld = 64;
    std::vector<double> h(n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) h[i + j * ld] = (i == j ? 70.0 : 1.0 / (1 + abs(i - j)));
    double *A, *W, *A0; CycleControl* ctl;
    cudaMalloc(&A, n * n * 8); cudaMalloc(&A0, n * n * 8); cudaMalloc(&W, n * n * 8); cudaMalloc(&ctl, sizeof(CycleControl));
    cudaMemcpy(A0, h.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemset(ctl, 0, sizeof(CycleControl));
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(A, A0, n * n * 8, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        potrf_diag_kernel<<<1, 256, kPotrfSmem>>>(A, W, ld, 0, ctl);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c[8]; cudaMemcpyFromSymbol(c, g_potrf_clk, sizeof c);
        printf("potrf total %.2f us (%s): factor loop %lld cyc, scale %lld, inv16 %lld, doubling %lld, store %lld\n", ms * 1e3, cudaGetErrorString(cudaGetLastError()),
               c[1] - c[0], c[2] - c[1], c[3] - c[2], c[4] - c[3], c[5] - c[4]);
    }
    // check: L L^T = A0, W L = I
    std::vector<double> L(n * n), Wh(n * n);
    cudaMemcpy(L.data(), A, n * n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Wh.data(), W, n * n * 8, cudaMemcpyDeviceToHost);
    double e1m = 0, e2m = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
        double s = 0, t = 0;
        for (int k = 0; k < n; ++k) s += L[i + k * ld] * L[j + k * ld], t += Wh[i + k * ld] * L[k + j * ld];
        e1m = fmax(e1m, fabs(s - h[i + j * ld])); e2m = fmax(e2m, fabs(t - (i == j)));
    }
    printf("max |LL^T - A| = %.3e, max |W L - I| = %.3e\n", e1m, e2m);
    // gemm tile K=64
    GemmTask t{A0, A0, W, ld, ld, ld, 0, 64, 1.0, 0.0}; GemmTask* dt; cudaMalloc(&dt, sizeof t); cudaMemcpy(dt, &t, sizeof t, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(gemm_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) gemm_tile_kernel<true><<<1, 256, kGemmSmem>>>(dt);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("gemm 64x64x64 tile: %.2f us per launch (20 back to back)\n", ms * 1e3 / 20);
    }
    return 0;
}
