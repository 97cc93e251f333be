// Timeline of the dataflow coarse factor (dense_factor.cuh) on a random SPD matrix. Not product code.
#define GMG_FACTOR_TRACE
#include "../gravo_mg_b200/csrc/dense_coarse.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace gmg;
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 783;
    // sparse-ish SPD operator as CSR: tridiagonal + a few long-range couplings, diagonally dominant
    std::vector<int> rp(n + 1), ci; std::vector<double> v;
    for (int i = 0; i < n; ++i) {
        rp[i] = (int)ci.size();
        for (int j = 0; j < n; ++j) {
            const int dist = abs(i - j);
            if (dist == 0) ci.push_back(j), v.push_back(8.0);
            else if (dist == 1 || dist == 17 || dist == 101) ci.push_back(j), v.push_back(-1.0);
        }
    }
    rp[n] = (int)ci.size();
    int *drp, *dci; double* dv; CycleControl* ctl;
    cudaMalloc(&drp, rp.size() * 4); cudaMalloc(&dci, ci.size() * 4); cudaMalloc(&dv, v.size() * 8); cudaMalloc(&ctl, sizeof(CycleControl));
    cudaMemcpy(drp, rp.data(), rp.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dci, ci.data(), ci.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, v.data(), v.size() * 8, cudaMemcpyHostToDevice); cudaMemset(ctl, 0, sizeof(CycleControl));
    DenseCoarseSolver s;
    s.setup(n, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        s.factor(drp, dci, dv, ctl, 0, false);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("factor n=%d: %.1f us (%s)\n", n, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    const int nb = (n + 63) / 64;
    std::vector<unsigned long long> tr(8192 * 4);
    cudaMemcpyFromSymbol(tr.data(), g_ftrace, tr.size() * 8);
    // task order: for c: chol (c..nb-1, c), then inverse (c, 0..c-1)
    unsigned long long t00 = tr[0];
    int t = 0;
    printf("times in us since the first claim: claim / inputs ready / posted(L) / end\n");
    for (int c = 0; c < nb; ++c) {
        for (int i = c; i < nb; ++i, ++t)
            if (i <= c + 1) printf("chol (%2d,%2d): %8.1f %8.1f %8.1f %8.1f\n", i, c, (tr[4*t]-t00)*1e-3, (tr[4*t+1]-t00)*1e-3, (tr[4*t+2]-t00)*1e-3, (tr[4*t+3]-t00)*1e-3);
        for (int j = 0; j < c; ++j, ++t)
            if (j == 0 || j == c - 1) printf("  inv (%2d,%2d): %8.1f %8.1f %8.1f %8.1f\n", c, j, (tr[4*t]-t00)*1e-3, (tr[4*t+1]-t00)*1e-3, (tr[4*t+2]-t00)*1e-3, (tr[4*t+3]-t00)*1e-3);
    }
    int err; cudaMemcpy(&err, &ctl->error, 4, cudaMemcpyDeviceToHost); printf("ctl error %d\n", err);
    return 0;
}
