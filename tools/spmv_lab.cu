// Kernel lab: times variants of the staged row-product kernel on a synthetic 1M-row, 7-entry
// banded matrix (the fine level of BASELINE config 2) outside the solver. Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo -I gravo_mg_b200/csrc tools/spmv_lab.cu -o tools/spmv_lab
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "sparse_kernels.cuh"

using namespace gmg;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Problem {
    int n; long nnz;
    int *rowptr, *colidx; double *vals, *x[3], *b, *dinv;
    std::vector<int> rp_h;
};

template <typename T> T* dev(const std::vector<T>& h, size_t pad = 16) {
    T* p; CK(cudaMalloc(&p, (h.size() + pad) * sizeof(T))); CK(cudaMemset(p, 0, (h.size() + pad) * sizeof(T)));
    CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return p;
}

Problem make(int side, int per_row_extra) {
    Problem P; P.n = side * side; const int n = P.n;
    std::vector<int> rp(n + 1), ci; std::vector<double> v;
    std::mt19937 rng(1); std::uniform_real_distribution<double> U(-1, 1);
    rp[0] = 0;
    for (int i = 0; i < n; ++i) {
        int offs[7] = {-side - 1, -side, -1, 0, 1, side, side + 1};
        std::vector<int> c;
        for (int o : offs) c.push_back(((i + o) % n + n) % n);
        for (int e = 0; e < per_row_extra; ++e) c.push_back(((i + (e + 2) * 7) % n + n) % n);
        std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end());
        for (int cc : c) { ci.push_back(cc); v.push_back(cc == i ? 6.0 : U(rng)); }
        rp[i + 1] = (int)ci.size();
    }
    P.nnz = ci.size(); P.rp_h = rp;
    P.rowptr = dev(rp); P.colidx = dev(ci); P.vals = dev(v);
    std::vector<double> h(n);
    for (auto& e : h) e = U(rng);
    for (int k = 0; k < 3; ++k) P.x[k] = dev(h);
    P.b = dev(h); for (auto& e : h) e = 1.0 / 6.0; P.dinv = dev(h);
    return P;
}

static bool g_pdl = false;
template <typename Kernel, typename Args>
void launch(Kernel kernel, int grid, int block, size_t smem, const Args& args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernel, args));
}

template <int LANES, int TPB, int STAGES>
float run_staged(const Problem& P, int ctas_per_sm_cap, int reps, const char* tag) {
    const int stage_rows = TPB / LANES;
    std::vector<int4> desc; int worst = 0;
    for (int r = 0; r < P.n; r += stage_rows) {
        int r1 = std::min(P.n, r + stage_rows);
        int p0 = P.rp_h[r] & ~3, p1 = (P.rp_h[r1] + 3) & ~3;
        desc.push_back(make_int4(r, r1, p0, p1)); worst = std::max(worst, p1 - p0);
    }
    int4* d_desc = dev(desc);
    SpmvArgs<double> a;
    a.n_rows = P.n; a.ld = 1; a.rowptr = P.rowptr; a.colidx = P.colidx; a.vals = P.vals;
    a.b = P.b; a.dinv = P.dinv; a.omega = 0.6; a.tile_desc = d_desc; a.n_tiles = (int)desc.size();
    a.stage_elems = (worst + 3) & ~3; a.stage_rows = stage_rows;
    auto kernel = spmv_staged_kernel<double, 1, EPI_JACOBI, LANES, TPB, STAGES>;
    size_t smem = 128 + (size_t)STAGES * staged_stage_bytes(stage_rows, a.stage_elems, 8);
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64));
    int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB + 32, smem));
    if (ctas_per_sm_cap > 0) per_sm = std::min(per_sm, ctas_per_sm_cap);
    int grid = std::min((int)desc.size(), per_sm * 148);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) { a.x = P.x[w % 3]; a.out = P.x[(w + 1) % 3]; launch(kernel, grid, TPB + 32, smem, a); }
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int w = 0; w < reps; ++w) { a.x = P.x[w % 3]; a.out = P.x[(w + 1) % 3]; launch(kernel, grid, TPB + 32, smem, a); }
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = 1e3 * ms / reps;
    const double bytes = P.nnz * 12.0 + P.n * 36.0;
    printf("%-28s lanes=%d tpb=%d stages=%d smem=%6zu ctas/sm=%d grid=%5d : %7.2f us  %6.0f GB/s (%.3f of 6551)\n", tag, LANES, TPB, STAGES, smem,
           per_sm, grid, us, bytes / us / 1e3, bytes / us / 1e3 / 6551.0);
    CK(cudaFree(d_desc));
    return (float)us;
}

// Upper bound: stream the same bytes with plain vectorised loads (no gathers, no row structure).
__global__ void stream_kernel(const double* __restrict__ vals, const int* __restrict__ col, const double* __restrict__ x,
                              const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ out, long nnz, int n) {
    const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x, nt = (long)gridDim.x * blockDim.x;
    double acc = 0;
    const double2* v2 = reinterpret_cast<const double2*>(vals);
    for (long i = tid; i < nnz / 2; i += nt) { double2 t = v2[i]; acc += t.x + t.y; }
    const int4* c4 = reinterpret_cast<const int4*>(col);
    for (long i = tid; i < nnz / 4; i += nt) { int4 t = c4[i]; acc += t.x + t.y + t.z + t.w; }
    for (long i = tid; i < n; i += nt) out[i] = x[i] + dinv[i] * b[i] + (acc == 12345.678 ? 1.0 : 0.0);
}

int main(int argc, char** argv) {
    const int side = argc > 1 ? atoi(argv[1]) : 1000;
    const int extra = argc > 2 ? atoi(argv[2]) : 0;
    const int reps = 60;
    Problem P = make(side, extra);
    printf("n=%d nnz=%ld (%.2f per row), algorithmic bytes per Jacobi sweep %.1f MB\n", P.n, P.nnz, (double)P.nnz / P.n, (P.nnz * 12.0 + P.n * 36.0) / 1e6);
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int g : {148 * 4, 148 * 8, 148 * 16}) {
            for (int w = 0; w < 3; ++w) stream_kernel<<<g, 512>>>(P.vals, P.colidx, P.x[0], P.b, P.dinv, P.x[1], P.nnz, P.n);
            CK(cudaEventRecord(e0));
            for (int w = 0; w < reps; ++w) stream_kernel<<<g, 512>>>(P.vals, P.colidx, P.x[w % 3], P.b, P.dinv, P.x[(w + 1) % 3], P.nnz, P.n);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = 1e3 * ms / reps, bytes = P.nnz * 12.0 + P.n * 32.0;
            printf("stream (LDG.128, no gather) grid=%5d : %7.2f us  %6.0f GB/s\n", g, us, bytes / us / 1e3);
        }
    }
    for (int pdl = 0; pdl < 2; ++pdl) {
        g_pdl = pdl;
        printf("---- programmatic dependent launch %s\n", pdl ? "ON" : "off");
        run_staged<1, 256, 2>(P, 0, reps, "staged");
        run_staged<1, 256, 3>(P, 0, reps, "staged");
        run_staged<1, 128, 2>(P, 0, reps, "staged");
        run_staged<1, 128, 3>(P, 0, reps, "staged");
        run_staged<1, 512, 2>(P, 0, reps, "staged");
        run_staged<2, 256, 2>(P, 0, reps, "staged");
        run_staged<2, 512, 2>(P, 0, reps, "staged");
    }
    return 0;
}
