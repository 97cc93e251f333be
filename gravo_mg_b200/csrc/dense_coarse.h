// Coarsest-level direct solve on the device (replaces Eigen::SimplicialLDLT at reference
// multigrid_solver.cpp:1401 (factor) and :1075 (solve)).
//
// The coarsest Galerkin operator has between low_bound and ~6*low_bound rows (1-6 k by
// default), so it is treated as a dense SPD matrix in fp64:
//   factor (once per solve):  A = L L^T by a blocked right-looking Cholesky (64-wide panels),
//                             then W = L^-1 by recursive doubling of triangular blocks;
//                             both are batched 64x64-tile fp64 GEMMs.
//   solve (once per cycle):   x = W^T (W b): two fully parallel triangular mat-vecs that
//                             stream W once each (HBM-bound, n_c^2 * 8 bytes per cycle) —
//                             no sequential substitution on the device.
#pragma once
#include "common.cuh"
#include "sparse_kernels.cuh"

namespace gmg {

class DenseCoarseSolver {
public:
    // Size the workspace for an n x n operator (idempotent for the same n).
    void setup(int n, cudaStream_t stream);
    // Densify the CSR operator and (re)compute L and W = L^-1. Sets ctl->error |= 4 on breakdown.
    void factor(const int* rowptr, const int* colidx, const double* vals, CycleControl* ctl, cudaStream_t stream,
                bool profile = false);
    // x = A^-1 b for K columns stored row-major with leading dimension ld. b and x may alias.
    void solve(const double* b, double* x, int K, int ld, const CycleControl* ctl, cudaStream_t stream);
    // true (default): one dataflow kernel (dense_factor.cuh); false: one kernel per phase
    void set_dataflow(bool on) { dataflow_ = on; }
    int size() const { return n_; }
    // factor storage for kernels that fuse the solve (tail_kernel.cuh): W = L^-1 (lower), Wt = W^T, scratch y
    const double* w() const { return W_.ptr; }
    const double* wt() const { return Wt_.ptr; }
    double* y() const { return y_.ptr; }
    int ld() const { return npad_; }
    int launches_per_solve() const { return 2; }
    int launches_per_factor() const { return factor_launches_; }
    size_t bytes_per_solve() const { return (size_t)n_ * (size_t)n_ * sizeof(double); }

private:
    struct Batch { int first, count; bool trans_b; };
    int n_ = 0, npad_ = 0, nb_ = 0;
    int factor_launches_ = 0;
    DeviceBuffer<double> L_, W_, Wt_, tmp_, y_;
    DeviceBuffer<unsigned char> tasks_;  // GemmTask array
    DeviceBuffer<unsigned char> ftasks_; // FactorTask array of the dataflow kernel
    DeviceBuffer<unsigned> fflags_;      // task counter + per-tile flags
    DeviceBuffer<double> rdiag_;
    int n_ftasks_ = 0, factor_grid_ = 1;
    unsigned epoch_ = 0;
    bool dataflow_ = true;
    std::vector<Batch> chol_panel_, chol_update_;  // per block column
    std::vector<Batch> inv_first_, inv_second_;    // per doubling level
};

}  // namespace gmg
