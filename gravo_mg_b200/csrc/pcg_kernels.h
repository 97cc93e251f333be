// Vector kernels of the conjugate-gradient wrapper around the V-cycle (the cycle as a
// preconditioner; `solverType == 4` of the reference, multigrid_solver.cpp:1453-1477, is plain CG through
// Eigen::ConjugateGradient with the identity preconditioner — option krylov = 2 here).
//
// All reductions are deterministic: every CTA adds its grid-strided share in a fixed tree, the last CTA to
// finish (ticket) adds the per-CTA partials in CTA order and derives the step lengths on the device, so an
// iteration needs no host round trip. K <= 4 right-hand sides, each with its own alpha / beta.
#pragma once
#include "common.cuh"
#include "sparse_kernels.cuh"

namespace gmg {

constexpr int kPcgMaxK = 4;
constexpr int kPcgBlocks = 148 * 4;

struct PcgScalars {
    double rz[kPcgMaxK];     // <r, z> of the previous iteration (0 before the first: beta = 0)
    double pq[kPcgMaxK];
    double alpha[kPcgMaxK];
    double beta[kPcgMaxK];
};

// mode 0: rz_new = <a, b>, beta = rz_new / rz (0 when rz == 0), rz = rz_new
// mode 1: pq = <a, b>, alpha = rz / pq (0 when pq == 0)
void launch_pcg_dot(int mode, int n, int K, const double* a, const double* b, double* partials, unsigned* ticket, PcgScalars* sc,
                    cudaStream_t s);
// p = z + beta p
void launch_pcg_direction(int n, int K, const double* z, double* p, const PcgScalars* sc, cudaStream_t s);
// x += alpha p, r -= alpha q; partial sums {sum w r^2, sum w b^2} per column for the stopping test (layout of the
// NORM epilogue: partials[cta][2K]); z0 = omega * dinv * r, the first smoothing sweep of the next cycle from a
// zero guess (z0 may be null). Returns the grid size.
int launch_pcg_update(int n, int K, double* x, double* r, const double* p, const double* q, const double* b, const double* weight,
                      const double* dinv, const double* omega, double* z0, const PcgScalars* sc, double* partials, cudaStream_t s);

}  // namespace gmg
