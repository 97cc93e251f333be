// Sparse kernels of the V-cycle (sm_100a). All of them are memory-bound CSR row products
//     acc_i = sum_j A_ij * x_j            (K right-hand sides, row-major N x K vectors)
// followed by a per-row epilogue, so there is ONE kernel template with five epilogues:
//
//   EPI_SPMV      out_i = acc_i  [, out2_i = omega * dinv_i * acc_i]   restriction r_c = U^T r
//                                                  (reference multigrid_solver.cpp:1069)
//   EPI_JACOBI    out_i = x_i + omega * dinv_i * (b_i - acc_i)          smoother sweep
//                                                  (replaces the GS sweep of :1194-1226)
//   EPI_RESIDUAL  out_i = b_i - acc_i                                   (:1066)
//   EPI_ADD       out_i = xin_i + acc_i                                 x += U e (:1082)
//   EPI_NORM      partial sums of w_i (acc_i - b_i)^2 and w_i b_i^2     residualCheck (:1228-1277)
//   EPI_RESNORM   EPI_NORM and EPI_RESIDUAL from the same row product: the fp64 defect of the
//                 mixed-precision cycle (fp32 levels correct an fp64 iterate) and its stopping norm
//   EPI_NORMJAC   EPI_NORM and EPI_JACOBI from the same row product: the stopping test of one
//                 cycle and the first pre-smoothing sweep of the next both need b - A x
//
// Two data paths:
//   staged  persistent CTAs walk row tiles; the tile's contiguous (colidx, vals) slab is
//           brought into shared memory by the TMA engine (cp.async.bulk + mbarrier, a
//           STAGES-deep ring), then LANES threads own one row and read its entries from
//           shared memory (consecutive lanes, consecutive entries: conflict free), gather x
//           through L1/L2 and combine with a shuffle butterfly. With LANES = 1 the products
//           are summed in CSR order, bit-identical to a sequential CPU loop (the library is
//           compiled with -fmad=false for that reason); LANES > 1 trades that for more rows
//           in flight per byte of shared memory (short tiles, many resident CTAs).
//   direct  LANES threads per row straight from global memory with a shuffle reduction;
//           used when a row does not fit a stage and as an independent cross-check.
#pragma once
#include "common.cuh"
#include "peer_fabric.cuh"

namespace gmg {

enum Epilogue { EPI_SPMV = 0, EPI_JACOBI = 1, EPI_RESIDUAL = 2, EPI_ADD = 3, EPI_NORM = 4, EPI_NORMJAC = 5, EPI_RESNORM = 6 };

// Device-resident loop state of one solve (multigrid_solver.cpp:1411-1417).
struct CycleControl {
    int iter;       // V-cycles completed
    int done;       // 1: every cycle kernel returns immediately
    int max_iter;
    int criterion;  // stoppingCriteria 0..3
    double tol;
    double residue;
    unsigned long long t_start_ns;
    int error;      // sticky: 1 bad diagonal, 2 non-finite residual, 4 coarse factor breakdown, 8 peer timeout, 16 factor stall
    int n_cols;
    // optional device timeline (option "trace"): one (globaltimer ns, tag) pair per kernel of the cycle
    unsigned long long* trace;
    int trace_n, trace_cap;
};

#ifdef __CUDACC__
// One thread per kernel: the time at which the kernel's dependencies were met, and what it is.
__device__ __forceinline__ void trace_mark(const CycleControl* ctl, unsigned long long tag) {
    if (ctl == nullptr || ctl->trace == nullptr) return;
    CycleControl* c = const_cast<CycleControl*>(ctl);
    const int i = atomicAdd(&c->trace_n, 1);
    if (i < c->trace_cap) c->trace[2 * i] = global_timer_ns(), c->trace[2 * i + 1] = tag;
}

// Stopping rule of the cycle loop (multigrid_solver.cpp:1228-1277 norms, :1413-1417 loop test) from
// the 2K sums {sum w r^2, sum w b^2} per right-hand side; one thread. record = 0: only the residue.
__device__ inline void apply_stopping_rule(const double* sums, int K, CycleControl* ctl, double* hist_res, double* hist_ms,
                                           int record, unsigned long long cond_handle) {
    double residue = 0.0;
    bool diverged = false;
    if (ctl->criterion == 3) {
        double tot = 0.0;
        for (int k = 0; k < K; ++k) tot += sums[2 * k];
        residue = sqrt(tot);
        diverged = !(residue <= 1.7976931348623157e308);
    } else {
        for (int k = 0; k < K; ++k) {  // maxCoeff over the right-hand sides
            const double rk = sqrt(sums[2 * k] / sums[2 * k + 1]);
            if (k == 0 || rk > residue || rk != rk) residue = rk;
            // 0/0 (an all-zero right-hand side) is NaN upstream too and simply ends the loop;
            // a non-finite residual of a non-zero system means the smoother diverged
            if (!(rk <= 1.7976931348623157e308) && sums[2 * k + 1] > 0.0) diverged = true;
        }
    }
    ctl->residue = residue;
    if (diverged) ctl->error |= 2;
    if (record) {
        const int it = ctl->iter;
        hist_res[it] = residue;
        hist_ms[it] = (double)(global_timer_ns() - ctl->t_start_ns) * 1e-6;
        ctl->iter = it + 1;
        const int keep_going = (residue > ctl->tol) && (it + 1 < ctl->max_iter);
        ctl->done = !keep_going;
        if (cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, keep_going ? 1u : 0u);
    }
}
#endif

template <typename T>
struct SpmvArgs {
    int n_rows = 0;                    // direct path: rows [row_begin, n_rows) are processed
    int row_begin = 0;
    int ld = 1;                        // leading dimension of every vector (= total K of the solve)
    const int* rowptr = nullptr;
    const int* colidx = nullptr;
    const T* vals = nullptr;
    const T* x = nullptr;              // gathered vector
    const T* b = nullptr;              // JACOBI / RESIDUAL / NORM
    const T* xin = nullptr;            // ADD: own-row input
    const T* dinv = nullptr;           // JACOBI; SPMV when out2 != nullptr
    const double* weight = nullptr;    // NORM: per-row weight (nullptr = 1)
    T* out = nullptr;
    T* out2 = nullptr;
    T omega = T(0);
    const T* omega_ptr = nullptr;      // device-resident damping of this sweep (overrides omega)
    double* partials = nullptr;        // NORM: [gridDim.x][2*K]
    const int4* tile_desc = nullptr;   // staged: per tile {first row, end row, first entry & ~3, (end entry + 3) & ~3}
    int n_tiles = 0;
    int stage_elems = 0;               // staged: capacity of one stage in entries (multiple of 4)
    int stage_rows = 0;                // staged: rows per tile (consumer threads / LANES)
    const CycleControl* ctl = nullptr; // loop state (error bits)
    // multi-GPU, fused halo exchange (peer_exchange.h): rows of `out` (`out2` when push_out2) that
    // peers gather are also stored into the peers' copies of the vector as they are produced, the
    // last CTA signals the exchange; a consumer first waits for the peers' signal of the vector it gathers
    const PeerFabric* fabric = nullptr;
    const unsigned char* send_mask = nullptr;  // [rows of out] bit q: peer q needs this row
    int push_out2 = 0;
    int wait_peers = 0;
    // stopping test fused into the NORM / NORMJAC kernel (single GPU, K <= 4): the last CTA reduces the
    // per-CTA partial sums and applies the stopping rule (what norm_finalize_kernel does otherwise)
    unsigned* fin_ticket = nullptr;
    double* hist_res = nullptr;
    double* hist_ms = nullptr;
    unsigned long long cond_handle = 0;
    int l2_hint = 0;                   // staged: L2 eviction priority of the operator slabs (0 none, 1 keep, 2 stream)
    int n_early = 0;                   // staged: tiles [0, n_early) hold every row that is pushed or gathers halo
                                       // entries; the exchange is signalled once they are done (rest overlaps)
    // cancellation-free row product (square operators): `vals` then holds the ROW SUM s_i in place of the
    // diagonal entry (extract_dinv_kernel) and the kernels evaluate
    //     acc_i = sum_{j != i} A_ij (x_j - x_i) + s_i x_i        (= sum_j A_ij x_j in exact arithmetic).
    // A Poisson system tau M + S has row sums ~1e-12 next to entries ~1 and an iterate that carries a
    // constant ~1/(tau sqrt N): in the plain form the products of that constant cancel and leave a
    // rounding floor ~1e-6 in the relative residual at >= 4 M vertices (tools/stall_study.py).
    int diff = 0;
};

#ifdef __CUDACC__
template <typename T, int K>
__device__ __forceinline__ void peer_push_row(const SpmvArgs<T>& a, T* vec, int row, const T (&v)[K]) {
    unsigned m = a.send_mask[row];
    if (!m) return;
    const size_t o = (size_t)row * a.ld;
    while (m) {
        const int q = __ffs(m) - 1;
        m &= m - 1;
        T* pv = on_peer(vec, *a.fabric, q) + o;
#pragma unroll
        for (int k = 0; k < K; ++k) pv[k] = v[k];
    }
    // ordered before the exchange signal by the CTA barrier + system fence of peer_signal_from_cta
}
#endif

constexpr int kStagedThreads = 256;
constexpr int kStagedStages = 2;
constexpr int kDirectThreads = 256;

template <int NV, int TPB>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double* out) {
    __shared__ double sh[NV][TPB / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double s = v[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh[j][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < TPB / 32; ++w) s += sh[threadIdx.x][w];
        out[threadIdx.x] = s;
    }
}

// NORM kernels with a.fin_ticket: the last CTA of the grid adds the per-CTA partial sums (lane-strided
// over the CTAs, then a shuffle butterfly: the order norm_finalize_kernel uses) and applies the stopping rule.
template <int K>
__device__ __forceinline__ void fused_stopping_test(const double* partials, unsigned* ticket, CycleControl* ctl,
                                                    double* hist_res, double* hist_ms, unsigned long long cond_handle) {
    __shared__ int sh_last;
    __shared__ double sh_sums[2 * K];
    __threadfence();  // this CTA's partial sums before its ticket
    __syncthreads();
    if (threadIdx.x == 0) sh_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int j = warp; j < 2 * K; j += n_warps) {
        double s = 0.0;
        for (int blk = lane; blk < (int)gridDim.x; blk += 32) s += __ldcg(partials + (size_t)blk * 2 * K + j);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sh_sums[j] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0;
        trace_mark(ctl, 100);
        apply_stopping_rule(sh_sums, K, ctl, hist_res, hist_ms, 1, cond_handle);
    }
}

// Own-row operands of the epilogue. The row index comes from the tile descriptor in the stage, so they are
// requested right after the stage's full-barrier wait, together with the gathers of the row (one round trip).
template <typename T, int K>
struct RowOperands {
    T b[K];
    T xo[K];
    T scale;     // omega * dinv
    double w;    // NORM weight
};

template <typename T, int K, int EPI>
__device__ __forceinline__ void load_row_operands(const SpmvArgs<T>& a, int row, RowOperands<T, K>& r) {
    const size_t o = (size_t)row * a.ld;
    if (EPI == EPI_JACOBI || EPI == EPI_NORMJAC || (EPI == EPI_SPMV && a.out2)) {
        const T om = a.omega_ptr ? *a.omega_ptr : a.omega;
        r.scale = om * a.dinv[row];
    }
    if (EPI == EPI_JACOBI || EPI == EPI_RESIDUAL || EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) {
#pragma unroll
        for (int k = 0; k < K; ++k) r.b[k] = a.b[o + k];
    }
    if (EPI == EPI_JACOBI || EPI == EPI_NORMJAC) {
#pragma unroll
        for (int k = 0; k < K; ++k) r.xo[k] = a.x[o + k];
    }
    if (EPI == EPI_ADD) {
#pragma unroll
        for (int k = 0; k < K; ++k) r.xo[k] = a.xin[o + k];
    }
    if (EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) r.w = a.weight ? a.weight[row] : 1.0;
}

template <typename T, int K, int EPI>
__device__ __forceinline__ void row_epilogue(const SpmvArgs<T>& a, int row, const T (&acc)[K], const RowOperands<T, K>& r,
                                             double (&nrm)[2 * K]) {
    const size_t o = (size_t)row * a.ld;
    T res[K];
    if (EPI == EPI_SPMV) {
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = acc[k];
        if (a.out2) {
#pragma unroll
            for (int k = 0; k < K; ++k) res[k] = r.scale * acc[k];
#pragma unroll
            for (int k = 0; k < K; ++k) a.out2[o + k] = res[k];
            if (a.send_mask && a.push_out2) peer_push_row<T, K>(a, a.out2, row, res);
        }
        if (a.send_mask && !a.push_out2) peer_push_row<T, K>(a, a.out, row, acc);
    }
    if (EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double bk = (double)r.b[k];
            const double d = (double)acc[k] - bk;
            nrm[2 * k] += r.w * d * d;
            nrm[2 * k + 1] += r.w * bk * bk;
        }
    }
    if (EPI != EPI_SPMV && EPI != EPI_NORM) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (EPI == EPI_JACOBI || EPI == EPI_NORMJAC) res[k] = r.xo[k] + r.scale * (r.b[k] - acc[k]);
            else if (EPI == EPI_RESIDUAL || EPI == EPI_RESNORM) res[k] = r.b[k] - acc[k];
            else res[k] = r.xo[k] + acc[k];  // EPI_ADD
        }
#pragma unroll
        for (int k = 0; k < K; ++k) a.out[o + k] = res[k];
        if (a.send_mask) peer_push_row<T, K>(a, a.out, row, res);
    }
}

// ---------------------------------------------------------------------------- staged path
// Shared memory of one stage (all offsets multiples of 16 bytes):
//   int4   desc      {first row, end row, first entry (aligned down to 4), unused}
//   int    rowptr[stage_rows + 8]   the tile's slice of rowptr, from row (first row & ~3)
//   T      vals[stage_elems]
//   int    colidx[stage_elems]
// Warp 0 is the producer: one lane waits for a stage to drain (empty barrier), writes the
// descriptor and issues three bulk copies that complete on the stage's full barrier. Warps
// 1..8 consume: they never block on each other, only on the stage they need, so a fast warp
// runs up to STAGES tiles ahead of a slow one.
__host__ __device__ inline size_t staged_stage_bytes(int stage_rows, int stage_elems, size_t value_size) {
    return 16 + (size_t)(stage_rows + 8) * sizeof(int) + (size_t)stage_elems * (value_size + sizeof(int));
}

template <typename T, int K, int EPI, int LANES, int TPB = kStagedThreads, int STAGES = kStagedStages>
__global__ void __launch_bounds__(TPB + 32) spmv_staged_kernel(const SpmvArgs<T> a) {
    // TPB consumer threads + one producer warp; STAGES-deep ring

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // first 128 bytes: 2 * STAGES barriers
    uint64_t* empty = full + STAGES;
    const size_t stage_bytes = staged_stage_bytes(a.stage_rows, a.stage_elems, sizeof(T));
    unsigned char* stage0 = smem_raw + 128;

    const int tid = threadIdx.x;
    const int n_my = (int)blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], TPB / 32);
        }
        mbar_init_fence();
    }
    __syncthreads();

    double nrm[2 * K];
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) nrm[j] = 0.0;

    if (tid < 32) {
        // ------------------------------------------------------------------ producer warp
        // The operator (rowptr, colidx, vals) is constant while a solve runs, so its slabs are
        // requested right away: under programmatic dependent launch this CTA may start while the
        // previous kernel in the stream is still draining, and the copies overlap that tail.
        // Lane l prefetches the descriptor of tile iteration (32 * block + l); lane 0 issues.
        const uint64_t policy = l2_policy(a.l2_hint);
        for (int it0 = 0; it0 < n_my; it0 += 32) {
            int4 mine = make_int4(0, 0, 0, 0);
            if (it0 + tid < n_my) mine = a.tile_desc[blockIdx.x + (it0 + tid) * gridDim.x];
            const int count = min(32, n_my - it0);
            for (int j = 0; j < count; ++j) {
                int4 d;
                d.x = __shfl_sync(0xffffffffu, mine.x, j);
                d.y = __shfl_sync(0xffffffffu, mine.y, j);
                d.z = __shfl_sync(0xffffffffu, mine.z, j);
                d.w = __shfl_sync(0xffffffffu, mine.w, j);
                if (tid == 0) {
                    const int it = it0 + j;
                    const int s = it % STAGES;
                    if (it >= STAGES) mbar_wait(&empty[s], (uint32_t)((it / STAGES - 1) & 1));
                    unsigned char* st = stage0 + s * stage_bytes;
                    int* rp = reinterpret_cast<int*>(st + 16);
                    unsigned char* sv = st + 16 + (size_t)(a.stage_rows + 8) * sizeof(int);
                    unsigned char* sc = sv + (size_t)a.stage_elems * sizeof(T);
                    *reinterpret_cast<int4*>(st) = d;
                    const int rp0 = d.x & ~3;
                    const uint32_t n_rp = (uint32_t)(((d.y + 1 + 3) & ~3) - rp0);
                    const uint32_t cnt = (uint32_t)(d.w - d.z);
                    mbar_arrive_expect_tx(&full[s], n_rp * 4u + cnt * (uint32_t)(sizeof(T) + sizeof(int)));
                    if (a.l2_hint) {
                        bulk_copy_g2s_hint(rp, a.rowptr + rp0, n_rp * 4u, &full[s], policy);
                        if (cnt) {
                            bulk_copy_g2s_hint(sv, a.vals + d.z, cnt * (uint32_t)sizeof(T), &full[s], policy);
                            bulk_copy_g2s_hint(sc, a.colidx + d.z, cnt * (uint32_t)sizeof(int), &full[s], policy);
                        }
                    } else {
                        bulk_copy_g2s(rp, a.rowptr + rp0, n_rp * 4u, &full[s]);
                        if (cnt) {
                            bulk_copy_g2s(sv, a.vals + d.z, cnt * (uint32_t)sizeof(T), &full[s]);
                            bulk_copy_g2s(sc, a.colidx + d.z, cnt * (uint32_t)sizeof(int), &full[s]);
                        }
                    }
                }
                __syncwarp();
            }
        }
        grid_dependency_wait();  // (only so that this CTA does not retire before its predecessor grid)
    } else {
        // ------------------------------------------------------------------ consumers
        // Vectors (x, b, out of the previous kernel) must not be touched before that kernel is done.
        grid_dependency_wait();
        grid_launch_dependents();
        if (a.wait_peers) peer_wait_warp(*a.fabric, const_cast<int*>(&a.ctl->error));
        const int ctid = tid - 32;
        const int lane = ctid % LANES;
        if (blockIdx.x == 0 && ctid == 0) trace_mark(a.ctl, ((unsigned long long)a.n_rows << 8) | (unsigned)EPI);
        // fused halo push: this CTA's share of the early (boundary) tiles comes first in its walk
        const int my_early = (EPI != EPI_NORM && a.send_mask)
                                 ? ((int)blockIdx.x < a.n_early ? (a.n_early - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0)
                                 : -1;
        if (my_early == 0 && ctid == 0) peer_signal_from_cta(*a.fabric, false);
        for (int it = 0; it < n_my; ++it) {
            const int s = it % STAGES;
            const unsigned char* st = stage0 + s * stage_bytes;
            const int* rp = reinterpret_cast<const int*>(st + 16);
            const T* sv = reinterpret_cast<const T*>(st + 16 + (size_t)(a.stage_rows + 8) * sizeof(int));
            const int* sc = reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(sv) + (size_t)a.stage_elems * sizeof(T));

            mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));

            const int4 d = *reinterpret_cast<const int4*>(st);
            const int row = d.x + ctid / LANES;
            const bool active = row < d.y;
            int ps = 0, pe = 0;
            RowOperands<T, K> ops;
            if (active) {
                const int o = row - (d.x & ~3);
                ps = rp[o] - d.z;
                pe = rp[o + 1] - d.z;
                if (lane == 0) load_row_operands<T, K, EPI>(a, row, ops);
            }
            T acc[K], xi[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = T(0), xi[k] = T(0);
            if (a.diff && active) {
#pragma unroll
                for (int k = 0; k < K; ++k) xi[k] = __ldg(a.x + (size_t)row * a.ld + k);
            }
#pragma unroll 8
            for (int p = ps + lane; p < pe; p += LANES) {
                const int c = sc[p];
                const T v = sv[p];
                const T* xp = a.x + (size_t)c * a.ld;
                const bool off = a.diff && c != row;  // plain form: x_j - 0 = x_j, bit for bit
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] += v * (__ldg(xp + k) - (off ? xi[k] : T(0)));
            }
            // every lane of the warp has read its part of stage s: hand it back to the producer
            __syncwarp();
            if ((ctid & 31) == 0) mbar_arrive(&empty[s]);
            if (LANES > 1) {
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                }
            }
            if (active && lane == 0) row_epilogue<T, K, EPI>(a, row, acc, ops, nrm);
            if (it + 1 == my_early) {
                // the boundary rows of this CTA are stored here and in the peers' HBM: count the CTA
                // in; the last one signals the exchange while interior tiles are still being computed
                asm volatile("bar.sync 1, %0;" ::"n"(TPB) : "memory");  // consumer warps only
                if (ctid == 0) peer_signal_from_cta(*a.fabric, true);
            }
        }
    }
    if (EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) {
        block_sum_store<2 * K, TPB + 32>(nrm, a.partials + (size_t)blockIdx.x * 2 * K);
        if (a.fin_ticket)
            fused_stopping_test<K>(a.partials, a.fin_ticket, const_cast<CycleControl*>(a.ctl), a.hist_res, a.hist_ms, a.cond_handle);
    }
}

// ---------------------------------------------------------------------------- direct path
template <typename T, int K, int EPI, int LANES>
__global__ void __launch_bounds__(kDirectThreads) spmv_direct_kernel(const SpmvArgs<T> a) {
    constexpr int TPB = kDirectThreads;
    grid_dependency_wait();
    grid_launch_dependents();
    if (a.wait_peers) peer_wait_warp(*a.fabric, const_cast<int*>(&a.ctl->error));
    if (blockIdx.x == 0 && threadIdx.x == 0) trace_mark(a.ctl, ((unsigned long long)a.n_rows << 8) | (unsigned)EPI);
    const int lane = threadIdx.x % LANES;
    const int rows_per_block = TPB / LANES;
    double nrm[2 * K];
#pragma unroll
    for (int j = 0; j < 2 * K; ++j) nrm[j] = 0.0;

    // block-uniform trip count so the shuffles below always see full warps
    for (int first = a.row_begin + blockIdx.x * rows_per_block; first < a.n_rows; first += gridDim.x * rows_per_block) {
        const int row = first + threadIdx.x / LANES;
        const bool active = row < a.n_rows;
        T acc[K], xi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = T(0), xi[k] = T(0);
        if (active) {
            if (a.diff) {
#pragma unroll
                for (int k = 0; k < K; ++k) xi[k] = __ldg(a.x + (size_t)row * a.ld + k);
            }
            const int pe = a.rowptr[row + 1];
            for (int p = a.rowptr[row] + lane; p < pe; p += LANES) {
                const int c = __ldg(a.colidx + p);
                const T v = __ldg(a.vals + p);
                const T* xp = a.x + (size_t)c * a.ld;
                const bool off = a.diff && c != row;
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] += v * (__ldg(xp + k) - (off ? xi[k] : T(0)));
            }
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        if (active && lane == 0) {
            RowOperands<T, K> ops;
            load_row_operands<T, K, EPI>(a, row, ops);
            row_epilogue<T, K, EPI>(a, row, acc, ops, nrm);
        }
    }
    if (EPI == EPI_NORM || EPI == EPI_NORMJAC || EPI == EPI_RESNORM) {
        block_sum_store<2 * K, TPB>(nrm, a.partials + (size_t)blockIdx.x * 2 * K);
        if (a.fin_ticket)
            fused_stopping_test<K>(a.partials, a.fin_ticket, const_cast<CycleControl*>(a.ctl), a.hist_res, a.hist_ms, a.cond_handle);
    }
    if (EPI != EPI_NORM && a.send_mask) {  // no tile order here: signal when the whole CTA is done
        __syncthreads();
        if (threadIdx.x == 0) peer_signal_from_cta(*a.fabric, true);
    }
}

}  // namespace gmg
