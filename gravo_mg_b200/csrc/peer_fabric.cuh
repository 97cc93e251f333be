// Peer-memory fabric of the multi-GPU V-cycle: what a kernel needs to store into a peer's HBM
// and to signal / await an exchange (see peer_exchange.h for the protocol).
#pragma once
#include "common.cuh"

namespace gmg {

constexpr int kMaxPeers = 8;
constexpr int kPeerNormSlots = 64;  // 2 sums x up to 32 right-hand sides per rank
constexpr unsigned long long kPeerTimeoutNs = 10000000000ull;  // 10 s: ranks drift by host work, never by this much

// Mailbox at the start of every rank's arena (written by peers).
struct PeerMailbox {
    unsigned long long flags[kMaxPeers];                  // flags[q]: last epoch rank q has signalled
    double norm[2][kMaxPeers][kPeerNormSlots];            // [cycle parity][source rank][2K sums]
};

// Per-rank view of the box.
struct PeerFabric {
    int rank = 0, world = 1;
    PeerMailbox* box = nullptr;                           // local mailbox (arena offset 0)
    long long peer_delta[kMaxPeers] = {0};                // peer arena base - local arena base, bytes
    unsigned long long* epoch = nullptr;                  // local: exchanges completed so far
    unsigned int* ticket = nullptr;                       // local: last-block detection
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <typename P>
__device__ __forceinline__ P* on_peer(P* local, const PeerFabric& f, int q) {
    return reinterpret_cast<P*>(reinterpret_cast<char*>(local) + f.peer_delta[q]);
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Producer side of an exchange, called by ONE thread of every CTA of the grid exactly once, after
// the CTA's threads have stored the rows peers need and met at a barrier. `pushed`: this CTA
// stored into peer memory. The last CTA of the grid advances this rank's epoch and publishes it.
//
// Memory order (PTX model): a pushing CTA releases its stores at gpu scope (fence + ticket
// increment); the last CTA acquires all tickets and then releases at SYSTEM scope (one fence.sys,
// measured 2.6 us on B200, then the flag stores). Release is cumulative, so the peer that acquires
// the flag also sees the other CTAs' rows. Only the last CTA pays the system fence.
// Returns true in the CTA that published.
__device__ __forceinline__ bool peer_signal_from_cta(const PeerFabric& f, bool pushed) {
    if (pushed) __threadfence();
    const unsigned t = atomicAdd(f.ticket, 1u);
    if (t != gridDim.x - 1) return false;
    __threadfence_system();
    *f.ticket = 0;
    const unsigned long long e = *f.epoch + 1;
    *f.epoch = e;
    for (int q = 0; q < f.world; ++q)
        if (q != f.rank) st_relaxed_sys(&on_peer(f.box, f, q)->flags[f.rank], e);
    return true;
}

// Consumer side, called by a full warp: returns when every peer has signalled this rank's
// current epoch, i.e. the rows peers pushed into the vector about to be gathered are in place.
// `error` gets bit 8 after kPeerTimeoutNs (and stops later waits of the solve).
__device__ __forceinline__ void peer_wait_warp(const PeerFabric& f, int* error) {
    const int q = threadIdx.x & 31;
    if (q < f.world && q != f.rank) {
        const unsigned long long e = *const_cast<volatile unsigned long long*>(f.epoch);
        const unsigned long long t0 = global_timer_ns();
        while (!(*const_cast<volatile int*>(error) & 8) && ld_acquire_sys(&f.box->flags[q]) < e) {
            if (global_timer_ns() - t0 > kPeerTimeoutNs) {
                atomicOr(error, 8);
                break;
            }
        }
    }
    __syncwarp();
}
#endif

}  // namespace gmg
