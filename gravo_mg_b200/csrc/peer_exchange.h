// Halo exchange over NVLink peer memory (multi-GPU V-cycle; new design, the reference
// is single-process).
//
// One process per GPU. Every rank puts the vectors of all levels plus a small mailbox into ONE
// cudaMalloc'ed arena with the same layout on every rank and exports it with CUDA IPC, so the
// address of any vector entry on a peer is `local address + peer_delta[q]`. Vectors are
// global-length and globally indexed (dist_plan.h), so a rank delivers the entries a peer needs by
// storing them at their global positions in the peer's copy — no pack / send / recv / unpack,
// no NCCL call inside the cycle:
//
//   peer_push_kernel   the rows of the freshly written vector that peers gather through A_k, R_k
//                      or U_k (send lists of dist_plan.h), or a whole row range (all-gather into
//                      the first replicated level), stored straight into the peers' HBM; then
//                      __threadfence_system, a release-store of this rank's epoch into every
//                      peer's mailbox, and an acquire-spin until every peer's epoch has arrived.
//   peer_norm_kernel   the stopping test: local partial sums of r^T M r and b^T M b are written
//                      into every peer's mailbox, same handshake, then every rank adds the slots in
//                      rank order (bitwise identical on all ranks) and applies the stopping rule.
//
// Every exchange is a handshake of all ranks (the box is one NVSwitch domain: <= 8 peers, each flag
// is one 8-byte NVLink store), which also orders the reuse of the ping-pong vectors: a rank can
// only be one exchange ahead of any other, and consecutive exchanges never touch the same vector.
// A spin that lasts longer than kPeerTimeoutNs sets CycleControl::error bit 8 instead of hanging.
#pragma once
#include "peer_fabric.cuh"
#include "sparse_kernels.h"

namespace gmg {

template <typename T>
struct PeerPushArgs {
    T* v = nullptr;                    // vector in the local arena
    T* v2 = nullptr;                   // optional second vector with the same rows
    int K = 1;
    const int* idx[kMaxPeers] = {nullptr};  // rows for peer q; nullptr: the contiguous range [first, first + count)
    int count[kMaxPeers] = {0};
    int first[kMaxPeers] = {0};
};

template <typename T>
void launch_peer_push(const PeerPushArgs<T>& args, const PeerFabric& fabric, CycleControl* ctl, cudaStream_t stream);

void launch_peer_norm(const double* partials, const NormChunks& chunks, int K, const PeerFabric& fabric, CycleControl* ctl,
                      double* hist_res, double* hist_ms, unsigned long long cond_handle, cudaStream_t stream);

}  // namespace gmg
