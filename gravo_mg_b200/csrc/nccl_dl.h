// NCCL bound at run time (dlopen), so libgravomg_b200.so loads on hosts without NCCL and shares
// the copy torch.distributed already mapped into the process (same SONAME) when there is one.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

namespace gmg {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// Throws std::runtime_error when libnccl.so.2 cannot be loaded.
const NcclApi& nccl();

void nccl_check(ncclResult_t r, const char* what, const char* file, int line);
#define GMG_NCCL(expr) ::gmg::nccl_check((expr), #expr, __FILE__, __LINE__)

}  // namespace gmg
