// Shared CUDA helpers: error checking, RAII device buffers, sm_100a async-copy primitives.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace gmg {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        std::snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorName(e), file, line, what);
        cudaGetLastError();  // clear the sticky-less error state
        throw CudaError(buf);
    }
}
#define GMG_CUDA(expr) ::gmg::cuda_check((expr), #expr, __FILE__, __LINE__)

// Device allocation owned by the solver handle. Arrays read by bulk copies are allocated with
// `pad` extra elements so a 16-byte aligned over-read at the tail stays inside the allocation.
template <typename T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    bool owned = true;   // false: a view into memory owned elsewhere (the multi-GPU peer arena)
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : ptr(o.ptr), count(o.count), owned(o.owned) { o.ptr = nullptr, o.count = 0, o.owned = true; }
    DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
        if (this != &o) {
            release();
            ptr = o.ptr, count = o.count, owned = o.owned;
            o.ptr = nullptr, o.count = 0, o.owned = true;
        }
        return *this;
    }
    ~DeviceBuffer() { release(); }
    void release() {
        if (ptr && owned) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
        owned = true;
    }
    // Point at `n` elements owned by someone else.
    void view(T* p, size_t n) {
        release();
        ptr = p, count = n, owned = false;
    }
    // Grow-only allocation; contents are not preserved.
    void ensure(size_t n, size_t pad = 0) {
        if (n + pad <= count && ptr && owned) return;
        release();
        GMG_CUDA(cudaMalloc((void**)&ptr, (n + pad ? n + pad : 1) * sizeof(T)));
        count = n + pad;
    }
    void zero(cudaStream_t s) { if (ptr) GMG_CUDA(cudaMemsetAsync(ptr, 0, count * sizeof(T), s)); }
    void upload(const T* host, size_t n, cudaStream_t s, size_t pad = 0) {
        ensure(n, pad);
        if (n) GMG_CUDA(cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
        if (pad) GMG_CUDA(cudaMemsetAsync(ptr + n, 0, pad * sizeof(T), s));
    }
    void upload(const std::vector<T>& host, cudaStream_t s, size_t pad = 0) { upload(host.data(), host.size(), s, pad); }
};

#ifdef __CUDACC__
// ---- mbarrier + bulk asynchronous copy (TMA engine, SASS UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
// Make barrier initialisation visible to the async proxy before the first bulk copy targets it.
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// One arrival that also announces `bytes` of pending bulk-copy traffic for the current phase.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Plain arrival (consumer releasing a stage back to the producer).
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy; dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// The same with an L2 eviction-priority hint for the source lines: hint 1 = evict_last (keep: the
// operator is re-read by the next kernels), 2 = evict_first (streamed once per cycle).
__device__ __forceinline__ uint64_t l2_policy(int hint) {
    uint64_t p = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_copy_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute
// starts while its predecessor in the stream drains; grid_dependency_wait() blocks until the
// predecessor has completed and its writes are visible (no-ops for a normal launch).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

}  // namespace gmg
