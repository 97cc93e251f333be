// Host <-> device transfers of caller-owned (pageable) buffers for the reference-facing call
// gmg_solve(host arrays in, host array out). The reference copies its arguments by value at the
// pybind11 boundary (core.cpp:21,68); here the copies go straight to HBM:
//
//   upload    worker threads copy 2 MB chunks of the caller's array into their own pinned slots
//             (two per thread) and issue the chunk's cudaMemcpyAsync on the solver stream, so the
//             host memcpy of chunk i+1 overlaps the DMA of chunk i and the link runs at pinned speed;
//   compare   the sparsity pattern of lhs is compared with the staged one by the same workers
//             (tasks of one parallel_for, so the comparison overlaps the value upload);
//   download  each worker pulls chunks over its own stream into a pinned slot and copies them out.
//
// A single pageable cudaMemcpy of the 56 MB of matrix values was the largest item of the
// end-to-end solve before this existed (bench.py `e2e`).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gmg {

class HostTransfer {
public:
    static constexpr size_t kChunk = 2u << 20;
    static constexpr size_t kDownChunk = 256u << 10;  // downloads: small chunks so that every worker copies out a share
    static constexpr int kSlots = 2;

    HostTransfer(int device, int threads) : device_(device) {
        n_threads_ = std::max(1, threads);
        workers_.resize(n_threads_);
        for (int t = 0; t < n_threads_; ++t) {
            Worker& w = workers_[t];
            GMG_CUDA(cudaHostAlloc((void**)&w.pinned, kChunk * kSlots, cudaHostAllocDefault));
            for (int s = 0; s < kSlots; ++s) GMG_CUDA(cudaEventCreateWithFlags(&w.ev[s], cudaEventDisableTiming));
            GMG_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        }
        // worker 0 is the calling thread; the others wait for jobs
        for (int t = 1; t < n_threads_; ++t) threads_.emplace_back([this, t] { loop(t); });
    }

    ~HostTransfer() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& th : threads_) th.join();
        cudaSetDevice(device_);
        for (Worker& w : workers_) {
            for (int s = 0; s < kSlots; ++s) cudaEventDestroy(w.ev[s]);
            cudaStreamDestroy(w.stream);
            cudaFreeHost(w.pinned);
        }
    }

    int threads() const { return n_threads_; }

    struct Copy {
        void* dev;
        const void* host;
        size_t bytes;
        cudaStream_t stream = nullptr;  // nullptr: the stream passed to upload_and_compare
    };
    struct Compare {
        const void* a;
        const void* b;
        size_t bytes;
    };

    // Issue the uploads on `stream` (asynchronously: returns when every chunk's copy has been
    // enqueued) and evaluate the comparisons meanwhile. Returns true when every pair compared equal.
    bool upload_and_compare(const std::vector<Copy>& copies, const std::vector<Compare>& compares, cudaStream_t stream) {
        struct Task {
            int kind;  // 0 copy, 1 compare
            size_t index, offset, bytes;
        };
        std::vector<Task> tasks;
        // comparisons are spread between the copy chunks so that both progress from the start
        std::vector<Task> cp, cm;
        for (size_t i = 0; i < copies.size(); ++i) {
            if (copies[i].bytes && page_locked(copies[i].host)) {
                // the caller's buffer is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory): the copy
                // engine reads it directly, no staging through the workers' pinned slots
                cudaStream_t cs = copies[i].stream ? copies[i].stream : stream;
                GMG_CUDA(cudaMemcpyAsync(copies[i].dev, copies[i].host, copies[i].bytes, cudaMemcpyHostToDevice, cs));
                continue;
            }
            for (size_t o = 0; o < copies[i].bytes; o += kChunk) cp.push_back({0, i, o, std::min(kChunk, copies[i].bytes - o)});
        }
        for (size_t i = 0; i < compares.size(); ++i)
            for (size_t o = 0; o < compares[i].bytes; o += kChunk) cm.push_back({1, i, o, std::min(kChunk, compares[i].bytes - o)});
        size_t a = 0, b = 0;
        while (a < cp.size() || b < cm.size()) {
            if (a < cp.size()) tasks.push_back(cp[a++]);
            if (b < cm.size() && (a * cm.size() >= b * cp.size() || a == cp.size())) tasks.push_back(cm[b++]);
        }
        std::atomic<bool> equal{true};
        parallel_for(tasks.size(), [&](size_t i, int t) {
            const Task& k = tasks[i];
            Worker& w = workers_[t];
            if (k.kind == 1) {
                if (!equal.load(std::memory_order_relaxed)) return;
                const Compare& c = compares[k.index];
                if (std::memcmp((const char*)c.a + k.offset, (const char*)c.b + k.offset, k.bytes) != 0) equal.store(false);
                return;
            }
            const Copy& c = copies[k.index];
            const int slot = w.next_slot;
            w.next_slot = (slot + 1) % kSlots;
            if (w.used[slot]) GMG_CUDA(cudaEventSynchronize(w.ev[slot]));
            char* pin = w.pinned + (size_t)slot * kChunk;
            std::memcpy(pin, (const char*)c.host + k.offset, k.bytes);
            cudaStream_t cs = c.stream ? c.stream : stream;
            GMG_CUDA(cudaMemcpyAsync((char*)c.dev + k.offset, pin, k.bytes, cudaMemcpyHostToDevice, cs));
            GMG_CUDA(cudaEventRecord(w.ev[slot], cs));
            w.used[slot] = true;
        });
        return equal.load();
    }

    // Synchronous device -> caller buffer copy. The producer of `dev` must have completed.
    void download(void* host, const void* dev, size_t bytes) {
        if (bytes && page_locked(host)) {  // page-locked destination: one direct copy
            GMG_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, workers_[0].stream));
            GMG_CUDA(cudaStreamSynchronize(workers_[0].stream));
            return;
        }
        const size_t n = (bytes + kDownChunk - 1) / kDownChunk;
        parallel_for(n, [&](size_t i, int t) {
            Worker& w = workers_[t];
            const size_t off = i * kDownChunk, len = std::min(kDownChunk, bytes - off);
            const int slot = w.next_slot;
            w.next_slot = (slot + 1) % kSlots;
            if (w.used[slot]) GMG_CUDA(cudaEventSynchronize(w.ev[slot]));
            w.used[slot] = false;
            char* pin = w.pinned + (size_t)slot * kChunk;
            GMG_CUDA(cudaMemcpyAsync(pin, (const char*)dev + off, len, cudaMemcpyDeviceToHost, w.stream));
            GMG_CUDA(cudaStreamSynchronize(w.stream));
            std::memcpy((char*)host + off, pin, len);
        });
    }

    // Is this host address inside page-locked memory known to CUDA?
    static bool page_locked(const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    }

private:
    struct Worker {
        char* pinned = nullptr;
        cudaEvent_t ev[kSlots] = {};
        bool used[kSlots] = {};
        int next_slot = 0;
        cudaStream_t stream = nullptr;
    };

    // fn(task index, worker index) over [0, n); the caller is worker 0. Exceptions of any worker
    // are rethrown on the calling thread after all workers have stopped.
    void parallel_for(size_t n, const std::function<void(size_t, int)>& fn) {
        if (n == 0) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &fn;
            job_n_ = n;
            next_.store(0);
            active_ = n_threads_ - 1;
            error_ = nullptr;
            ++generation_;
        }
        cv_.notify_all();
        run(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return active_ == 0; });
        job_ = nullptr;
        if (error_) std::rethrow_exception(error_);
    }

    void run(int t) {
        try {
            for (;;) {
                const size_t i = next_.fetch_add(1);
                if (i >= job_n_) break;
                (*job_)(i, t);
            }
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu_);
            if (!error_) error_ = std::current_exception();
            next_.store(job_n_);
        }
    }

    void loop(int t) {
        cudaSetDevice(device_);
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || generation_ != seen; });
                if (quit_) return;
                seen = generation_;
            }
            run(t);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --active_;
            }
            done_cv_.notify_one();
        }
    }

    int device_ = 0, n_threads_ = 1;
    std::vector<Worker> workers_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t, int)>* job_ = nullptr;
    size_t job_n_ = 0;
    std::atomic<size_t> next_{0};
    int active_ = 0;
    uint64_t generation_ = 0;
    bool quit_ = false;
    std::exception_ptr error_;
};

}  // namespace gmg
