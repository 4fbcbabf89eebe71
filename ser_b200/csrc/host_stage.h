// Host-side staging of PAGEABLE caller buffers (what a numpy caller of the Python mirror passes:
// ser_b200/dsp.py, data_loader.py, feature_extractor.py).  cudaMemcpyAsync from pageable memory is
// staged by the driver on the calling thread at the speed of one core and does not overlap the
// kernels; here the caller's bytes are gathered into a ring of pinned slots by a few worker threads
// and each slot travels as one asynchronous copy, so the transfer of chunk k + 1 overlaps the
// kernels of chunk k exactly as it does for pinned callers (bench.py: e2e.pageable).
//
// Reference side: the reference hands numpy arrays from soundfile / librosa.load
// (ser/_internal/utils/audio_utils.py:63-113) -- pageable by construction.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace serb {

// true for plain malloc / numpy memory, false for cudaHostAlloc / cudaHostRegister / managed memory
inline bool host_pointer_is_pageable(const void* p) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// memcpy of a task list on `threads` threads (the caller is one of them)
class CopyPool {
public:
    struct Task { void* dst; const void* src; size_t bytes; };

    explicit CopyPool(int threads) {
        for (int i = 1; i < threads; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
    CopyPool(const CopyPool&) = delete;
    CopyPool& operator=(const CopyPool&) = delete;

    void run(const std::vector<Task>& tasks) {
        // blocks of at most 256 KiB, handed out one at a time: files of any size balance
        blocks_.clear();
        for (const Task& t : tasks)
            for (size_t off = 0; off < t.bytes; off += kBlock)
                blocks_.push_back(Task{static_cast<char*>(t.dst) + off, static_cast<const char*>(t.src) + off,
                                       std::min(kBlock, t.bytes - off)});
        if (blocks_.empty()) return;
        if (workers_.empty() || blocks_.size() < 4) {
            for (const Task& b : blocks_) std::memcpy(b.dst, b.src, b.bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lock(mu_);
            next_ = 0;
            finished_ = 0;
            ++generation_;
        }
        cv_work_.notify_all();
        drain();
        std::unique_lock<std::mutex> lock(mu_);
        cv_done_.wait(lock, [this] { return finished_ == workers_.size(); });
    }

private:
    static constexpr size_t kBlock = 256u << 10;
    void drain() {
        for (;;) {
            size_t i;
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (next_ >= blocks_.size()) return;
                i = next_++;
            }
            std::memcpy(blocks_[i].dst, blocks_[i].src, blocks_[i].bytes);
        }
    }
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_work_.wait(lock, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
            }
            drain();
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (++finished_ == workers_.size()) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::vector<Task> blocks_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    size_t next_ = 0, finished_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

// ring of pinned slots; a slot is reused once the copy that read it has completed
struct StageRing {
    static constexpr size_t kSlotBytes = 32u << 20;
    static constexpr int kSlots = 3;
    char* base = nullptr;
    cudaEvent_t done[kSlots] = {};
    bool in_flight[kSlots] = {};
    int next = 0;
    CopyPool* pool = nullptr;

    cudaError_t ensure(int threads) {
        if (base) return cudaSuccess;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&base), kSlotBytes * kSlots, cudaHostAllocDefault);
        if (e != cudaSuccess) { base = nullptr; return e; }
        for (int i = 0; i < kSlots; ++i)
            if ((e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)) != cudaSuccess) {
                destroy();          // nothing half-built stays behind for the next call to trust
                return e;
            }
        pool = new CopyPool(std::max(1, threads));
        return cudaSuccess;
    }
    void destroy() {
        delete pool;
        pool = nullptr;
        if (!base) return;
        for (int i = 0; i < kSlots; ++i) {
            if (done[i]) cudaEventDestroy(done[i]);
            done[i] = nullptr;
            in_flight[i] = false;
        }
        cudaFreeHost(base);
        base = nullptr;
    }
};

// Appends (device destination, pageable source, bytes) runs; destinations that follow each other
// within a few bytes (the 16-byte alignment gaps of the device layouts) share a slot and one copy.
struct StagedWriter {
    StageRing* ring;
    cudaStream_t copy_stream;
    int slot = -1;
    char* dev_base = nullptr;
    size_t used = 0;
    std::vector<CopyPool::Task> tasks;

    cudaError_t add(void* dst_dev, const void* src, size_t bytes) {
        char* dst = static_cast<char*>(dst_dev);
        const char* s = static_cast<const char*>(src);
        while (bytes > 0) {
            if (slot < 0) {
                slot = ring->next;
                ring->next = (ring->next + 1) % StageRing::kSlots;
                if (ring->in_flight[slot]) {
                    const cudaError_t e = cudaEventSynchronize(ring->done[slot]);
                    if (e != cudaSuccess) return e;
                    ring->in_flight[slot] = false;
                }
                dev_base = dst;
                used = 0;
            }
            const ptrdiff_t off = dst - dev_base;
            if (off < static_cast<ptrdiff_t>(used) || off > static_cast<ptrdiff_t>(used) + 64 ||
                static_cast<size_t>(off) >= StageRing::kSlotBytes) {
                const cudaError_t e = flush();
                if (e != cudaSuccess) return e;
                continue;
            }
            const size_t n = std::min(bytes, StageRing::kSlotBytes - static_cast<size_t>(off));
            tasks.push_back(CopyPool::Task{ring->base + slot * StageRing::kSlotBytes + off, s, n});
            used = static_cast<size_t>(off) + n;
            dst += n;
            s += n;
            bytes -= n;
            if (used == StageRing::kSlotBytes) {
                const cudaError_t e = flush();
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    }
    cudaError_t flush() {
        if (slot < 0) return cudaSuccess;
        ring->pool->run(tasks);
        tasks.clear();
        cudaError_t e = cudaMemcpyAsync(dev_base, ring->base + slot * StageRing::kSlotBytes, used, cudaMemcpyHostToDevice, copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ring->done[slot], copy_stream);
        ring->in_flight[slot] = (e == cudaSuccess);
        slot = -1;
        return e;
    }
};

}  // namespace serb
