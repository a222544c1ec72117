// host_copy_pool.h -- pageable -> pinned staging copies for libb200expect (host code only; included by b200_expect.cu
// and, on its own, by tests/cuda_emul/copy_pool_test.cpp which runs it under ThreadSanitizer / AddressSanitizer).
#pragma once
#include <emmintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace b200e_host {

// dst (64-byte aligned) = src with non-temporal stores: the pinned mirror is read next by the DMA engine, not by a
// core, so the lines need neither be fetched for ownership (a plain store reads the destination line first: three
// memory transfers per byte instead of two) nor kept in the caches.
inline void stream_copy(char *dst, const char *src, size_t bytes) {
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 63) == 0) {
        for (; i + 64 <= bytes; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
        }
        _mm_sfence();
    }
    if (i < bytes) memcpy(dst + i, src + i, bytes - i);
}

class CopyPool {
public:
    static CopyPool &get() { static CopyPool p; return p; }
    // dst[0, bytes) = src[0, bytes) in `parts` slices: slice 0 on the calling thread, slice w + 1 on worker w.
    // One copy at a time (callers on other handles wait their turn).
    void copy(void *dst, const void *src, size_t bytes) {
        const size_t kMinSlice = 256u << 10;
        const int parts = (int)std::min<size_t>(workers_.size() + 1, (bytes + kMinSlice - 1) / kMinSlice);
        if (parts <= 1) { stream_copy((char *)dst, (const char *)src, bytes); return; }
        std::lock_guard<std::mutex> user(user_mu_);
        const Job job{(char *)dst, (const char *)src, bytes, parts};
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = job;
            pending_.store(parts - 1);
            generation_.fetch_add(1);
        }
        cv_.notify_all();
        slice(job, 0);
        // the slices are of equal size: the others finish within microseconds of this one
        for (int spin = 0; pending_.load(std::memory_order_acquire) != 0; spin++)
            if (spin > 2000) std::this_thread::yield();
    }

private:
    struct Job { char *dst; const char *src; size_t bytes; int parts; };
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) hw = (unsigned)CPU_COUNT(&set);
        // measured on the 16-core B200 box, 160 MB per call (pinned: 3.07 ms): 4 copy streams 5.3 ms, 8 3.8 ms, 12 3.55 ms
        // one process per GPU (torchrun exports LOCAL_WORLD_SIZE): the ranks share the host's cores
        int share = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = std::max(1, atoi(e));
        int want = std::min(11, (int)hw * 3 / 4 / share - 1);  // + the caller
        if (const char *e = getenv("B200E_COPY_THREADS")) want = std::max(0, atoi(e) - 1);
        const int n = std::max(0, std::min<int>(want, (int)hw - 1));
        for (int w = 0; w < n; w++) workers_.emplace_back([this, w] { run(w); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    static void slice(const Job &j, int i) {
        const size_t per = ((j.bytes + j.parts - 1) / j.parts + 63) & ~(size_t)63;
        const size_t lo = std::min(j.bytes, per * (size_t)i), hi = std::min(j.bytes, lo + per);
        if (hi > lo) stream_copy(j.dst + lo, j.src + lo, hi - lo);
    }
    void run(int w) {
        unsigned long long seen = 0;
        for (;;) {
            Job job;
            // a call copies piece after piece: stay on the core for a short while before sleeping on the condition
            // variable (a wake-up costs tens of microseconds, a 4 MB piece takes about a hundred)
            const auto t0 = std::chrono::steady_clock::now();
            while (generation_.load(std::memory_order_acquire) == seen &&
                   std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(300)) {
#if defined(__x86_64__)
                _mm_pause();
#endif
            }
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_.load() != seen; });
                if (stop_) return;
                seen = generation_.load();
                job = job_;
            }
            // a worker without a slice in this job does nothing; one with a slice is waited for by copy(), so the
            // job it captured is the current one until it reports back
            if (w + 1 >= job.parts) continue;
            slice(job, w + 1);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, user_mu_;
    std::condition_variable cv_;
    Job job_{nullptr, nullptr, 0, 0};
    std::atomic<int> pending_{0};
    std::atomic<unsigned long long> generation_{0};
    bool stop_ = false;
};

}  // namespace b200e_host
