// copy_pool_test.cpp -- TEST INFRASTRUCTURE: the library's pageable -> pinned staging pool (csrc/host_copy_pool.h) on its
// own, for ThreadSanitizer / AddressSanitizer runs (tests/test_sanitizers.py).  Copies of many sizes and alignments, back
// to back (generation hand-over between jobs, workers with and without a slice), from two caller threads at once.
#include "host_copy_pool.h"

#include <cstdio>

static int run_caller(unsigned seed, int rounds) {
    std::vector<size_t> sizes = {1, 63, 64, 65, 4096, (256u << 10) - 8, 256u << 10, (256u << 10) + 8, 1u << 20, (3u << 20) + 40, 4u << 20};
    int bad = 0;
    for (int r = 0; r < rounds; r++) {
        for (size_t bytes : sizes) {
            for (size_t mis : {size_t(0), size_t(8), size_t(24)}) {
                std::vector<char> src(bytes + 64), dstv(bytes + 128);
                for (size_t i = 0; i < src.size(); i++) src[i] = (char)((i * 131 + seed + r) & 0xff);
                char *dst = dstv.data();
                dst += (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;  // the pinned mirrors are 64-byte aligned
                std::memset(dst, 0x5a, bytes + 32);
                b200e_host::CopyPool::get().copy(dst, src.data() + mis, bytes);
                if (std::memcmp(dst, src.data() + mis, bytes) != 0) bad++;
                for (int k = 0; k < 32; k++) if (dst[bytes + k] != 0x5a) bad++;  // nothing written past the end
            }
        }
    }
    return bad;
}

int main() {
    int bad1 = 0, bad2 = 0;
    std::thread a([&] { bad1 = run_caller(1, 3); });
    std::thread b([&] { bad2 = run_caller(2, 3); });
    a.join();
    b.join();
    std::this_thread::sleep_for(std::chrono::milliseconds(2));  // let the workers fall asleep on the condition variable
    const int bad3 = run_caller(3, 1);                            // ... and wake them again
    std::printf("copy pool: %d mismatches\n", bad1 + bad2 + bad3);
    return bad1 + bad2 + bad3 ? 1 : 0;
}
