// cuda_emul.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the kernels of
// scimlexpectations.jl_b200/csrc/kernels/expect_kernels.cuh on the CPU box, one OS thread per CUDA thread, one block
// at a time.  It exists so that the "not gpu" suite exercises the SAME kernel source the B200 runs (start of a
// trajectory, step functions, refill logic, work distribution, block reduction, last-block fold) against the oracle;
// it is never part of the product (the library has no CPU path) and nothing outside tests/ includes it.
//
// Warp intrinsics are collective: every lane of the warp deposits its operand, a 32-party barrier separates the
// deposit from the reads and a second one the reads from the next deposit.  The kernels only call them in
// warp-uniform control flow with the full mask, which is what this model requires.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __grid_constant__

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;
struct double2 { double x, y; };

struct EmuWarp {
    std::barrier<> bar{32};
    unsigned long long slot[32];
};
inline thread_local EmuWarp *emu_warp = nullptr;
inline thread_local int emu_lane = 0;
inline std::barrier<> *emu_block_bar = nullptr;

template <class T>
inline unsigned long long emu_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T emu_from_bits(unsigned long long b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

// deposit v, then read the value lane `src` deposited (own value when src is outside the warp)
template <class T>
inline T emu_exchange(T v, int src) {
    emu_warp->slot[emu_lane] = emu_bits(v);
    emu_warp->bar.arrive_and_wait();
    const T r = (src >= 0 && src < 32) ? emu_from_bits<T>(emu_warp->slot[src]) : v;
    emu_warp->bar.arrive_and_wait();
    return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int off) { return emu_exchange(v, emu_lane + off); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, emu_lane ^ m); }
inline unsigned __ballot_sync(unsigned, bool pred) {
    emu_warp->slot[emu_lane] = pred ? 1ull : 0ull;
    emu_warp->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= (unsigned)(emu_warp->slot[l] & 1ull) << l;
    emu_warp->bar.arrive_and_wait();
    return m;
}
inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    emu_warp->slot[emu_lane] = v;
    emu_warp->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m = emu_warp->slot[l] > m ? (unsigned)emu_warp->slot[l] : m;
    emu_warp->bar.arrive_and_wait();
    return m;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    emu_warp->slot[emu_lane] = v;
    emu_warp->bar.arrive_and_wait();
    unsigned s = 0;
    for (int l = 0; l < 32; l++) s += (unsigned)emu_warp->slot[l];
    emu_warp->bar.arrive_and_wait();
    return s;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp->bar.arrive_and_wait(); }
#ifdef EMU_DROP_BLOCK_BARRIER  // negative control of the ThreadSanitizer run: without the barrier it must report the race
inline void __syncthreads() {}
#else
inline void __syncthreads() { emu_block_bar->arrive_and_wait(); }
#endif
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __float_as_int(float x) { int b; std::memcpy(&b, &x, 4); return b; }
inline float __int_as_float(int b) { float x; std::memcpy(&x, &b, 4); return x; }
using std::max;
using std::min;
inline double __longlong_as_double(long long b) { double x; std::memcpy(&x, &b, 8); return x; }
inline long long __double_as_longlong(double x) { long long b; std::memcpy(&b, &x, 8); return b; }

// run kernel(P) for every block of the grid, blocks one after the other, `block` OS threads each
template <class Kernel, class Params>
void emu_launch(Kernel kernel, const Params &P, unsigned grid, unsigned block) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; b++) {
        std::barrier<> bar(block);
        emu_block_bar = &bar;
        std::vector<std::unique_ptr<EmuWarp>> warps;
        for (unsigned w = 0; w < block / 32; w++) warps.emplace_back(new EmuWarp());
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t] {
                threadIdx.x = t;
                blockIdx.x = b;
                emu_warp = warps[t / 32].get();
                emu_lane = (int)(t % 32);
                kernel(P);
            });
        for (auto &x : th) x.join();
    }
}
