// simt.h -- TEST-ONLY thread-per-lane emulation of the handful of CUDA warp intrinsics the closed-loop
// kernel uses, so that csrc/mpc_sim.cuh (the exact kernel source) can be executed on a CPU by the
// not-gpu test-suite: one std::thread per lane, a std::barrier per warp, shuffles through a shared slot
// array.  Slow (a barrier per intrinsic) but faithful: per-lane registers, divergence-free control flow
// and every __syncwarp()/shuffle of the real kernel are exercised.  Never part of the product.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define MPC_SIMT_EMULATION 1

struct double2 { double x, y; };
struct SimtDim3 { unsigned x = 0, y = 0, z = 0; };
struct SimtWarp {
    std::barrier<> bar{32};
    double xd[32];
    long long xi[32];
    unsigned vote[32];
    // sub-warp groups (mpc_nmpc_group.cuh): a barrier per lane mask, created on first use, so that the groups of a warp may diverge
    std::mutex mu;
    std::map<unsigned, std::unique_ptr<std::barrier<>>> group_bars;
    std::barrier<> &bar_of(unsigned mask) {
        if (mask == 0xffffffffu) return bar;
        std::lock_guard<std::mutex> lock(mu);
        auto &b = group_bars[mask];
        if (!b) b = std::make_unique<std::barrier<>>((std::ptrdiff_t)__builtin_popcount(mask));
        return *b;
    }
};
static thread_local SimtDim3 threadIdx, blockIdx, blockDim;
static thread_local SimtWarp *simt_warp = nullptr;
static thread_local std::barrier<> *simt_block = nullptr;
static inline void __syncthreads() { simt_block->arrive_and_wait(); }
static inline long long clock64() { return 0; }

static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt_warp->bar_of(mask).arrive_and_wait(); }

template <typename T>
static inline T simt_xchg(T v, int src, unsigned mask = 0xffffffffu) {
    static_assert(sizeof(T) <= 8, "");
    long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    std::barrier<> &b = simt_warp->bar_of(mask);
    simt_warp->xi[threadIdx.x & 31] = raw;
    b.arrive_and_wait();
    long long got = simt_warp->xi[src & 31];
    b.arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = threadIdx.x & 31;
    const int base = lane & ~(width - 1);
    return simt_xchg(v, base + (src & (width - 1)), mask);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned member_mask, T v, int mask, int width = 32) {
    const int lane = threadIdx.x & 31;
    (void)width;
    return simt_xchg(v, lane ^ mask, member_mask);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = threadIdx.x & 31;
    const int pos = lane & (width - 1);
    const int src = (pos >= (int)delta) ? lane - (int)delta : lane;  // CUDA: out-of-segment lanes keep their own value
    return simt_xchg(v, src);
}
static inline int __any_sync(unsigned, int pred) {
    simt_warp->vote[threadIdx.x & 31] = pred ? 1u : 0u;
    simt_warp->bar.arrive_and_wait();
    int r = 0;
    for (int i = 0; i < 32; ++i) r |= (int)simt_warp->vote[i];
    simt_warp->bar.arrive_and_wait();
    return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    std::barrier<> &b = simt_warp->bar_of(mask);
    simt_warp->vote[threadIdx.x & 31] = pred ? 1u : 0u;
    b.arrive_and_wait();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i) & 1u) r |= simt_warp->vote[i] << i;
    b.arrive_and_wait();
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    simt_warp->vote[threadIdx.x & 31] = v;
    simt_warp->bar.arrive_and_wait();
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; ++i) r = simt_warp->vote[i] < r ? simt_warp->vote[i] : r;
    simt_warp->bar.arrive_and_wait();
    return r;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    simt_warp->vote[threadIdx.x & 31] = v;
    simt_warp->bar.arrive_and_wait();
    unsigned r = 0u;
    for (int i = 0; i < 32; ++i) r |= simt_warp->vote[i];
    simt_warp->bar.arrive_and_wait();
    return r;
}
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __reduce_xor_sync(unsigned, unsigned v) {
    simt_warp->vote[threadIdx.x & 31] = v;
    simt_warp->bar.arrive_and_wait();
    unsigned r = 0u;
    for (int i = 0; i < 32; ++i) r ^= simt_warp->vote[i];
    simt_warp->bar.arrive_and_wait();
    return r;
}
static inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; std::memcpy(&r, &l, 8); return r; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline double __ldg(const double *p) { return *p; }
static inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// run fn() as one warp of 32 lanes
static inline void simt_run_warp(const std::function<void()> &fn, unsigned block = 0) {
    SimtWarp w;
    std::vector<std::thread> th;
    for (unsigned l = 0; l < 32; ++l)
        th.emplace_back([&, l]() {
            threadIdx.x = l;
            blockIdx.x = block;
            simt_warp = &w;
            fn();
        });
    for (auto &t : th) t.join();
}

// run fn() as one CTA of `nthreads` threads (a multiple of 32): __syncthreads across the block, warp intrinsics inside each
// group of 32 consecutive threads
static inline void simt_run_block(const std::function<void()> &fn, unsigned nthreads) {
    std::barrier<> bar((std::ptrdiff_t)nthreads);
    std::vector<SimtWarp> warps((nthreads + 31) / 32);
    std::vector<std::thread> th;
    for (unsigned l = 0; l < nthreads; ++l)
        th.emplace_back([&, l]() {
            threadIdx.x = l;
            blockDim.x = nthreads;
            blockIdx.x = 0;
            simt_block = &bar;
            simt_warp = &warps[l / 32];
            fn();
        });
    for (auto &t : th) t.join();
}
