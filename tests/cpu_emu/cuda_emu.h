// Host stand-in for the handful of CUDA constructs csrc/gnn_train.cuh and csrc/two_opt.cuh use: one OS thread per CUDA
// thread, pthread barriers for __syncthreads / __syncwarp / barrier.cluster, warp shuffles through a per-warp exchange
// buffer, a heap block per CTA for dynamic shared memory.
//
// TEST INFRASTRUCTURE ONLY.  It lets the *same kernel source* that nvcc compiles for sm_100a run in the CPU test
// suite (tests/test_gnn_train_emu.py), where its indexing, phase ordering and -- under ThreadSanitizer -- its barrier
// placement are checked against torch autograd without a GPU.  Nothing under deepaco_b200/ includes or loads this.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <thread>
#include <vector>

#define DEEPACO_CPU_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct alignas(16) float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct alignas(8) float2 {
    float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

namespace emu {
struct Dim3 {
    unsigned x, y, z;
};
struct Ctx {
    Dim3 tid, bid, bdim;
    unsigned rank, ncta;
    unsigned char* smem;
    pthread_barrier_t* cta_bar;
    pthread_barrier_t* cluster_bar;
    pthread_barrier_t* warp_bars;   // one per warp of this CTA
    uint32_t* warp_slots;           // [warps][32] shuffle exchange
    int* cta_flag;                  // __syncthreads_or accumulator
};
inline thread_local Ctx ctx;

// run `kernel(params)` for n_clusters clusters of `ncta` CTAs x `nth` threads (clusters one after another).
// grid_x > 0: blockIdx = (linear % grid_x, linear / grid_x) for kernels launched on a 2-D grid (ncta must be 1).
template <class Kernel, class Params>
void launch(Kernel kernel, const Params& params, int n_clusters, int ncta, int nth, size_t smem_bytes, int grid_x = 0) {
    for (int cl = 0; cl < n_clusters; ++cl) {
        std::vector<unsigned char*> smem(ncta);
        std::vector<pthread_barrier_t> bars(ncta);
        const int nwarps = (nth + 31) / 32;
        std::vector<pthread_barrier_t> warp_bars((size_t)ncta * nwarps);
        std::vector<uint32_t> warp_slots((size_t)ncta * nwarps * 32, 0u);
        std::vector<int> cta_flags(ncta, 0);
        pthread_barrier_t cluster_bar;
        pthread_barrier_init(&cluster_bar, nullptr, (unsigned)(ncta * nth));
        for (int r = 0; r < ncta; ++r) {
            smem[r] = static_cast<unsigned char*>(aligned_alloc(128, (smem_bytes + 127) / 128 * 128));
            memset(smem[r], 0xff, smem_bytes);       // NaN pattern: reads of never-written shared memory show up
            pthread_barrier_init(&bars[r], nullptr, (unsigned)nth);
            for (int w = 0; w < nwarps; ++w)
                pthread_barrier_init(&warp_bars[(size_t)r * nwarps + w], nullptr, (unsigned)(nth - w * 32 < 32 ? nth - w * 32 : 32));
        }
        std::vector<std::thread> threads;
        threads.reserve((size_t)ncta * nth);
        for (int r = 0; r < ncta; ++r)
            for (int t = 0; t < nth; ++t)
                threads.emplace_back([&, r, t]() {
                    ctx.tid = {(unsigned)t, 0, 0};
                    ctx.bid = {(unsigned)(cl * ncta + r), 0, 0};
                    if (grid_x > 0) ctx.bid = {(unsigned)(cl % grid_x), (unsigned)(cl / grid_x), 0};
                    ctx.bdim = {(unsigned)nth, 1, 1};
                    ctx.rank = (unsigned)r;
                    ctx.ncta = (unsigned)ncta;
                    ctx.smem = smem[r];
                    ctx.cta_bar = &bars[r];
                    ctx.cluster_bar = &cluster_bar;
                    ctx.warp_bars = &warp_bars[(size_t)r * nwarps];
                    ctx.warp_slots = &warp_slots[(size_t)r * nwarps * 32];
                    ctx.cta_flag = &cta_flags[r];
                    kernel(params);
                });
        for (auto& th : threads) th.join();
        for (int r = 0; r < ncta; ++r) {
            free(smem[r]);
            pthread_barrier_destroy(&bars[r]);
            for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&warp_bars[(size_t)r * nwarps + w]);
        }
        pthread_barrier_destroy(&cluster_bar);
    }
}
}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define DACO_DYN_SMEM(name) unsigned char* name = emu::ctx.smem
static inline void __syncthreads() { pthread_barrier_wait(emu::ctx.cta_bar); }

// ---- warp level (every lane of the warp must take part, as with a full mask on the device) ----
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::ctx.warp_bars[emu::ctx.tid.x >> 5]); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t* slots = emu::ctx.warp_slots + (emu::ctx.tid.x >> 5) * 32;
    uint32_t bits;
    memcpy(&bits, &v, 4);
    slots[emu::ctx.tid.x & 31] = bits;
    __syncwarp();
    bits = slots[src_lane & 31];
    __syncwarp();                      // nobody overwrites a slot before every lane has read
    T r;
    memcpy(&r, &bits, 4);
    return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, (int)(emu::ctx.tid.x & 31) ^ lane_mask); }
static inline int __syncthreads_or(int predicate) {
    if (predicate) __atomic_fetch_or(emu::ctx.cta_flag, 1, __ATOMIC_ACQ_REL);
    __syncthreads();
    const int r = __atomic_load_n(emu::ctx.cta_flag, __ATOMIC_ACQUIRE);
    __syncthreads();
    if (emu::ctx.tid.x == 0) __atomic_store_n(emu::ctx.cta_flag, 0, __ATOMIC_RELEASE);
    __syncthreads();
    return r;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline float atomicAdd(float* p, float v) {
    uint32_t old = __atomic_load_n(reinterpret_cast<uint32_t*>(p), __ATOMIC_ACQUIRE), want;
    float f;
    do {
        memcpy(&f, &old, 4);
        f += v;
        memcpy(&want, &f, 4);
    } while (!__atomic_compare_exchange_n(reinterpret_cast<uint32_t*>(p), &old, want, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE));
    return f - v;
}
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_ACQ_REL); }
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {}
    return old;
}
// warp votes / reductions: every lane contributes through the exchange buffer, then reads all 32 slots
template <class F>
static inline uint32_t emu_warp_fold(uint32_t mine, F f) {
    uint32_t* slots = emu::ctx.warp_slots + (emu::ctx.tid.x >> 5) * 32;
    slots[emu::ctx.tid.x & 31] = mine;
    __syncwarp();
    uint32_t acc = slots[0];
    for (int l = 1; l < 32; ++l) acc = f(acc, slots[l]);
    __syncwarp();
    return acc;
}
static inline uint32_t __ballot_sync(unsigned, int pred) {
    return emu_warp_fold(pred ? (1u << (emu::ctx.tid.x & 31)) : 0u, [](uint32_t a, uint32_t b) { return a | b; });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline uint32_t __reduce_max_sync(unsigned, uint32_t v) { return emu_warp_fold(v, [](uint32_t a, uint32_t b) { return a > b ? a : b; }); }
static inline uint32_t __reduce_min_sync(unsigned, uint32_t v) { return emu_warp_fold(v, [](uint32_t a, uint32_t b) { return a < b ? a : b; }); }
static inline uint32_t __reduce_or_sync(unsigned, uint32_t v) { return emu_warp_fold(v, [](uint32_t a, uint32_t b) { return a | b; }); }
template <class T>
static inline T __shfl_up_sync(unsigned m, T v, unsigned delta) {     // lanes below `delta` keep their own value
    const int lane = (int)(emu::ctx.tid.x & 31);
    const T got = __shfl_sync(m, v, lane >= (int)delta ? lane - (int)delta : lane);
    return lane >= (int)delta ? got : v;
}
static inline uint32_t __match_any_sync(unsigned, int v) {           // mask of the lanes holding the same value
    uint32_t* slots = emu::ctx.warp_slots + (emu::ctx.tid.x >> 5) * 32;
    slots[emu::ctx.tid.x & 31] = (uint32_t)v;
    __syncwarp();
    uint32_t m = 0;
    for (int l = 0; l < 32; ++l) m |= (slots[l] == (uint32_t)v) ? (1u << l) : 0u;
    __syncwarp();
    return m;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __float_as_uint(float f) { uint32_t v; memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(uint32_t v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline void* __cvta_shared_to_generic(uint32_t addr) { return emu::ctx.smem + addr; }
#define __grid_constant__

// ---- mbarrier + bulk-copy stand-in: the 64-bit barrier word is (pending tx bytes << 32 | completed phases); a copy
// happens at issue time and completes the phase when it brings the pending bytes to zero (one issuing thread per barrier)
static inline void emu_mbar_init(uint64_t* bar) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_RELEASE); }
static inline void emu_mbar_expect_tx(uint64_t* bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t)bytes << 32, __ATOMIC_ACQ_REL); }
static inline void emu_bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    memcpy(dst, src, bytes);
    const uint64_t now = __atomic_sub_fetch(bar, (uint64_t)bytes << 32, __ATOMIC_ACQ_REL);
    if ((now >> 32) == 0) __atomic_fetch_add(bar, (uint64_t)1, __ATOMIC_ACQ_REL);
}
static inline void emu_mbar_wait(uint64_t* bar, uint32_t parity) {   // done when the current phase parity != `parity`
    while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1u) == parity) sched_yield();
}
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __fadd_rn(float a, float b) { return a + b; }      // built with -ffp-contract=off
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }

namespace deepaco {
namespace gnnt {
static inline unsigned cta_rank() { return emu::ctx.rank; }
static inline unsigned cta_count() { return emu::ctx.ncta; }
static inline void cluster_barrier() { pthread_barrier_wait(emu::ctx.cluster_bar); }
static inline void counter_barrier(unsigned* ctr, unsigned target) {   // same protocol as the device version
    __syncthreads();
    if (threadIdx.x == 0) {
        __atomic_fetch_add(ctr, 1u, __ATOMIC_ACQ_REL);
        while (__atomic_load_n(ctr, __ATOMIC_ACQUIRE) < target) sched_yield();
    }
    __syncthreads();
}
static inline float ld_cg(const float* p) { return *p; }
static inline float4 ld_cg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
}  // namespace gnnt
}  // namespace deepaco
