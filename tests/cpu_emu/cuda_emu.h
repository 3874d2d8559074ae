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
#define DACO_NOINLINE __attribute__((noinline))   /* not `__noinline__`: libstdc++ spells its own attributes that way */

struct alignas(16) float4 {
    float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

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

// run `kernel(params)` for n_clusters clusters of `ncta` CTAs x `nth` threads (clusters one after another)
template <class Kernel, class Params>
void launch(Kernel kernel, const Params& params, int n_clusters, int ncta, int nth, size_t smem_bytes) {
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
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __fadd_rn(float a, float b) { return a + b; }      // built with -ffp-contract=off
static inline float __fsub_rn(float a, float b) { return a - b; }

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
