// Host stand-in for the handful of CUDA constructs csrc/gnn_train.cuh uses: one OS thread per CUDA thread, pthread
// barriers for __syncthreads / barrier.cluster, a heap block per CTA for dynamic shared memory.
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
};
inline thread_local Ctx ctx;

// run `kernel(params)` for n_clusters clusters of `ncta` CTAs x `nth` threads (clusters one after another)
template <class Kernel, class Params>
void launch(Kernel kernel, const Params& params, int n_clusters, int ncta, int nth, size_t smem_bytes) {
    for (int cl = 0; cl < n_clusters; ++cl) {
        std::vector<unsigned char*> smem(ncta);
        std::vector<pthread_barrier_t> bars(ncta);
        pthread_barrier_t cluster_bar;
        pthread_barrier_init(&cluster_bar, nullptr, (unsigned)(ncta * nth));
        for (int r = 0; r < ncta; ++r) {
            smem[r] = static_cast<unsigned char*>(aligned_alloc(128, (smem_bytes + 127) / 128 * 128));
            memset(smem[r], 0xff, smem_bytes);       // NaN pattern: reads of never-written shared memory show up
            pthread_barrier_init(&bars[r], nullptr, (unsigned)nth);
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
                    kernel(params);
                });
        for (auto& th : threads) th.join();
        for (int r = 0; r < ncta; ++r) {
            free(smem[r]);
            pthread_barrier_destroy(&bars[r]);
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
