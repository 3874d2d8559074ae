#include "host_util.h"

#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>

namespace deepaco {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const DeviceInfo* device_info() {
    static DeviceInfo cache[64];
    static bool have[64] = {};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        set_error("cudaGetDevice failed: no CUDA device available");
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev]) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
            set_error("cudaGetDeviceProperties failed");
            return nullptr;
        }
        cache[dev] = {dev, p.multiProcessorCount, p.maxThreadsPerMultiProcessor, (int)p.sharedMemPerBlockOptin, p.major};
        have[dev] = true;
    }
    return &cache[dev];
}

cudaError_t StreamScratch::alloc(size_t bytes, cudaStream_t stream) {
    static bool tuned[64] = {};
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> lk(mu);
        if (!tuned[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            tuned[dev] = true;
        }
    }
    st = stream;
    return cudaMallocAsync(&ptr, bytes, stream);
}

StreamScratch::~StreamScratch() {
    if (ptr) cudaFreeAsync(ptr, st);
}

static int last_pow2(int64_t v) {
    int p = 1;
    while ((int64_t)p * 2 <= v) p *= 2;
    return p;
}

// ATen/native/cuda/Reduce.cuh: setReduceConfig + ReduceConfig::set_block_dimension, specialised to a
// contiguous fp32 [n_rows][row_len] input reduced over its last dimension.
SumPlan aten_sum_plan(int row_len, int n_rows) {
    SumPlan sp{};
    int64_t dim0 = row_len, dim1 = n_rows;
    sp.vectorized = row_len >= 128;
    if (sp.vectorized) dim0 /= 4;
    const int max_threads = 512;
    const int d0 = dim0 < max_threads ? last_pow2(dim0) : max_threads;
    const int d1 = dim1 < max_threads ? last_pow2(dim1) : max_threads;
    int bw = std::min(d0, 32);
    const int bh = std::min(d1, max_threads / bw);
    bw = std::min(d0, max_threads / bh);
    sp.block_width = bw;
    const int vpt = (row_len + bw - 1) / bw;
    const bool split_warps = vpt >= std::min(bh * 16, 256);
    sp.exact = !split_warps && bw <= 32;
    return sp;
}

// ATen/native/cuda/DistributionTemplates.h: calc_execution_policy (block 256, unroll 4)
DrawPlan torch_draw_plan(int64_t numel, const DeviceInfo& di) {
    DrawPlan dp{};
    if (numel <= 0) {
        dp.threads = 256;
        dp.single = 1;
        dp.increment = 0;
        return dp;
    }
    const uint64_t block = 256;
    uint64_t grid = ((uint64_t)numel + block - 1) / block;
    const uint64_t cap = (uint64_t)di.sm_count * (uint64_t)(di.max_threads_per_sm / (int)block);
    grid = std::min(grid, cap);
    dp.threads = (uint32_t)(grid * block);
    dp.single = (uint64_t)numel <= grid * block;
    dp.increment = (((uint64_t)numel - 1) / (block * grid * 4) + 1) * 4;
    return dp;
}

}  // namespace deepaco

extern "C" {

const char* deepaco_last_error(void) { return deepaco::g_err; }

int deepaco_version(void) { return 100; }

long long deepaco_kernel_launches(void) { return deepaco::g_launches.load(); }

int deepaco_torch_draw_geometry(int64_t numel, uint32_t* threads_out, uint64_t* offset_increment_out) {
    const deepaco::DeviceInfo* di = deepaco::device_info();
    if (!di) return DEEPACO_ENODEV;
    deepaco::DrawPlan dp = deepaco::torch_draw_plan(numel, *di);
    if (threads_out) *threads_out = dp.threads;
    if (offset_increment_out) *offset_increment_out = dp.increment;
    return DEEPACO_OK;
}

int deepaco_aten_sum_plan(int row_len, int n_rows, int* block_width_out, int* vectorized_out, int* exact_out) {
    DACO_CHECK_ARG(row_len > 0 && n_rows > 0, "deepaco_aten_sum_plan: row_len and n_rows must be positive");
    deepaco::SumPlan sp = deepaco::aten_sum_plan(row_len, n_rows);
    if (block_width_out) *block_width_out = sp.block_width;
    if (vectorized_out) *vectorized_out = sp.vectorized;
    if (exact_out) *exact_out = sp.exact;
    return DEEPACO_OK;
}

}  // extern "C"
