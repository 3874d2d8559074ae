// host-side helpers shared by the C-ABI translation units
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deepaco_b200.h"

namespace deepaco {

void set_error(const char* fmt, ...);

struct DeviceInfo {
    int device;
    int sm_count;
    int max_threads_per_sm;
    int max_smem_optin;
    int cc_major;
};
// properties of the CURRENT device (cached per device ordinal); returns nullptr on failure
const DeviceInfo* device_info();

// ATen sum plan for sum over the last dim of a contiguous [n_rows][row_len] fp32 tensor
struct SumPlan {
    int block_width;   // lanes cooperating on one row (power of two)
    int vectorized;    // ATen "vectorize along input" path (row_len >= 128)
    int exact;         // 1 if our kernels reproduce this order exactly
};
SumPlan aten_sum_plan(int row_len, int n_rows);

struct DrawPlan {
    uint32_t threads;
    uint32_t single;
    uint64_t increment;
};
DrawPlan torch_draw_plan(int64_t numel, const DeviceInfo& di);

// Stream-ordered scratch memory (cudaMallocAsync on the current device's default pool; the pool's release threshold is
// raised once per device so freed blocks stay cached).  The free is queued on the same stream behind the consumer, so
// the buffer is private to (device, stream, call): no process-global workspace.
struct StreamScratch {
    void* ptr = nullptr;
    cudaStream_t st = nullptr;
    StreamScratch() = default;
    StreamScratch(const StreamScratch&) = delete;
    StreamScratch& operator=(const StreamScratch&) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t stream);
    ~StreamScratch();
};

void count_launch();   // every kernel launch of the library passes through DACO_CHECK_LAUNCH

// (count_launch is declared before the macros that use it)
#define DACO_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            deepaco::set_error(__VA_ARGS__); \
            return DEEPACO_EINVAL;           \
        }                                    \
    } while (0)

#define DACO_CHECK_CUDA(expr)                                                                        \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            deepaco::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DEEPACO_ECUDA;                                                                    \
        }                                                                                            \
    } while (0)

#define DACO_CHECK_LAUNCH()                   \
    do {                                      \
        deepaco::count_launch();              \
        DACO_CHECK_CUDA(cudaGetLastError());  \
    } while (0)

}  // namespace deepaco
