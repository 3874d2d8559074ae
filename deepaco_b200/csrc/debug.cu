// Probe entry points: regenerate torch's CUDA `exponential_` / `randint` streams and ATen's row sum so
// the tests can confirm, on the box, the parity assumptions of SURVEY.md Appendix A before relying on them.
#include "common.cuh"
#include "host_util.h"

namespace deepaco {

__global__ void debug_exponential_kernel(uint64_t seed, uint64_t offset, int64_t numel, DrawGeom g, float* out) {
    for (int64_t li = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; li < numel; li += (int64_t)gridDim.x * blockDim.x)
        out[li] = exp1_from_word(torch_philox_word(seed, offset, (uint64_t)li, g));
}

__global__ void debug_randint_kernel(uint64_t seed, uint64_t offset, int64_t numel, uint32_t high, DrawGeom g, int64_t* out) {
    for (int64_t li = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; li < numel; li += (int64_t)gridDim.x * blockDim.x)
        out[li] = (int64_t)(torch_philox_word(seed, offset, (uint64_t)li, g) % high);
}

__global__ void debug_row_sum_kernel(const float* x, int rows, int len, int lbw, int vec, float* out) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* row = x + (size_t)r * len;
    const float s = aten_row_sum_fn([&](int k) { return row[k]; }, len, lbw, vec != 0, threadIdx.x & 31,
                                    vec ? (int)(((unsigned)r * (unsigned)len) & 3u) : 0);
    if ((threadIdx.x & 31) == 0) out[r] = s;
}

}  // namespace deepaco

using namespace deepaco;

extern "C" int deepaco_debug_exponential(uint64_t seed, uint64_t offset, int64_t numel, float* out, void* stream) {
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(numel > 0 && out, "deepaco_debug_exponential: bad arguments");
    const DrawPlan dp = torch_draw_plan(numel, *di);
    debug_exponential_kernel<<<(unsigned)std::min<int64_t>((numel + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(
        seed, offset, numel, DrawGeom{dp.threads, dp.single}, out);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

extern "C" int deepaco_debug_randint(uint64_t seed, uint64_t offset, int64_t numel, int64_t high, int64_t* out, void* stream) {
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(numel > 0 && out && high > 0 && high < (1ll << 28), "deepaco_debug_randint: bad arguments");
    const DrawPlan dp = torch_draw_plan(numel, *di);
    debug_randint_kernel<<<(unsigned)std::min<int64_t>((numel + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(
        seed, offset, numel, (uint32_t)high, DrawGeom{dp.threads, dp.single}, out);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

extern "C" int deepaco_debug_row_sum(const float* x, int n_rows, int row_len, float* out, void* stream) {
    DACO_CHECK_ARG(x && out && n_rows > 0 && row_len > 0, "deepaco_debug_row_sum: bad arguments");
    const SumPlan sp = aten_sum_plan(row_len, n_rows);
    int bw = sp.block_width > 32 ? 32 : sp.block_width, lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    debug_row_sum_kernel<<<(n_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, n_rows, row_len, lbw, sp.vectorized, out);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

namespace deepaco {
__global__ void exp_guard_probe_kernel(unsigned long long* bad, uint32_t base, uint32_t stride) {
    const uint32_t x = base + (blockIdx.x * blockDim.x + threadIdx.x) * stride;
    if (__float_as_uint(exp1_from_word(x)) != __float_as_uint(exp1_from_word_guarded(x))) atomicAdd(bad, 1ull);
}
}  // namespace deepaco

// Counts the 32-bit Philox words for which the shortened Exp(1) transform (common.cuh exp1_from_word) differs from the
// literal ATen form: exhaustive over the top 2^20 words (the only region where they can differ) + every 256th word.
extern "C" int deepaco_debug_exp_guard(uint64_t* mismatches, void* stream) {
    DACO_CHECK_ARG(mismatches, "deepaco_debug_exp_guard: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    DACO_CHECK_CUDA(cudaMemsetAsync(mismatches, 0, sizeof(uint64_t), st));
    deepaco::exp_guard_probe_kernel<<<(1 << 20) / 256, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(mismatches), 0xFFF00000u, 1u);
    DACO_CHECK_LAUNCH();
    deepaco::exp_guard_probe_kernel<<<(1 << 24) / 256, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(mismatches), 0u, 256u);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
