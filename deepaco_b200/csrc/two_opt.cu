// K4 -- batched 2-opt and NLS local search: launchers and C ABI (kernels in two_opt.cuh).
#include "two_opt.cuh"
#include "host_util.h"

#include <stdlib.h>

using namespace deepaco;

static int launch_two_opt(const float* dist, const float* heu_dist, uint16_t* tours, int n, int A, int B, int mode, int maxt,
                          int T_nls, int T_p, float* costs_out, int32_t* passes_out, cudaStream_t st) {
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(dist && tours, "deepaco_two_opt: NULL argument");
    DACO_CHECK_ARG(n >= 4 && n <= 65535 && A >= 1 && B >= 1 && B <= 65535, "deepaco_two_opt: bad sizes (n >= 4)");
    DACO_CHECK_ARG(maxt >= 0 && T_nls >= 0 && T_p >= 0, "deepaco_two_opt: negative iteration count");
    int variant = two_opt_variant(n);
    if (const char* e = getenv("DEEPACO_2OPT_LEGACY")) { if (atoi(e) != 0) variant = 0; }
    int W = two_opt_warps(variant, n);
    if (variant == 0) { if (const char* e = getenv("DEEPACO_2OPT_WARPS")) { const int w = atoi(e); if (w == 4 || w == 8 || w == 12 || w == 16) W = w; } }
    const size_t smem = two_opt_smem_bytes(W, n);
    DACO_CHECK_ARG(smem <= (size_t)di->max_smem_optin - 1024, "deepaco_two_opt: n=%d does not fit shared memory", n);
    dim3 grid(A, B);
#define DACO_LAUNCH_2OPT(K)                                                                                                  \
    do {                                                                                                                     \
        DACO_CHECK_CUDA(cudaFuncSetAttribute(two_opt_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        two_opt_kernel<K><<<grid, W * 32, smem, st>>>(dist, heu_dist, tours, n, A, mode, maxt, T_nls, T_p, costs_out, passes_out); \
    } while (0)
    switch (variant) {
        case 4: DACO_LAUNCH_2OPT(4); break;
        case 8: DACO_LAUNCH_2OPT(8); break;
        case 16: DACO_LAUNCH_2OPT(16); break;
        default: DACO_LAUNCH_2OPT(0); break;
    }
#undef DACO_LAUNCH_2OPT
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

extern "C" int deepaco_two_opt(const float* distances, uint16_t* tours, int n, int n_ants, int n_colonies,
                               int max_iterations, int32_t* passes_out, void* stream) {
    return launch_two_opt(distances, nullptr, tours, n, n_ants, n_colonies, 0, max_iterations, 0, 0, nullptr, passes_out,
                          (cudaStream_t)stream);
}

extern "C" int deepaco_tsp_nls(const float* distances, const float* heuristic_dist, uint16_t* tours, int n, int n_ants,
                               int n_colonies, int max_iterations, int T_nls, int T_p, float* costs_out,
                               int32_t* passes_out, void* stream) {
    DACO_CHECK_ARG(heuristic_dist != nullptr, "deepaco_tsp_nls: heuristic_dist is NULL");
    return launch_two_opt(distances, heuristic_dist, tours, n, n_ants, n_colonies, 1, max_iterations, T_nls, T_p, costs_out,
                          passes_out, (cudaStream_t)stream);
}

// tour layout conversion helpers for the Python boundary: paths int64 [B][n][A] <-> tours u16 [B][A][n]
namespace deepaco {
__global__ void paths_to_tours_kernel(const int64_t* __restrict__ paths, uint16_t* __restrict__ tours, int n, int A, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)n * A), r = i % ((size_t)n * A);
        const int a = (int)(r / n), k = (int)(r % n);
        tours[i] = (uint16_t)paths[(b * n + k) * A + a];
    }
}
__global__ void tours_to_paths_kernel(const uint16_t* __restrict__ tours, int64_t* __restrict__ paths, int n, int A, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / ((size_t)n * A), r = i % ((size_t)n * A);
        const int k = (int)(r / A), a = (int)(r % A);
        paths[i] = (int64_t)tours[(b * A + a) * n + k];
    }
}
}  // namespace deepaco

extern "C" int deepaco_paths_to_tours(const int64_t* paths, uint16_t* tours, int n, int n_ants, int n_colonies, void* stream) {
    DACO_CHECK_ARG(paths && tours && n >= 1 && n_ants >= 1 && n_colonies >= 1, "deepaco_paths_to_tours: bad arguments");
    const size_t total = (size_t)n_colonies * n * n_ants;
    paths_to_tours_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(paths, tours, n, n_ants, total);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

extern "C" int deepaco_tours_to_paths(const uint16_t* tours, int64_t* paths, int n, int n_ants, int n_colonies, void* stream) {
    DACO_CHECK_ARG(paths && tours && n >= 1 && n_ants >= 1 && n_colonies >= 1, "deepaco_tours_to_paths: bad arguments");
    const size_t total = (size_t)n_colonies * n * n_ants;
    tours_to_paths_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(tours, paths, n, n_ants, total);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
