// Backward of ACO.sample()'s log-probabilities: C ABI (kernels in backward.cuh).
#include "backward.cuh"
#include "host_util.h"

using namespace deepaco;

extern "C" int deepaco_logp_backward(const float* pheromone_pow, const float* heuristic_pow, const int64_t* paths,
                                     const float* grad_log_probs, int n, int n_ants, int path_rows, const float* demand,
                                     float capacity, float* grad_heuristic, float* grad_pheromone, void* stream) {
    DACO_CHECK_ARG(pheromone_pow && heuristic_pow && paths && grad_log_probs && grad_heuristic, "deepaco_logp_backward: NULL argument");
    DACO_CHECK_ARG(n >= 2 && n <= 1024 && n_ants >= 1 && path_rows >= 2 && path_rows < 65535,
                   "deepaco_logp_backward: bad sizes (n <= 1024, path_rows < 65535)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t steps = (size_t)(path_rows - 1) * n_ants;
    // stream-ordered scratch: coef | gact | rem (f32 each) | when u16 [A][n] | depot_steps u16 [A][rows] | depot_count | dok
    const size_t off_when = 3 * steps * sizeof(float);
    const size_t off_depot = off_when + (((size_t)n_ants * n * 2 + 15) & ~(size_t)15);
    const size_t off_cnt = off_depot + (((size_t)n_ants * path_rows * 2 + 15) & ~(size_t)15);
    const size_t off_dok = off_cnt + (((size_t)n_ants * 4 + 15) & ~(size_t)15);
    const size_t total = off_dok + steps;
    StreamScratch ws;
    DACO_CHECK_CUDA(ws.alloc(total, st));
    char* base = static_cast<char*>(ws.ptr);
    BackwardParams p{pheromone_pow, heuristic_pow, paths, grad_log_probs, grad_heuristic, grad_pheromone, demand, capacity, n, n_ants,
                     path_rows, reinterpret_cast<float*>(base), reinterpret_cast<float*>(base) + steps,
                     reinterpret_cast<float*>(base) + 2 * steps, reinterpret_cast<uint8_t*>(base + off_dok),
                     reinterpret_cast<uint16_t*>(base + off_when), reinterpret_cast<uint16_t*>(base + off_depot),
                     reinterpret_cast<int32_t*>(base + off_cnt)};
    DACO_CHECK_CUDA(cudaMemsetAsync(p.when, 0xff, (size_t)n_ants * n * 2, st));
    logp_backward_prepare_kernel<<<(n_ants + 7) / 8, 256, 0, st>>>(p);
    DACO_CHECK_LAUNCH();
    logp_backward_rows_kernel<<<n, 128, 0, st>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
