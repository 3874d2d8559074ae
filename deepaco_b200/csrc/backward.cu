// Backward of ACO.sample()'s log-probabilities: C ABI (kernel in backward.cuh).
#include "backward.cuh"
#include "host_util.h"

using namespace deepaco;

extern "C" int deepaco_logp_backward(const float* pheromone_pow, const float* heuristic_pow, const int64_t* paths,
                                     const float* grad_log_probs, int n, int n_ants, int path_rows, const float* demand,
                                     float capacity, float* grad_heuristic, float* grad_pheromone, void* stream) {
    DACO_CHECK_ARG(pheromone_pow && heuristic_pow && paths && grad_log_probs && grad_heuristic, "deepaco_logp_backward: NULL argument");
    DACO_CHECK_ARG(n >= 2 && n <= 1024 && n_ants >= 1 && path_rows >= 2, "deepaco_logp_backward: bad sizes (n <= 1024)");
    BackwardParams p{pheromone_pow, heuristic_pow, paths, grad_log_probs, grad_heuristic, grad_pheromone, demand, capacity, n, n_ants, path_rows};
    logp_backward_kernel<<<(n_ants + 7) / 8, 256, 0, (cudaStream_t)stream>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
