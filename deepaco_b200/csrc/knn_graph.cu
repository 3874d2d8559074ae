// Instance -> graph front end: C ABI (kernel in knn_graph.cuh).
#include "knn_graph.cuh"
#include "host_util.h"

using namespace deepaco;

extern "C" int deepaco_knn_graph(const float* coords, const float* distances_in, int n, int n_instances, int k, float diag,
                                 float* distances_out, int32_t* nbr_index, float* nbr_value, int64_t* edge_index, void* stream) {
    DACO_CHECK_ARG((coords != nullptr) != (distances_in != nullptr), "deepaco_knn_graph: pass coords or distances_in, not both");
    DACO_CHECK_ARG(n >= 1 && n <= 8192 && n_instances >= 1 && k >= 0 && k <= n, "deepaco_knn_graph: need 1 <= n <= 8192 and 0 <= k <= n");
    DACO_CHECK_ARG(k > 0 || distances_out, "deepaco_knn_graph: nothing to compute (k = 0 and no distances_out)");
    DACO_CHECK_ARG(k == 0 || nbr_index || nbr_value || edge_index, "deepaco_knn_graph: k > 0 needs an output for the neighbours");
    const long rows = (long)n_instances * n;
    int warps = 8;                                                   // one warp per row, its row of n floats in shared memory
    while (warps > 1 && (size_t)warps * n * sizeof(float) > 200 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * n * sizeof(float);
    DACO_CHECK_CUDA(cudaFuncSetAttribute(knn_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const KnnGraphParams p{coords, distances_in, distances_out, nbr_index, nbr_value, edge_index, n, n_instances, k, diag};
    knn_graph_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, smem, (cudaStream_t)stream>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
