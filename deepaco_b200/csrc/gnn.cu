// K3 -- heuristic network forward, eval mode, one CTA per instance: C ABI (kernel in gnn.cuh).
#include "gnn.cuh"
#include "host_util.h"

using namespace deepaco;

extern "C" int64_t deepaco_gnn_weight_count(int feats) { return (int64_t)U * feats + 3 * U + (int64_t)kDepth * kLayerFloats + 2 * (U * U + U) + U + 1; }

extern "C" int deepaco_gnn_forward(const float* x, const int32_t* row_ptr, const int32_t* dst_sorted, const float* attr_sorted,
                                   const int32_t* order, const float* weights, int n_nodes, int n_edges, int feats,
                                   int n_instances, float* node_ws, float* edge_ws, float* heu_out, float* dense_out,
                                   float dense_eps, const int32_t* src_sorted, void* stream) {
    DACO_CHECK_ARG(x && row_ptr && dst_sorted && attr_sorted && order && weights && node_ws && edge_ws && (heu_out || dense_out),
                   "deepaco_gnn_forward: NULL argument");
    DACO_CHECK_ARG(n_nodes >= 1 && n_edges >= 1 && feats >= 1 && feats <= 8 && n_instances >= 1, "deepaco_gnn_forward: bad sizes");
    GnnParams p{x, row_ptr, dst_sorted, attr_sorted, order, weights, node_ws, edge_ws, heu_out, dense_out, dense_eps, n_nodes, n_edges, feats,
                src_sorted};
    const int threads = 512;
    const size_t smem = gnn_forward_smem(feats, threads);
    DACO_CHECK_CUDA(cudaFuncSetAttribute(gnn_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gnn_forward_kernel<<<n_instances, threads, smem, (cudaStream_t)stream>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
