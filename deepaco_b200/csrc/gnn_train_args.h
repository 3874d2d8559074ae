// deepaco_gnn_train_args (C ABI) -> kernel parameter block, with the argument checks.  Plain C++ (no CUDA) so the
// host harness in tests/cpu_emu/ validates exactly what the library validates.
#pragma once
#include "../../include/deepaco_b200.h"

namespace deepaco {
namespace gnnt {

enum GroupCall { kTrainForward, kTrainBackward, kEvalForward };

// returns NULL on success, otherwise the reason
inline const char* gnn_train_params(const deepaco_gnn_train_args* a, GroupCall call, TrainParams& p) {
    if (!a) return "NULL args";
    if (!(a->x && a->row_ptr && a->src_sorted && a->dst_sorted && a->attr_sorted && a->order && a->weights && a->xs && a->ws &&
          a->node_ws && a->sync_ws))
        return "NULL argument";
    if (call != kEvalForward && !(a->zv && a->ze && a->stats && a->red)) return "NULL argument";
    if (call != kTrainBackward && !a->heu_out) return "NULL heu_out";
    if (call == kTrainBackward && !(a->grad_heu && a->grad_weights && a->edge_ws && a->col_ptr && a->in_edges))
        return "NULL backward argument";
    if (!(a->n_nodes >= 1 && a->n_edges >= 1 && a->feats >= 1 && a->feats <= 8 && a->n_instances >= 1)) return "bad sizes";
    const int c = a->ctas_per_instance;
    if (!(c == 1 || c == 2 || c == 4 || c == 8 || c == 16 || c == 32 || c == 64)) return "ctas_per_instance must be 1, 2, 4, 8, 16, 32 or 64";
    if (!(a->bn_eps > 0.f)) return "bn_eps must be positive";
    p.x_in = a->x; p.row_ptr = a->row_ptr; p.src = a->src_sorted; p.dst = a->dst_sorted; p.attr = a->attr_sorted;
    p.order = a->order; p.col_ptr = a->col_ptr; p.in_edges = a->in_edges; p.weights = a->weights;
    p.XS = a->xs; p.WS = a->ws; p.ZV = a->zv; p.ZE = a->ze; p.stats = a->stats;
    p.node_ws = a->node_ws; p.edge_ws = a->edge_ws; p.red = a->red;
    p.sync_ctr = a->sync_ws; p.grid_ctas = c > 8 ? c : 0; p.b0 = 0;
    p.out = a->heu_out; p.g_out = a->grad_heu; p.grad_w = a->grad_weights;
    p.n = a->n_nodes; p.E = a->n_edges; p.feats = a->feats; p.bn_eps = a->bn_eps;
    p.wc = (long long)U * a->feats + 3 * U + (long long)kDepth * kLayerFloats + kHeadFloats;
    return nullptr;
}

}  // namespace gnnt
}  // namespace deepaco
