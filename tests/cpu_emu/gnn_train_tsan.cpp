// ThreadSanitizer run of the training-mode GNN kernel source on a small synthetic graph (irregular degrees):
// every cross-thread shared- or global-memory dependency that is not ordered by a barrier is reported as a race.
// Test infrastructure only (see cuda_emu.h).  Exit code 0 = no race, finite outputs.
#include "gnn_train_emu.cpp"

#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>

int main(int argc, char** argv) {
    const int ctas = argc > 1 ? atoi(argv[1]) : 2, nth_f = argc > 2 ? atoi(argv[2]) : 64, nth_b = argc > 3 ? atoi(argv[3]) : 128;
    const int n = argc > 4 ? atoi(argv[4]) : 97, F = 2, B = 2;
    std::mt19937 rng(5);
    std::uniform_real_distribution<float> uni(-0.3f, 0.3f);
    // irregular graph: node i has (i % 5) + 1 out-edges, node 7 none
    std::vector<int32_t> src, dst;
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < (i == 7 ? 0 : (i % 5) + 1); ++k) { src.push_back(i); dst.push_back((i * 7 + k * 3 + 1) % n); }
    const int E = (int)src.size();
    std::vector<int32_t> row_ptr(n + 1, 0), col_ptr(n + 1, 0), order(E), in_edges(E);
    for (int e = 0; e < E; ++e) { row_ptr[src[e] + 1]++; col_ptr[dst[e] + 1]++; }
    for (int i = 0; i < n; ++i) { row_ptr[i + 1] += row_ptr[i]; col_ptr[i + 1] += col_ptr[i]; }
    std::iota(order.begin(), order.end(), 0);
    std::iota(in_edges.begin(), in_edges.end(), 0);
    std::stable_sort(in_edges.begin(), in_edges.end(), [&](int a, int b) { return dst[a] < dst[b]; });
    auto rep = [&](const std::vector<int32_t>& v) { std::vector<int32_t> r; for (int b = 0; b < B; ++b) r.insert(r.end(), v.begin(), v.end()); return r; };
    std::vector<int32_t> row_ptr_b = rep(row_ptr), col_ptr_b = rep(col_ptr), src_b = rep(src), dst_b = rep(dst), order_b = rep(order), in_b = rep(in_edges);
    const long long wc = 32LL * F + 96 + 12LL * deepaco::gnnt::kLayerFloats + deepaco::gnnt::kHeadFloats;
    std::vector<float> weights(wc), x((size_t)B * n * F), attr((size_t)B * E), g_heu((size_t)B * E);
    for (auto& v : weights) v = uni(rng);
    for (auto& v : x) v = uni(rng) + 0.5f;
    for (auto& v : attr) v = uni(rng) + 0.5f;
    for (auto& v : g_heu) v = uni(rng);
    std::vector<float> xs((size_t)B * 13 * n * 32), ws((size_t)B * 13 * E * 32), zv((size_t)B * 12 * n * 32), ze((size_t)B * 12 * E * 32),
        stats((size_t)B * 12 * 6 * 32), node_ws((size_t)B * n * 224), edge_ws((size_t)B * E * 96), red((size_t)B * 36 * 64 * 128),
        heu((size_t)B * E), grad((size_t)B * ctas * wc, 0.f);
    deepaco_gnn_train_args a = {};
    a.n_nodes = n; a.n_edges = E; a.feats = F; a.n_instances = B; a.ctas_per_instance = ctas; a.bn_eps = 1e-5f;
    a.x = x.data(); a.row_ptr = row_ptr_b.data(); a.src_sorted = src_b.data(); a.dst_sorted = dst_b.data(); a.attr_sorted = attr.data();
    a.order = order_b.data(); a.col_ptr = col_ptr_b.data(); a.in_edges = in_b.data(); a.weights = weights.data();
    a.xs = xs.data(); a.ws = ws.data(); a.zv = zv.data(); a.ze = ze.data(); a.stats = stats.data(); a.node_ws = node_ws.data();
    std::vector<uint32_t> sync_ws(B, 0);
    a.sync_ws = sync_ws.data();
    a.edge_ws = edge_ws.data(); a.red = red.data(); a.heu_out = heu.data(); a.grad_heu = g_heu.data(); a.grad_weights = grad.data();
    {   // eval-mode group forward first (ping-pong state in the first two layers of xs / ws; positive invstd slots)
        for (int l = 0; l < 12; ++l)
            for (int k = 0; k < 2; ++k)
                for (int f = 0; f < 32; ++f) {
                    float& istd = weights[32 * F + 96 + (size_t)l * deepaco::gnnt::kLayerFloats + 5 * deepaco::gnnt::LIN + k * 128 + 96 + f];
                    istd = std::fabs(istd) + 0.5f;
                }
        if (const char* err = emu_gnn_forward_group(&a, nth_f)) { printf("eval forward: %s\n", err); return 2; }
        double es = 0;
        for (float v : heu) es += v;
        if (!std::isfinite(es)) { printf("non-finite eval output\n"); return 3; }
    }
    if (const char* err = emu_gnn_train_forward(&a, nth_f)) { printf("forward: %s\n", err); return 2; }
    if (const char* err = emu_gnn_train_backward(&a, nth_b)) { printf("backward: %s\n", err); return 2; }
    double hs = 0, gs = 0;
    for (float v : heu) hs += v;
    for (float v : grad) gs += std::fabs(v);
    if (!std::isfinite(hs) || !std::isfinite(gs)) { printf("non-finite output\n"); return 3; }
    printf("tsan run ok: ctas=%d threads=%d/%d  n=%d E=%d  sum(heu)=%.6f  sum|grad|=%.6f\n", ctas, nth_f, nth_b, n, E, hs, gs);
    return 0;
}
