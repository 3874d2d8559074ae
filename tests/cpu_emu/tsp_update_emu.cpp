// csrc/tsp_update.cuh and csrc/cvrp_update.cuh (tour cost, neighbour table, fused evaporate + ordered deposit) compiled
// for the host (see cuda_emu.h) behind entry points shaped like deepaco_{tsp,cvrp}_cost / _update.  Test infrastructure only.
#include "cuda_emu.h"

#include <algorithm>
#include <vector>
#include <cmath>
using std::min;

struct uint4 {
    uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#define DACO_NOINLINE __attribute__((noinline))
#define DACO_DYN_SMEM128(name) unsigned char* name = emu::ctx.smem
#define DACO_DYN_SMEM16(name) unsigned char* name = emu::ctx.smem
#define __shared__ static

namespace deepaco {   // bulk-copy stand-ins (cuda_emu.h)
static inline void mbar_init(uint64_t* bar, uint32_t) { emu_mbar_init(bar); }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_expect_tx(bar, bytes); }
static inline void fence_barrier_init() {}
static inline void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { emu_bulk_copy(dst, src, bytes, bar); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu_mbar_wait(bar, parity); }
}  // namespace deepaco
#include "../../deepaco_b200/csrc/tsp_update.cuh"
#include "../../deepaco_b200/csrc/cvrp_update.cuh"
#include "../../deepaco_b200/csrc/backward.cuh"
#include "../../deepaco_b200/csrc/knn_graph.cuh"

using namespace deepaco;

namespace {
struct CostArgs {
    const float* dist; const int64_t* paths; const uint16_t* tours; int n, A, lbw, vec; float* costs; uint32_t* nbr;
};
struct UpdArgs {
    float* ph; const uint32_t* nbr; const float* costs; int n, A; float decay; int elitist, min_max; float ph_min;
    const float* ph_max; const float* scale; const float* heu; float* prod;
};
}  // namespace

// tile != 0: tsp_cost_tile_kernel (compact tours), else tsp_cost_kernel (paths or tours).  One colony per call.
extern "C" const char* emu_tsp_cost(const float* dist, const int64_t* paths, const uint16_t* tours, int n, int A, int lbw, int vec,
                                    float* costs, uint32_t* nbr, int tile) {
    if (!dist || ((paths != nullptr) == (tours != nullptr)) || n < 2 || A < 1) return "bad arguments";
    const CostArgs a{dist, paths, tours, n, A, lbw, vec, costs, nbr};
    if (tile) {
        if (!tours) return "the tile kernel needs compact tours";
        emu::launch([](const CostArgs& q) { tsp_cost_tile_kernel(q.dist, q.tours, q.n, q.A, q.lbw, q.vec, q.costs, q.nbr); }, a,
                    (A + 31) / 32, 1, 256, (size_t)32 * n * 4);
    } else {
        const int W = 8;
        emu::launch([](const CostArgs& q) { tsp_cost_kernel(q.dist, q.paths, q.tours, q.n, q.A, q.lbw, q.vec, q.costs, q.nbr); }, a,
                    (A + W - 1) / W, 1, W * 32, 16);
    }
    return nullptr;
}

extern "C" const char* emu_tsp_update(float* ph, const uint32_t* nbr, const float* costs, int n, int A, float decay, int elitist,
                                      int min_max, float ph_min, const float* ph_max, const float* scale, const float* heu, float* prod) {
    if (!ph || !nbr || !costs || n < 2 || A < 1 || (min_max && !ph_max)) return "bad arguments";
    const int W = 4;
    const size_t per_warp = (((size_t)2 * A * 4 + (size_t)(2 * n + 1) * 4) + 15) & ~(size_t)15;
    const size_t smem = per_warp * W + (((size_t)A * 4 + 15) & ~(size_t)15);
    const UpdArgs a{ph, nbr, costs, n, A, decay, elitist, min_max, ph_min, ph_max, scale, heu, prod};
    emu::launch([](const UpdArgs& q) { tsp_update_kernel(q.ph, q.nbr, q.costs, q.n, q.A, q.decay, q.elitist, q.min_max, q.ph_min,
                                                        q.ph_max, q.scale, q.heu, q.prod); },
                a, (n + W - 1) / W, 1, W * 32, smem);
    return nullptr;
}

// tsp_update_row_kernel (one CTA per row, ants in chunks of CH, W warps): the many-ants form of the same update
extern "C" const char* emu_tsp_update_rows(float* ph, const uint32_t* nbr, const float* costs, int n, int A, int CH, int W, float decay,
                                           int min_max, float ph_min, const float* ph_max, const float* scale, const float* heu,
                                           float* prod) {
    if (!ph || !nbr || !costs || n < 2 || A < 1 || CH < 1 || W < 1 || W > 8 || (min_max && !ph_max)) return "bad arguments";
    struct RowArgs {
        UpdArgs u; int CH;
    };
    const RowArgs a{{ph, nbr, costs, n, A, decay, 0, min_max, ph_min, ph_max, scale, heu, prod}, CH};
    const size_t smem = (size_t)n * 4 + (size_t)(n + 1) * 4 + (size_t)W * n * 4 + (size_t)4 * CH * 4;
    emu::launch([](const RowArgs& r) { const UpdArgs& q = r.u; tsp_update_row_kernel(q.ph, q.nbr, q.costs, q.n, q.A, r.CH, q.decay, q.min_max,
                                                                                 q.ph_min, q.ph_max, q.scale, q.heu, q.prod); },
                a, n, 1, W * 32, smem);
    return nullptr;
}

// tsp_update_seq_kernel (one CTA per colony, matrix in shared memory, ants one after another)
extern "C" const char* emu_tsp_update_seq(float* ph, const uint16_t* tours, const float* costs, int n, int A, int threads, float decay,
                                          int elitist, int min_max, float ph_min, const float* ph_max, const float* scale,
                                          const float* heu, float* prod) {
    if (!ph || !tours || !costs || n < 3 || A < 1 || threads < n || (min_max && !ph_max)) return "bad arguments";
    struct SeqArgs {
        float* ph; const uint16_t* t; const float* c; int n, A; float decay; int el, mm; float mn; const float* mx; const float* sc;
        const float* heu; float* prod;
    };
    const SeqArgs a{ph, tours, costs, n, A, decay, elitist, min_max, ph_min, ph_max, scale, heu, prod};
    emu::launch([](const SeqArgs& q) { tsp_update_seq_kernel(q.ph, q.t, q.c, q.n, q.A, q.decay, q.el, q.mm, q.mn, q.mx, q.sc, q.heu, q.prod); },
                a, 1, 1, threads, ((size_t)n * n + A) * 4 + (size_t)2 * (16 * n + 2) * 2);
    return nullptr;
}

// tsp_tail_kernel: cost + best tracking + ant-sequential update of one colony in one launch
extern "C" const char* emu_tsp_tail(float* ph, const uint16_t* tours, const float* dist, const float* heu, float* prod, float* costs,
                                    float* lowest, int64_t* shortest, float* ph_max, int n, int A, float decay, int elitist, int min_max,
                                    float ph_min, int lbw, int vec) {
    if (!ph || !tours || !dist || !heu || !prod || !costs || !lowest || !shortest || n < 3 || A < 1) return "bad arguments";
    const TailParams p{ph, tours, dist, heu, prod, costs, lowest, shortest, ph_max, n, A, decay, elitist, min_max, ph_min, lbw, vec};
    emu::launch(tsp_tail_kernel, p, 1, 1, 256, ((size_t)2 * n * n + A) * 4 + (size_t)2 * (16 * n + 2) * 2);
    return nullptr;
}

// knn_refresh_kernel: candidate lists (columns of the 32 largest entries per row) from a product matrix [rows][n]
extern "C" const char* emu_knn_refresh(const float* prod, uint8_t* knn, int n, int rows) {
    if (!prod || !knn || n <= 32 || n > 256 || rows < 1) return "bad arguments";
    struct A { const float* p; uint8_t* k; int n, rows; };
    const A a{prod, knn, n, rows};
    emu::launch([](const A& q) { knn_refresh_kernel(q.p, q.k, q.n, q.rows); }, a, (rows + 7) / 8, 1, 256, 16);
    return nullptr;
}

// knn_graph_kernel: distance matrix + k smallest entries per row + edge_index for a batch (one warp per row)
extern "C" const char* emu_knn_graph(const float* coords, const float* dist_in, int n, int B, int k, float diag, float* dist_out,
                                     int32_t* idx, float* val, int64_t* edge_index) {
    if ((coords != nullptr) == (dist_in != nullptr) || n < 1 || B < 1 || k < 0 || k > n) return "bad arguments";
    const KnnGraphParams p{coords, dist_in, dist_out, idx, val, edge_index, n, B, k, diag};
    const long rows = (long)B * n;
    emu::launch(knn_graph_kernel, p, (int)((rows + 7) / 8), 1, 256, (size_t)8 * n * 4);
    return nullptr;
}

// ---- CVRP (one colony per call) ----
namespace {
struct CvrpCostArgs {
    const float* dist; const int64_t* paths; const uint16_t* tours; int N, A, rows_in, T; float* costs; uint32_t* nbr;
};
}  // namespace

extern "C" const char* emu_cvrp_cost(const float* dist, const int64_t* paths, const uint16_t* tours, int N, int A, int rows_in, int T,
                                     float* costs, uint32_t* nbr) {
    if (!dist || ((paths != nullptr) == (tours != nullptr)) || N < 2 || A < 1 || T < 1 || T > rows_in) return "bad arguments";
    const CvrpCostArgs a{dist, paths, tours, N, A, rows_in, T, costs, nbr};
    const int W = 8;
    emu::launch([](const CvrpCostArgs& q) { cvrp_cost_kernel(q.dist, q.paths, q.tours, q.N, q.A, q.rows_in, nullptr, q.T, q.costs, q.nbr); },
                a, (A + W - 1) / W, 1, W * 32, 16);
    return nullptr;
}

extern "C" const char* emu_cvrp_update(float* ph, const uint32_t* nbr, const float* costs, int N, int A, float decay, int elitist,
                                       int min_max, float ph_min, const float* ph_max, const float* scale) {
    if (!ph || !nbr || !costs || N < 2 || A < 1 || (min_max && !ph_max)) return "bad arguments";
    const UpdArgs a{ph, nbr, costs, N, A, decay, elitist, min_max, ph_min, ph_max, scale, nullptr, nullptr};
    const int threads = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    emu::launch([](const UpdArgs& q) { cvrp_update_kernel(q.ph, q.nbr, q.costs, q.n, q.A, q.decay, q.elitist, q.min_max, q.ph_min,
                                                         q.ph_max, q.scale, q.heu, q.prod); },
                a, N, 1, threads, (size_t)A * 8);
    return nullptr;
}

// ---- analytic backward of the log-probabilities (deepaco_logp_backward): prepare + ordered row accumulation ----
extern "C" const char* emu_logp_backward(const float* ph, const float* heu, const int64_t* paths, const float* glogp, int n, int A,
                                         int rows, const float* demand, float capacity, float* g_heu, float* g_ph) {
    if (!ph || !heu || !paths || !glogp || !g_heu || n < 2 || A < 1 || rows < 2) return "bad arguments";
    const size_t steps = (size_t)(rows - 1) * A;
    std::vector<float> coef(steps), gact(steps), rem(steps);
    std::vector<uint8_t> dok(steps);
    std::vector<uint16_t> when((size_t)A * n, 0xffff), dsteps((size_t)A * rows);
    std::vector<int32_t> dcnt(A);
    const BackwardParams p{ph, heu, paths, glogp, g_heu, g_ph, demand, capacity, n, A, rows, coef.data(), gact.data(), rem.data(),
                           dok.data(), when.data(), dsteps.data(), dcnt.data()};
    emu::launch(logp_backward_prepare_kernel, p, (A + 7) / 8, 1, 256, 16);
    emu::launch(logp_backward_rows_kernel, p, n, 1, 128, 16);
    return nullptr;
}
