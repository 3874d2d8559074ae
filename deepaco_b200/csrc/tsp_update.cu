// K2 -- tour cost and fused evaporate + deposit for TSP colonies: launchers and C ABI (kernels in tsp_update.cuh).
#include "tsp_update.cuh"
#include "host_util.h"

#include <stdlib.h>

#include <algorithm>

using namespace deepaco;

static const int kRowKernelMinAnts = 2048;   // from here on a row's 2A events are worth a whole CTA

extern "C" int deepaco_tsp_cost(const float* distances, const int64_t* paths, const uint16_t* tours, int n, int n_ants,
                                int n_colonies, float* costs, uint32_t* neighbours, void* stream) {
    DACO_CHECK_ARG(distances && (costs || neighbours), "deepaco_tsp_cost: NULL distances / no output requested");
    DACO_CHECK_ARG((paths != nullptr) != (tours != nullptr), "deepaco_tsp_cost: pass exactly one of paths / tours");
    DACO_CHECK_ARG(n >= 2 && n <= 65535 && n_ants >= 1 && n_colonies >= 1, "deepaco_tsp_cost: bad sizes");
    const SumPlan sp = aten_sum_plan(n, n_ants);
    int bw = sp.block_width > 32 ? 32 : sp.block_width, lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    const size_t tile_smem = (size_t)32 * n * 4;
    if (tours && tile_smem <= 96 * 1024 && !getenv("DEEPACO_COST_SIMPLE")) {
        DACO_CHECK_CUDA(cudaFuncSetAttribute(tsp_cost_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
        dim3 grid((n_ants + 31) / 32, n_colonies);
        tsp_cost_tile_kernel<<<grid, 256, tile_smem, (cudaStream_t)stream>>>(distances, tours, n, n_ants, lbw, sp.vectorized, costs,
                                                                           neighbours);
        DACO_CHECK_LAUNCH();
        return DEEPACO_OK;
    }
    const int W = 8;
    dim3 grid((n_ants + W - 1) / W, n_colonies);
    tsp_cost_kernel<<<grid, W * 32, 0, (cudaStream_t)stream>>>(distances, paths, tours, n, n_ants, lbw, sp.vectorized,
                                                               costs, neighbours);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

static int launch_tsp_update(float* pheromone, const uint32_t* neighbours, const float* costs, int n, int n_ants,
                             int n_colonies, float decay, int elitist, int min_max, float ph_min, const float* ph_max,
                             const float* scale, const float* heuristic, float* product, cudaStream_t st) {
    if (!elitist && n_ants >= kRowKernelMinAnts && n <= 4096 && !getenv("DEEPACO_UPDATE_WARP_ROWS")) {
        // many ants per colony: one CTA per matrix row, ants in chunks (tsp_update_row_kernel)
        const int Wr = 8, CH = std::min(n_ants, 4096);
        const size_t smem = (size_t)n * 4 + (size_t)(n + 1) * 4 + (size_t)Wr * n * 4 + (size_t)4 * CH * 4;
        DACO_CHECK_CUDA(cudaFuncSetAttribute(tsp_update_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(n, n_colonies);
        tsp_update_row_kernel<<<grid, Wr * 32, smem, st>>>(pheromone, neighbours, costs, n, n_ants, CH, decay, min_max, ph_min,
                                                          ph_max, scale, heuristic, product);
        DACO_CHECK_LAUNCH();
        return DEEPACO_OK;
    }
    int W = 4;   // rows (warps) per CTA; fewer when a row's 2 * n_ants deposit events need more shared memory
    const size_t per_warp = (((size_t)2 * n_ants * 4 + (size_t)(2 * n + 1) * 4) + 15) & ~(size_t)15;
    const size_t inv_bytes = ((size_t)n_ants * 4 + 15) & ~(size_t)15;
    while (W > 1 && per_warp * W + inv_bytes > 200 * 1024) W >>= 1;
    const size_t smem = per_warp * W + inv_bytes;
    DACO_CHECK_ARG(smem <= 200 * 1024, "deepaco_tsp_update: n_ants=%d / n=%d too large for one pass", n_ants, n);
    DACO_CHECK_CUDA(cudaFuncSetAttribute(tsp_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((n + W - 1) / W, n_colonies);
    tsp_update_kernel<<<grid, W * 32, smem, st>>>(pheromone, neighbours, costs, n, n_ants, decay, elitist, min_max, ph_min,
                                                  ph_max, scale, heuristic, product);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

namespace deepaco {
// ant-sequential update straight from the compact tours (no neighbour table): small / medium colonies whose pheromone
// matrix fits in shared memory
bool tsp_update_seq_ok(int n, int n_ants) {
    return n >= 3 && n <= 224 && n_ants <= 1024 && !getenv("DEEPACO_UPDATE_ROWS_ONLY");
}
// ... and worth it: the kernel is one sequential chain per colony (a shared-memory read-modify-write + barrier per ant),
// so it pays when many colonies run side by side (several CTAs per SM: n <= 128) and loses to the row-parallel kernel
// when only a few do.  Measured on 256 x TSP-100 x 512: the whole tail (costs + best + update, tsp_tail_kernel) 90 us
// against 87 + 5 + 149 us for the three row-parallel launches; one colony: 69 us against 15 for the update alone.
bool tsp_update_seq_preferred(int n, int n_ants, int n_colonies) {
    if (getenv("DEEPACO_UPDATE_SEQ")) return tsp_update_seq_ok(n, n_ants);
    return tsp_update_seq_ok(n, n_ants) && n <= 128 && n_colonies >= 64;
}
int tsp_update_seq_launch(float* pheromone, const uint16_t* tours, const float* costs, int n, int n_ants, int n_colonies,
                          float decay, int elitist, int min_max, float ph_min, const float* ph_max, const float* scale,
                          const float* heuristic, float* product, cudaStream_t st) {
    const size_t smem = ((size_t)n * n + n_ants) * sizeof(float) + (size_t)2 * (16 * n + 2) * sizeof(uint16_t);
    DACO_CHECK_ARG(tsp_update_seq_ok(n, n_ants) && smem <= 220 * 1024, "tsp_update_seq: n=%d / n_ants=%d outside its range", n, n_ants);
    DACO_CHECK_CUDA(cudaFuncSetAttribute(tsp_update_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = n <= 64 ? 128 : 256;     // >= 2n threads when n <= 128: one cell per thread and ant
    tsp_update_seq_kernel<<<n_colonies, threads, smem, st>>>(pheromone, tours, costs, n, n_ants, decay, elitist, min_max, ph_min, ph_max,
                                                            scale, heuristic, product);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
// cost + best tracking + ant-sequential update of one iteration in one launch per colony (tsp_tail_kernel)
int tsp_tail_launch(float* pheromone, const uint16_t* tours, const float* distances, const float* heuristic, float* product,
                    float* costs, float* lowest, int64_t* shortest, float* ph_max, int n, int n_ants, int n_colonies, float decay,
                    int elitist, int min_max, float ph_min, cudaStream_t st) {
    const size_t smem = ((size_t)2 * n * n + n_ants) * sizeof(float) + (size_t)2 * (16 * n + 2) * sizeof(uint16_t);
    DACO_CHECK_ARG(tsp_update_seq_ok(n, n_ants) && smem <= 220 * 1024, "tsp_tail: n=%d / n_ants=%d outside its range", n, n_ants);
    const SumPlan sp = aten_sum_plan(n, n_ants);
    int bw = sp.block_width > 32 ? 32 : sp.block_width, lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    TailParams p{pheromone, tours, distances, heuristic, product, costs, lowest, shortest, ph_max, n, n_ants, decay, elitist, min_max,
                 ph_min, lbw, sp.vectorized};
    DACO_CHECK_CUDA(cudaFuncSetAttribute(tsp_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tsp_tail_kernel<<<n_colonies, 256, smem, st>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
bool tsp_tail_ok(int n, int n_ants, int n_colonies) {
    return tsp_update_seq_preferred(n, n_ants, n_colonies) && ((size_t)2 * n * n + n_ants) * 4 + 4 * (16 * n + 2) <= 220 * 1024 &&
           !getenv("DEEPACO_NO_TAIL_KERNEL");
}
int knn_refresh_launch(const float* product, uint8_t* knn, int n, int n_colonies, cudaStream_t st) {
    DACO_CHECK_ARG(n > 32 && n <= 256, "knn refresh: 32 < n <= 256");
    const int rows = n * n_colonies;
    knn_refresh_kernel<<<(rows + 7) / 8, 256, 0, st>>>(product, knn, n, rows);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
int tsp_update_launch(float* pheromone, const uint32_t* neighbours, const float* costs, int n, int n_ants, int n_colonies,
                      float decay, int elitist, int min_max, float ph_min, const float* ph_max, const float* scale,
                      const float* heuristic, float* product, cudaStream_t st) {
    return launch_tsp_update(pheromone, neighbours, costs, n, n_ants, n_colonies, decay, elitist, min_max, ph_min, ph_max,
                             scale, heuristic, product, st);
}
int tsp_cost_launch(const float* distances, const uint16_t* tours, int n, int n_ants, int n_colonies, float* costs,
                    uint32_t* neighbours, cudaStream_t st) {
    return deepaco_tsp_cost(distances, nullptr, tours, n, n_ants, n_colonies, costs, neighbours, st);
}
}  // namespace deepaco

extern "C" int deepaco_tsp_update(float* pheromone, const uint32_t* neighbours, const float* costs, int n, int n_ants,
                                  int n_colonies, float decay, int elitist, int min_max, float ph_min,
                                  const float* ph_max, void* stream) {
    DACO_CHECK_ARG(pheromone && neighbours && costs, "deepaco_tsp_update: NULL argument");
    DACO_CHECK_ARG(n >= 2 && n <= 65535 && n_ants >= 1 && n_colonies >= 1, "deepaco_tsp_update: bad sizes");
    DACO_CHECK_ARG(!min_max || ph_max, "deepaco_tsp_update: min_max needs ph_max");
    return launch_tsp_update(pheromone, neighbours, costs, n, n_ants, n_colonies, decay, elitist, min_max, ph_min, ph_max,
                             nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int deepaco_tsp_update_tours(float* pheromone, const uint16_t* tours, const float* costs, int n, int n_ants,
                                        int n_colonies, float decay, int elitist, int min_max, float ph_min,
                                        const float* ph_max, void* stream) {
    DACO_CHECK_ARG(pheromone && tours && costs, "deepaco_tsp_update_tours: NULL argument");
    DACO_CHECK_ARG(n >= 3 && n_ants >= 1 && n_colonies >= 1, "deepaco_tsp_update_tours: bad sizes");
    DACO_CHECK_ARG(!min_max || ph_max, "deepaco_tsp_update_tours: min_max needs ph_max");
    DACO_CHECK_ARG(tsp_update_seq_ok(n, n_ants), "deepaco_tsp_update_tours: needs 3 <= n <= 224 and n_ants <= 1024 (use deepaco_tsp_cost's "
                   "neighbour table + deepaco_tsp_update beyond that)");
    return tsp_update_seq_launch(pheromone, tours, costs, n, n_ants, n_colonies, decay, elitist, min_max, ph_min, ph_max, nullptr, nullptr,
                                 nullptr, (cudaStream_t)stream);
}
