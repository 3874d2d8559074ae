// CVRP construction, cost and pheromone update (reference cvrp/aco.py:106-205, adaptive=False).
//
// gen_path (:138-165): every ant starts at the depot 0; per step the candidate weights are
//   P[cur][j] * visit_mask[j] * capacity_mask[j]                                 (pick_move :167-174)
//   visit rule    (:176-180): visited customers are masked; the depot is allowed except when the ant is AT
//                             the depot while customers remain;
//   capacity rule (:182-202): used = (cur == 0 ? 0 : used) + demand[cur]; j masked iff demand[j] > capacity - used;
//   done          (:204-205): all customers visited and the ant is back at the depot.
// The reference loops until the slowest ant is done and keeps drawing node 0 for finished ants; a finished
// ant has a single non-zero candidate, so its draws do not depend on the noise and each ant can stop on
// its own.  Path rows past an ant's end are 0, log-probs are log(1 - eps).
// Same machinery as the TSP list kernel: P staged in shared memory by TMA, unvisited-customer list per
// warp, approximate-then-verified arg-max with exact_step() as the tie fallback.
#include "common.cuh"
#include "host_util.h"
#include "sample_common.cuh"
#include "list_kernel.cuh"
#include "cvrp_update.cuh"

#include <stdlib.h>


namespace deepaco {
__global__ void hadamard3_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        o[i] = __fmul_rn(a[i], b[i]);
}
}  // namespace deepaco

using namespace deepaco;

extern "C" int deepaco_cvrp_sample(const float* pheromone, const float* heuristic, const float* demand, float capacity,
                                   int n_nodes, int n_ants, int n_colonies, uint64_t seed, uint64_t offset,
                                   const uint64_t* offsets, const float* noise, int path_rows, int64_t* paths,
                                   float* log_probs, uint16_t* tours, int32_t* lens, int32_t* tmax, void* stream) {
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(pheromone && demand && lens && tmax, "deepaco_cvrp_sample: NULL argument");
    DACO_CHECK_ARG(n_nodes >= 2 && n_ants >= 1 && n_colonies >= 1 && n_colonies <= 65535, "deepaco_cvrp_sample: bad sizes");
    DACO_CHECK_ARG(path_rows == 2 * n_nodes, "deepaco_cvrp_sample: path_rows must be 2 * n_nodes (=%d)", 2 * n_nodes);
    cudaStream_t st = (cudaStream_t)stream;
    ListParams p{};
    p.ph = pheromone; p.heu = heuristic; p.demand = demand; p.capacity = capacity;
    p.n = n_nodes; p.A = n_ants; p.B = n_colonies; p.rows = path_rows;
    p.start_node = 0; p.double_norm = 0;
    p.seed = seed; p.offset = offset; p.offsets = offsets; p.noise = noise;
    p.keys.init(seed);
    p.paths = paths; p.logp = log_probs; p.tours = tours; p.lens = lens; p.tmax = tmax;
    const SumPlan sp = aten_sum_plan(n_nodes, n_ants);
    int bw = sp.block_width > 32 ? 32 : sp.block_width, lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    p.lbw = lbw; p.vec = sp.vectorized;
    const DrawPlan dn = torch_draw_plan((int64_t)n_ants * n_nodes, *di);
    DACO_CHECK_ARG(dn.single && (uint64_t)n_ants * n_nodes < (1ull << 32),
                   "deepaco_cvrp_sample: n_ants * n_nodes = %ld exceeds the single-draw Philox geometry", (long)n_ants * n_nodes);
    p.g_noise = {dn.threads, dn.single};
    p.g_start = p.g_noise;
    p.step_increment = (uint32_t)dn.increment;

    const long total_ants = (long)n_ants * n_colonies;
    int W = total_ants <= (long)di->sm_count * 4 ? 4 : 8;
    if (const char* e = getenv("DEEPACO_TSP_WARPS")) {
        const int w = atoi(e);
        if (w >= 1 && w <= 16) W = w;
    }
    const size_t cap = (size_t)di->max_smem_optin - 1024;
    const bool global_p = !(n_nodes <= 256 && list_kernel_smem(n_nodes, path_rows, W, true) <= cap);
    DACO_CHECK_ARG(n_nodes <= 1024, "deepaco_cvrp_sample: n_nodes=%d exceeds the supported maximum of 1024", n_nodes);
    StreamScratch prod_ws;   // stream-ordered scratch, private to this call (freed behind the kernel on return)
    if (global_p) {
        W = 4;
        if (heuristic) {   // product once per call into a scratch matrix; rows are then gathered from L2
            const size_t cnt = (size_t)n_colonies * n_nodes * n_nodes;
            DACO_CHECK_CUDA(prod_ws.alloc(cnt * sizeof(float), st));
            float* ws = static_cast<float*>(prod_ws.ptr);
            hadamard3_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(pheromone, heuristic, ws, cnt);
            DACO_CHECK_LAUNCH();
            p.ph = ws; p.heu = nullptr;
        }
    }
    auto need = [&](int w) { return list_kernel_smem(n_nodes, path_rows, w, true, global_p); };
    if (!global_p && total_ants > (long)di->sm_count * 4)
        while (W < 16 && (cap / need(W)) * W < 32 && need(W * 2) <= cap) W *= 2;
    DACO_CHECK_CUDA(cudaMemsetAsync(tmax, 0, sizeof(int32_t) * n_colonies, st));
    dim3 grid((n_ants + W - 1) / W, n_colonies);
    const int epl = (n_nodes + 31) / 32;
    const size_t sm = need(W);
#define DACO_LAUNCH1(KFN)                                                                                          \
    do {                                                                                                           \
        DACO_CHECK_CUDA(cudaFuncSetAttribute(KFN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));          \
        DACO_CHECK_CUDA(cudaFuncSetAttribute(KFN, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
        KFN<<<grid, W * 32, sm, st>>>(p);                                                                          \
        DACO_CHECK_LAUNCH();                                                                                       \
        return DEEPACO_OK;                                                                                         \
    } while (0)
#define DACO_LIST(E, G)                                                              \
    do {                                                                             \
        if (noise) {                                                                 \
            if (log_probs) DACO_LAUNCH1((aco_list_kernel<E, true, true, true, G>));  \
            DACO_LAUNCH1((aco_list_kernel<E, true, false, true, G>));                \
        }                                                                            \
        if (log_probs) DACO_LAUNCH1((aco_list_kernel<E, true, true, false, G>));     \
        DACO_LAUNCH1((aco_list_kernel<E, true, false, false, G>));                   \
    } while (0)
    if (global_p) {
        if (epl <= 8) DACO_LIST(8, true);
        if (epl <= 16) DACO_LIST(16, true);
        DACO_LIST(32, true);
    }
    if (epl <= 1) DACO_LIST(1, false);
    if (epl <= 2) DACO_LIST(2, false);
    if (epl <= 4) DACO_LIST(4, false);
    DACO_LIST(8, false);
#undef DACO_LAUNCH1
#undef DACO_LIST
}

extern "C" uint64_t deepaco_cvrp_step_offset_increment(int n_nodes, int n_ants) {
    const DeviceInfo* di = device_info();
    if (!di) return 0;
    return torch_draw_plan((int64_t)n_ants * n_nodes, *di).increment;
}

extern "C" int deepaco_cvrp_cost(const float* distances, const int64_t* paths, const uint16_t* tours, int n_nodes,
                                 int n_ants, int n_colonies, int rows_in, const int32_t* tmax, int T_fixed, float* costs,
                                 uint32_t* neighbours, void* stream) {
    DACO_CHECK_ARG(distances && (costs || neighbours), "deepaco_cvrp_cost: NULL distances / no output requested");
    DACO_CHECK_ARG((paths != nullptr) != (tours != nullptr), "deepaco_cvrp_cost: pass exactly one of paths / tours");
    DACO_CHECK_ARG(n_nodes >= 2 && n_nodes <= 65535 && n_ants >= 1 && n_colonies >= 1 && rows_in >= 1, "deepaco_cvrp_cost: bad sizes");
    DACO_CHECK_ARG(tmax || (T_fixed >= 1 && T_fixed < rows_in + 1), "deepaco_cvrp_cost: need tmax or a valid T_fixed");
    const int W = 8;
    dim3 grid((n_ants + W - 1) / W, n_colonies);
    cvrp_cost_kernel<<<grid, W * 32, 0, (cudaStream_t)stream>>>(distances, paths, tours, n_nodes, n_ants, rows_in, tmax, T_fixed,
                                                                costs, neighbours);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

extern "C" int deepaco_cvrp_update(float* pheromone, const uint32_t* neighbours, const float* costs, int n_nodes, int n_ants,
                                   int n_colonies, float decay, int elitist, int min_max, float ph_min, const float* ph_max,
                                   void* stream) {
    DACO_CHECK_ARG(pheromone && neighbours && costs, "deepaco_cvrp_update: NULL argument");
    DACO_CHECK_ARG(n_nodes >= 2 && n_ants >= 1 && n_colonies >= 1, "deepaco_cvrp_update: bad sizes");
    DACO_CHECK_ARG(!min_max || ph_max, "deepaco_cvrp_update: min_max needs ph_max");
    const size_t smem = (size_t)n_ants * 8;
    DACO_CHECK_ARG(smem <= 200 * 1024, "deepaco_cvrp_update: n_ants=%d too large for one pass", n_ants);
    DACO_CHECK_CUDA(cudaFuncSetAttribute(cvrp_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = n_nodes <= 64 ? 64 : (n_nodes <= 128 ? 128 : 256);
    dim3 grid(n_nodes, n_colonies);
    cvrp_update_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(pheromone, neighbours, costs, n_nodes, n_ants, decay,
                                                                      elitist, min_max, ph_min, ph_max, nullptr, nullptr, nullptr);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

// ---- ACO.run for CVRP colonies (cvrp/aco.py:72-104, adaptive = False) without host round trips ---------------
namespace deepaco {
int best_launch(const float* costs, const uint16_t* tours, const float* ph, int n, int tour_len, int A, int B, int min_max,
                float* lowest, int64_t* shortest, float* ph_max, float* scale, const int32_t* tmax, int32_t* shortest_rows,
                cudaStream_t st);


// the reference draws one [A, N] exponential_ per construction step until the slowest ant is done: the next
// iteration's Philox offset is data dependent, so it is advanced on the device
__global__ void advance_offsets_kernel(uint64_t* offsets, const int32_t* tmax, uint64_t step_increment, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) offsets[b] += (uint64_t)tmax[b] * step_increment;
}
}  // namespace deepaco

extern "C" int deepaco_cvrp_run(const deepaco_cvrp_run_args* a, int n_iterations, void* stream) {
    DACO_CHECK_ARG(a && n_iterations >= 0, "deepaco_cvrp_run: bad arguments");
    DACO_CHECK_ARG(a->pheromone && a->heuristic && a->distances && a->demand && a->product && a->tours && a->costs &&
                       a->neighbours && a->lens && a->tmax && a->offsets && a->lowest_cost && a->shortest_path,
                   "deepaco_cvrp_run: NULL buffer");
    DACO_CHECK_ARG(!a->min_max || (a->ph_max && a->scale), "deepaco_cvrp_run: min_max needs ph_max and scale buffers");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = a->n_nodes, A = a->n_ants, B = a->n_colonies, R = 2 * N;
    const uint64_t inc = deepaco_cvrp_step_offset_increment(N, A);
    const size_t cnt = (size_t)B * N * N;
    if (n_iterations > 0 && !a->product_valid) {
        hadamard3_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(a->pheromone, a->heuristic, a->product, cnt);
        DACO_CHECK_LAUNCH();
    }
    for (int it = 0; it < n_iterations; ++it) {
        int rc = deepaco_cvrp_sample(a->product, nullptr, a->demand, a->capacity, N, A, B, a->seed, 0, a->offsets, nullptr, R,
                                     nullptr, nullptr, a->tours, a->lens, a->tmax, st);
        if (rc) return rc;
        rc = deepaco_cvrp_cost(a->distances, nullptr, a->tours, N, A, B, R, a->tmax, 0, a->costs, a->neighbours, st);
        if (rc) return rc;
        rc = best_launch(a->costs, a->tours, a->pheromone, N, R, A, B, a->min_max, a->lowest_cost, a->shortest_path, a->ph_max,
                         a->min_max ? a->scale : nullptr, a->tmax, a->shortest_rows, st);
        if (rc) return rc;
        const size_t smem = (size_t)A * 8;
        DACO_CHECK_CUDA(cudaFuncSetAttribute(cvrp_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int threads = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
        cvrp_update_kernel<<<dim3(N, B), threads, smem, st>>>(a->pheromone, a->neighbours, a->costs, N, A, a->decay, a->elitist,
                                                              a->min_max, a->ph_min, a->ph_max, a->min_max ? a->scale : nullptr,
                                                              a->heuristic, a->product);
        DACO_CHECK_LAUNCH();
        advance_offsets_kernel<<<(B + 127) / 128, 128, 0, st>>>(a->offsets, a->tmax, inc, B);
        DACO_CHECK_LAUNCH();
    }
    return DEEPACO_OK;
}
