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

#include <stdlib.h>

namespace deepaco {

// ---- cost (cvrp/aco.py:132-136) + neighbour table ------------------------------------------------
// costs[a] = sum_{k < T} dist[u_k][u_{k+1}], T = tmax[b] (path rows - 1), padding pairs (0,0) included,
// in ATen's order for a contiguous [A][T] input.
// neighbours[b][u][a] for customers u >= 1: (pred << 16) | succ ; for u = 0: 1 if the ant's padded path
// contains a (0,0) pair (i.e. it finished before the slowest ant), else 0.
struct CvrpTourView {
    const int64_t* paths;   // [rows_in][A] of this colony or null
    const uint16_t* tour;   // [rows_in] of this ant or null
    int A, a, rows_in;
    __device__ __forceinline__ int at(int k) const {
        if (k >= rows_in) return 0;
        return paths ? (int)paths[(size_t)k * A + a] : (int)tour[k];
    }
};

__global__ void __launch_bounds__(256) cvrp_cost_kernel(const float* __restrict__ dist, const int64_t* __restrict__ paths,
                                                        const uint16_t* __restrict__ tours, int N, int A, int rows_in,
                                                        const int32_t* __restrict__ tmax, int T_fixed,
                                                        float* __restrict__ costs, uint32_t* __restrict__ nbr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + warp;
    const int b = blockIdx.y;
    if (a >= A) return;
    const int T = tmax ? tmax[b] : T_fixed;
    const float* D = dist + (size_t)b * N * N;
    CvrpTourView tv{paths ? paths + (size_t)b * rows_in * A : nullptr,
                    tours ? tours + ((size_t)b * A + a) * rows_in : nullptr, A, a, rows_in};
    if (costs) {
        int lbw, vec;
        aten_sum_plan_dev(T, A, &lbw, &vec);
        auto edge = [&](int k) -> float { return __ldg(D + (size_t)tv.at(k) * N + tv.at(k + 1)); };
        const float c = aten_row_sum_fn(edge, T, lbw, vec != 0, lane, vec ? (int)(((unsigned)a * (unsigned)T) & 3u) : 0);
        if (lane == 0) costs[(size_t)b * A + a] = c;
    }
    if (nbr) {
        uint32_t* Nb = nbr + (size_t)b * N * A;
        bool pad = false;
        for (int k = lane; k < T; k += 32) {
            const int u = tv.at(k), v = tv.at(k + 1);
            if (u == 0 && v == 0) pad = true;
            if (v != 0) {   // customer v: predecessor u, successor at k+2
                const int w = tv.at(k + 2);
                Nb[(size_t)v * A + a] = ((uint32_t)u << 16) | (uint32_t)w;
            }
        }
        pad = __any_sync(DACO_FULL, pad);
        if (lane == 0) Nb[a] = pad ? 1u : 0u;
    }
}

// grid (N rows, B); cvrp/aco.py:106-130: ph *= decay; per ant (in order) ph[path[k], path[k+1]] += 1/cost
// (index_put without accumulate: repeated (0,0) pairs count once); optional min_max clamp; 1e-10 floor.
__global__ void __launch_bounds__(256) cvrp_update_kernel(float* __restrict__ ph, const uint32_t* __restrict__ nbr,
                                                          const float* __restrict__ costs, int N, int A, float decay,
                                                          int elitist, int min_max, float ph_min,
                                                          const float* __restrict__ ph_max, const float* __restrict__ scale,
                                                          const float* __restrict__ heu, float* __restrict__ prod) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* nb_s = reinterpret_cast<uint32_t*>(smem);
    float* w_s = reinterpret_cast<float*>(smem) + A;
    __shared__ int best_ant;
    const int u = blockIdx.x, b = blockIdx.y;
    const uint32_t* Nb = nbr + (size_t)b * N * A;
    const float* C = costs + (size_t)b * A;
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        nb_s[a] = Nb[(size_t)u * A + a];
        w_s[a] = __fdiv_rn(1.0f, C[a]);
    }
    if (threadIdx.x < 32) {
        float bc = INFINITY;
        int bi = 0x7fffffff;
        for (int a = threadIdx.x; a < A; a += 32) {
            const float c = C[a];
            if (c < bc) { bc = c; bi = a; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
            const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
            if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        }
        if (threadIdx.x == 0) best_ant = bi;
    }
    __syncthreads();
    float* row = ph + ((size_t)b * N + u) * N;
    const float hi = min_max ? ph_max[b] : 0.f;
    const int a_lo = elitist ? best_ant : 0, a_hi = elitist ? best_ant + 1 : A;
    for (int v = threadIdx.x; v < N; v += blockDim.x) {
        float val = row[v];
        if (scale) val = __fmul_rn(val, scale[b]);   // MMAS rescale on the first improvement (cvrp/aco.py:90-92)
        val = __fmul_rn(val, decay);
        if (u != 0) {
            for (int a = a_lo; a < a_hi; ++a)
                if ((int)(nb_s[a] & 0xffffu) == v) val = __fadd_rn(val, w_s[a]);          // (u -> succ_a(u))
        } else if (v != 0) {
            for (int a = a_lo; a < a_hi; ++a)
                if ((Nb[(size_t)v * A + a] >> 16) == 0u) val = __fadd_rn(val, w_s[a]);     // (0 -> v): pred_a(v) == 0
        } else {
            for (int a = a_lo; a < a_hi; ++a)
                if (nb_s[a] != 0u) val = __fadd_rn(val, w_s[a]);                           // padded (0,0), once per ant
        }
        if (min_max) {
            const float gate = __fmul_rn(val > 1e-9f ? 1.0f : 0.0f, val);
            if (gate < ph_min) val = ph_min;
            if (val > hi) val = hi;
        }
        if (val < 1e-10f) val = 1e-10f;   // cvrp/aco.py:130
        row[v] = val;
        if (prod) prod[((size_t)b * N + u) * N + v] = __fmul_rn(val, heu[((size_t)b * N + u) * N + v]);
    }
}

}  // namespace deepaco

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
    static float* cvrp_prod_ws = nullptr;
    static size_t cvrp_prod_ws_bytes = 0;
    if (global_p) {
        W = 4;
        if (heuristic) {   // product once per call into a scratch matrix; rows are then gathered from L2
            const size_t need_b = (size_t)n_colonies * n_nodes * n_nodes * sizeof(float);
            if (need_b > cvrp_prod_ws_bytes) {
                if (cvrp_prod_ws) cudaFree(cvrp_prod_ws);
                cvrp_prod_ws = nullptr; cvrp_prod_ws_bytes = 0;
                DACO_CHECK_CUDA(cudaMalloc(&cvrp_prod_ws, need_b));
                cvrp_prod_ws_bytes = need_b;
            }
            const size_t cnt = (size_t)n_colonies * n_nodes * n_nodes;
            hadamard3_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(pheromone, heuristic, cvrp_prod_ws, cnt);
            DACO_CHECK_LAUNCH();
            p.ph = cvrp_prod_ws; p.heu = nullptr;
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
