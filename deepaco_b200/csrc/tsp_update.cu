// K2 -- tour cost and fused evaporate + deposit for TSP colonies.
//
// cost:   ACO.gen_path_costs (reference tsp/aco.py:120-132): sum_k dist[u_k][u_{k-1}], accumulated in
//         the order ATen's sum kernel uses for a contiguous [n_ants][n] input, so costs are bit-equal.
//         As a by-product each ant writes, per node u, its tour predecessor and successor.
// update: ACO.update_pheronome (tsp/aco.py:94-118).  The reference adds 1/cost_a to cells
//         (u, pred_a(u)) and (u, succ_a(u)) one ant at a time (index_put, non-accumulating), so every
//         matrix cell sees its additions in ant order.  One CTA per matrix row replays exactly that
//         order per cell from the neighbour table: deterministic, atomics-free, and every row is read
//         and written once, coalesced, with the evaporation folded in.
#include "common.cuh"
#include "host_util.h"

namespace deepaco {

struct TourView {
    const int64_t* paths;    // [n][A] of this colony, or null
    const uint16_t* tour;    // [n] of this ant, or null
    int A, a;
    __device__ __forceinline__ int at(int k) const {
        return paths ? (int)paths[(size_t)k * A + a] : (int)tour[k];
    }
};

__global__ void __launch_bounds__(256) tsp_cost_kernel(const float* __restrict__ dist, const int64_t* __restrict__ paths,
                                                       const uint16_t* __restrict__ tours, int n, int A, int lbw, int vec,
                                                       float* __restrict__ costs, uint32_t* __restrict__ nbr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + warp;
    const int b = blockIdx.y;
    if (a >= A) return;
    const float* D = dist + (size_t)b * n * n;
    TourView tv{paths ? paths + (size_t)b * n * A : nullptr, tours ? tours + ((size_t)b * A + a) * n : nullptr, A, a};
    auto edge = [&](int k) -> float {
        const int u = tv.at(k);
        const int v = tv.at(k == 0 ? n - 1 : k - 1);
        return __ldg(D + (size_t)u * n + v);
    };
    if (costs) {
        const float c = aten_row_sum_fn(edge, n, lbw, vec != 0, lane, vec ? (int)(((unsigned)a * (unsigned)n) & 3u) : 0);
        if (lane == 0) costs[(size_t)b * A + a] = c;
    }
    if (nbr) {
        uint32_t* N = nbr + (size_t)b * n * A;
        for (int k = lane; k < n; k += 32) {
            const int u = tv.at(k);
            const int pr = tv.at(k == 0 ? n - 1 : k - 1);
            const int su = tv.at(k == n - 1 ? 0 : k + 1);
            N[(size_t)u * A + a] = ((uint32_t)pr << 16) | (uint32_t)su;
        }
    }
}

// grid (n rows, B colonies); dynamic smem: A * (uint32 nbr + float w)
__global__ void __launch_bounds__(256) tsp_update_kernel(float* __restrict__ ph, const uint32_t* __restrict__ nbr,
                                                         const float* __restrict__ costs, int n, int A, float decay,
                                                         int elitist, int min_max, float ph_min,
                                                         const float* __restrict__ ph_max) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* nb_s = reinterpret_cast<uint32_t*>(smem);
    float* w_s = reinterpret_cast<float*>(smem) + A;
    __shared__ int best_ant;
    const int u = blockIdx.x, b = blockIdx.y;
    const uint32_t* N = nbr + ((size_t)b * n + u) * A;
    const float* C = costs + (size_t)b * A;
    for (int a = threadIdx.x; a < A; a += blockDim.x) {
        nb_s[a] = N[a];
        w_s[a] = __fdiv_rn(1.0f, C[a]);   // `1.0 / cost` = reciprocal(cost) * 1.0
    }
    if (elitist && threadIdx.x < 32) {
        // costs.min(dim=0): first index of the minimum
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
    float* row = ph + ((size_t)b * n + u) * n;
    const float hi = min_max ? ph_max[b] : 0.f;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        float val = __fmul_rn(row[v], decay);
        if (elitist) {
            const uint32_t e = nb_s[best_ant];
            const float w = w_s[best_ant];
            if ((int)(e >> 16) == v) val = __fadd_rn(val, w);
            if ((int)(e & 0xffffu) == v) val = __fadd_rn(val, w);
        } else {
            for (int a = 0; a < A; ++a) {
                const uint32_t e = nb_s[a];
                const float w = w_s[a];
                if ((int)(e >> 16) == v) val = __fadd_rn(val, w);       // statement 1: (path[k], path[k-1])
                if ((int)(e & 0xffffu) == v) val = __fadd_rn(val, w);   // statement 2: (path[k-1], path[k])
            }
        }
        if (min_max) {
            // ph[(ph > 1e-9) * ph < min] = min ; ph[ph > max] = max   (tsp/aco.py:117-118)
            const float gate = __fmul_rn(val > 1e-9f ? 1.0f : 0.0f, val);
            if (gate < ph_min) val = ph_min;
            if (val > hi) val = hi;
        }
        row[v] = val;
    }
}

}  // namespace deepaco

using namespace deepaco;

extern "C" int deepaco_tsp_cost(const float* distances, const int64_t* paths, const uint16_t* tours, int n, int n_ants,
                                int n_colonies, float* costs, uint32_t* neighbours, void* stream) {
    DACO_CHECK_ARG(distances && (costs || neighbours), "deepaco_tsp_cost: NULL distances / no output requested");
    DACO_CHECK_ARG((paths != nullptr) != (tours != nullptr), "deepaco_tsp_cost: pass exactly one of paths / tours");
    DACO_CHECK_ARG(n >= 2 && n <= 65535 && n_ants >= 1 && n_colonies >= 1, "deepaco_tsp_cost: bad sizes");
    const SumPlan sp = aten_sum_plan(n, n_ants);
    int bw = sp.block_width > 32 ? 32 : sp.block_width, lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    const int W = 8;
    dim3 grid((n_ants + W - 1) / W, n_colonies);
    tsp_cost_kernel<<<grid, W * 32, 0, (cudaStream_t)stream>>>(distances, paths, tours, n, n_ants, lbw, sp.vectorized,
                                                               costs, neighbours);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

extern "C" int deepaco_tsp_update(float* pheromone, const uint32_t* neighbours, const float* costs, int n, int n_ants,
                                  int n_colonies, float decay, int elitist, int min_max, float ph_min,
                                  const float* ph_max, void* stream) {
    DACO_CHECK_ARG(pheromone && neighbours && costs, "deepaco_tsp_update: NULL argument");
    DACO_CHECK_ARG(n >= 2 && n <= 65535 && n_ants >= 1 && n_colonies >= 1, "deepaco_tsp_update: bad sizes");
    DACO_CHECK_ARG(!min_max || ph_max, "deepaco_tsp_update: min_max needs ph_max");
    const size_t smem = (size_t)n_ants * 8;
    DACO_CHECK_ARG(smem <= 200 * 1024, "deepaco_tsp_update: n_ants=%d too large for one pass", n_ants);
    DACO_CHECK_CUDA(cudaFuncSetAttribute(tsp_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
    dim3 grid(n, n_colonies);
    tsp_update_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(pheromone, neighbours, costs, n, n_ants, decay, elitist,
                                                                     min_max, ph_min, ph_max);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
