// CVRP tour cost, neighbour table and pheromone update kernels (reference cvrp/aco.py:106-136); launchers and the C ABI
// are in cvrp.cu.  Plain CUDA C++ plus warp intrinsics, so tests/cpu_emu compiles the same text for the host.
#pragma once
#include "common.cuh"

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
    DACO_DYN_SMEM16(smem);
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
