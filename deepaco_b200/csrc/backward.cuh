// Backward of ACO.sample()'s log-probabilities with respect to the heuristic matrix.
//
// The reference trains the heuristic network with REINFORCE through
//   log_probs[t, a] = log(clamp(p_c, eps, 1 - eps)),  p_c = x_c / sum_k x_k,  x_k = ph[u,k]^alpha * heu[u,k]^beta * mask_k
// (tsp/aco.py:165-177, tsp_nls/aco.py:205-211, cvrp/aco.py:167-174), autograd doing the rest.  Here the gradient is
// computed analytically by replaying every ant's path (the masks are a function of the path prefix):
//   d logp / d w[u,k] = [eps < p_c < 1 - eps] * ( delta_{k,c} / w[u,c]  -  mask_k * other[u,k] / S )
// for w = the (powered) heuristic and other = the (powered) pheromone -- and symmetrically for the pheromone.
// One warp per ant; contributions are accumulated with fp32 atomics (the reference's autograd scatter is atomic,
// too), so the result matches autograd to fp32 rounding, not bit for bit.
//
// Kernel source only (the C ABI is in backward.cu); plain CUDA C++ plus warp shuffles, so tests/cpu_emu compiles the
// same text for the host.
#pragma once
#include "common.cuh"

namespace deepaco {

struct BackwardParams {
    const float* ph;       // [n][n]  pheromone ** alpha
    const float* heu;      // [n][n]  heuristic ** beta
    const int64_t* paths;  // [rows][A]
    const float* glogp;    // [rows-1][A] upstream gradient
    float* g_heu;          // [n][n] accumulated (caller zeroes)
    float* g_ph;           // [n][n] or null
    const float* demand;   // CVRP [n] or null (TSP)
    float capacity;
    int n, A, rows;
};

__global__ void __launch_bounds__(256) logp_backward_kernel(const BackwardParams p) {
    __shared__ uint32_t vis_all[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + warp;
    if (a >= p.A) return;
    uint32_t* vis = vis_all[warp];
    const int n = p.n;
    const bool cvrp = p.demand != nullptr;
    const float eps = 1.1920928955078125e-07f;
    vis[lane] = 0u;
    __syncwarp();
    int cur = (int)p.paths[a];
    if (lane == 0) vis[cur >> 5] |= 1u << (cur & 31);
    __syncwarp();
    float used = cvrp ? p.demand[0] : 0.f;
    int left = cvrp ? n - 1 : n - 1;   // unvisited (customers for CVRP)
    for (int t = 0; t + 1 < p.rows; ++t) {
        const int c = (int)p.paths[(size_t)(t + 1) * p.A + a];
        const float g = p.glogp[(size_t)t * p.A + a];
        const float* rph = p.ph + (size_t)cur * n;
        const float* rh = p.heu + (size_t)cur * n;
        const float remaining = p.capacity - used;
        // admissible set of this step
        float s = 0.f;
        for (int k = lane; k < n; k += 32) {
            bool ok = !((vis[k >> 5] >> (k & 31)) & 1u);
            if (cvrp) ok = (k == 0) ? (cur != 0 || left == 0) : (ok && !(p.demand[k] > remaining));
            if (ok) s += rph[k] * rh[k];
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(DACO_FULL, s, off);
        const float xc = rph[c] * rh[c];
        const float pc = xc / s;
        if (g != 0.f && pc > eps && pc < 1.0f - eps) {     // clamp passes no gradient outside (eps, 1 - eps)
            const float gs = g / s;
            for (int k = lane; k < n; k += 32) {
                bool ok = !((vis[k >> 5] >> (k & 31)) & 1u);
                if (cvrp) ok = (k == 0) ? (cur != 0 || left == 0) : (ok && !(p.demand[k] > remaining));
                if (ok) {
                    atomicAdd(p.g_heu + (size_t)cur * n + k, -gs * rph[k]);
                    if (p.g_ph) atomicAdd(p.g_ph + (size_t)cur * n + k, -gs * rh[k]);
                }
            }
            if (lane == 0) {
                atomicAdd(p.g_heu + (size_t)cur * n + c, g / rh[c]);
                if (p.g_ph) atomicAdd(p.g_ph + (size_t)cur * n + c, g / rph[c]);
            }
        }
        __syncwarp();
        // advance the replay
        if (cvrp) {
            if (c == 0) {
                used = p.demand[0];
            } else {
                used += p.demand[c];
                if (!((vis[c >> 5] >> (c & 31)) & 1u)) --left;
            }
        }
        if (lane == 0) vis[c >> 5] |= 1u << (c & 31);
        __syncwarp();
        cur = c;
    }
}

}  // namespace deepaco
