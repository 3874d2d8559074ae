// Backward of ACO.sample()'s log-probabilities with respect to the heuristic matrix.
//
// The reference trains the heuristic network with REINFORCE through
//   log_probs[t, a] = log(clamp(p_c, eps, 1 - eps)),  p_c = x_c / sum_k x_k,  x_k = ph[u,k]^alpha * heu[u,k]^beta * mask_k
// (tsp/aco.py:165-177, tsp_nls/aco.py:205-211, cvrp/aco.py:167-174), autograd doing the rest.  Here the gradient is
// computed analytically by replaying every ant's path (the masks are a function of the path prefix):
//   d logp / d w[u,k] = [eps < p_c < 1 - eps] * ( delta_{k,c} / w[u,c]  -  mask_k * other[u,k] / S )
// for w = the (powered) heuristic and other = the (powered) pheromone -- and symmetrically for the pheromone.
//
// DETERMINISTIC, like the pheromone deposit (K2): no atomics.  Two kernels:
//   1. logp_backward_prepare_kernel, one warp per ant: replays the path and records, per step, the coefficient g / S
//      (0 where the clamp or a zero upstream gradient kills it), g itself, the remaining capacity and the depot rule
//      (CVRP), and per node the step at which the ant first ARRIVED there (`when`), which is all a later reader needs
//      to rebuild the visited mask of any step: node k is unvisited at step t  <=>  when[a][k] > t.
//   2. logp_backward_rows_kernel, one CTA per matrix row u: walks the ants in ascending order and, for every step an
//      ant spent at u (once for a TSP node or a customer; every depot visit for row 0 of CVRP), adds that step's
//      contribution to the row held in registers -- a fixed order per matrix element, so the result is bit-identical
//      run to run (the reference's autograd scatter is atomic and is not).
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
    float* g_heu;          // [n][n] accumulated into (caller zeroes)
    float* g_ph;           // [n][n] or null
    const float* demand;   // CVRP [n] or null (TSP)
    float capacity;
    int n, A, rows;
    // scratch written by the prepare kernel
    float* coef;           // [rows-1][A]  g / S, or 0 when the step passes no gradient
    float* gact;           // [rows-1][A]  g, or 0 when the step passes no gradient
    float* rem;            // [rows-1][A]  CVRP: remaining capacity before the step
    uint8_t* dok;          // [rows-1][A]  CVRP: depot admissible at the step
    uint16_t* when;        // [A][n]       step at which the ant first arrived at node k (0xffff: never); caller fills 0xff
    uint16_t* depot_steps; // [A][rows]    CVRP: steps spent at the depot (ascending)
    int32_t* depot_count;  // [A]
};

__global__ void __launch_bounds__(256) logp_backward_prepare_kernel(const BackwardParams p) {
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
    if (lane == 0) {
        vis[cur >> 5] |= 1u << (cur & 31);
        p.when[(size_t)a * n + cur] = 0;
    }
    __syncwarp();
    float used = cvrp ? p.demand[0] : 0.f;
    int left = n - 1;      // unvisited (customers for CVRP)
    int ndepot = 0;
    for (int t = 0; t + 1 < p.rows; ++t) {
        const int c = (int)p.paths[(size_t)(t + 1) * p.A + a];
        const float g = p.glogp[(size_t)t * p.A + a];
        const float* rph = p.ph + (size_t)cur * n;
        const float* rh = p.heu + (size_t)cur * n;
        const float remaining = p.capacity - used;
        const bool depot_ok = cur != 0 || left == 0;
        float s = 0.f;      // normaliser over the admissible set of this step
        for (int k = lane; k < n; k += 32) {
            bool ok = !((vis[k >> 5] >> (k & 31)) & 1u);
            if (cvrp) ok = (k == 0) ? depot_ok : (ok && !(p.demand[k] > remaining));
            if (ok) s += rph[k] * rh[k];
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(DACO_FULL, s, off);
        const float pc = rph[c] * rh[c] / s;
        const bool active = g != 0.f && pc > eps && pc < 1.0f - eps;     // clamp passes no gradient outside (eps, 1 - eps)
        if (lane == 0) {
            const size_t i = (size_t)t * p.A + a;
            p.coef[i] = active ? g / s : 0.f;
            p.gact[i] = active ? g : 0.f;
            if (cvrp) {
                p.rem[i] = remaining;
                p.dok[i] = depot_ok ? 1 : 0;
                if (cur == 0) p.depot_steps[(size_t)a * p.rows + ndepot] = (uint16_t)t;
            }
        }
        if (cvrp && cur == 0) ++ndepot;
        // advance the replay
        const bool fresh = !((vis[c >> 5] >> (c & 31)) & 1u);
        __syncwarp();
        if (cvrp) {
            if (c == 0) {
                used = p.demand[0];
            } else {
                used += p.demand[c];
                if (fresh) --left;
            }
        }
        if (lane == 0) {
            if (fresh) p.when[(size_t)a * n + c] = (uint16_t)(t + 1);
            vis[c >> 5] |= 1u << (c & 31);
        }
        __syncwarp();
        cur = c;
    }
    if (cvrp && lane == 0) p.depot_count[a] = ndepot;
}

// One CTA per matrix row u; thread i owns columns i, i + blockDim.x, ... (n <= 1024, at most 8 per thread at 128 threads).
__global__ void __launch_bounds__(128) logp_backward_rows_kernel(const BackwardParams p) {
    constexpr int KMAX = 8;
    const int u = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const int n = p.n, A = p.A;
    const bool cvrp = p.demand != nullptr;
    const float* rph = p.ph + (size_t)u * n;
    const float* rh = p.heu + (size_t)u * n;
    float acc_h[KMAX], acc_p[KMAX], vph[KMAX], vh[KMAX], dem[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        const int k = tid + j * nth;
        acc_h[j] = acc_p[j] = 0.f;
        vph[j] = k < n ? rph[k] : 0.f;
        vh[j] = k < n ? rh[k] : 0.f;
        dem[j] = (cvrp && k < n) ? p.demand[k] : 0.f;
    }
    for (int a = 0; a < A; ++a) {                      // ascending ants: the fixed accumulation order
        const uint16_t* W = p.when + (size_t)a * n;
        const int visits = (cvrp && u == 0) ? p.depot_count[a] : 1;
        for (int v = 0; v < visits; ++v) {
            const int t = (cvrp && u == 0) ? (int)p.depot_steps[(size_t)a * p.rows + v] : (int)W[u];
            if (t >= p.rows - 1) continue;             // 0xffff (never visited) or the final node: no move from here
            const size_t i = (size_t)t * A + a;
            const float gs = p.coef[i];
            if (gs == 0.f) continue;                   // uniform: the step passes no gradient
            const float g = p.gact[i];
            const int c = (int)p.paths[(size_t)(t + 1) * A + a];
            const float remaining = cvrp ? p.rem[i] : 0.f;
            const bool depot_ok = cvrp && p.dok[i] != 0;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                const int k = tid + j * nth;
                if (k < n) {
                    bool ok = (int)W[k] > t;
                    if (cvrp) ok = (k == 0) ? depot_ok : (ok && !(dem[j] > remaining));
                    if (ok) {
                        acc_h[j] += -gs * vph[j];
                        acc_p[j] += -gs * vh[j];
                    }
                    if (k == c) {
                        acc_h[j] += g / vh[j];
                        acc_p[j] += g / vph[j];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        const int k = tid + j * nth;
        if (k < n) {
            p.g_heu[(size_t)u * n + k] += acc_h[j];
            if (p.g_ph) p.g_ph[(size_t)u * n + k] += acc_p[j];
        }
    }
}

}  // namespace deepaco
