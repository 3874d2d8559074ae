// One construction step with caller-supplied masks: ACO.pick_move of the reference (tsp/aco.py:165-177,
// cvrp/aco.py:167-174).  The fused construction kernels (list_kernel.cuh) never call this -- they keep the masks in
// registers -- but pick_move is part of the class surface, so code that drives the construction step by step (a custom
// gen_path with its own feasibility masks) gets the same draw the reference would make on this GPU:
//
//   x      = ((pheromone**alpha)[prev] * (heuristic**beta)[prev]) * mask (* capacity_mask)      fp32, left to right
//   probs  = x / x.sum(-1)                        Categorical.__init__; the sum in ATen's order (common.cuh)
//   action = argmax(probs / q), q = exponential_(1) over the [n_ants, n] tensor at the generator offset
//            (torch.multinomial's one-sample path; lowest index on ties)
//   logp   = log(clamp(probs, eps, 1 - eps))[action]
//
// One warp per ant; the two matrix rows and the mask rows are read once from global memory, coalesced.
//
// Kernel source only (the C ABI is in pick_move.cu); plain CUDA C++ plus the helpers of common.cuh, so tests/cpu_emu
// compiles the same text for the host.
#pragma once
#include "common.cuh"

namespace deepaco {

struct PickMoveParams {
    const float* php;      // [n][n] pheromone ** alpha
    const float* heup;     // [n][n] heuristic ** beta, or null (ones)
    const int64_t* prev;   // [A]
    const float* mask;     // [A][n]
    const float* mask2;    // [A][n] or null
    int n, A, lbw, vec;
    uint64_t seed, offset;
    DrawGeom g;
    int64_t* actions;      // [A]
    float* logp;           // [A] or null
    int* bad_prev;         // set to 1 if a prev index is outside [0, n)
};

__global__ void __launch_bounds__(256) pick_move_kernel(PickMoveParams p) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= p.A) return;
    const int n = p.n;
    const int64_t u = p.prev[a];
    if (u < 0 || u >= n) {
        if (lane == 0) { *p.bad_prev = 1; p.actions[a] = 0; if (p.logp) p.logp[a] = 0.f; }
        return;
    }
    const float* pr = p.php + (size_t)u * n;
    const float* hr = p.heup ? p.heup + (size_t)u * n : nullptr;
    const float* m1 = p.mask + (size_t)a * n;
    const float* m2 = p.mask2 ? p.mask2 + (size_t)a * n : nullptr;
    auto xval = [&](int k) -> float {
        float x = hr ? __fmul_rn(__ldg(pr + k), __ldg(hr + k)) : __ldg(pr + k);
        x = __fmul_rn(x, __ldg(m1 + k));
        if (m2) x = __fmul_rn(x, __ldg(m2 + k));
        return x;
    };
    const uint64_t base = (uint64_t)a * n;                    // element offset of this row in the [A, n] tensors
    const int shift = p.vec ? (int)(base & 3u) : 0;
    const float S = aten_row_sum_fn(xval, n, p.lbw, p.vec != 0, lane, shift);
    float best = 0.f, bestp = 0.f;
    uint32_t bestj = 0xffffffffu;
    for (int k = lane; k < n; k += 32) {
        const float pn = __fdiv_rn(xval(k), S);
        const float q = exp1_from_word(torch_philox_word(p.seed, p.offset, base + k, p.g));
        const float v = __fdiv_rn(pn, q);
        if (bestj == 0xffffffffu || v > best) {
            best = v;
            bestj = k;
            bestp = pn;
        }
    }
    const uint32_t jstar = warp_argmax_nonneg(best, bestj);
    if (bestj == jstar) {                                      // exactly one lane owns the winner
        p.actions[a] = (int64_t)jstar;
        if (p.logp) {
            const float eps = 1.1920928955078125e-07f;         // torch.finfo(float32).eps
            p.logp[a] = logf(fminf(fmaxf(bestp, eps), 1.0f - eps));
        }
    }
}

}  // namespace deepaco
