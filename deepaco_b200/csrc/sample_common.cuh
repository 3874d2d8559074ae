// Pieces shared by the TSP and CVRP construction kernels.
#pragma once
#include "common.cuh"

namespace deepaco {

// Stage P = pheromone (.) heuristic of colony b into shared memory (TMA bulk copy + mbarrier).
__device__ __forceinline__ void stage_product(float* Psm, const float* ph, const float* heu, int n, int b, uint64_t* bar) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const size_t nn = (size_t)n * n;
    const float* src = ph + (size_t)b * nn;
    const uint32_t total = (uint32_t)(nn * 4);
    const uint32_t bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (total & ~15u) : 0u;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0 && bulk) {
        mbar_expect_tx(bar, bulk);
        constexpr uint32_t kChunk = 32768;
        for (uint32_t off = 0; off < bulk; off += kChunk) {
            const uint32_t sz = (bulk - off < kChunk) ? (bulk - off) : kChunk;
            tma_bulk_g2s(reinterpret_cast<char*>(Psm) + off, reinterpret_cast<const char*>(src) + off, sz, bar);
        }
    }
    for (size_t i = bulk / 4 + tid; i < nn; i += nthreads) Psm[i] = src[i];
    if (bulk) mbar_wait(bar, 0);
    __syncthreads();
    if (heu) {
        const float* h = heu + (size_t)b * nn;
        if (((reinterpret_cast<uintptr_t>(h) & 15) == 0) && (nn & 3) == 0) {
            // float4 loads, four in flight per thread, so the L2 latency of the heuristic read is overlapped
            const float4* h4 = reinterpret_cast<const float4*>(h);
            float4* P4 = reinterpret_cast<float4*>(Psm);
            const int n4 = (int)(nn >> 2);
            for (int i = tid; i < n4; i += 4 * nthreads) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u * nthreads < n4) v[u] = __ldg(h4 + i + u * nthreads);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u * nthreads < n4) {
                        float4 q = P4[i + u * nthreads];
                        q.x = __fmul_rn(q.x, v[u].x);
                        q.y = __fmul_rn(q.y, v[u].y);
                        q.z = __fmul_rn(q.z, v[u].z);
                        q.w = __fmul_rn(q.w, v[u].w);
                        P4[i + u * nthreads] = q;
                    }
            }
        } else {
            for (size_t i = tid; i < nn; i += nthreads) Psm[i] = __fmul_rn(Psm[i], __ldg(h + i));
        }
        __syncthreads();
    }
}

// One construction step with the reference's exact arithmetic (ATen summation order, IEEE divisions,
// lowest-index tie break).  `alive` is a 1024-bit map of unvisited nodes.  Rarely executed from the list
// kernel (near-ties), so it is kept out of line.  Returns the chosen node in every lane; *pn_out receives
// the normalised probability of the chosen node (valid in the lane that owns it, broadcast by caller).
static __device__ DACO_NOINLINE uint32_t exact_step(const float* row, const uint32_t* alive, int n, int lbw, int vec,
                                            int double_norm, const float* nz, uint64_t seed, uint64_t off_step,
                                            uint32_t sub_base, DrawGeom g, float* pn_out) {
    const int lane = threadIdx.x & 31;
    const int shift = vec ? (int)(sub_base & 3u) : 0;   // sub_base = ant * n = element offset of this probs row
    auto xval = [&](int k) -> float { return ((alive[k >> 5] >> (k & 31)) & 1u) ? row[k] : 0.f; };
    float S = aten_row_sum_fn(xval, n, lbw, vec != 0, lane, shift);
    float S2 = 1.f;
    if (double_norm) S2 = aten_row_sum_fn([&](int k) { return __fdiv_rn(xval(k), S); }, n, lbw, vec != 0, lane, shift);
    float best = 0.f, bestp = 0.f;
    uint32_t bestj = 0xffffffffu;
    for (int k = lane; k < n; k += 32) {
        float pn = __fdiv_rn(xval(k), S);
        if (double_norm) pn = __fdiv_rn(pn, S2);
        const float q = nz ? nz[k] : exp1_from_word(torch_philox_word(seed, off_step, (uint64_t)sub_base + k, g));
        const float v = __fdiv_rn(pn, q);
        if (bestj == 0xffffffffu || v > best) {
            best = v;
            bestj = k;
            bestp = pn;
        }
    }
    const uint32_t jstar = warp_argmax_nonneg(best, bestj);
    const uint32_t owner = __ffs(__ballot_sync(DACO_FULL, bestj == jstar)) - 1;
    *pn_out = __shfl_sync(DACO_FULL, bestp, owner);
    return jstar;
}

}  // namespace deepaco
