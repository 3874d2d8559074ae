// Roulette-wheel tour construction: the sampler of the reference's INFERENCE path (tsp_nls/aco.py:260-297
// `_inference_sample` / `inference_batch_sample`, used by `sample(inference=True)` :81-85 and `run(.., inference=True)`
// :106-110): per step  prob = probmat[last] * mask;  rand = U[0,1) * sum(prob);  next = first k with
// prob[0] + .. + prob[k] >= rand.  The reference draws U from numba's private generator, which nothing outside numba
// can reproduce, so parity with it is statistical (tests/test_gpu_roulette.py: chi-square of the first-step and
// pair-transition frequencies against `inference_batch_sample` on the same probmat); here U comes from Philox4x32-10 at
// the caller's (seed, offset), one word per (ant, step).
//
// One warp per ant.  Lane l holds columns l, l + 32, ... of the current row (registers), the running sum over the row
// is a warp scan per 32 columns carried across chunks in column order -- the same left-to-right order in which the
// reference subtracts prob[k] from rand -- and the pick is a ballot over "prefix >= rand and prob > 0".
#include "common.cuh"
#include "host_util.h"

namespace deepaco {

struct RouletteParams {
    const float* prob;      // [B][n][n]  pheromone^alpha (.) heuristic^beta
    int n, A, B;
    int start_node;         // >= 0 fixed start (the reference passes 0); -1: uniform random start per ant
    uint64_t seed, offset;
    const uint64_t* offsets;
    uint16_t* tours;        // [B][A][n] or null
    int64_t* paths;         // [B][n][A] or null
};

// word `step` of ant `a`: Philox counter offset/4 + step/4, subsequence a, output word step%4
__device__ __forceinline__ uint32_t roulette_word(uint64_t seed, uint64_t offset, uint32_t a, uint32_t step) {
    const uint64_t ctr = (offset >> 2) + (step >> 2);
    const uint4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), a, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t c = step & 3u;
    return c == 0 ? r.x : (c == 1 ? r.y : (c == 2 ? r.z : r.w));
}

template <int EPL>
__global__ void __launch_bounds__(128) aco_roulette_kernel(const RouletteParams p) {
    const int n = p.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int a = blockIdx.x * W + warp, b = blockIdx.y;
    if (a >= p.A) return;
    const float* P = p.prob + (size_t)b * n * n;
    const uint64_t off = (p.offsets ? p.offsets[b] : 0ull) + p.offset;
    uint32_t alive = 0;                                   // bit k: column lane + 32 k unvisited
#pragma unroll
    for (int k = 0; k < EPL; ++k)
        if (lane + 32 * k < n) alive |= 1u << k;
    int cur = p.start_node >= 0 ? p.start_node : (int)(roulette_word(p.seed, off, (uint32_t)a, 0u) % (uint32_t)n);
    uint16_t* tour = p.tours ? p.tours + ((size_t)b * p.A + a) * n : nullptr;
    int64_t* path = p.paths ? p.paths + (size_t)b * n * p.A + a : nullptr;
    for (int step = 1;; ++step) {
        if ((cur & 31) == lane) alive &= ~(1u << (cur >> 5));
        if (lane == 0) {
            if (tour) tour[step - 1] = (uint16_t)cur;
            if (path) path[(size_t)(step - 1) * p.A] = cur;
        }
        if (step == n) break;
        const float* row = P + (size_t)cur * n;
        float x[EPL], pre[EPL];
        float carry = 0.f;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int j = lane + 32 * k;
            x[k] = ((alive >> k) & 1u) ? __ldg(row + j) : 0.f;
            float s = x[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float y = __shfl_up_sync(DACO_FULL, s, o);
                if (lane >= o) s += y;
            }
            pre[k] = carry + s;
            carry = __shfl_sync(DACO_FULL, pre[k], 31);
        }
        // U in (0, 1): never 0, so a column with prob 0 (visited) is never picked while any mass is left
        const float u = ((float)(roulette_word(p.seed, off, (uint32_t)a, (uint32_t)step) >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float rnd = u * carry;
        int pick = -1, last_alive = -1;
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const uint32_t hit = __ballot_sync(DACO_FULL, x[k] > 0.f && pre[k] >= rnd);
            const uint32_t any = __ballot_sync(DACO_FULL, (alive >> k) & 1u);
            if (pick < 0 && hit) pick = 32 * k + __ffs(hit) - 1;
            if (any) last_alive = 32 * k + 31 - __clz(any);
        }
        cur = pick >= 0 ? pick : last_alive;            // rounding left rnd above the total (or an all-zero row): last column
    }
}

}  // namespace deepaco

using namespace deepaco;

extern "C" uint64_t deepaco_tsp_roulette_offset_increment(int n, int n_ants) {
    (void)n_ants;
    return n < 1 ? 0 : 4ull * (uint64_t)((n + 3) / 4);   // one Philox word per (ant, step): ceil(n / 4) counters per tour
}

extern "C" int deepaco_tsp_roulette_sample(const float* prob, int n, int n_ants, int n_colonies, int start_node, uint64_t seed,
                                           uint64_t offset, const uint64_t* offsets, uint16_t* tours, int64_t* paths, void* stream) {
    DACO_CHECK_ARG(prob && (tours || paths), "deepaco_tsp_roulette_sample: NULL prob / no output requested");
    DACO_CHECK_ARG(n >= 2 && n <= DEEPACO_MAX_NODES && n_ants >= 1 && n_colonies >= 1 && n_colonies <= 65535 && start_node < n,
                   "deepaco_tsp_roulette_sample: bad sizes (n=%d, n_ants=%d, start=%d)", n, n_ants, start_node);
    RouletteParams p{prob, n, n_ants, n_colonies, start_node, seed, offset, offsets, tours, paths};
    const int W = 4;
    dim3 grid((n_ants + W - 1) / W, n_colonies);
    cudaStream_t st = (cudaStream_t)stream;
    const int epl = (n + 31) / 32;
    if (epl <= 4) aco_roulette_kernel<4><<<grid, W * 32, 0, st>>>(p);
    else if (epl <= 8) aco_roulette_kernel<8><<<grid, W * 32, 0, st>>>(p);
    else if (epl <= 16) aco_roulette_kernel<16><<<grid, W * 32, 0, st>>>(p);
    else aco_roulette_kernel<32><<<grid, W * 32, 0, st>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
