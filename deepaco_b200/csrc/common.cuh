// deepaco_b200 -- shared device helpers (sm_100a).
//
// Everything numerically significant in here exists to reproduce, bit for bit, what the reference
// gets out of ATen on the same GPU:
//   * Philox4x32-10 counter layout of torch's `distribution_elementwise_grid_stride_kernel`
//     (ATen/native/cuda/DistributionTemplates.h:50-89) used by `exponential_` and `randint`;
//   * the exponential transform of ATen/core/TransformationHelper.h:129-146 (CUDA branch);
//   * the summation ORDER of ATen's `reduce_kernel` for a sum over the last, contiguous dimension
//     (ATen/native/cuda/Reduce.cuh: thread_reduce_impl / input_vectorized_thread_reduce_impl /
//     block_x_reduce), so that `p / p.sum(-1)` and `sum(dist[u,v], 1)` round identically.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace deepaco {

#define DACO_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (same constants as curand_philox4x32_x.h; round keys hoisted by the caller)
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct PhiloxKey {
    uint32_t k0, k1;   // seed lo / hi
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
        const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += kPhiloxW0; k1 += kPhiloxW1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// torch draw geometry: `threads` = grid*256 of the ATen launch for a tensor of `numel` elements
// (calc_execution_policy).  Element li is produced by ATen thread (li % threads) on its
// (li / threads / 4)-th curand4 call, component (li / threads) % 4.
struct DrawGeom {
    uint32_t threads;   // grid.x * 256
    uint32_t single;    // 1 when numel <= threads (every element = component .x of call 0)
};

// raw 32-bit Philox word that ATen's kernel hands to the transform for linear element `li`
// of a draw launched at generator offset `offset` (a multiple of 4) with `seed`.
__device__ __forceinline__ uint32_t torch_philox_word(uint64_t seed, uint64_t offset, uint64_t li,
                                                      const DrawGeom& g) {
    uint64_t sub = li, call = 0;
    uint32_t comp = 0;
    if (!g.single) {
        sub = li % g.threads;
        const uint64_t chunk = li / g.threads;
        call = chunk >> 2;
        comp = (uint32_t)(chunk & 3);
    }
    const uint64_t ctr = (offset >> 2) + call;      // curand_init skipahead(offset) + call-th curand4
    const uint4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)sub, (uint32_t)(sub >> 32),
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    return comp == 0 ? r.x : (comp == 1 ? r.y : (comp == 2 ? r.z : r.w));
}

// curand_uniform (curand_uniform.h:69-72) followed by transformation::exponential, lambda = 1.
__device__ __forceinline__ float exp1_from_word(uint32_t x) {
    const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
    const float lg = (u >= 1.0f - 1.1920928955078125e-07f / 2.0f) ? -(1.1920928955078125e-07f / 2.0f) : logf(u);
    return (-1.0f / 1.0f) * lg;
}

// ---------------------------------------------------------------------------------------------
// ATen-ordered row sums.  The caller holds one row distributed over the 32 lanes of a warp.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_tree_sum(float v) {
    // block_x_reduce: value += shfl_down(value, off), off = 16..1.  IEEE add is commutative, so the
    // xor butterfly leaves lane 0's exact result in every lane.
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(DACO_FULL, v, off);
    return v;
}

// Strided layout (ATen non-vectorised path, row length < 128): lane l (< bw) holds x[l + bw*k].
// vt0 = 4 accumulators, element k goes to accumulator k & 3.  Lanes >= bw must hold zeros.
template <int EPL>
__device__ __forceinline__ float aten_sum_strided(const float (&x)[EPL]) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < EPL; ++k) acc[k & 3] = acc[k & 3] + x[k];
    return warp_tree_sum(((acc[0] + acc[1]) + acc[2]) + acc[3]);
}

// Vectorised layout (row length >= 128, multiple of 4, 16-byte aligned rows): lane l holds the float4
// at vector index l + 32*m in x[4m .. 4m+3]; accumulator i sums component i over m.
template <int EPL>
__device__ __forceinline__ float aten_sum_vec4(const float (&x)[EPL]) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int m = 0; m < EPL / 4; ++m) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = acc[i] + x[4 * m + i];
    }
    return warp_tree_sum(((acc[0] + acc[1]) + acc[2]) + acc[3]);
}

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) used to stage matrices / rows in smem
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// warp arg-max over non-negative floats with lowest-index tie break (ATen ArgMaxOps semantics):
// returns the winning index in every lane.
__device__ __forceinline__ uint32_t warp_argmax_nonneg(float v, uint32_t idx) {
    const uint32_t bits = __float_as_uint(v);
    const uint32_t top = __reduce_max_sync(DACO_FULL, bits);
    return __reduce_min_sync(DACO_FULL, bits == top ? idx : 0xffffffffu);
}

}  // namespace deepaco
