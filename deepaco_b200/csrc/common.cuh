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
#ifndef DEEPACO_CPU_EMU
#include <cuda_runtime.h>
// spelled through macros so that tests/cpu_emu can compile the kernel sources for the host (test infrastructure only;
// under nvcc they expand to exactly the tokens they replace)
#define DACO_NOINLINE __noinline__
#define DACO_DYN_SMEM128(name) extern __shared__ __align__(128) unsigned char name[]
#define DACO_DYN_SMEM16(name) extern __shared__ __align__(16) unsigned char name[]
#define DACO_STS_U8(addr, v) asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory")
#endif
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

// Fast path used by the sampling kernels when every element of the draw is component .x of call 0
// (numel <= grid*256, true for every configuration of BASELINE.json): the ten round keys are hoisted
// by the caller, the subsequence fits 32 bits and only output word .x is produced (38 instructions).
struct PhiloxRoundKeys {
    uint32_t a[10], b[10];
    __host__ __device__ __forceinline__ void init(uint64_t seed) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            a[r] = (uint32_t)seed + (uint32_t)r * kPhiloxW0;
            b[r] = (uint32_t)(seed >> 32) + (uint32_t)r * kPhiloxW1;
        }
    }
};

__device__ __forceinline__ uint32_t philox_word_x(uint32_t ctr_lo, uint32_t ctr_hi, uint32_t sub, const PhiloxRoundKeys& K) {
    uint32_t c0 = ctr_lo, c1 = ctr_hi, c2 = sub, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const uint32_t h0 = __umulhi(kPhiloxM0, c0), l0 = kPhiloxM0 * c0;
        const uint32_t h1 = __umulhi(kPhiloxM1, c2), l1 = kPhiloxM1 * c2;
        c0 = h1 ^ c1 ^ K.a[r];
        c2 = h0 ^ c3 ^ K.b[r];
        c1 = l1;
        c3 = l0;
    }
    return __umulhi(kPhiloxM1, c2) ^ c1 ^ K.a[9];
}

#ifndef DEEPACO_CPU_EMU   // MUFU-based pieces: tests/cpu_emu supplies libm stand-ins
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// curand_uniform (curand_uniform.h:69-72) followed by transformation::exponential, lambda = 1.
// at::log<float> on device is the fast `__logf` (ATen/NumericUtils.h:149-160) = lg2.approx.f32 * ln2.
// u >= 2^-33 is never denormal, so the .ftz form of the MUFU instruction returns the same bits without
// the denormal pre-scaling code the non-ftz intrinsic expands to.
__device__ __forceinline__ float exp1_from_word_guarded(uint32_t x) {      // literal form of the ATen transform
    const float u = fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);   // product is exact: fma == mul + add
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    const float lg = (u >= 1.0f - 1.1920928955078125e-07f / 2.0f) ? -(1.1920928955078125e-07f / 2.0f)
                                                                   : __fmul_rn(l2, 0.693147182464599609375f);
    return -lg;
}

// Same bits with one instruction less: the guard only fires for u == 1 (lg2 = 0 -> max picks 2^-24), and for the
// largest u below 1 (1 - 2^-24) MUFU.LG2 * ln2 is not below 2^-24 on sm_100 -- checked exhaustively over the top
// 2^20 Philox words by deepaco_debug_exp_guard (tests/test_gpu_probes.py); everywhere else -log(u) >= 2^-23.
__device__ __forceinline__ float exp1_from_word(uint32_t x) {
    const float u = fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    return fmaxf(-__fmul_rn(l2, 0.693147182464599609375f), 1.1920928955078125e-07f / 2.0f);
}

#endif  // DEEPACO_CPU_EMU

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

// Run-time-length version: ATen-ordered sum of f(k), k in [0, len), by one warp; result in every lane.
// `vec` selects the vectorised order (len >= 128), `lbw` = log2(block_width) for the strided order.
// `shift` = (element offset of the row start from a 16-byte boundary) & 3 -- ATen peels 4-shift head
// elements into accumulator 0 of lanes shift..3 before the aligned float4 loop
// (Reduce.cuh input_vectorized_thread_reduce_impl); rows of a fresh contiguous [rows][len] tensor have
// shift = (row * len) & 3.
template <typename F>
__device__ __forceinline__ float aten_row_sum_fn(F f, int len, int lbw, bool vec, int lane, int shift = 0) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
        int head = 0, end = len;
        if (shift > 0) {
            if (lane >= shift && lane < 4) acc[0] = __fadd_rn(acc[0], f(lane - shift));
            head = 4 - shift;
            end = len - head;
        }
        for (int idx = lane; idx * 4 + 3 < end; idx += 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = __fadd_rn(acc[i], f(head + 4 * idx + i));
        }
        const int t = end - (end & 3) + lane;                       // ATen tail -> accumulator 0
        if (t < end) acc[0] = __fadd_rn(acc[0], f(head + t));
    } else if (lane < (1 << lbw)) {
        int i = 0;
        for (int k = lane; k < len; k += (1 << lbw), ++i) acc[i & 3] = __fadd_rn(acc[i & 3], f(k));
    }
    return warp_tree_sum(__fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]));
}

// Device copy of host aten_sum_plan() for row lengths only known on the device (CVRP path length).
__device__ __forceinline__ void aten_sum_plan_dev(int row_len, int n_rows, int* lbw, int* vec) {
    int dim0 = row_len;
    *vec = row_len >= 128;
    if (*vec) dim0 >>= 2;
    auto lp2 = [](int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; };
    const int d0 = dim0 < 512 ? lp2(dim0) : 512;
    const int d1 = n_rows < 512 ? lp2(n_rows) : 512;
    int bw = d0 < 32 ? d0 : 32;
    const int bh = d1 < 512 / bw ? d1 : 512 / bw;
    bw = d0 < 512 / bh ? d0 : 512 / bh;
    if (bw > 32) bw = 32;   // wider blocks (n_rows < 16) are not reproduced exactly
    int l = 0;
    while ((1 << l) < bw) ++l;
    *lbw = l;
}

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) used to stage matrices / rows in smem
// ---------------------------------------------------------------------------------------------
#ifndef DEEPACO_CPU_EMU
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

#endif  // DEEPACO_CPU_EMU

// warp arg-max over non-negative floats with lowest-index tie break (ATen ArgMaxOps semantics):
// returns the winning index in every lane.
__device__ __forceinline__ uint32_t warp_argmax_nonneg(float v, uint32_t idx) {
    const uint32_t bits = __float_as_uint(v);
    const uint32_t top = __reduce_max_sync(DACO_FULL, bits);
    return __reduce_min_sync(DACO_FULL, bits == top ? idx : 0xffffffffu);
}

}  // namespace deepaco
