// K4 -- batched 2-opt and NLS local search (reference tsp_nls/two_opt.py:6-49 and tsp_nls/aco.py:234-258).
//
// One CTA per ant tour; the whole local search of that ant (every pass of every 2-opt call, and for NLS
// all T_nls perturbation rounds) runs inside one launch -- ants never interact.
//
// two_opt_once (two_opt.py:6-28): scan all 1 <= i < j <= n-1, change = d[p,nj] + d[ni,nx] - d[p,ni] - d[nj,nx]
// (fp32, left to right), keep the first strict minimum in (i,j) order, reverse tour[i..j] iff it is
// < -1e-6.  Here warp w owns a contiguous band of i (balanced by pair count); for each i it stages the two
// distance rows d[p,:] and d[ni,:] in shared memory with coalesced loads (the row of ni is reused as the
// row of p for i+1), lanes sweep j, and the (change, i*n+j) minimum is reduced over the CTA with the same
// tie rule.  The terms d[p,ni] and d[nj,nx] are tour edges, cached per pass.
//
// n <= 510 runs two_opt_call_v2: with G(a, m) = d[tour[a], tour[m]] the candidate is
// change(i, j) = G(i-1, j) + G(i, j+1) - edge[i] - edge[j+1], so the row of gathers made for i is reused as the first
// term of i+1.  Lane l owns the tour positions m = 32k + l for the whole pass (node ids and edge[m+1] in registers),
// gathers g_i[m] once per row from the staged distance row, keeps g_{i-1}[m] in registers and takes g_i[m+1] from its
// neighbour lane by shuffle: one shared-memory gather per candidate instead of two plus three index / edge loads.
// Rows are dealt to warps in mirrored pairs (i with n-1-i), so every warp gets the same number of candidates AND the
// same number of row fetches (contiguous bands by candidate count leave the last warp with n/4 nearly empty rows, each
// a full L2 round trip).
//
// Kernel source only (launchers and the C ABI are in two_opt.cu): apart from the helpers in the first #ifndef block it
// uses plain CUDA C++ (thread indices, __syncthreads / __syncwarp, warp shuffles, mbarrier / TMA helpers of common.cuh),
// so tests/cpu_emu compiles the same text for the host.
#pragma once
#ifndef DEEPACO_CPU_EMU
#include "common.cuh"
#endif

namespace deepaco {

#ifndef DEEPACO_CPU_EMU
// the handful of device-only constructs of this file; tests/cpu_emu/cuda_emu.h supplies host stand-ins so the same
// kernel source runs in the CPU test suite (test infrastructure only)
__device__ __forceinline__ void cp_async_16(float* dst_smem, const float* src) {     // LDGSTS.128
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// shared-window loads by 32-bit address (the generic pointers inside TwoOptShared cost a 64-bit add per access)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
#define DACO_2OPT_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// numpy float32 pairwise summation of one contiguous row (numpy/_core/src/umath/loops_utils.h.src
// @TYPE@_pairwise_sum, PW_BLOCKSIZE = 128) -- tsp_nls/aco.py:171-182 compares tours by np.sum(dist[u, v], axis=1).
template <typename F>
__device__ float numpy_pairwise_sum(F f, int lo, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, f(lo + i));
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = f(lo + k);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], f(lo + i + k));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, f(lo + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(numpy_pairwise_sum(f, lo, n2), numpy_pairwise_sum(f, lo + n2, n - n2));
}

struct TwoOptShared {
    uint64_t* bars;   // [W][3]  mbarriers of the row buffers (two_opt_call_v2, TMA rows)
    uint32_t* phase;  // [W]     their parities, carried from call to call
    uint16_t* tour;   // [n + 1] tour[n] mirrors tour[0]
    float* edge;      // [n]  edge[k] = d[tour[k-1], tour[k]], edge[0] = d[tour[n-1], tour[0]]
    float* rows;      // [W][3][n]  triple-buffered distance rows (cp.async prefetch of the next row)
    float* red_c;     // [W]
    uint32_t* red_k;  // [W]
    int* band;        // [W+1]
};

// one 2-opt call: up to max_iterations passes on the tour in shared memory; returns passes done
__device__ DACO_NOINLINE int two_opt_call(const float* __restrict__ D, int n, int max_iterations, const TwoOptShared& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    float* rowbuf = S.rows + (size_t)warp * 3 * n;
    const bool vec16 = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    // asynchronous global -> shared copy of one distance row (LDGSTS); completion via cp.async groups
    auto prefetch_row = [&](float* dst, int node) {
        const float* src = D + (size_t)node * n;
        if (vec16) {
            for (int c = lane * 4; c < n; c += 128) cp_async_16(dst + c, src + c);
        } else {
            for (int c = lane; c < n; c += 32) cp_async_4(dst + c, src + c);
        }
        cp_async_commit();
    };
    int it = 0;
    while (it < max_iterations) {
        __syncthreads();
        for (int k = tid; k < n; k += blockDim.x) {
            const int a = S.tour[k == 0 ? n - 1 : k - 1], b = S.tour[k];
            S.edge[k] = __ldg(D + (size_t)a * n + b);
        }
        __syncthreads();
        float best = 0.f;                 // delta starts at 0 (two_opt.py:10)
        uint32_t bestkey = 0xffffffffu;
        const int lo = S.band[warp], hi = S.band[warp + 1];
        if (lo < hi) {
            prefetch_row(rowbuf, S.tour[lo - 1]);          // d[tour[i-1], :] of the first i
            prefetch_row(rowbuf + n, S.tour[lo]);          // d[tour[i], :]
        }
        for (int i = lo; i < hi; ++i) {
            const int slot = (i - lo) % 3;
            const float* rp = rowbuf + (size_t)slot * n;               // d[tour[i-1], :]
            const float* ri = rowbuf + (size_t)((slot + 1) % 3) * n;   // d[tour[i], :]
            cp_async_wait<0>();
            __syncwarp();
            if (i + 1 < hi) prefetch_row(rowbuf + (size_t)((slot + 2) % 3) * n, S.tour[i + 1]);   // overlaps the sweep below
            const int ni = S.tour[i];
            const int p = S.tour[i - 1];
            const float e_i = S.edge[i];
            for (int j = i + 1 + lane; j < n; j += 32) {
                const int nj = S.tour[j];
                const int jn = (j + 1 == n) ? 0 : j + 1;
                const int nx = S.tour[jn];
                if (p == nj || nx == ni) continue;
                const float change = __fsub_rn(__fsub_rn(__fadd_rn(rp[nj], ri[nx]), e_i), S.edge[jn]);
                if (change < best) {      // strict: first minimum in (i, j) order within this lane's sweep
                    best = change;
                    bestkey = (uint32_t)i * (uint32_t)n + (uint32_t)j;
                }
            }
            __syncwarp();
        }
        // CTA arg-min with lowest key on ties (== first strict minimum of the sequential scan)
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(DACO_FULL, best, off);
            const uint32_t ok = __shfl_xor_sync(DACO_FULL, bestkey, off);
            if (oc < best || (oc == best && ok < bestkey)) { best = oc; bestkey = ok; }
        }
        if (lane == 0) { S.red_c[warp] = best; S.red_k[warp] = bestkey; }
        __syncthreads();
        best = S.red_c[0];
        bestkey = S.red_k[0];
        for (int w = 1; w < W; ++w) {
            const float oc = S.red_c[w];
            const uint32_t ok = S.red_k[w];
            if (oc < best || (oc == best && ok < bestkey)) { best = oc; bestkey = ok; }
        }
        ++it;
        if (!((double)best < -1e-6)) break;   // two_opt.py:24,36
        const int i = (int)(bestkey / (uint32_t)n), j = (int)(bestkey % (uint32_t)n);
        __syncthreads();
        for (int k = tid; k < (j - i + 1) / 2; k += blockDim.x) {
            const uint16_t x = S.tour[i + k];
            S.tour[i + k] = S.tour[j - k];
            S.tour[j - k] = x;
        }
    }
    __syncthreads();
    return it;
}

// same contract as two_opt_call for a PERMUTATION tour with n + 1 <= 32 * KMAX (see the header comment).  For a
// permutation the reference's skip test (node_prev == node_j or node_next == node_i, two_opt.py:16) can never fire
// for 1 <= i < j <= n-1, so it is not evaluated here; the kernel checks the tour once and sends anything else to
// two_opt_call.
struct TwoOptBest {   // two running (change, key) minima per lane (even / odd k: two short dependency chains)
    float c0, c1;
    uint32_t k0, k1;
};

// one row r of the sweep: cur[k] = g_r[m] for the positions m = 32k + lane >= r+1, then the candidates (i = r, j = m)
// from prev = g_{r-1}.  Steps come in blocks of four k (straight-line inside a block, so four gathers / shuffles /
// compare chains overlap); blocks entirely below k0 = (r+1)/32 are skipped with one warp-uniform branch.  One code
// copy serves every row -- a variant per k0 is faster per row but sixteen warps in sixteen variants thrash the
// instruction cache.  e_i = -inf turns the row into a pure gather (first row of a run).
template <int KMAX>
__device__ __forceinline__ void two_opt_row(uint32_t row, int r, uint32_t keybase, int lane, const uint32_t (&off)[KMAX / 2],
                                            const float (&en)[KMAX], const float (&prev)[KMAX + 1], float (&cur)[KMAX + 1],
                                            float e_i, TwoOptBest& B) {
    const int k0 = (r + 1) >> 5;
#pragma unroll
    for (int kb = 0; kb < KMAX; kb += 4) {
        if (kb + 3 >= k0) {
#pragma unroll
            for (int k = kb; k < kb + 4; ++k) cur[k] = lds_f32(row + ((k & 1) ? off[k / 2] >> 16 : off[k / 2] & 0xffffu));
        }
    }
    const int next = (lane + 1) & 31;
    const bool first = lane == 0;
    const int rl = r - lane;                                      // 32k + lane > r  <=>  32k > rl
#pragma unroll
    for (int kb = 0; kb < KMAX; kb += 4) {
        if (kb + 3 >= k0) {
            float nb[4];
#pragma unroll
            for (int k = kb; k < kb + 4; ++k) nb[k - kb] = __shfl_sync(DACO_FULL, first ? cur[k + 1] : cur[k], next);   // g_r[m+1]
#pragma unroll
            for (int k = kb; k < kb + 4; ++k) {
                const float change = __fsub_rn(__fsub_rn(__fadd_rn(prev[k], nb[k - kb]), e_i), en[k]);
                // strict <: first minimum in (i, j) order
                const bool take = (32 * k > rl) & (change < ((k & 1) ? B.c1 : B.c0));
                if (k & 1) {
                    B.c1 = take ? change : B.c1;
                    B.k1 = take ? keybase + 32u * k : B.k1;
                } else {
                    B.c0 = take ? change : B.c0;
                    B.k0 = take ? keybase + 32u * k : B.k0;
                }
            }
        }
    }
}

// 4-byte cp.async copy of one row (n % 4 != 0 or unaligned matrices); out of line: not on the TMA path
__device__ DACO_NOINLINE void two_opt_row_copy_async(float* dst, const float* src, int n, int lane) {
    for (int c = lane; c < n; c += 32) cp_async_4(dst + c, src + c);
    cp_async_commit();
}

template <int KMAX>
__device__ DACO_NOINLINE int two_opt_call_v2(const float* __restrict__ D, int n, int max_iterations, const TwoOptShared& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    uint16_t* const tour = S.tour;
    float* const edge = S.edge;
    float* const rowbuf = S.rows + (size_t)warp * 3 * n;
    uint64_t* const bars = S.bars + warp * 3;
    const uint32_t rows_s = smem_u32(rowbuf), tour_s = smem_u32(tour), edge_s = smem_u32(edge);
    // distance rows arrive by TMA bulk copy (one instruction per row, completion on an mbarrier per buffer) when the
    // rows are 16-byte granular, else by 4-byte cp.async
    const bool tma = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
    uint32_t ph = S.phase[warp];          // parity of the next completion of each of the three row buffers
    auto issue = [&](int slot, int node) {
        float* dst = rowbuf + slot * n;
        const float* src = D + (size_t)node * n;
        if (tma) {
            if (lane == 0) {
                mbar_expect_tx(bars + slot, 4u * n);
                tma_bulk_g2s(dst, src, 4u * n, bars + slot);
            }
        } else {
            two_opt_row_copy_async(dst, src, n, lane);
        }
    };
    // rows 1 .. n-2 in mirrored pairs: row i has n-1-i candidates, row n-1-i has i
    const int F = (n - 2) / 2;
    const int lo1 = 1 + F * warp / W;
    int hi1 = 1 + F * (warp + 1) / W;
    const int lo2 = n - hi1, hi2 = n - lo1;
    if (warp == W - 1 && ((n - 2) & 1)) hi1 = F + 2;   // the unpaired middle row
    const float ninf = __int_as_float(0xff800000);
    int it = 0;
    while (it < max_iterations) {
        __syncthreads();
        for (int k = tid + 1; k <= n; k += blockDim.x)   // tour[n] == tour[0]: edge[n] closes the tour
            edge[k] = __ldg(D + (size_t)tour[k - 1] * n + tour[k]);
        __syncthreads();
        uint32_t off[KMAX / 2];  // byte offsets of tour[m] in a distance row, m = 32k + lane, two per register
        float en[KMAX];          // edge[m+1]; -inf where m is not a candidate position (change becomes +inf)
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            const int m = 32 * k + lane;
            const uint32_t o = m <= n ? 4u * tour[m] : 0u;
            off[k / 2] = (k & 1) ? off[k / 2] | (o << 16) : o;
            en[k] = m <= n - 1 ? edge[m + 1] : ninf;
        }
        TwoOptBest B = {0.f, 0.f, 0xffffffffu, 0xffffffffu};   // delta starts at 0 (two_opt.py:10)
#pragma unroll 1
        for (int run = 0; run < 2; ++run) {
            const int lo = run ? lo2 : lo1, hi = run ? hi2 : hi1;
            if (lo >= hi) continue;
            float ga[KMAX + 1], gb[KMAX + 1];   // g of the previous row (ga on entry of a row pair) and of the current one
#pragma unroll
            for (int k = 0; k <= KMAX; ++k) ga[k] = gb[k] = 0.f;
            int slot = 0;                       // buffer of row r; r+1 is in flight in slot+1, r+2 goes to slot+2 (mod 3)
            issue(0, tour[lo - 1]);
            issue(1, tour[lo]);
            auto stage = [&](int r) -> uint32_t {         // row r landed and visible; row r+2 on its way
                if (tma) {
                    mbar_wait(bars + slot, (ph >> slot) & 1u);
                    ph ^= 1u << slot;
                } else {
                    cp_async_wait<1>();
                }
                __syncwarp();
                const int s2 = slot == 0 ? 2 : slot - 1;
                if (r + 2 < hi) issue(s2, lds_u16(tour_s + 2 * (r + 2)));
                else if (!tma) cp_async_commit();
                const uint32_t row = rows_s + 4u * slot * n;
                slot = slot == 2 ? 0 : slot + 1;
                return row;
            };
            // row lo-1 is a pure gather, then the candidates of rows lo .. hi-1; ga / gb alternate as previous / current
#pragma unroll 1
            for (int r = lo - 1; r < hi; r += 2) {
                uint32_t row = stage(r);
                two_opt_row<KMAX>(row, r, (uint32_t)r * (uint32_t)n + (uint32_t)lane, lane, off, en, ga, gb,
                                  r >= lo ? lds_f32(edge_s + 4 * r) : ninf, B);
                if (r + 1 < hi) {
                    row = stage(r + 1);
                    two_opt_row<KMAX>(row, r + 1, (uint32_t)(r + 1) * (uint32_t)n + (uint32_t)lane, lane, off, en, gb, ga,
                                      lds_f32(edge_s + 4 * (r + 1)), B);
                }
            }
            if (!tma) cp_async_wait<0>();
            __syncwarp();
        }
        float best = B.c0;
        uint32_t bestkey = B.k0;
        if (B.c1 < best || (B.c1 == best && B.k1 < bestkey)) { best = B.c1; bestkey = B.k1; }
        // CTA arg-min with lowest key on ties (== first strict minimum of the sequential scan)
        for (int o = 16; o > 0; o >>= 1) {
            const float oc = __shfl_xor_sync(DACO_FULL, best, o);
            const uint32_t ok = __shfl_xor_sync(DACO_FULL, bestkey, o);
            if (oc < best || (oc == best && ok < bestkey)) { best = oc; bestkey = ok; }
        }
        if (lane == 0) { S.red_c[warp] = best; S.red_k[warp] = bestkey; }
        __syncthreads();
        best = S.red_c[0];
        bestkey = S.red_k[0];
        for (int w = 1; w < W; ++w) {
            const float oc = S.red_c[w];
            const uint32_t ok = S.red_k[w];
            if (oc < best || (oc == best && ok < bestkey)) { best = oc; bestkey = ok; }
        }
        ++it;
        if (!((double)best < -1e-6)) break;   // two_opt.py:24,36
        const int i = (int)(bestkey / (uint32_t)n), j = (int)(bestkey % (uint32_t)n);
        __syncthreads();
        for (int k = tid; k < (j - i + 1) / 2; k += blockDim.x) {
            const uint16_t x = tour[i + k];
            tour[i + k] = tour[j - k];
            tour[j - k] = x;
        }
    }
    if (lane == 0) S.phase[warp] = ph;
    __syncthreads();
    return it;
}

__device__ float tour_cost_numpy(const float* __restrict__ D, int n, const uint16_t* tour) {
    auto f = [&](int k) -> float { return __ldg(D + (size_t)tour[k] * n + tour[k == 0 ? n - 1 : k - 1]); };
    return numpy_pairwise_sum(f, 0, n);
}

// mode 0: one 2-opt call (ACO.two_opt); mode 1: NLS (ACO.nls).  KMAX = 0: two_opt_call (any n that fits shared memory,
// up to 16 warps); KMAX > 0: two_opt_call_v2 (n + 1 <= 32 * KMAX, 8 warps, two CTAs per SM).
template <int KMAX>
__global__ void __launch_bounds__(KMAX ? 256 : 512, KMAX ? 2 : 1)
two_opt_kernel(const float* __restrict__ dist, const float* __restrict__ heu_dist, uint16_t* __restrict__ tours, int n, int A,
               int mode, int maxt, int T_nls, int T_p, float* __restrict__ costs_out, int32_t* __restrict__ passes_out) {
    DACO_2OPT_SMEM(smem);
    const int tid = threadIdx.x, W = blockDim.x >> 5;
    const int a = blockIdx.x, b = blockIdx.y;
    TwoOptShared S;
    S.bars = reinterpret_cast<uint64_t*>(smem);
    S.phase = reinterpret_cast<uint32_t*>(S.bars + 3 * W);
    S.rows = reinterpret_cast<float*>(S.phase + W);      // W is a multiple of 4: 16-byte aligned
    S.edge = S.rows + (size_t)W * 3 * n;   // [n + 1]
    S.red_c = S.edge + n + 1;
    S.red_k = reinterpret_cast<uint32_t*>(S.red_c + W);
    S.band = reinterpret_cast<int*>(S.red_k + W);
    S.tour = reinterpret_cast<uint16_t*>(S.band + W + 1);
    uint16_t* best_tour = S.tour + n + 1;   // NLS only; tour[n] mirrors tour[0] (position 0 never moves)
    __shared__ float s_best_cost, s_new_cost;

    const float* D = dist + (size_t)b * n * n;
    const float* H = heu_dist ? heu_dist + (size_t)b * n * n : nullptr;
    uint16_t* T = tours + ((size_t)b * A + a) * n;
    for (int k = tid; k <= n; k += blockDim.x) S.tour[k] = T[k == n ? 0 : k];
    if (tid < 3 * W) mbar_init(S.bars + tid, 1);
    if (tid < W) S.phase[tid] = 0;
    fence_barrier_init();
    if (tid == 0) {
        // bands of i in [1, n-1) with ~equal numbers of (i, j) pairs
        const long total = (long)(n - 2) * (n - 1) / 2;
        int i = 1;
        long acc = 0;
        S.band[0] = 1;
        for (int w = 1; w <= W; ++w) {
            const long target = total * w / W;
            while (i < n - 1 && acc < target) { acc += n - 1 - i; ++i; }
            S.band[w] = (w == W) ? n - 1 : i;
        }
    }
    __syncthreads();
    bool permutation = false;
    if constexpr (KMAX > 0) {            // every node exactly once?  (reversals keep it that way)
        int* seen = reinterpret_cast<int*>(S.edge);
        for (int k = tid; k < n; k += blockDim.x) seen[k] = 0;
        __syncthreads();
        for (int k = tid; k < n; k += blockDim.x) {
            const int t = S.tour[k];
            if (t < n) atomicAdd(&seen[t], 1);
        }
        __syncthreads();
        int bad = 0;
        for (int k = tid; k < n; k += blockDim.x) bad |= seen[k] != 1;
        permutation = !__syncthreads_or(bad);
    }
    auto call = [&](const float* M, int max_iterations) -> int {
        if constexpr (KMAX > 0) {
            if (permutation) return two_opt_call_v2<KMAX>(M, n, max_iterations, S);
        }
        return two_opt_call(M, n, max_iterations, S);
    };
    int passes = call(D, maxt);
    if (mode == 1) {
        for (int k = tid; k < n; k += blockDim.x) best_tour[k] = S.tour[k];
        if (tid == 0) s_best_cost = tour_cost_numpy(D, n, S.tour);
        __syncthreads();
        for (int r = 0; r < T_nls; ++r) {
            passes += call(H, T_p);      // perturbation on the heuristic "distance"
            passes += call(D, maxt);
            if (tid == 0) s_new_cost = tour_cost_numpy(D, n, S.tour);
            __syncthreads();
            if (s_new_cost < s_best_cost) {            // tsp_nls/aco.py:252-254
                for (int k = tid; k < n; k += blockDim.x) best_tour[k] = S.tour[k];
                __syncthreads();
                if (tid == 0) s_best_cost = s_new_cost;
            }
            __syncthreads();
        }
        for (int k = tid; k < n; k += blockDim.x) T[k] = best_tour[k];
        if (costs_out && tid == 0) costs_out[(size_t)b * A + a] = s_best_cost;
    } else {
        for (int k = tid; k < n; k += blockDim.x) T[k] = S.tour[k];
    }
    if (passes_out && tid == 0) passes_out[(size_t)b * A + a] = passes;
}

// launch geometry, shared by the launcher in two_opt.cu and the host harness in tests/cpu_emu
inline int two_opt_variant(int n) { return n + 1 <= 128 ? 4 : (n + 1 <= 256 ? 8 : (n + 1 <= 512 ? 16 : 0)); }   // KMAX; 0 = band kernel
inline int two_opt_warps(int variant, int n) { return variant ? 8 : (n >= 256 ? 16 : 8); }   // band kernel: more warps once a pass can feed them
inline size_t two_opt_smem_bytes(int W, int n) {
    return (size_t)W * 28 + ((size_t)W * 3 * n + n + 1 + 2 * W) * 4 + (size_t)(W + 1) * 4 + (size_t)(2 * n + 1) * 2 + 16;
}

}  // namespace deepaco
