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
//
// Kernel source only (launchers and the C ABI are in tsp_update.cu); plain CUDA C++ plus warp intrinsics, so
// tests/cpu_emu compiles the same text for the host.
#pragma once
#include "common.cuh"

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

// Tile version for compact tours: one CTA handles 32 consecutive ants.  Tours are staged in shared memory, each
// ant's inverse permutation is built there, and the neighbour table is written node-major with the 32 ants of a
// node side by side (128-byte coalesced stores) instead of one scattered 4-byte store per (ant, node).
//   smem: tours u16 [32][n] | pos u16 [32][n]
__global__ void __launch_bounds__(256) tsp_cost_tile_kernel(const float* __restrict__ dist, const uint16_t* __restrict__ tours,
                                                            int n, int A, int lbw, int vec, float* __restrict__ costs,
                                                            uint32_t* __restrict__ nbr) {
    DACO_DYN_SMEM16(smem);
    uint16_t* t_s = reinterpret_cast<uint16_t*>(smem);
    uint16_t* pos_s = t_s + (size_t)32 * n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int a0 = blockIdx.x * 32, b = blockIdx.y;
    const int na = min(32, A - a0);
    const uint16_t* T = tours + ((size_t)b * A + a0) * n;
    for (int i = tid; i < na * n; i += blockDim.x) t_s[i] = T[i];
    __syncthreads();
    const float* D = dist + (size_t)b * n * n;
    for (int al = warp; al < na; al += W) {
        const uint16_t* tour = t_s + (size_t)al * n;
        if (costs) {
            auto edge = [&](int k) -> float { return __ldg(D + (size_t)tour[k] * n + tour[k == 0 ? n - 1 : k - 1]); };
            const int a = a0 + al;
            const float c = aten_row_sum_fn(edge, n, lbw, vec != 0, lane, vec ? (int)(((unsigned)a * (unsigned)n) & 3u) : 0);
            if (lane == 0) costs[(size_t)b * A + a] = c;
        }
        if (nbr)
            for (int k = lane; k < n; k += 32) pos_s[(size_t)al * n + tour[k]] = (uint16_t)k;
    }
    __syncthreads();
    if (nbr) {
        uint32_t* N = nbr + (size_t)b * n * A;
        for (int i = tid; i < 32 * n; i += blockDim.x) {
            const int u = i >> 5, al = i & 31;
            if (al < na) {
                const uint16_t* tour = t_s + (size_t)al * n;
                const int k = pos_s[(size_t)al * n + u];
                const uint32_t pr = tour[k == 0 ? n - 1 : k - 1], su = tour[k == n - 1 ? 0 : k + 1];
                N[(size_t)u * A + a0 + al] = (pr << 16) | su;
            }
        }
    }
}

// One warp per matrix row u.  The 2A deposit events of the row (ant a: first its predecessor cell, then its
// successor cell -- the reference's statement order) are bucketed by cell with a stable counting sort in
// shared memory, so each cell then adds its own few weights in ant order: work per row is O(A + n) instead
// of O(A * n), and the row is read and written exactly once, coalesced.
//   smem: inv[A] f32 (1 / cost, shared by the CTA's rows) | per warp: w_sorted[2A] f32 | start[n+1] i32 | cursor[n] i32
__global__ void __launch_bounds__(128) tsp_update_kernel(float* __restrict__ ph, const uint32_t* __restrict__ nbr,
                                                         const float* __restrict__ costs, int n, int A, float decay,
                                                         int elitist, int min_max, float ph_min,
                                                         const float* __restrict__ ph_max, const float* __restrict__ scale,
                                                         const float* __restrict__ heu, float* __restrict__ prod) {
    DACO_DYN_SMEM16(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int u = blockIdx.x * W + warp, b = blockIdx.y;
    const size_t per_warp = (size_t)2 * A * 4 + (size_t)(2 * n + 1) * 4;
    float* inv = reinterpret_cast<float*>(smem);
    float* w_sorted = reinterpret_cast<float*>(smem + (((size_t)A * 4 + 15) & ~(size_t)15) + warp * ((per_warp + 15) & ~(size_t)15));
    int* start = reinterpret_cast<int*>(w_sorted + 2 * A);
    int* cursor = start + n + 1;
    const uint32_t* N = nbr + ((size_t)b * n + u) * A;
    const float* C = costs + (size_t)b * A;
    for (int a = threadIdx.x; a < A; a += blockDim.x) inv[a] = __fdiv_rn(1.0f, C[a]);   // `1.0 / cost` = reciprocal(cost) * 1.0
    __syncthreads();
    if (u >= n) return;

    int a_lo = 0, a_hi = A;
    if (elitist) {   // costs.min(dim=0): first index of the minimum
        float bc = INFINITY;
        int bi = 0x7fffffff;
        for (int a = lane; a < A; a += 32) {
            const float c = C[a];
            if (c < bc) { bc = c; bi = a; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
            const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
            if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        }
        a_lo = bi;
        a_hi = bi + 1;
    }
    const int E = 2 * (a_hi - a_lo);   // events, key e -> ant a_lo + e/2, statement e&1
    for (int v = lane; v <= n; v += 32) start[v] = 0;
    __syncwarp();
    for (int e = lane; e < E; e += 32) {
        const uint32_t nb = N[a_lo + (e >> 1)];
        const int cell = (e & 1) ? (int)(nb & 0xffffu) : (int)(nb >> 16);
        atomicAdd(&start[cell + 1], 1);
    }
    __syncwarp();
    // inclusive scan of start[1..n] (start[0] = 0) -> bucket offsets
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
        const int v = base + lane;
        int x = (v < n) ? start[v + 1] : 0;
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(DACO_FULL, x, off);
            if (lane >= off) x += y;
        }
        x += carry;
        if (v < n) { start[v + 1] = x; }
        carry = __shfl_sync(DACO_FULL, x, 31);
    }
    __syncwarp();
    for (int v = lane; v < n; v += 32) cursor[v] = start[v];
    __syncwarp();
    for (int e0 = 0; e0 < E; e0 += 32) {
        const int e = e0 + lane;
        int cell = -1 - lane;   // distinct dummies for idle lanes
        float w = 0.f;
        if (e < E) {
            const int a = a_lo + (e >> 1);
            const uint32_t nb = N[a];
            cell = (e & 1) ? (int)(nb & 0xffffu) : (int)(nb >> 16);
            w = inv[a];
        }
        const uint32_t grp = __match_any_sync(DACO_FULL, cell);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        if (e < E) w_sorted[cursor[cell] + rank] = w;
        __syncwarp();
        if (e < E && rank == 0) cursor[cell] += __popc(grp);
        __syncwarp();
    }
    float* row = ph + ((size_t)b * n + u) * n;
    const float hi = min_max ? ph_max[b] : 0.f;
    const float sc = scale ? scale[b] : 1.0f;
    for (int base = 0; base < n; base += 32) {      // warp-uniform loop: every lane takes part in the reductions
        const int v = base + lane;
        const bool valid = v < n;
        float val = valid ? row[v] : 0.f;
        if (scale) val = __fmul_rn(val, sc);   // MMAS rescale on the first improvement (tsp/aco.py:86-87)
        val = __fmul_rn(val, decay);
        if (valid)      // this cell's deposits in ant order
            for (int i = start[v]; i < start[v + 1]; ++i) val = __fadd_rn(val, w_sorted[i]);
        if (min_max) {
            // ph[(ph > 1e-9) * ph < min] = min ; ph[ph > max] = max   (tsp/aco.py:117-118)
            const float gate = __fmul_rn(val > 1e-9f ? 1.0f : 0.0f, val);
            if (gate < ph_min) val = ph_min;
            if (val > hi) val = hi;
        }
        if (valid) {
            row[v] = val;
            if (prod) prod[((size_t)b * n + u) * n + v] = __fmul_rn(val, heu[((size_t)b * n + u) * n + v]);
        }
    }
}

// The deposit chain of the ant-sequential update (shared by tsp_update_seq_kernel and tsp_tail_kernel): all threads of the
// CTA; M = pheromone matrix in shared memory, inv[a] = 1 / cost of ant a, ring = u16 [2][16 * n + 2] scratch.
// Tours reach the chain through the double-buffered ring in blocks of 16 ants: the global loads of block j + 1 are issued
// (into registers) before block j is processed and parked in shared memory after it, so their latency is covered by 16
// ant steps; inside a block the next ant's edge is read from shared memory one step ahead of the barrier.
// (Measured alternative, dropped: a node-major neighbour table per block in shared memory with thread u as the sole
// writer of matrix row u -- one barrier per 16 ants instead of one per ant -- is slower: 16 x 2 dependent shared-memory
// read-modify-writes per thread and block, plus building the table, cost more than 16 barriers.)
template <bool SPLIT>
__device__ __forceinline__ void seq_deposit_chain_impl(float* __restrict__ M, const float* __restrict__ inv, uint16_t* __restrict__ ring,
                                                       const uint16_t* __restrict__ tours, int b, int n, int A, int a_lo, int a_hi) {
    const int tid = threadIdx.x, nth = blockDim.x;
    constexpr int kBlk = 16;
    const uint16_t* T = tours + ((size_t)b * A + a_lo) * n;          // ants [a_lo, a_hi) are contiguous
    const int count = a_hi - a_lo;
    const int words = (kBlk * n + 1) / 2;                             // 32-bit words per block (tour rows are contiguous)
    // thread roles: with at least 2n threads the lower half adds to cell (t_k, t_k+1) and the upper half to (t_k+1, t_k)
    // -- one load / add / store per thread and ant; with fewer, thread k does both cells of edge k
    constexpr bool split = SPLIT;
    const int half = nth >> 1;
    const int ke = split ? (tid >= half ? tid - half : tid) : tid;   // edge index of this thread
    const bool flip = split && tid >= half;
    const bool on = ke < n;
    const int k0 = flip ? (ke + 1 == n ? 0 : ke + 1) : ke;            // row endpoint, column endpoint of this thread's cell
    const int k1 = flip ? ke : (ke + 1 == n ? 0 : ke + 1);
    constexpr int kRegs = 8;                                          // words per thread per block: 8 * 256 * 2 >= kBlk * 224
    uint32_t stage[kRegs];
    auto fetch = [&](int blk) {                                       // global -> registers (block `blk`), 4-byte loads
        const size_t base = (size_t)blk * kBlk * n;                   // in uint16 units; even because kBlk is
        const uint32_t* src = reinterpret_cast<const uint32_t*>(T + base);
        const int elems = min(kBlk, count - blk * kBlk) * n, full = elems >> 1;   // an odd last element gets a 2-byte load
#pragma unroll
        for (int q = 0; q < kRegs; ++q) {
            const int wi = tid + q * nth;
            uint32_t w = 0u;
            if (wi < full) w = src[wi];
            else if (wi == full && (elems & 1)) w = T[base + elems - 1];
            stage[q] = w;
        }
    };
    auto fetch16 = [&](int blk) {                                     // same with 2-byte loads (T not 4-byte aligned)
        const uint16_t* src = T + (size_t)blk * kBlk * n;
        const int avail = min(kBlk, count - blk * kBlk) * n;
#pragma unroll
        for (int q = 0; q < kRegs; ++q) {
            const int wi = tid + q * nth;
            uint32_t lo = 0, hi16 = 0;
            if (wi < words) {
                if (2 * wi < avail) lo = src[2 * wi];
                if (2 * wi + 1 < avail) hi16 = src[2 * wi + 1];
            }
            stage[q] = lo | (hi16 << 16);
        }
    };
    auto park = [&](int blk) {                                        // registers -> ring slot blk & 1
        uint32_t* dst = reinterpret_cast<uint32_t*>(ring + (size_t)(blk & 1) * (2 * words));
#pragma unroll
        for (int q = 0; q < kRegs; ++q) {
            const int wi = tid + q * nth;
            if (wi < words) dst[wi] = stage[q];
        }
    };
    const int nblk = (count + kBlk - 1) / kBlk;
    const bool aligned = ((reinterpret_cast<uintptr_t>(T) & 3) == 0);
    if (nblk > 0) {
        if (aligned) fetch(0); else fetch16(0);
        park(0);
    }
    __syncthreads();
    for (int blk = 0; blk < nblk; ++blk) {
        if (blk + 1 < nblk) { if (aligned) fetch(blk + 1); else fetch16(blk + 1); }
        const uint16_t* tb = ring + (size_t)(blk & 1) * (2 * words);
        const int m = min(kBlk, count - blk * kBlk);
        if (m == kBlk) {
            // full block: the 16 cells of this thread (and their weights) go to registers first -- independent loads, no
            // barrier between them -- and the dependent part of an ant step is load cell / add / store / barrier
            float* cu[kBlk];
            float* cv[kBlk];
            float w[kBlk];
#pragma unroll
            for (int i = 0; i < kBlk; ++i) {
                const int u = on ? tb[i * n + k0] : 0, v = on ? tb[i * n + k1] : 0;
                cu[i] = M + (u * n + v);
                cv[i] = M + (v * n + u);
                w[i] = inv[a_lo + blk * kBlk + i];
            }
#pragma unroll
            for (int i = 0; i < kBlk; ++i) {
                if (on) *cu[i] = __fadd_rn(*cu[i], w[i]);
                if (!split && on) *cv[i] = __fadd_rn(*cv[i], w[i]);
                __syncthreads();
            }
        } else {
            int u = on ? tb[k0] : 0, v = on ? tb[k1] : 0;
            for (int i = 0; i < m; ++i) {
                const float w = inv[a_lo + blk * kBlk + i];
                int un = 0, vn = 0;
                if (on && i + 1 < m) {                                // next ant's edge, ahead of the barrier
                    un = tb[(i + 1) * n + k0];
                    vn = tb[(i + 1) * n + k1];
                }
                if (on) {
                    M[u * n + v] = __fadd_rn(M[u * n + v], w);
                    if (!split) M[v * n + u] = __fadd_rn(M[v * n + u], w);
                }
                __syncthreads();
                u = un;
                v = vn;
            }
        }
        if (blk + 1 < nblk) park(blk + 1);                            // slot (blk + 1) & 1 was last read in block blk - 1
        __syncthreads();
    }
}

__device__ __forceinline__ void seq_deposit_chain(float* __restrict__ M, const float* __restrict__ inv, uint16_t* __restrict__ ring,
                                                  const uint16_t* __restrict__ tours, int b, int n, int A, int a_lo, int a_hi) {
    if ((int)blockDim.x >= 2 * n) seq_deposit_chain_impl<true>(M, inv, ring, tours, b, n, A, a_lo, a_hi);
    else seq_deposit_chain_impl<false>(M, inv, ring, tours, b, n, A, a_lo, a_hi);
}

// The same update the way the reference states it (tsp/aco.py:101-114): ants one after another, each ant's 2n cells in
// parallel.  One CTA per colony keeps the whole pheromone matrix in shared memory; thread k owns tour edge k of the
// current ant and adds 1 / cost to cells (t_k, t_k+1) and (t_k+1, t_k) -- the cells of one ant are pairwise distinct for
// n >= 3, so there is nothing to order inside an ant, and the per-ant barrier gives every cell its additions in ant
// order: bit-identical to the reference, at a few dozen warp-instructions per ant instead of a counting sort per matrix
// row (4.2 k instructions for each of the n rows).  No neighbour table in global memory is needed.  Used when the matrix fits in shared memory and the colony has at most ~1 k ants (the
// chain is sequential in the ants; big colonies take tsp_update_row_kernel).
//   smem: M f32 [n][n] | inv f32 [A] | tour ring u16 [2][16 n + 2]
__global__ void __launch_bounds__(256) tsp_update_seq_kernel(float* __restrict__ ph, const uint16_t* __restrict__ tours,
                                                             const float* __restrict__ costs, int n, int A, float decay, int elitist,
                                                             int min_max, float ph_min, const float* __restrict__ ph_max,
                                                             const float* __restrict__ scale, const float* __restrict__ heu,
                                                             float* __restrict__ prod) {
    DACO_DYN_SMEM16(smem);
    __shared__ int s_best;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    float* M = reinterpret_cast<float*>(smem);
    float* inv = M + (size_t)n * n;
    float* P = ph + (size_t)b * n * n;
    const float* C = costs + (size_t)b * A;
    const float sc = scale ? scale[b] : 1.0f;
    for (int i = tid; i < n * n; i += nth) {
        float x = P[i];
        if (scale) x = __fmul_rn(x, sc);      // MMAS rescale on the first improvement (tsp/aco.py:86-87)
        M[i] = __fmul_rn(x, decay);
    }
    for (int a = tid; a < A; a += nth) inv[a] = __fdiv_rn(1.0f, C[a]);   // `1.0 / cost` = reciprocal(cost) * 1.0
    int a_lo = 0, a_hi = A;
    if (elitist) {   // costs.min(dim=0): first index of the minimum
        if (warp == 0) {
            float bc = INFINITY;
            int bi = 0x7fffffff;
            for (int a = lane; a < A; a += 32) {
                const float c = C[a];
                if (c < bc) { bc = c; bi = a; }
            }
            for (int off = 16; off > 0; off >>= 1) {
                const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
                const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
                if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
            }
            if (lane == 0) s_best = bi;
        }
        __syncthreads();
        a_lo = s_best;
        a_hi = a_lo + 1;
    }
    __syncthreads();
    seq_deposit_chain(M, inv, reinterpret_cast<uint16_t*>(inv + A), tours, b, n, A, a_lo, a_hi);
    const float hi = min_max ? ph_max[b] : 0.f;
    for (int i = tid; i < n * n; i += nth) {
        float x = M[i];
        if (min_max) {
            // ph[(ph > 1e-9) * ph < min] = min ; ph[ph > max] = max   (tsp/aco.py:117-118)
            const float gate = __fmul_rn(x > 1e-9f ? 1.0f : 0.0f, x);
            if (gate < ph_min) x = ph_min;
            if (x > hi) x = hi;
        }
        P[i] = x;
        if (prod) prod[(size_t)b * n * n + i] = __fmul_rn(x, heu[(size_t)b * n * n + i]);
    }
}

// The whole iteration tail of ACO.run (tsp/aco.py:76-90) for one colony in ONE launch: tour costs (ATen summation order,
// distance matrix staged in shared memory) -> iteration best vs running best, MMAS bookkeeping (what tsp_best_kernel
// does) -> evaporation + the ant-sequential deposit chain -> clamp, pheromone and next product matrix out.  Same bits as
// tsp_cost_tile_kernel + tsp_best_kernel + tsp_update_seq_kernel, two launches and one pass over the tours fewer.
//   smem: M f32 [n][n] | Dm f32 [n][n] | cs f32 [A] (costs, then 1 / cost) | ring u16 [2][16 n + 2] (phase 1: one tour per warp)
struct TailParams {
    float* ph;                 // [B][n][n] in / out
    const uint16_t* tours;     // [B][A][n]
    const float* dist;         // [B][n][n]
    const float* heu;          // [B][n][n]
    float* prod;               // [B][n][n] out
    float* costs;              // [B][A] out
    float* lowest;             // [B] in / out
    int64_t* shortest;         // [B][n] in / out
    float* ph_max;             // [B] in / out (min_max)
    int n, A;
    float decay;
    int elitist, min_max;
    float ph_min;
    int lbw, vec;              // ATen summation plan of a [A][n] row sum
};

__global__ void __launch_bounds__(256) tsp_tail_kernel(const TailParams p) {
    DACO_DYN_SMEM16(smem);
    __shared__ float s_c[8], s_m[8];
    __shared__ int s_i[8];
    __shared__ int s_improved, s_first, s_bi;
    __shared__ float s_bc, s_scale;
    __shared__ uint64_t s_bar;
    const int n = p.n, A = p.A;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, W = nth >> 5, b = blockIdx.x;
    float* M = reinterpret_cast<float*>(smem);
    float* Dm = M + (size_t)n * n;
    float* cs = Dm + (size_t)n * n;
    uint16_t* ring = reinterpret_cast<uint16_t*>(cs + A);
    float* P = p.ph + (size_t)b * n * n;
    const float* D = p.dist + (size_t)b * n * n;
    const uint16_t* T = p.tours + (size_t)b * A * n;
    const float* H = p.heu + (size_t)b * n * n;
    // matrices move by TMA bulk copies when 16-byte granular (n even): pheromone and distances now; the heuristic later,
    // into the distance matrix's place once the costs are done, so that its read hides behind the deposit chain
    const uint32_t mat_bytes = (uint32_t)n * n * 4;
    const bool bulk = (mat_bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(D) |
                                                  reinterpret_cast<uintptr_t>(H)) & 15) == 0;
    auto bulk_load = [&](float* dst, const float* src) {   // thread 0 only
        constexpr uint32_t kChunk = 32768;
        for (uint32_t off = 0; off < mat_bytes; off += kChunk)
            tma_bulk_g2s(reinterpret_cast<char*>(dst) + off, reinterpret_cast<const char*>(src) + off,
                         mat_bytes - off < kChunk ? mat_bytes - off : kChunk, &s_bar);
    };
    if (bulk) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            fence_barrier_init();
            mbar_expect_tx(&s_bar, 2 * mat_bytes);
            bulk_load(M, P);
            bulk_load(Dm, D);
        }
        __syncthreads();
        mbar_wait(&s_bar, 0);
    } else {
        for (int i = tid; i < n * n; i += nth) {
            M[i] = P[i];
            Dm[i] = D[i];
        }
    }
    __syncthreads();
    // ---- costs: a warp per ant, the next ant's tour already in flight
    if (!p.vec && p.lbw == 5) {
        // 32 <= n < 128, ATen's plain plan: lane l sums elements l, l + 32, l + 64, l + 96 into one accumulator each
        // (aten_row_sum_fn's non-vectorised branch) -- exactly the elements the lane loaded, so the tour stays in
        // registers and the predecessor t[k - 1] comes from the lane below by shuffle
        // Tours come straight from L2: kAhead ants per warp are in flight (one ant's arithmetic is ~150 cycles, the load
        // latency several hundred).
        constexpr int KT = 4, kAhead = 4;
        uint32_t q[kAhead][KT];
        auto load = [&](uint32_t (&dst)[KT], int a) {
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                const int k = lane + 32 * j;
                dst[j] = (a < A && k < n) ? (uint32_t)T[(size_t)a * n + k] : 0u;
            }
        };
        const int jl = (n - 1) >> 5, ll = (n - 1) & 31;
#pragma unroll
        for (int d = 0; d < kAhead; ++d) load(q[d], warp + d * W);
        for (int a0 = warp; a0 < A; a0 += kAhead * W) {
#pragma unroll
            for (int d = 0; d < kAhead; ++d) {
                const int a = a0 + d * W;
                if (a >= A) break;                                  // warp-uniform
                uint32_t cur[KT];
#pragma unroll
                for (int j = 0; j < KT; ++j) cur[j] = q[d][j];
                load(q[d], a + kAhead * W);
                uint32_t carry = __shfl_sync(DACO_FULL, jl == 0 ? cur[0] : jl == 1 ? cur[1] : jl == 2 ? cur[2] : cur[3], ll);   // t[n - 1]
                float acc[KT];
#pragma unroll
                for (int j = 0; j < KT; ++j) {
                    const uint32_t up = __shfl_up_sync(DACO_FULL, cur[j], 1);
                    const uint32_t pred = lane == 0 ? carry : up;
                    carry = __shfl_sync(DACO_FULL, cur[j], 31);
                    // branch-free: lanes past the end hold tour value 0, read cell (0, pred) and mask the bits to +0
                    const uint32_t keep = (uint32_t)((lane + 32 * j - n) >> 31);       // all ones iff lane + 32 j < n
                    const float e = Dm[cur[j] * (uint32_t)n + pred];
                    acc[j] = __fadd_rn(0.f, __uint_as_float(__float_as_uint(e) & keep));
                }
                const float c = warp_tree_sum(__fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]));
                if (lane == 0) {
                    cs[a] = c;
                    p.costs[(size_t)b * A + a] = c;
                }
            }
        }
    } else {
        // general plan: the tour staged in the warp's slice of `ring`
        uint16_t* tw = ring + (size_t)warp * n;
        constexpr int KT = 8;                               // n <= 224 < 8 * 32
        uint16_t nxt[KT];
        auto load = [&](int a) {
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                const int k = lane + 32 * j;
                nxt[j] = (a < A && k < n) ? T[(size_t)a * n + k] : (uint16_t)0;
            }
        };
        load(warp);
        for (int a = warp; a < A; a += W) {
#pragma unroll
            for (int j = 0; j < KT; ++j) {
                const int k = lane + 32 * j;
                if (k < n) tw[k] = nxt[j];
            }
            __syncwarp();
            load(a + W);
            auto edge = [&](int k) -> float { return Dm[(int)tw[k] * n + (int)tw[k == 0 ? n - 1 : k - 1]]; };
            const float c = aten_row_sum_fn(edge, n, p.lbw, p.vec != 0, lane, p.vec ? (int)(((unsigned)a * (unsigned)n) & 3u) : 0);
            if (lane == 0) {
                cs[a] = c;
                p.costs[(size_t)b * A + a] = c;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (bulk && tid == 0) {                                  // distances are done with: the heuristic takes their place
        mbar_expect_tx(&s_bar, mat_bytes);
        bulk_load(Dm, H);
    }
    __syncthreads();
    // ---- iteration best -> running best, MMAS bookkeeping (tsp/aco.py:79-88)
    {
        float bc = INFINITY;
        int bi = 0x7fffffff;
        for (int a = tid; a < A; a += nth) {
            const float c = cs[a];
            if (c < bc) { bc = c; bi = a; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
            const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
            if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        }
        if (lane == 0) { s_c[warp] = bc; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bc = lane < W ? s_c[lane] : INFINITY;
            bi = lane < W ? s_i[lane] : 0x7fffffff;
            for (int off = 16; off > 0; off >>= 1) {
                const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
                const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
                if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
            }
            if (lane == 0) {
                s_improved = bc < p.lowest[b];                                   // tsp/aco.py:81
                s_first = p.min_max && !(p.ph_max[b] > 0.f);                     // MMAS max not set yet (marker 0)
                s_bc = bc;
                s_bi = bi;
                s_scale = 1.0f;
            }
        }
        __syncthreads();
    }
    float hi = p.min_max ? p.ph_max[b] : 0.f;
    if (s_improved) {
        const float bc = s_bc;
        const uint16_t* t = T + (size_t)s_bi * n;
        for (int k = tid; k < n; k += nth) p.shortest[(size_t)b * n + k] = (int64_t)t[k];
        if (p.min_max) {
            // max = problem_size / lowest_cost  ==  reciprocal(lowest) * n   (Tensor.__rtruediv__)
            const float new_max = __fmul_rn(__fdiv_rn(1.0f, bc), (float)n);
            if (s_first) {   // self.pheromone *= max / self.pheromone.max()
                float m = -INFINITY;
                for (int i = tid; i < n * n; i += nth) m = fmaxf(m, M[i]);
                for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(DACO_FULL, m, off));
                if (lane == 0) s_m[warp] = m;
                __syncthreads();
                if (warp == 0) {
                    m = lane < W ? s_m[lane] : -INFINITY;
                    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(DACO_FULL, m, off));
                    if (lane == 0) s_scale = __fdiv_rn(new_max, m);
                }
            }
            hi = new_max;
        }
        __syncthreads();                                     // every thread has read lowest / ph_max before they change
        if (tid == 0) {
            p.lowest[b] = bc;
            if (p.min_max) p.ph_max[b] = __fmul_rn(__fdiv_rn(1.0f, bc), (float)n);
        }
    }
    __syncthreads();
    // ---- evaporation (+ MMAS rescale), reciprocal costs, deposit chain
    const float sc = s_scale;
    for (int i = tid; i < n * n; i += nth) {
        float x = M[i];
        if (p.min_max) x = __fmul_rn(x, sc);
        M[i] = __fmul_rn(x, p.decay);
    }
    for (int a = tid; a < A; a += nth) cs[a] = __fdiv_rn(1.0f, cs[a]);
    const int a_lo = p.elitist ? s_bi : 0, a_hi = p.elitist ? s_bi + 1 : A;   // costs.min(dim=0): first index of the minimum
    __syncthreads();
    seq_deposit_chain(M, cs, ring, p.tours, b, n, A, a_lo, a_hi);
    if (bulk) mbar_wait(&s_bar, 1);
    const float* Hs = bulk ? Dm : H;
    for (int i = tid; i < n * n; i += nth) {
        float x = M[i];
        if (p.min_max) {
            // ph[(ph > 1e-9) * ph < min] = min ; ph[ph > max] = max   (tsp/aco.py:117-118)
            const float gate = __fmul_rn(x > 1e-9f ? 1.0f : 0.0f, x);
            if (gate < p.ph_min) x = p.ph_min;
            if (x > hi) x = hi;
        }
        P[i] = x;
        p.prod[(size_t)b * n * n + i] = __fmul_rn(x, Hs[i]);
    }
}

// Candidate lists for the kNN construction kernel, re-derived from the CURRENT product matrix: columns of the 32 largest
// entries of every row (ties: lower column first), in any order.  The lists are only a performance hint -- the kernel
// bounds the unlisted columns by their actual maximum -- but as the pheromone evolves, edges outside the heuristic's
// top 32 get reinforced and lists taken from the heuristic alone send more and more steps to the dense fallback.
// One warp per row, n <= 256, entries >= 0: the bit pattern of a non-negative float is monotonic in its value, so the
// 32nd largest value is found by a 31-step bisection on the bits (count of entries >= trial via ballots); everything
// above it is listed, and the remaining places go to the entries equal to it in column order.
__global__ void __launch_bounds__(256) knn_refresh_kernel(const float* __restrict__ prod, uint8_t* __restrict__ knn, int n,
                                                          int rows_total) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= rows_total) return;
    const float* P = prod + (size_t)row * n;
    const int K = (n + 31) >> 5;
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = lane + 32 * k;                     // column of slot k; padding slots never qualify (trial >= 1)
        v[k] = (k < K && idx < n) ? __float_as_uint(fmaxf(P[idx], 0.f)) : 0u;
    }
    uint32_t thr = 0u;                                     // largest t with  #{v >= t} >= 32
    for (int bit = 30; bit >= 0; --bit) {
        const uint32_t trial = thr | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < K) cnt += __popc(__ballot_sync(DACO_FULL, v[k] >= trial));
        if (cnt >= 32) thr = trial;
    }
    // thr == 0: fewer than 32 positive entries -- zeros (and only real columns) fill up in column order
    int above = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < K) above += __popc(__ballot_sync(DACO_FULL, v[k] > thr && lane + 32 * k < n));
    int slot_hi = 0, slot_eq = above;                      // next free place for "> thr" and for "== thr" entries
    const uint32_t lt = (1u << lane) - 1u;
    uint8_t* out = knn + (size_t)row * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k < K) {
            const int idx = lane + 32 * k;
            const bool real = idx < n;
            const uint32_t hi = __ballot_sync(DACO_FULL, real && v[k] > thr);
            const uint32_t eq = __ballot_sync(DACO_FULL, real && v[k] == thr);
            if (real && v[k] > thr) out[slot_hi + __popc(hi & lt)] = (uint8_t)idx;
            if (real && v[k] == thr) {
                const int pos = slot_eq + __popc(eq & lt);
                if (pos < 32) out[pos] = (uint8_t)idx;
            }
            slot_hi += __popc(hi);
            slot_eq += __popc(eq);
        }
    }
}

// Same update for colonies with MANY ants (thousands: the ant-sharded path), where a row's 2A deposit events are too
// many for one warp: one CTA per matrix row.  Ants are processed in chunks of `CH` (ant order, so the per-cell add order
// is kept across chunks); inside a chunk warp w owns a contiguous ant range and buckets its events by cell into its own
// sub-bucket -- buckets are laid out cell-major, warp-minor, so reading a cell's bucket front to back is ant order again.
//   smem: val f32 [n] | cellstart i32 [n+1] | cnt i32 [W][n] (counts, then cursors) | w_sorted f32 [2*CH] |
//         nb u32 [CH] | inv f32 [CH]  (the chunk's neighbour words and 1 / cost, staged by the whole CTA)
__global__ void __launch_bounds__(256) tsp_update_row_kernel(float* __restrict__ ph, const uint32_t* __restrict__ nbr,
                                                             const float* __restrict__ costs, int n, int A, int CH, float decay,
                                                             int min_max, float ph_min, const float* __restrict__ ph_max,
                                                             const float* __restrict__ scale, const float* __restrict__ heu,
                                                             float* __restrict__ prod) {
    DACO_DYN_SMEM16(smem);
    const int tid = threadIdx.x, nthreads = blockDim.x, warp = tid >> 5, lane = tid & 31, W = nthreads >> 5;
    const int u = blockIdx.x, b = blockIdx.y;
    float* val = reinterpret_cast<float*>(smem);
    int* cellstart = reinterpret_cast<int*>(val + n);
    int* cnt = cellstart + n + 1;
    float* w_sorted = reinterpret_cast<float*>(cnt + (size_t)W * n);
    uint32_t* nb_s = reinterpret_cast<uint32_t*>(w_sorted + (size_t)2 * CH);
    float* inv_s = reinterpret_cast<float*>(nb_s + CH);
    const uint32_t* N = nbr + ((size_t)b * n + u) * A;
    const float* C = costs + (size_t)b * A;
    float* row = ph + ((size_t)b * n + u) * n;
    const float sc = scale ? scale[b] : 1.0f;
    for (int v = tid; v < n; v += nthreads) {
        float x = row[v];
        if (scale) x = __fmul_rn(x, sc);   // MMAS rescale on the first improvement (tsp/aco.py:86-87)
        val[v] = __fmul_rn(x, decay);
    }
    int* mycnt = cnt + (size_t)warp * n;
    for (int a0 = 0; a0 < A; a0 += CH) {
        const int ca = min(CH, A - a0);
        const int per = (ca + W - 1) / W;
        const int wa0 = min(ca, warp * per), wa1 = min(ca, wa0 + per);
        const int e_lo = 2 * wa0, e_hi = 2 * wa1;          // this warp's events of the chunk: e -> ant a0 + e/2, statement e&1
        for (int i = tid; i < W * n; i += nthreads) cnt[i] = 0;
        if (tid == 0) cellstart[0] = 0;
        for (int i0 = tid; i0 < ca; i0 += 4 * nthreads) {   // coalesced; eight loads per thread in flight before the first use
            uint32_t nb4[4];
            float c4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * nthreads;
                nb4[q] = i < ca ? N[a0 + i] : 0u;
                c4[q] = i < ca ? C[a0 + i] : 1.0f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * nthreads;
                if (i < ca) {
                    nb_s[i] = nb4[q];
                    inv_s[i] = __fdiv_rn(1.0f, c4[q]);   // `1.0 / cost` = reciprocal(cost) * 1.0
                }
            }
        }
        __syncthreads();
        for (int e0 = e_lo; e0 < e_hi; e0 += 32) {   // most ants share a few cells: count per group, not per lane
            const int e = e0 + lane;
            int cell = -1 - lane;
            if (e < e_hi) {
                const uint32_t nb = nb_s[e >> 1];
                cell = (e & 1) ? (int)(nb & 0xffffu) : (int)(nb >> 16);
            }
            const uint32_t grp = __match_any_sync(DACO_FULL, cell);
            if (e < e_hi && (grp & ((1u << lane) - 1u)) == 0u) mycnt[cell] += __popc(grp);
            __syncwarp();
        }
        __syncthreads();
        for (int v = tid; v < n; v += nthreads) {
            int tot = 0;
            for (int w = 0; w < W; ++w) tot += cnt[(size_t)w * n + v];
            cellstart[v + 1] = tot;
        }
        __syncthreads();
        if (warp == 0) {   // inclusive scan of cellstart[1..n]
            int carry = 0;
            for (int base = 0; base < n; base += 32) {
                const int v = base + lane;
                int x = (v < n) ? cellstart[v + 1] : 0;
                for (int off = 1; off < 32; off <<= 1) {
                    const int y = __shfl_up_sync(DACO_FULL, x, off);
                    if (lane >= off) x += y;
                }
                x += carry;
                if (v < n) cellstart[v + 1] = x;
                carry = __shfl_sync(DACO_FULL, x, 31);
            }
        }
        __syncthreads();
        for (int v = tid; v < n; v += nthreads) {   // counts -> cursors (cell-major, warp-minor)
            int run = cellstart[v];
            for (int w = 0; w < W; ++w) {
                const int t = cnt[(size_t)w * n + v];
                cnt[(size_t)w * n + v] = run;
                run += t;
            }
        }
        __syncthreads();
        for (int e0 = e_lo; e0 < e_hi; e0 += 32) {   // stable placement of this warp's events
            const int e = e0 + lane;
            int cell = -1 - lane;   // distinct dummies for idle lanes
            float w = 0.f;
            if (e < e_hi) {
                const uint32_t nb = nb_s[e >> 1];
                cell = (e & 1) ? (int)(nb & 0xffffu) : (int)(nb >> 16);
                w = inv_s[e >> 1];
            }
            const uint32_t grp = __match_any_sync(DACO_FULL, cell);
            const int rank = __popc(grp & ((1u << lane) - 1u));
            if (e < e_hi) w_sorted[mycnt[cell] + rank] = w;
            __syncwarp();
            if (e < e_hi && rank == 0) mycnt[cell] += __popc(grp);
            __syncwarp();
        }
        __syncthreads();
        for (int v = tid; v < n; v += nthreads) {   // this cell's deposits of the chunk, in ant order
            float x = val[v];
            int i = cellstart[v];
            const int end = cellstart[v + 1];
            // the add chain is sequential by definition (fp32, ant order); keep the loads of the next eight ahead of it
            if (i + 8 <= end) {
                float c[8], d[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) c[k] = w_sorted[i + k];
                i += 8;
                for (; i + 16 <= end; i += 16) {      // two batches per trip: the buffers swap roles without register moves
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[k] = w_sorted[i + k];
#pragma unroll
                    for (int k = 0; k < 8; ++k) x = __fadd_rn(x, c[k]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) c[k] = w_sorted[i + 8 + k];
#pragma unroll
                    for (int k = 0; k < 8; ++k) x = __fadd_rn(x, d[k]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) x = __fadd_rn(x, c[k]);
            }
            for (; i < end; ++i) x = __fadd_rn(x, w_sorted[i]);
            val[v] = x;
        }
        __syncthreads();
    }
    const float hi = min_max ? ph_max[b] : 0.f;
    for (int v = tid; v < n; v += nthreads) {
        float x = val[v];
        if (min_max) {
            // ph[(ph > 1e-9) * ph < min] = min ; ph[ph > max] = max   (tsp/aco.py:117-118)
            const float gate = __fmul_rn(x > 1e-9f ? 1.0f : 0.0f, x);
            if (gate < ph_min) x = ph_min;
            if (x > hi) x = hi;
        }
        row[v] = x;
        if (prod) prod[((size_t)b * n + u) * n + v] = __fmul_rn(x, heu[((size_t)b * n + u) * n + v]);
    }
}

}  // namespace deepaco
