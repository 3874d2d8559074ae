// Instance -> graph front end: Euclidean distance matrix + k-nearest-neighbour edges of a batch of instances.
//
// Replaces, for CUDA inputs, the reference's eager op chain (tsp/utils.py:4-36, tsp_nls/utils.py:5-46; the distance part
// also cvrp/utils.py:18-22):
//     distances = torch.norm(coords[:, None] - coords, dim=2, p=2);  distances[diag] = 1e9 (TSP) | 1e-10 (CVRP)
//     topk_values, topk_indices = torch.topk(distances, k, dim=1, largest=False)
//     edge_index = stack([repeat_interleave(arange(n), k), flatten(topk_indices)]);  edge_attr = topk_values
// One warp per (instance, row).  Distances are bit-identical to ATen's: the subtraction is its own rounded op, and the
// norm over the pair (dx, dy) is a two-lane reduction -- each lane squares its element (fma(x, x, 0) = fl(x * x)), the
// lanes combine with one add, then sqrt: sqrt(fl(fl(dx * dx) + fl(dy * dy))), no fused multiply-add across the two
// (probed on the B200: tools/probes/norm_formula.py, 0 mismatches for n = 20 .. 1000; the fma forms differ in ~8 % of the
// entries).  The k smallest entries of the row come out in ascending order, k rounds of a warp arg-min over an
// order-preserving integer key; bit-equal distances inside a row go lowest column first (torch.topk's order among equal
// values is unspecified -- the probe shows neither ascending nor descending -- so edge order can differ from torch's only
// inside such a tie, and the edge SET only if the tie straddles rank k).
//
// Kernel source only (C ABI in knn_graph.cu); plain CUDA C++ + warp intrinsics, compiled for the host by tests/cpu_emu.
#pragma once
#include "common.cuh"

namespace deepaco {

struct KnnGraphParams {
    const float* coords;    // [B][n][2] or null
    const float* dist_in;   // [B][n][n] or null (exactly one of the two inputs)
    float* dist_out;        // [B][n][n] or null
    int32_t* nbr_index;     // [B][n][k] or null
    float* nbr_value;       // [B][n][k] or null
    int64_t* edge_index;    // [B][2][n * k] or null: row 0 = source (repeat_interleave), row 1 = neighbour
    int n, B, k;
    float diag;             // value of distances[i][i] (the reference overwrites the diagonal)
};

// monotone map float -> uint32 (total order of the finite values and infinities, -0 < +0)
__device__ __forceinline__ uint32_t float_order_key(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) knn_graph_kernel(const KnnGraphParams p) {
    DACO_DYN_SMEM16(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n, k = p.k;
    const long r = (long)blockIdx.x * (blockDim.x >> 5) + warp;      // global row = b * n + i
    if (r >= (long)p.B * n) return;
    const int b = (int)(r / n), i = (int)(r - (long)b * n);
    float* row = reinterpret_cast<float*>(smem) + (size_t)warp * n;
    if (p.coords) {
        const float* C = p.coords + (size_t)b * n * 2;
        const float xi = C[2 * i], yi = C[2 * i + 1];
        for (int j = lane; j < n; j += 32) {
            const float dx = __fsub_rn(xi, C[2 * j]), dy = __fsub_rn(yi, C[2 * j + 1]);
            const float d = j == i ? p.diag : __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            row[j] = d;
            if (p.dist_out) p.dist_out[(size_t)r * n + j] = d;
        }
    } else {
        for (int j = lane; j < n; j += 32) {
            const float d = p.dist_in[(size_t)r * n + j];
            row[j] = d;
            if (p.dist_out) p.dist_out[(size_t)r * n + j] = d;
        }
    }
    __syncwarp();
    const size_t E = (size_t)n * k;
    for (int t = 0; t < k; ++t) {
        uint32_t best = 0xffffffffu, bj = 0xffffffffu;               // key 0xffffffff = taken (only NaN maps there)
        for (int j = lane; j < n; j += 32) {
            const uint32_t key = float_order_key(row[j]);
            if (key < best) { best = key; bj = (uint32_t)j; }         // ascending j per lane: first of equal keys stays
        }
        const uint32_t top = __reduce_min_sync(DACO_FULL, best);
        const uint32_t jj = __reduce_min_sync(DACO_FULL, best == top ? bj : 0xffffffffu);
        if (jj == 0xffffffffu) break;                                 // fewer than k selectable entries (host checks k <= n)
        const float v = row[jj];
        __syncwarp();
        if (lane == 0) {
            row[jj] = __uint_as_float(0x7fffffffu);                   // NaN pattern: key 0xffffffff, never selected again
            if (p.nbr_index) p.nbr_index[(size_t)r * k + t] = (int32_t)jj;
            if (p.nbr_value) p.nbr_value[(size_t)r * k + t] = v;
            if (p.edge_index) {
                int64_t* ei = p.edge_index + (size_t)b * 2 * E;
                ei[(size_t)i * k + t] = i;
                ei[E + (size_t)i * k + t] = (int64_t)jj;
            }
        }
        __syncwarp();
    }
}

}  // namespace deepaco
