// K3 -- heuristic network forward (reference tsp/net.py:8-102; identical in tsp_nls/ and cvrp/ up to `feats`).
//
//   EmbNet (net.py:27-45):  x = silu(Lin_v0 x);  w = silu(Lin_e0 e)
//     12 x {  x1..x4 = Lin_1..4(x);  w1 = Lin_e(w);  w2 = sigmoid(w)
//             x <- x + silu(BN_v(x1 + mean_{e: src(e)=i} w2_e * x2[dst(e)]))         (global_mean_pool over edge_index[0])
//             w <- w + silu(BN_e(w1 + x3[src] + x4[dst])) }
//   ParNet (net.py:48-75):  heu = sigmoid(Lin(silu(Lin(silu(Lin(w))))))  -> one value per edge
//
// Edge-parallel gather / scatter-free formulation: edges arrive sorted by source node (CSR), so the mean
// aggregation is a fixed-order segmented sum (deterministic, no atomics -- the reference's torch_scatter path
// uses atomics).  One CTA per instance runs all 12 layers: the per-layer weight block (22 KB) is double
// buffered in shared memory by TMA bulk copies, node-side intermediates live in a small global scratch
// (L2-resident), the edge state w[E][32] streams through L2 once per layer, and the 32x32 linears are fp32
// FMAs with the weight row broadcast from shared memory (M is tiny; a tensor-core tile would be mostly padding).
// BatchNorm runs in eval mode (running statistics), as in the reference's test drivers.
//
// Kernel source only (the C ABI is in gnn.cu); apart from the TMA / mbarrier helpers of common.cuh it is plain CUDA
// C++, so tests/cpu_emu compiles the same text for the host.
#pragma once
#include "common.cuh"

namespace deepaco {

constexpr int U = 32;                                            // units
constexpr int kLayerFloats = 4 * (U * U + U) + (U * U + U) + 8 * U;   // 5536
constexpr int kDepth = 12;

struct GnnParams {
    const float* x_in;      // [B][n][feats]
    const int32_t* row_ptr; // [B][n+1]   CSR by source node
    const int32_t* dst;     // [B][E]     destination of the sorted edge
    const float* attr;      // [B][E]     edge attribute of the sorted edge
    const int32_t* order;   // [B][E]     original edge id of the sorted edge (output scatter)
    const float* weights;   // packed, see deepaco_b200/net.py pack_weights()
    float* node_ws;         // [B][n][6*32] scratch: x | x1 | x3 | agg | x2 | x4
    float* edge_ws;         // [B][E][32]   scratch: w
    float* out;             // [B][E]       heuristic per ORIGINAL edge id (may be null when dense_out is given)
    float* dense_out;       // [B][n][n] or null: Net.reshape(pyg, heu) + eps  (zero-padded matrix, tsp/net.py:95-102)
    float dense_eps;
    int n, E, feats;
};

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// out[o] = b[o] + sum_k W[o][k] * in[k]   (W, b in shared memory: broadcast reads)
__device__ __forceinline__ void linear32(const float* __restrict__ Wb, const float (&in)[U], float (&out)[U]) {
#pragma unroll
    for (int o = 0; o < U; ++o) {
        float acc = Wb[U * U + o];
#pragma unroll
        for (int k = 0; k < U; ++k) acc = fmaf(Wb[o * U + k], in[k], acc);
        out[o] = acc;
    }
}

__global__ void __launch_bounds__(512) gnn_forward_kernel(const GnnParams p) {
    DACO_DYN_SMEM128(smem);
    __shared__ uint64_t bars[2];
    float* wbuf = reinterpret_cast<float*>(smem);                 // [2][kLayerFloats]
    float* head = wbuf + 2 * kLayerFloats;                        // 2*(U*U+U) + U + 1
    float* w0s = head + 2 * (U * U + U) + U + 1;                  // v_lin0 [U][feats] + b[U], e_lin0 [U] + b[U]
    const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
    const int n = p.n, E = p.E, F = p.feats;
    const int32_t* rp = p.row_ptr + (size_t)b * (n + 1);
    const int32_t* dst = p.dst + (size_t)b * E;
    float* NW = p.node_ws + (size_t)b * n * 6 * U;
    float* X = NW, *X1 = NW + (size_t)n * U, *X3 = NW + (size_t)2 * n * U, *AG = NW + (size_t)3 * n * U;
    float* X2 = NW + (size_t)4 * n * U, *X4 = NW + (size_t)5 * n * U;
    float* Wst = p.edge_ws + (size_t)b * E * U;
    const float* Wg = p.weights;
    const int off_layers = U * F + U + U + U;                     // after v_lin0 and e_lin0
    const float* layers_g = Wg + off_layers;
    const float* head_g = layers_g + (size_t)kDepth * kLayerFloats;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < off_layers; i += nth) w0s[i] = Wg[i];
    for (int i = tid; i < 2 * (U * U + U) + U + 1; i += nth) head[i] = head_g[i];
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bars[0], kLayerFloats * 4);
        tma_bulk_g2s(wbuf, layers_g, kLayerFloats * 4, &bars[0]);
    }
    // ---- input embeddings
    for (int t = tid; t < n * U; t += nth) {
        const int i = t / U, o = t % U;
        float acc = w0s[U * F + o];
        for (int k = 0; k < F; ++k) acc = fmaf(w0s[o * F + k], p.x_in[((size_t)b * n + i) * F + k], acc);
        X[t] = silu_f(acc);
    }
    {
        const float* We0 = w0s + U * F + U;
        const float* attr = p.attr + (size_t)b * E;
        for (int t = tid; t < E * U; t += nth) {
            const int e = t / U, o = t % U;
            Wst[t] = silu_f(fmaf(We0[o], attr[e], We0[U + o]));
        }
    }
    __syncthreads();

    for (int l = 0; l < kDepth; ++l) {
        const float* Wl = wbuf + (size_t)(l & 1) * kLayerFloats;
        if (tid == 0 && l + 1 < kDepth) {      // prefetch next layer's weights into the other buffer
            mbar_expect_tx(&bars[(l + 1) & 1], kLayerFloats * 4);
            tma_bulk_g2s(wbuf + (size_t)((l + 1) & 1) * kLayerFloats, layers_g + (size_t)(l + 1) * kLayerFloats,
                         kLayerFloats * 4, &bars[(l + 1) & 1]);
        }
        mbar_wait(&bars[l & 1], (l >> 1) & 1);
        const float* Wv = Wl;                                   // 4 x (W[32][32], b[32])
        const float* We = Wl + 4 * (U * U + U);
        const float* bnv = We + (U * U + U);                    // gamma, beta, mean, invstd
        const float* bne = bnv + 4 * U;

        // ---- node linears: task = (node, which linear)
        for (int t = tid; t < n * 4; t += nth) {
            const int i = t >> 2, q = t & 3;
            float in[U], out[U];
#pragma unroll
            for (int k = 0; k < U; k += 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(X + (size_t)i * U + k));
                in[k] = v.x; in[k + 1] = v.y; in[k + 2] = v.z; in[k + 3] = v.w;
            }
            linear32(Wv + q * (U * U + U), in, out);
            float* dstp = (q == 0 ? X1 : q == 1 ? X2 : q == 2 ? X3 : X4) + (size_t)i * U;
#pragma unroll
            for (int k = 0; k < U; k += 4) *reinterpret_cast<float4*>(dstp + k) = make_float4(out[k], out[k + 1], out[k + 2], out[k + 3]);
        }
        __syncthreads();
        // ---- aggregation: task = (node, feature); fixed edge order within the node's CSR segment
        for (int t = tid; t < n * U; t += nth) {
            const int i = t / U, f = t % U;
            const int e0 = rp[i], e1 = rp[i + 1];
            float s = 0.f;
            for (int e = e0; e < e1; ++e)
                s += sigmoid_f(__ldcg(Wst + (size_t)e * U + f)) * __ldcg(X2 + (size_t)dst[e] * U + f);
            const int deg = e1 - e0;
            AG[t] = s / (float)(deg > 0 ? deg : 1);
        }
        __syncthreads();
        // ---- edge update: task = edge
        for (int e = tid; e < E; e += nth) {
            // source node of sorted edge e: binary search in row_ptr
            int lo = 0, hi = n;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (rp[mid] <= e) lo = mid; else hi = mid;
            }
            const int src = lo, d = dst[e];
            float in[U], out[U];
#pragma unroll
            for (int k = 0; k < U; k += 4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(Wst + (size_t)e * U + k));
                in[k] = v.x; in[k + 1] = v.y; in[k + 2] = v.z; in[k + 3] = v.w;
            }
            linear32(We, in, out);
#pragma unroll
            for (int k = 0; k < U; k += 4) {
                const float4 a3 = __ldcg(reinterpret_cast<const float4*>(X3 + (size_t)src * U + k));
                const float4 a4 = __ldcg(reinterpret_cast<const float4*>(X4 + (size_t)d * U + k));
                const float z[4] = {out[k] + a3.x + a4.x, out[k + 1] + a3.y + a4.y, out[k + 2] + a3.z + a4.z, out[k + 3] + a3.w + a4.w};
                float r[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int f = k + c;
                    const float y = (z[c] - bne[2 * U + f]) * bne[3 * U + f] * bne[f] + bne[U + f];
                    r[c] = in[f] + silu_f(y);
                }
                *reinterpret_cast<float4*>(Wst + (size_t)e * U + k) = make_float4(r[0], r[1], r[2], r[3]);
            }
        }
        // ---- node update: task = (node, feature)
        for (int t = tid; t < n * U; t += nth) {
            const int f = t % U;
            const float z = __ldcg(X1 + t) + __ldcg(AG + t);
            const float y = (z - bnv[2 * U + f]) * bnv[3 * U + f] * bnv[f] + bnv[U + f];
            X[t] = __ldcg(X + t) + silu_f(y);
        }
        __syncthreads();
    }
    if (p.dense_out) {          // background of the dense matrix: 0 + eps off-graph
        float* M = p.dense_out + (size_t)b * n * n;
        for (int i = tid; i < n * n; i += nth) M[i] = p.dense_eps;
        __syncthreads();
    }
    // ---- head MLP per edge
    const float* H0 = head, *H1 = head + (U * U + U), *H2 = head + 2 * (U * U + U);
    const int32_t* order = p.order + (size_t)b * E;
    for (int e = tid; e < E; e += nth) {
        float in[U], h[U];
#pragma unroll
        for (int k = 0; k < U; k += 4) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(Wst + (size_t)e * U + k));
            in[k] = v.x; in[k + 1] = v.y; in[k + 2] = v.z; in[k + 3] = v.w;
        }
        linear32(H0, in, h);
#pragma unroll
        for (int k = 0; k < U; ++k) in[k] = silu_f(h[k]);
        linear32(H1, in, h);
        float acc = H2[U];
#pragma unroll
        for (int k = 0; k < U; ++k) acc = fmaf(H2[k], silu_f(h[k]), acc);
        const float hv = sigmoid_f(acc);
        if (p.out) p.out[(size_t)b * E + order[e]] = hv;
        if (p.dense_out) {
            int lo = 0, hi = n;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (rp[mid] <= e) lo = mid; else hi = mid;
            }
            p.dense_out[((size_t)b * n + lo) * n + dst[e]] = hv + p.dense_eps;
        }
    }
}

}  // namespace deepaco
