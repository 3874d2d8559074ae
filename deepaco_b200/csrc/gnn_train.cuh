// K3t -- heuristic network in TRAINING mode: forward with batch-statistics BatchNorm + analytic backward
// (reference tsp/net.py:27-45 EmbNet.forward, :62-75 MLP/ParNet.forward as driven by train_instance,
// tsp/train.ipynb cell 1 / tsp_nls/train.py:15-44; PyG BatchNorm == nn.BatchNorm1d over all nodes / all edges of
// the one graph of a forward call, biased variance, eps 1e-5).
//
// One GROUP of CTAs per instance: a thread-block cluster (2, 4, 8 CTAs; hardware barrier.cluster) or, for large graphs,
// 16 / 32 / 64 co-resident CTAs of a cooperative launch synchronised through a per-instance arrival counter.
// Rows (nodes / edges) are split over the group's threads, the per-feature BatchNorm reductions go CTA (shared memory,
// fixed order) -> group (global scratch, fixed rank order) so results do not depend on timing.  No atomics on data:
// the gather side of every scatter is walked through the CSR (by source) / CSC (by destination) edge lists.
// The 32x32 linears are fp32 FMAs with the weights broadcast from shared memory; weight gradients are
// tile-staged outer products (4x4 register blocks) reduced in a fixed order.
//
// The kernel bodies only use threadIdx/blockIdx/blockDim, __syncthreads and the five helpers below, so the same
// source also compiles for the host thread-per-CUDA-thread harness in tests/cpu_emu/ (test infrastructure:
// checks indexing, phase ordering and races under ThreadSanitizer without a GPU; never part of the product).
#pragma once
#ifndef DEEPACO_CPU_EMU
#include <cuda_runtime.h>
#include <stdint.h>
#define DACO_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

namespace deepaco {
namespace gnnt {

constexpr int U = 32;                          // units
constexpr int LIN = U * U + U;                 // one 32x32 linear: W[out][in] then b[out]
constexpr int kLayerFloats = 5 * LIN + 8 * U;  // 4 node linears | edge linear | v_bn (g, b, -, -) | e_bn (g, b, -, -)
constexpr int kDepth = 12;
constexpr int kHeadFloats = 2 * LIN + U + 1;
constexpr int TS = 36;                         // tile row stride in floats (16-byte aligned, conflict-free float4 rows)
constexpr int kRedStride = 128;                // floats per (slot, rank) of the group reduction scratch
constexpr int kRedSlots = 3 * kDepth;          // forward uses 2 per layer, backward 1 per layer
constexpr int kMaxCtas = 64;
constexpr int kStatFloats = 6 * U;             // per layer: mean_v, invstd_v, var_v, mean_e, invstd_e, var_e

#ifndef DEEPACO_CPU_EMU
__device__ __forceinline__ unsigned cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cta_count() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {   // all threads of all CTAs of the cluster; release / acquire
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// arrival-counter barrier over the `ncta` co-resident CTAs of one instance (cooperative launch); *ctr starts at 0 and
// only grows: the k-th barrier completes when it reaches k * ncta.
__device__ __forceinline__ void counter_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
#endif

struct TrainParams {
    // graph, per instance, edges sorted by source node (stable)
    const float* x_in;        // [B][n][feats]
    const int32_t* row_ptr;   // [B][n+1]  CSR by source
    const int32_t* src;       // [B][E]    source of sorted edge
    const int32_t* dst;       // [B][E]    destination of sorted edge
    const float* attr;        // [B][E]
    const int32_t* order;     // [B][E]    original edge id of sorted edge
    const int32_t* col_ptr;   // [B][n+1]  CSC by destination (backward only)
    const int32_t* in_edges;  // [B][E]    sorted-edge ids grouped by destination (backward only)
    const float* weights;     // packed as net.py:pack_weights (mean / invstd slots of the BN blocks unused)
    // saved by the forward pass for the backward pass
    float* XS;                // [B][13][n][32]  x_l   (input of layer l; x_12 = output)
    float* WS;                // [B][13][E][32]  w_l
    float* ZV;                // [B][12][n][32]  pre-BatchNorm node activations
    float* ZE;                // [B][12][E][32]  pre-BatchNorm edge activations
    float* stats;             // [B][12][6][32]  batch mean / invstd / biased var, nodes then edges
    float* node_ws;           // [B][n][7*32] scratch: fwd x1|x2|x3|x4 ; bwd x2|GX|GZV|G1|G2|G3|G4
    float* edge_ws;           // [B][E][3*32] scratch (backward): GW | GZ | GM
    float* red;               // [B][kRedSlots][kMaxCtas][kRedStride] group reduction scratch
    unsigned* sync_ctr;       // [B] arrival counters (zero before the launch); used when grid_ctas > 0
    int grid_ctas;            // 0: the group is the thread-block cluster; > 0: CTAs per instance of a cooperative launch
    int b0;                   // first instance of this launch
    float* out;               // forward:  [B][E] heuristic per ORIGINAL edge id
    const float* g_out;       // backward: [B][E] gradient w.r.t. `out`
    float* grad_w;            // backward: [B][ctas][weight_count] partial parameter gradients (zero-initialised by the caller)
    int n, E, feats;
    float bn_eps;
    long long wc;             // weight_count
};

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float dsilu_f(float v) {    // d/dv [v * sigmoid(v)]
    const float s = sigmoid_f(v);
    return s * fmaf(v, 1.0f - s, 1.0f);
}

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
// out[k] = sum_o W[o][k] * g[o]
__device__ __forceinline__ void linear32_t(const float* __restrict__ Wb, const float (&g)[U], float (&out)[U]) {
#pragma unroll
    for (int k = 0; k < U; ++k) out[k] = 0.f;
#pragma unroll
    for (int o = 0; o < U; ++o) {
#pragma unroll
        for (int k = 0; k < U; ++k) out[k] = fmaf(Wb[o * U + k], g[o], out[k]);
    }
}
__device__ __forceinline__ void load_row(const float* __restrict__ src, float (&v)[U]) {
#pragma unroll
    for (int k = 0; k < U; k += 4) {
        const float4 t = ld_cg4(src + k);
        v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&v)[U]) {
#pragma unroll
    for (int k = 0; k < U; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}

struct Cta {
    int tid, nth;
    unsigned rank, ncta;
    int gt, gn;          // group-wide thread index / thread count
    int b;               // instance
    unsigned* ctr;       // arrival counter of this instance (cooperative-launch groups), else NULL
    unsigned epoch;      // barriers passed so far
    __device__ __forceinline__ void init(const int grid_ctas, unsigned* sync_ctr, int b0) {
        tid = threadIdx.x; nth = blockDim.x;
        if (grid_ctas > 0) { ncta = (unsigned)grid_ctas; rank = blockIdx.x % ncta; }
        else { ncta = cta_count(); rank = cta_rank(); }
        b = b0 + (int)(blockIdx.x / ncta);
        ctr = grid_ctas > 0 ? sync_ctr + b : nullptr;
        epoch = 0;
        gt = (int)rank * nth + tid; gn = (int)ncta * nth;
    }
    // all threads of all CTAs of the group; orders global-memory accesses across it
    __device__ __forceinline__ void sync_all() {
        if (ncta == 1) __syncthreads();
        else if (ctr) counter_barrier(ctr, ++epoch * ncta);
        else cluster_barrier();
    }
};

// ---- per-feature reductions ------------------------------------------------------------------------------------
// every thread contributes a 32-vector; dst[32] (shared memory) = sum over the CTA's threads, fixed order.
// stage: >= nth * 33 floats, part: >= nth floats.  Ends with a __syncthreads (dst readable by all).
__device__ __forceinline__ void cta_colsum_vec(const Cta& c, const float (&v)[U], float* stage, float* part, float* dst) {
    __syncthreads();                               // stage / part may still be read by a previous reduction
#pragma unroll
    for (int k = 0; k < U; ++k) stage[c.tid * (U + 1) + k] = v[k];
    __syncthreads();
    {
        const int g = c.tid >> 5, f = c.tid & 31;
        float s = 0.f;
        for (int r = 0; r < 32; ++r) s += stage[(g * 32 + r) * (U + 1) + f];
        part[c.tid] = s;
    }
    __syncthreads();
    if (c.tid < U) {
        float t = 0.f;
        for (int g = 0; g < (c.nth >> 5); ++g) t += part[g * 32 + c.tid];
        dst[c.tid] = t;
    }
    __syncthreads();
}
// thread tid contributes one value of feature tid % 32
__device__ __forceinline__ void cta_colsum_scalar(const Cta& c, float v, float* part, float* dst) {
    __syncthreads();
    part[c.tid] = v;
    __syncthreads();
    if (c.tid < U) {
        float t = 0.f;
        for (int g = 0; g < (c.nth >> 5); ++g) t += part[g * 32 + c.tid];
        dst[c.tid] = t;
    }
    __syncthreads();
}
// S[K*32] holds this CTA's sums; on return it holds the group-wide sums (identical bits in every CTA).
__device__ __forceinline__ void cluster_sum(Cta& c, float* S, int K, float* red_slot) {
    if (c.ncta == 1) return;
    if (c.tid < K * U) red_slot[c.rank * kRedStride + c.tid] = S[c.tid];
    c.sync_all();
    if (c.tid < K * U) {
        float t = 0.f;
        for (unsigned r = 0; r < c.ncta; ++r) t += ld_cg(red_slot + r * kRedStride + c.tid);
        S[c.tid] = t;
    }
    __syncthreads();
}

// ---- weight-gradient outer products ----------------------------------------------------------------------------
// dW[o][k] += sum_r TG[r][o] * TX[r][k], db[o] += sum_r TG[r][o] over `rows` tile rows (tiles in shared memory, row
// stride TS).  Thread = 4x4 block `cell` (64 cells) x row group (nth/64 groups).
struct OuterAcc {
    float a[16];
    float b[4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = 0.f;
    }
};
__device__ __forceinline__ void outer_accum(const Cta& c, const float* TG, const float* TX, int rows, OuterAcc& acc) {
    const int cell = c.tid & 63, grp = c.tid >> 6, ngrp = c.nth >> 6;
    const int ob = (cell >> 3) * 4, kb = (cell & 7) * 4;
    const bool bias = (cell & 7) == 0;
    for (int r = grp; r < rows; r += ngrp) {
        const float4 g = *reinterpret_cast<const float4*>(TG + r * TS + ob);
        const float4 x = *reinterpret_cast<const float4*>(TX + r * TS + kb);
        const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc.a[i * 4 + j] = fmaf(gv[i], xv[j], acc.a[i * 4 + j]);
        }
        if (bias) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc.b[i] += gv[i];
        }
    }
}
// reduce the row groups in order and STORE dW (1024) + db (32) to gdst; scratch >= (nth/64) * LIN floats of shared memory
__device__ __forceinline__ void outer_flush(const Cta& c, OuterAcc& acc, float* scratch, float* gdst) {
    const int cell = c.tid & 63, grp = c.tid >> 6, ngrp = c.nth >> 6;
    const int ob = (cell >> 3) * 4, kb = (cell & 7) * 4;
    __syncthreads();                               // the tiles (which scratch may alias) are no longer read
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) scratch[grp * LIN + (ob + i) * U + kb + j] = acc.a[i * 4 + j];
    }
    if ((cell & 7) == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) scratch[grp * LIN + U * U + ob + i] = acc.b[i];
    }
    __syncthreads();
    for (int idx = c.tid; idx < LIN; idx += c.nth) {
        float t = 0.f;
        for (int g = 0; g < ngrp; ++g) t += scratch[g * LIN + idx];
        gdst[idx] = t;
    }
    __syncthreads();
    acc.clear();
}
__device__ __forceinline__ void tile_put(float* T, int row, const float (&v)[U]) {
#pragma unroll
    for (int k = 0; k < U; k += 4) *reinterpret_cast<float4*>(T + row * TS + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}

// shared-memory carve-up (floats).  fwd: stage = nth*33;  bwd: two tiles of nth rows (stage and the flush scratch alias them)
struct Smem {
    float *wl, *head, *w0s, *S, *part, *tile;
};
__device__ __forceinline__ Smem carve(float* sm, int nth) {
    Smem s;
    s.wl = sm;
    s.head = s.wl + kLayerFloats;               // 5536
    s.w0s = s.head + 2176;                      // kHeadFloats = 2145, padded
    s.S = s.w0s + 352;                          // v_lin0 (<= 32*8 + 32) + e_lin0 (64)
    s.part = s.S + 8 * U;
    s.tile = s.part + nth;                      // 16-byte aligned: all sizes above are multiples of 4 floats
    return s;
}
inline size_t smem_floats_fwd(int nth) { return (size_t)kLayerFloats + 2176 + 352 + 8 * U + nth + (size_t)nth * (U + 1); }
inline size_t smem_floats_bwd(int nth) {
    size_t tiles = (size_t)2 * nth * TS, stage = (size_t)nth * (U + 1), flush = (size_t)(nth / 64) * LIN;
    size_t m = tiles > stage ? tiles : stage;
    if (flush > m) m = flush;
    return (size_t)kLayerFloats + 2176 + 352 + 8 * U + nth + m;
}

// =================================================================================================================
// forward.  TRAIN: batch-statistics BatchNorm, every activation the backward needs is saved (XS / WS hold all 13
// layers, ZV / ZE the pre-BatchNorm values).  !TRAIN (eval mode, deepaco_gnn_forward_group): running statistics from
// the mean / invstd slots of the packed weights, so a layer is two phases instead of four with no group reduction;
// XS / WS are two-layer ping-pong buffers; ZV / ZE / stats / red are not touched.  The point of the eval variant is
// latency: the reference's inference drivers run one instance at a time (tsp/test.ipynb infer_instance), and one
// instance on one CTA (gnn.cu) leaves 147 SMs idle.
// =================================================================================================================
template <bool TRAIN>
__global__ void __launch_bounds__(512) gnn_group_forward_kernel(const TrainParams p) {
    DACO_DYN_SMEM(smem_raw);
    Cta c;
    c.init(p.grid_ctas, p.sync_ctr, p.b0);
    const int b = c.b;
    const Smem s = carve(reinterpret_cast<float*>(smem_raw), c.nth);
    const int n = p.n, E = p.E, F = p.feats, tid = c.tid, nth = c.nth;
    const int32_t* rp = p.row_ptr + (size_t)b * (n + 1);
    const int32_t* srcs = p.src + (size_t)b * E;
    const int32_t* dsts = p.dst + (size_t)b * E;
    constexpr int kKept = TRAIN ? kDepth + 1 : 2;              // layers of x / w kept per instance
    float* XS = p.XS + (size_t)b * kKept * n * U;
    float* WS = p.WS + (size_t)b * kKept * E * U;
    float* ZV = TRAIN ? p.ZV + (size_t)b * kDepth * n * U : nullptr;
    float* ZE = TRAIN ? p.ZE + (size_t)b * kDepth * E * U : nullptr;
    float* NW = p.node_ws + (size_t)b * n * (TRAIN ? 7 : 4) * U;
    float* X1 = NW, *X2 = NW + (size_t)n * U, *X3 = NW + (size_t)2 * n * U, *X4 = NW + (size_t)3 * n * U;
    float* red = TRAIN ? p.red + (size_t)b * kRedSlots * kMaxCtas * kRedStride : nullptr;
    const int off_layers = U * F + 3 * U;
    const float* layers_g = p.weights + off_layers;
    const float* head_g = layers_g + (size_t)kDepth * kLayerFloats;
    const float inv_n = 1.0f / (float)n, inv_E = 1.0f / (float)E;
    // pre-BatchNorm node activation of (node i, feature f): x1 + mean over the out-edges of sigmoid(w) * x2[dst]
    auto node_pre = [&](const float* Wl, int i, int f) -> float {
        const int e0 = rp[i], e1 = rp[i + 1];
        float a = 0.f;
        int e = e0;
        for (; e + 8 <= e1; e += 8) {                          // 8 independent gathers in flight; same summation order
            int d[8];
            float wv[8], xv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = dsts[e + j];
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = ld_cg(Wl + (size_t)(e + j) * U + f);
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = ld_cg(X2 + (size_t)d[j] * U + f);
#pragma unroll
            for (int j = 0; j < 8; ++j) a += sigmoid_f(wv[j]) * xv[j];
        }
        for (; e < e1; ++e) a += sigmoid_f(ld_cg(Wl + (size_t)e * U + f)) * ld_cg(X2 + (size_t)dsts[e] * U + f);
        const int deg = e1 - e0;
        return ld_cg(X1 + (size_t)i * U + f) + a / (float)(deg > 0 ? deg : 1);
    };

    for (int i = tid; i < off_layers; i += nth) s.w0s[i] = p.weights[i];
    for (int i = tid; i < kHeadFloats; i += nth) s.head[i] = head_g[i];
    __syncthreads();
    // ---- input embeddings (net.py:28-30)
    for (int t = c.gt; t < n * U; t += c.gn) {
        const int i = t / U, o = t % U;
        float acc = s.w0s[U * F + o];
        for (int k = 0; k < F; ++k) acc = fmaf(s.w0s[o * F + k], p.x_in[((size_t)b * n + i) * F + k], acc);
        XS[t] = silu_f(acc);
    }
    {
        const float* We0 = s.w0s + U * F + U;
        const float* attr = p.attr + (size_t)b * E;
        for (int t = c.gt; t < E * U; t += c.gn) {
            const int e = t / U, o = t % U;
            WS[t] = silu_f(fmaf(We0[o], attr[e], We0[U + o]));
        }
    }

    for (int l = 0; l < kDepth; ++l) {
        const float* Xl = XS + (size_t)(TRAIN ? l : l & 1) * n * U;
        float* Xn = XS + (size_t)(TRAIN ? l + 1 : (l + 1) & 1) * n * U;
        const float* Wl = WS + (size_t)(TRAIN ? l : l & 1) * E * U;
        float* Wn = WS + (size_t)(TRAIN ? l + 1 : (l + 1) & 1) * E * U;
        float* Zv = TRAIN ? ZV + (size_t)l * n * U : nullptr;
        float* Ze = TRAIN ? ZE + (size_t)l * E * U : nullptr;
        // eval: x_12 is never read (EmbNet.forward returns w, tsp/net.py:45) -> no node update in the last layer
        const bool node_update = TRAIN || l + 1 < kDepth;
        c.sync_all();                                          // x_l, w_l complete; s.wl no longer read
        for (int i = tid; i < kLayerFloats; i += nth) s.wl[i] = layers_g[(size_t)l * kLayerFloats + i];
        __syncthreads();
        const float* We = s.wl + 4 * LIN;
        const float* bnv = We + LIN;                           // gamma, beta
        const float* bne = bnv + 4 * U;
        // ---- P1 node linears: task = (node, which linear)
        for (int t = c.gt; t < n * 4; t += c.gn) {
            const int i = t >> 2, q = t & 3;
            if (!node_update && q < 2) continue;               // x1, x2 only feed the node update
            float in[U], out[U];
            load_row(Xl + (size_t)i * U, in);
            linear32(s.wl + q * LIN, in, out);
            store_row((q == 0 ? X1 : q == 1 ? X2 : q == 2 ? X3 : X4) + (size_t)i * U, out);
        }
        c.sync_all();
        if constexpr (!TRAIN) {
            // ---- eval: pre-activation, BatchNorm with the running statistics, activation and residual in one phase
            const float* mean_v = bnv + 2 * U, *istd_v = bnv + 3 * U, *mean_e = bne + 2 * U, *istd_e = bne + 3 * U;
            if (node_update) {
                for (int t = c.gt; t < n * U; t += c.gn) {
                    const int i = t / U, f = t % U;
                    const float z = node_pre(Wl, i, f);
                    // same expression as gnn.cu (no fused multiply-add): both eval kernels return the same bits
                    Xn[t] = ld_cg(Xl + t) + silu_f((z - mean_v[f]) * istd_v[f] * bnv[f] + bnv[U + f]);
                }
            }
            for (int e = c.gt; e < E; e += c.gn) {
                float w[U], z[U], a3[U], a4[U];
                load_row(Wl + (size_t)e * U, w);
                linear32(We, w, z);
                load_row(X3 + (size_t)srcs[e] * U, a3);
                load_row(X4 + (size_t)dsts[e] * U, a4);
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const float zk = z[k] + a3[k] + a4[k];
                    w[k] += silu_f((zk - mean_e[k]) * istd_e[k] * bne[k] + bne[U + k]);
                }
                store_row(Wn + (size_t)e * U, w);
            }
            continue;                                          // the barrier at the top of the next iteration orders the stores
        }
        // ---- P2 pre-BatchNorm activations + per-feature sums
        float sv = 0.f;
        for (int t = c.gt; t < n * U; t += c.gn) {             // gn is a multiple of 32: feature = tid % 32
            const float z = node_pre(Wl, t / U, t % U);
            Zv[t] = z;
            sv += z;
        }
        float se[U];
#pragma unroll
        for (int k = 0; k < U; ++k) se[k] = 0.f;
        for (int e = c.gt; e < E; e += c.gn) {
            float in[U], z[U], a3[U], a4[U];
            load_row(Wl + (size_t)e * U, in);
            linear32(We, in, z);
            load_row(X3 + (size_t)srcs[e] * U, a3);
            load_row(X4 + (size_t)dsts[e] * U, a4);
#pragma unroll
            for (int k = 0; k < U; ++k) { z[k] = z[k] + a3[k] + a4[k]; se[k] += z[k]; }
            store_row(Ze + (size_t)e * U, z);
        }
        cta_colsum_scalar(c, sv, s.part, s.S);
        cta_colsum_vec(c, se, s.tile, s.part, s.S + U);
        cluster_sum(c, s.S, 2, red + (size_t)(2 * l) * kMaxCtas * kRedStride);
        if (tid < 2 * U) s.S[2 * U + tid] = s.S[tid] * (tid < U ? inv_n : inv_E);     // means -> S[64..127]
        __syncthreads();
        const float* mean_v = s.S + 2 * U, *mean_e = s.S + 3 * U;
        // ---- P3 biased variance (second pass over this thread's own rows)
        sv = 0.f;
        for (int t = c.gt; t < n * U; t += c.gn) {
            const float d = Zv[t] - mean_v[t % U];
            sv = fmaf(d, d, sv);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) se[k] = 0.f;
        for (int e = c.gt; e < E; e += c.gn) {
            float z[U];
            load_row(Ze + (size_t)e * U, z);
#pragma unroll
            for (int k = 0; k < U; ++k) { const float d = z[k] - mean_e[k]; se[k] = fmaf(d, d, se[k]); }
        }
        cta_colsum_scalar(c, sv, s.part, s.S);
        cta_colsum_vec(c, se, s.tile, s.part, s.S + U);
        cluster_sum(c, s.S, 2, red + (size_t)(2 * l + 1) * kMaxCtas * kRedStride);
        if (tid < 2 * U) {
            const float var = s.S[tid] * (tid < U ? inv_n : inv_E);
            const float istd = 1.0f / sqrtf(var + p.bn_eps);
            s.S[4 * U + tid] = istd;                                                  // invstd -> S[128..191]
            if (c.rank == 0) {
                float* st = p.stats + ((size_t)b * kDepth + l) * kStatFloats + (tid < U ? 0 : 3 * U) + (tid & 31);
                st[0] = s.S[2 * U + tid];
                st[U] = istd;
                st[2 * U] = var;
            }
        }
        __syncthreads();
        const float* istd_v = s.S + 4 * U, *istd_e = s.S + 5 * U;
        // ---- P4 normalise, activate, residual (net.py:41-44)
        for (int t = c.gt; t < n * U; t += c.gn) {
            const int f = t % U;
            const float y = fmaf((Zv[t] - mean_v[f]) * istd_v[f], bnv[f], bnv[U + f]);
            Xn[t] = ld_cg(Xl + t) + silu_f(y);
        }
        for (int e = c.gt; e < E; e += c.gn) {
            float z[U], w[U];
            load_row(Ze + (size_t)e * U, z);
            load_row(Wl + (size_t)e * U, w);
#pragma unroll
            for (int k = 0; k < U; ++k) w[k] += silu_f(fmaf((z[k] - mean_e[k]) * istd_e[k], bne[k], bne[U + k]));
            store_row(Wn + (size_t)e * U, w);
        }
        // the barrier at the top of the next iteration orders these stores before their readers
    }
    // ---- head MLP per edge (net.py:62-75); every thread reads only the w_12 rows it wrote itself
    const float* H0 = s.head, *H1 = s.head + LIN, *H2 = s.head + 2 * LIN;
    const int32_t* order = p.order + (size_t)b * E;
    const float* W12 = WS + (size_t)(TRAIN ? kDepth : kDepth & 1) * E * U;
    for (int e = c.gt; e < E; e += c.gn) {
        float in[U], h[U];
        load_row(W12 + (size_t)e * U, in);
        linear32(H0, in, h);
#pragma unroll
        for (int k = 0; k < U; ++k) in[k] = silu_f(h[k]);
        linear32(H1, in, h);
        float acc = H2[U];
#pragma unroll
        for (int k = 0; k < U; ++k) acc = fmaf(H2[k], silu_f(h[k]), acc);
        p.out[(size_t)b * E + order[e]] = sigmoid_f(acc);
    }
}

// =================================================================================================================
// backward: gradients of sum(out * g_out) w.r.t. every parameter
// =================================================================================================================
__global__ void __launch_bounds__(256) gnn_train_backward_kernel(const TrainParams p) {
    DACO_DYN_SMEM(smem_raw);
    Cta c;
    c.init(p.grid_ctas, p.sync_ctr, p.b0);
    const int b = c.b;
    const Smem s = carve(reinterpret_cast<float*>(smem_raw), c.nth);
    const int n = p.n, E = p.E, F = p.feats, tid = c.tid, nth = c.nth;
    const int32_t* rp = p.row_ptr + (size_t)b * (n + 1);
    const int32_t* srcs = p.src + (size_t)b * E;
    const int32_t* dsts = p.dst + (size_t)b * E;
    const int32_t* cp = p.col_ptr + (size_t)b * (n + 1);
    const int32_t* ine = p.in_edges + (size_t)b * E;
    const float* XS = p.XS + (size_t)b * (kDepth + 1) * n * U;
    const float* WS = p.WS + (size_t)b * (kDepth + 1) * E * U;
    const float* ZV = p.ZV + (size_t)b * kDepth * n * U;
    const float* ZE = p.ZE + (size_t)b * kDepth * E * U;
    float* NW = p.node_ws + (size_t)b * n * 7 * U;
    float* X2 = NW, *GX = NW + (size_t)n * U, *GZV = NW + (size_t)2 * n * U, *G14 = NW + (size_t)3 * n * U;   // G14: [4][n][32]
    float* GW = p.edge_ws + (size_t)b * E * 3 * U;
    float* GZ = GW + (size_t)E * U, *GM = GW + (size_t)2 * E * U;
    float* red = p.red + (size_t)b * kRedSlots * kMaxCtas * kRedStride;
    float* grad = p.grad_w + ((size_t)b * c.ncta + c.rank) * p.wc;     // this CTA's partial gradient buffer
    const int off_layers = U * F + 3 * U;
    const float* layers_g = p.weights + off_layers;
    const float* head_g = layers_g + (size_t)kDepth * kLayerFloats;
    float* grad_layers = grad + off_layers;
    float* grad_head = grad_layers + (size_t)kDepth * kLayerFloats;
    const float inv_n = 1.0f / (float)n, inv_E = 1.0f / (float)E;
    float* TG = s.tile, *TX = s.tile + (size_t)nth * TS;
    OuterAcc accA, accB;
    accA.clear(); accB.clear();

    for (int i = tid; i < off_layers; i += nth) s.w0s[i] = p.weights[i];
    for (int i = tid; i < kHeadFloats; i += nth) s.head[i] = head_g[i];
    for (int t = c.gt; t < n * U; t += c.gn) GX[t] = 0.f;      // x_12 does not reach the output (net.py:45 returns w)
    __syncthreads();

    // ---- head backward: thread = edge, in tiles of nth edges (uniform trip count: the tile syncs are CTA-wide)
    {
        const float* H0 = s.head, *H1 = s.head + LIN, *H2 = s.head + 2 * LIN;
        const int32_t* order = p.order + (size_t)b * E;
        const float* W12 = WS + (size_t)kDepth * E * U;
        float acc_h2[U], acc_b2 = 0.f;
#pragma unroll
        for (int k = 0; k < U; ++k) acc_h2[k] = 0.f;
        for (int base = (int)c.rank * nth; base < E; base += c.gn) {
            const int e = base + tid;
            const bool on = e < E;
            const int rows = (E - base) < nth ? (E - base) : nth;
            float w[U], h0[U], t0[U], h1[U], g1[U];
            if (on) load_row(W12 + (size_t)e * U, w);
            else {
#pragma unroll
                for (int k = 0; k < U; ++k) w[k] = 0.f;
            }
            linear32(H0, w, h0);
#pragma unroll
            for (int k = 0; k < U; ++k) t0[k] = silu_f(h0[k]);
            linear32(H1, t0, h1);
            float a = H2[U];
#pragma unroll
            for (int k = 0; k < U; ++k) a = fmaf(H2[k], silu_f(h1[k]), a);
            const float heu = sigmoid_f(a);
            const float ga = on ? p.g_out[(size_t)b * E + order[e]] * heu * (1.0f - heu) : 0.f;
            acc_b2 += ga;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                acc_h2[k] = fmaf(ga, silu_f(h1[k]), acc_h2[k]);
                g1[k] = ga * H2[k] * dsilu_f(h1[k]);                       // d/d h1
            }
            tile_put(TG, tid, g1);
            tile_put(TX, tid, t0);
            __syncthreads();
            outer_accum(c, TG, TX, rows, accA);                            // dH1 += g1 (x) silu(h0)
            __syncthreads();
            linear32_t(H1, g1, t0);                                        // d/d silu(h0)
#pragma unroll
            for (int k = 0; k < U; ++k) g1[k] = t0[k] * dsilu_f(h0[k]);    // d/d h0
            tile_put(TG, tid, g1);
            tile_put(TX, tid, w);
            __syncthreads();
            outer_accum(c, TG, TX, rows, accB);                            // dH0 += g0 (x) w_12
            __syncthreads();
            linear32_t(H0, g1, t0);
            if (on) store_row(GW + (size_t)e * U, t0);                     // d/d w_12
        }
        outer_flush(c, accA, s.tile, grad_head + LIN);
        outer_flush(c, accB, s.tile, grad_head);
        cta_colsum_vec(c, acc_h2, s.tile, s.part, s.S);
#pragma unroll
        for (int k = 0; k < U; ++k) acc_h2[k] = k == 0 ? acc_b2 : 0.f;
        cta_colsum_vec(c, acc_h2, s.tile, s.part, s.S + U);
        if (tid < U) grad_head[2 * LIN + tid] = s.S[tid];
        if (tid == 0) grad_head[2 * LIN + U] = s.S[U];
    }

    for (int l = kDepth - 1; l >= 0; --l) {
        const float* Xl = XS + (size_t)l * n * U;
        const float* Wl = WS + (size_t)l * E * U;
        const float* Zv = ZV + (size_t)l * n * U;
        const float* Ze = ZE + (size_t)l * E * U;
        const float* st = p.stats + ((size_t)b * kDepth + l) * kStatFloats;
        float* gl = grad_layers + (size_t)l * kLayerFloats;
        c.sync_all();                                          // GX / GW of layer l+1 complete; s.wl free
        for (int i = tid; i < kLayerFloats; i += nth) s.wl[i] = layers_g[(size_t)l * kLayerFloats + i];
        if (tid < 2 * U) {                                     // S[256+..]: mean_v, mean_e ; S[320+..]: invstd_v, invstd_e
            s.S[4 * U + tid] = st[(tid < U ? 0 : 3 * U) + (tid & 31)];
            s.S[6 * U + tid] = st[(tid < U ? 0 : 3 * U) + U + (tid & 31)];
        }
        __syncthreads();
        const float* We = s.wl + 4 * LIN;
        const float* bnv = We + LIN, *bne = bnv + 4 * U;
        const float* mean_v = s.S + 4 * U, *mean_e = s.S + 5 * U, *istd_v = s.S + 6 * U, *istd_e = s.S + 7 * U;
        // ---- A: x2 = Lin2(x_l) recomputed; gy = g * silu'(BN(z)); per-feature sums of gy and gy * xhat
        for (int i = c.gt; i < n; i += c.gn) {
            float in[U], out[U];
            load_row(Xl + (size_t)i * U, in);
            linear32(s.wl + LIN, in, out);
            store_row(X2 + (size_t)i * U, out);
        }
        float v1 = 0.f, v2 = 0.f;
        for (int t = c.gt; t < n * U; t += c.gn) {
            const int f = t % U;
            const float xh = (Zv[t] - mean_v[f]) * istd_v[f];
            const float gy = ld_cg(GX + t) * dsilu_f(fmaf(xh, bnv[f], bnv[U + f]));
            GZV[t] = gy;
            v1 += gy;
            v2 = fmaf(gy, xh, v2);
        }
        float e1[U], e2[U];
#pragma unroll
        for (int k = 0; k < U; ++k) { e1[k] = 0.f; e2[k] = 0.f; }
        for (int e = c.gt; e < E; e += c.gn) {
            float z[U], g[U];
            load_row(Ze + (size_t)e * U, z);
            load_row(GW + (size_t)e * U, g);
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float xh = (z[k] - mean_e[k]) * istd_e[k];
                g[k] = g[k] * dsilu_f(fmaf(xh, bne[k], bne[U + k]));
                e1[k] += g[k];
                e2[k] = fmaf(g[k], xh, e2[k]);
            }
            store_row(GZ + (size_t)e * U, g);
        }
        cta_colsum_scalar(c, v1, s.part, s.S);
        cta_colsum_scalar(c, v2, s.part, s.S + U);
        cta_colsum_vec(c, e1, s.tile, s.part, s.S + 2 * U);
        cta_colsum_vec(c, e2, s.tile, s.part, s.S + 3 * U);
        cluster_sum(c, s.S, 4, red + (size_t)(2 * kDepth + l) * kMaxCtas * kRedStride);
        // S[0..31] = sum gy_v (= d beta_v), S[32..63] = sum gy_v*xhat (= d gamma_v), S[64..], S[96..] same for edges
        if (c.rank == 0 && tid < U) {
            gl[5 * LIN + tid] = s.S[U + tid];
            gl[5 * LIN + U + tid] = s.S[tid];
            gl[5 * LIN + 4 * U + tid] = s.S[3 * U + tid];
            gl[5 * LIN + 5 * U + tid] = s.S[2 * U + tid];
        }
        // ---- B: BatchNorm backward (batch statistics): gz = gamma * invstd * (gy - mean(gy) - xhat * mean(gy * xhat))
        for (int t = c.gt; t < n * U; t += c.gn) {
            const int f = t % U;
            const float xh = (Zv[t] - mean_v[f]) * istd_v[f];
            GZV[t] = bnv[f] * istd_v[f] * (GZV[t] - s.S[f] * inv_n - xh * (s.S[U + f] * inv_n));
        }
        for (int base = (int)c.rank * nth; base < E; base += c.gn) {
            const int e = base + tid;
            const bool on = e < E;
            const int rows = (E - base) < nth ? (E - base) : nth;
            float z[U], g[U], w[U];
            if (on) {
                load_row(Ze + (size_t)e * U, z);
                load_row(GZ + (size_t)e * U, g);
                load_row(Wl + (size_t)e * U, w);
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const float xh = (z[k] - mean_e[k]) * istd_e[k];
                    g[k] = bne[k] * istd_e[k] * (g[k] - s.S[2 * U + k] * inv_E - xh * (s.S[3 * U + k] * inv_E));
                }
                store_row(GZ + (size_t)e * U, g);
            } else {
#pragma unroll
                for (int k = 0; k < U; ++k) { g[k] = 0.f; w[k] = 0.f; }
            }
            tile_put(TG, tid, g);
            tile_put(TX, tid, w);
            __syncthreads();
            outer_accum(c, TG, TX, rows, accA);                            // dWe += gz (x) w_l
            __syncthreads();
            if (on) {                                                      // residual + linear path of w_l's gradient
                linear32_t(We, g, z);
                load_row(GW + (size_t)e * U, g);
#pragma unroll
                for (int k = 0; k < U; ++k) g[k] += z[k];
                store_row(GW + (size_t)e * U, g);
            }
        }
        outer_flush(c, accA, s.tile, gl + 4 * LIN);
        c.sync_all();                                          // GZV, GZ, X2 complete
        // ---- C: gate path of w_l's gradient; node-side gathers of the four linear outputs' gradients
        for (int e = c.gt; e < E; e += c.gn) {
            const int i = srcs[e], j = dsts[e];
            const int deg = rp[i + 1] - rp[i];
            const float idg = 1.0f / (float)(deg > 0 ? deg : 1);
            float w[U], g[U], gi[U], xj[U];
            load_row(Wl + (size_t)e * U, w);
            load_row(GW + (size_t)e * U, g);
            load_row(GZV + (size_t)i * U, gi);
            load_row(X2 + (size_t)j * U, xj);
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float sg = sigmoid_f(w[k]), m = gi[k] * idg;
                g[k] = fmaf(m * xj[k], sg * (1.0f - sg), g[k]);
                gi[k] = m * sg;                                            // message gradient towards x2[dst]
            }
            store_row(GW + (size_t)e * U, g);
            store_row(GM + (size_t)e * U, gi);
        }
        c.sync_all();                                          // GM complete
        for (int t = c.gt; t < n * U; t += c.gn) {
            const int i = t / U, f = t % U;
            float g3 = 0.f, g4 = 0.f, g2 = 0.f;
            int e = rp[i];
            const int e1 = rp[i + 1], q1 = cp[i + 1];
            for (; e + 8 <= e1; e += 8) {                      // 8 independent loads in flight; same summation order
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = ld_cg(GZ + (size_t)(e + j) * U + f);
#pragma unroll
                for (int j = 0; j < 8; ++j) g3 += v[j];
            }
            for (; e < e1; ++e) g3 += ld_cg(GZ + (size_t)e * U + f);
            int q = cp[i];
            for (; q + 8 <= q1; q += 8) {
                int ie[8];
                float v[8], m[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) ie[j] = ine[q + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = ld_cg(GZ + (size_t)ie[j] * U + f); m[j] = ld_cg(GM + (size_t)ie[j] * U + f); }
#pragma unroll
                for (int j = 0; j < 8; ++j) { g4 += v[j]; g2 += m[j]; }
            }
            for (; q < q1; ++q) {
                const int ie = ine[q];
                g4 += ld_cg(GZ + (size_t)ie * U + f);
                g2 += ld_cg(GM + (size_t)ie * U + f);
            }
            G14[t] = ld_cg(GZV + t);
            G14[(size_t)n * U + t] = g2;
            G14[(size_t)2 * n * U + t] = g3;
            G14[(size_t)3 * n * U + t] = g4;
        }
        c.sync_all();                                          // G14 complete
        // ---- D: x_l's gradient through the four node linears (task = node x 8-feature slice) and their weight gradients
        for (int t = c.gt; t < n * 4; t += c.gn) {
            const int i = t >> 2, k0 = (t & 3) * 8;
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
            for (int q = 0; q < 4; ++q) {
                const float* Wq = s.wl + q * LIN;
                float g[U];
                load_row(G14 + ((size_t)q * n + i) * U, g);
#pragma unroll
                for (int o = 0; o < U; ++o) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(Wq[o * U + k0 + k], g[o], acc[k]);
                }
            }
            float* gx = GX + (size_t)i * U + k0;
#pragma unroll
            for (int k = 0; k < 8; ++k) gx[k] = ld_cg(gx + k) + acc[k];
        }
        for (int q = 0; q < 4; ++q) {
            for (int base = (int)c.rank * nth; base < n; base += c.gn) {
                const int rows = (n - base) < nth ? (n - base) : nth;
                __syncthreads();
                for (int idx = tid; idx < rows * 8; idx += nth) {          // 8 float4 per row, coalesced
                    const int r = idx >> 3, k = (idx & 7) * 4;
                    *reinterpret_cast<float4*>(TG + r * TS + k) = ld_cg4(G14 + ((size_t)q * n + base + r) * U + k);
                    *reinterpret_cast<float4*>(TX + r * TS + k) = ld_cg4(Xl + (size_t)(base + r) * U + k);
                }
                __syncthreads();
                outer_accum(c, TG, TX, rows, accA);
            }
            outer_flush(c, accA, s.tile, gl + q * LIN);
        }
    }
    c.sync_all();                                              // GX, GW now hold d/d x_0, d/d w_0
    // ---- input embeddings: x_0 = silu(Lin_v0 x_in), w_0 = silu(Lin_e0 attr)
    {
        float gb = 0.f, gw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) gw[k] = 0.f;
        for (int t = c.gt; t < n * U; t += c.gn) {
            const int i = t / U, o = t % U;
            const float* xi = p.x_in + ((size_t)b * n + i) * F;
            float pre = s.w0s[U * F + o];
            for (int k = 0; k < F; ++k) pre = fmaf(s.w0s[o * F + k], xi[k], pre);
            const float g = ld_cg(GX + t) * dsilu_f(pre);
            gb += g;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < F) gw[k] = fmaf(g, xi[k], gw[k]);
        }
        cta_colsum_scalar(c, gb, s.part, s.S);
        if (tid < U) grad[U * F + tid] = s.S[tid];
        for (int k = 0; k < F; ++k) {
            cta_colsum_scalar(c, gw[k], s.part, s.S);
            if (tid < U) grad[tid * F + k] = s.S[tid];
        }
        const float* We0 = s.w0s + U * F + U;
        const float* attr = p.attr + (size_t)b * E;
        float ew[U], eb[U];
#pragma unroll
        for (int k = 0; k < U; ++k) { ew[k] = 0.f; eb[k] = 0.f; }
        for (int e = c.gt; e < E; e += c.gn) {
            float g[U];
            load_row(GW + (size_t)e * U, g);
            const float a = attr[e];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float gp = g[k] * dsilu_f(fmaf(We0[k], a, We0[U + k]));
                eb[k] += gp;
                ew[k] = fmaf(gp, a, ew[k]);
            }
        }
        cta_colsum_vec(c, ew, s.tile, s.part, s.S);
        cta_colsum_vec(c, eb, s.tile, s.part, s.S + U);
        if (tid < 2 * U) grad[U * F + U + tid] = s.S[tid];
    }
}

}  // namespace gnnt
}  // namespace deepaco
