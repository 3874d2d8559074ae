// K1 -- TSP tour construction (one warp per ant, all n-1 steps inside one launch).
//
// Replaces the Python step loop of ACO.gen_path / pick_move (reference tsp/aco.py:134-177 and the
// fixed-start variant tsp_nls/aco.py:184-220): per step the reference gathers the pheromone and
// heuristic rows of the current node, multiplies them with the visited mask, normalises
// (`Categorical`), and draws with `torch.multinomial(probs, 1)` == argmax(probs / q), q ~ Exp(1)
// from `exponential_`.  Here a warp owns an ant: the product matrix is staged once per CTA into shared
// memory with TMA bulk copies, the row sum uses ATen's exact summation order (common.cuh), the Exp(1)
// noise is regenerated in registers from torch's Philox stream, and arg-max is two warp REDUX ops.
#include "common.cuh"
#include "host_util.h"

#include <stdlib.h>

namespace deepaco {

struct TspSampleParams {
    const float* ph;      // [B][n][n]
    const float* heu;     // [B][n][n] or null
    int n, A, B;
    int start_node;       // >= 0 fixed; -1 -> `start` tensor or torch randint stream
    int double_norm;
    uint64_t seed, offset;
    const uint64_t* rng;  // [B][2] or null
    const float* noise;   // [B][n-1][A][n] or null
    const int64_t* start; // [B][A] or null
    int64_t* paths;       // [B][n][A] or null
    float* logp;          // [B][n-1][A] or null
    uint16_t* tours;      // [B][A][n] or null
    int lbw;              // log2(ATen block_width) for the strided layout
    DrawGeom g_noise, g_start;
    uint32_t start_increment;
};

template <int EPL, bool VEC>
__device__ __forceinline__ uint32_t elem_index(int k, int lane, int lbw) {
    if (VEC) return 4u * (uint32_t)(lane + 32 * (k >> 2)) + (uint32_t)(k & 3);
    return (uint32_t)lane + ((uint32_t)k << lbw);
}

template <int EPL, bool VEC, bool SMEMP>
__global__ void __launch_bounds__(512) tsp_sample_kernel(const TspSampleParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const int n = p.n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int W = nthreads >> 5, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int a0 = blockIdx.x * W;
    const int a = a0 + warp;
    const size_t nn = (size_t)n * n;
    float* Psm = reinterpret_cast<float*>(smem);
    const size_t pbytes = SMEMP ? ((nn * 4 + 15) & ~(size_t)15) : 0;
    uint16_t* tour_all = reinterpret_cast<uint16_t*>(smem + pbytes);
    uint16_t* tour_sm = tour_all + (size_t)warp * n;

    if (SMEMP) {
        // ---- stage P = pheromone (.) heuristic of colony b into shared memory (TMA bulk + mbarrier)
        const float* src = p.ph + (size_t)b * nn;
        const uint32_t total = (uint32_t)(nn * 4);
        const uint32_t bulk = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (total & ~15u) : 0u;
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (tid == 0 && bulk) {
            mbar_expect_tx(&bar, bulk);
            constexpr uint32_t kChunk = 32768;
            for (uint32_t off = 0; off < bulk; off += kChunk) {
                const uint32_t sz = (bulk - off < kChunk) ? (bulk - off) : kChunk;
                tma_bulk_g2s(reinterpret_cast<char*>(Psm) + off, reinterpret_cast<const char*>(src) + off, sz, &bar);
            }
        }
        for (size_t i = bulk / 4 + tid; i < nn; i += nthreads) Psm[i] = src[i];
        if (bulk) mbar_wait(&bar, 0);
        __syncthreads();
        if (p.heu) {
            const float* h = p.heu + (size_t)b * nn;
            for (size_t i = tid; i < nn; i += nthreads) Psm[i] = __fmul_rn(Psm[i], __ldg(h + i));
            __syncthreads();
        }
    }

    if (a < p.A) {
        const int lbw = p.lbw;
        const bool lane_on = VEC || lane < (1 << lbw);
        const uint64_t seed = p.rng ? p.rng[2 * b] : p.seed;
        const uint64_t offset0 = p.rng ? p.rng[2 * b + 1] : p.offset;
        const float* Pg = SMEMP ? Psm : (p.ph + (size_t)b * nn);   // !SMEMP: caller passes the product in `ph`

        int cur;
        uint64_t off_noise = offset0;
        if (p.start_node >= 0) {
            cur = p.start_node;
        } else if (p.start) {
            cur = (int)p.start[(size_t)b * p.A + a];
        } else {
            // torch.randint(0, n, (A,)): element a <- curand4().x % n  (random_from_to_kernel, 32-bit branch)
            cur = (int)(torch_philox_word(seed, offset0, (uint64_t)a, p.g_start) % (uint32_t)n);
            off_noise += p.start_increment;
        }

        uint32_t vis = 0;
#pragma unroll
        for (int k = 0; k < EPL; ++k)
            if (elem_index<EPL, VEC>(k, lane, lbw) == (uint32_t)cur && lane_on) vis |= 1u << k;
        if (lane == 0) tour_sm[0] = (uint16_t)cur;

        for (int step = 0; step < n - 1; ++step) {
            const float* row = Pg + (size_t)cur * n;
            float x[EPL];
            if (VEC) {
#pragma unroll
                for (int m = 0; m < EPL / 4; ++m) {
                    const int vi = lane + 32 * m;
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (4 * vi < n) t = SMEMP ? *reinterpret_cast<const float4*>(row + 4 * vi)
                                              : __ldg(reinterpret_cast<const float4*>(row + 4 * vi));
                    x[4 * m + 0] = ((vis >> (4 * m + 0)) & 1u) ? 0.f : t.x;
                    x[4 * m + 1] = ((vis >> (4 * m + 1)) & 1u) ? 0.f : t.y;
                    x[4 * m + 2] = ((vis >> (4 * m + 2)) & 1u) ? 0.f : t.z;
                    x[4 * m + 3] = ((vis >> (4 * m + 3)) & 1u) ? 0.f : t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < EPL; ++k) {
                    const uint32_t j = elem_index<EPL, VEC>(k, lane, lbw);
                    const bool ok = lane_on && j < (uint32_t)n && !((vis >> k) & 1u);
                    x[k] = ok ? (SMEMP ? row[j] : __ldg(row + j)) : 0.f;
                }
            }
            float S = VEC ? aten_sum_vec4<EPL>(x) : aten_sum_strided<EPL>(x);
            if (p.double_norm) {   // tsp_nls/aco.py:206 then Categorical normalises again
#pragma unroll
                for (int k = 0; k < EPL; ++k) x[k] = __fdiv_rn(x[k], S);
                S = VEC ? aten_sum_vec4<EPL>(x) : aten_sum_strided<EPL>(x);
            }

            float best = 0.f, bestp = 0.f;
            uint32_t bestj = 0xffffffffu;
            const uint64_t off_step = off_noise + 4ull * (uint64_t)step;
            const float* nz = p.noise ? p.noise + (((size_t)b * (n - 1) + step) * p.A + a) * (size_t)n : nullptr;
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const uint32_t j = elem_index<EPL, VEC>(k, lane, lbw);
                if (lane_on && j < (uint32_t)n) {
                    const float pn = __fdiv_rn(x[k], S);
                    const float q = nz ? nz[j]
                                       : exp1_from_word(torch_philox_word(seed, off_step, (uint64_t)a * n + j, p.g_noise));
                    const float v = __fdiv_rn(pn, q);
                    if (bestj == 0xffffffffu || v > best) {
                        best = v;
                        bestj = j;
                        bestp = pn;
                    }
                }
            }
            const uint32_t jstar = warp_argmax_nonneg(best, bestj);
            if (p.logp && bestj == jstar) {
                // Categorical.log_prob: log(clamp(probs, eps, 1 - eps))[action]
                const float eps = 1.1920928955078125e-07f;
                const float c = fminf(fmaxf(bestp, eps), 1.0f - eps);
                p.logp[((size_t)b * (n - 1) + step) * p.A + a] = logf(c);
            }
            if (lane == 0) tour_sm[step + 1] = (uint16_t)jstar;
#pragma unroll
            for (int k = 0; k < EPL; ++k)
                if (elem_index<EPL, VEC>(k, lane, lbw) == jstar && lane_on) vis |= 1u << k;
            cur = (int)jstar;
        }
    }
    __syncthreads();

    // ---- cooperative output: reference layout paths[b][s][a] (int64, step-major) and compact tours
    const int wvalid = min(W, p.A - a0);
    if (p.paths) {
        int64_t* out = p.paths + (size_t)b * n * p.A;
        for (int i = tid; i < n * W; i += nthreads) {
            const int s = i / W, w = i - s * W;
            if (w < wvalid) out[(size_t)s * p.A + a0 + w] = (int64_t)tour_all[(size_t)w * n + s];
        }
    }
    if (p.tours) {
        uint16_t* out = p.tours + ((size_t)b * p.A + a0) * n;
        for (int i = tid; i < n * wvalid; i += nthreads) out[i] = tour_all[i];
    }
}

template <int EPL, bool VEC, bool SMEMP>
static int launch_variant(const TspSampleParams& p, int W, size_t smem, cudaStream_t st) {
    auto kfn = tsp_sample_kernel<EPL, VEC, SMEMP>;
    DACO_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.A + W - 1) / W, p.B);
    kfn<<<grid, W * 32, smem, st>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

// element-wise product into a scratch matrix (only for colonies too large for shared memory)
__global__ void hadamard_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        o[i] = __fmul_rn(a[i], b[i]);
}

static float* g_prod_ws = nullptr;
static size_t g_prod_ws_bytes = 0;

}  // namespace deepaco

using namespace deepaco;

extern "C" uint64_t deepaco_tsp_sample_offset_increment(int n, int n_ants, int start_node) {
    const DeviceInfo* di = device_info();
    if (!di || n < 2 || n_ants < 1) return 0;
    uint64_t inc = (uint64_t)(n - 1) * torch_draw_plan((int64_t)n_ants * n, *di).increment;
    if (start_node < 0) inc += torch_draw_plan(n_ants, *di).increment;
    return inc;
}

extern "C" int deepaco_tsp_sample(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                                  int start_node, int double_norm, uint64_t seed, uint64_t offset,
                                  const uint64_t* rng, const float* noise, const int64_t* start, int64_t* paths,
                                  float* log_probs, uint16_t* tours, void* stream) {
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(pheromone != nullptr, "deepaco_tsp_sample: pheromone is NULL");
    DACO_CHECK_ARG(n >= 2 && n <= DEEPACO_MAX_NODES, "deepaco_tsp_sample: n=%d outside [2, %d]", n, DEEPACO_MAX_NODES);
    DACO_CHECK_ARG(n_ants >= 1 && n_colonies >= 1 && n_colonies <= 65535, "deepaco_tsp_sample: bad n_ants/n_colonies");
    DACO_CHECK_ARG(start_node < n, "deepaco_tsp_sample: start_node %d >= n", start_node);
    cudaStream_t st = (cudaStream_t)stream;

    TspSampleParams p{};
    p.ph = pheromone; p.heu = heuristic;
    p.n = n; p.A = n_ants; p.B = n_colonies;
    p.start_node = start_node; p.double_norm = double_norm;
    p.seed = seed; p.offset = offset; p.rng = rng;
    p.noise = noise; p.start = start;
    p.paths = paths; p.logp = log_probs; p.tours = tours;

    const SumPlan sp = aten_sum_plan(n, n_ants);
    const bool vec = sp.vectorized && (n % 4 == 0);
    int bw = sp.block_width > 32 ? 32 : sp.block_width;
    int lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    p.lbw = lbw;
    const DrawPlan dn = torch_draw_plan((int64_t)n_ants * n, *di);
    const DrawPlan ds = torch_draw_plan(n_ants, *di);
    p.g_noise = {dn.threads, dn.single};
    p.g_start = {ds.threads, ds.single};
    p.start_increment = (uint32_t)ds.increment;

    int epl_needed = vec ? 4 * ((n + 127) / 128) : (n + bw - 1) / bw;
    int epl = vec ? 8 : 1;
    while (epl < epl_needed) epl *= 2;
    DACO_CHECK_ARG(epl <= 32, "deepaco_tsp_sample: n=%d needs %d elements per lane (max 32)", n, epl);

    // warps per CTA: one warp per SM sub-partition while the job is small, 8-16 when it is not
    const long total_ants = (long)n_ants * n_colonies;
    int W = total_ants <= (long)di->sm_count * 4 ? 4 : 8;
    if (const char* e = getenv("DEEPACO_TSP_WARPS")) {
        const int w = atoi(e);
        if (w >= 1 && w <= 16) W = w;
    }
    const size_t pbytes = (((size_t)n * n * 4) + 15) & ~(size_t)15;
    size_t smem = pbytes + (size_t)W * n * 2;
    bool smemp = smem + 1024 <= (size_t)di->max_smem_optin;
    if (smemp && W < 16 && total_ants > (long)di->sm_count * 4) {
        // keep >= 32 resident warps per SM when shared memory allows only few CTAs
        const size_t per_sm = (size_t)di->max_smem_optin;
        while (W < 16 && (per_sm / (pbytes + (size_t)W * n * 2 + 1024)) * W < 32) W *= 2;
        smem = pbytes + (size_t)W * n * 2;
        smemp = smem + 1024 <= per_sm;
    }
    if (!smemp) {
        smem = (size_t)W * n * 2;
        if (heuristic) {   // product once per call into a scratch matrix, rows then come from L2
            const size_t need = (size_t)n_colonies * n * n * sizeof(float);
            if (need > g_prod_ws_bytes) {
                if (g_prod_ws) cudaFree(g_prod_ws);
                g_prod_ws = nullptr; g_prod_ws_bytes = 0;
                DACO_CHECK_CUDA(cudaMalloc(&g_prod_ws, need));
                g_prod_ws_bytes = need;
            }
            const size_t cnt = (size_t)n_colonies * n * n;
            hadamard_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(pheromone, heuristic, g_prod_ws, cnt);
            DACO_CHECK_LAUNCH();
            p.ph = g_prod_ws; p.heu = nullptr;
        }
    }

#define DACO_LAUNCH(E, V, S) return launch_variant<E, V, S>(p, W, smem, st)
    if (vec) {
        if (epl == 8) { if (smemp) DACO_LAUNCH(8, true, true); else DACO_LAUNCH(8, true, false); }
        if (epl == 16) DACO_LAUNCH(16, true, false);
        DACO_LAUNCH(32, true, false);
    } else {
        if (epl == 1) DACO_LAUNCH(1, false, true);
        if (epl == 2) DACO_LAUNCH(2, false, true);
        if (epl == 4) DACO_LAUNCH(4, false, true);
        if (epl == 8) { if (smemp) DACO_LAUNCH(8, false, true); else DACO_LAUNCH(8, false, false); }
        if (epl == 16) DACO_LAUNCH(16, false, false);
        DACO_LAUNCH(32, false, false);
    }
#undef DACO_LAUNCH
}
