// K1 -- TSP tour construction (one warp per ant, all n-1 steps inside one launch).
//
// Replaces the Python step loop of ACO.gen_path / pick_move (reference tsp/aco.py:134-177 and the
// fixed-start variant tsp_nls/aco.py:184-220): per step the reference gathers the pheromone and
// heuristic rows of the current node, multiplies them with the visited mask, normalises
// (`Categorical`), and draws with `torch.multinomial(probs, 1)` == argmax(probs / q), q ~ Exp(1)
// from `exponential_`.
//
// Two kernels:
//   tsp_sample_list_kernel  (n*n*4 bytes fit in shared memory, the benchmark sizes n <= ~230)
//     The product matrix P = pheromone (.) heuristic is staged once per CTA into shared memory with
//     TMA bulk copies.  Each warp keeps the list of its ant's unvisited nodes; per step only those
//     are evaluated: Philox word -> Exp(1) noise in registers, score x/q.  The arg-max is decided on
//     approximate scores (one MUFU.RCP + FMUL per candidate); the winner is certain -- identical to the
//     reference's exactly rounded argmax((x/S)/q) -- unless a second candidate lies within 2^-18
//     relative, in which case (probability ~1e-6 per step) the step is redone by exact_step() with
//     ATen's exact arithmetic.  The row normaliser S is only computed when log-probs are requested.
//   tsp_sample_dense_kernel (larger n: rows come from L2; also any draw geometry)
//     Dense register layout, every step evaluated with the exact arithmetic.
#include "common.cuh"
#include "host_util.h"
#include "sample_common.cuh"
#include "list_kernel.cuh"

#include <stdlib.h>

#include <algorithm>

namespace deepaco {

struct TspSampleParams {
    const float* ph;      // [B][n][n]
    const float* heu;     // [B][n][n] or null
    int n, A, B;
    int start_node;       // >= 0 fixed; -1 -> `start` tensor or torch randint stream
    int double_norm;
    uint64_t seed, offset;
    const uint64_t* offsets;  // [B] per-colony Philox offsets or null
    const float* noise;   // [B][n-1][A][n] or null
    const int64_t* start; // [B][A] or null
    int64_t* paths;       // [B][n][A] or null
    float* logp;          // [B][n-1][A] or null
    uint16_t* tours;      // [B][A][n] or null
    int lbw;              // log2(ATen block_width) for the strided summation order
    int vec;              // ATen vectorised summation order (n >= 128)
    DrawGeom g_noise, g_start;
    uint32_t start_increment, step_increment;
    int ant_base;
    uint16_t* peer_tours[8];   // fused exchange (ant sharding): [B][A_total][n] tour buffer of each of the n_peers GPUs
    int n_peers, A_total;
};

// ---------------------------------------------------------------------------------------------
// dense kernel (exact arithmetic every step)
// ---------------------------------------------------------------------------------------------
template <int EPL, bool VEC>
__device__ __forceinline__ uint32_t elem_index(int k, int lane, int lbw) {
    if (VEC) return 4u * (uint32_t)(lane + 32 * (k >> 2)) + (uint32_t)(k & 3);
    return (uint32_t)lane + ((uint32_t)k << lbw);
}

template <int EPL, bool VEC, bool SMEMP>
__global__ void __launch_bounds__(512) tsp_sample_dense_kernel(const TspSampleParams p) {
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

    if (SMEMP) stage_product(Psm, p.ph, p.heu, n, b, &bar);

    if (a < p.A) {
        const int lbw = p.lbw;
        const bool lane_on = VEC || lane < (1 << lbw);
        const uint64_t seed = p.seed;
        const uint64_t offset0 = (p.offsets ? p.offsets[b] : 0ull) + p.offset;
        const float* Pg = SMEMP ? Psm : (p.ph + (size_t)b * nn);   // !SMEMP: caller passes the product in `ph`

        int cur;
        uint64_t off_noise = offset0;
        if (p.start_node >= 0) {
            cur = p.start_node;
        } else if (p.start) {
            cur = (int)p.start[(size_t)b * p.A + a];
        } else {
            cur = (int)(torch_philox_word(seed, offset0, (uint64_t)(a + p.ant_base), p.g_start) % (uint32_t)n);
            off_noise += p.start_increment;
        }

        uint32_t vis = 0;
#pragma unroll
        for (int k = 0; k < EPL; ++k)
            if (elem_index<EPL, VEC>(k, lane, lbw) == (uint32_t)cur && lane_on) vis |= 1u << k;
        if (lane == 0) tour_sm[0] = (uint16_t)cur;

#pragma unroll 1
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
            const uint64_t off_step = off_noise + (uint64_t)p.step_increment * (uint64_t)step;
            const float* nz = p.noise ? p.noise + (((size_t)b * (n - 1) + step) * p.A + a) * (size_t)n : nullptr;
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const uint32_t j = elem_index<EPL, VEC>(k, lane, lbw);
                if (lane_on && j < (uint32_t)n) {
                    const float pn = __fdiv_rn(x[k], S);
                    const float q = nz ? nz[j]
                                       : exp1_from_word(torch_philox_word(seed, off_step, (uint64_t)(a + p.ant_base) * n + j, p.g_noise));
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
                const float eps = 1.1920928955078125e-07f;
                p.logp[((size_t)b * (n - 1) + step) * p.A + a] = logf(fminf(fmaxf(bestp, eps), 1.0f - eps));
            }
            if (lane == 0) tour_sm[step + 1] = (uint16_t)jstar;
#pragma unroll
            for (int k = 0; k < EPL; ++k)
                if (elem_index<EPL, VEC>(k, lane, lbw) == jstar && lane_on) vis |= 1u << k;
            cur = (int)jstar;
        }
    }
    __syncthreads();

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
    for (int r = 0; r < p.n_peers; ++r) {
        uint16_t* out = p.peer_tours[r] + ((size_t)b * p.A_total + p.ant_base + a0) * n;
        for (int i = tid; i < n * wvalid; i += nthreads) out[i] = tour_all[i];
    }
}

// Warps per CTA for the kNN kernel.  A tour is a chain of n-1 dependent steps; measured (profiles/r02_k1_*): a step takes
// about 700 + 30 x (resident warps per SM) cycles -- latency of the dependent chain plus issue contention, ~1180 cycles
// at 16 warps, ~1660 at 32.  Estimated time = waves x that; ties go to the larger CTA (per-CTA staging shared by more ants).
// Returns 0 when no size fits shared memory.
static int knn_pick_warps(int n, int n_ants, int n_colonies, int sm_count, size_t cap) {
    int best_w = 0;
    double best_t = 0.0;
    for (int w = 4; w <= 32; w *= 2) {
        const size_t sm = knn_kernel_smem(n, w);
        if (sm > cap) continue;
        long cps = (long)(cap / sm);                       // CTAs per SM: shared memory, then registers (64 / thread -> 32 warps)
        cps = std::max<long>(std::min<long>(cps, 32 / w), 1);
        const long ctas = (long)((n_ants + w - 1) / w) * n_colonies;
        const long slots = (long)sm_count * cps;
        const long waves = (ctas + slots - 1) / slots;
        const long resident = std::min<long>(cps, (ctas + sm_count - 1) / sm_count) * w;
        const double t = (double)waves * (700.0 + 30.0 * (double)resident);
        if (best_w == 0 || t <= best_t) { best_w = w; best_t = t; }
    }
    return best_w;
}

template <typename KFn, typename P>
static int launch_kernel(KFn kfn, const P& p, int W, size_t smem, cudaStream_t st) {
    DACO_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DACO_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
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

}  // namespace deepaco

using namespace deepaco;

extern "C" uint64_t deepaco_tsp_sample_offset_increment(int n, int n_ants, int start_node) {
    const DeviceInfo* di = device_info();
    if (!di || n < 2 || n_ants < 1) return 0;
    uint64_t inc = (uint64_t)(n - 1) * torch_draw_plan((int64_t)n_ants * n, *di).increment;
    if (start_node < 0) inc += torch_draw_plan(n_ants, *di).increment;
    return inc;
}

static int tsp_sample_impl(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                           int start_node, int double_norm, uint64_t seed, uint64_t offset,
                           const uint64_t* offsets, const float* noise, const int64_t* start, int64_t* paths,
                           float* log_probs, uint16_t* tours, const uint8_t* knn, int ant_base, int n_ants_total, void* stream,
                           const float* fuse_dist = nullptr, float* fuse_costs = nullptr, uint32_t* fuse_nbr = nullptr,
                           int* fused_out = nullptr, const uint64_t* peer_tours = nullptr, int n_peers = 0) {
    const DeviceInfo* di = device_info();
    if (fused_out) *fused_out = 0;
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(pheromone != nullptr, "deepaco_tsp_sample: pheromone is NULL");
    DACO_CHECK_ARG(n >= 2 && n <= DEEPACO_MAX_NODES, "deepaco_tsp_sample: n=%d outside [2, %d]", n, DEEPACO_MAX_NODES);
    DACO_CHECK_ARG(n_ants >= 1 && n_colonies >= 1 && n_colonies <= 65535, "deepaco_tsp_sample: bad n_ants/n_colonies");
    DACO_CHECK_ARG(start_node < n, "deepaco_tsp_sample: start_node %d >= n", start_node);
    DACO_CHECK_ARG(ant_base >= 0 && n_ants_total >= ant_base + n_ants, "deepaco_tsp_sample: ant shard [%d, %d) outside the colony's %d ants",
                   ant_base, ant_base + n_ants, n_ants_total);
    cudaStream_t st = (cudaStream_t)stream;
    StreamScratch prod_ws;   // product matrix for colonies too large for shared memory; freed (stream-ordered) on return

    TspSampleParams p{};
    p.ph = pheromone; p.heu = heuristic;
    p.n = n; p.A = n_ants; p.B = n_colonies;
    p.start_node = start_node; p.double_norm = double_norm;
    p.seed = seed; p.offset = offset; p.offsets = offsets;
    p.noise = noise; p.start = start;
    p.paths = paths; p.logp = log_probs; p.tours = tours;

    const SumPlan sp = aten_sum_plan(n, n_ants_total);   // the reference sums [n_ants_total, n] tensors
    const bool vec = sp.vectorized && (n % 4 == 0);
    int bw = sp.block_width > 32 ? 32 : sp.block_width;
    int lbw = 0;
    while ((1 << lbw) < bw) ++lbw;
    p.lbw = lbw;
    p.vec = sp.vectorized ? 1 : 0;   // runtime ATen sums handle n % 4 != 0 through the row shift
    const DrawPlan dn = torch_draw_plan((int64_t)n_ants_total * n, *di);
    const DrawPlan ds = torch_draw_plan(n_ants_total, *di);
    p.ant_base = ant_base;
    p.g_noise = {dn.threads, dn.single};
    p.g_start = {ds.threads, ds.single};
    p.start_increment = (uint32_t)ds.increment;
    p.step_increment = (uint32_t)dn.increment;

    // warps per CTA: one warp per SM sub-partition while the job is small, 8-16 when it is not
    const long total_ants = (long)n_ants * n_colonies;
    int W = total_ants <= (long)di->sm_count * 4 ? 4 : 8;
    if (const char* e = getenv("DEEPACO_TSP_WARPS")) {
        const int w = atoi(e);
        if (w >= 1 && w <= 16) W = w;
    }
    const size_t pbytes = (((size_t)n * n * 4) + 15) & ~(size_t)15;
    auto list_smem = [&](int w) { return list_kernel_smem(n, n, w, false); };
    const size_t cap = (size_t)di->max_smem_optin - 1024;
    const bool force_dense = getenv("DEEPACO_TSP_DENSE") != nullptr;
    const bool knn_ok = knn && !noise && !log_probs && !paths && !start && (tours || n_peers) && n > 32 && n <= 256 &&
                        (uint64_t)n_ants_total * n < (1ull << 32) && ds.single && !getenv("DEEPACO_TSP_NO_KNN");
    if (!force_dense && (dn.single || knn_ok) && (uint64_t)n_ants_total * n < (1ull << 32) && n <= 256 && (list_smem(W) <= cap || knn_ok)) {
        // keep >= 32 resident warps per SM when shared memory allows only few CTAs
        if (total_ants > (long)di->sm_count * 4)
            while (W < 16 && (cap / list_smem(W)) * W < 32 && list_smem(W * 2) <= cap) W *= 2;
        ListParams q{};
        q.ph = pheromone; q.heu = heuristic; q.n = n; q.A = n_ants; q.B = n_colonies; q.rows = n;
        q.start_node = start_node; q.double_norm = double_norm; q.seed = seed; q.offset = offset; q.offsets = offsets;
        q.keys.init(seed);
        q.ant_base = ant_base;
        q.A_total = n_ants_total;
        q.n_peers = n_peers;
        for (int r = 0; r < n_peers && r < 8; ++r) q.peer_tours[r] = reinterpret_cast<uint16_t*>(peer_tours[r]);
        q.noise = noise; q.start = start; q.paths = paths; q.logp = log_probs; q.tours = tours;
        q.lbw = p.lbw; q.vec = p.vec; q.g_noise = p.g_noise; q.g_start = p.g_start;
        q.start_increment = p.start_increment; q.step_increment = p.step_increment;
        if (knn_ok) {
            // sparse product: one candidate per lane (kNN kernel)
            // 4 warps while the job is tiny, 16 when many CTAs queue per SM (the per-CTA staging is shared by more ants)
            int Wk = knn_pick_warps(n, n_ants, n_colonies, di->sm_count, cap);
            if (const char* e = getenv("DEEPACO_TSP_WARPS")) { const int w = atoi(e); if (w >= 1 && w <= 32) Wk = w; }
            if (Wk > 0 && knn_kernel_smem(n, Wk) <= cap) {
                q.knn = knn;
                // small jobs are launch-latency bound: fold cost + neighbour table into the construction kernel's epilogue
                const bool gen = !dn.single;   // more elements than torch's draw launch has threads: general Philox geometry
                // (measured for big jobs too, costs only: the fused instantiation runs 4.5 % slower -- 87 us on the 256-colony
                // step -- than the plain one, more than the separate cost kernel's 67 us)
                const bool fuse = !gen && fuse_dist && fuse_costs && ant_base == 0 && n_ants_total == n_ants &&
                                  (getenv("DEEPACO_TSP_FUSE_COST") || total_ants <= (long)di->sm_count * 16);
                if (fuse) {
                    q.dist = fuse_dist; q.costs = fuse_costs; q.nbr = fuse_nbr;
                    if (fused_out) *fused_out = 1;
                }
                dim3 grid((n_ants + Wk - 1) / Wk, n_colonies);
                const size_t ksm = knn_kernel_smem(n, Wk);
#define DACO_KNN(F, M, G)                                                                                                       \
    do {                                                                                                                        \
        DACO_CHECK_CUDA(cudaFuncSetAttribute(aco_knn_kernel<F, M, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm));  \
        DACO_CHECK_CUDA(cudaFuncSetAttribute(aco_knn_kernel<F, M, G>, cudaFuncAttributePreferredSharedMemoryCarveout,           \
                                             cudaSharedmemCarveoutMaxShared));                                                  \
        aco_knn_kernel<F, M, G><<<grid, Wk * 32, ksm, st>>>(q);                                                                 \
    } while (0)
                if (gen) { if (Wk <= 8) DACO_KNN(false, 8, true); else if (Wk <= 16) DACO_KNN(false, 16, true); else DACO_KNN(false, 32, true); }
                else if (fuse) { if (Wk <= 8) DACO_KNN(true, 8, false); else if (Wk <= 16) DACO_KNN(true, 16, false); else DACO_KNN(true, 32, false); }
                else { if (Wk <= 8) DACO_KNN(false, 8, false); else if (Wk <= 16) DACO_KNN(false, 16, false); else DACO_KNN(false, 32, false); }
#undef DACO_KNN
                DACO_CHECK_LAUNCH();
                return DEEPACO_OK;
            }
        }
        if (dn.single && list_smem(W) <= cap) {
        const int epl = (n - 1 + 31) / 32;
        const size_t sm = list_smem(W);
#define DACO_LIST(E)                                                                                        \
        do {                                                                                                \
            if (noise) {                                                                                    \
                if (log_probs) return launch_kernel(aco_list_kernel<E, false, true, true>, q, W, sm, st);   \
                return launch_kernel(aco_list_kernel<E, false, false, true>, q, W, sm, st);                 \
            }                                                                                               \
            if (log_probs) return launch_kernel(aco_list_kernel<E, false, true, false>, q, W, sm, st);      \
            return launch_kernel(aco_list_kernel<E, false, false, false>, q, W, sm, st);                    \
        } while (0)
        if (epl <= 1) DACO_LIST(1);
        if (epl <= 2) DACO_LIST(2);
        if (epl <= 4) DACO_LIST(4);
        DACO_LIST(8);
#undef DACO_LIST
        }
    }

    // ---- list kernel with the product in global memory (n too large for shared memory)
    if (!force_dense && dn.single && (uint64_t)n_ants_total * n < (1ull << 32) && n - 1 <= 32 * 32) {
        const float* prod = pheromone;
        if (heuristic) {   // product once per call into a scratch matrix, rows then come from L2
            const size_t cnt = (size_t)n_colonies * n * n;
            DACO_CHECK_CUDA(prod_ws.alloc(cnt * sizeof(float), st));   // stream-ordered: private to this call
            float* ws = static_cast<float*>(prod_ws.ptr);
            hadamard_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(pheromone, heuristic, ws, cnt);
            DACO_CHECK_LAUNCH();
            prod = ws;
        }
        ListParams q{};
        q.ph = prod; q.heu = nullptr; q.n = n; q.A = n_ants; q.B = n_colonies; q.rows = n;
        q.start_node = start_node; q.double_norm = double_norm; q.seed = seed; q.offset = offset; q.offsets = offsets;
        q.keys.init(seed);
        q.ant_base = ant_base;
        q.A_total = n_ants_total;
        q.n_peers = n_peers;
        for (int r = 0; r < n_peers && r < 8; ++r) q.peer_tours[r] = reinterpret_cast<uint16_t*>(peer_tours[r]);
        q.noise = noise; q.start = start; q.paths = paths; q.logp = log_probs; q.tours = tours;
        q.lbw = p.lbw; q.vec = p.vec; q.g_noise = p.g_noise; q.g_start = p.g_start;
        q.start_increment = p.start_increment; q.step_increment = p.step_increment;
        const int Wg = 4;
        const size_t sm = list_kernel_smem(n, n, Wg, false, true);
        const int epl = (n - 1 + 31) / 32;
#define DACO_GLIST(E)                                                                                               \
        do {                                                                                                        \
            if (noise) {                                                                                            \
                if (log_probs) return launch_kernel(aco_list_kernel<E, false, true, true, true>, q, Wg, sm, st);    \
                return launch_kernel(aco_list_kernel<E, false, false, true, true>, q, Wg, sm, st);                  \
            }                                                                                                       \
            if (log_probs) return launch_kernel(aco_list_kernel<E, false, true, false, true>, q, Wg, sm, st);       \
            return launch_kernel(aco_list_kernel<E, false, false, false, true>, q, Wg, sm, st);                     \
        } while (0)
        if (epl <= 8) DACO_GLIST(8);
        if (epl <= 16) DACO_GLIST(16);
        DACO_GLIST(32);
#undef DACO_GLIST
    }

    // ---- dense exact kernel
    p.n_peers = n_peers; p.A_total = n_ants_total;
    for (int r = 0; r < n_peers && r < 8; ++r) p.peer_tours[r] = reinterpret_cast<uint16_t*>(peer_tours[r]);
    int epl_needed = vec ? 4 * ((n + 127) / 128) : (n + bw - 1) / bw;
    int epl = vec ? 8 : 1;
    while (epl < epl_needed) epl *= 2;
    DACO_CHECK_ARG(epl <= 32, "deepaco_tsp_sample: n=%d needs %d elements per lane (max 32)", n, epl);
    size_t smem = pbytes + (size_t)W * n * 2;
    bool smemp = smem <= cap;
    if (!smemp) {
        smem = (size_t)W * n * 2;
        if (heuristic) {   // product once per call into a scratch matrix, rows then come from L2
            const size_t cnt = (size_t)n_colonies * n * n;
            DACO_CHECK_CUDA(prod_ws.alloc(cnt * sizeof(float), st));   // stream-ordered: private to this call
            float* ws = static_cast<float*>(prod_ws.ptr);
            hadamard_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(pheromone, heuristic, ws, cnt);
            DACO_CHECK_LAUNCH();
            p.ph = ws; p.heu = nullptr;
        }
    }

#define DACO_LAUNCH(E, V, S) return launch_kernel(tsp_sample_dense_kernel<E, V, S>, p, W, smem, st)
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

extern "C" int deepaco_tsp_sample(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                                  int start_node, int double_norm, uint64_t seed, uint64_t offset,
                                  const uint64_t* offsets, const float* noise, const int64_t* start, int64_t* paths,
                                  float* log_probs, uint16_t* tours, const uint8_t* knn, void* stream) {
    return tsp_sample_impl(pheromone, heuristic, n, n_ants, n_colonies, start_node, double_norm, seed, offset, offsets, noise,
                           start, paths, log_probs, tours, knn, 0, n_ants, stream);
}

// Ant-sharded variant: this launch constructs ants [ant_base, ant_base + n_ants) of colonies that have n_ants_total
// ants.  Noise words and summation plans are those of the full colony, so an ant's tour does not depend on how the
// colony is split across GPUs.  Output buffers are indexed by the LOCAL ant number.
extern "C" int deepaco_tsp_sample_shard(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                                        int start_node, int double_norm, uint64_t seed, uint64_t offset,
                                        const uint64_t* offsets, int64_t* paths, float* log_probs, uint16_t* tours,
                                        const uint8_t* knn, int ant_base, int n_ants_total, void* stream) {
    return tsp_sample_impl(pheromone, heuristic, n, n_ants, n_colonies, start_node, double_norm, seed, offset, offsets, nullptr,
                           nullptr, paths, log_probs, tours, knn, ant_base, n_ants_total, stream);
}

namespace deepaco {
// internal: sampling with the cost / neighbour-table epilogue fused when the kNN kernel is selected (*fused = 1)
int tsp_sample_fused(const float* product, int n, int n_ants, int n_colonies, int start_node, int double_norm, uint64_t seed,
                     uint64_t offset, const uint64_t* offsets, uint16_t* tours, const uint8_t* knn, const float* dist, float* costs,
                     uint32_t* nbr, int* fused, cudaStream_t st) {
    return tsp_sample_impl(product, nullptr, n, n_ants, n_colonies, start_node, double_norm, seed, offset, offsets, nullptr, nullptr,
                           nullptr, nullptr, tours, knn, 0, n_ants, st, dist, costs, nbr, fused);
}
// internal: ant-sharded sampling from a ready product matrix with the fused peer store (deepaco_tsp_run_shard)
int tsp_sample_peers(const float* product, int n, int n_ants_local, int n_colonies, int start_node, int double_norm, uint64_t seed,
                     uint64_t offset, const uint64_t* offsets, const uint8_t* knn, int ant_base, int n_ants_total,
                     const uint64_t* peer_tours_host, int n_peers, cudaStream_t st) {
    return tsp_sample_impl(product, nullptr, n, n_ants_local, n_colonies, start_node, double_norm, seed, offset, offsets, nullptr,
                           nullptr, nullptr, nullptr, nullptr, knn, ant_base, n_ants_total, st, nullptr, nullptr, nullptr, nullptr,
                           peer_tours_host, n_peers);
}
}  // namespace deepaco

// Ant-sharded construction with the exchange fused into the kernel: each finished tour is written by the building
// warp into the [B][n_ants_total][n] uint16 tour buffer of EVERY rank (`peer_tours_host[r]` = peer-mapped device
// pointer of rank r's buffer, r < n_peers <= 8, our own included), so the transfer over NVLink overlaps the
// construction of the remaining ants.  The caller then needs only a barrier, no collective.
extern "C" int deepaco_tsp_sample_shard_p2p(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                                            int start_node, int double_norm, uint64_t seed, uint64_t offset,
                                            const uint64_t* offsets, const uint8_t* knn, int ant_base, int n_ants_total,
                                            const uint64_t* peer_tours_host, int n_peers, void* stream) {
    DACO_CHECK_ARG(peer_tours_host && n_peers >= 1 && n_peers <= 8, "deepaco_tsp_sample_shard_p2p: need 1..8 peer buffers");
    return tsp_sample_impl(pheromone, heuristic, n, n_ants, n_colonies, start_node, double_norm, seed, offset, offsets, nullptr,
                           nullptr, nullptr, nullptr, nullptr, knn, ant_base, n_ants_total, stream, nullptr, nullptr, nullptr,
                           nullptr, peer_tours_host, n_peers);
}
