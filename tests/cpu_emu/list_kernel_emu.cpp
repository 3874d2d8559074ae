// csrc/list_kernel.cuh (+ common.cuh, sample_common.cuh) compiled for the host (see cuda_emu.h): the tour-construction
// kernels of TSP / TSP-NLS / CVRP -- aco_list_kernel in all its template variants and aco_knn_kernel -- behind entry
// points shaped like deepaco_tsp_sample / deepaco_cvrp_sample.  Test infrastructure only.
//
// What the host build can and cannot reproduce: with EXTERNAL noise every value the kernels compute is IEEE fp32
// (the ranking filter uses 1/q where the device uses rcp.approx -- a tighter approximation, so the filter's proof of
// "same winner as the exactly rounded ranking" still holds), hence tours and log-probs must equal the reference's
// goldens.  In Philox mode the Exp(1) transform uses libm log2 instead of MUFU.LG2, so noise differs in the last bits
// from a GPU's; kernels are then compared with each other (kNN == list kernel), not with a GPU run.
#include "cuda_emu.h"

#include <algorithm>
#include <cmath>
using std::min;

struct uint4 {
    uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

#define DACO_NOINLINE __attribute__((noinline))
#define DACO_DYN_SMEM128(name) unsigned char* name = emu::ctx.smem
#define DACO_STS_U8(addr, v) (emu::ctx.smem[(addr)] = (unsigned char)(v))
#define __shared__ static            /* one CTA at a time */

namespace deepaco {
static inline float rcp_approx(float x) { return 1.0f / x; }
static inline float exp1_from_word_guarded(uint32_t x) {
    const float u = fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    const float lg = (u >= 1.0f - 1.1920928955078125e-07f / 2.0f) ? -(1.1920928955078125e-07f / 2.0f) : log2f(u) * 0.693147182464599609375f;
    return -lg;
}
static inline float exp1_from_word(uint32_t x) {
    const float u = fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    return fmaxf(-(log2f(u) * 0.693147182464599609375f), 1.1920928955078125e-07f / 2.0f);
}
static inline uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const unsigned char*>(p) - emu::ctx.smem); }
static inline void mbar_init(uint64_t* bar, uint32_t) { emu_mbar_init(bar); }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_expect_tx(bar, bytes); }
static inline void fence_barrier_init() {}
static inline void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { emu_bulk_copy(dst, src, bytes, bar); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu_mbar_wait(bar, parity); }
static inline uint32_t lds_s8(uint32_t a) { return (uint32_t)(int32_t)(int8_t)emu::ctx.smem[a]; }
static inline uint32_t lds_u8(uint32_t a) { return emu::ctx.smem[a]; }
static inline uint32_t lds_u16(uint32_t a) { uint16_t v; memcpy(&v, emu::ctx.smem + a, 2); return v; }
static inline uint32_t lds_u32(uint32_t a) { uint32_t v; memcpy(&v, emu::ctx.smem + a, 4); return v; }
static inline float lds_f32(uint32_t a) { float v; memcpy(&v, emu::ctx.smem + a, 4); return v; }
static inline void sts_u16(uint32_t a, uint32_t v) { const uint16_t h = (uint16_t)v; memcpy(emu::ctx.smem + a, &h, 2); }
static inline uint32_t pin_u32(uint32_t v) { return v; }
}  // namespace deepaco

#include "../../deepaco_b200/csrc/list_kernel.cuh"
#include "../../deepaco_b200/csrc/pick_move.cuh"

using namespace deepaco;

namespace {
template <int E, bool CVRP, bool GLOBAL = false>
void run_list(const ListParams& p, int W, bool logp, bool ext) {
    const size_t sm = list_kernel_smem(p.n, p.rows, W, CVRP, GLOBAL);
    const int gx = (p.A + W - 1) / W;
    auto go = [&](auto kernel) { emu::launch(kernel, p, gx * p.B, 1, W * 32, sm, gx); };
    if (ext) { if (logp) go(aco_list_kernel<E, CVRP, true, true, GLOBAL>); else go(aco_list_kernel<E, CVRP, false, true, GLOBAL>); }
    else { if (logp) go(aco_list_kernel<E, CVRP, true, false, GLOBAL>); else go(aco_list_kernel<E, CVRP, false, false, GLOBAL>); }
}
template <bool CVRP>
const char* dispatch_list(const ListParams& p, int W) {
    const int epl = CVRP ? (p.n + 31) / 32 : (p.n - 1 + 31) / 32;
    const bool logp = p.logp != nullptr, ext = p.noise != nullptr;
    if (p.heu == nullptr && p.n > 256) {       // product matrix in global memory (colonies too large for shared memory)
        if (epl <= 16) run_list<16, CVRP, true>(p, W, logp, ext);
        else return "n too large for the host harness (n <= 512)";
        return nullptr;
    }
    if (epl <= 1) run_list<1, CVRP>(p, W, logp, ext);
    else if (epl <= 2) run_list<2, CVRP>(p, W, logp, ext);
    else if (epl <= 4) run_list<4, CVRP>(p, W, logp, ext);
    else if (epl <= 8) run_list<8, CVRP>(p, W, logp, ext);
    else return "n too large for the host harness (n <= 256)";
    return nullptr;
}
void fill_common(ListParams& p, const float* ph, const float* heu, int n, int A, int B, uint64_t seed, uint64_t offset,
                 const float* noise, int lbw, int vec, uint32_t g_threads, uint32_t g_single, uint32_t increment) {
    p.ph = ph; p.heu = heu; p.n = n; p.A = A; p.B = B; p.seed = seed; p.offset = offset; p.offsets = nullptr;
    p.keys.init(seed);
    p.noise = noise; p.lbw = lbw; p.vec = vec;
    p.g_noise = {g_threads, g_single};
    p.g_start = p.g_noise;
    p.step_increment = increment; p.start_increment = increment;
    p.ant_base = 0; p.A_total = A; p.n_peers = 0;
}
}  // namespace

// kernel: 0 = aco_list_kernel, 1 = aco_knn_kernel (needs `knn`, Philox noise, tours out).  Draw geometry
// (g_threads, g_single, increment) is whatever the caller wants the Philox stream to look like.
extern "C" const char* emu_tsp_sample(const float* ph, const float* heu, int n, int A, int B, int start_node, int double_norm,
                                      uint64_t seed, uint64_t offset, const float* noise, const int64_t* start, const uint8_t* knn,
                                      int64_t* paths, float* logp, uint16_t* tours, int lbw, int vec, uint32_t g_threads,
                                      uint32_t g_single, uint32_t increment, int kernel, int W) {
    if (!ph || n < 2 || A < 1 || B < 1 || W < 1 || W > 16) return "bad arguments";
    ListParams p{};
    fill_common(p, ph, heu, n, A, B, seed, offset, noise, lbw, vec, g_threads, g_single, increment);
    p.rows = n; p.start_node = start_node; p.double_norm = double_norm; p.start = start;
    p.paths = paths; p.logp = logp; p.tours = tours;
    if (kernel == 1 || kernel == 2) {   // 2 = the general-geometry instantiation (any g_threads / increment)
        if (!knn || noise || logp || paths || start || !tours || n <= 32 || n > 256 || (kernel == 1 && (!g_single || increment != 4)))
            return "the kNN kernel needs candidate lists, Philox noise, tours out, 32 < n <= 256";
        p.knn = knn;
        const int gx = (A + W - 1) / W;
        const size_t sm = knn_kernel_smem(n, W);
        if (kernel == 2) {
            if (W <= 8) emu::launch(aco_knn_kernel<false, 8, true>, p, gx * B, 1, W * 32, sm, gx);
            else emu::launch(aco_knn_kernel<false, 16, true>, p, gx * B, 1, W * 32, sm, gx);
        } else {
            if (W <= 8) emu::launch(aco_knn_kernel<false, 8>, p, gx * B, 1, W * 32, sm, gx);
            else emu::launch(aco_knn_kernel<false, 16>, p, gx * B, 1, W * 32, sm, gx);
        }
        return nullptr;
    }
    return dispatch_list<false>(p, W);
}

// Exp(1) variates of a torch `exponential_` draw of `numel` elements at (seed, offset) for an arbitrary launch geometry,
// element by element through torch_philox_word (the literal layout rule): the independent reference for the
// general-geometry kNN kernel.
extern "C" void emu_exponential_general(uint64_t seed, uint64_t offset, int64_t numel, uint32_t g_threads, uint32_t g_single, float* out) {
    const DrawGeom g{g_threads, g_single};
    for (int64_t li = 0; li < numel; ++li) out[li] = exp1_from_word(torch_philox_word(seed, offset, (uint64_t)li, g));
}

extern "C" const char* emu_cvrp_sample(const float* ph, const float* heu, const float* demand, float capacity, int n, int A, int B,
                                       uint64_t seed, uint64_t offset, const float* noise, int64_t* paths, float* logp,
                                       uint16_t* tours, int32_t* lens, int32_t* tmax, int lbw, int vec, uint32_t g_threads,
                                       uint32_t g_single, uint32_t increment, int W) {
    if (!ph || !demand || !lens || !tmax || n < 2 || A < 1 || B < 1 || W < 1 || W > 16) return "bad arguments";
    ListParams p{};
    fill_common(p, ph, heu, n, A, B, seed, offset, noise, lbw, vec, g_threads, g_single, increment);
    p.rows = 2 * n; p.start_node = 0; p.double_norm = 0; p.demand = demand; p.capacity = capacity;
    p.paths = paths; p.logp = logp; p.tours = tours; p.lens = lens; p.tmax = tmax;
    for (int b = 0; b < B; ++b) tmax[b] = 0;
    return dispatch_list<true>(p, W);
}

// ACO.pick_move (deepaco_pick_move): one step for caller-held masks, Philox noise with the given draw geometry
extern "C" const char* emu_pick_move(const float* php, const float* heup, const int64_t* prev, const float* mask, const float* mask2,
                                     int n, int A, uint64_t seed, uint64_t offset, int64_t* actions, float* logp, int* bad_prev,
                                     int lbw, int vec, uint32_t g_threads, uint32_t g_single) {
    if (!php || !prev || !mask || !actions || !bad_prev || n < 1 || A < 1) return "bad arguments";
    PickMoveParams p{};
    p.php = php; p.heup = heup; p.prev = prev; p.mask = mask; p.mask2 = mask2; p.n = n; p.A = A; p.lbw = lbw; p.vec = vec;
    p.seed = seed; p.offset = offset; p.g = {g_threads, g_single}; p.actions = actions; p.logp = logp; p.bad_prev = bad_prev;
    const int W = 8;
    emu::launch(pick_move_kernel, p, (A + W - 1) / W, 1, W * 32, 16);
    return nullptr;
}
