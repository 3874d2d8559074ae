// csrc/two_opt.cuh compiled for the host (see cuda_emu.h) behind one entry point shaped like deepaco_two_opt /
// deepaco_tsp_nls.  Test infrastructure only: it lets the CPU suite run the 2-opt kernel SOURCE (register-carry
// variants and band kernel, TMA and cp.async row staging, NLS composition) against the C oracle.
#include "cuda_emu.h"

#include <algorithm>
using std::min;

#define DACO_FULL 0xffffffffu
#define DACO_NOINLINE __attribute__((noinline))
#define __shared__ static            /* one CTA at a time (cluster size 1, clusters run one after another) */
#define DACO_2OPT_SMEM(name) unsigned char* name = emu::ctx.smem

namespace deepaco {
// 32-bit shared-window addresses are offsets into the CTA's heap block
static inline uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const unsigned char*>(p) - emu::ctx.smem); }
static inline float lds_f32(uint32_t addr) { float v; memcpy(&v, emu::ctx.smem + addr, 4); return v; }
static inline uint32_t lds_u16(uint32_t addr) { uint16_t v; memcpy(&v, emu::ctx.smem + addr, 2); return v; }
// cp.async: the copy happens at issue time (groups complete immediately)
static inline void cp_async_16(float* dst, const float* src) { memcpy(dst, src, 16); }
static inline void cp_async_4(float* dst, const float* src) { memcpy(dst, src, 4); }
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}
// mbarrier + TMA bulk copy (cuda_emu.h)
static inline void mbar_init(uint64_t* bar, uint32_t) { emu_mbar_init(bar); }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_expect_tx(bar, bytes); }
static inline void fence_barrier_init() {}
static inline void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { emu_bulk_copy(dst, src, bytes, bar); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu_mbar_wait(bar, parity); }
}  // namespace deepaco

#include "../../deepaco_b200/csrc/two_opt.cuh"

namespace {
struct Args {
    const float* dist;
    const float* heu_dist;
    uint16_t* tours;
    int n, A, mode, maxt, T_nls, T_p;
    float* costs_out;
    int32_t* passes_out;
};
template <int K>
void run(const Args& a, int W) {
    emu::launch([](const Args& q) { deepaco::two_opt_kernel<K>(q.dist, q.heu_dist, q.tours, q.n, q.A, q.mode, q.maxt, q.T_nls, q.T_p,
                                                               q.costs_out, q.passes_out); },
                a, a.A, 1, W * 32, deepaco::two_opt_smem_bytes(W, a.n));
}
}  // namespace

// variant: -1 = what the library picks for this n, 0 = band kernel, 4 / 8 / 16 = register-carry kernel with that KMAX
extern "C" const char* emu_two_opt(const float* dist, const float* heu_dist, uint16_t* tours, int n, int n_ants, int mode,
                                   int max_iterations, int T_nls, int T_p, int variant, float* costs_out, int32_t* passes_out) {
    if (!dist || !tours || n < 4 || n_ants < 1) return "bad arguments";
    if (mode == 1 && !heu_dist) return "heuristic_dist is NULL";
    if (variant < 0) variant = deepaco::two_opt_variant(n);
    if (variant && n + 1 > 32 * variant) return "n does not fit this KMAX";
    const int W = deepaco::two_opt_warps(variant, n);
    const Args a{dist, heu_dist, tours, n, n_ants, mode, max_iterations, T_nls, T_p, costs_out, passes_out};
    switch (variant) {
        case 4: run<4>(a, W); break;
        case 8: run<8>(a, W); break;
        case 16: run<16>(a, W); break;
        case 0: run<0>(a, W); break;
        default: return "variant must be -1, 0, 4, 8 or 16";
    }
    return nullptr;
}
