// csrc/gnn.cuh (eval-mode heuristic network, one CTA per instance, TMA double-buffered layer weights) compiled for the
// host (see cuda_emu.h) behind an entry point shaped like deepaco_gnn_forward.  Test infrastructure only.
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
#define DACO_DYN_SMEM16(name) unsigned char* name = emu::ctx.smem
#define __shared__ static

namespace deepaco {
static inline uint32_t smem_u32(const void* p) { return (uint32_t)(static_cast<const unsigned char*>(p) - emu::ctx.smem); }
static inline void mbar_init(uint64_t* bar, uint32_t) { emu_mbar_init(bar); }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_mbar_expect_tx(bar, bytes); }
static inline void fence_barrier_init() {}
static inline void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { emu_bulk_copy(dst, src, bytes, bar); }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu_mbar_wait(bar, parity); }
}  // namespace deepaco

#include "../../deepaco_b200/csrc/gnn.cuh"

using namespace deepaco;

extern "C" const char* emu_gnn_forward(const float* x, const int32_t* row_ptr, const int32_t* dst_sorted, const float* attr_sorted,
                                       const int32_t* order, const float* weights, int n, int E, int feats, int B, float* node_ws,
                                       float* edge_ws, float* heu_out, float* dense_out, float dense_eps, int threads) {
    if (!(x && row_ptr && dst_sorted && attr_sorted && order && weights && node_ws && edge_ws && (heu_out || dense_out))) return "NULL argument";
    if (!(n >= 1 && E >= 1 && feats >= 1 && feats <= 8 && B >= 1 && threads >= 32 && threads % 32 == 0)) return "bad sizes";
    const GnnParams p{x, row_ptr, dst_sorted, attr_sorted, order, weights, node_ws, edge_ws, heu_out, dense_out, dense_eps, n, E, feats,
                      nullptr};
    emu::launch(gnn_forward_kernel, p, B, 1, threads, gnn_forward_smem(feats, threads));
    return nullptr;
}
