// ACO.pick_move: C ABI (kernel in pick_move.cuh).
#include "pick_move.cuh"
#include "host_util.h"

using namespace deepaco;

extern "C" uint64_t deepaco_pick_move_offset_increment(int n, int n_ants) {
    const DeviceInfo* di = device_info();
    if (!di || n < 1 || n_ants < 1) return 0;
    return torch_draw_plan((int64_t)n_ants * n, *di).increment;
}

extern "C" int deepaco_pick_move(const float* pheromone_pow, const float* heuristic_pow, const int64_t* prev, const float* mask,
                                 const float* mask2, int n, int n_ants, uint64_t seed, uint64_t offset, int64_t* actions,
                                 float* log_probs, int* bad_prev, void* stream) {
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    DACO_CHECK_ARG(pheromone_pow && prev && mask && actions && bad_prev, "deepaco_pick_move: NULL argument");
    DACO_CHECK_ARG(n >= 1 && n_ants >= 1, "deepaco_pick_move: bad sizes n=%d n_ants=%d", n, n_ants);
    PickMoveParams p{};
    p.php = pheromone_pow; p.heup = heuristic_pow; p.prev = prev; p.mask = mask; p.mask2 = mask2;
    p.n = n; p.A = n_ants; p.seed = seed; p.offset = offset; p.actions = actions; p.logp = log_probs; p.bad_prev = bad_prev;
    const SumPlan sp = aten_sum_plan(n, n_ants);              // the reference sums a contiguous [n_ants, n] tensor
    int bw = sp.block_width > 32 ? 32 : sp.block_width;
    while ((1 << p.lbw) < bw) ++p.lbw;
    p.vec = sp.vectorized ? 1 : 0;
    const DrawPlan dn = torch_draw_plan((int64_t)n_ants * n, *di);
    p.g = {dn.threads, dn.single};
    const int W = 8;
    pick_move_kernel<<<(n_ants + W - 1) / W, W * 32, 0, (cudaStream_t)stream>>>(p);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}
