// ACO.run for TSP colonies whose ANTS are split over the GPUs of one box (BASELINE north_star:
// "ants ... shard across the 8 GPUs ... one exchange per ACO iteration").  The reference has no multi-GPU code; the
// single-GPU semantics being reproduced are tsp/aco.py:74-92 (run), :134-177 (gen_path / pick_move), :94-132.
//
// Every rank enqueues, per iteration and without a host sync:
//   K1 shard   ants [ant_base, ant_base + n_local) are built with the Philox words and ATen summation plans of the FULL
//              colony; each finished tour is stored by the building warp into the tour buffer of EVERY rank over
//              NVLink (peer-mapped st.global), so the exchange overlaps the construction of the other ants;
//   barrier    one tiny kernel: release-store of the iteration number into every peer's flag word, then acquire-spin
//              on our own flag words (no NCCL call, no host involvement);
//   K2 replay  cost + neighbour table, best tracking and the ordered deposit over ALL tours, on every rank -> the
//              pheromone (and with it every later tour) is bit-identical to the single-GPU run for any world size.
// Tour buffers are double-buffered by iteration parity: a rank can only be one barrier ahead of its slowest peer, and
// the peer finished reading buffer k of iteration t before it signalled iteration t+1, so writing iteration t+2 into
// buffer k needs no second barrier.
#include "common.cuh"
#include "host_util.h"

namespace deepaco {

int tsp_update_launch(float* pheromone, const uint32_t* neighbours, const float* costs, int n, int n_ants, int n_colonies,
                      float decay, int elitist, int min_max, float ph_min, const float* ph_max, const float* scale,
                      const float* heuristic, float* product, cudaStream_t st);
int tsp_cost_launch(const float* distances, const uint16_t* tours, int n, int n_ants, int n_colonies, float* costs,
                    uint32_t* neighbours, cudaStream_t st);
int knn_refresh_launch(const float* product, uint8_t* knn, int n, int n_colonies, cudaStream_t st);
bool tsp_update_seq_preferred(int n, int n_ants, int n_colonies);
bool tsp_tail_ok(int n, int n_ants, int n_colonies);
int tsp_tail_launch(float* pheromone, const uint16_t* tours, const float* distances, const float* heuristic, float* product,
                    float* costs, float* lowest, int64_t* shortest, float* ph_max, int n, int n_ants, int n_colonies, float decay,
                    int elitist, int min_max, float ph_min, cudaStream_t st);
int tsp_update_seq_launch(float* pheromone, const uint16_t* tours, const float* costs, int n, int n_ants, int n_colonies,
                          float decay, int elitist, int min_max, float ph_min, const float* ph_max, const float* scale,
                          const float* heuristic, float* product, cudaStream_t st);
int tsp_sample_peers(const float* product, int n, int n_ants_local, int n_colonies, int start_node, int double_norm, uint64_t seed,
                     uint64_t offset, const uint64_t* offsets, const uint8_t* knn, int ant_base, int n_ants_total,
                     const uint64_t* peer_tours_host, int n_peers, cudaStream_t st);
int best_launch(const float* costs, const uint16_t* tours, const float* ph, int n, int tour_len, int A, int B, int min_max,
                float* lowest, int64_t* shortest, float* ph_max, float* scale, const int32_t* tmax, int32_t* shortest_rows,
                cudaStream_t st);
int hadamard_launch(const float* a, const float* b, float* o, size_t n, cudaStream_t st);

struct BarrierParams {
    uint32_t* peer_flags[8];   // peer-mapped address of rank r's flag words (uint32 [8]); [rank] is our own
    int rank, world;
    uint32_t value;            // iteration number being signalled (monotonic, wrap-safe compare)
    unsigned long long timeout_ns;
    int32_t* status;           // set to 1 + (rank waited for) when a wait times out; once set, later barriers do not wait
};

// <<<1, 32>>>: lane r signals rank r and waits for rank r.
__global__ void shard_barrier_kernel(const BarrierParams p) {
    const int r = threadIdx.x;
    if (r >= p.world) return;
    // the tours stored by the sampling kernel before us (same stream) must be visible to a peer that sees the flag
    __threadfence_system();
    uint32_t* remote = p.peer_flags[r] + p.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(p.value) : "memory");
    if (*reinterpret_cast<volatile int32_t*>(p.status) != 0) return;
    const uint32_t* mine = p.peer_flags[p.rank] + r;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - p.value) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > p.timeout_ns) {
            atomicExch(p.status, 1 + r);
            break;
        }
    }
}

}  // namespace deepaco

using namespace deepaco;

extern "C" int deepaco_tsp_run_shard(const deepaco_tsp_run_args* a, const deepaco_shard_args* sh, int n_iterations, void* stream) {
    DACO_CHECK_ARG(a != nullptr && sh != nullptr && n_iterations >= 0, "deepaco_tsp_run_shard: bad arguments");
    DACO_CHECK_ARG(a->pheromone && a->heuristic && a->distances && a->product && a->costs && a->neighbours && a->lowest_cost &&
                       a->shortest_path,
                   "deepaco_tsp_run_shard: NULL buffer");
    DACO_CHECK_ARG(!a->min_max || (a->ph_max && a->scale), "deepaco_tsp_run_shard: min_max needs ph_max and scale buffers");
    DACO_CHECK_ARG(a->local_search == 0, "deepaco_tsp_run_shard: local search is not available on the ant-sharded path");
    DACO_CHECK_ARG(sh->world >= 1 && sh->world <= 8 && sh->rank >= 0 && sh->rank < sh->world,
                   "deepaco_tsp_run_shard: rank %d / world %d (1..8 ranks)", sh->rank, sh->world);
    DACO_CHECK_ARG(sh->peer_tours_host && sh->peer_flags_host && sh->status, "deepaco_tsp_run_shard: NULL peer tables / status");
    DACO_CHECK_ARG(sh->ant_base >= 0 && sh->n_ants_local >= 0 && sh->ant_base + sh->n_ants_local <= a->n_ants,
                   "deepaco_tsp_run_shard: ant shard [%d, %d) outside the colony's %d ants", sh->ant_base,
                   sh->ant_base + sh->n_ants_local, a->n_ants);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = a->n, A = a->n_ants, B = a->n_colonies;
    const uint64_t inc = deepaco_tsp_sample_offset_increment(n, A, a->start_node);
    if (n_iterations > 0 && !a->product_valid) {
        const int rc = hadamard_launch(a->pheromone, a->heuristic, a->product, (size_t)B * n * n, st);
        if (rc) return rc;
    }
    BarrierParams bp{};
    for (int r = 0; r < sh->world; ++r) bp.peer_flags[r] = reinterpret_cast<uint32_t*>(sh->peer_flags_host[r]);
    bp.rank = sh->rank;
    bp.world = sh->world;
    bp.status = sh->status;
    bp.timeout_ns = (unsigned long long)(sh->timeout_ms ? sh->timeout_ms : 2000u) * 1000000ull;
    for (int it = 0; it < n_iterations; ++it) {
        const uint32_t e = sh->epoch + (uint32_t)it;
        const int buf = (int)(e & 1u);
        uint64_t peers[8];
        for (int r = 0; r < sh->world; ++r) peers[r] = sh->peer_tours_host[2 * r + buf];
        const uint16_t* tours = reinterpret_cast<const uint16_t*>(peers[sh->rank]);
        if (a->ev_sample_begin) DACO_CHECK_CUDA(cudaEventRecord((cudaEvent_t)a->ev_sample_begin, st));
        int rc = DEEPACO_OK;
        if (sh->n_ants_local > 0)
            rc = tsp_sample_peers(a->product, n, sh->n_ants_local, B, a->start_node, a->double_norm, a->seed,
                                  a->offset + (uint64_t)it * inc, a->offsets, a->knn, sh->ant_base, A, peers, sh->world, st);
        if (rc) return rc;
        if (a->ev_sample_end) DACO_CHECK_CUDA(cudaEventRecord((cudaEvent_t)a->ev_sample_end, st));
        if (sh->world > 1) {
            bp.value = e + 1u;
            shard_barrier_kernel<<<1, 32, 0, st>>>(bp);
            DACO_CHECK_LAUNCH();
        }
        const bool seq = tsp_update_seq_preferred(n, A, B);
        if (seq && tsp_tail_ok(n, A, B)) {   // cost + best tracking + update in one launch per colony
            rc = tsp_tail_launch(a->pheromone, tours, a->distances, a->heuristic, a->product, a->costs, a->lowest_cost, a->shortest_path,
                                 a->ph_max, n, A, B, a->decay, a->elitist, a->min_max, a->ph_min, st);
            if (rc) return rc;
            if (a->knn && a->knn_refresh > 0 && n > 32 && n <= 256 && (a->knn_iteration0 + it + 1) % a->knn_refresh == 0) {
                rc = knn_refresh_launch(a->product, const_cast<uint8_t*>(a->knn), n, B, st);
                if (rc) return rc;
            }
            continue;
        }
        rc = tsp_cost_launch(a->distances, tours, n, A, B, a->costs, seq ? nullptr : a->neighbours, st);
        if (rc) return rc;
        rc = best_launch(a->costs, tours, a->pheromone, n, n, A, B, a->min_max, a->lowest_cost, a->shortest_path, a->ph_max,
                         a->min_max ? a->scale : nullptr, nullptr, nullptr, st);
        if (rc) return rc;
        rc = seq ? tsp_update_seq_launch(a->pheromone, tours, a->costs, n, A, B, a->decay, a->elitist, a->min_max, a->ph_min, a->ph_max,
                                         a->min_max ? a->scale : nullptr, a->heuristic, a->product, st)
                 : tsp_update_launch(a->pheromone, a->neighbours, a->costs, n, A, B, a->decay, a->elitist, a->min_max, a->ph_min,
                                     a->ph_max, a->min_max ? a->scale : nullptr, a->heuristic, a->product, st);
        if (rc) return rc;
        if (a->knn && a->knn_refresh > 0 && n > 32 && n <= 256 && (a->knn_iteration0 + it + 1) % a->knn_refresh == 0) {
            rc = knn_refresh_launch(a->product, const_cast<uint8_t*>(a->knn), n, B, st);
            if (rc) return rc;
        }
    }
    return DEEPACO_OK;
}
