/* deepaco_b200 -- C ABI of the B200-native DeepACO rollout engine (libdeepaco_b200.so).
 *
 * The reference (henry-yeh/DeepACO) has no FFI of its own: its boundary is the duck-typed Python
 * class `ACO` of each problem directory.  Every entry point below replaces the ATen call sequence
 * of one reference method; the Python classes in deepaco_b200/{tsp,tsp_nls,cvrp}/aco.py bind them
 * with ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - matrices are dense row-major fp32, batched over `n_colonies` independent instances:
 *       pheromone / heuristic / distances : [n_colonies][n][n]
 *       paths  (reference layout)          : [n_colonies][n_rows][n_ants] int64, step-major
 *       tours  (compact layout)            : [n_colonies][n_ants][n] uint16, ant-major
 *   - return value 0 on success, negative DEEPACO_E* otherwise; deepaco_last_error() gives the text;
 *   - no call synchronises the device unless documented ("host" entry points do).
 */
#ifndef DEEPACO_B200_H_
#define DEEPACO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEEPACO_OK 0
#define DEEPACO_EINVAL (-1)   /* bad argument / unsupported size */
#define DEEPACO_ECUDA (-2)    /* CUDA runtime error */
#define DEEPACO_ENODEV (-3)   /* no sm_100 device */

#define DEEPACO_MAX_NODES 1024 /* largest n the sampling kernels are instantiated for */

const char* deepaco_last_error(void);
int deepaco_version(void);
/* number of CUDA kernels this library has launched so far in this process (monotonic counter) */
long long deepaco_kernel_launches(void);

/* Geometry of torch's Philox draw for a tensor of `numel` elements on the current device
 * (ATen/native/cuda/DistributionTemplates.h:50-62): returns grid*256 threads and the generator
 * offset increment one such draw consumes.  Host helper, no device work. */
int deepaco_torch_draw_geometry(int64_t numel, uint32_t* threads_out, uint64_t* offset_increment_out);

/* block_width ATen picks for sum(x[n_rows][row_len], dim=-1) (ATen/native/cuda/Reduce.cuh
 * setReduceConfig); the kernels reproduce that summation order.  *exact_out = 1 when the order is
 * reproduced exactly for this shape. */
int deepaco_aten_sum_plan(int row_len, int n_rows, int* block_width_out, int* vectorized_out, int* exact_out);

/* ---- tour construction ----------------------------------------------------------------------
 * Replaces ACO.gen_path + pick_move: tsp/aco.py:134-177 (start_node = -1, double_norm = 0) and
 * tsp_nls/aco.py:184-220 (start_node = 0, double_norm = 1).
 * Noise: when `noise` is NULL the kernel regenerates, in registers, exactly the Philox words torch's
 * `randint` / `exponential_` kernels would draw from (seed, offset); batched colonies share the seed and
 * take their Philox offset from offsets[b] + offset (offsets: device uint64 [B], NULL = 0) --
 * i.e. colony b sees the stream the reference would have reached had it processed the colonies one
 * after another under one generator.  The caller advances its generator by
 * deepaco_tsp_sample_offset_increment() per colony.
 * With `noise` != NULL ([B][n-1][A][n] Exp(1) draws) and `start` ([B][A], or start_node >= 0) the
 * result is a pure function of its inputs (cross-device parity mode).
 * Outputs (each may be NULL): paths int64 [B][n][A]; log_probs fp32 [B][n-1][A]; tours u16 [B][A][n].
 * knn (optional, uint8 [B][n][32], 32 < n <= 256): per row, 32 distinct columns holding the row's largest
 * products (e.g. topk of the heuristic).  Purely a performance hint for sparse products -- the result is
 * identical with or without it (see csrc/list_kernel.cuh). */
int deepaco_tsp_sample(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                       int start_node, int double_norm, uint64_t seed, uint64_t offset,
                       const uint64_t* offsets, const float* noise, const int64_t* start, int64_t* paths,
                       float* log_probs, uint16_t* tours, const uint8_t* knn, void* stream);
uint64_t deepaco_tsp_sample_offset_increment(int n, int n_ants, int start_node);
/* Ant-sharded construction (multi-GPU): ants [ant_base, ant_base + n_ants) of colonies with n_ants_total ants.
 * Philox subsequences and ATen summation plans are those of the full colony: a given ant's tour is the same
 * whichever GPU builds it.  Outputs are indexed by the local ant number.  The offset increment to advance the
 * generator by is that of the full colony: deepaco_tsp_sample_offset_increment(n, n_ants_total, start_node). */
int deepaco_tsp_sample_shard(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                             int start_node, int double_norm, uint64_t seed, uint64_t offset, const uint64_t* offsets,
                             int64_t* paths, float* log_probs, uint16_t* tours, const uint8_t* knn, int ant_base,
                             int n_ants_total, void* stream);

/* Same, with the multi-GPU exchange fused into the kernel: every finished tour is stored by the building warp into
 * the uint16 [B][n_ants_total][n] tour buffer of each of the n_peers ranks (peer_tours_host[r] = peer-mapped device
 * address of rank r's buffer, our own included; n_peers <= 8, n <= 256).  Afterwards only a barrier is needed. */
int deepaco_tsp_sample_shard_p2p(const float* pheromone, const float* heuristic, int n, int n_ants, int n_colonies,
                                 int start_node, int double_norm, uint64_t seed, uint64_t offset, const uint64_t* offsets,
                                 const uint8_t* knn, int ant_base, int n_ants_total, const uint64_t* peer_tours_host,
                                 int n_peers, void* stream);

/* ---- tour cost  (ACO.gen_path_costs, tsp/aco.py:120-132) --------------------------------------
 * costs[b][a] = sum_k dist[u_k][u_{k-1}] in ATen's summation order.  Input tours either as
 * `paths` (int64 [B][n][A]) or `tours` (u16 [B][A][n]); exactly one non-NULL.  `costs` may be NULL.  Optionally emits
 * neighbours[b][u][a] = (pred << 16) | succ of node u in ant a's tour (uint32 [B][n][A]). */
int deepaco_tsp_cost(const float* distances, const int64_t* paths, const uint16_t* tours, int n, int n_ants,
                     int n_colonies, float* costs, uint32_t* neighbours, void* stream);

/* ---- evaporate + deposit  (ACO.update_pheronome, tsp/aco.py:94-118) ---------------------------
 * In place on `pheromone`.  Deposit order = ant order per matrix cell (bit-exact with the reference's
 * sequential index_put loop).  elitist: only the first arg-min ant deposits.  min_max: clamp to
 * [ph_min, ph_max[b]] afterwards (ph_max device fp32 [B]). */
int deepaco_tsp_update(float* pheromone, const uint32_t* neighbours, const float* costs, int n, int n_ants,
                       int n_colonies, float decay, int elitist, int min_max, float ph_min, const float* ph_max,
                       void* stream);

/* Same update straight from compact tours (u16 [B][A][n]), no neighbour table: ants one after another, each ant's cells in
 * parallel on a pheromone matrix held in shared memory -- the reference's own loop structure (tsp/aco.py:109-114), same
 * bits.  For 3 <= n <= 224 and n_ants <= 1024; deepaco_tsp_run picks it automatically. */
int deepaco_tsp_update_tours(float* pheromone, const uint16_t* tours, const float* costs, int n, int n_ants, int n_colonies,
                             float decay, int elitist, int min_max, float ph_min, const float* ph_max, void* stream);

/* ---- ACO.run for TSP (tsp/aco.py:74-92; tsp_nls/aco.py:104-129 with local_search=None) -------------
 * n_iterations x { sample -> cost -> best tracking -> evaporate+deposit } on `stream`, no host sync.
 * All buffers are caller-owned device memory:
 *   pheromone [B][n][n] in/out; heuristic, distances [B][n][n] in;
 *   product [B][n][n] scratch holding pheromone (.) heuristic (product_valid = 1 if already up to date);
 *   tours u16 [B][A][n], costs f32 [B][A], neighbours u32 [B][n][A]: scratch, hold the LAST iteration (neighbours is
 *   only written when the row-parallel update is used: few colonies, n > 128 or more than 1024 ants);
 *   lowest_cost f32 [B] in/out (+inf before the first run); shortest_path i64 [B][n] in/out;
 *   ph_max f32 [B] in/out (min_max only; 0 = not set yet); scale f32 [B] scratch (min_max only).
 * Iteration t of colony b consumes the Philox stream at offsets[b] + offset + t * increment, increment =
 * deepaco_tsp_sample_offset_increment(): exactly the reference's generator consumption. */
typedef struct {
    int n, n_ants, n_colonies;
    int start_node;   /* -1: torch.randint start (tsp/);  0: fixed start (tsp_nls/) */
    int double_norm;  /* 1 for tsp_nls/ */
    float decay;
    int elitist, min_max;
    float ph_min;
    uint64_t seed, offset;
    const uint64_t* offsets;
    float* pheromone;
    const float* heuristic;
    const float* distances;
    float* product;
    int product_valid;
    uint16_t* tours;
    float* costs;
    uint32_t* neighbours;
    float* lowest_cost;
    int64_t* shortest_path;
    float* ph_max;
    float* scale;
    const uint8_t* knn;    /* optional candidate lists, see deepaco_tsp_sample (rewritten in place when knn_refresh > 0) */
    int local_search;      /* tsp_nls/aco.py:97-102 between construction and cost: 0 none, 1 2-opt, 2 NLS */
    int ls_max_iterations; /* 2-opt passes per call (n/4 in training, 10000 at inference) */
    int T_nls, T_p;        /* NLS rounds / perturbation passes (10 / 20 in the reference) */
    const float* heuristic_dist; /* NLS only: [B][n][n], 1 / (heuristic / rowmax + 1e-5), tsp_nls/aco.py:230-232 */
    void* ev_sample_begin; /* optional cudaEvent_t recorded before / after each sampling launch (profiling) */
    void* ev_sample_end;
    int knn_refresh;       /* R > 0: after every R-th iteration (counted from knn_iteration0) the candidate lists `knn` are
                              REWRITTEN in place from the current product matrix (columns of each row's 32 largest
                              entries); a pure performance measure, results do not depend on it.  0: lists stay as given */
    int knn_iteration0;    /* iterations already run on these lists (the caller's running count) */
    int roulette;          /* 1: construct with the roulette-wheel sampler (run(.., inference=True), tsp_nls/aco.py:106-110);
                              iteration t then consumes the Philox stream at offset + t * deepaco_tsp_roulette_offset_increment() */
} deepaco_tsp_run_args;
int deepaco_tsp_run(const deepaco_tsp_run_args* args, int n_iterations, void* stream);
/* Same with HOST matrices ([B][n][n] fp32 each) -- `ACO(distances, heuristic=...).run(T)` for a caller whose data lives
 * on the host: H2D of distances, heuristic and pheromone into the device buffers of `args`, n_iterations, D2H of
 * lowest_cost [B], shortest_path [B][n] and, if copy_back_pheromone, the pheromone; synchronises `stream` before
 * returning.  pheromone_host may be NULL: the pheromone then starts as ACO.__init__ creates it (ones, times ph_min under
 * min_max; tsp/aco.py:37-40) and is not uploaded.  Uploads run on an internal copy stream, chunk after chunk, while the
 * caller's stream computes the chunks that have landed.
 * Bytes moved: (2 or 3) * B*n*n*4 in; B*4 + B*n*8 (+ B*n*n*4) out. */
int deepaco_tsp_run_host(const deepaco_tsp_run_args* args, int n_iterations, const float* distances_host,
                         const float* heuristic_host, float* pheromone_host, float* lowest_cost_host,
                         int64_t* shortest_path_host, int copy_back_pheromone, void* stream);

/* ---- roulette-wheel construction: the reference's INFERENCE sampler (tsp_nls/aco.py:260-297 _inference_sample /
 * inference_batch_sample; used by sample(inference=True) :81-85 and run(.., inference=True) :106-110).
 * prob = pheromone^alpha (.) heuristic^beta, fp32 [B][n][n].  Per step: rand = U * sum(prob[last] * mask), next = first
 * column whose running sum reaches rand.  U comes from Philox at (seed, offset) -- the reference uses numba's private
 * generator, so parity is statistical.  start_node >= 0: fixed start (the reference passes 0); -1: random start.
 * Outputs (either may be NULL): tours u16 [B][A][n], paths i64 [B][n][A].  Advance the generator by
 * deepaco_tsp_roulette_offset_increment(n, n_ants) per call. */
int deepaco_tsp_roulette_sample(const float* prob, int n, int n_ants, int n_colonies, int start_node, uint64_t seed,
                                uint64_t offset, const uint64_t* offsets, uint16_t* tours, int64_t* paths, void* stream);
uint64_t deepaco_tsp_roulette_offset_increment(int n, int n_ants);

/* ---- ACO.run with the ANTS of every colony split over the GPUs of one box (no reference counterpart: the reference is
 * single-device; semantics reproduced: tsp/aco.py:74-92, bit-identical to deepaco_tsp_run on one GPU for any world size).
 * One process per GPU calls this with the same `args` values (its own device buffers) and its own shard description.
 * Per iteration: this rank builds ants [ant_base, ant_base + n_ants_local) of args->n_ants and the sampling kernel stores
 * each finished tour into the tour buffer of EVERY rank over NVLink (peer-mapped addresses); a one-CTA flag barrier
 * (release / acquire at system scope, no host sync, no NCCL) follows; then every rank replays cost, best tracking and the
 * ordered deposit on all tours.  args->tours is not used: the tour buffers are the peer-mapped ones below.
 *   peer_tours_host [world][2]  host array: address (valid on THIS device) of rank r's tour buffer k, uint16
 *                               [n_colonies][n_ants][n] each; iteration e uses buffer (epoch + e) & 1;
 *   peer_flags_host [world]     host array: address of rank r's flag words, uint32 [8], zero before the first call;
 *   epoch                       iterations already run on these flags (0 for the first call, then += n_iterations);
 *   status                      device int32, zero at entry; non-zero afterwards = a peer did not arrive within
 *                               timeout_ms (default 2000) and the results are invalid;
 * With world = 1 the call degenerates to deepaco_tsp_run on the supplied buffer. */
typedef struct {
    int rank, world;
    int ant_base, n_ants_local;
    const uint64_t* peer_tours_host;
    const uint64_t* peer_flags_host;
    uint32_t epoch;
    uint32_t timeout_ms;
    int32_t* status;
} deepaco_shard_args;
int deepaco_tsp_run_shard(const deepaco_tsp_run_args* args, const deepaco_shard_args* shard, int n_iterations, void* stream);

/* ---- local search (tsp_nls/two_opt.py:6-49, tsp_nls/aco.py:234-258) ---------------------------------
 * In place on compact tours (u16 [B][A][n]); one CTA per tour, bit-exact with the reference's numba code
 * (first strict minimum in (i,j) scan order, fp32 left-to-right delta, threshold -1e-6).
 * deepaco_two_opt  = batched_two_opt_python(dist, tours, max_iterations).
 * deepaco_tsp_nls  = ACO.nls: 2-opt(dist, max_iterations), then T_nls x { 2-opt(heuristic_dist, T_p),
 *                    2-opt(dist, max_iterations), keep if numpy-summed cost strictly improves }.
 * passes_out (optional int32 [B][A]) = total 2-opt passes executed; costs_out (optional f32 [B][A]) = the
 * numpy-order cost of the kept tour. */
int deepaco_two_opt(const float* distances, uint16_t* tours, int n, int n_ants, int n_colonies, int max_iterations,
                    int32_t* passes_out, void* stream);
int deepaco_tsp_nls(const float* distances, const float* heuristic_dist, uint16_t* tours, int n, int n_ants,
                    int n_colonies, int max_iterations, int T_nls, int T_p, float* costs_out, int32_t* passes_out,
                    void* stream);
/* layout conversion between the reference's paths (int64 [B][n][A]) and compact tours (u16 [B][A][n]) */
int deepaco_paths_to_tours(const int64_t* paths, uint16_t* tours, int n, int n_ants, int n_colonies, void* stream);
int deepaco_tours_to_paths(const uint16_t* tours, int64_t* paths, int n, int n_ants, int n_colonies, void* stream);

/* ---- instance -> graph front end (gen_distance_matrix + gen_pyg_data's topk / edge_index: tsp/utils.py:4-36,
 * tsp_nls/utils.py:5-46; the distance part also cvrp/utils.py:18-22) for a batch of instances in one launch.
 * Input: coords f32 [B][n][2] (distances computed with ATen's rounding: sqrt(fl(fl(dx*dx) + fl(dy*dy))), diagonal := diag,
 * 1e9 for TSP / 1e-10 for CVRP) OR distances_in f32 [B][n][n] (taken as is) -- exactly one of them.
 * Outputs, each optional: distances_out f32 [B][n][n]; the k smallest entries of every row in ascending order
 * (torch.topk(distances, k, dim=1, largest=False)) as nbr_index int32 [B][n][k] / nbr_value f32 [B][n][k]; edge_index
 * int64 [B][2][n*k] (row 0: repeat_interleave(arange(n), k), row 1: the flattened neighbour indices).  Bit-equal distances
 * within a row: lowest column first (torch's order among equal values is unspecified).  0 <= k <= n <= 8192. */
int deepaco_knn_graph(const float* coords, const float* distances_in, int n, int n_instances, int k, float diag,
                      float* distances_out, int32_t* nbr_index, float* nbr_value, int64_t* edge_index, void* stream);

/* ---- heuristic network forward, eval mode (Net.forward, tsp/net.py:84-88 -> EmbNet :27-45 -> ParNet :74-75)
 * Graph per instance in CSR-by-source form: row_ptr int32 [B][n+1]; for the edges sorted by source:
 * dst_sorted int32 [B][E], attr_sorted f32 [B][E], order int32 [B][E] (original edge id, used to write
 * heu_out[b][order[e]]).  x: node features f32 [B][n][feats].  weights: packed fp32 (layout in
 * deepaco_b200/net.py:pack_weights; deepaco_gnn_weight_count(feats) values).  node_ws f32 [B][n][192] and
 * edge_ws f32 [B][E][32] are scratch.  heu_out f32 [B][E] = Net.forward(pyg) per original edge (may be NULL).
 * dense_out (optional f32 [B][n][n]) = Net.reshape(pyg, heu) + dense_eps, i.e. the heuristic matrix the drivers
 * hand to ACO (tsp/test.ipynb cell 1), written by the same launch.
 * src_sorted (optional int32 [B][E]): source node of each sorted edge; NULL = found by binary search in row_ptr.
 * The 32x32 linears (tsp/net.py:36-43,62-75) run on the tensor cores (mma.sync TF32 with the 3xTF32 split: fp32-level
 * accuracy, fp32 accumulate); everything else is fp32. */
int64_t deepaco_gnn_weight_count(int feats);
int deepaco_gnn_forward(const float* x, const int32_t* row_ptr, const int32_t* dst_sorted, const float* attr_sorted,
                        const int32_t* order, const float* weights, int n_nodes, int n_edges, int feats,
                        int n_instances, float* node_ws, float* edge_ws, float* heu_out, float* dense_out,
                        float dense_eps, const int32_t* src_sorted, void* stream);

/* ---- heuristic network, TRAINING mode: forward with batch-statistics BatchNorm, and its backward
 * (Net.forward under net.train() as driven by train_instance: tsp/train.ipynb cell 1, tsp_nls/train.py:15-44,
 * cvrp/train.py; PyG BatchNorm(32) == nn.BatchNorm1d over all nodes / all edges of the one graph, eps = bn_eps).
 * Replaces the autograd graph torch builds through tsp/net.py:27-45,62-75 by two launches.
 * Every instance b of the batch is an independent forward call (own batch statistics).
 * Graph arrays as for deepaco_gnn_forward plus src_sorted int32 [B][E] (source of each sorted edge) and, for the
 * backward pass, the same edges grouped by destination: col_ptr int32 [B][n+1], in_edges int32 [B][E] (indices into
 * the source-sorted order).  ctas_per_instance: CTAs cooperating on one graph -- 1, or a thread-block cluster of 2 / 4 / 8
 * (hardware cluster barrier), or 16 / 32 / 64 co-resident CTAs of a cooperative launch (arrival-counter barrier; instances
 * are launched in as many waves as co-residency requires).
 * Buffers (fp32, device):  xs [B][13][n][32], ws [B][13][E][32], zv [B][12][n][32], ze [B][12][E][32] -- activations
 * saved by the forward for the backward;  stats [B][12][6][32] -- per layer the batch mean, 1/sqrt(var + eps) and
 * biased variance of the node BatchNorm, then of the edge BatchNorm (the caller updates running_mean / running_var
 * from them);  node_ws [B][n][224], edge_ws [B][E][96] (backward only), red [B][36][64][128], sync_ws uint32 [B]
 * (zeroed by the call) -- scratch.
 * forward:  heu_out [B][E] = Net.forward(pyg) per ORIGINAL edge id.
 * backward: grad_heu [B][E] = dL/d heu_out;  grad_weights [B][ctas_per_instance][deepaco_gnn_weight_count(feats)],
 *           zero-initialised by the caller, receives partial parameter gradients in the packed weight layout: the
 *           gradient of instance b is the sum over its ctas_per_instance slices (running-stat slots stay 0). */
typedef struct deepaco_gnn_train_args {
    int32_t n_nodes, n_edges, feats, n_instances, ctas_per_instance;
    float bn_eps;
    const float* x;
    const int32_t* row_ptr;
    const int32_t* src_sorted;
    const int32_t* dst_sorted;
    const float* attr_sorted;
    const int32_t* order;
    const int32_t* col_ptr;
    const int32_t* in_edges;
    const float* weights;
    float* xs;
    float* ws;
    float* zv;
    float* ze;
    float* stats;
    float* node_ws;
    float* edge_ws;
    float* red;
    uint32_t* sync_ws;
    float* heu_out;
    const float* grad_heu;
    float* grad_weights;
} deepaco_gnn_train_args;
int deepaco_gnn_train_forward(const deepaco_gnn_train_args* args, void* stream);
int deepaco_gnn_train_backward(const deepaco_gnn_train_args* args, void* stream);
/* Eval-mode Net.forward (tsp/net.py:84-88, running-statistics BatchNorm) by a group of ctas_per_instance CTAs per
 * graph: the low-latency form of deepaco_gnn_forward for one or a few instances (the reference's inference drivers
 * process one instance at a time).  Same argument block; `weights` in the eval packing (mean / invstd slots filled),
 * xs [B][2][n][32], ws [B][2][E][32], node_ws [B][n][128], sync_ws as above, heu_out [B][E]; zv / ze / stats / red /
 * edge_ws / col_ptr / in_edges / grad_* are not used and may be NULL. */
int deepaco_gnn_forward_group(const deepaco_gnn_train_args* args, void* stream);

/* ---- CVRP (cvrp/aco.py:106-205, adaptive = False) ----------------------------------------------
 * Node 0 is the depot; n_nodes = customers + 1; demand fp32 [B][n_nodes] (demand[0] = 0).
 * deepaco_cvrp_sample replaces ACO.gen_path + pick_move + update_visit_mask + update_capacity_mask +
 * check_done (cvrp/aco.py:138-205).  Path buffers have path_rows = 2 * n_nodes rows (upper bound of the
 * data-dependent length); rows past an ant's end are 0.  lens[b][a] = steps the ant took, tmax[b] = max over
 * ants = (rows of the reference's `paths`) - 1.  The reference consumes one [n_ants, n_nodes] exponential_
 * draw per step until the slowest ant is done: advance the generator by tmax * step_offset_increment.
 * noise (optional): [B][path_rows-1][n_ants][n_nodes]. */
int deepaco_cvrp_sample(const float* pheromone, const float* heuristic, const float* demand, float capacity,
                        int n_nodes, int n_ants, int n_colonies, uint64_t seed, uint64_t offset,
                        const uint64_t* offsets, const float* noise, int path_rows, int64_t* paths,
                        float* log_probs, uint16_t* tours, int32_t* lens, int32_t* tmax, void* stream);
uint64_t deepaco_cvrp_step_offset_increment(int n_nodes, int n_ants);

/* ACO.gen_path_costs (cvrp/aco.py:132-136): costs[b][a] = sum_{k<T} dist[u_k][u_{k+1}] over the padded
 * path, T = tmax[b] (device) or T_fixed when tmax is NULL; rows_in = rows of the supplied paths/tours.
 * neighbours (uint32 [B][n_nodes][n_ants], optional): per customer (pred << 16) | succ; row 0 = 1 if the
 * ant's padded path contains a (0,0) pair. */
int deepaco_cvrp_cost(const float* distances, const int64_t* paths, const uint16_t* tours, int n_nodes,
                      int n_ants, int n_colonies, int rows_in, const int32_t* tmax, int T_fixed, float* costs,
                      uint32_t* neighbours, void* stream);

/* ACO.update_pheronome (cvrp/aco.py:106-130): in place; one-directional deposit in ant order, repeated
 * (0,0) pairs count once per ant, optional min_max clamp, final floor at 1e-10. */
int deepaco_cvrp_update(float* pheromone, const uint32_t* neighbours, const float* costs, int n_nodes, int n_ants,
                        int n_colonies, float decay, int elitist, int min_max, float ph_min, const float* ph_max,
                        void* stream);

/* ACO.run for CVRP (cvrp/aco.py:72-104, adaptive = False) on the device, like deepaco_tsp_run.  Path lengths are
 * data dependent, hence so is the generator consumption: `offsets` (device uint64 [B], REQUIRED, in/out) holds each
 * colony's current Philox offset and is advanced by tmax * step_increment after every iteration; read it back to
 * resynchronise a host-side generator.  tours u16 [B][A][2N], lens i32 [B][A], tmax i32 [B], shortest_path i64
 * [B][2N] (zero padded; its length is the tmax of the iteration that produced it + 1). */
typedef struct {
    int n_nodes, n_ants, n_colonies;
    float capacity;
    float decay;
    int elitist, min_max;
    float ph_min;
    uint64_t seed;
    uint64_t* offsets;
    float* pheromone;
    const float* heuristic;
    const float* distances;
    const float* demand;
    float* product;
    int product_valid;
    uint16_t* tours;
    float* costs;
    uint32_t* neighbours;
    int32_t* lens;
    int32_t* tmax;
    float* lowest_cost;
    int64_t* shortest_path;
    int32_t* shortest_rows; /* optional i32 [B]: rows (tmax + 1) of the iteration that produced shortest_path */
    float* ph_max;
    float* scale;
} deepaco_cvrp_run_args;
int deepaco_cvrp_run(const deepaco_cvrp_run_args* args, int n_iterations, void* stream);

/* ---- one construction step with caller-supplied masks: ACO.pick_move (tsp/aco.py:165-177, cvrp/aco.py:167-174)
 * pheromone_pow / heuristic_pow: fp32 [n][n], already raised to alpha / beta (heuristic_pow may be NULL: ones);
 * prev int64 [n_ants]; mask fp32 [n_ants][n]; mask2 fp32 [n_ants][n] or NULL (the CVRP capacity mask).
 * actions int64 [n_ants]; log_probs fp32 [n_ants] or NULL; *bad_prev (device int, caller zeroes) is set if a prev
 * index is outside [0, n).  The draw is the one `Categorical(x).sample()` makes for the [n_ants, n] tensor at generator
 * (seed, offset); the caller advances the generator by deepaco_pick_move_offset_increment(n, n_ants). */
int deepaco_pick_move(const float* pheromone_pow, const float* heuristic_pow, const int64_t* prev, const float* mask,
                      const float* mask2, int n, int n_ants, uint64_t seed, uint64_t offset, int64_t* actions,
                      float* log_probs, int* bad_prev, void* stream);
uint64_t deepaco_pick_move_offset_increment(int n, int n_ants);

/* ---- backward of the log-probabilities (ACO.sample -> REINFORCE loss; tsp/aco.py:165-177, cvrp/aco.py:167-174)
 * Analytic gradient of sum_{t,a} grad_log_probs[t][a] * log_probs[t][a] with respect to the (powered) heuristic
 * and optionally the (powered) pheromone, by replaying the paths (int64 [path_rows][n_ants]).  demand = NULL:
 * TSP masks; demand != NULL: CVRP visit + capacity masks.  Gradients are ACCUMULATED into grad_heuristic /
 * grad_pheromone ([n][n], caller zeroes; grad_pheromone may be NULL) in a fixed order per matrix element (ants
 * ascending, steps ascending; no atomics): the result is bit-identical run to run. */
int deepaco_logp_backward(const float* pheromone_pow, const float* heuristic_pow, const int64_t* paths,
                          const float* grad_log_probs, int n, int n_ants, int path_rows, const float* demand,
                          float capacity, float* grad_heuristic, float* grad_pheromone, void* stream);

/* ---- debug / probe entry points (used by tests to validate the torch-parity assumptions) ------ */
int deepaco_debug_exponential(uint64_t seed, uint64_t offset, int64_t numel, float* out, void* stream);
int deepaco_debug_randint(uint64_t seed, uint64_t offset, int64_t numel, int64_t high, int64_t* out, void* stream);
int deepaco_debug_row_sum(const float* x, int n_rows, int row_len, float* out, void* stream);
/* Number of 32-bit Philox words for which the kernels' shortened Exp(1) transform differs from the literal
 * ATen form (TransformationHelper.h:129-146); must be 0.  mismatches: device uint64. */
int deepaco_debug_exp_guard(uint64_t* mismatches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPACO_B200_H_ */
