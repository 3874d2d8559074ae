// ACO.run for TSP colonies without returning to the host between iterations
// (reference tsp/aco.py:74-92 and tsp_nls/aco.py:104-129 without local search).
//
// Per iteration, on one stream:  K1 sample (tours)  ->  cost + neighbour table  ->  best tracking
// (lowest_cost / shortest_path / MMAS max, all kept on the device)  ->  K2 evaporate + deposit, which also
// emits next iteration's product matrix P = pheromone (.) heuristic so that K1 stages a single matrix.
// The reference's `if best_cost < self.lowest_cost` is a host-side branch on a device value (one sync per
// iteration); here the comparison runs in best_kernel.
#include "common.cuh"
#include "host_util.h"

#include <stdlib.h>

#include <algorithm>

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
int tsp_sample_fused(const float* product, int n, int n_ants, int n_colonies, int start_node, int double_norm, uint64_t seed,
                     uint64_t offset, const uint64_t* offsets, uint16_t* tours, const uint8_t* knn, const float* dist, float* costs,
                     uint32_t* nbr, int* fused, cudaStream_t st);

__global__ void hadamard2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        o[i] = __fmul_rn(a[i], b[i]);
}

int hadamard_launch(const float* a, const float* b, float* o, size_t n, cudaStream_t st) {
    hadamard2_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(a, b, o, n);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

// one CTA per colony: iteration best -> running best, MMAS bookkeeping
__global__ void __launch_bounds__(256) tsp_best_kernel(const float* __restrict__ costs, const uint16_t* __restrict__ tours,
                                                       const float* __restrict__ ph, int n, int tour_len, int A, int min_max,
                                                       float* __restrict__ lowest, int64_t* __restrict__ shortest,
                                                       float* __restrict__ ph_max, float* __restrict__ scale,
                                                       const int32_t* __restrict__ tmax, int32_t* __restrict__ shortest_rows) {
    __shared__ float s_c[32];
    __shared__ int s_i[32];
    __shared__ int s_improved, s_first;
    __shared__ float s_m[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* C = costs + (size_t)b * A;
    float bc = INFINITY;
    int bi = 0x7fffffff;
    for (int a = tid; a < A; a += blockDim.x) {
        const float c = C[a];
        if (c < bc) { bc = c; bi = a; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
        const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
        if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
    if (lane == 0) { s_c[warp] = bc; s_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bc = lane < (blockDim.x >> 5) ? s_c[lane] : INFINITY;
        bi = lane < (blockDim.x >> 5) ? s_i[lane] : 0x7fffffff;
        for (int off = 16; off > 0; off >>= 1) {
            const float oc = __shfl_xor_sync(DACO_FULL, bc, off);
            const int oi = __shfl_xor_sync(DACO_FULL, bi, off);
            if (oc < bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
        }
        if (lane == 0) {
            s_improved = bc < lowest[b];   // tsp/aco.py:81
            s_first = min_max && !(ph_max[b] > 0.f);   // MMAS max not set yet (marker 0)
            s_c[0] = bc;
            s_i[0] = bi;
        }
    }
    __syncthreads();
    if (scale && tid == 0) scale[b] = 1.0f;
    if (!s_improved) return;
    bc = s_c[0];
    bi = s_i[0];
    const uint16_t* t = tours + ((size_t)b * A + bi) * tour_len;
    for (int k = tid; k < tour_len; k += blockDim.x) shortest[(size_t)b * tour_len + k] = (int64_t)t[k];
    if (tid == 0) {
        lowest[b] = bc;
        if (shortest_rows) shortest_rows[b] = tmax[b] + 1;   // CVRP: rows of that iteration's `paths`
    }
    if (min_max) {
        // max = problem_size / lowest_cost  ==  reciprocal(lowest) * n   (Tensor.__rtruediv__)
        const float new_max = __fmul_rn(__fdiv_rn(1.0f, bc), (float)n);
        if (s_first) {
            // self.pheromone *= max / self.pheromone.max()
            float m = -INFINITY;
            const float* P = ph + (size_t)b * n * n;
            for (int i = tid; i < n * n; i += blockDim.x) m = fmaxf(m, P[i]);
            for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(DACO_FULL, m, off));
            if (lane == 0) s_m[warp] = m;
            __syncthreads();
            if (warp == 0) {
                m = lane < (blockDim.x >> 5) ? s_m[lane] : -INFINITY;
                for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(DACO_FULL, m, off));
                if (lane == 0) scale[b] = __fdiv_rn(new_max, m);
            }
        }
        if (tid == 0) ph_max[b] = new_max;
    }
}

int best_launch(const float* costs, const uint16_t* tours, const float* ph, int n, int tour_len, int A, int B, int min_max,
                float* lowest, int64_t* shortest, float* ph_max, float* scale, const int32_t* tmax, int32_t* shortest_rows,
                cudaStream_t st) {
    tsp_best_kernel<<<B, 256, 0, st>>>(costs, tours, ph, n, tour_len, A, min_max, lowest, shortest, ph_max, scale, tmax,
                                       shortest_rows);
    DACO_CHECK_LAUNCH();
    return DEEPACO_OK;
}

}  // namespace deepaco

using namespace deepaco;

extern "C" int deepaco_tsp_run(const deepaco_tsp_run_args* a, int n_iterations, void* stream) {
    DACO_CHECK_ARG(a != nullptr && n_iterations >= 0, "deepaco_tsp_run: bad arguments");
    DACO_CHECK_ARG(a->pheromone && a->heuristic && a->distances && a->product && a->tours && a->costs && a->neighbours &&
                       a->lowest_cost && a->shortest_path,
                   "deepaco_tsp_run: NULL buffer");
    DACO_CHECK_ARG(!a->min_max || (a->ph_max && a->scale), "deepaco_tsp_run: min_max needs ph_max and scale buffers");
    DACO_CHECK_ARG(a->local_search >= 0 && a->local_search <= 2 && (a->local_search != 2 || a->heuristic_dist),
                   "deepaco_tsp_run: bad local_search / missing heuristic_dist");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = a->n, A = a->n_ants, B = a->n_colonies;
    const uint64_t inc = a->roulette ? deepaco_tsp_roulette_offset_increment(n, A) : deepaco_tsp_sample_offset_increment(n, A, a->start_node);
    const size_t cnt = (size_t)B * n * n;
    if (n_iterations > 0 && !a->product_valid) {
        hadamard2_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(a->pheromone, a->heuristic,
                                                                                                a->product, cnt);
        DACO_CHECK_LAUNCH();
    }
    for (int it = 0; it < n_iterations; ++it) {
        if (a->ev_sample_begin) DACO_CHECK_CUDA(cudaEventRecord((cudaEvent_t)a->ev_sample_begin, st));
        const bool seq = tsp_update_seq_preferred(n, A, B);      // ant-sequential update from the tours: no neighbour table needed
        int fused = 0;
        int rc = a->roulette ? deepaco_tsp_roulette_sample(a->product, n, A, B, a->start_node, a->seed, a->offset + (uint64_t)it * inc,
                                                           a->offsets, a->tours, nullptr, st)
                             : tsp_sample_fused(a->product, n, A, B, a->start_node, a->double_norm, a->seed, a->offset + (uint64_t)it * inc,
                                  a->offsets, a->tours, a->knn, a->local_search ? nullptr : a->distances, a->costs,
                                  seq ? nullptr : a->neighbours, &fused, st);
        if (rc) return rc;
        if (a->ev_sample_end) DACO_CHECK_CUDA(cudaEventRecord((cudaEvent_t)a->ev_sample_end, st));
        if (a->local_search == 1) {
            rc = deepaco_two_opt(a->distances, a->tours, n, A, B, a->ls_max_iterations, nullptr, st);
            if (rc) return rc;
        } else if (a->local_search == 2) {
            rc = deepaco_tsp_nls(a->distances, a->heuristic_dist, a->tours, n, A, B, a->ls_max_iterations, a->T_nls, a->T_p, nullptr,
                                 nullptr, st);
            if (rc) return rc;
        }
        if (!fused && seq && tsp_tail_ok(n, A, B)) {   // cost + best tracking + update in one launch per colony
            rc = tsp_tail_launch(a->pheromone, a->tours, a->distances, a->heuristic, a->product, a->costs, a->lowest_cost,
                                 a->shortest_path, a->ph_max, n, A, B, a->decay, a->elitist, a->min_max, a->ph_min, st);
            if (rc) return rc;
            if (a->knn && a->knn_refresh > 0 && n > 32 && n <= 256 && (a->knn_iteration0 + it + 1) % a->knn_refresh == 0) {
                rc = knn_refresh_launch(a->product, const_cast<uint8_t*>(a->knn), n, B, st);
                if (rc) return rc;
            }
            continue;
        }
        if (!fused) {
            rc = tsp_cost_launch(a->distances, a->tours, n, A, B, a->costs, seq ? nullptr : a->neighbours, st);
            if (rc) return rc;
        }
        tsp_best_kernel<<<B, 256, 0, st>>>(a->costs, a->tours, a->pheromone, n, n, A, a->min_max, a->lowest_cost, a->shortest_path,
                                           a->ph_max, a->min_max ? a->scale : nullptr, nullptr, nullptr);
        DACO_CHECK_LAUNCH();
        rc = seq ? tsp_update_seq_launch(a->pheromone, a->tours, a->costs, n, A, B, a->decay, a->elitist, a->min_max, a->ph_min,
                                         a->ph_max, a->min_max ? a->scale : nullptr, a->heuristic, a->product, st)
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

// Host-buffer entry point (what a CPU-side caller of the reference's `ACO(distances, heuristic=...).run(T)` would bind):
// copies the matrices of every colony to the device buffers named in `a`, runs, copies the results back, synchronises.
// Colonies are independent, so the batch is cut into chunks and pipelined over four streams:
//   upload stream    H2D of chunk after chunk, back to back (PCIe never idles, no two uploads compete);
//   caller's stream + one internal compute stream, alternating: chunk c starts as soon as its upload has landed, and the
//                    tail of one chunk's launches overlaps the head of the next;
//   download stream  D2H of a chunk's results while later chunks still compute.
// The first and last chunks are small (only their upload / download is exposed), the middle ones large.
struct AuxStream {   // one per device ordinal, created on first use
    cudaStream_t up = nullptr, comp = nullptr, down = nullptr;
    cudaEvent_t fork = nullptr, comp_done = nullptr, down_done = nullptr;
    cudaEvent_t ready[8] = {}, done[8] = {};
};
static AuxStream g_aux[64];

__global__ void fill_kernel(float* __restrict__ o, float v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = v;
}

extern "C" int deepaco_tsp_run_host(const deepaco_tsp_run_args* a, int n_iterations, const float* distances_host,
                                    const float* heuristic_host, float* pheromone_host, float* lowest_cost_host,
                                    int64_t* shortest_path_host, int copy_back_pheromone, void* stream) {
    DACO_CHECK_ARG(a && distances_host && heuristic_host && lowest_cost_host && shortest_path_host,
                   "deepaco_tsp_run_host: NULL argument");
    DACO_CHECK_ARG(pheromone_host || !copy_back_pheromone, "deepaco_tsp_run_host: copy_back_pheromone needs pheromone_host");
    cudaStream_t st = (cudaStream_t)stream;
    const DeviceInfo* di = device_info();
    if (!di) return DEEPACO_ENODEV;
    AuxStream& aux = g_aux[di->device];
    if (!aux.up) {
        DACO_CHECK_CUDA(cudaStreamCreateWithFlags(&aux.up, cudaStreamNonBlocking));
        DACO_CHECK_CUDA(cudaStreamCreateWithFlags(&aux.comp, cudaStreamNonBlocking));
        DACO_CHECK_CUDA(cudaStreamCreateWithFlags(&aux.down, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&aux.fork, &aux.comp_done, &aux.down_done}) DACO_CHECK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (auto& e : aux.ready) DACO_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : aux.done) DACO_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int B = a->n_colonies, n = a->n, A = a->n_ants;
    // chunk boundaries: 1/16 | 4/16 | 5/16 | 5/16 | 1/16 of a large batch (the best of the measured partitions,
    // profiles/r02_e2e_probe.txt), halves of a medium one
    int cuts[9] = {0, B, 0, 0, 0, 0, 0, 0, 0};
    int chunks = 1;
    if (B >= 64) { chunks = 5; cuts[1] = B / 16; cuts[2] = 5 * B / 16; cuts[3] = 10 * B / 16; cuts[4] = B - B / 16; cuts[5] = B; }
    else if (B >= 8) { chunks = 2; cuts[1] = B / 2; cuts[2] = B; }
    if (const char* e = getenv("DEEPACO_HOST_CHUNKS")) {   // equal chunks, for experiments
        const int c = atoi(e);
        if (c >= 1 && c <= 8 && c <= B) { chunks = c; for (int k = 0; k <= c; ++k) cuts[k] = (int)((long)B * k / c); }
    }
    if (const char* e = getenv("DEEPACO_HOST_CUTS")) {     // explicit ascending boundaries "16,72,224", for experiments
        int c = 0, prev = 0;
        bool ok = true;
        int tmp[9] = {0};
        for (const char* q = e; ok && *q && c < 7;) {
            char* end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q || v <= prev || v >= B) { ok = false; break; }
            tmp[++c] = prev = (int)v;
            q = (*end == ',') ? end + 1 : end;
            if (*end && *end != ',') ok = false;
        }
        if (ok && c >= 1) { chunks = c + 1; tmp[chunks] = B; for (int k = 0; k <= chunks; ++k) cuts[k] = tmp[k]; }
    }
    const size_t mat1 = (size_t)n * n;
    // the internal streams start after everything already queued on the caller's stream (the buffers may be in use)
    DACO_CHECK_CUDA(cudaEventRecord(aux.fork, st));
    for (cudaStream_t s : {aux.up, aux.comp, aux.down}) DACO_CHECK_CUDA(cudaStreamWaitEvent(s, aux.fork, 0));
    for (int c = 0; c < chunks; ++c) {
        const int b0 = cuts[c], nb = cuts[c + 1] - cuts[c];
        const size_t bytes = (size_t)nb * mat1 * sizeof(float);
        DACO_CHECK_CUDA(cudaMemcpyAsync(const_cast<float*>(a->distances) + b0 * mat1, distances_host + b0 * mat1, bytes, cudaMemcpyHostToDevice, aux.up));
        DACO_CHECK_CUDA(cudaMemcpyAsync(const_cast<float*>(a->heuristic) + b0 * mat1, heuristic_host + b0 * mat1, bytes, cudaMemcpyHostToDevice, aux.up));
        if (pheromone_host)
            DACO_CHECK_CUDA(cudaMemcpyAsync(a->pheromone + b0 * mat1, pheromone_host + b0 * mat1, bytes, cudaMemcpyHostToDevice, aux.up));
        DACO_CHECK_CUDA(cudaEventRecord(aux.ready[c], aux.up));
    }
    for (int c = 0; c < chunks; ++c) {
        const int b0 = cuts[c], nb = cuts[c + 1] - cuts[c];
        cudaStream_t cs = (c & 1) ? aux.comp : st;
        deepaco_tsp_run_args b = *a;
        b.n_colonies = nb;
        b.product_valid = 0;
        b.offsets = a->offsets ? a->offsets + b0 : nullptr;
        b.pheromone = a->pheromone + b0 * mat1;
        b.heuristic = a->heuristic + b0 * mat1;
        b.distances = a->distances + b0 * mat1;
        b.product = a->product + b0 * mat1;
        b.tours = a->tours + (size_t)b0 * A * n;
        b.costs = a->costs + (size_t)b0 * A;
        b.neighbours = a->neighbours + (size_t)b0 * n * A;
        b.lowest_cost = a->lowest_cost + b0;
        b.shortest_path = a->shortest_path + (size_t)b0 * n;
        b.ph_max = a->ph_max ? a->ph_max + b0 : nullptr;
        b.scale = a->scale ? a->scale + b0 : nullptr;
        b.knn = a->knn ? a->knn + (size_t)b0 * n * 32 : nullptr;
        b.heuristic_dist = a->heuristic_dist ? a->heuristic_dist + b0 * mat1 : nullptr;
        b.ev_sample_begin = b.ev_sample_end = nullptr;
        const size_t cnt = (size_t)nb * mat1;
        if (!pheromone_host) {   // ACO.__init__ (tsp/aco.py:37-40): pheromone = ones (* min under min_max)
            fill_kernel<<<(unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 8), 256, 0, cs>>>(b.pheromone, a->min_max ? a->ph_min : 1.0f, cnt);
            DACO_CHECK_LAUNCH();
        }
        DACO_CHECK_CUDA(cudaStreamWaitEvent(cs, aux.ready[c], 0));
        const int rc = deepaco_tsp_run(&b, n_iterations, cs);
        if (rc) return rc;
        DACO_CHECK_CUDA(cudaEventRecord(aux.done[c], cs));
        DACO_CHECK_CUDA(cudaStreamWaitEvent(aux.down, aux.done[c], 0));
        if (copy_back_pheromone)
            DACO_CHECK_CUDA(cudaMemcpyAsync(pheromone_host + b0 * mat1, b.pheromone, cnt * sizeof(float), cudaMemcpyDeviceToHost, aux.down));
        DACO_CHECK_CUDA(cudaMemcpyAsync(lowest_cost_host + b0, b.lowest_cost, sizeof(float) * nb, cudaMemcpyDeviceToHost, aux.down));
        DACO_CHECK_CUDA(cudaMemcpyAsync(shortest_path_host + (size_t)b0 * n, b.shortest_path, sizeof(int64_t) * nb * n,
                                        cudaMemcpyDeviceToHost, aux.down));
    }
    // join: the caller's stream continues after the internal compute stream and the downloads
    DACO_CHECK_CUDA(cudaEventRecord(aux.comp_done, aux.comp));
    DACO_CHECK_CUDA(cudaEventRecord(aux.down_done, aux.down));
    DACO_CHECK_CUDA(cudaStreamWaitEvent(st, aux.comp_done, 0));
    DACO_CHECK_CUDA(cudaStreamWaitEvent(st, aux.down_done, 0));
    DACO_CHECK_CUDA(cudaStreamSynchronize(st));
    return DEEPACO_OK;
}
