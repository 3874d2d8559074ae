// Tour-construction kernel shared by TSP and CVRP (included by tsp_sample.cu and cvrp.cu).
//
// One warp per ant.  The product matrix P = pheromone (.) heuristic of the ant's colony sits in shared
// memory (TMA bulk copy per CTA).  The ant's candidate list (unvisited nodes; for CVRP slot 0 is the depot)
// lives in REGISTERS: lane l owns slots l, l+32, ...; removing the chosen node moves the last list entry
// into its slot.  Per step and candidate j the score is x_j / q_j with x_j = P[cur][j] and q_j the Exp(1)
// variate torch's `exponential_` would have written at element (ant, j) of that step's [n_ants, n] draw
// (Philox4x32-10 regenerated in registers).  Because q depends only on (step, ant, j) and not on the
// previous choice, the reciprocal noise of step t+1 is computed while the warp reductions of step t are
// in flight: the dependent chain per step is LDS -> FMUL -> REDUX -> VOTE -> SHFL.
//
// Exactness.  The reference takes argmax_j fl(fl(x_j / S) / q_j) (S = ATen row sum).  We rank by
// A_j = x_j * rcp.approx(q_j) (relative error < 2^-21 against the reference value scaled by S) and accept
// the top candidate only when no other lies within 2^-18 relative -- then the exactly-rounded ranking
// provably has the same winner.  Otherwise (probability ~1e-6 per step, or an all-zero row) the step is
// replayed by exact_step() with the reference's arithmetic.  S is computed only for log-probs.
#pragma once
#include "common.cuh"
#include "sample_common.cuh"

namespace deepaco {

struct ListParams {
    const float* ph;        // [B][n][n]
    const float* heu;       // [B][n][n] or null (then `ph` already is the product)
    int n, A, B;
    int rows;               // rows of the path buffers: n (TSP) or 2n (CVRP)
    int start_node;         // TSP: >= 0 fixed start; -1 -> `start` tensor or torch randint stream
    int double_norm;        // TSP-NLS pre-normalisation
    uint64_t seed, offset;  // one seed for the launch; per-colony Philox offsets from `offsets` when non-null
    const uint64_t* offsets;   // [B] or null
    PhiloxRoundKeys keys;   // round keys of `seed`, host-computed: read straight from the constant bank
    const float* noise;     // [B][rows-1][A][n] external Exp(1) draws, or null
    const uint8_t* knn;     // [B][n][32] per-row candidate columns for the kNN kernel, or null
    const float* dist;      // optional fused epilogue (kNN kernel): distances [B][n][n] ->
    float* costs;           //   costs [B][A] in ATen summation order and
    uint32_t* nbr;          //   neighbour table [B][n][A] (pred << 16 | succ), as deepaco_tsp_cost would produce
    uint16_t* peer_tours[8]; // fused exchange (ant sharding over NVLink): every finished tour is stored straight into the
    int n_peers;             //   [B][A_total][n] tour buffer of each of the n_peers GPUs (peer-mapped pointers, incl. our own)
    int A_total;
    int ant_base;           // index of this launch's ant 0 in the colony (ant sharding across GPUs); Philox uses global indices
    const int64_t* start;   // [B][A] or null
    int64_t* paths;         // [B][rows][A] or null
    float* logp;            // [B][rows-1][A] or null
    uint16_t* tours;        // [B][A][rows] or null
    int32_t* lens;          // CVRP [B][A]
    int32_t* tmax;          // CVRP [B]
    const float* demand;    // CVRP [B][n]
    float capacity;
    int lbw, vec;           // ATen summation plan for a length-n row
    DrawGeom g_noise, g_start;
    uint32_t start_increment, step_increment;
};

#ifndef DEEPACO_CPU_EMU
__device__ __forceinline__ uint32_t lds_s8(uint32_t addr) {
    int32_t v;
    asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return (uint32_t)v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}

#endif  // DEEPACO_CPU_EMU

// reciprocal Exp(1) noise of element `sub` of the draw at Philox counter (ctr_lo, ctr_hi)
__device__ __forceinline__ float noise_rcp(uint32_t ctr_lo, uint32_t ctr_hi, uint32_t sub, const PhiloxRoundKeys& K) {
    return rcp_approx(exp1_from_word(philox_word_x(ctr_lo, ctr_hi, sub, K)));
}

// Component COMP (0..3 = .x .. .w) of the Philox4x32-10 block at (counter, subsequence): torch's draw kernels hand
// component (li / threads) % 4 of call (li / threads) / 4 to element li once a tensor has more elements than the launch
// has threads (DistributionTemplates.h:65-89).  COMP = 0 compiles to the same instructions as philox_word_x.
template <int COMP>
__device__ __forceinline__ uint32_t philox_word_comp(uint32_t ctr_lo, uint32_t ctr_hi, uint32_t sub, const PhiloxRoundKeys& K) {
    if (COMP == 0) return philox_word_x(ctr_lo, ctr_hi, sub, K);
    uint32_t c0 = ctr_lo, c1 = ctr_hi, c2 = sub, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(kPhiloxM0, c0), l0 = kPhiloxM0 * c0;
        const uint32_t h1 = __umulhi(kPhiloxM1, c2), l1 = kPhiloxM1 * c2;
        c0 = h1 ^ c1 ^ K.a[r];
        c2 = h0 ^ c3 ^ K.b[r];
        c1 = l1;
        c3 = l0;
    }
    return COMP == 1 ? c1 : (COMP == 2 ? c2 : c3);
}
template <int COMP>
__device__ __forceinline__ float noise_rcp_comp(uint32_t ctr_lo, uint32_t ctr_hi, uint32_t sub, const PhiloxRoundKeys& K) {
    return rcp_approx(exp1_from_word(philox_word_comp<COMP>(ctr_lo, ctr_hi, sub, K)));
}
// Any geometry, run-time: element li = q0 * threads + r0 + j of a draw whose call 0 has counter (ctr_lo, ctr_hi).
// An ant's n consecutive elements cross a multiple of `threads` at most once, so no division is needed per column.
__device__ __forceinline__ float noise_rcp_general(uint32_t ctr_lo, uint32_t ctr_hi, uint32_t q0, uint32_t r0, uint32_t j,
                                                   uint32_t threads, const PhiloxRoundKeys& K) {
    uint32_t sub = r0 + j, q = q0;
    if (sub >= threads) {
        sub -= threads;
        q += 1u;
    }
    const uint32_t call = q >> 2, comp = q & 3u;
    const uint32_t lo = ctr_lo + call, hi = ctr_hi + (lo < ctr_lo ? 1u : 0u);
    uint32_t c0 = lo, c1 = hi, c2 = sub, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(kPhiloxM0, c0), l0 = kPhiloxM0 * c0;
        const uint32_t h1 = __umulhi(kPhiloxM1, c2), l1 = kPhiloxM1 * c2;
        c0 = h1 ^ c1 ^ K.a[r];
        c2 = h0 ^ c3 ^ K.b[r];
        c1 = l1;
        c3 = l0;
    }
    const uint32_t w = comp == 0u ? c0 : (comp == 1u ? c1 : (comp == 2u ? c2 : c3));
    return rcp_approx(exp1_from_word(w));
}

#ifndef DEEPACO_CPU_EMU
// opaque register copy: stops the compiler from re-deriving a shared-memory address inside the step loop
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}

#endif  // DEEPACO_CPU_EMU

// GLOBAL_P: the product matrix stays in global memory (L2-resident) instead of shared memory -- colonies with more
// than ~230 nodes; `p.ph` must then already be the product (p.heu == nullptr).
template <int EPL, bool CVRP, bool WANT_LOGP, bool EXT_NOISE, bool GLOBAL_P = false>
__global__ void __launch_bounds__(GLOBAL_P ? 256 : 512, GLOBAL_P ? 1 : 2) aco_list_kernel(const __grid_constant__ ListParams p) {
    DACO_DYN_SMEM128(smem);
    __shared__ uint64_t bar;
    const int n = p.n, R = p.rows;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int W = nthreads >> 5, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int a0 = blockIdx.x * W;
    const int a = a0 + warp;
    // shared layout: P [n*n f32] | demand [n f32] (CVRP) | tours [W][R] u16 | cand [W][n] u16 | alive [W][32] u32
    const size_t pbytes = GLOBAL_P ? 0 : ((((size_t)n * n * 4) + 15) & ~(size_t)15);
    const float* Psm = GLOBAL_P ? p.ph + (size_t)b * n * n : reinterpret_cast<const float*>(smem);
    const size_t dbytes = CVRP ? ((((size_t)n * 4) + 15) & ~(size_t)15) : 0;
    float* dem = reinterpret_cast<float*>(smem + pbytes);
    uint16_t* tour_all = reinterpret_cast<uint16_t*>(smem + pbytes + dbytes);
    uint16_t* cand_all = tour_all + (size_t)W * R;
    uint32_t* alive_all = reinterpret_cast<uint32_t*>(smem + pbytes + dbytes + (((size_t)W * (R + n) * 2 + 15) & ~(size_t)15));
    uint16_t* tour_sm = tour_all + (size_t)warp * R;
    uint32_t* alive = alive_all + warp * 32;
    const uint32_t P_addr = GLOBAL_P ? 0u : pin_u32(smem_u32(smem));
    const uint32_t dem_addr = pin_u32(smem_u32(dem));
    const uint32_t cand_addr = pin_u32(smem_u32(cand_all + (size_t)warp * n));
    const uint32_t tour_addr = pin_u32(smem_u32(tour_sm));

    if (CVRP)
        for (int i = tid; i < n; i += nthreads) dem[i] = p.demand[(size_t)b * n + i];
    if (GLOBAL_P) __syncthreads();
    else stage_product(reinterpret_cast<float*>(smem), p.ph, p.heu, n, b, &bar);   // ends with __syncthreads()

    if (a < p.A) {
        const uint64_t seed = p.seed;
        const uint64_t offset0 = (p.offsets ? p.offsets[b] : 0ull) + p.offset;
        const PhiloxRoundKeys& K = p.keys;
        const uint32_t sub_base = (uint32_t)(a + p.ant_base) * (uint32_t)n;
        const float eps = 1.1920928955078125e-07f;
        const float kGap = 1.0f - 3.814697265625e-06f;   // 1 - 2^-18

        int cur = 0;
        uint64_t off_noise = offset0;
        if (!CVRP) {
            if (p.start_node >= 0) {
                cur = p.start_node;
            } else if (p.start) {
                cur = (int)p.start[(size_t)b * p.A + a];
            } else {
                // torch.randint(0, n, (A,)): element a <- curand4().x % n  (random_from_to_kernel, 32-bit branch)
                cur = (int)(torch_philox_word(seed, offset0, (uint64_t)(a + p.ant_base), p.g_start) % (uint32_t)n);
                off_noise += p.start_increment;
            }
        }
        // candidate list: TSP = all nodes but the start (index order); CVRP = depot in slot 0, then customers
        int cnt = CVRP ? n : n - 1;
        uint32_t cj[EPL];
        float r[EPL];
        const uint32_t ctr0_lo = (uint32_t)(off_noise >> 2), ctr0_hi = (uint32_t)(off_noise >> 34);
#pragma unroll
        for (int k = 0; k < EPL; ++k) {
            const int s = lane + 32 * k;
            cj[k] = CVRP ? (uint32_t)s : (uint32_t)(s < cur ? s : s + 1);
            if (s < cnt) sts_u16(cand_addr + 2 * s, cj[k]);
            r[k] = 0.f;
            if (s < cnt && !EXT_NOISE) r[k] = noise_rcp(ctr0_lo, ctr0_hi, sub_base + cj[k], K);
        }
        if (WANT_LOGP && !CVRP) {
            uint32_t m = 0;
            for (int k = 0; k < 32; ++k) {
                const int j = lane * 32 + k;
                if (j < n && j != cur) m |= 1u << k;
            }
            alive[lane] = m;
        }
        if (lane == 0) sts_u16(tour_addr, (uint32_t)cur);
        __syncwarp();

        float used = 0.f;
        if (CVRP) used = __fadd_rn(0.f, lds_f32(dem_addr));   // update_capacity_mask(cur = depot)
        int step = 0;
        const int max_steps = R - 1;
#pragma unroll 1
        while (CVRP ? (!(cnt == 1 && cur == 0) && step < max_steps) : (step < n - 1)) {
            const uint32_t row_addr = P_addr + (uint32_t)cur * (uint32_t)n * 4u;
            const uint64_t off_step = off_noise + (uint64_t)p.step_increment * (uint64_t)step;
            const float* nz = EXT_NOISE ? p.noise + (((size_t)b * (R - 1) + step) * p.A + a) * (size_t)n : nullptr;
            const uint32_t lastj = lds_u16(cand_addr + 2 * (cnt - 1));   // entry that fills the freed slot
            const float remaining = CVRP ? __fsub_rn(p.capacity, used) : 0.f;
            const bool depot_ok = CVRP && ((cur != 0) || (cnt == 1));

            // ---- scores of this step
            float bestA = 0.f, second = 0.f, bestx = 0.f;
            uint32_t bestsj = 0xffffffffu;   // (slot << 16) | node
            bool okk[EPL];
#pragma unroll
            for (int k = 0; k < EPL; ++k) {
                const int s = lane + 32 * k;
                bool ok = s < cnt;
                if (CVRP && ok) ok = (s == 0) ? depot_ok : !(lds_f32(dem_addr + 4 * cj[k]) > remaining);
                okk[k] = ok;
                const float x = ok ? (GLOBAL_P ? __ldg(Psm + (size_t)cur * n + cj[k]) : lds_f32(row_addr + 4 * cj[k])) : 0.f;
                const float rr = EXT_NOISE ? (ok ? rcp_approx(nz[cj[k]]) : 0.f) : r[k];
                const float A = __fmul_rn(x, rr);
                if (A > bestA) {
                    second = bestA;
                    bestA = A;
                    bestsj = ((uint32_t)s << 16) | cj[k];
                    bestx = x;
                } else if (A > second) {
                    second = A;
                }
            }
            const uint32_t mybits = __float_as_uint(bestA);
            const uint32_t topbits = __reduce_max_sync(DACO_FULL, mybits);

            // ---- noise of the NEXT step for the current list (independent of this step's outcome)
            float rn[EPL];
            {
                const uint64_t off_next = off_step + p.step_increment;
                const uint32_t nlo = (uint32_t)(off_next >> 2), nhi = (uint32_t)(off_next >> 34);
#pragma unroll
                for (int k = 0; k < EPL; ++k) {
                    rn[k] = 0.f;
                    if (32 * k < cnt && !EXT_NOISE) rn[k] = noise_rcp(nlo, nhi, sub_base + cj[k], K);
                }
            }
            const int kl = (cnt - 1) >> 5;   // warp-uniform: pick the register, then one shuffle
            float lastr = rn[0];
#pragma unroll
            for (int k = 1; k < EPL; ++k) lastr = (k == kl) ? rn[k] : lastr;
            lastr = __shfl_sync(DACO_FULL, lastr, (cnt - 1) & 31);

            const float thr = __fmul_rn(__uint_as_float(topbits), kGap);
            const bool is_top = mybits == topbits;
            const uint32_t tops = __ballot_sync(DACO_FULL, is_top);
            const uint32_t nears = __ballot_sync(DACO_FULL, (second >= thr) || (bestA >= thr && !is_top));
            uint32_t jstar, sstar;
            float pn_win = 0.f;
            const bool need_exact = (nears != 0u) || (__popc(tops) != 1);
            if (need_exact || (WANT_LOGP && CVRP)) {
                // bitmap of this step's admissible nodes
                alive[lane] = 0u;
                __syncwarp();
#pragma unroll
                for (int k = 0; k < EPL; ++k)
                    if (okk[k]) atomicOr(&alive[cj[k] >> 5], 1u << (cj[k] & 31));
                __syncwarp();
            }
            if (need_exact) {
                jstar = exact_step(Psm + (size_t)cur * n, alive, n, p.lbw, p.vec, p.double_norm, nz, seed, off_step,
                                   sub_base, p.g_noise, &pn_win);
                uint32_t found = 0;
#pragma unroll
                for (int k = 0; k < EPL; ++k)
                    if (lane + 32 * k < cnt && cj[k] == jstar) found = (uint32_t)(lane + 32 * k) + 1;
                sstar = __reduce_max_sync(DACO_FULL, found);
                // an all-masked argmax can land on a visited node (reference quirk): leave the list untouched
                sstar = sstar ? sstar - 1 : 0xffffffffu;
            } else {
                const int winner = __ffs(tops) - 1;
                const uint32_t sj = __shfl_sync(DACO_FULL, bestsj, winner);
                jstar = sj & 0xffffu;
                sstar = sj >> 16;
                if (WANT_LOGP) {
                    const float* row = Psm + (size_t)cur * n;
                    auto xval = [&](int k) -> float { return ((alive[k >> 5] >> (k & 31)) & 1u) ? row[k] : 0.f; };
                    const int shift = p.vec ? (int)(sub_base & 3u) : 0;
                    const float S = aten_row_sum_fn(xval, n, p.lbw, p.vec != 0, lane, shift);
                    float pn = __fdiv_rn(__shfl_sync(DACO_FULL, bestx, winner), S);
                    if (p.double_norm) {
                        const float S2 = aten_row_sum_fn([&](int k) { return __fdiv_rn(xval(k), S); }, n, p.lbw,
                                                         p.vec != 0, lane, shift);
                        pn = __fdiv_rn(pn, S2);
                    }
                    pn_win = pn;
                }
            }
            if (WANT_LOGP && lane == 0)   // Categorical.log_prob: log(clamp(probs, eps, 1 - eps))[action]
                p.logp[((size_t)b * (R - 1) + step) * p.A + a] = logf(fminf(fmaxf(pn_win, eps), 1.0f - eps));

            // ---- state update
            bool remove = sstar < (uint32_t)cnt;
            if (CVRP) {
                if (jstar == 0u) {
                    used = __fadd_rn(0.f, lds_f32(dem_addr));
                    remove = false;
                } else {
                    used = __fadd_rn(used, lds_f32(dem_addr + 4 * jstar));
                }
            }
            if (remove) {
#pragma unroll
                for (int k = 0; k < EPL; ++k)
                    if ((uint32_t)(lane + 32 * k) == sstar) {
                        cj[k] = lastj;
                        rn[k] = lastr;
                    }
                if (lane == 0) sts_u16(cand_addr + 2 * sstar, lastj);
                --cnt;
                if (WANT_LOGP && !CVRP && lane == 0) alive[jstar >> 5] &= ~(1u << (jstar & 31));
            }
#pragma unroll
            for (int k = 0; k < EPL; ++k) r[k] = rn[k];
            ++step;
            if (lane == 0) sts_u16(tour_addr + 2 * step, jstar);
            __syncwarp();
            cur = (int)jstar;
        }
        if (CVRP) {
            // finished ants sit at the depot; the reference keeps drawing node 0 with probability exactly 1
            for (int k = step + 1 + lane; k < R; k += 32) sts_u16(tour_addr + 2 * k, 0u);
            if (WANT_LOGP) {
                const float lp1 = logf(1.0f - eps);
                for (int k = step + lane; k < R - 1; k += 32) p.logp[((size_t)b * (R - 1) + k) * p.A + a] = lp1;
            }
            if (lane == 0) {
                p.lens[(size_t)b * p.A + a] = step;
                atomicMax(&p.tmax[b], step);
            }
        }
    }
    __syncthreads();

    // ---- cooperative output: reference layout paths[b][s][a] (int64, step-major) and compact tours
    const int wvalid = min(W, p.A - a0);
    if (p.paths) {
        int64_t* out = p.paths + (size_t)b * R * p.A;
        for (int i = tid; i < R * W; i += nthreads) {
            const int s = i / W, w = i - s * W;
            if (w < wvalid) out[(size_t)s * p.A + a0 + w] = (int64_t)tour_all[(size_t)w * R + s];
        }
    }
    if (p.tours) {
        uint16_t* out = p.tours + ((size_t)b * p.A + a0) * R;
        for (int i = tid; i < R * wvalid; i += nthreads) out[i] = tour_all[i];
    }
    for (int r = 0; r < p.n_peers; ++r) {   // fused exchange (ant sharding): this CTA's tours go to every GPU's buffer
        uint16_t* out = p.peer_tours[r] + ((size_t)b * p.A_total + p.ant_base + a0) * R;
        for (int i = tid; i < R * wvalid; i += nthreads) out[i] = tour_all[i];
    }
}

// ---------------------------------------------------------------------------------------------
// kNN variant for sparse products (DeepACO's learned heuristic: a few edges per row carry almost all the
// mass, the rest sit at the 1e-10 floor).  Lane l evaluates candidate knn[cur][l] only (one Philox per lane
// and step).  Let T = max of the row over the columns NOT in knn[cur] (computed per row at staging).  A
// column outside the list can only win if x_j / q_j > A_top, and q_j >= q_min = 2^-24 for every possible
// noise word, so when T * 2^24 < A_top no unlisted column can win whatever its noise: the listed arg-max IS the
// row's arg-max and the step is done.  Otherwise (about a fifth of the steps on the pretrained TSP-100
// network) the step is evaluated densely over all unvisited nodes; near-ties go to exact_step() as before.
// ---------------------------------------------------------------------------------------------
#ifndef DEEPACO_CPU_EMU
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

#endif  // DEEPACO_CPU_EMU

// Rare tail of the fallback step: a tie or a near-tie among the approximate scores -> exact arithmetic in ATen order.
// (ctr_lo, ctr_hi) = Philox counter of call 0 of this step's draw; li0 = linear element index of the ant's column 0.
static __device__ DACO_NOINLINE uint32_t knn_exact_tail(const ListParams& p, uint32_t row_addr, uint32_t wbase, uint32_t ctr_lo, uint32_t ctr_hi,
                                                       uint32_t li0) {
    const int lane = threadIdx.x & 31;
    uint32_t* alive_scratch = reinterpret_cast<uint32_t*>(__cvta_shared_to_generic(wbase + 256u));
    for (int w = 0; w < 8; ++w) {       // alive bitmap for exact_step from the alive byte map
        const uint32_t bits = __ballot_sync(DACO_FULL, lds_u8(wbase + (uint32_t)(w * 32 + lane)) != 0u);
        if (lane == 0) alive_scratch[w] = bits;
    }
    for (int w = 8 + lane; w < 32; w += 32) alive_scratch[w] = 0u;
    __syncwarp();
    float pn;
    const uint64_t off_step = (((uint64_t)ctr_hi << 32) | ctr_lo) << 2;
    return exact_step(reinterpret_cast<const float*>(__cvta_shared_to_generic(row_addr)), alive_scratch, p.n, p.lbw, p.vec, p.double_norm, nullptr,
                      p.seed, off_step, li0, p.g_noise, &pn);
}

// Fallback step of the kNN kernel: evaluate every unvisited column of row `cur` (row_addr = shared address of that
// row of P), commit the winner (alive byte, tour slot at shared address `slot`) and return it.  Everything is passed by value / as
// 32-bit shared addresses so that the call marshals few registers.  GEN: any torch draw geometry (element li = q0 *
// threads + r0 + j); otherwise the single-launch geometry (q0 = 0, every element is component .x of call 0).
template <bool GEN>
static __device__ DACO_NOINLINE uint32_t knn_dense_step(const ListParams& p, uint32_t row_addr, uint32_t wbase, uint32_t slot, uint32_t n,
                                                       uint32_t ctr_lo, uint32_t ctr_hi, uint32_t q0, uint32_t r0) {
    const uint32_t lane = threadIdx.x & 31u;
    const PhiloxRoundKeys& K = p.keys;   // the compiler clones this function for its kernel: constant-bank operands
    // Compact the unvisited columns first (ids = the not-yet-written tail of this ant's tour buffer, one slot per
    // unvisited node by construction): the Philox work shrinks from ceil(n/32) rounds to ceil(alive/32).
    // Lane l looks at columns 128r + 4l .. 4l+3 (one 32-bit load of the alive bytes: 0xff = unvisited; bytes >= n are 0);
    // the order of the compacted list is irrelevant -- any tie goes to the exact path.
    const uint32_t ids_addr = slot;
    uint32_t cnt = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t r = 0; r * 128u < n; ++r) {
        const uint32_t v4 = lds_u32(wbase + (r * 32u + lane) * 4u);
#pragma unroll
        for (uint32_t s = 0; s < 4; ++s) {
            const bool alive = ((v4 >> (8 * s)) & 0xffu) != 0u;
            const uint32_t bits = __ballot_sync(DACO_FULL, alive);
            if (alive) sts_u16(ids_addr + 2u * (cnt + __popc(bits & lt)), r * 128u + 4u * lane + s);
            cnt += __popc(bits);
        }
    }
    __syncwarp();
    float bestA = 0.f, second = 0.f;
    uint32_t bestj = 0xffffffffu;
    auto score = [&](uint32_t i, uint32_t& j) -> float {
        const bool on = i < cnt;
        j = on ? lds_u16(ids_addr + 2u * i) : 0u;
        const float rq = GEN ? noise_rcp_general(ctr_lo, ctr_hi, q0, r0, j, p.g_noise.threads, K) : noise_rcp(ctr_lo, ctr_hi, r0 + j, K);
        const float A = __fmul_rn(lds_f32(row_addr + 4u * j), rq);
        return on ? A : 0.f;
    };
    auto keep = [&](float A, uint32_t j) {
        if (A > bestA) {
            second = bestA;
            bestA = A;
            bestj = j;
        } else if (A > second) {
            second = A;
        }
    };
    uint32_t base = 0;
#pragma unroll 1
    for (; base + 32u < cnt; base += 64u) {   // two independent Philox chains per lane in flight (the step is latency bound)
        uint32_t ja, jb;
        const float Aa = score(base + lane, ja);
        const float Ab = score(base + 32u + lane, jb);
        keep(Aa, ja);
        keep(Ab, jb);
    }
    if (base < cnt) {                         // warp-uniform remainder round
        uint32_t ja;
        const float Aa = score(base + lane, ja);
        keep(Aa, ja);
    }
    const uint32_t mybits = __float_as_uint(bestA);
    const uint32_t topbits = __reduce_max_sync(DACO_FULL, mybits);
    const float thr = __fmul_rn(__uint_as_float(topbits), 1.0f - 3.814697265625e-06f);
    // lanes holding a score within 2^-18 of the top (the top lane included), or a runner-up that close
    const uint32_t close = __ballot_sync(DACO_FULL, bestA >= thr);
    const uint32_t nears = __ballot_sync(DACO_FULL, second >= thr);
    uint32_t jstar;
    if (nears == 0u && __popc(close) == 1) jstar = __shfl_sync(DACO_FULL, bestj, 31 - __clz(close));
    else jstar = knn_exact_tail(p, row_addr, wbase, ctr_lo, ctr_hi, GEN ? q0 * p.g_noise.threads + r0 : r0);
    __syncwarp();
    DACO_STS_U8(wbase + jstar, 0u);
    sts_u16(slot, jstar);
    __syncwarp();
    return jstar;
}

// Fast steps of one tour: lane l evaluates column knn[cur][l] only.  Runs until the tour is complete or a step needs
// the dense / exact treatment (returns with `slot` at that step).  COMP = Philox output word of this ant's elements.
// COMP = 4: the ant's elements straddle two Philox blocks -- word / call chosen per column at run time (q0, threads).
template <int COMP>
__device__ __forceinline__ void knn_fast_steps(uint32_t& slot, uint32_t& ctr_lo, int& cur, const uint32_t slot_end, const uint32_t ctr_step,
                                               const uint32_t ctr_hi, const uint32_t sub0, const uint32_t knn_lane, const uint32_t wbase,
                                               const uint32_t T_addr, const uint32_t P_addr, const uint32_t n, const PhiloxRoundKeys& K,
                                               const uint32_t q0 = 0u, const uint32_t threads = 0u) {
#pragma unroll 1
    for (; slot < slot_end; slot += 2u, ctr_lo += ctr_step) {
        const uint32_t j = lds_u8(knn_lane + (uint32_t)cur * 32u);
        const uint32_t alive = lds_s8(wbase + j);     // sign-extended: all ones while column j is unvisited
        const float T = lds_f32(T_addr + 4u * (uint32_t)cur);
        const float x = __uint_as_float(__float_as_uint(lds_f32(P_addr + ((uint32_t)cur * n + j) * 4u)) & alive);
        const float A = __fmul_rn(x, COMP == 4 ? noise_rcp_general(ctr_lo, ctr_hi, q0, sub0, j, threads, K)
                                               : noise_rcp_comp<(COMP & 3)>(ctr_lo, ctr_hi, sub0 + j, K));
        const uint32_t mybits = __float_as_uint(A);
        const uint32_t topbits = __reduce_max_sync(DACO_FULL, mybits);
        const float top = __uint_as_float(topbits);
        // lanes within 2^-18 (relative) of the top score, the top lane included: the step is decided here
        // only when that is exactly one lane and no unlisted column can beat it
        const uint32_t close = __ballot_sync(DACO_FULL, A >= __fmul_rn(top, 1.0f - 3.814697265625e-06f));
        if (!(__popc(close) == 1 && T < top)) break;
        const uint32_t jstar = __shfl_sync(DACO_FULL, j, 31 - __clz(close));
        // Every lane stores the same two values to the same addresses: one wavefront, no predicate to
        // maintain, and each lane later reads what it wrote itself, so no warp-level fence between steps.
        DACO_STS_U8(wbase + jstar, 0u);
        sts_u16(slot, jstar);
        cur = (int)jstar;
    }
}

// TSP only, 32 < n <= 256, no log-probs, Philox noise, compact tours out.
// Shared layout (fixed offsets keep the address arithmetic of the step loop in two registers):
//   per warp w (kKnnWarpBytes each, kKnnMaxWarps slots): visited bytes [256] | scratch u32 [32] | tour u16 [256]
//   then: bound f32 [256] | knn u8 [n][32] | P f32 [n][n]
constexpr int kKnnWarpBytes = 256 + 128 + 512;
constexpr int kKnnBoundBytes = 1024;

// GEN = false: single-launch draw geometry (n_ants * n <= torch's grid * 256 threads: every element is word .x of call 0).
// GEN = true : any geometry (big colonies): an ant's elements share one (call, word) pair unless its n consecutive
//              elements straddle a multiple of `threads` -- those few ants take every step through the fallback.
template <bool FUSE_COST, int MAXW, bool GEN = false>
static __global__ void __launch_bounds__(MAXW * 32, 1024 / (MAXW * 32)) aco_knn_kernel(const __grid_constant__ ListParams p) {
    constexpr int kKnnFixed = kKnnWarpBytes * MAXW;
    DACO_DYN_SMEM128(smem);
    __shared__ uint64_t bar;
    const int n = p.n;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int W = nthreads >> 5, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    uint8_t* wblk = smem + warp * kKnnWarpBytes;
    uint8_t* vis = wblk;
    uint32_t* scratch = reinterpret_cast<uint32_t*>(wblk + 256);
    uint16_t* tour_sm = reinterpret_cast<uint16_t*>(wblk + 384);
    float* Tsm = reinterpret_cast<float*>(smem + kKnnFixed);                 // fixed offsets: immediates in the step loop
    uint8_t* knn_sm = smem + kKnnFixed + kKnnBoundBytes;
    float* Psm = reinterpret_cast<float*>(knn_sm + (size_t)n * 32);

    {   // candidate lists of this colony (static per instance)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.knn + (size_t)b * n * 32);
        uint32_t* dst = reinterpret_cast<uint32_t*>(knn_sm);
        for (int i = tid; i < n * 8; i += nthreads) dst[i] = __ldg(src + i);
    }
    stage_product(Psm, p.ph, p.heu, n, b, &bar);   // ends with __syncthreads()
    // bound[u] = (max of row u over the columns outside knn[u]) / q_min, with a margin for the approximate scores
    for (int u = warp; u < n; u += W) {
        const uint32_t jj = knn_sm[u * 32 + lane];
        float m = 0.f;
        for (int k = 0; k * 32 < n; ++k) {
            const uint32_t listed = __reduce_or_sync(DACO_FULL, (jj >> 5) == (uint32_t)k ? (1u << (jj & 31)) : 0u);
            const int j = lane + 32 * k;
            if (j < n && !((listed >> lane) & 1u)) m = fmaxf(m, Psm[(size_t)u * n + j]);
        }
        const uint32_t tb = __reduce_max_sync(DACO_FULL, __float_as_uint(m));
        if (lane == 0) Tsm[u] = __fmul_rn(__uint_as_float(tb), 16777216.0f * 1.0001f);
    }
    __syncthreads();

    const int a = blockIdx.x * W + warp;
    if (a >= p.A) return;
    const uint32_t wbase = pin_u32(smem_u32(wblk));                       // vis at +0, tour at +384
    const uint32_t knn_lane = smem_u32(knn_sm) + (uint32_t)lane;
    const uint32_t T_addr = smem_u32(Tsm);
    const uint32_t P_addr = pin_u32(smem_u32(Psm));
    const uint64_t offset0 = (p.offsets ? p.offsets[b] : 0ull) + p.offset;
    const PhiloxRoundKeys& K = p.keys;
    const uint32_t sub_base = (uint32_t)(a + p.ant_base) * (uint32_t)n;

    int cur;
    uint64_t ctr = offset0 >> 2;                                          // Philox counter of the first noise draw (call 0)
    if (p.start_node >= 0) {
        cur = p.start_node;
    } else {
        cur = (int)(torch_philox_word(p.seed, offset0, (uint64_t)(a + p.ant_base), p.g_start) % (uint32_t)n);
        ctr += p.start_increment >> 2;
    }
    // every draw advances the generator offset by step_increment (4 for the single-launch geometry, host-checked)
    const uint32_t ctr_step = GEN ? (p.step_increment >> 2) : 1u;
    uint32_t q0 = 0u, r0 = sub_base;                                      // element li = q0 * threads + r0 + j
    bool straddle = false;
    if (GEN) {
        q0 = sub_base / p.g_noise.threads;
        r0 = sub_base - q0 * p.g_noise.threads;
        straddle = r0 + (uint32_t)n > p.g_noise.threads;                  // this ant's row straddles two Philox blocks
    }
    for (int k = lane; k < 64; k += 32) {          // alive bytes: 0xff = unvisited, 0 = visited or column >= n
        const int j0 = 4 * k;
        uint32_t v = 0u;
#pragma unroll
        for (int s = 0; s < 4; ++s) v |= (j0 + s < n) ? (0xffu << (8 * s)) : 0u;
        reinterpret_cast<uint32_t*>(vis)[k] = v;
    }
    __syncwarp();
    if (lane == 0) {
        vis[cur] = 0;
        tour_sm[0] = (uint16_t)cur;
    }
    __syncwarp();

    // Fast steps run in an inner loop that contains no call, so its registers are not constrained by the calling
    // convention; a step that needs the dense / exact treatment leaves it, is handled, and the fast loop resumes.
    // The loop state is the shared address of the tour slot being filled (no step counter, no per-step address sum)
    // and the low word of the Philox counter.  Its high word is loop-invariant unless the low word wraps during this
    // tour: such a warp (one in ~2^32/n) takes every step through the fallback, which gets the carried high word.
    uint32_t slot = wbase + 384u + 2u;
    const uint32_t slot_end = wbase + 384u + 2u * (uint32_t)n;
    // fast loop: counter of this ant's own call (of call 0 for a straddling ant, which picks call / word per column)
    const uint64_t ctr_fast = straddle ? ctr : ctr + (q0 >> 2);
    const uint32_t ctr_lo0 = (uint32_t)ctr_fast, ctr_hi0 = (uint32_t)(ctr_fast >> 32);
    const bool slow = ctr_lo0 > 0xffffffffu - ((uint32_t)n * ctr_step + 8u + (q0 >> 2));
    uint32_t ctr_lo = ctr_lo0;
#pragma unroll 1
    while (slot < slot_end) {
        if (!slow) {
            if (GEN && straddle)
                knn_fast_steps<4>(slot, ctr_lo, cur, slot_end, ctr_step, ctr_hi0, r0, knn_lane, wbase, T_addr, P_addr, (uint32_t)n, K, q0,
                                  p.g_noise.threads);
            else if (!GEN || (q0 & 3u) == 0u)
                knn_fast_steps<0>(slot, ctr_lo, cur, slot_end, ctr_step, ctr_hi0, r0, knn_lane, wbase, T_addr, P_addr, (uint32_t)n, K);
            else if ((q0 & 3u) == 1u)
                knn_fast_steps<1>(slot, ctr_lo, cur, slot_end, ctr_step, ctr_hi0, r0, knn_lane, wbase, T_addr, P_addr, (uint32_t)n, K);
            else if ((q0 & 3u) == 2u)
                knn_fast_steps<2>(slot, ctr_lo, cur, slot_end, ctr_step, ctr_hi0, r0, knn_lane, wbase, T_addr, P_addr, (uint32_t)n, K);
            else
                knn_fast_steps<3>(slot, ctr_lo, cur, slot_end, ctr_step, ctr_hi0, r0, knn_lane, wbase, T_addr, P_addr, (uint32_t)n, K);
        }
        if (slot < slot_end) {
            // counter of call 0 of this step's draw (the fallback derives call / word per column itself)
            const uint64_t c = ctr + (uint64_t)((slot - (wbase + 384u + 2u)) >> 1) * ctr_step;
            cur = (int)knn_dense_step<GEN>(p, P_addr + (uint32_t)cur * (uint32_t)n * 4u, wbase, slot, (uint32_t)n, (uint32_t)c,
                                           (uint32_t)(c >> 32), q0, r0);
            slot += 2u;
            ctr_lo += ctr_step;
        }
    }
    __syncwarp();
    if (p.tours) {   // warp-local, coalesced: this ant's row of the compact layout
        uint16_t* out = p.tours + ((size_t)b * p.A + a) * n;
        for (int k = lane; k < n; k += 32) out[k] = tour_sm[k];
    }
    for (int r = 0; r < p.n_peers; ++r) {   // fused exchange: the finished tour goes straight to every GPU's buffer
        uint16_t* out = p.peer_tours[r] + ((size_t)b * p.A_total + p.ant_base + a) * n;   // while other ants still build
        for (int k = lane; k < n; k += 32) out[k] = tour_sm[k];
    }
    if (FUSE_COST) {   // fused ACO.gen_path_costs + neighbour table (same arithmetic as tsp_cost_kernel)
        const float* D = p.dist + (size_t)b * n * n;
        auto edge = [&](int k) -> float { return __ldg(D + (size_t)tour_sm[k] * n + tour_sm[k == 0 ? n - 1 : k - 1]); };
        const float c = aten_row_sum_fn(edge, n, p.lbw, p.vec != 0, lane, p.vec ? (int)(((unsigned)a * (unsigned)n) & 3u) : 0);
        if (lane == 0) p.costs[(size_t)b * p.A + a] = c;
        if (p.nbr) {   // neighbour table only for the row-parallel update (the ant-sequential one reads the tours)
            uint32_t* N = p.nbr + (size_t)b * n * p.A;
            for (int k = lane; k < n; k += 32) {
                const uint32_t u = tour_sm[k], pr = tour_sm[k == 0 ? n - 1 : k - 1], su = tour_sm[k == n - 1 ? 0 : k + 1];
                N[(size_t)u * p.A + a] = (pr << 16) | su;
            }
        }
    }
}

inline size_t knn_kernel_smem(int n, int W) {
    return (size_t)kKnnWarpBytes * (W <= 8 ? 8 : W <= 16 ? 16 : 32) + kKnnBoundBytes + (size_t)n * 32 + ((((size_t)n * n * 4) + 15) & ~(size_t)15);
}

inline size_t list_kernel_smem(int n, int rows, int W, bool cvrp, bool global_p = false) {
    const size_t pbytes = global_p ? 0 : ((((size_t)n * n * 4) + 15) & ~(size_t)15);
    const size_t dbytes = cvrp ? ((((size_t)n * 4) + 15) & ~(size_t)15) : 0;
    return pbytes + dbytes + (((size_t)W * (rows + n) * 2 + 15) & ~(size_t)15) + (size_t)W * 128;
}

}  // namespace deepaco
