/* ORACLE (test infrastructure only -- never linked into or called by the product path).
 *
 * Plain-C restatement of the reference's numba 2-opt (tsp_nls/two_opt.py:6-39) and of the NLS composition
 * (tsp_nls/aco.py:241-258), plus numpy's float32 pairwise row sum which the reference uses to compare
 * tours inside NLS (tsp_nls/aco.py:171-182 -> np.sum(..., axis=1)).
 * Pinned against golden vectors produced by the unmodified reference (tests/golden/two_opt_n60.npz,
 * tsp_nls_n200_a16.npz) in tests/test_oracle_golden.py.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* two_opt.py:6-28 -- one pass, in place; returns delta (float32) or 0 */
static float two_opt_once(const float* d, int n, uint16_t* tour) {
    int p = 0, q = 0;
    double delta = 0.0;                       /* numba types `delta` float64, `change` float32 */
    for (int i = 1; i < n - 1; ++i) {
        for (int j = i + 1; j < n; ++j) {
            const int ni = tour[i], nj = tour[j];
            const int np_ = tour[i - 1], nx = tour[(j + 1) % n];
            if (np_ == nj || nx == ni) continue;
            float change = d[(size_t)np_ * n + nj];
            change = change + d[(size_t)ni * n + nx];
            change = change - d[(size_t)np_ * n + ni];
            change = change - d[(size_t)nj * n + nx];
            if ((double)change < delta) { p = i; q = j; delta = (double)change; }
        }
    }
    if (delta < -1e-6) {
        while (p < q) { uint16_t t = tour[p]; tour[p] = tour[q]; tour[q] = t; ++p; --q; }
        return (float)delta;
    }
    return 0.0f;
}

/* two_opt.py:31-39 */
void oracle_two_opt(const float* d, int n, uint16_t* tour, int64_t max_iterations) {
    int64_t it = 0;
    float min_change = -1.0f;
    while ((double)min_change < -1e-6 && it < max_iterations) {
        min_change = two_opt_once(d, n, tour);
        ++it;
    }
}

/* two_opt.py:41-49 (the thread pool is only a scheduler: tours are independent) */
void oracle_batched_two_opt(const float* d, int n, uint16_t* tours, int count, int64_t max_iterations) {
    for (int a = 0; a < count; ++a) oracle_two_opt(d, n, tours + (size_t)a * n, max_iterations);
}

/* numpy FLOAT_pairwise_sum (PW_BLOCKSIZE 128, 8 accumulators) over a contiguous float32 row */
float oracle_numpy_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return oracle_numpy_pairwise_sum(a, n2) + oracle_numpy_pairwise_sum(a + n2, n - n2);
    }
}

/* tsp_nls/aco.py:171-182 for one tour: sum_k d[u_k, u_{k-1}] */
float oracle_tour_cost_numpy(const float* d, int n, const uint16_t* tour) {
    float* e = (float*)malloc(sizeof(float) * (size_t)n);
    for (int k = 0; k < n; ++k) e[k] = d[(size_t)tour[k] * n + tour[k == 0 ? n - 1 : k - 1]];
    const float s = oracle_numpy_pairwise_sum(e, n);
    free(e);
    return s;
}

/* tsp_nls/aco.py:241-258 */
void oracle_nls(const float* d, const float* heu_dist, int n, uint16_t* tours, int count, int64_t maxt, int T_nls, int T_p) {
    uint16_t* cur = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)n);
    for (int a = 0; a < count; ++a) {
        uint16_t* best = tours + (size_t)a * n;
        oracle_two_opt(d, n, best, maxt);
        float best_cost = oracle_tour_cost_numpy(d, n, best);
        memcpy(cur, best, sizeof(uint16_t) * (size_t)n);
        for (int r = 0; r < T_nls; ++r) {
            oracle_two_opt(heu_dist, n, cur, T_p);
            oracle_two_opt(d, n, cur, maxt);
            const float c = oracle_tour_cost_numpy(d, n, cur);
            if (c < best_cost) { memcpy(best, cur, sizeof(uint16_t) * (size_t)n); best_cost = c; }
        }
    }
    free(cur);
}
