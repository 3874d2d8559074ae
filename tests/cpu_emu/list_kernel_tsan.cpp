// ThreadSanitizer run of the list kernel source (TSP with log-probs, CVRP) on small synthetic instances, Philox noise.
// aco_knn_kernel is not run here: its fast loop deliberately lets all 32 lanes store the same bytes to the same
// addresses (benign on the device, a write-write race to a race detector).  Test infrastructure only.
#include "list_kernel_emu.cpp"

#include <cstdio>
#include <random>
#include <vector>

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 40, A = 8, W = 4;
    std::mt19937 rng(11);
    std::uniform_real_distribution<float> uni(0.05f, 1.0f);
    std::vector<float> ph((size_t)n * n), heu((size_t)n * n), demand(n);
    for (auto& v : ph) v = uni(rng) + 0.5f;
    for (auto& v : heu) v = uni(rng);
    for (int i = 0; i < n; ++i) demand[i] = i == 0 ? 0.f : (float)(1 + (int)(uni(rng) * 8.9f));
    std::vector<int64_t> paths((size_t)2 * n * A, -1);
    std::vector<float> logp((size_t)2 * n * A, 0.f);
    std::vector<uint16_t> tours((size_t)2 * n * A, 0);
    std::vector<int32_t> lens(A, 0), tmax(1, 0);
    int lbw = 0;
    while ((2 << lbw) <= (n < 32 ? n : 32)) ++lbw;   // largest power of two <= min(n, 32) lanes per row
    const char* err = emu_tsp_sample(ph.data(), heu.data(), n, A, 1, -1, 0, 7, 16, nullptr, nullptr, nullptr, paths.data(), logp.data(),
                                     tours.data(), lbw, 0, 1u << 20, 1, 4, 0, W);
    if (err) { printf("tsp: %s\n", err); return 2; }
    long sum = 0;
    for (int i = 0; i < n * A; ++i) sum += paths[i];
    if (sum != (long)A * n * (n - 1) / 2) { printf("tsp: not permutations\n"); return 3; }
    err = emu_cvrp_sample(ph.data(), heu.data(), demand.data(), 20.f, n, A, 1, 7, 16, nullptr, paths.data(), logp.data(), tours.data(),
                          lens.data(), tmax.data(), lbw, 0, 1u << 20, 1, 4, W);
    if (err) { printf("cvrp: %s\n", err); return 2; }
    if (tmax[0] < n - 1 || tmax[0] >= 2 * n) { printf("cvrp: implausible length %d\n", tmax[0]); return 3; }
    printf("tsan run ok: n=%d tsp + cvrp (T=%d)\n", n, tmax[0]);
    return 0;
}
