// ThreadSanitizer run of the 2-opt / NLS kernel source: every cross-thread shared-memory dependency that is not ordered
// by a barrier, a warp-level sync point or an mbarrier wait is reported as a race.  Test infrastructure only.
//   two_opt_tsan <n> <variant> [nls]
#include "two_opt_emu.cpp"

#include <cstdio>
#include <random>
#include <vector>

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 44, variant = argc > 2 ? atoi(argv[2]) : -1, nls = argc > 3 ? atoi(argv[3]) : 0;
    const int A = 2;
    std::mt19937 rng(3);
    std::uniform_real_distribution<float> uni(0.05f, 1.0f);
    std::vector<float> dist((size_t)n * n), hd((size_t)n * n);
    for (auto& v : dist) v = uni(rng);
    for (auto& v : hd) v = uni(rng) * 3.f;
    for (int i = 0; i < n; ++i) dist[(size_t)i * n + i] = 1e9f;
    std::vector<uint16_t> tours((size_t)A * n);
    for (int a = 0; a < A; ++a) {
        std::vector<uint16_t> p(n);
        for (int i = 0; i < n; ++i) p[i] = (uint16_t)i;
        std::shuffle(p.begin() + 1, p.end(), rng);
        std::copy(p.begin(), p.end(), tours.begin() + (size_t)a * n);
    }
    std::vector<int32_t> passes(A, 0);
    std::vector<float> costs(A, 0.f);
    const char* err = emu_two_opt(dist.data(), nls ? hd.data() : nullptr, tours.data(), n, A, nls, 6, 2, 3, variant, costs.data(), passes.data());
    if (err) { printf("%s\n", err); return 2; }
    long sum = 0;
    for (auto v : tours) sum += v;
    if (sum != (long)A * n * (n - 1) / 2) { printf("not permutations\n"); return 3; }
    printf("tsan run ok: n=%d variant=%d nls=%d passes=%d,%d\n", n, variant, nls, passes[0], passes[1]);
    return 0;
}
