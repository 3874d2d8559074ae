"""Developer timing probe (not the contract bench): CUDA-event time of the K1/K2 kernels at C2-like sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepaco_b200 import _engine as E

dev = "cuda"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


for (n, A, B, nls) in [(100, 512, 1, False), (100, 512, 8, False), (100, 512, 64, False), (100, 512, 256, False),
                       (200, 256, 1, False), (200, 256, 64, False), (500, 256, 1, True), (20, 8, 1, False)]:
    torch.manual_seed(0)
    coords = torch.rand(B, n, 2, device=dev)
    dist = torch.cdist(coords, coords)
    dist[:, torch.arange(n), torch.arange(n)] = 1e9
    k = max(2, n // 5)
    _, idx = torch.topk(dist, k, dim=2, largest=False)
    heu = torch.full_like(dist, 1e-10)
    heu.scatter_(2, idx, torch.rand(B, n, k, device=dev) * 0.9 + 0.05)
    ph = torch.ones_like(dist)
    offsets = torch.tensor([4000 * b for b in range(B)], dtype=torch.int64, device=dev)
    out = {}

    def samp():
        out["p"], _, out["t"] = E.tsp_sample(ph, heu, A, seed=5, offsets=offsets, start_node=0 if nls else -1, double_norm=nls,
                                             want_paths=True, want_tours=True)
    t_s = timeit(samp)

    def cost():
        out["c"], out["nb"] = E.tsp_cost(dist, tours=out["t"], want_neighbours=True)
    t_c = timeit(cost)

    def upd():
        E.tsp_update_(ph, out["nb"], out["c"], decay=0.9)
    t_u = timeit(upd)
    tours = A * B
    print(f"n={n} A={A} B={B}: sample {t_s:9.1f} us  cost {t_c:7.1f} us  update {t_u:7.1f} us  -> "
          f"{tours / (t_s + t_c + t_u):8.2f} M tours/s (sample-only {tours / t_s:8.2f} M/s, "
          f"{tours / t_s * 1e6 * (2 * 4 * n * (n - 1) + 8 * n) / 1e9:8.1f} GB/s alg)", flush=True)
