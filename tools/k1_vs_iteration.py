"""K1 time as the pheromone evolves (256 x TSP-100 x 512, pretrained heuristic): per-iteration sampling-kernel time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from bench_legs import tsp_instances
from deepaco_b200 import _engine as E
from deepaco_b200.heuristics import tsp_heuristic

dev = torch.device("cuda:0")
B, n, A = 256, 100, 512
coords, d = tsp_instances(B, n, 1234, dev)
heu, _ = tsp_heuristic(coords, d, 20)
r = E.TspRunner(d, heu, torch.ones_like(d), A)
offs = torch.tensor([b * 4_000_000 for b in range(B)], dtype=torch.int64, device=dev)
refresh = len(sys.argv) > 1 and sys.argv[1] == "refresh"
evs = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
out = []
for it in range(80):
    if refresh and it > 0:
        r.knn = E.sparse_candidates(r.product)
    r.run(1, 1234, it * r.increment, offs, sample_events=evs)
    torch.cuda.synchronize()
    out.append(evs[0].elapsed_time(evs[1]))
print("refresh" if refresh else "static", " ".join(f"{v:.2f}" for v in out[:12]), "...", " ".join(f"{v:.2f}" for v in out[20::10]))
print("mean best cost", float(r.lowest_cost.mean()))
