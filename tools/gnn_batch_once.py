"""Batched heuristic front end (instance -> kNN graph -> Net -> dense heuristic), 256 x TSP-100 and 64 x TSP-200:
CUDA-event time per batch; with an argument: a few launches only (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from bench_legs import timed, tsp_instances
from deepaco_b200.heuristics import load_net

dev = torch.device("cuda:0")
net = load_net("tsp", dev)
for B, n, k in ((256, 100, 20), (64, 200, 20)):
    coords, d = tsp_instances(B, n, 1, dev)
    if len(sys.argv) > 1:
        for _ in range(3):
            net.heuristic_matrices(coords, d, k)
        torch.cuda.synchronize()
        break
    t = timed(lambda: net.heuristic_matrices(coords, d, k), 10)
    print(f"GNN batched front end {B} x TSP-{n} (k={k}): {t:8.3f} ms -> {t / B * 1e3:8.1f} us / instance", flush=True)
