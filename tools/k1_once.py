"""A few K1 launches for ncu: python tools/k1_once.py <shape>   shape: c5shard (8 x TSP-200 x 256) | ant1024 | ant8192 | c2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from bench_legs import tsp_instances
from deepaco_b200 import _engine as E
from deepaco_b200.heuristics import tsp_heuristic

shape = sys.argv[1] if len(sys.argv) > 1 else "c5shard"
dev = torch.device("cuda:0")
B, n, A, cnt = {"c5shard": (8, 200, 256, 256), "c5full": (64, 200, 256, 256), "ant1024": (1, 200, 8192, 1024),
                "ant8192": (1, 200, 8192, 8192), "c2": (256, 100, 512, 512)}[shape]
coords, d = tsp_instances(B, n, 2005, dev)
heu, _ = tsp_heuristic(coords, d, 20)
r = E.TspRunner(d, heu, torch.ones_like(d), A)
r.run(3, 1)                                     # evolve the pheromone a little (product = ph * heu)
knn = r.knn
for i in range(3):
    E.tsp_sample_shard(r.product, None, cnt, 0, A, seed=1, offset=4000 * i, knn=knn)
torch.cuda.synchronize()
