"""One-GPU probe of the ant-sharded iteration's parts (informational): sampling of A/W ants for W = 1, 2, 4, 8, and the
replayed cost / update over all A ants, for TSP-200 colonies with A = 8192 .. 32768 ants."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from bench_legs import timed, tsp_instances
from deepaco_b200 import _engine as E
from deepaco_b200.heuristics import tsp_heuristic

dev = torch.device("cuda:0")
n = 200
coords, d = tsp_instances(1, n, 8192, dev)
heu, _ = tsp_heuristic(coords, d, 20)
ph = torch.ones_like(d)
knn = E.sparse_candidates(heu)
for A in (8192, 16384, 32768):
    row = [f"A={A}:"]
    for W in (1, 2, 4, 8):
        cnt = A // W
        t = timed(lambda: E.tsp_sample_shard(ph, heu, cnt, 0, A, seed=1, offset=0, knn=knn), 10)
        row.append(f"sample A/{W} {t * 1e3:7.1f} us")
    tours = E.tsp_sample_shard(ph, heu, A, 0, A, seed=1, offset=0, knn=knn)
    out = {}

    def cost():
        out["c"], out["n"] = E.tsp_cost(d, tours=tours, want_neighbours=True)

    t = timed(cost, 10)
    row.append(f"cost {t * 1e3:6.1f} us")
    p2 = ph.clone()
    t = timed(lambda: E.tsp_update_(p2, out["n"], out["c"], decay=0.9), 10)
    row.append(f"update {t * 1e3:6.1f} us")
    r = E.TspRunner(d, heu, ph, A)
    t = timed(lambda: r.run(10, 1), 3) / 10
    row.append(f"run/iter {t * 1e3:7.1f} us")
    print("  ".join(row), flush=True)
