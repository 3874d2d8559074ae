"""A few launches of the many-ants pheromone update for ncu: TSP-200 x 8192 ants (tsp_update_row_kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from bench_legs import tsp_instances
from deepaco_b200 import _engine as E
from deepaco_b200.heuristics import tsp_heuristic

dev = torch.device("cuda:0")
A = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
coords, d = tsp_instances(1, 200, 8192, dev)
heu, _ = tsp_heuristic(coords, d, 20)
r = E.TspRunner(d, heu, torch.ones_like(d), A)
r.run(4, 1)
torch.cuda.synchronize()
nbr = r.neighbours[0]
cnt = torch.bincount((nbr.long() & 0xffff).flatten() + 200 * torch.arange(200, device=dev).repeat_interleave(A), minlength=40000)
print("max successor-cell events per row cell:", int(cnt.max()), "of", A)
