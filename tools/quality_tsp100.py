"""End-to-end drop-in check against numbers the reference itself publishes (tsp/train.ipynb cell 7 output, epoch 4,
TSP100, 20 ants, k_sparse = 20, T = 5 ACO iterations, 100 validation instances):

        (avg sample cost, best sample cost, best ACO cost) = (9.711, 9.027, 8.639)

This script is the reference's `validation` loop (train.ipynb cell 1-2) with `ACO` / `Net` imported from
deepaco_b200, the pretrained tsp100 checkpoint and the reference's validation set (fixtures under tests/golden).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepaco_b200.net import load_npz_state_dict
from deepaco_b200.tsp.aco import ACO
from deepaco_b200.tsp.net import Net
from deepaco_b200.tsp.utils import gen_pyg_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = "cuda:0"
EPS = 1e-10
torch.manual_seed(1234)                                    # tsp/train.ipynb cell 0
coords = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/val_tsp100_coords.npz"))["coords"]).to(dev)
net = Net().to(dev)
net.load_state_dict(load_npz_state_dict(os.path.join(ROOT, "deepaco_b200/data/weights_tsp100.npz"), dev))
net.eval()
tot = np.zeros(3)
for inst in coords:
    pyg, distances = gen_pyg_data(inst, k_sparse=20)
    with torch.no_grad():
        heu_mat = net.reshape(pyg, net(pyg)) + EPS
    aco = ACO(n_ants=20, heuristic=heu_mat, distances=distances, device=dev)
    costs, _ = aco.sample()
    aco.run(n_iterations=5)
    tot += np.array([costs.mean().item(), costs.min().item(), float(aco.lowest_cost)])
print({"instances": len(coords), "avg_sample_cost": tot[0] / len(coords), "best_sample_cost": tot[1] / len(coords),
       "best_aco_cost_T5": tot[2] / len(coords), "reference_published": [9.711, 9.027, 8.639]})
