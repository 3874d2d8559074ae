"""Heuristic matrices for benchmark / smoke instances (set-up code, outside the timed hot path)."""
from __future__ import annotations

import os

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tsp_heuristic(coords, dist, k_sparse):
    """[B, n, n] heuristic for a batch of TSP instances: the DeepACO heuristic network on the k-nearest-neighbour
    graph (+1e-10 off-graph, tsp/test.ipynb cell 1) when its weights are available, else a synthetic matrix with
    the same sparsity structure.  Returns (heuristic, description)."""
    B, n = dist.shape[0], dist.shape[1]
    try:
        from .tsp.net import Net, load_npz_state_dict
        wpath = os.path.join(_ROOT, "tests", "golden", f"weights_tsp{n}.npz")
        if os.path.exists(wpath):
            net = Net().to(dist.device)
            net.load_state_dict(load_npz_state_dict(wpath, dist.device))
            net.eval()
            out = net.heuristic_matrices(coords, dist, k_sparse, 1e-10)
            return out, f"Net(pretrained tsp{n} weights) on k={k_sparse} graph + 1e-10"
    except ImportError:
        pass
    g = torch.Generator(device="cpu").manual_seed(4321)
    _, idx = torch.topk(dist, k_sparse, dim=2, largest=False)
    heu = torch.full_like(dist, 1e-10)
    heu.scatter_(2, idx, (torch.rand((B, n, k_sparse), generator=g) * 0.9 + 0.05).to(dist.device))
    return heu, f"synthetic sigmoid-range values on the k={k_sparse} nearest-neighbour graph, 1e-10 elsewhere"
