"""Heuristic matrices for benchmark / smoke / driver instances (set-up code, outside the timed hot path), and the
location of the converted reference checkpoints shipped as package data (`deepaco_b200/data/weights_*.npz`, produced
from `pretrained/{tsp,tsp_nls,cvrp}/*.pt` by tests/golden/make_golden.py)."""
from __future__ import annotations

import os

import torch

from ._lib import DeepAcoError

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def weights_path(name: str) -> str:
    """Path of a packaged checkpoint (`weights_tsp100`, `weights_tsp_nls500`, `weights_cvrp100`); raises if absent --
    a benchmark must never silently change its workload."""
    path = os.path.join(DATA_DIR, name + ".npz")
    if not os.path.exists(path):
        raise DeepAcoError(f"packaged checkpoint {path} is missing")
    return path


def load_net(kind: str, device):
    """`Net` of problem `kind` ('tsp' | 'tsp_nls' | 'cvrp') with its pretrained checkpoint, eval mode."""
    from .net import load_npz_state_dict
    if kind == "tsp":
        from .tsp.net import Net
        name = "weights_tsp100"
    elif kind == "tsp_nls":
        from .tsp_nls.net import Net
        name = "weights_tsp_nls500"
    elif kind == "cvrp":
        from .cvrp.net import Net
        name = "weights_cvrp100"
    else:
        raise DeepAcoError(f"unknown problem kind {kind!r}")
    net = Net().to(device)
    r = net.load_state_dict(load_npz_state_dict(weights_path(name), device))
    if r.missing_keys or r.unexpected_keys:
        raise DeepAcoError(f"checkpoint {name} does not match Net: {r}")
    return net.eval()


def tsp_heuristic(coords, dist, k_sparse, kind="tsp"):
    """[B, n, n] heuristic for a batch of TSP instances: the DeepACO heuristic network (pretrained TSP-100 checkpoint for
    kind='tsp', TSP-500 NLS checkpoint for kind='tsp_nls') on the k-nearest-neighbour graph, +1e-10 off-graph
    (tsp/test.ipynb cell 1).  Returns (heuristic, description).  Raises when the checkpoint is missing."""
    net = load_net(kind, dist.device)
    feats = coords
    if kind == "tsp_nls":                     # one-hot start-node feature (tsp_nls/utils.py:37-43), start node 0
        feats = torch.zeros((dist.shape[0], dist.shape[1], 1), dtype=torch.float32, device=dist.device)
        feats[:, 0, 0] = 1.0
    with torch.no_grad():
        out = net.heuristic_matrices(feats, dist, k_sparse, 1e-10)
    ck = "tsp100" if kind == "tsp" else "tsp_nls500"
    return out, f"Net(pretrained {ck} weights) on k={k_sparse} graph + 1e-10"
