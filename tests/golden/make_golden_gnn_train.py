"""Known answers for the TRAINING-mode heuristic network, produced by the UNMODIFIED reference Net
(/root/reference/{tsp,cvrp}/net.py with tests/golden/pyg_shim) on CPU: train-mode output, the gradient of
L = sum(c * heu_vec) w.r.t. every parameter, and the BatchNorm running statistics after that one forward.

    python tests/golden/make_golden_gnn_train.py        (build container only)

Weights: the pretrained checkpoints already stored as weights_tsp100.npz / weights_cvrp100.npz.
"""
import os

import numpy as np
import torch

from make_golden import REF, load_ref, save


def run(model, pyg, seed):
    model.train()
    torch.manual_seed(seed)
    c = torch.rand(pyg.edge_index.shape[1]) - 0.3
    heu = model(pyg)
    (heu * c).sum().backward()
    out = {"c": c, "heu_vec": heu, "edge_index": pyg.edge_index, "edge_attr": pyg.edge_attr, "x": pyg.x}
    for name, p in model.named_parameters():
        if p.grad is not None:
            out["grad__" + name.replace(".", "__")] = p.grad
    for name, buf in model.named_buffers():
        out["buf__" + name.replace(".", "__")] = buf
    return out


def main():
    torch.set_num_threads(4)
    utils, net_mod = load_ref("tsp", "utils"), load_ref("tsp", "net")
    torch.manual_seed(4242)
    coords = torch.rand(40, 2)
    pyg, _ = utils.gen_pyg_data(coords, k_sparse=8)
    model = net_mod.Net()
    print(model.load_state_dict(torch.load(os.path.join(REF, "pretrained/tsp/tsp100.pt"), map_location="cpu")))
    save("tsp_n40_gnn_train_grads", coords=coords, **run(model, pyg, 7))

    utils, net_mod = load_ref("cvrp", "utils"), load_ref("cvrp", "net")
    torch.manual_seed(2468)
    demand, dist = utils.gen_instance(14, "cpu")
    pyg = utils.gen_pyg_data(demand, dist, "cpu")
    model = net_mod.Net()
    print(model.load_state_dict(torch.load(os.path.join(REF, "pretrained/cvrp/cvrp100.pt"), map_location="cpu")))
    save("cvrp_n14_gnn_train_grads", demand=demand, dist=dist, **run(model, pyg, 8))


if __name__ == "__main__":
    main()
