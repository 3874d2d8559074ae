"""Eval-mode golden of the heuristic network at the C3 size (BASELINE.json configs[2]: TSP-NLS n = 500, k = 50 graph,
start node 0, pretrained/tsp_nls/tsp500.pt) from the UNMODIFIED reference net.py + utils.py (with the 3-symbol
torch_geometric shim), on CPU.  Stores coordinates and the edge heuristic as float32 plus a float64 evaluation of the same
network (the arbiter tests/test_gpu_gnn.py uses).

    python tests/golden/make_golden_c3.py      ->  tests/golden/tsp_nls_n500_gnn.npz      (build container only)
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, load_ref  # noqa: E402


def main():
    utils = load_ref("tsp_nls", "utils")
    net_mod = load_ref("tsp_nls", "net")
    torch.manual_seed(500)
    coords = torch.rand(500, 2)
    pyg, dist = utils.gen_pyg_data(coords, k_sparse=50, start_node=0)
    model = net_mod.Net()
    print(model.load_state_dict(torch.load(os.path.join(REF, "pretrained/tsp_nls/tsp500.pt"), map_location="cpu")))
    model.eval()
    with torch.no_grad():
        heu_vec = model(pyg)
        m64 = copy.deepcopy(model).double()
        p64 = copy.copy(pyg)
        p64.x, p64.edge_attr = pyg.x.double(), pyg.edge_attr.double()
        heu64 = m64(p64)
    rel = ((heu_vec.double() - heu64).abs() / heu64.abs()).max()
    print("reference fp32 vs its own fp64 evaluation: max relative error", float(rel), " range", float(heu64.min()), float(heu64.max()))
    np.savez_compressed(os.path.join(HERE, "tsp_nls_n500_gnn.npz"), coords=coords.numpy(), heu_vec=heu_vec.numpy(),
                        heu_vec_fp64=heu64.numpy(), edge_index=pyg.edge_index.numpy().astype(np.int32))


if __name__ == "__main__":
    main()
