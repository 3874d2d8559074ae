"""Generate the golden vectors in this directory by running the UNMODIFIED reference
(/root/reference, henry-yeh/DeepACO @ 9a756a3) on CPU under fixed seeds.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or known-answer vectors of its own, so these files are the
pins for `oracle/` (tests/test_oracle_golden.py) and, through the oracle, for the CUDA kernels.
`torch_geometric` is not installable here; `tests/golden/pyg_shim` supplies the three symbols the
reference's net.py/utils.py import so that those files load unmodified.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "pyg_shim"))


def load_ref(subdir, name):
    """Import /root/reference/<subdir>/<name>.py under a private module name (the reference dirs are
    script bundles with clashing module names)."""
    d = os.path.join(REF, subdir)
    sys.path.insert(0, d)
    try:
        for clash in ("aco", "net", "utils", "two_opt"):
            sys.modules.pop(clash, None)
        spec = importlib.util.spec_from_file_location(f"ref_{subdir}_{name}", os.path.join(d, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path.remove(d)


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def save(name, **arrays):
    # converted checkpoints are package data (deepaco_b200/data); everything else is a test fixture
    where = os.path.join(os.path.dirname(os.path.dirname(HERE)), "deepaco_b200", "data") if name.startswith("weights_") else HERE
    path = os.path.join(where, name + ".npz")
    np.savez_compressed(path, **{k: npy(v) for k, v in arrays.items()})
    print(f"{name}.npz  {os.path.getsize(path)/1024:.1f} KiB")


def pack_state_dict(sd):
    return {k.replace(".", "__"): npy(v) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
def gen_tsp():
    aco_mod = load_ref("tsp", "aco")
    utils = load_ref("tsp", "utils")
    net_mod = load_ref("tsp", "net")

    # C1: n=20, 8 ants, heuristic = 1/dist (BASELINE.json configs[0])
    torch.manual_seed(1234)
    coords = torch.rand(20, 2)
    dist = utils.gen_distance_matrix(coords)
    out = {"coords": coords, "dist": dist}
    torch.manual_seed(12345)
    aco = aco_mod.ACO(dist, n_ants=8)
    paths = aco.gen_path(require_prob=False)
    out["paths_seed12345"] = paths
    out["costs_seed12345"] = aco.gen_path_costs(paths)
    aco.update_pheronome(paths, out["costs_seed12345"])
    out["pheromone_after_update"] = aco.pheromone
    torch.manual_seed(777)
    aco = aco_mod.ACO(dist, n_ants=8)
    costs, logp = aco.sample()
    out["sample_costs_seed777"] = costs
    out["sample_logp_seed777"] = logp
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, n_ants=8)
    out["run5_lowest_seed4321"] = aco.run(5)
    out["run5_pheromone_seed4321"] = aco.pheromone
    out["run5_shortest_seed4321"] = aco.shortest_path
    # variants: elitist, min_max, sparsify
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, n_ants=8, elitist=True)
    out["run5_elitist_lowest"] = aco.run(5)
    out["run5_elitist_pheromone"] = aco.pheromone
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, n_ants=8, min_max=True)
    out["run5_minmax_lowest"] = aco.run(5)
    out["run5_minmax_pheromone"] = aco.pheromone
    aco = aco_mod.ACO(dist, n_ants=8)
    aco.sparsify(5)
    out["sparsify5_heuristic"] = aco.heuristic
    save("tsp_n20_a8", **out)

    # C2-shaped: n=100 with the pretrained GNN heuristic (k_sparse=20), small ant count for file size
    torch.manual_seed(1234)
    coords = torch.rand(100, 2)
    pyg, dist = utils.gen_pyg_data(coords, k_sparse=20)
    model = net_mod.Net()
    sd = torch.load(os.path.join(REF, "pretrained/tsp/tsp100.pt"), map_location="cpu")
    print("tsp100.pt:", model.load_state_dict(sd))
    model.eval()
    with torch.no_grad():
        heu_vec = model(pyg)
        heu = model.reshape(pyg, heu_vec) + 1e-10          # tsp/test.ipynb cell 1 (EPS)
    out = {"coords": coords, "dist": dist, "heu_vec": heu_vec, "heuristic": heu,
           "edge_index": pyg.edge_index, "edge_attr": pyg.edge_attr}
    torch.manual_seed(12345)
    aco = aco_mod.ACO(dist, n_ants=32, heuristic=heu)
    paths = aco.gen_path(require_prob=False)
    out["paths_seed12345"] = paths.to(torch.int16)
    out["costs_seed12345"] = aco.gen_path_costs(paths)
    torch.manual_seed(777)
    aco = aco_mod.ACO(dist, n_ants=32, heuristic=heu)
    costs, logp = aco.sample()
    out["sample_costs_seed777"] = costs
    out["sample_logp_seed777"] = logp
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, n_ants=32, heuristic=heu)
    out["run3_lowest_seed4321"] = aco.run(3)
    out["run3_pheromone_seed4321"] = aco.pheromone
    out["run3_shortest_seed4321"] = aco.shortest_path
    save("tsp_n100_a32_gnn", **out)
    save("weights_tsp100", **pack_state_dict(sd))

    # training-mode BatchNorm forward (train_instance path, tsp/train.ipynb cell 1)
    model.train()
    with torch.no_grad():
        heu_vec_train = model(pyg)
    save("tsp_n100_gnn_trainmode", heu_vec=heu_vec_train)


def gen_tsp_nls():
    aco_mod = load_ref("tsp_nls", "aco")
    utils = load_ref("tsp_nls", "utils")
    net_mod = load_ref("tsp_nls", "net")
    batched_two_opt = aco_mod.batched_two_opt_python   # tsp_nls/two_opt.py:41-49

    torch.manual_seed(1234)
    coords = torch.rand(200, 2)
    pyg, dist = utils.gen_pyg_data(coords, k_sparse=20, start_node=0)
    model = net_mod.Net()
    sd = torch.load(os.path.join(REF, "pretrained/tsp_nls/tsp500.pt"), map_location="cpu")
    print("tsp_nls/tsp500.pt:", model.load_state_dict(sd))
    model.eval()
    with torch.no_grad():
        heu_vec = model(pyg)
        heu = model.reshape(pyg, heu_vec) + 1e-10          # tsp_nls/test.py:20 (EPS)
    out = {"coords": coords, "dist": dist, "heu_vec": heu_vec, "heuristic": heu}
    torch.manual_seed(12345)
    aco = aco_mod.ACO(dist, n_ants=16, heuristic=heu, local_search=None)
    paths = aco.gen_path(require_prob=False)
    out["paths_seed12345"] = paths.to(torch.int16)
    torch.manual_seed(777)
    aco = aco_mod.ACO(dist, n_ants=16, heuristic=heu, local_search=None)
    costs, logp, p2 = aco.sample()
    out["sample_costs_seed777"] = costs
    out["sample_logp_seed777"] = logp
    out["sample_paths_seed777"] = p2.to(torch.int16)
    # 2-opt / NLS on those tours (deterministic, tsp_nls/aco.py:234-258)
    aco = aco_mod.ACO(dist, n_ants=16, heuristic=heu, local_search="nls")
    out["two_opt_train"] = aco.two_opt(paths, inference=False).to(torch.int16)      # max_iterations = n//4
    out["two_opt_inference"] = aco.two_opt(paths, inference=True).to(torch.int16)   # max_iterations = 10000
    out["nls_train"] = aco.nls(paths, inference=False).to(torch.int16)
    out["heuristic_dist"] = aco.heuristic_dist
    out["two_opt_heudist_20"] = batched_two_opt(
        aco.heuristic_dist, paths.T.numpy(), max_iterations=20).astype(np.int16)
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, n_ants=16, heuristic=heu, local_search="nls")
    out["run2_nls_lowest_seed4321"] = np.float64(aco.run(2))
    out["run2_nls_pheromone_seed4321"] = aco.pheromone
    save("tsp_nls_n200_a16", **out)
    save("weights_tsp_nls500", **pack_state_dict(sd))

    # small 2-opt known answers on random tours (no GNN involved)
    rng = np.random.default_rng(7)
    torch.manual_seed(99)
    coords = torch.rand(60, 2)
    dist = utils.gen_distance_matrix(coords).numpy()
    tours = np.stack([np.concatenate(([0], 1 + rng.permutation(59))) for _ in range(12)]).astype(np.uint16)
    o = {"dist": dist, "tours": tours.astype(np.int16)}
    for it in (1, 5, 1000):
        o[f"out_it{it}"] = batched_two_opt(dist, tours, max_iterations=it).astype(np.int16)
    save("two_opt_n60", **o)


def gen_cvrp():
    aco_mod = load_ref("cvrp", "aco")
    utils = load_ref("cvrp", "utils")
    net_mod = load_ref("cvrp", "net")

    # small: n=20 customers, heuristic = 1/dist
    torch.manual_seed(123456)
    demand, dist = utils.gen_instance(20, "cpu")
    out = {"demand": demand, "dist": dist}
    torch.manual_seed(12345)
    aco = aco_mod.ACO(dist, demand, n_ants=16)
    paths = aco.gen_path(require_prob=False)
    out["paths_seed12345"] = paths.to(torch.int16)
    out["costs_seed12345"] = aco.gen_path_costs(paths)
    aco.update_pheronome(paths, out["costs_seed12345"])
    out["pheromone_after_update"] = aco.pheromone
    torch.manual_seed(777)
    aco = aco_mod.ACO(dist, demand, n_ants=16)
    costs, logp = aco.sample()
    out["sample_costs_seed777"] = costs
    out["sample_logp_seed777"] = logp
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, demand, n_ants=16)
    out["run4_lowest_seed4321"] = aco.run(4)
    out["run4_pheromone_seed4321"] = aco.pheromone
    out["run4_shortest_seed4321"] = aco.shortest_path.to(torch.int16)
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, demand, n_ants=16, elitist=True)
    out["run4_elitist_lowest"] = aco.run(4)
    out["run4_elitist_pheromone"] = aco.pheromone
    save("cvrp_n20_a16", **out)

    # C4-shaped: n=100 customers with the pretrained GNN (dense graph)
    torch.manual_seed(123456)
    demand, dist = utils.gen_instance(100, "cpu")
    pyg = utils.gen_pyg_data(demand, dist, "cpu")
    model = net_mod.Net()
    sd = torch.load(os.path.join(REF, "pretrained/cvrp/cvrp100.pt"), map_location="cpu")
    print("cvrp100.pt:", model.load_state_dict(sd))
    model.eval()
    with torch.no_grad():
        heu_vec = model(pyg)
        heu = heu_vec.reshape((101, 101)) + 1e-10          # cvrp/test.py:19-20
    out = {"demand": demand, "dist": dist, "heu_vec": heu_vec, "heuristic": heu}
    torch.manual_seed(12345)
    aco = aco_mod.ACO(dist, demand, n_ants=32, heuristic=heu)
    paths = aco.gen_path(require_prob=False)
    out["paths_seed12345"] = paths.to(torch.int16)
    out["costs_seed12345"] = aco.gen_path_costs(paths)
    torch.manual_seed(4321)
    aco = aco_mod.ACO(dist, demand, n_ants=32, heuristic=heu)
    out["run3_lowest_seed4321"] = aco.run(3)
    out["run3_pheromone_seed4321"] = aco.pheromone
    save("cvrp_n100_a32_gnn", **out)
    save("weights_cvrp100", **pack_state_dict(sd))


if __name__ == "__main__":
    torch.set_num_threads(8)
    gen_tsp()
    gen_tsp_nls()
    gen_cvrp()


def gen_val_fixture():
    """data/tsp/valDataset-100.pt (100 instances, float32 [100, 100, 2]) as a fixture for the quality check of
    tools/quality_tsp100.py against the reference's own published validation numbers (tsp/train.ipynb cell 7)."""
    v = torch.load(os.path.join(REF, "data/tsp/valDataset-100.pt"))
    save("val_tsp100_coords", coords=v.to(torch.float32))


if __name__ == "__main__":
    gen_val_fixture()
