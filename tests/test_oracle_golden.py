"""Pin oracle/aco_torch.py (op-for-op restatement) against golden vectors produced by the unmodified
reference on CPU (tests/golden/make_golden.py).  Bit-exact: same ATen ops, same seed, same device."""
import numpy as np
import torch

from oracle import aco_torch as O


def T(x):
    return torch.from_numpy(np.asarray(x))


def test_tsp_n20_gen_path_costs_update(golden):
    g = golden("tsp_n20_a8")
    dist = T(g["dist"])
    ph, heu = torch.ones_like(dist), 1 / dist
    torch.manual_seed(12345)
    paths = O.tsp_gen_path(ph, heu, 8)
    assert np.array_equal(paths.numpy(), g["paths_seed12345"])
    costs = O.tsp_path_costs(dist, paths)
    assert np.array_equal(costs.numpy(), g["costs_seed12345"])
    ph2 = O.tsp_update_pheromone(ph, paths, costs)
    assert np.array_equal(ph2.numpy(), g["pheromone_after_update"])


def test_tsp_n20_sample_logp(golden):
    g = golden("tsp_n20_a8")
    dist = T(g["dist"])
    torch.manual_seed(777)
    paths, logp = O.tsp_gen_path(torch.ones_like(dist), 1 / dist, 8, require_prob=True)
    assert np.array_equal(logp.numpy(), g["sample_logp_seed777"])
    assert np.array_equal(O.tsp_path_costs(dist, paths).numpy(), g["sample_costs_seed777"])


def test_tsp_n20_explicit_noise_identity(golden):
    """Categorical.sample() == argmax(probs / Exp(1) noise): same tours, same generator consumption."""
    g = golden("tsp_n20_a8")
    dist = T(g["dist"])
    torch.manual_seed(12345)
    log = []
    paths = O.tsp_gen_path(torch.ones_like(dist), 1 / dist, 8, noise_log=log)
    assert np.array_equal(paths.numpy(), g["paths_seed12345"])
    assert len(log) == 19 and log[0].shape == (8, 20)


def test_tsp_n20_run_variants(golden):
    g = golden("tsp_n20_a8")
    dist = T(g["dist"])
    for kw, key in (({}, "run5"), ({"elitist": True}, "run5_elitist"), ({"min_max": True}, "run5_minmax")):
        torch.manual_seed(4321)
        col = O.TspColony(dist, 8, **kw)
        low = col.run(5)
        suffix = "_seed4321" if key == "run5" else ""
        assert np.array_equal(np.asarray(low), g[f"{key}_lowest{suffix}"])
        assert np.array_equal(col.pheromone.numpy(), g[f"{key}_pheromone{suffix}"])
    torch.manual_seed(4321)
    col = O.TspColony(dist, 8)
    col.run(5)
    assert np.array_equal(col.shortest_path.numpy(), g["run5_shortest_seed4321"])


def test_tsp_n100_gnn_heuristic(golden):
    g = golden("tsp_n100_a32_gnn")
    dist, heu = T(g["dist"]), T(g["heuristic"])
    torch.manual_seed(12345)
    paths = O.tsp_gen_path(torch.ones_like(dist), heu, 32)
    assert np.array_equal(paths.numpy(), g["paths_seed12345"].astype(np.int64))
    assert np.array_equal(O.tsp_path_costs(dist, paths).numpy(), g["costs_seed12345"])
    torch.manual_seed(777)
    paths, logp = O.tsp_gen_path(torch.ones_like(dist), heu, 32, require_prob=True)
    assert np.array_equal(logp.numpy(), g["sample_logp_seed777"])
    torch.manual_seed(4321)
    col = O.TspColony(dist, 32, heuristic=heu)
    low = col.run(3)
    assert np.array_equal(np.asarray(low), g["run3_lowest_seed4321"])
    assert np.array_equal(col.pheromone.numpy(), g["run3_pheromone_seed4321"])
    assert np.array_equal(col.shortest_path.numpy(), g["run3_shortest_seed4321"])


def test_tsp_nls_gen_path(golden):
    g = golden("tsp_nls_n200_a16")
    dist, heu = T(g["dist"]), T(g["heuristic"])
    torch.manual_seed(12345)
    paths = O.tsp_gen_path(torch.ones_like(dist), heu, 16, nls_variant=True)
    assert np.array_equal(paths.numpy(), g["paths_seed12345"].astype(np.int64))
    torch.manual_seed(777)
    paths, logp = O.tsp_gen_path(torch.ones_like(dist), heu, 16, require_prob=True, nls_variant=True)
    assert np.array_equal(paths.numpy(), g["sample_paths_seed777"].astype(np.int64))
    assert np.array_equal(logp.numpy(), g["sample_logp_seed777"])


def test_cvrp_n20(golden):
    g = golden("cvrp_n20_a16")
    dist, demand = T(g["dist"]), T(g["demand"])
    ph, heu = torch.ones_like(dist), 1 / dist
    torch.manual_seed(12345)
    paths = O.cvrp_gen_path(ph, heu, demand, 50, 16)
    assert np.array_equal(paths.numpy(), g["paths_seed12345"].astype(np.int64))
    costs = O.cvrp_path_costs(dist, paths)
    assert np.array_equal(costs.numpy(), g["costs_seed12345"])
    assert np.array_equal(O.cvrp_update_pheromone(ph, paths, costs).numpy(), g["pheromone_after_update"])
    torch.manual_seed(777)
    paths, logp = O.cvrp_gen_path(ph, heu, demand, 50, 16, require_prob=True)
    assert np.array_equal(logp.numpy(), g["sample_logp_seed777"])
    for kw, key, sfx in (({}, "run4", "_seed4321"), ({"elitist": True}, "run4_elitist", "")):
        torch.manual_seed(4321)
        col = O.CvrpColony(dist, demand, 16, **kw)
        low = col.run(4)
        assert np.array_equal(np.asarray(low), g[f"{key}_lowest{sfx}"])
        assert np.array_equal(col.pheromone.numpy(), g[f"{key}_pheromone{sfx}"])


def test_cvrp_n100_gnn(golden):
    g = golden("cvrp_n100_a32_gnn")
    dist, demand, heu = T(g["dist"]), T(g["demand"]), T(g["heuristic"])
    torch.manual_seed(12345)
    paths = O.cvrp_gen_path(torch.ones_like(dist), heu, demand, 50, 32)
    assert np.array_equal(paths.numpy(), g["paths_seed12345"].astype(np.int64))
    torch.manual_seed(4321)
    col = O.CvrpColony(dist, demand, 32, heuristic=heu)
    low = col.run(3)
    assert np.array_equal(np.asarray(low), g["run3_lowest_seed4321"])
    assert np.array_equal(col.pheromone.numpy(), g["run3_pheromone_seed4321"])


# ---- 2-opt / NLS oracle (plain C restatement) vs the reference's numba implementation ---------------
def test_two_opt_oracle_matches_reference(golden):
    from oracle import two_opt as T2
    g = golden("two_opt_n60")
    for it in (1, 5, 1000):
        out = T2.batched_two_opt(g["dist"], g["tours"], it)
        assert np.array_equal(out.astype(np.int16), g[f"out_it{it}"])


def test_two_opt_and_nls_oracle_on_gnn_instance(golden):
    from oracle import two_opt as T2
    g = golden("tsp_nls_n200_a16")
    tours = g["paths_seed12345"].T.astype(np.uint16)
    n = tours.shape[1]
    assert np.array_equal(T2.batched_two_opt(g["dist"], tours, n // 4).astype(np.int16), g["two_opt_train"].T)
    assert np.array_equal(T2.batched_two_opt(g["dist"], tours, 10000).astype(np.int16), g["two_opt_inference"].T)
    assert np.array_equal(T2.batched_two_opt(g["heuristic_dist"], tours, 20).astype(np.int16), g["two_opt_heudist_20"])
    assert np.array_equal(T2.nls(g["dist"], g["heuristic_dist"], tours, n // 4).astype(np.int16), g["nls_train"].T)


def test_numpy_pairwise_sum_restatement():
    from oracle import two_opt as T2
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 60, 127, 128, 129, 200, 500, 1000, 1031):
        for _ in range(5):
            row = rng.random(n, dtype=np.float32) * 3
            assert T2.numpy_pairwise_sum(row) == np.sum(row.reshape(1, -1), axis=1)[0]
