"""Parity of the CVRP CUDA path (C ABI) against the oracle: reference goldens via external noise,
same-device stream parity, run() parity, and structural validity at benchmark size."""
import numpy as np
import pytest
import torch

from oracle import aco_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.mark.parametrize("name,n_ants", [("cvrp_n20_a16", 16), ("cvrp_n100_a32_gnn", 32)])
def test_external_noise_reproduces_reference_golden(golden, name, n_ants):
    from deepaco_b200 import _engine as E
    g = golden(name)
    dist, demand = T(g["dist"]), T(g["demand"])
    heu = T(g["heuristic"]) if "heuristic" in g else 1 / dist
    N = dist.shape[0]
    torch.manual_seed(12345)
    log = []
    paths_cpu = O.cvrp_gen_path(torch.ones_like(dist), heu, demand, 50, n_ants, noise_log=log)
    assert np.array_equal(paths_cpu.numpy(), g["paths_seed12345"].astype(np.int64))
    Tn = len(log)
    noise = torch.ones((2 * N - 1, n_ants, N))
    noise[:Tn] = torch.stack(log)
    out = E.cvrp_sample(torch.ones_like(dist).to(DEV), heu.to(DEV), demand.to(DEV), 50, n_ants, noise=noise.to(DEV))
    assert int(out["tmax"][0]) == Tn
    assert torch.equal(out["paths"][0, :Tn + 1].cpu(), paths_cpu)
    assert (out["paths"][0, Tn + 1:] == 0).all()
    costs, _ = E.cvrp_cost(dist.to(DEV), paths=out["paths"][0, :Tn + 1])
    assert torch.allclose(costs.cpu(), T(g["costs_seed12345"]), rtol=1e-6)


def _instance(n, seed=123456, gnn_like=False):
    torch.manual_seed(seed)
    loc = torch.rand(n, 2, device=DEV)
    demand = torch.cat((torch.zeros(1, device=DEV), torch.randint(1, 10, (n,), device=DEV).float()))
    allc = torch.cat((torch.tensor([[0.5, 0.5]], device=DEV), loc))
    dist = torch.norm(allc[:, None] - allc, dim=2, p=2)
    dist[torch.arange(n + 1), torch.arange(n + 1)] = 1e-10
    heu = 1 / dist
    if gnn_like:
        heu = torch.rand(n + 1, n + 1, device=DEV) * 0.98 + 1e-10
    return demand, dist, heu


@pytest.mark.parametrize("n,n_ants,gnn_like", [(20, 16, False), (20, 20, True), (50, 64, True), (100, 512, True),
                                               (100, 64, False), (31, 33, True), (200, 64, True), (300, 16, True), (500, 8, False)])
def test_stream_parity_same_device(n, n_ants, gnn_like):
    from deepaco_b200 import _engine as E
    demand, dist, heu = _instance(n, gnn_like=gnn_like)
    N = n + 1
    ph = torch.rand(N, N, device=DEV) + 0.5
    g = torch.cuda.default_generators[0]
    torch.manual_seed(31337)
    ref_paths, ref_logp = O.cvrp_gen_path(ph, heu, demand, 50, n_ants, require_prob=True)
    ref_off = g.get_offset()
    torch.manual_seed(31337)
    seed, off = int(g.initial_seed()), int(g.get_offset())
    out = E.cvrp_sample(ph, heu, demand, 50, n_ants, seed=seed, offset=off, want_logp=True, want_tours=True)
    Tn = int(out["tmax"][0])
    assert Tn + 1 == ref_paths.shape[0]
    assert torch.equal(out["paths"][0, :Tn + 1], ref_paths)
    assert off + Tn * E.cvrp_step_offset_increment(N, n_ants) == ref_off
    _, _, exact = E.aten_sum_plan(N, n_ants)
    if exact:
        assert torch.equal(out["logp"][0, :Tn], ref_logp)
    else:
        assert torch.allclose(out["logp"][0, :Tn], ref_logp, rtol=1e-6, atol=1e-6)
    # cost (device-side T and host T agree) + update
    ref_costs = O.cvrp_path_costs(dist, ref_paths)
    c1, nbr = E.cvrp_cost(dist, paths=ref_paths, want_neighbours=True)
    c2, _ = E.cvrp_cost(dist, tours=out["tours"][0], tmax=out["tmax"])
    _, _, exact_c = E.aten_sum_plan(Tn, n_ants)
    if exact_c:
        assert torch.equal(c1, ref_costs)
    else:
        assert torch.allclose(c1, ref_costs, rtol=1e-6)
    assert torch.equal(c1, c2)
    for elitist in (False, True):
        mine = E.cvrp_update_(ph.clone(), nbr, ref_costs, decay=0.9, elitist=elitist)
        ref = O.cvrp_update_pheromone(ph, ref_paths, ref_costs, 0.9, elitist)
        assert torch.equal(mine, ref)


@pytest.mark.parametrize("kw", [{}, {"elitist": True}])
def test_run_matches_oracle_same_seed(kw):
    from deepaco_b200.cvrp.aco import ACO
    demand, dist, heu = _instance(100, gnn_like=True)
    torch.manual_seed(11)
    aco = ACO(dist, demand, n_ants=64, heuristic=heu, device=DEV, **kw)
    low = aco.run(5)
    torch.manual_seed(11)
    ref = O.CvrpColony(dist, demand, 64, heuristic=heu, **kw)
    ref_low = ref.run(5)
    assert torch.equal(aco.pheromone, ref.pheromone)
    assert float(low) == float(ref_low)
    assert torch.equal(aco.shortest_path, ref.shortest_path)


def test_routes_valid_at_benchmark_size():
    """cvrp_nls/test.py:20-37 style validation: every customer once, every route within capacity."""
    from deepaco_b200 import _engine as E
    B, n, A = 4, 100, 512
    demand, dist, heu = _instance(n, gnn_like=True)
    N = n + 1
    offsets = torch.tensor([4000 * b for b in range(B)], dtype=torch.int64, device=DEV)
    out = E.cvrp_sample(torch.ones(B, N, N, device=DEV), heu.expand(B, N, N).contiguous(),
                        demand.expand(B, N).contiguous(), 50, A, seed=9, offsets=offsets, want_tours=True)
    paths = out["paths"].cpu().numpy()
    dem = demand.cpu().numpy()
    lens = out["lens"].cpu().numpy()
    assert (out["tmax"].cpu().numpy() == lens.max(axis=1)).all()
    for b in range(B):
        for a in range(0, A, 37):
            p = paths[b, :, a]
            L = lens[b, a]
            assert p[0] == 0 and p[L] == 0 and (p[L:] == 0).all()
            cust = p[p != 0]
            assert len(cust) == n and len(set(cust.tolist())) == n
            load = 0.0
            for v in p[:L + 1]:
                load = 0.0 if v == 0 else load + dem[v]
                assert load <= 50
    assert torch.equal(out["tours"].to(torch.int64).transpose(1, 2), out["paths"])
