"""K4 parity: GPU 2-opt / NLS vs golden vectors of the reference's numba code and vs the C oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_two_opt(dist, tours_np, it):
    from deepaco_b200 import _engine as E
    t = torch.from_numpy(np.ascontiguousarray(tours_np).astype(np.uint16)).to(DEV)
    E.two_opt_(torch.from_numpy(dist).to(DEV), t, it)
    return t.cpu().numpy().astype(np.int16)


def test_two_opt_matches_reference_golden(golden):
    g = golden("two_opt_n60")
    for it in (1, 5, 1000):
        assert np.array_equal(_run_two_opt(g["dist"], g["tours"], it), g[f"out_it{it}"])


def test_two_opt_and_nls_on_gnn_instance_match_reference_golden(golden):
    from deepaco_b200 import _engine as E
    g = golden("tsp_nls_n200_a16")
    tours = g["paths_seed12345"].T
    n = tours.shape[1]
    assert np.array_equal(_run_two_opt(g["dist"], tours, n // 4), g["two_opt_train"].T)
    assert np.array_equal(_run_two_opt(g["dist"], tours, 10000), g["two_opt_inference"].T)
    assert np.array_equal(_run_two_opt(g["heuristic_dist"], tours, 20), g["two_opt_heudist_20"])
    t = torch.from_numpy(np.ascontiguousarray(tours).astype(np.uint16)).to(DEV)
    E.tsp_nls_(torch.from_numpy(g["dist"]).to(DEV), torch.from_numpy(g["heuristic_dist"]).to(DEV), t, n // 4)
    assert np.array_equal(t.cpu().numpy().astype(np.int16), g["nls_train"].T)


@pytest.mark.parametrize("n,count,it", [(5, 7, 100), (33, 20, 3), (120, 40, 1000), (500, 24, 125), (1000, 4, 30)])
def test_two_opt_matches_c_oracle_on_random_tours(n, count, it):
    from oracle import two_opt as T2
    rng = np.random.default_rng(n)
    xy = rng.random((n, 2), dtype=np.float32)
    dist = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1)).astype(np.float32)
    np.fill_diagonal(dist, 1e9)
    tours = np.stack([np.concatenate(([0], 1 + rng.permutation(n - 1))) for _ in range(count)]).astype(np.uint16)
    ref = T2.batched_two_opt(dist, tours, it)
    assert np.array_equal(_run_two_opt(dist, tours, it), ref.astype(np.int16))


def test_nls_matches_c_oracle_and_class_api():
    from deepaco_b200.tsp_nls.aco import ACO
    from oracle import two_opt as T2
    n, A = 100, 16
    torch.manual_seed(3)
    xy = torch.rand(n, 2, device=DEV)
    dist = torch.norm(xy[:, None] - xy, dim=2, p=2)
    dist[torch.arange(n), torch.arange(n)] = 1e9
    k = 10
    _, idx = torch.topk(dist, k, dim=1, largest=False)
    heu = torch.full_like(dist, 1e-10)
    heu.scatter_(1, idx, torch.rand(n, k, device=DEV) * 0.9 + 0.05)
    aco = ACO(dist, n_ants=A, heuristic=heu, device=DEV, local_search='nls')
    paths = aco.gen_path()
    assert (paths[0] == 0).all()
    out = aco.nls(paths)
    ref = T2.nls(dist.cpu().numpy(), aco.heuristic_dist.cpu().numpy(), paths.T.cpu().numpy(), n // 4)
    assert np.array_equal(out.T.cpu().numpy(), ref.astype(np.int64))
    out2 = aco.two_opt(paths, inference=True)
    ref2 = T2.batched_two_opt(dist.cpu().numpy(), paths.T.cpu().numpy(), 10000)
    assert np.array_equal(out2.T.cpu().numpy(), ref2.astype(np.int64))
    # 2-opt never lengthens a tour and keeps it a permutation starting at node 0
    c0, c1 = aco.gen_path_costs(paths), aco.gen_path_costs(out2)
    assert (c1 <= c0 + 1e-5).all()
    assert torch.equal(torch.sort(out2, dim=0).values, torch.arange(n, device=DEV)[:, None].expand(n, A))
    low = aco.run(2)
    assert isinstance(low, float) and low <= float(c1.min()) * 1.5


@pytest.mark.parametrize("mode", ["2opt", "nls"])
def test_device_side_run_with_local_search_matches_stepwise_composition(mode):
    """ACO.run of tsp_nls (construction -> local search -> cost -> best -> update, all enqueued on the device) equals
    the same iteration composed step by step from the individually verified pieces (kernel construction == reference
    ops, 2-opt / NLS == numba, cost / update == reference ops)."""
    from deepaco_b200 import _engine as E
    from deepaco_b200.tsp_nls.aco import ACO
    n, A = 60, 24
    torch.manual_seed(8)
    xy = torch.rand(n, 2, device=DEV)
    dist = torch.norm(xy[:, None] - xy, dim=2, p=2)
    dist[torch.arange(n), torch.arange(n)] = 1e9
    _, idx = torch.topk(dist, 8, dim=1, largest=False)
    heu = torch.full_like(dist, 1e-10).scatter_(1, idx, torch.rand(n, 8, device=DEV) * 0.9 + 0.05)
    torch.manual_seed(123)
    aco = ACO(dist, n_ants=A, heuristic=heu, device=DEV, local_search=mode)
    low = aco.run(3)
    # step-by-step composition
    torch.manual_seed(123)
    ref = ACO(dist, n_ants=A, heuristic=heu, device=DEV, local_search=mode)
    ph = torch.ones_like(dist)
    best = float("inf")
    for _ in range(3):
        ref.pheromone = ph
        paths = ref.gen_path()
        paths = ref.local_search(paths)
        costs = ref.gen_path_costs(paths)
        best = min(best, float(costs.min()))
        ref.update_pheronome(paths, costs)
        ph = ref.pheromone
    assert torch.equal(aco.pheromone, ph)
    assert low == best


def test_two_opt_and_nls_properties_at_benchmark_size():
    """C3 (TSP-500 x 256 ants, too large for the CPU oracle in test time): size-independent properties -- the result is
    a permutation that still starts at node 0, no tour gets longer, a converged 2-opt is a fixed point, and NLS is at
    least as good as the plain 2-opt it starts with."""
    from deepaco_b200 import _engine as E
    n, A, k = 500, 256, 50
    torch.manual_seed(0)
    xy = torch.rand(n, 2, device=DEV)
    dist = torch.norm(xy[:, None] - xy, dim=2, p=2)
    dist[torch.arange(n), torch.arange(n)] = 1e9
    _, idx = torch.topk(dist, k, dim=1, largest=False)
    heu = torch.full_like(dist, 1e-10).scatter_(1, idx, torch.rand(n, k, device=DEV) * 0.9 + 0.05)
    base = E.tsp_sample(torch.ones_like(dist), heu, A, start_node=0, double_norm=True, seed=3, want_paths=False, want_tours=True)[2]
    hd = (1 / (heu / heu.max(-1, keepdim=True).values + 1e-5)).contiguous()

    def costs(t):
        return E.tsp_cost(dist, tours=t)[0]

    def is_perm(t):
        paths = E.tours_to_paths(t)                                  # int64 [n, A]
        return bool((torch.sort(paths, dim=0).values == torch.arange(n, device=DEV)[:, None]).all()) and bool((paths[0] == 0).all())

    c0 = costs(base)
    t1, passes = E.two_opt_(dist, base.clone(), n // 4, want_passes=True)
    c1 = costs(t1)
    assert is_perm(t1) and bool((c1 <= c0 + 1e-4).all()) and int(passes.max()) <= n // 4
    conv, p_conv = E.two_opt_(dist, base.clone(), 10000, want_passes=True)
    assert is_perm(conv) and int(p_conv.max()) < 10000
    again, p_again = E.two_opt_(dist, conv.clone(), 10000, want_passes=True)
    assert torch.equal(again, conv) and bool((p_again == 1).all())          # one pass that finds nothing
    t2 = E.tsp_nls_(dist, hd, base.clone(), n // 4)
    assert is_perm(t2) and bool((costs(t2) <= c1 + 1e-4).all())


def test_two_opt_on_tours_with_repeated_nodes_follows_the_reference_skip_rule():
    """A tour that is not a permutation can make the reference's skip test (two_opt.py:16) fire; the kernel detects
    such tours and runs them through the band kernel, which evaluates the test: same output as the oracle."""
    from oracle import two_opt as T2
    n, count = 40, 12
    rng = np.random.default_rng(7)
    xy = rng.random((n, 2), dtype=np.float32)
    dist = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1)).astype(np.float32)
    np.fill_diagonal(dist, 1e9)
    tours = np.stack([np.concatenate(([0], 1 + rng.permutation(n - 1))) for _ in range(count)]).astype(np.uint16)
    for a in range(0, count, 2):                     # every other tour: three positions repeat earlier nodes
        pos = rng.choice(np.arange(2, n), size=3, replace=False)
        tours[a, pos] = tours[a, pos - 2]
    ref = T2.batched_two_opt(dist, tours, 15)
    assert np.array_equal(_run_two_opt(dist, tours, 15), ref.astype(np.int16))


def test_tsp_nls_run_with_general_exponents_composes_the_verified_steps():
    """alpha / beta != 1: ACO.run of tsp_nls goes through the per-step methods (powers by torch.pow) and must equal the
    explicit composition of those methods under the same seed."""
    from deepaco_b200.tsp_nls.aco import ACO
    n, A = 40, 16
    torch.manual_seed(4)
    xy = torch.rand(n, 2, device=DEV)
    dist = torch.norm(xy[:, None] - xy, dim=2, p=2)
    dist[torch.arange(n), torch.arange(n)] = 1e9
    heu = torch.rand(n, n, device=DEV) * 0.9 + 0.05
    kw = dict(n_ants=A, heuristic=heu, device=DEV, local_search="2opt", alpha=1.5, beta=2)
    torch.manual_seed(21)
    aco = ACO(dist, **kw)
    low = aco.run(2)
    torch.manual_seed(21)
    ref = ACO(dist, **kw)
    best = float("inf")
    for _ in range(2):
        paths = ref.local_search(ref.gen_path())
        costs = ref.gen_path_costs(paths)
        best = min(best, float(costs.min()))
        ref.update_pheronome(paths, costs)
    assert isinstance(low, float) and low == best
    assert torch.equal(aco.pheromone, ref.pheromone)
