"""Edge cases of the boundary on the GPU: smallest / largest sizes, ragged ant counts, error behaviour,
and the ant-sharded multi-rank protocol running the real CUDA backend (two gloo ranks sharing cuda:0)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import aco_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inst(n, seed=0):
    torch.manual_seed(seed)
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    return d, 1 / d


@pytest.mark.parametrize("n,n_ants", [(2, 1), (3, 5), (4, 33), (31, 7), (32, 9), (33, 1), (257, 3), (1024, 8)])
def test_extreme_sizes_match_oracle_stream(n, n_ants):
    from deepaco_b200 import _engine as E
    d, heu = _inst(n)
    ph = torch.rand(n, n, device=DEV) + 0.5
    g = torch.cuda.default_generators[0]
    torch.manual_seed(1)
    ref = O.tsp_gen_path(ph, heu, n_ants)
    torch.manual_seed(1)
    paths, _, tours = E.tsp_sample(ph, heu, n_ants, seed=int(g.initial_seed()), offset=int(g.get_offset()), want_tours=True)
    _, _, exact = E.aten_sum_plan(n, n_ants)
    if exact:
        assert torch.equal(paths, ref)
    else:   # ATen widens the reduction block for < 16 rows: only 1-ulp near-ties could differ
        assert (paths == ref).float().mean() > 0.999
    assert torch.equal(torch.sort(paths, dim=0).values, torch.arange(n, device=DEV)[:, None].expand(n, n_ants))
    costs, nbr = E.tsp_cost(d, tours=tours, want_neighbours=True)
    assert torch.allclose(costs, O.tsp_path_costs(d, paths), rtol=1e-6)
    if n >= 3:
        upd = E.tsp_update_(ph.clone(), nbr, costs)
        assert torch.equal(upd, O.tsp_update_pheromone(ph, paths, costs)) or not exact


def test_error_behaviour():
    from deepaco_b200 import DeepAcoError
    from deepaco_b200 import _engine as E
    from deepaco_b200.tsp.aco import ACO
    d, heu = _inst(10)
    with pytest.raises(DeepAcoError):
        ACO(d.cpu(), n_ants=4)                                  # no CPU path
    with pytest.raises(DeepAcoError):
        E.tsp_sample(torch.ones(1, 1, device=DEV), None, 4)     # n < 2
    with pytest.raises(DeepAcoError):
        E.tsp_sample(torch.ones(2000, 2000, device=DEV), None, 4)   # n > DEEPACO_MAX_NODES
    with pytest.raises(DeepAcoError):
        E.tsp_cost(d, paths=torch.zeros(9, 4, dtype=torch.int64, device=DEV))   # wrong number of rows
    with pytest.raises(DeepAcoError):
        E.two_opt_(d, torch.zeros(4, 10, dtype=torch.int64, device=DEV), 3)     # tours must be uint16


def _rank(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deepaco_b200.dist import AntShardedColony, CudaTspBackend
        torch.cuda.set_device(0)
        torch.manual_seed(0)
        n, A = 60, 50
        xy = torch.rand(n, 2, device=DEV)
        d = torch.norm(xy[:, None] - xy, dim=2, p=2)
        d[torch.arange(n), torch.arange(n)] = 1e9
        _, idx = torch.topk(d, 12, dim=1, largest=False)
        heu = torch.full_like(d, 1e-10).scatter_(1, idx, torch.rand(n, 12, device=DEV) * 0.9 + 0.05)
        col = AntShardedColony(CudaTspBackend(d, heu), torch.ones(n, n, device=DEV), A)
        low = col.run(4, seed=4242, offset=16)
        ret[rank] = (col.pheromone.cpu().numpy(), float(low), col.shortest_path.cpu().numpy(), col.collectives)
    finally:
        dist.destroy_process_group()


def _spawn(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank, args=(world, port, ret), nprocs=world, join=True)
    return dict(ret)


def test_ant_sharded_cuda_backend_is_world_size_invariant_and_matches_reference():
    one, three = _spawn(1), _spawn(3)
    for r in range(3):
        assert np.array_equal(three[r][0], one[0][0]) and three[r][1] == one[0][1]
        assert np.array_equal(three[r][2], one[0][2]) and three[r][3] == 4
    # and equals the reference semantics under the same Philox stream (seed 4242, offset 16)
    torch.manual_seed(0)
    n, A = 60, 50
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    _, idx = torch.topk(d, 12, dim=1, largest=False)
    heu = torch.full_like(d, 1e-10).scatter_(1, idx, torch.rand(n, 12, device=DEV) * 0.9 + 0.05)
    torch.manual_seed(4242)
    torch.cuda.default_generators[0].set_offset(16)
    ref = O.TspColony(d, A, heuristic=heu)
    ref.run(4)
    assert np.array_equal(ref.pheromone.cpu().numpy(), one[0][0])
    assert float(ref.lowest_cost) == one[0][1]


def test_run_restarts_from_state_rebound_or_mutated_between_calls():
    """The reference keeps no cached device state: `aco.heuristic = X`, `aco.pheromone = Y`, an in-place edit of either,
    or `aco.lowest_cost = c` between run() calls take effect on the next call (tsp/, tsp_nls/ and cvrp/ alike)."""
    from deepaco_b200.cvrp.aco import ACO as CvrpACO
    from deepaco_b200.cvrp.utils import gen_instance
    from deepaco_b200.tsp.aco import ACO as TspACO
    torch.manual_seed(1)
    n = 40
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    heu2 = torch.rand(n, n, device=DEV) + 0.1

    def tsp_case(edit):
        torch.manual_seed(5)
        a = TspACO(d, n_ants=16, device=DEV)
        a.run(2)
        edit(a)
        a.run(2)
        torch.manual_seed(5)
        b = TspACO(d, n_ants=16, device=DEV)
        b.run(2)
        edit(b)
        b._runner = None                                   # what a fresh runner built from the edited state gives
        b.run(2)
        assert torch.equal(a.pheromone, b.pheromone) and float(a.lowest_cost) == float(b.lowest_cost)
        return a

    tsp_case(lambda a: setattr(a, "heuristic", heu2))
    tsp_case(lambda a: a.heuristic.mul_(heu2))
    tsp_case(lambda a: a.pheromone.fill_(1.0))
    demand, dist = gen_instance(20, DEV)

    def cvrp_case(edit):
        out = []
        for drop in (False, True):
            torch.manual_seed(9)
            a = CvrpACO(dist, demand, n_ants=16, device=DEV)
            a.run(2)
            edit(a)
            if drop:
                a._runner = None
            a.run(2)
            out.append((a.pheromone.clone(), float(a.lowest_cost)))
        assert torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]

    cvrp_case(lambda a: setattr(a, "pheromone", torch.ones_like(dist)))
    cvrp_case(lambda a: setattr(a, "heuristic", torch.rand_like(dist) + 0.1))
    cvrp_case(lambda a: setattr(a, "lowest_cost", float("inf")))
