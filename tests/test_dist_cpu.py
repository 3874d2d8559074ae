"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic in deepaco_b200/dist.py: shard arithmetic, the
one-collective-per-iteration ant-sharding protocol (with an injected CPU backend built from the oracle, so the
result can be compared with the single-process reference semantics), and the colony result gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepaco_b200.dist import (AntShardedColony, DeviceShardedColony, PeerMemory, colony_offsets, gather_colony_results,
                               gather_colony_results_packed, shard_range, shard_tables)


def test_shard_range_covers_everything():
    for total in (1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert colony_offsets(3, 2, 10, 400) == [12000, 16000]


class _OracleBackend:
    """CPU stand-in for CudaTspBackend: noise-explicit oracle arithmetic; the Exp(1) noise of an iteration is a
    function of (seed, offset) for the FULL colony and each rank uses only the rows of its own ants -- the same
    contract as the Philox subsequence = global ant index rule of the CUDA kernels."""

    def __init__(self, distances, heuristic):
        self.distances, self.heuristic = distances, heuristic

    def _noise(self, n, n_ants_total, seed, offset):
        g = torch.Generator().manual_seed(int(seed) * 1000003 + int(offset))
        start = torch.randint(0, n, (n_ants_total,), generator=g)
        q = torch.empty((n - 1, n_ants_total, n)).exponential_(1, generator=g)
        return start, q

    def sample(self, pheromone, a0, count, n_ants_total, seed, offset):
        n = pheromone.shape[0]
        start, q = self._noise(n, n_ants_total, seed, offset)
        cur = start[a0:a0 + count].clone()
        alive = torch.ones((count, n))
        rows = torch.arange(count)
        alive[rows, cur] = 0
        tour = [cur]
        for s in range(n - 1):
            w = pheromone[cur] * self.heuristic[cur] * alive
            p = w / w.sum(-1, keepdim=True)
            cur = torch.argmax(p / q[s, a0:a0 + count], dim=-1)
            tour.append(cur)
            alive[rows, cur] = 0
        return torch.stack(tour, dim=1).to(torch.int16)        # [count, n]; gloo-friendly dtype

    def cost_and_neighbours(self, tours):
        from oracle import aco_torch as O
        paths = tours.to(torch.int64).T
        return O.tsp_path_costs(self.distances, paths), paths

    def update_(self, pheromone, paths, costs, decay, elitist):
        from oracle import aco_torch as O
        return O.tsp_update_pheromone(pheromone, paths, costs, decay, elitist)

    def increment(self, n, n_ants_total):
        return 4 * n


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        n, A = 12, 10                                  # ragged split for world = 3, even for 2
        xy = torch.rand(n, 2)
        d = torch.norm(xy[:, None] - xy, dim=2) + torch.eye(n) * 1e9
        col = AntShardedColony(_OracleBackend(d, 1 / d), torch.ones(n, n), A)
        low = col.run(3, seed=7)
        # colony-sharded result gather (ragged)
        s, c = shard_range(5, world, rank)
        lc = torch.arange(s, s + c, dtype=torch.float32)
        sp = torch.arange(s, s + c)[:, None].repeat(1, 4)
        counts = [shard_range(5, world, r)[1] for r in range(world)]
        glc, gsp = gather_colony_results(lc, sp, counts)
        plc, psp = gather_colony_results_packed(lc + 0.25, sp, counts)      # one collective, same result
        assert torch.equal(plc, glc + 0.25) and torch.equal(psp, gsp)
        ret[rank] = (col.pheromone.numpy().copy(), float(low), col.shortest_path.numpy().copy(), col.collectives,
                     glc.numpy().copy(), gsp.numpy().copy())
    finally:
        dist.destroy_process_group()


def _run(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    return dict(ret)


def test_ant_sharding_is_independent_of_world_size():
    one, two = _run(1), _run(2)
    ph1, low1, sp1, ncoll1, _, _ = one[0]
    for r in range(2):
        ph, low, sp, ncoll, glc, gsp = two[r]
        assert np.array_equal(ph, ph1), "pheromone differs between 1 and 2 ranks"
        assert low == low1 and np.array_equal(sp, sp1)
        assert ncoll == 3                               # exactly one collective per iteration
        assert np.array_equal(glc, np.arange(5, dtype=np.float32))
        assert np.array_equal(gsp[:, 0], np.arange(5))
    assert np.array_equal(two[0][0], two[1][0])


class _FakeRunner:
    """Records what DeviceShardedColony hands to the C entry (host logic only: shard, epoch, tables)."""

    def __init__(self, n_ants):
        self.n_ants, self.dev, self.calls, self.lowest_cost, self.increment = n_ants, "cpu", [], torch.tensor([1.0]), 400

    def run_shard(self, T, seed, peer, a0, count, epoch, status, timeout_ms, **kw):
        self.calls.append((T, seed, a0, count, epoch, shard_tables(peer), kw["offset"]))


def test_device_sharded_colony_host_logic():
    world = 3
    tour_ptrs = [[1000 * r + 10 * k for k in range(2)] for r in range(world)]
    flag_ptrs = [9000 + r for r in range(world)]
    for rank in range(world):
        fr = _FakeRunner(500)
        col = DeviceShardedColony(fr, PeerMemory(rank, world, tour_ptrs, flag_ptrs, None))
        assert (col.a0, col.count) == shard_range(500, world, rank)
        col.run(4, seed=9, offset=0)
        col.run(3, seed=9, offset=4 * fr.increment)
        (T0, _, a0, cnt, e0, tabs, off0), (T1, _, _, _, e1, _, off1) = fr.calls
        assert (T0, e0, off0) == (4, 0, 0) and (T1, e1, off1) == (3, 4, 1600) and col.epoch == 7
        assert tabs == ([0, 10, 1000, 1010, 2000, 2010], [9000, 9001, 9002]) and (a0, cnt) == (col.a0, col.count)
