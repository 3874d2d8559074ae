"""Device-side ant-sharded ACO.run (deepaco_tsp_run_shard, csrc/tsp_shard.cu): fused peer stores + flag barrier +
double-buffered tours, exercised on ONE GPU with `world` virtual ranks on separate streams (plain device pointers are
their own peer mapping).  The result must be bit-identical to the single-GPU deepaco_tsp_run -- which
tests/test_gpu_tsp.py pins to the reference op sequence (tsp/aco.py:74-92) -- for every world size, for sparse (kNN
kernel) and dense (list kernel) heuristics, batched colonies, ragged ant splits, elitist and min-max variants."""
import pytest
import torch

from deepaco_b200 import _engine as E
from deepaco_b200._lib import DeepAcoError
from deepaco_b200.dist import DeviceShardedColony, local_peer_memory, shard_range, warm_virtual_ranks

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _instances(B, n, k, sparse, seed=3):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand((B, n, 2), generator=g).to(DEV)
    d = torch.cdist(xy, xy)
    idx = torch.arange(n, device=DEV)
    d[:, idx, idx] = 1e9
    if sparse:
        _, nn = torch.topk(d, k, dim=2, largest=False)
        heu = torch.full_like(d, 1e-10).scatter_(2, nn, (torch.rand((B, n, k), generator=g) * 0.9 + 0.05).to(DEV))
    else:
        heu = 1.0 / d
    return d.contiguous(), heu.contiguous()


def _run_sharded(d, heu, A, world, T, seed, offsets, calls=1, **kw):
    B, n = d.shape[0], d.shape[1]
    # one process drives all virtual ranks: load this shape's kernels before any barrier can spin (dist.local_peer_memory)
    warm_virtual_ranks(lambda: E.TspRunner(d, heu, torch.ones_like(d), A, **kw), world)
    peers = local_peer_memory(B, A, n, DEV, world)
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    cols = [DeviceShardedColony(E.TspRunner(d, heu, torch.ones_like(d), A, **kw), peers[r], timeout_ms=4000)
            for r in range(world)]
    torch.cuda.synchronize()
    inc = cols[0].runner.increment
    for c in range(calls):              # several calls: the epoch / buffer parity carries over
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                cols[r].run(T, seed, offset=c * T * inc, offsets=offsets)
    torch.cuda.synchronize()
    for col in cols:
        col.check()
    return cols


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("sparse", [True, False])
def test_sharded_run_equals_single_gpu_run(world, sparse):
    B, n, A, T = 2, 100, 250, 4          # 250 ants: ragged split at world 3 / 4
    d, heu = _instances(B, n, 20, sparse)
    offsets = [0, 4_000_000]
    single = E.TspRunner(d, heu, torch.ones_like(d), A)
    single.run(2 * T, 99, offsets=offsets)
    cols = _run_sharded(d, heu, A, world, T, 99, offsets, calls=2)
    torch.cuda.synchronize()
    for col in cols:
        r = col.runner
        assert torch.equal(r.pheromone, single.pheromone)
        assert torch.equal(r.lowest_cost, single.lowest_cost)
        assert torch.equal(r.shortest_path, single.shortest_path)
        assert col.epoch == 2 * T and col.collectives == 0


@pytest.mark.parametrize("kw", [dict(elitist=True), dict(min_max=True, ph_min=0.1)])
def test_sharded_run_variants(kw):
    B, n, A, T = 1, 64, 96, 5
    d, heu = _instances(B, n, 12, True, seed=8)
    single = E.TspRunner(d, heu, torch.ones_like(d), A, **kw)
    single.run(T, 5)
    cols = _run_sharded(d, heu, A, 3, T, 5, None, **kw)
    for col in cols:
        assert torch.equal(col.runner.pheromone, single.pheromone)
        assert torch.equal(col.runner.lowest_cost, single.lowest_cost)
        assert torch.equal(col.runner.shortest_path, single.shortest_path)


def test_sharded_run_large_colony_many_ants():
    """TSP-200 x 2048 ants (the ant-sharded bench leg's shape, scaled down): update kernel with few rows per CTA."""
    B, n, A, T = 1, 200, 2048, 2
    d, heu = _instances(B, n, 20, True, seed=11)
    single = E.TspRunner(d, heu, torch.ones_like(d), A)
    single.run(T, 21)
    cols = _run_sharded(d, heu, A, 4, T, 21, None)
    for col in cols:
        assert torch.equal(col.runner.pheromone, single.pheromone)
        assert torch.equal(col.runner.shortest_path, single.shortest_path)


def test_missing_peer_times_out_and_is_reported():
    """A rank whose peer never launches must not hang: the barrier gives up after timeout_ms and check() raises."""
    B, n, A = 1, 64, 64
    d, heu = _instances(B, n, 12, True)
    peers = local_peer_memory(B, A, n, DEV, 2)
    col = DeviceShardedColony(E.TspRunner(d, heu, torch.ones_like(d), A), peers[0], timeout_ms=50)
    col.run(3, 1)
    torch.cuda.synchronize()
    with pytest.raises(DeepAcoError):
        col.check()


def test_shard_split_matches_engine():
    assert [shard_range(250, 4, r) for r in range(4)] == [(0, 63), (63, 63), (126, 62), (188, 62)]
