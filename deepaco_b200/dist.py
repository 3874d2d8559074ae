"""Multi-GPU execution of the rollout path (one process per GPU, `torch.distributed`).

Two independent ways to shard:

* **colonies** -- instances are independent: rank r owns a contiguous block of the batch, no data-path
  collective; `gather_colony_results` collects best costs / tours once at the end.  Colony b always consumes the
  Philox range of its GLOBAL index, so results do not depend on the number of ranks.  This is what `bench.py
  --gpus N` measures (weak scaling).
* **ants of one colony** -- ants are independent within an iteration.  Rank r builds ants [a0, a0 + A_r) with the
  noise words of their global indices, ONE all-gather per iteration exchanges the compact tours (2 bytes per node),
  and every rank replays cost, best tracking and the ordered deposit on the full set, so the pheromone -- and with
  it every later tour -- is bit-identical to the single-GPU run whatever the world size.  (An all-reduce of
  pheromone deltas would need less replicated work but changes the floating-point summation order.)

The compute backend is injectable so that the protocol can be exercised with `gloo` on CPU in the test-suite;
the product backend is the CUDA engine.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int):
    """Contiguous balanced split of `total` items: -> (start, count)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def colony_offsets(first_colony: int, count: int, iterations: int, increment: int, base_offset: int = 0):
    """Philox offsets of colonies [first_colony, first_colony + count) when each consumes
    `iterations * increment` of the stream in global colony order (the reference's sequential instance loop)."""
    return [base_offset + (first_colony + b) * iterations * increment for b in range(count)]


def gather_colony_results(lowest_cost: torch.Tensor, shortest_path: torch.Tensor, counts, group=None):
    """All-gather the per-colony results of every rank: -> (lowest_cost [B_total], shortest_path [B_total, n]).
    `counts[r]` = number of colonies of rank r (ragged shards are padded for the collective)."""
    world = dist.get_world_size(group)
    mx = max(counts)
    n = shortest_path.shape[-1]
    lc = torch.full((mx,), float("inf"), dtype=lowest_cost.dtype, device=lowest_cost.device)
    sp = torch.zeros((mx, n), dtype=shortest_path.dtype, device=shortest_path.device)
    lc[:lowest_cost.shape[0]] = lowest_cost
    sp[:shortest_path.shape[0]] = shortest_path
    lcs = [torch.empty_like(lc) for _ in range(world)]
    sps = [torch.empty_like(sp) for _ in range(world)]
    dist.all_gather(lcs, lc, group=group)
    dist.all_gather(sps, sp, group=group)
    return (torch.cat([lcs[r][:counts[r]] for r in range(world)]),
            torch.cat([sps[r][:counts[r]] for r in range(world)]))


class CudaTspBackend:
    """Product backend: the sm_100a kernels through the C ABI."""

    def __init__(self, distances, heuristic, *, start_node=-1, double_norm=False, use_knn=True):
        from . import _engine as E
        self.E = E
        self.distances, self.heuristic = distances, heuristic
        self.start_node, self.double_norm = start_node, double_norm
        self.knn = E.sparse_candidates(heuristic) if use_knn else None

    def sample(self, pheromone, a0, count, n_ants_total, seed, offset):
        return self.E.tsp_sample_shard(pheromone, self.heuristic, count, a0, n_ants_total, start_node=self.start_node,
                                       double_norm=self.double_norm, seed=seed, offset=offset, knn=self.knn)[0]

    def sample_p2p(self, pheromone, a0, count, n_ants_total, seed, offset, peer_ptrs):
        self.E.tsp_sample_shard_p2p(pheromone, self.heuristic, count, a0, n_ants_total, peer_ptrs,
                                    start_node=self.start_node, double_norm=self.double_norm, seed=seed, offset=offset,
                                    knn=self.knn)

    def cost_and_neighbours(self, tours):
        return self.E.tsp_cost(self.distances, tours=tours, want_neighbours=True)

    def update_(self, pheromone, neighbours, costs, decay, elitist):
        return self.E.tsp_update_(pheromone, neighbours, costs, decay=decay, elitist=elitist)

    def increment(self, n, n_ants_total):
        return self.E.tsp_sample_offset_increment(n, n_ants_total, self.start_node)


class PeerTourBuffer:
    """Symmetric-memory tour buffer (`torch.distributed._symmetric_memory`): every rank's [A, n] uint16 buffer is
    peer-mapped on all ranks (NVLink / NVSwitch), so the sampling kernel of rank r can store its ants' tours into
    all of them while it is still building the other ants.  `barrier()` (signal pads, no data) then replaces the
    collective."""

    def __init__(self, n_ants, n, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.buf = symm_mem.empty((n_ants, n * 2), dtype=torch.uint8, device=device)     # uint16 tours as bytes
        self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.tours = self.buf.view(torch.uint16).view(n_ants, n)

    def barrier(self):
        self.handle.barrier(channel=0)


class AntShardedColony:
    """One TSP colony whose ants are split over the ranks of `group` (see module docstring).
    exchange='nccl': one all-gather per iteration.  exchange='p2p': the exchange is fused into the sampling kernel
    (peer stores over NVLink into a symmetric-memory buffer) and only barriers remain."""

    def __init__(self, backend, pheromone, n_ants, *, decay=0.9, elitist=False, group=None, exchange="nccl"):
        self.backend, self.group = backend, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = pheromone.shape[-1]
        self.n_ants = n_ants
        self.a0, self.count = shard_range(n_ants, self.world, self.rank)
        self.counts = [shard_range(n_ants, self.world, r)[1] for r in range(self.world)]
        self.pheromone = pheromone.clone()
        self.decay, self.elitist = decay, elitist
        self.lowest_cost = torch.tensor(float("inf"), device=pheromone.device)
        self.shortest_path = None
        self.collectives = 0
        self.exchange = exchange
        self.peer = PeerTourBuffer(n_ants, self.n, pheromone.device, group) if exchange == "p2p" else None

    def _all_gather_tours(self, local):
        mx = max(self.counts)
        buf = torch.zeros((mx, self.n), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
        # torch has no uint16 collectives on every backend: exchange the raw bytes
        send = buf.view(torch.uint8)
        recv = [torch.empty_like(send) for _ in range(self.world)]
        dist.all_gather(recv, send, group=self.group)       # the ONE collective of the iteration
        self.collectives += 1
        return torch.cat([recv[r].view(local.dtype)[:self.counts[r]] for r in range(self.world)])

    def iterate(self, seed, offset):
        """One ACO iteration (tsp/aco.py:75-90).  Every rank ends with identical state."""
        if self.peer is not None:
            self.peer.barrier()      # every rank has finished reading the previous iteration's tours
            self.backend.sample_p2p(self.pheromone, self.a0, self.count, self.n_ants, seed, offset, self.peer.ptrs)
            self.peer.barrier()      # every rank's stores have landed everywhere
            tours = self.peer.tours
        else:
            local = self.backend.sample(self.pheromone, self.a0, self.count, self.n_ants, seed, offset)
            tours = self._all_gather_tours(local)
        costs, nbr = self.backend.cost_and_neighbours(tours)
        best = torch.argmin(costs)
        if costs[best] < self.lowest_cost:
            self.lowest_cost = costs[best].clone()
            self.shortest_path = tours[best].to(torch.int64)
        self.pheromone = self.backend.update_(self.pheromone, nbr, costs, self.decay, self.elitist)
        return self.lowest_cost

    def run(self, n_iterations, seed, offset=0):
        inc = self.backend.increment(self.n, self.n_ants)
        for t in range(n_iterations):
            self.iterate(seed, offset + t * inc)
        return self.lowest_cost
