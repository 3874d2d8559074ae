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

`DeviceShardedColony` is the product path for ant sharding: all T iterations are enqueued by ONE C call per rank
(`deepaco_tsp_run_shard`): the sampling kernel stores finished tours into every rank's peer-mapped buffer, a one-CTA
flag barrier replaces the collective, and there is no host sync and no Python between iterations.
`AntShardedColony` is the same protocol driven from Python with an injectable compute backend, so that it can be
exercised with `gloo` on CPU in the test-suite (and with NCCL all-gathers where peer mapping is unavailable).
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


def gather_colony_results_packed(lowest_cost: torch.Tensor, shortest_path: torch.Tensor, counts, group=None):
    """Same result as `gather_colony_results` with ONE collective: best cost (fp32 bits) and best tour (int64) of every
    colony travel in one int64 [count, 1 + n] block per rank (`all_gather_into_tensor`; ragged shards are padded)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = shortest_path.shape[-1]
    mx = max(counts)
    block = torch.zeros((mx, n + 1), dtype=torch.int64, device=shortest_path.device)
    b = lowest_cost.shape[0]
    block[:b, 0] = lowest_cost.contiguous().view(torch.int32).to(torch.int64)
    block[:b, 1:] = shortest_path
    if world == 1:
        allb = block[None]
    else:
        allb = torch.empty((world * mx, n + 1), dtype=torch.int64, device=block.device)   # rank-major concatenation
        dist.all_gather_into_tensor(allb, block, group=group)
        allb = allb.view(world, mx, n + 1)
    if all(c == mx for c in counts):
        flat = allb.reshape(world * mx, n + 1)
    else:
        flat = torch.cat([allb[r, :counts[r]] for r in range(world)])
    return flat[:, 0].to(torch.int32).view(torch.float32), flat[:, 1:]


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


# ---- device-side ant-sharded run (deepaco_tsp_run_shard) ------------------------------------------------------------
TOUR_BUFFERS = 2          # double buffered by iteration parity (see csrc/tsp_shard.cu)
FLAG_BYTES = 128          # uint32 [8] flag words, padded to a cache line


class PeerMemory:
    """Tour buffers + flag words of every rank as addresses valid on THIS rank's device.
    tour_ptrs[r][k]: rank r's tour buffer k (uint16 [B, A, n]); flag_ptrs[r]: rank r's flag words (uint32 [8])."""

    def __init__(self, rank, world, tour_ptrs, flag_ptrs, keep_alive, barrier=None):
        self.rank, self.world = int(rank), int(world)
        self.tour_ptrs, self.flag_ptrs = tour_ptrs, flag_ptrs
        self._keep, self._barrier = keep_alive, barrier

    def host_barrier(self):
        if self._barrier is not None:
            self._barrier()


def _peer_layout(B, A, n):
    tour_bytes = (B * A * n * 2 + 255) // 256 * 256
    return tour_bytes, TOUR_BUFFERS * tour_bytes + FLAG_BYTES


def _init_peer_block(buf, B, A, n, tour_bytes):
    """Flag words zero; tour buffers hold identity tours, so that a rank whose peer never arrives (barrier time-out,
    reported through `status`) still replays valid permutations instead of indexing with garbage."""
    buf.zero_()
    ident = torch.arange(n, dtype=torch.int32, device=buf.device).to(torch.uint16).view(torch.uint8).repeat(B * A)
    for k in range(TOUR_BUFFERS):
        buf[k * tour_bytes:k * tour_bytes + ident.numel()] = ident


def symmetric_peer_memory(B, A, n, device, group=None) -> PeerMemory:
    """One symmetric-memory allocation per rank (`torch.distributed._symmetric_memory`): [tours 0 | tours 1 | flags],
    peer-mapped on all ranks of `group` over NVLink / NVSwitch."""
    import torch.distributed._symmetric_memory as symm_mem
    group = group if group is not None else dist.group.WORLD
    tour_bytes, total = _peer_layout(B, A, n)
    buf = symm_mem.empty((total,), dtype=torch.uint8, device=device)
    _init_peer_block(buf, B, A, n, tour_bytes)
    handle = symm_mem.rendezvous(buf, group)
    torch.cuda.synchronize(device)
    handle.barrier(channel=0)               # every rank's flag words are zero before anyone signals
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    base = [int(p) for p in handle.buffer_ptrs]
    return PeerMemory(rank, world, [[base[r] + k * tour_bytes for k in range(TOUR_BUFFERS)] for r in range(world)],
                      [base[r] + TOUR_BUFFERS * tour_bytes for r in range(world)], (buf, handle),
                      barrier=lambda: handle.barrier(channel=0))


def local_peer_memory(B, A, n, device, world):
    """`world` virtual ranks on ONE device (plain device pointers are their own peer mapping): exercises the complete
    device-side protocol -- fused peer stores, flag barrier, double buffering -- on a single GPU, each virtual rank on its
    own stream.  -> list of PeerMemory, one per virtual rank.

    One process drives all virtual ranks, so the FIRST launch of a kernel must not happen while another rank's barrier
    kernel is already spinning: CUDA loads kernels lazily and loading synchronises with the device, i.e. the host would
    wait for a barrier that waits for launches the host has not issued yet (until the barrier's time-out).  Run the same
    shapes once with world = 1 first (`warm_virtual_ranks`); real multi-GPU runs (one process per GPU) have no such
    coupling."""
    tour_bytes, total = _peer_layout(B, A, n)
    bufs = [torch.empty((total,), dtype=torch.uint8, device=device) for _ in range(world)]
    for b in bufs:
        _init_peer_block(b, B, A, n, tour_bytes)
    base = [b.data_ptr() for b in bufs]
    tour_ptrs = [[base[r] + k * tour_bytes for k in range(TOUR_BUFFERS)] for r in range(world)]
    flag_ptrs = [base[r] + TOUR_BUFFERS * tour_bytes for r in range(world)]
    return [PeerMemory(r, world, tour_ptrs, flag_ptrs, bufs) for r in range(world)]


def warm_virtual_ranks(make_runner, world=1, n_iterations=1, seed=0):
    """Load every kernel of the sharded path (see `local_peer_memory`): one world-1 run of a runner built by
    `make_runner()` (cost / best / update variants of the full colony), plus one construction launch per distinct shard
    size of a `world`-way split (the construction kernel's instantiation depends on the ants per launch); synchronised."""
    from . import _engine as E
    r = make_runner()
    col = DeviceShardedColony(r, local_peer_memory(r.B, r.n_ants, r.n, r.dev, 1)[0])
    col.run(n_iterations, seed)
    for count in {shard_range(r.n_ants, world, k)[1] for k in range(world)} - {0, r.n_ants}:
        E.tsp_sample_shard(r.product, None, count, 0, r.n_ants, start_node=r.start_node, double_norm=r.double_norm, seed=seed,
                           knn=r.knn)
    torch.cuda.synchronize(r.dev)
    col.check()


def shard_tables(peer: PeerMemory):
    """Flattened host tables of deepaco_shard_args: (peer_tours_host [world * 2], peer_flags_host [world])."""
    return [p for r in range(peer.world) for p in peer.tour_ptrs[r]], list(peer.flag_ptrs)


class DeviceShardedColony:
    """TSP colonies whose ants are split over the ranks of `peer` (deepaco_tsp_run_shard).  State (pheromone,
    lowest_cost, shortest_path) lives in `runner` (an `_engine.TspRunner` with the FULL ant count) and ends up identical
    on every rank -- and identical to the single-GPU `TspRunner.run`."""

    def __init__(self, runner, peer: PeerMemory, timeout_ms=2000):
        self.runner, self.peer = runner, peer
        self.a0, self.count = shard_range(runner.n_ants, peer.world, peer.rank)
        self.epoch = 0
        self.timeout_ms = int(timeout_ms)
        self.status = torch.zeros(1, dtype=torch.int32, device=runner.dev)
        self.collectives = 0                 # NCCL collectives issued on the data path: none

    def run(self, n_iterations, seed, offset=0, offsets=None, sample_events=None):
        """Enqueue n_iterations on the current stream (no host sync).  Call `check()` after synchronising."""
        self.runner.run_shard(n_iterations, seed, self.peer, self.a0, self.count, self.epoch, self.status, self.timeout_ms,
                              offset=offset, offsets=offsets, sample_events=sample_events)
        self.epoch += int(n_iterations)
        return self.runner.lowest_cost

    def check(self):
        """Host-side check (synchronises): raises if a peer never arrived at a barrier."""
        s = int(self.status.item())
        if s:
            from ._lib import DeepAcoError
            raise DeepAcoError(f"ant-sharded run: rank {self.peer.rank} timed out waiting for rank {s - 1}")
