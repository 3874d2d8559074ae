"""Thin tensor-level wrappers over the C ABI (one function per entry point of include/deepaco_b200.h).

Tensors are torch CUDA tensors; a leading colony dimension is optional (a 2-D matrix is one colony).
Nothing here computes: it validates shapes, allocates outputs and forwards raw pointers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, f32c, lib, ptr, require_cuda, stream_ptr


def _colonies(mat: torch.Tensor):
    """-> (B, n) for a [n, n] or [B, n, n] matrix."""
    if mat.dim() == 2:
        return 1, mat.shape[0]
    if mat.dim() == 3:
        return mat.shape[0], mat.shape[1]
    raise _lib.DeepAcoError(f"expected [n,n] or [B,n,n] matrix, got shape {tuple(mat.shape)}")


def _offsets(offsets, B, dev):
    """Per-colony Philox offsets as a device int64 [B] tensor (None -> NULL: by-value offset for all)."""
    if offsets is None:
        return None
    t = torch.as_tensor(offsets, dtype=torch.int64, device=dev).contiguous()
    if t.numel() != B:
        raise _lib.DeepAcoError(f"offsets must hold one Philox offset per colony ({B})")
    return t


KNN_WIDTH = 32


def state_key(*objs):
    """Fingerprint of the tensors / scalars a cached runner was built from: (storage address, version counter, shape)
    per tensor, the value per Python scalar.  A caller that rebinds or mutates `aco.pheromone`, `aco.heuristic`,
    `aco.lowest_cost` ... between run() calls changes it, and run() then restarts from the new state, as the reference
    (which keeps no cached state) does."""
    key = []
    for o in objs:
        if isinstance(o, torch.Tensor):
            key.append((o.data_ptr(), o._version, tuple(o.shape)))
        else:
            key.append(o)
    return tuple(key)


def sparse_candidates(heuristic, ratio=1e-6):
    """Candidate lists for the kNN sampling kernel: uint8 [..., n, 32] = columns of the 32 largest heuristic
    values per row, or None when the heuristic is not sparse (the 33rd largest value of a typical row is not
    `ratio` times smaller than the largest) or n is outside (32, 256].  Set-up code, run once per instance."""
    n = heuristic.shape[-1]
    if n <= KNN_WIDTH or n > 256:
        return None
    vals, idx = torch.topk(heuristic.detach().to(torch.float32), KNN_WIDTH + 1, dim=-1)
    rel = vals[..., KNN_WIDTH] / vals[..., 0].clamp(min=1e-30)
    if float(rel.median()) > ratio:
        return None
    return idx[..., :KNN_WIDTH].to(torch.uint8).contiguous()


def aten_sum_plan(row_len: int, n_rows: int):
    bw, vec, exact = C.c_int(), C.c_int(), C.c_int()
    check(lib().deepaco_aten_sum_plan(row_len, n_rows, C.byref(bw), C.byref(vec), C.byref(exact)), "aten_sum_plan")
    return bw.value, bool(vec.value), bool(exact.value)


def tsp_sample(pheromone, heuristic, n_ants, *, start_node=-1, double_norm=False, seed=0, offset=0, offsets=None,
               noise=None, start=None, want_paths=True, want_logp=False, want_tours=False, knn=None):
    """deepaco_tsp_sample.  Returns (paths|None, log_probs|None, tours|None) shaped like the inputs'
    batching: single colony -> paths [n, A]; batched -> [B, n, A]."""
    pheromone = f32c(require_cuda(pheromone, "pheromone"))
    B, n = _colonies(pheromone)
    batched = pheromone.dim() == 3
    dev = pheromone.device
    if heuristic is not None:
        heuristic = f32c(require_cuda(heuristic, "heuristic"))
        if heuristic.shape != pheromone.shape:
            raise _lib.DeepAcoError("heuristic / pheromone shape mismatch")
    if noise is not None:
        noise = f32c(require_cuda(noise, "noise"))
        if noise.numel() != B * (n - 1) * n_ants * n:
            raise _lib.DeepAcoError(f"noise must hold [B][n-1][A][n] = {B * (n - 1) * n_ants * n} values")
    if start is not None:
        start = require_cuda(start, "start").to(torch.int64).contiguous()
    paths = torch.empty((B, n, n_ants), dtype=torch.int64, device=dev) if want_paths else None
    logp = torch.empty((B, n - 1, n_ants), dtype=torch.float32, device=dev) if want_logp else None
    tours = torch.empty((B, n_ants, n), dtype=torch.uint16, device=dev) if want_tours else None
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_sample(ptr(pheromone), ptr(heuristic), n, n_ants, B, int(start_node), int(double_norm),
                                       int(seed), int(offset), ptr(_offsets(offsets, B, dev)), ptr(noise), ptr(start), ptr(paths), ptr(logp),
                                       ptr(tours), ptr(knn), stream_ptr(dev)), "deepaco_tsp_sample")
    if not batched:
        paths = None if paths is None else paths[0]
        logp = None if logp is None else logp[0]
        tours = None if tours is None else tours[0]
    return paths, logp, tours


def tsp_sample_shard(pheromone, heuristic, n_ants_local, ant_base, n_ants_total, *, start_node=-1, double_norm=False,
                     seed=0, offset=0, offsets=None, knn=None):
    """deepaco_tsp_sample_shard: compact tours (uint16 [B, n_ants_local, n]) of ants
    [ant_base, ant_base + n_ants_local) of colonies with n_ants_total ants."""
    pheromone = f32c(require_cuda(pheromone, "pheromone"))
    B, n = _colonies(pheromone)
    dev = pheromone.device
    heuristic = None if heuristic is None else f32c(require_cuda(heuristic, "heuristic"))
    tours = torch.empty((B, n_ants_local, n), dtype=torch.uint16, device=dev)
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_sample_shard(ptr(pheromone), ptr(heuristic), n, n_ants_local, B, int(start_node),
                                             int(double_norm), int(seed), int(offset), ptr(_offsets(offsets, B, dev)),
                                             None, None, ptr(tours), ptr(knn), int(ant_base), int(n_ants_total),
                                             stream_ptr(dev)), "deepaco_tsp_sample_shard")
    return tours


def tsp_sample_shard_p2p(pheromone, heuristic, n_ants_local, ant_base, n_ants_total, peer_ptrs, *, start_node=-1,
                         double_norm=False, seed=0, offset=0, offsets=None, knn=None):
    """deepaco_tsp_sample_shard_p2p: tours of ants [ant_base, ant_base + n_ants_local) are written by the kernel
    into every rank's [B, n_ants_total, n] uint16 buffer (`peer_ptrs` = peer-mapped device addresses)."""
    pheromone = f32c(require_cuda(pheromone, "pheromone"))
    B, n = _colonies(pheromone)
    dev = pheromone.device
    heuristic = None if heuristic is None else f32c(require_cuda(heuristic, "heuristic"))
    arr = (C.c_uint64 * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_sample_shard_p2p(ptr(pheromone), ptr(heuristic), n, n_ants_local, B, int(start_node),
                                                 int(double_norm), int(seed), int(offset), ptr(_offsets(offsets, B, dev)),
                                                 ptr(knn), int(ant_base), int(n_ants_total), C.cast(arr, C.c_void_p),
                                                 len(peer_ptrs), stream_ptr(dev)), "deepaco_tsp_sample_shard_p2p")


def tsp_roulette_sample(prob, n_ants, *, start_node=0, seed=0, offset=0, offsets=None, want_paths=True, want_tours=False):
    """deepaco_tsp_roulette_sample (the reference's inference sampler, tsp_nls/aco.py:260-297) -> (paths | None, tours | None)."""
    prob = f32c(require_cuda(prob, "prob"))
    B, n = _colonies(prob)
    batched = prob.dim() == 3
    dev = prob.device
    paths = torch.empty((B, n, n_ants), dtype=torch.int64, device=dev) if want_paths else None
    tours = torch.empty((B, n_ants, n), dtype=torch.uint16, device=dev) if want_tours else None
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_roulette_sample(ptr(prob), n, n_ants, B, int(start_node), int(seed), int(offset),
                                                ptr(_offsets(offsets, B, dev)), ptr(tours), ptr(paths), stream_ptr(dev)),
              "deepaco_tsp_roulette_sample")
    if not batched:
        paths = None if paths is None else paths[0]
        tours = None if tours is None else tours[0]
    return paths, tours


def tsp_roulette_offset_increment(n, n_ants) -> int:
    return int(lib().deepaco_tsp_roulette_offset_increment(n, n_ants))


def tsp_sample_offset_increment(n, n_ants, start_node=-1) -> int:
    return int(lib().deepaco_tsp_sample_offset_increment(n, n_ants, int(start_node)))


def tsp_cost(distances, paths=None, tours=None, *, want_costs=True, want_neighbours=False):
    """deepaco_tsp_cost -> (costs [A] | [B, A], neighbours uint32 [n, A] | [B, n, A])."""
    distances = f32c(require_cuda(distances, "distances"))
    B, n = _colonies(distances)
    batched = distances.dim() == 3
    dev = distances.device
    if (paths is None) == (tours is None):
        raise _lib.DeepAcoError("pass exactly one of paths / tours")
    if paths is not None:
        paths = require_cuda(paths, "paths")
        if paths.dtype != torch.int64:
            paths = paths.to(torch.int64)
        paths = paths.contiguous()
        n_ants = paths.shape[-1]
        if paths.shape[-2] != n:
            raise _lib.DeepAcoError(f"paths has {paths.shape[-2]} rows, expected problem_size {n}")
    else:
        tours = require_cuda(tours, "tours").contiguous()
        if tours.dtype != torch.uint16:
            raise _lib.DeepAcoError("tours must be uint16")
        n_ants = tours.shape[-2]
    costs = torch.empty((B, n_ants), dtype=torch.float32, device=dev) if want_costs else None
    nbr = torch.empty((B, n, n_ants), dtype=torch.int32, device=dev) if want_neighbours else None
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_cost(ptr(distances), ptr(paths), ptr(tours), n, n_ants, B, ptr(costs), ptr(nbr),
                                     stream_ptr(dev)), "deepaco_tsp_cost")
    if not batched:
        costs = None if costs is None else costs[0]
        nbr = None if nbr is None else nbr[0]
    return costs, nbr


def tsp_update_(pheromone, neighbours, costs, *, decay=0.9, elitist=False, min_max=False, ph_min=0.0, ph_max=None):
    """deepaco_tsp_update, in place on `pheromone` (must be fp32 contiguous CUDA)."""
    require_cuda(pheromone, "pheromone")
    if pheromone.dtype != torch.float32 or not pheromone.is_contiguous():
        raise _lib.DeepAcoError("pheromone must be contiguous fp32 for the in-place update")
    B, n = _colonies(pheromone)
    n_ants = costs.shape[-1]
    costs = f32c(costs)
    dev = pheromone.device
    if min_max:
        ph_max = f32c(torch.as_tensor(ph_max, device=dev).reshape(-1))
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_update(ptr(pheromone), ptr(neighbours), ptr(costs), n, n_ants, B, float(decay),
                                       int(elitist), int(min_max), float(ph_min), ptr(ph_max) if min_max else None,
                                       stream_ptr(dev)), "deepaco_tsp_update")
    return pheromone


def tsp_update_tours_(pheromone, tours, costs, *, decay=0.9, elitist=False, min_max=False, ph_min=0.0, ph_max=None):
    """deepaco_tsp_update_tours, in place on `pheromone`: the ant-sequential update from compact uint16 tours."""
    require_cuda(pheromone, "pheromone")
    if pheromone.dtype != torch.float32 or not pheromone.is_contiguous():
        raise _lib.DeepAcoError("pheromone must be contiguous fp32 for the in-place update")
    B, n = _colonies(pheromone)
    _check_tours(tours, n)
    n_ants = tours.shape[-2]
    costs = f32c(costs)
    dev = pheromone.device
    if min_max:
        ph_max = f32c(torch.as_tensor(ph_max, device=dev).reshape(-1))
    with torch.cuda.device(dev):
        check(lib().deepaco_tsp_update_tours(ptr(pheromone), ptr(tours), ptr(costs), n, n_ants, B, float(decay), int(elitist),
                                             int(min_max), float(ph_min), ptr(ph_max) if min_max else None, stream_ptr(dev)),
              "deepaco_tsp_update_tours")
    return pheromone


class TspRunner:
    """Device-resident state of `ACO.run` for one or many TSP colonies (deepaco_tsp_run).

    Owns the pheromone (a private contiguous copy), the product matrix and the per-iteration scratch, and
    keeps lowest_cost / shortest_path on the device, so T iterations are launched without a host sync.
    """

    def __init__(self, distances, heuristic, pheromone, n_ants, *, decay=0.9, elitist=False, min_max=False,
                 ph_min=0.0, start_node=-1, double_norm=False, use_knn=True):
        self.distances = f32c(require_cuda(distances, "distances"))
        self.B, self.n = _colonies(self.distances)
        self.batched = self.distances.dim() == 3
        dev = self.dev = self.distances.device
        shp = (self.B, self.n, self.n)
        self.heuristic = f32c(require_cuda(heuristic, "heuristic").detach())
        self.pheromone = pheromone.detach().to(torch.float32).clone(memory_format=torch.contiguous_format)
        if self.heuristic.numel() != self.B * self.n * self.n or self.pheromone.numel() != self.B * self.n * self.n:
            raise _lib.DeepAcoError("distances / heuristic / pheromone shape mismatch")
        # all state is kept with an explicit leading colony dimension
        self.distances, self.heuristic, self.pheromone = (t.reshape(shp) for t in (self.distances, self.heuristic, self.pheromone))
        self.n_ants = int(n_ants)
        self.product = torch.empty(shp, dtype=torch.float32, device=dev)
        self.product_valid = False
        self.tours = torch.empty((self.B, self.n_ants, self.n), dtype=torch.uint16, device=dev)
        self.costs = torch.empty((self.B, self.n_ants), dtype=torch.float32, device=dev)
        self.neighbours = torch.empty((self.B, self.n, self.n_ants), dtype=torch.int32, device=dev)
        self.lowest_cost = torch.full((self.B,), float("inf"), dtype=torch.float32, device=dev)
        self.shortest_path = torch.zeros((self.B, self.n), dtype=torch.int64, device=dev)
        self.ph_max = torch.zeros((self.B,), dtype=torch.float32, device=dev)
        self.scale = torch.ones((self.B,), dtype=torch.float32, device=dev)
        self.decay, self.elitist, self.min_max, self.ph_min = float(decay), bool(elitist), bool(min_max), float(ph_min)
        self.start_node, self.double_norm = int(start_node), bool(double_norm)
        self.increment = tsp_sample_offset_increment(self.n, self.n_ants, self.start_node)
        self.knn = sparse_candidates(self.heuristic) if use_knn else None
        # optional local search between construction and cost (tsp_nls): 0 none, 1 2-opt, 2 NLS
        self.local_search, self.ls_max_iterations, self.T_nls, self.T_p, self.heuristic_dist = 0, 0, 10, 20, None
        self.roulette = False           # True: roulette-wheel construction (tsp_nls run(.., inference=True))
        # candidate lists are re-derived from the product matrix every `knn_refresh` iterations (the runner owns `knn`)
        self.knn_refresh, self.iterations_done = 4, 0

    def set_local_search(self, mode, max_iterations, heuristic_dist=None, T_nls=10, T_p=20):
        self.local_search = {None: 0, "2opt": 1, "nls": 2}[mode]
        self.ls_max_iterations, self.T_nls, self.T_p = int(max_iterations), int(T_nls), int(T_p)
        if self.local_search == 2:
            self.heuristic_dist = f32c(require_cuda(heuristic_dist, "heuristic_dist")).reshape(self.B, self.n, self.n)

    def _args(self, seed, offset, offs, events=None):
        ev0 = ev1 = None
        if events:
            for ev in events:      # torch creates the cudaEvent_t lazily on the first record()
                ev.record(torch.cuda.current_stream(self.dev))
            ev0, ev1 = events[0].cuda_event, events[1].cuda_event
        return _lib.TspRunArgs(self.n, self.n_ants, self.B, self.start_node, int(self.double_norm), self.decay,
                               int(self.elitist), int(self.min_max), self.ph_min, int(seed), int(offset), ptr(offs),
                               ptr(self.pheromone), ptr(self.heuristic), ptr(self.distances), ptr(self.product),
                               int(self.product_valid), ptr(self.tours), ptr(self.costs), ptr(self.neighbours),
                               ptr(self.lowest_cost), ptr(self.shortest_path), ptr(self.ph_max), ptr(self.scale), ptr(self.knn),
                               self.local_search, self.ls_max_iterations, self.T_nls, self.T_p, ptr(self.heuristic_dist),
                               ev0, ev1, int(self.knn_refresh if self.knn is not None else 0), int(self.iterations_done),
                               int(self.roulette))

    def run(self, n_iterations, seed, offset=0, offsets=None, sample_events=None):
        """Launch n_iterations ACO iterations; colony b consumes offsets[b] + offset + t * self.increment.
        sample_events: optional (begin, end) torch.cuda.Event pair recorded around each sampling launch."""
        offs = _offsets(offsets, self.B, self.dev)
        a = self._args(seed, offset, offs, sample_events)
        with torch.cuda.device(self.dev):
            check(lib().deepaco_tsp_run(C.byref(a), int(n_iterations), stream_ptr(self.dev)), "deepaco_tsp_run")
        if n_iterations > 0:
            self.product_valid = True
        self.iterations_done += int(n_iterations)
        return self.lowest_cost

    def run_shard(self, n_iterations, seed, peer, ant_base, n_ants_local, epoch, status, timeout_ms=2000, *, offset=0,
                  offsets=None, sample_events=None):
        """deepaco_tsp_run_shard: this rank builds ants [ant_base, ant_base + n_ants_local) of every colony; `peer`
        (dist.PeerMemory) names every rank's tour buffers / flag words.  No host sync."""
        from .dist import shard_tables
        offs = _offsets(offsets, self.B, self.dev)
        a = self._args(seed, offset, offs, sample_events)
        tours_tab, flags_tab = shard_tables(peer)
        t_arr = (C.c_uint64 * len(tours_tab))(*tours_tab)
        f_arr = (C.c_uint64 * len(flags_tab))(*flags_tab)
        sh = _lib.ShardArgs(peer.rank, peer.world, int(ant_base), int(n_ants_local), C.cast(t_arr, C.c_void_p),
                            C.cast(f_arr, C.c_void_p), int(epoch) & 0xffffffff, int(timeout_ms), ptr(status))
        with torch.cuda.device(self.dev):
            check(lib().deepaco_tsp_run_shard(C.byref(a), C.byref(sh), int(n_iterations), stream_ptr(self.dev)),
                  "deepaco_tsp_run_shard")
        if n_iterations > 0:
            self.product_valid = True
        self.iterations_done += int(n_iterations)
        return self.lowest_cost

    def run_host(self, n_iterations, seed, distances_h, heuristic_h, pheromone_h, lowest_h, shortest_h, offset=0,
                 offsets=None, copy_back_pheromone=True):
        """deepaco_tsp_run_host: pinned HOST tensors in/out (pheromone_h is updated in place when
        copy_back_pheromone; pheromone_h = None: start from ones like ACO.__init__); synchronous."""
        for t, nm in ((distances_h, "distances"), (heuristic_h, "heuristic"), (pheromone_h, "pheromone")):
            if t is None and nm == "pheromone":
                copy_back_pheromone = False      # pheromone starts as ACO.__init__ creates it (ones), nothing to upload
                continue
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.B * self.n * self.n:
                raise _lib.DeepAcoError(f"run_host: `{nm}` must be a contiguous fp32 host tensor [B, n, n]")
        offs = _offsets(offsets, self.B, self.dev)
        a = self._args(seed, offset, offs)
        with torch.cuda.device(self.dev):
            check(lib().deepaco_tsp_run_host(C.byref(a), int(n_iterations), ptr(distances_h), ptr(heuristic_h),
                                             ptr(pheromone_h), ptr(lowest_h), ptr(shortest_h), int(copy_back_pheromone),
                                             stream_ptr(self.dev)),
                  "deepaco_tsp_run_host")
        self.product_valid = n_iterations > 0
        self.iterations_done += int(n_iterations)
        return lowest_h


# ---- local search -------------------------------------------------------------------------------
def knn_graph(coords=None, distances=None, k=0, *, diag=1e9, want_distances=True, want_neighbours=True, want_edge_index=False):
    """deepaco_knn_graph: distance matrix and k-nearest-neighbour edges of one instance ([n, 2] | [n, n]) or a batch
    ([B, n, 2] | [B, n, n]) in one launch (tsp/utils.py:4-36: torch.norm + diagonal, torch.topk(largest=False), edge_index).
    Pass `coords` (distances are computed, diagonal = diag) or `distances` (taken as is).
    Returns (distances | None, nbr_index int32 [.., n, k] | None, nbr_value f32 [.., n, k] | None, edge_index int64 [.., 2, n*k] | None)."""
    if (coords is None) == (distances is None):
        raise _lib.DeepAcoError("knn_graph: pass coords or distances, not both")
    src = f32c(require_cuda(coords if coords is not None else distances, "coords" if coords is not None else "distances"))
    batched = src.dim() == 3
    if not batched:
        src = src[None]
    B, n = src.shape[0], src.shape[1]
    if src.shape[2] != (2 if coords is not None else n):
        raise _lib.DeepAcoError("knn_graph: coords must be [.., n, 2], distances [.., n, n]")
    k = int(k)
    dev = src.device
    want_distances = bool(want_distances) and coords is not None
    d_out = torch.empty((B, n, n), dtype=torch.float32, device=dev) if want_distances else None
    idx = torch.empty((B, n, k), dtype=torch.int32, device=dev) if (k > 0 and want_neighbours) else None
    val = torch.empty((B, n, k), dtype=torch.float32, device=dev) if (k > 0 and want_neighbours) else None
    ei = torch.empty((B, 2, n * k), dtype=torch.int64, device=dev) if (k > 0 and want_edge_index) else None
    with torch.cuda.device(dev):
        check(lib().deepaco_knn_graph(ptr(src) if coords is not None else None, ptr(src) if coords is None else None, n, B, k,
                                      float(diag), ptr(d_out), ptr(idx), ptr(val), ptr(ei), stream_ptr(dev)), "deepaco_knn_graph")
    if coords is None:
        d_out = src
    pick = (lambda t: t) if batched else (lambda t: None if t is None else t[0])
    return pick(d_out), pick(idx), pick(val), pick(ei)


def paths_to_tours(paths):
    """int64 [n, A] | [B, n, A] -> uint16 [A, n] | [B, A, n]."""
    paths = require_cuda(paths, "paths").to(torch.int64).contiguous()
    batched = paths.dim() == 3
    B = paths.shape[0] if batched else 1
    n, A = paths.shape[-2], paths.shape[-1]
    tours = torch.empty((B, A, n), dtype=torch.uint16, device=paths.device)
    with torch.cuda.device(paths.device):
        check(lib().deepaco_paths_to_tours(ptr(paths), ptr(tours), n, A, B, stream_ptr(paths.device)), "paths_to_tours")
    return tours if batched else tours[0]


def tours_to_paths(tours):
    tours = require_cuda(tours, "tours").contiguous()
    batched = tours.dim() == 3
    B = tours.shape[0] if batched else 1
    A, n = tours.shape[-2], tours.shape[-1]
    paths = torch.empty((B, n, A), dtype=torch.int64, device=tours.device)
    with torch.cuda.device(tours.device):
        check(lib().deepaco_tours_to_paths(ptr(tours), ptr(paths), n, A, B, stream_ptr(tours.device)), "tours_to_paths")
    return paths if batched else paths[0]


def _check_tours(tours, n):
    require_cuda(tours, "tours")
    if tours.dtype != torch.uint16 or not tours.is_contiguous() or tours.shape[-1] != n:
        raise _lib.DeepAcoError("tours must be a contiguous uint16 tensor [..., n_ants, n] (in-place local search)")


def two_opt_(distances, tours, max_iterations, *, want_passes=False):
    """deepaco_two_opt in place on uint16 tours [A, n] | [B, A, n]."""
    distances = f32c(require_cuda(distances, "distances"))
    B, n = _colonies(distances)
    _check_tours(tours, n)
    A = tours.shape[-2]
    passes = torch.empty((B, A), dtype=torch.int32, device=tours.device) if want_passes else None
    with torch.cuda.device(tours.device):
        check(lib().deepaco_two_opt(ptr(distances), ptr(tours), n, A, B, int(max_iterations), ptr(passes),
                                    stream_ptr(tours.device)), "deepaco_two_opt")
    return (tours, passes) if want_passes else tours


def tsp_nls_(distances, heuristic_dist, tours, max_iterations, T_nls=10, T_p=20, *, want_passes=False):
    """deepaco_tsp_nls in place on uint16 tours."""
    distances = f32c(require_cuda(distances, "distances"))
    heuristic_dist = f32c(require_cuda(heuristic_dist, "heuristic_dist"))
    B, n = _colonies(distances)
    _check_tours(tours, n)
    A = tours.shape[-2]
    passes = torch.empty((B, A), dtype=torch.int32, device=tours.device) if want_passes else None
    with torch.cuda.device(tours.device):
        check(lib().deepaco_tsp_nls(ptr(distances), ptr(heuristic_dist), ptr(tours), n, A, B, int(max_iterations),
                                    int(T_nls), int(T_p), None, ptr(passes), stream_ptr(tours.device)), "deepaco_tsp_nls")
    return (tours, passes) if want_passes else tours


# ---- CVRP -------------------------------------------------------------------------------------
def cvrp_sample(pheromone, heuristic, demand, capacity, n_ants, *, seed=0, offset=0, offsets=None, noise=None,
                want_paths=True, want_logp=False, want_tours=False):
    """deepaco_cvrp_sample -> dict(paths, logp, tours, lens, tmax); buffers have 2N rows (not yet sliced)."""
    pheromone = f32c(require_cuda(pheromone, "pheromone"))
    B, N = _colonies(pheromone)
    dev = pheromone.device
    if heuristic is not None:
        heuristic = f32c(require_cuda(heuristic, "heuristic"))
    demand = f32c(require_cuda(demand, "demand"))
    if demand.numel() != B * N:
        raise _lib.DeepAcoError("demand must hold n_nodes values per colony")
    R = 2 * N
    if noise is not None:
        noise = f32c(require_cuda(noise, "noise"))
        if noise.numel() != B * (R - 1) * n_ants * N:
            raise _lib.DeepAcoError("noise must hold [B][2N-1][A][N] values")
    paths = torch.empty((B, R, n_ants), dtype=torch.int64, device=dev) if want_paths else None
    logp = torch.empty((B, R - 1, n_ants), dtype=torch.float32, device=dev) if want_logp else None
    tours = torch.empty((B, n_ants, R), dtype=torch.uint16, device=dev) if want_tours else None
    lens = torch.empty((B, n_ants), dtype=torch.int32, device=dev)
    tmax = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().deepaco_cvrp_sample(ptr(pheromone), ptr(heuristic), ptr(demand), float(capacity), N, n_ants, B,
                                        int(seed), int(offset), ptr(_offsets(offsets, B, dev)), ptr(noise), R, ptr(paths), ptr(logp),
                                        ptr(tours), ptr(lens), ptr(tmax), stream_ptr(dev)), "deepaco_cvrp_sample")
    return {"paths": paths, "logp": logp, "tours": tours, "lens": lens, "tmax": tmax, "batched": pheromone.dim() == 3}


def pick_move(pheromone_pow, heuristic_pow, prev, mask, mask2=None, *, seed=0, offset=0, want_logp=False):
    """deepaco_pick_move: one construction step for explicit masks -> (actions int64 [A], log_probs [A] | None).
    Raises (after one host read of a flag) if a `prev` index is outside the matrix."""
    pheromone_pow = f32c(require_cuda(pheromone_pow, "pheromone"))
    heuristic_pow = None if heuristic_pow is None else f32c(require_cuda(heuristic_pow, "heuristic"))
    n = pheromone_pow.shape[-1]
    dev = pheromone_pow.device
    prev = require_cuda(prev, "prev").to(torch.int64).contiguous()
    A = prev.shape[0]
    mask = f32c(require_cuda(mask, "mask"))
    if tuple(mask.shape) != (A, n):
        raise _lib.DeepAcoError(f"pick_move: mask must be [{A}, {n}], got {tuple(mask.shape)}")
    if mask2 is not None:
        mask2 = f32c(require_cuda(mask2, "capacity_mask"))
        if tuple(mask2.shape) != (A, n):
            raise _lib.DeepAcoError(f"pick_move: capacity_mask must be [{A}, {n}], got {tuple(mask2.shape)}")
    actions = torch.empty(A, dtype=torch.int64, device=dev)
    logp = torch.empty(A, dtype=torch.float32, device=dev) if want_logp else None
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().deepaco_pick_move(ptr(pheromone_pow), ptr(heuristic_pow), ptr(prev), ptr(mask), ptr(mask2), n, A,
                                      int(seed), int(offset), ptr(actions), ptr(logp), ptr(bad), stream_ptr(dev)),
              "deepaco_pick_move")
    if int(bad.item()):
        raise IndexError("pick_move: `prev` holds a node index outside the problem")      # ATen raises from the gather
    return actions, logp


def pick_move_offset_increment(n, n_ants) -> int:
    return int(lib().deepaco_pick_move_offset_increment(n, n_ants))


def pick_move_for(aco, prev, mask, mask2, require_prob):
    """ACO.pick_move of the reference classes (tsp/aco.py:165-177, cvrp/aco.py:167-174) on top of deepaco_pick_move:
    consumes the default CUDA generator like `Categorical(...).sample()`.  When autograd needs the log-probabilities
    (heuristic / pheromone require grad) they are re-expressed, for the action the kernel drew, with the tensor ops
    Categorical itself applies, so the gradient reaches the heuristic exactly as in the reference."""
    ph, heu = aco._weights()
    n = ph.shape[-1]
    gen, seed, offset = _lib.generator_state(aco.device)
    differentiable = require_prob and torch.is_grad_enabled() and (ph.requires_grad or heu.requires_grad)
    actions, logp = pick_move(ph.detach(), heu.detach(), prev, mask, mask2, seed=seed, offset=offset,
                              want_logp=require_prob and not differentiable)
    gen.set_offset(offset + pick_move_offset_increment(n, prev.shape[0]))
    if differentiable:
        x = ph[prev] * heu[prev] * mask
        if mask2 is not None:
            x = x * mask2
        probs = x / x.sum(-1, keepdim=True)
        eps = torch.finfo(probs.dtype).eps
        logp = torch.log(probs.clamp(min=eps, max=1 - eps)).gather(-1, actions[:, None]).squeeze(-1)
    return actions, logp


def cvrp_step_offset_increment(n_nodes, n_ants) -> int:
    return int(lib().deepaco_cvrp_step_offset_increment(n_nodes, n_ants))


def cvrp_cost(distances, paths=None, tours=None, *, tmax=None, T=None, want_costs=True, want_neighbours=False):
    """deepaco_cvrp_cost.  paths: int64 [rows, A] | [B, rows, A]; tours: uint16 [A, rows] | [B, A, rows].
    Path length either from the device tensor `tmax` ([B] int32) or the host int `T` (= rows - 1)."""
    distances = f32c(require_cuda(distances, "distances"))
    B, N = _colonies(distances)
    batched = distances.dim() == 3
    dev = distances.device
    if (paths is None) == (tours is None):
        raise _lib.DeepAcoError("pass exactly one of paths / tours")
    if paths is not None:
        paths = require_cuda(paths, "paths").to(torch.int64).contiguous()
        rows_in, n_ants = paths.shape[-2], paths.shape[-1]
    else:
        tours = require_cuda(tours, "tours").contiguous()
        n_ants, rows_in = tours.shape[-2], tours.shape[-1]
    if tmax is None and T is None:
        T = rows_in - 1
    costs = torch.empty((B, n_ants), dtype=torch.float32, device=dev) if want_costs else None
    nbr = torch.empty((B, N, n_ants), dtype=torch.int32, device=dev) if want_neighbours else None
    with torch.cuda.device(dev):
        check(lib().deepaco_cvrp_cost(ptr(distances), ptr(paths), ptr(tours), N, n_ants, B, rows_in, ptr(tmax),
                                      int(T or 0), ptr(costs), ptr(nbr), stream_ptr(dev)), "deepaco_cvrp_cost")
    if not batched:
        costs = None if costs is None else costs[0]
        nbr = None if nbr is None else nbr[0]
    return costs, nbr


def cvrp_update_(pheromone, neighbours, costs, *, decay=0.9, elitist=False, min_max=False, ph_min=0.0, ph_max=None):
    require_cuda(pheromone, "pheromone")
    if pheromone.dtype != torch.float32 or not pheromone.is_contiguous():
        raise _lib.DeepAcoError("pheromone must be contiguous fp32 for the in-place update")
    B, N = _colonies(pheromone)
    n_ants = costs.shape[-1]
    costs = f32c(costs)
    dev = pheromone.device
    if min_max:
        ph_max = f32c(torch.as_tensor(ph_max, device=dev).reshape(-1))
    with torch.cuda.device(dev):
        check(lib().deepaco_cvrp_update(ptr(pheromone), ptr(neighbours), ptr(costs), N, n_ants, B, float(decay),
                                        int(elitist), int(min_max), float(ph_min), ptr(ph_max) if min_max else None,
                                        stream_ptr(dev)), "deepaco_cvrp_update")
    return pheromone


class CvrpRunner:
    """Device-resident state of `ACO.run` for CVRP colonies (deepaco_cvrp_run); see TspRunner."""

    def __init__(self, distances, demand, heuristic, pheromone, n_ants, *, capacity=50, decay=0.9, elitist=False,
                 min_max=False, ph_min=0.0):
        self.distances = f32c(require_cuda(distances, "distances"))
        self.B, self.N = _colonies(self.distances)
        dev = self.dev = self.distances.device
        shp = (self.B, self.N, self.N)
        self.distances = self.distances.reshape(shp)
        self.heuristic = f32c(require_cuda(heuristic, "heuristic").detach()).reshape(shp)
        self.pheromone = pheromone.detach().to(torch.float32).clone(memory_format=torch.contiguous_format).reshape(shp)
        self.demand = f32c(require_cuda(demand, "demand")).reshape(self.B, self.N)
        self.n_ants, self.capacity = int(n_ants), float(capacity)
        R = 2 * self.N
        self.product = torch.empty(shp, dtype=torch.float32, device=dev)
        self.product_valid = False
        self.tours = torch.empty((self.B, self.n_ants, R), dtype=torch.uint16, device=dev)
        self.costs = torch.empty((self.B, self.n_ants), dtype=torch.float32, device=dev)
        self.neighbours = torch.empty((self.B, self.N, self.n_ants), dtype=torch.int32, device=dev)
        self.lens = torch.empty((self.B, self.n_ants), dtype=torch.int32, device=dev)
        self.tmax = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self.lowest_cost = torch.full((self.B,), float("inf"), dtype=torch.float32, device=dev)
        self.shortest_path = torch.zeros((self.B, R), dtype=torch.int64, device=dev)
        self.shortest_rows = torch.zeros((self.B,), dtype=torch.int32, device=dev)
        self.ph_max = torch.zeros((self.B,), dtype=torch.float32, device=dev)
        self.scale = torch.ones((self.B,), dtype=torch.float32, device=dev)
        self.offsets = torch.zeros((self.B,), dtype=torch.int64, device=dev)
        self.decay, self.elitist, self.min_max, self.ph_min = float(decay), bool(elitist), bool(min_max), float(ph_min)

    def run(self, n_iterations, seed, offsets):
        """`offsets`: per-colony Philox offsets at entry; self.offsets holds the advanced values afterwards."""
        self.offsets.copy_(torch.as_tensor(offsets, dtype=torch.int64).reshape(-1).to(self.dev))
        a = _lib.CvrpRunArgs(self.N, self.n_ants, self.B, self.capacity, self.decay, int(self.elitist), int(self.min_max),
                             self.ph_min, int(seed), ptr(self.offsets), ptr(self.pheromone), ptr(self.heuristic),
                             ptr(self.distances), ptr(self.demand), ptr(self.product), int(self.product_valid),
                             ptr(self.tours), ptr(self.costs), ptr(self.neighbours), ptr(self.lens), ptr(self.tmax),
                             ptr(self.lowest_cost), ptr(self.shortest_path), ptr(self.shortest_rows), ptr(self.ph_max),
                             ptr(self.scale))
        with torch.cuda.device(self.dev):
            check(lib().deepaco_cvrp_run(C.byref(a), int(n_iterations), stream_ptr(self.dev)), "deepaco_cvrp_run")
        if n_iterations > 0:
            self.product_valid = True
        return self.lowest_cost


# ---- autograd -----------------------------------------------------------------------------------
def logp_backward(pheromone_pow, heuristic_pow, paths, grad_logp, *, demand=None, capacity=0.0, want_pheromone_grad=False):
    """deepaco_logp_backward -> (grad wrt heuristic_pow, grad wrt pheromone_pow | None), single colony."""
    ph = f32c(require_cuda(pheromone_pow, "pheromone").detach())
    heu = f32c(require_cuda(heuristic_pow, "heuristic").detach())
    n = ph.shape[-1]
    paths = paths.to(torch.int64).contiguous()
    g = f32c(grad_logp)
    dem = None if demand is None else f32c(demand)
    g_heu = torch.zeros_like(heu)
    g_ph = torch.zeros_like(ph) if want_pheromone_grad else None
    with torch.cuda.device(ph.device):
        check(lib().deepaco_logp_backward(ptr(ph), ptr(heu), ptr(paths), ptr(g), n, paths.shape[1], paths.shape[0], ptr(dem),
                                          float(capacity), ptr(g_heu), ptr(g_ph), stream_ptr(ph.device)), "deepaco_logp_backward")
    return g_heu, g_ph


class SampleLogProbs(torch.autograd.Function):
    """paths, log_probs = construct(pheromone_pow, heuristic_pow); differentiable in both matrices.
    `construct` is a closure running the forward kernels; `replay` carries the mask rule for the backward."""

    @staticmethod
    def forward(ctx, pheromone_pow, heuristic_pow, construct, demand, capacity):
        paths, logp = construct(pheromone_pow.detach(), heuristic_pow.detach())
        ctx.save_for_backward(pheromone_pow, heuristic_pow, paths)
        ctx.demand, ctx.capacity = demand, capacity
        ctx.mark_non_differentiable(paths)
        return paths, logp

    @staticmethod
    def backward(ctx, _gpaths, glogp):
        ph, heu, paths = ctx.saved_tensors
        need_ph = ctx.needs_input_grad[0]
        g_heu, g_ph = logp_backward(ph, heu, paths, glogp, demand=ctx.demand, capacity=ctx.capacity,
                                    want_pheromone_grad=need_ph)
        return (g_ph if need_ph else None), (g_heu if ctx.needs_input_grad[1] else None), None, None, None


# ---- probes ------------------------------------------------------------------------------------
def debug_exponential(seed, offset, numel, device):
    out = torch.empty((numel,), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        check(lib().deepaco_debug_exponential(int(seed), int(offset), numel, ptr(out), stream_ptr(out.device)), "debug_exponential")
    return out


def debug_randint(seed, offset, numel, high, device):
    out = torch.empty((numel,), dtype=torch.int64, device=device)
    with torch.cuda.device(out.device):
        check(lib().deepaco_debug_randint(int(seed), int(offset), numel, int(high), ptr(out), stream_ptr(out.device)), "debug_randint")
    return out


def debug_exp_guard(device):
    """deepaco_debug_exp_guard: Philox words whose shortened Exp(1) transform differs from the literal ATen form (0)."""
    out = torch.zeros(1, dtype=torch.int64, device=device)
    with torch.cuda.device(out.device):
        check(lib().deepaco_debug_exp_guard(ptr(out), stream_ptr(out.device)), "debug_exp_guard")
    return int(out.item())


def debug_row_sum(x):
    x = f32c(require_cuda(x, "x"))
    out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().deepaco_debug_row_sum(ptr(x), x.shape[0], x.shape[1], ptr(out), stream_ptr(x.device)), "debug_row_sum")
    return out
