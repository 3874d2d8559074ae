"""`ACO` for TSP with neural-guided local search: the class surface of reference tsp_nls/aco.py:10-258.

Differences from `deepaco_b200.tsp.aco.ACO` (as in the reference): every ant starts at node 0, the
probability row is normalised once before `Categorical` normalises it again (tsp_nls/aco.py:205-207),
`sample()` also returns the paths, and `run()` applies 2-opt / NLS to the sampled tours before the
pheromone update.  The local search runs on the GPU (deepaco_two_opt / deepaco_tsp_nls), bit-exact with
the reference's numba code, so no tour ever crosses to the host.
"""
from __future__ import annotations

import torch

from .. import _engine as E
from .._lib import generator_state
from ..tsp.aco import ACO as _TspACO


class ACO(_TspACO):
    _START_NODE = 0          # tsp_nls/aco.py:191
    _DOUBLE_NORM = True      # tsp_nls/aco.py:206 + Categorical's own normalisation

    def __init__(self, distances, n_ants=20, decay=0.9, alpha=1, beta=1, elitist=False, min_max=False,
                 pheromone=None, heuristic=None, min=None, two_opt=False, device='cpu', local_search='nls'):
        super().__init__(distances, n_ants=n_ants, decay=decay, alpha=alpha, beta=beta, elitist=elitist,
                         min_max=min_max, pheromone=pheromone, heuristic=heuristic, min=min, device=device)
        assert local_search in [None, "2opt", "nls"]
        self.local_search_type = '2opt' if two_opt else local_search
        self._heuristic_dist = None

    # ---- sampling ----------------------------------------------------------------------------
    def sample(self, inference=False):
        '''tsp_nls/aco.py:80-90.  inference=True: the roulette-wheel sampler of `inference_batch_sample` (:81-85) on the
        device (deepaco_tsp_roulette_sample), start node 0, no log-probs; its uniforms come from the default CUDA
        generator (the reference's come from numba's private generator: statistical parity).'''
        if inference:
            paths = self._roulette_paths()
            return self.gen_path_costs(paths), None, paths
        paths, log_probs = self.gen_path(require_prob=True)
        return self.gen_path_costs(paths), log_probs, paths

    @torch.no_grad()
    def _roulette_paths(self):
        ph, heu = self._weights()
        probmat = (ph.detach() * heu.detach()).to(torch.float32)              # tsp_nls/aco.py:82
        gen, seed, offset = generator_state(self.device)
        paths, _ = E.tsp_roulette_sample(probmat, self.n_ants, start_node=0, seed=seed, offset=offset)
        gen.set_offset(offset + E.tsp_roulette_offset_increment(self.problem_size, self.n_ants))
        return paths

    @property
    def distances_numpy(self):
        '''tsp_nls/aco.py:222-224 (host copy for callers that want it; nothing in this class computes on it).'''
        return self.distances.detach().cpu().numpy().astype("float32")

    @property
    def heuristic_numpy(self):
        '''tsp_nls/aco.py:226-228.'''
        return self.heuristic.detach().cpu().numpy().astype("float32")

    def gen_numpy_path_costs(self, paths, numpy_distances):
        '''tsp_nls/aco.py:171-182: closed-tour lengths of host tours, paths numpy [n_ants, problem_size] (note the
        shape) -> numpy [n_ants].  A host-array helper by definition (the reference's NLS compares tours with it on
        the CPU); the device-side NLS reproduces the same float32 pairwise sum inside deepaco_tsp_nls.'''
        import numpy as np
        assert paths.shape == (self.n_ants, self.problem_size)
        return np.sum(numpy_distances[paths, np.roll(paths, shift=1, axis=1)], axis=1)

    def sample_2opt(self, paths):
        paths = self.local_search(paths)
        return self.gen_path_costs(paths), paths

    # ---- local search --------------------------------------------------------------------------
    @property
    def heuristic_dist(self):
        '''1 / (heuristic / rowmax + 1e-5)  (tsp_nls/aco.py:230-232), fp32 on the device.'''
        if self._heuristic_dist is None:
            h = self.heuristic.detach().to(torch.float32)
            self._heuristic_dist = (1 / (h / h.max(-1, keepdim=True).values + 1e-5)).contiguous()
        return self._heuristic_dist

    def _max_passes(self, inference):
        return 10000 if inference else self.problem_size // 4      # tsp_nls/aco.py:235

    @torch.no_grad()
    def two_opt(self, paths, inference=False):
        tours = E.paths_to_tours(paths)
        E.two_opt_(self.distances, tours, self._max_passes(inference))
        return E.tours_to_paths(tours)

    @torch.no_grad()
    def nls(self, paths, inference=False, T_nls=10, T_p=20):
        tours = E.paths_to_tours(paths)
        E.tsp_nls_(self.distances, self.heuristic_dist, tours, self._max_passes(inference), T_nls, T_p)
        return E.tours_to_paths(tours)

    def local_search(self, paths, inference=False):
        if self.local_search_type == "2opt":
            return self.two_opt(paths, inference)
        if self.local_search_type == "nls":
            return self.nls(paths, inference)
        return paths

    # ---- run -----------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, n_iterations, inference=False):
        '''tsp_nls/aco.py:104-129 on the device: construction, local search (2-opt / NLS), cost, best tracking and
        pheromone update of every iteration are enqueued without a host round trip (the reference crosses to the CPU
        for the numba local search each iteration).  Returns lowest_cost as a Python float like the reference (:120).'''
        if self.alpha != 1 or self.beta != 1 or self.min_max:
            # general exponents: torch.pow supplies the powers.  min_max: the reference keeps lowest_cost as a Python float
            # (:120), so max = n / lowest is computed in double and the MMAS rescale is reciprocal(ph.max()) * scalar --
            # the stepwise path reproduces exactly that; the device loop's fp32 bookkeeping is the tsp/ variant.
            return self._run_stepwise(n_iterations, inference)
        self._check_runner()
        if self._runner is None:
            self._runner = self._make_runner()
        r = self._runner
        r.set_local_search(self.local_search_type, self._max_passes(inference),
                           self.heuristic_dist if self.local_search_type == "nls" else None)
        r.roulette = bool(inference)            # tsp_nls/aco.py:106-110: inference constructs with the roulette sampler
        gen, seed, offset = generator_state(self.device)
        r.run(n_iterations, seed, offset)
        inc = E.tsp_roulette_offset_increment(self.problem_size, self.n_ants) if inference else r.increment
        gen.set_offset(offset + n_iterations * inc)
        self._pheromone = r.pheromone[0].clone()
        self._shortest_path = r.shortest_path[0].clone()
        self._lowest_cost = float(r.lowest_cost[0].item())
        self._runner_key = self._state_key()
        return self._lowest_cost

    def _run_stepwise(self, n_iterations, inference=False):
        '''The same iteration composed from the per-step methods (one host read of the best cost per iteration, as the
        reference has at tsp_nls/aco.py:120); used when alpha / beta are not 1.'''
        for _ in range(n_iterations):
            paths = self.local_search(self._roulette_paths() if inference else self.gen_path(require_prob=False), inference)
            costs = self.gen_path_costs(paths)
            best = torch.argmin(costs)
            best_cost = float(costs[best].item())
            if best_cost < self._lowest_cost:
                self._shortest_path, self._lowest_cost = paths[:, best], best_cost
                if self.min_max:
                    new_max = self.problem_size / self._lowest_cost
                    if self.max is None:
                        self._pheromone = self._pheromone * (new_max / self._pheromone.max())
                    self.max = new_max
            self.update_pheronome(paths, costs)
        return self._lowest_cost


def inference_batch_sample(probmat, count=1, startnode=None):
    '''tsp_nls/aco.py:277-297: `count` roulette-wheel tours over the numpy matrix `probmat` [n, n] -> uint16 [count, n]
    (startnode None: uniform random start per tour; an int: that start).  Runs deepaco_tsp_roulette_sample on the
    current CUDA device with the default CUDA generator; host arrays in and out like the reference function.'''
    import numpy as np
    dev = torch.device("cuda", torch.cuda.current_device())
    prob = torch.as_tensor(np.asarray(probmat, dtype=np.float32), device=dev)
    n = prob.shape[0]
    gen, seed, offset = generator_state(dev)
    _, tours = E.tsp_roulette_sample(prob, int(count), start_node=-1 if startnode is None else int(startnode), seed=seed,
                                     offset=offset, want_paths=False, want_tours=True)
    gen.set_offset(offset + E.tsp_roulette_offset_increment(n, int(count)))
    return tours.cpu().numpy()
