"""`ACO` for TSP with the constructor, methods and attributes of the reference class
(reference tsp/aco.py:4-177), backed by the sm_100a kernels in libdeepaco_b200.so.

Drop-in use: `from deepaco_b200.tsp.aco import ACO` instead of `from aco import ACO` in
tsp/test.ipynb / tsp/train.ipynb.  Tensors must live on a CUDA device (the engine has no CPU path).
Under one `torch.manual_seed` the tours are the ones the reference produces on the same GPU: the
kernels consume the default CUDA generator exactly as `torch.randint` + `Categorical.sample` would.

`run()` keeps the whole iteration loop on the device (deepaco_tsp_run): `lowest_cost` is returned as a
0-d CUDA tensor, like the reference does, and is only synchronised when the caller reads it.
"""
from __future__ import annotations

import torch

from .. import _engine as E
from .._lib import DeepAcoError, generator_state, require_cuda


class ACO:
    _START_NODE = -1       # tsp/aco.py:141 draws the start with torch.randint
    _DOUBLE_NORM = False

    def __init__(self, distances, n_ants=20, decay=0.9, alpha=1, beta=1, elitist=False, min_max=False,
                 pheromone=None, heuristic=None, min=None, device='cpu'):
        require_cuda(distances, "distances")
        self.problem_size = len(distances)
        self.distances = distances
        self.n_ants = n_ants
        self.decay = decay
        self.alpha = alpha
        self.beta = beta
        self.elitist = elitist
        self.min_max = min_max
        if min_max:                                    # tsp/aco.py:29-35
            if min is not None:
                assert min > 1e-9
            else:
                min = 0.1
            self.min = min
            self.max = None
        if pheromone is None:
            pheromone = torch.ones_like(self.distances)
            if min_max:
                pheromone = pheromone * self.min
        self._pheromone = pheromone
        self.heuristic = 1 / distances if heuristic is None else heuristic
        self._shortest_path = None
        self._lowest_cost = float('inf')
        self.device = distances.device if str(device) == 'cpu' else torch.device(device)
        if self.device.type != 'cuda':
            raise DeepAcoError("deepaco_b200 ACO needs a CUDA device")
        self._runner = None      # device-resident run() state, created lazily

    # ---- state that run() keeps on the device ------------------------------------------------
    @property
    def pheromone(self):
        return self._pheromone

    @pheromone.setter
    def pheromone(self, value):
        self._pheromone = value
        self._runner = None      # externally replaced: run() restarts from the new matrix

    @property
    def lowest_cost(self):
        return self._lowest_cost

    @lowest_cost.setter
    def lowest_cost(self, value):
        self._lowest_cost = value
        self._runner = None

    @property
    def shortest_path(self):
        return self._shortest_path

    @shortest_path.setter
    def shortest_path(self, value):
        self._shortest_path = value

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sparsify(self, k_sparse):
        '''Vanilla-ACO heuristic: keep the k nearest per row, others 1/1e10 (tsp/aco.py:51-67).
        One-off set-up outside the hot path: plain tensor ops.'''
        _, nearest = torch.topk(self.distances, k=k_sparse, dim=1, largest=False)
        rows = torch.arange(len(self.distances), device=self.distances.device).repeat_interleave(k_sparse)
        cols = nearest.flatten()
        kept = torch.full_like(self.distances, 1e10)
        kept[rows, cols] = self.distances[rows, cols]
        self.heuristic = 1 / kept
        self._runner = None

    def sample(self):
        paths, log_probs = self.gen_path(require_prob=True)
        return self.gen_path_costs(paths), log_probs

    def _candidates(self):
        """Per-row candidate columns for the sparse-heuristic kernel (None for dense heuristics); cached."""
        key = (self.heuristic.data_ptr(), self.heuristic._version)
        if getattr(self, "_knn_key", None) != key:
            self._knn, self._knn_key = E.sparse_candidates(self.heuristic), key
        return self._knn

    def _weights(self):
        """(pheromone ** alpha, heuristic ** beta); alpha = beta = 1 (every reference driver) costs nothing."""
        ph = self._pheromone if self.alpha == 1 else self._pheromone ** self.alpha
        heu = self.heuristic if self.beta == 1 else self.heuristic ** self.beta
        return ph, heu

    def gen_path(self, require_prob=False):
        '''Tour construction for all ants (tsp/aco.py:134-163; tsp_nls/aco.py:184-220 for the subclass).
        Returns paths [problem_size, n_ants] int64 (and log_probs [problem_size-1, n_ants]).'''
        ph, heu = self._weights()
        gen, seed, offset = generator_state(self.device)
        gen.set_offset(offset + E.tsp_sample_offset_increment(self.problem_size, self.n_ants, self._START_NODE))

        def construct(ph_, heu_):
            paths_, logp_, _ = E.tsp_sample(ph_, heu_, self.n_ants, start_node=self._START_NODE,
                                            double_norm=self._DOUBLE_NORM, seed=seed, offset=offset, want_logp=require_prob,
                                            knn=None if require_prob else self._candidates())
            return paths_, logp_

        if require_prob and torch.is_grad_enabled() and (ph.requires_grad or heu.requires_grad):
            # REINFORCE training path: log-probs differentiable w.r.t. heuristic (and pheromone)
            return E.SampleLogProbs.apply(ph, heu, construct, None, 0.0)
        paths, logp = construct(ph.detach(), heu.detach())
        return (paths, logp) if require_prob else paths

    def pick_move(self, prev, mask, require_prob):
        '''One construction step for caller-held state (tsp/aco.py:165-177): prev [n_ants] previous nodes, mask
        [n_ants, problem_size] with 0 for visited cities -> (actions [n_ants], log_probs [n_ants] | None).  gen_path
        does not go through here (its masks never leave the kernel).'''
        return E.pick_move_for(self, prev, mask, None, require_prob)

    @torch.no_grad()
    def gen_path_costs(self, paths):
        assert paths.shape == (self.problem_size, self.n_ants)
        costs, _ = E.tsp_cost(self.distances, paths=paths)
        return costs

    @torch.no_grad()
    def update_pheronome(self, paths, costs):
        '''Evaporate + deposit (tsp/aco.py:94-118).  Like the reference, binds a NEW pheromone tensor.'''
        _, nbr = E.tsp_cost(self.distances, paths=paths, want_costs=False, want_neighbours=True)
        ph = self._pheromone.detach().to(torch.float32).clone(memory_format=torch.contiguous_format)
        E.tsp_update_(ph, nbr, costs, decay=self.decay, elitist=self.elitist, min_max=self.min_max,
                      ph_min=self.min if self.min_max else 0.0, ph_max=self.max if self.min_max else None)
        self.pheromone = ph

    # ------------------------------------------------------------------------------------------
    def _make_runner(self):
        if self.alpha != 1 or self.beta != 1:
            return None      # general exponents: per-iteration path below (torch.pow supplies the powers)
        r = E.TspRunner(self.distances, self.heuristic, self._pheromone, self.n_ants, decay=self.decay,
                        elitist=self.elitist, min_max=self.min_max, ph_min=self.min if self.min_max else 0.0,
                        start_node=self._START_NODE, double_norm=self._DOUBLE_NORM)
        if not isinstance(self._lowest_cost, float) or self._lowest_cost != float('inf'):
            r.lowest_cost.fill_(float(self._lowest_cost))
            if self._shortest_path is not None:
                r.shortest_path[0].copy_(self._shortest_path)
        if self.min_max and self.max is not None:
            r.ph_max.fill_(float(self.max))
        return r

    @torch.no_grad()
    def run(self, n_iterations):
        '''tsp/aco.py:74-92 for n_iterations, entirely on the device; returns lowest_cost (0-d tensor).'''
        self._check_runner()
        if self._runner is None:
            self._runner = self._make_runner()
        r = self._runner
        if r is None:
            return self._run_stepwise(n_iterations)
        gen, seed, offset = generator_state(self.device)
        r.run(n_iterations, seed, offset)
        gen.set_offset(offset + n_iterations * r.increment)
        # snapshots: like the reference, earlier results are never modified by later run() calls
        self._pheromone = r.pheromone[0].clone()
        self._lowest_cost = r.lowest_cost[0].clone()
        self._shortest_path = r.shortest_path[0].clone()
        if self.min_max:
            self.max = r.ph_max[0].clone()
        self._runner_key = self._state_key()
        return self._lowest_cost

    def _state_key(self):
        return E.state_key(self._pheromone, self.heuristic, self.distances, self._lowest_cost,
                           self.max if self.min_max else None, self.n_ants, self.decay, self.elitist)

    def _check_runner(self):
        """Drop the cached device state when anything it was built from has been rebound or mutated in place since the
        last run() (`aco.heuristic = ...`, `aco.pheromone.mul_(2)`, `aco.n_ants = ...`): the reference keeps no cache."""
        if self._runner is not None and getattr(self, "_runner_key", None) != self._state_key():
            self._runner = None

    def _run_stepwise(self, n_iterations):
        for _ in range(n_iterations):
            paths = self.gen_path(require_prob=False)
            costs = self.gen_path_costs(paths)
            best = torch.argmin(costs)
            if costs[best] < self._lowest_cost:
                self._shortest_path, self._lowest_cost = paths[:, best], costs[best]
                if self.min_max:
                    new_max = self.problem_size / self._lowest_cost
                    if self.max is None:
                        self._pheromone = self._pheromone * (new_max / self._pheromone.max())
                    self.max = new_max
            self.update_pheronome(paths, costs)
            self._runner = None
        return self._lowest_cost
