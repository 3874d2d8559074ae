"""`ACO` for TSP with the constructor, methods and attributes of the reference class
(reference tsp/aco.py:4-177), backed by the sm_100a kernels in libdeepaco_b200.so.

Drop-in use: `from deepaco_b200.tsp.aco import ACO` instead of `from aco import ACO` in
tsp/test.ipynb / tsp/train.ipynb.  Tensors must live on a CUDA device (the engine has no CPU path).
Under one `torch.manual_seed` the tours are the ones the reference produces on the same GPU: the
kernels consume the default CUDA generator exactly as `torch.randint` + `Categorical.sample` would.
"""
from __future__ import annotations

import torch

from .. import _engine as E
from .._lib import DeepAcoError, generator_state, require_cuda


class ACO:

    def __init__(self,
                 distances,
                 n_ants=20,
                 decay=0.9,
                 alpha=1,
                 beta=1,
                 elitist=False,
                 min_max=False,
                 pheromone=None,
                 heuristic=None,
                 min=None,
                 device='cpu'):
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
            self.pheromone = torch.ones_like(self.distances)
            if min_max:
                self.pheromone = self.pheromone * self.min
        else:
            self.pheromone = pheromone

        self.heuristic = 1 / distances if heuristic is None else heuristic

        self.shortest_path = None
        self.lowest_cost = float('inf')

        self.device = distances.device if str(device) == 'cpu' else torch.device(device)
        if self.device.type != 'cuda':
            raise DeepAcoError("deepaco_b200.tsp.ACO needs a CUDA device")

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sparsify(self, k_sparse):
        '''Vanilla-ACO heuristic: keep the k nearest per row, others 1/1e10 (tsp/aco.py:51-67).
        One-off set-up outside the hot path: plain tensor ops.'''
        _, topk_indices = torch.topk(self.distances, k=k_sparse, dim=1, largest=False)
        rows = torch.arange(len(self.distances), device=self.distances.device).repeat_interleave(k_sparse)
        cols = topk_indices.flatten()
        sparse = torch.full_like(self.distances, 1e10)
        sparse[rows, cols] = self.distances[rows, cols]
        self.heuristic = 1 / sparse

    def sample(self):
        paths, log_probs = self.gen_path(require_prob=True)
        costs = self.gen_path_costs(paths)
        return costs, log_probs

    # ------------------------------------------------------------------------------------------
    def _weights(self):
        """(pheromone ** alpha, heuristic ** beta); alpha = beta = 1 (every reference driver) is free."""
        ph = self.pheromone if self.alpha == 1 else self.pheromone ** self.alpha
        heu = self.heuristic if self.beta == 1 else self.heuristic ** self.beta
        return ph, heu

    def gen_path(self, require_prob=False):
        '''Tour construction for all ants (tsp/aco.py:134-163).
        Returns paths [problem_size, n_ants] int64 (and log_probs [problem_size-1, n_ants]).'''
        ph, heu = self._weights()
        gen, seed, offset = generator_state(self.device)
        paths, logp, _ = E.tsp_sample(ph.detach(), heu.detach(), self.n_ants, start_node=-1, seed=seed, offset=offset,
                                      want_logp=require_prob)
        gen.set_offset(offset + E.tsp_sample_offset_increment(self.problem_size, self.n_ants, -1))
        if require_prob:
            return paths, logp
        return paths

    @torch.no_grad()
    def gen_path_costs(self, paths):
        assert paths.shape == (self.problem_size, self.n_ants)
        costs, _ = E.tsp_cost(self.distances, paths=paths)
        return costs

    @torch.no_grad()
    def update_pheronome(self, paths, costs):
        '''Evaporate + deposit (tsp/aco.py:94-118).  Like the reference, binds a NEW pheromone tensor.'''
        _, nbr = E.tsp_cost(self.distances, paths=paths, want_costs=False, want_neighbours=True)
        ph = self.pheromone.detach().to(torch.float32).clone(memory_format=torch.contiguous_format)
        E.tsp_update_(ph, nbr, costs, decay=self.decay, elitist=self.elitist, min_max=self.min_max,
                      ph_min=self.min if self.min_max else 0.0, ph_max=self.max if self.min_max else None)
        self.pheromone = ph

    @torch.no_grad()
    def run(self, n_iterations):
        for _ in range(n_iterations):
            paths = self.gen_path(require_prob=False)
            costs = self.gen_path_costs(paths)

            best_cost, best_idx = costs.min(dim=0)
            if best_cost < self.lowest_cost:
                self.shortest_path = paths[:, best_idx]
                self.lowest_cost = best_cost
                if self.min_max:
                    max = self.problem_size / self.lowest_cost
                    if self.max is None:
                        self.pheromone *= max / self.pheromone.max()
                    self.max = max

            self.update_pheronome(paths, costs)

        return self.lowest_cost
