"""`ACO` for CVRP with the reference class surface (reference cvrp/aco.py:9-205, adaptive=False),
backed by libdeepaco_b200.so.  Drop-in for `from aco import ACO` in cvrp/test.py / cvrp/train.ipynb.

The "adaptive elitist AS" comparison code of the reference (cvrp/aco.py:207-384) is out of scope
(SURVEY.md section 2): `adaptive=True` raises.
"""
from __future__ import annotations

import torch

from .. import _engine as E
from .._lib import DeepAcoError, generator_state, require_cuda

CAPACITY = 50


class ACO:

    def __init__(self,  # 0: depot
                 distances,  # (n, n)
                 demand,  # (n, )
                 n_ants=20,
                 decay=0.9,
                 alpha=1,
                 beta=1,
                 elitist=False,
                 min_max=False,
                 pheromone=None,
                 heuristic=None,
                 min=None,
                 device='cpu',
                 adaptive=False,
                 capacity=CAPACITY):
        if adaptive:
            raise DeepAcoError("adaptive elitist AS (cvrp/aco.py:207-384) is not part of the DeepACO path and is not provided")
        require_cuda(distances, "distances")
        self.problem_size = len(distances)
        self.distances = distances
        self.capacity = capacity
        self.demand = demand
        self.n_ants = n_ants
        self.decay = decay
        self.alpha = alpha
        self.beta = beta
        self.elitist = elitist
        self.min_max = min_max
        self.adaptive = False
        if min_max:
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

    def sample(self):
        paths, log_probs = self.gen_path(require_prob=True)
        costs = self.gen_path_costs(paths)
        return costs, log_probs

    def _weights(self):
        ph = self.pheromone if self.alpha == 1 else self.pheromone ** self.alpha
        heu = self.heuristic if self.beta == 1 else self.heuristic ** self.beta
        return ph, heu

    def gen_path(self, require_prob=False):
        '''cvrp/aco.py:138-165.  Returns paths [T+1, n_ants] int64 (T data dependent) and log_probs [T, n_ants].'''
        ph, heu = self._weights()
        gen, seed, offset = generator_state(self.device)
        out = E.cvrp_sample(ph.detach(), heu.detach(), self.demand, self.capacity, self.n_ants, seed=seed,
                            offset=offset, want_logp=require_prob)
        T = int(out["tmax"][0].item())          # one host sync per construction (the reference has one per step)
        gen.set_offset(offset + T * E.cvrp_step_offset_increment(self.problem_size, self.n_ants))
        paths = out["paths"][0, :T + 1]
        if require_prob:
            return paths, out["logp"][0, :T]
        return paths

    @torch.no_grad()
    def gen_path_costs(self, paths):
        costs, _ = E.cvrp_cost(self.distances, paths=paths)
        return costs

    @torch.no_grad()
    def update_pheronome(self, paths, costs):
        _, nbr = E.cvrp_cost(self.distances, paths=paths, want_costs=False, want_neighbours=True)
        ph = self.pheromone.detach().to(torch.float32).clone(memory_format=torch.contiguous_format)
        E.cvrp_update_(ph, nbr, costs, decay=self.decay, elitist=self.elitist, min_max=self.min_max,
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
