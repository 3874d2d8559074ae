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
        self._runner = None
        self._runner_key = None

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

        def construct(ph_, heu_):
            out = E.cvrp_sample(ph_, heu_, self.demand, self.capacity, self.n_ants, seed=seed, offset=offset,
                                want_logp=require_prob)
            T = int(out["tmax"][0].item())      # one host sync per construction (the reference has one per step)
            gen.set_offset(offset + T * E.cvrp_step_offset_increment(self.problem_size, self.n_ants))
            return out["paths"][0, :T + 1].contiguous(), (out["logp"][0, :T].contiguous() if require_prob else None)

        if require_prob and torch.is_grad_enabled() and (ph.requires_grad or heu.requires_grad):
            return E.SampleLogProbs.apply(ph, heu, construct, self.demand, float(self.capacity))
        paths, logp = construct(ph.detach(), heu.detach())
        return (paths, logp) if require_prob else paths

    # ---- the step-level pieces of the reference's gen_path (cvrp/aco.py:167-205), for callers that drive the
    # construction themselves; gen_path above keeps all of this state inside one kernel ----------------------------
    def pick_move(self, prev, visit_mask, capacity_mask, require_prob):
        return E.pick_move_for(self, prev, visit_mask, capacity_mask, require_prob)

    def update_visit_mask(self, visit_mask, actions):
        '''cvrp/aco.py:176-180, in place: the chosen node becomes unavailable, the depot stays available unless the
        ant sits on it while customers remain.'''
        ants = torch.arange(self.n_ants, device=visit_mask.device)
        visit_mask[ants, actions] = 0
        customers_left = (visit_mask[:, 1:] != 0).any(dim=1)
        visit_mask[:, 0] = torch.where((actions == 0) & customers_left, 0.0, 1.0).to(visit_mask.dtype)
        return visit_mask

    def update_capacity_mask(self, cur_nodes, used_capacity):
        '''cvrp/aco.py:182-202: -> (used_capacity, capacity_mask [n_ants, problem_size]); the load resets at the depot
        (in place, like the reference), then customers whose demand exceeds the remaining capacity are masked.'''
        used_capacity[cur_nodes == 0] = 0
        used_capacity = used_capacity + self.demand[cur_nodes]
        remaining = (self.capacity - used_capacity).unsqueeze(-1)
        capacity_mask = (~(self.demand.unsqueeze(0) > remaining)).to(torch.float32)
        return used_capacity, capacity_mask

    def check_done(self, visit_mask, actions):
        '''cvrp/aco.py:204-205: every customer served and every ant back at the depot.'''
        return (visit_mask[:, 1:] == 0).all() and (actions == 0).all()

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
        self._runner = None

    @torch.no_grad()
    def run(self, n_iterations):
        '''cvrp/aco.py:72-104 (adaptive=False) on the device: no host round trip per construction step or per
        iteration; the generator offset (data dependent: one draw per step of the slowest ant) is read back once.'''
        if self.alpha != 1 or self.beta != 1:
            return self._run_stepwise(n_iterations)
        if self._runner is not None and self._runner_key != self._state_key():
            self._runner = None             # pheromone / heuristic / lowest_cost / ... were replaced or mutated since
        if self._runner is None:
            r = E.CvrpRunner(self.distances, self.demand, self.heuristic, self.pheromone, self.n_ants,
                             capacity=self.capacity, decay=self.decay, elitist=self.elitist, min_max=self.min_max,
                             ph_min=self.min if self.min_max else 0.0)
            if not isinstance(self.lowest_cost, float) or self.lowest_cost != float('inf'):
                r.lowest_cost.fill_(float(self.lowest_cost))
            if self.min_max and self.max is not None:
                r.ph_max.fill_(float(self.max))
            self._runner = r
        r = self._runner
        gen, seed, offset = generator_state(self.device)
        r.run(n_iterations, seed, [offset])
        gen.set_offset(int(r.offsets[0].item()))
        self.pheromone = r.pheromone[0].clone()
        self.lowest_cost = r.lowest_cost[0].clone()
        rows = int(r.shortest_rows[0].item())
        if rows > 0:
            self.shortest_path = r.shortest_path[0, :rows].clone()
        if self.min_max:
            self.max = r.ph_max[0].clone()
        self._runner_key = self._state_key()
        return self.lowest_cost

    def _state_key(self):
        return E.state_key(self.pheromone, self.heuristic, self.distances, self.demand, self.lowest_cost,
                           self.max if self.min_max else None, self.n_ants, self.decay, self.elitist, self.capacity)

    def _run_stepwise(self, n_iterations):
        for _ in range(n_iterations):
            paths = self.gen_path(require_prob=False)
            costs = self.gen_path_costs(paths)
            best = torch.argmin(costs)
            if costs[best] < self.lowest_cost:
                self.shortest_path, self.lowest_cost = paths[:, best], costs[best]
                if self.min_max:
                    new_max = self.problem_size / self.lowest_cost
                    if self.max is None:
                        self.pheromone = self.pheromone * (new_max / self.pheromone.max())
                    self.max = new_max
            self.update_pheronome(paths, costs)
        return self.lowest_cost
