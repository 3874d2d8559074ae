"""ORACLE (test infrastructure, never imported by the product path).

Op-for-op restatement of the reference ACO hot path in plain torch tensor ops, device agnostic.
It issues the same ATen calls in the same order as the reference (`Categorical(...).sample()`,
`torch.randint`, `index_put`, `sum`), so under one `torch.manual_seed` it consumes the global
generator identically and produces identical tensors on the same device.  Uses:

  * pinned against golden vectors produced by the *unmodified* reference on CPU
    (`tests/golden/make_golden.py` -> `tests/golden/*.npz`, checked in `tests/test_oracle_golden.py`);
  * same-device stream parity for the CUDA kernels (run with device='cuda' on the B200 box, where
    /root/reference does not exist);
  * the `cpu_baseline` / `--impl reference` leg of `bench.py` (device='cpu', all host threads).

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import torch
from torch.distributions import Categorical


# ----------------------------------------------------------------------------------------------
# TSP  (tsp/aco.py) and fixed-start TSP (tsp_nls/aco.py)
# ----------------------------------------------------------------------------------------------
def tsp_gen_path(pheromone, heuristic, n_ants, alpha=1, beta=1, require_prob=False,
                 nls_variant=False, noise_log=None):
    """tsp/aco.py:134-163 (+ pick_move :165-177); nls_variant=True follows tsp_nls/aco.py:184-220
    (start node 0, prob_mat built once, explicit pre-normalisation, validate_args=False).

    noise_log: optional list; when given, sampling is done through the identity
    Categorical(p).sample() == argmax((p/sum p)/q), q = empty_like(p).exponential_(1)
    (torch multinomial n_sample==1 fast path) and every q is appended -- same generator consumption.
    """
    dev = pheromone.device
    n = pheromone.shape[0]
    rows = torch.arange(n_ants, device=dev)
    if nls_variant:
        cur = torch.zeros((n_ants,), dtype=torch.long, device=dev)          # tsp_nls/aco.py:191
        prob_mat = (pheromone ** alpha) * (heuristic ** beta)                 # tsp_nls/aco.py:195
    else:
        cur = torch.randint(low=0, high=n, size=(n_ants,), device=dev)       # tsp/aco.py:141
    alive = torch.ones((n_ants, n), device=dev)
    alive[rows, cur] = 0
    visited_order = [cur]
    logps = []
    for _ in range(n - 1):
        if nls_variant:
            w = prob_mat[cur] * alive                                         # tsp_nls/aco.py:205
            w = w / w.sum(dim=-1, keepdim=True)                               # :206
            cat = Categorical(w, validate_args=False)                         # :207
        else:
            w = (pheromone[cur] ** alpha) * (heuristic[cur] ** beta) * alive  # tsp/aco.py:171-173
            cat = Categorical(w)                                              # :174
        if noise_log is None:
            nxt = cat.sample()                                                # :175
        else:
            q = torch.empty_like(cat.probs).exponential_(1)
            noise_log.append(q)
            nxt = torch.argmax(cat.probs / q, dim=-1)
        visited_order.append(nxt)
        if require_prob:
            logps.append(cat.log_prob(nxt))                                   # :176
            alive = alive.clone()                                             # :156
        cur = nxt
        alive[rows, nxt] = 0                                                  # :158
    paths = torch.stack(visited_order)
    if require_prob:
        return paths, torch.stack(logps)
    return paths


def tsp_path_costs(distances, paths):
    """tsp/aco.py:120-132: closed-tour length, sum over dist[u_k, u_{k-1}]."""
    u = paths.T
    v = torch.roll(u, shifts=1, dims=1)
    return torch.sum(distances[u, v], dim=1)


def tsp_update_pheromone(pheromone, paths, costs, decay=0.9, elitist=False, min_max=False,
                         ph_min=None, ph_max=None):
    """tsp/aco.py:94-118: evaporate, then per-ant symmetric deposit in ant order."""
    pheromone = pheromone * decay                                             # :101
    if elitist:
        best_cost, best_idx = costs.min(dim=0)                                # :104
        tour = paths[:, best_idx]
        prev = torch.roll(tour, shifts=1)
        pheromone[tour, prev] += 1.0 / best_cost                              # :106
        pheromone[prev, tour] += 1.0 / best_cost                              # :107
    else:
        for a in range(paths.shape[1]):                                       # :110
            tour = paths[:, a]
            prev = torch.roll(tour, shifts=1)
            w = 1.0 / costs[a]
            pheromone[tour, prev] += w                                        # :113
            pheromone[prev, tour] += w                                        # :114
    if min_max:
        pheromone[(pheromone > 1e-9) * (pheromone) < ph_min] = ph_min          # :117
        pheromone[pheromone > ph_max] = ph_max                                # :118
    return pheromone


class TspColony:
    """State machine of tsp/aco.py:74-92 (`run`) built from the functions above."""

    def __init__(self, distances, n_ants, heuristic=None, pheromone=None, decay=0.9, alpha=1, beta=1,
                 elitist=False, min_max=False, ph_min=None, nls_variant=False):
        self.distances = distances
        self.n = distances.shape[0]
        self.n_ants = n_ants
        self.decay, self.alpha, self.beta = decay, alpha, beta
        self.elitist, self.min_max = elitist, min_max
        self.nls_variant = nls_variant
        if min_max:
            self.ph_min = 0.1 if ph_min is None else ph_min                   # tsp/aco.py:29-35
            self.ph_max = None
        else:
            self.ph_min = self.ph_max = None
        if pheromone is None:
            pheromone = torch.ones_like(distances)
            if min_max:
                pheromone = pheromone * self.ph_min
        self.pheromone = pheromone
        self.heuristic = 1 / distances if heuristic is None else heuristic    # tsp/aco.py:44
        self.lowest_cost = float("inf")
        self.shortest_path = None

    @torch.no_grad()
    def run(self, n_iterations):
        for _ in range(n_iterations):
            paths = tsp_gen_path(self.pheromone, self.heuristic, self.n_ants, self.alpha, self.beta,
                                 nls_variant=self.nls_variant)
            costs = tsp_path_costs(self.distances, paths)
            best_cost, best_idx = costs.min(dim=0)
            if best_cost < self.lowest_cost:                                  # tsp/aco.py:81
                self.shortest_path = paths[:, best_idx]
                self.lowest_cost = best_cost
                if self.min_max:                                              # :84-88
                    new_max = self.n / self.lowest_cost
                    if self.ph_max is None:
                        self.pheromone *= new_max / self.pheromone.max()
                    self.ph_max = new_max
            self.pheromone = tsp_update_pheromone(self.pheromone, paths, costs, self.decay, self.elitist,
                                                  self.min_max, self.ph_min, self.ph_max)
        return self.lowest_cost


# ----------------------------------------------------------------------------------------------
# CVRP  (cvrp/aco.py:1-205)
# ----------------------------------------------------------------------------------------------
def _cvrp_visit_rule(alive, nxt, rows):
    """cvrp/aco.py:176-180."""
    alive[rows, nxt] = 0
    alive[:, 0] = 1
    alive[(nxt == 0) * (alive[:, 1:] != 0).any(dim=1), 0] = 0
    return alive


def _cvrp_capacity_rule(cur, used, demand, capacity, n_ants, n):
    """cvrp/aco.py:182-202."""
    cap_ok = torch.ones((n_ants, n), device=cur.device)
    used[cur == 0] = 0                                                        # :194
    used = used + demand[cur]                                                 # :195
    remaining = capacity - used
    remaining_rep = remaining.unsqueeze(-1).repeat(1, n)
    demand_rep = demand.unsqueeze(0).repeat(n_ants, 1)
    cap_ok[demand_rep > remaining_rep] = 0                                    # :200
    return used, cap_ok


def cvrp_gen_path(pheromone, heuristic, demand, capacity, n_ants, alpha=1, beta=1, require_prob=False,
                  noise_log=None):
    """cvrp/aco.py:138-165 (+ pick_move :167-174, check_done :204-205)."""
    dev = pheromone.device
    n = pheromone.shape[0]
    rows = torch.arange(n_ants, device=dev)
    cur = torch.zeros((n_ants,), dtype=torch.long, device=dev)
    alive = torch.ones((n_ants, n), device=dev)
    alive = _cvrp_visit_rule(alive, cur, rows)
    used = torch.zeros((n_ants,), device=dev)
    used, cap_ok = _cvrp_capacity_rule(cur, used, demand, capacity, n_ants, n)
    seq = [cur]
    logps = []

    def finished():
        return (alive[:, 1:] == 0).all() and (cur == 0).all()                 # :205

    while not finished():
        w = (pheromone[cur] ** alpha) * (heuristic[cur] ** beta) * alive * cap_ok   # :170
        cat = Categorical(w)
        if noise_log is None:
            nxt = cat.sample()
        else:
            q = torch.empty_like(cat.probs).exponential_(1)
            noise_log.append(q)
            nxt = torch.argmax(cat.probs / q, dim=-1)
        seq.append(nxt)
        if require_prob:
            logps.append(cat.log_prob(nxt))
            alive = alive.clone()
        cur = nxt
        alive = _cvrp_visit_rule(alive, cur, rows)
        used, cap_ok = _cvrp_capacity_rule(cur, used, demand, capacity, n_ants, n)
    paths = torch.stack(seq)
    if require_prob:
        return paths, torch.stack(logps)
    return paths


def cvrp_path_costs(distances, paths):
    """cvrp/aco.py:132-136: open path, sum dist[u_k, u_{k+1}], k < T."""
    u = paths.permute(1, 0)
    v = torch.roll(u, shifts=-1, dims=1)
    return torch.sum(distances[u[:, :-1], v[:, :-1]], dim=1)


def cvrp_update_pheromone(pheromone, paths, costs, decay=0.9, elitist=False, min_max=False,
                          ph_min=None, ph_max=None):
    """cvrp/aco.py:106-130: evaporate, one-directional deposit, 1e-10 floor."""
    pheromone = pheromone * decay
    if elitist:
        best_cost, best_idx = costs.min(dim=0)
        tour = paths[:, best_idx]
        pheromone[tour[:-1], torch.roll(tour, shifts=-1)[:-1]] += 1.0 / best_cost
    else:
        for a in range(paths.shape[1]):
            tour = paths[:, a]
            pheromone[tour[:-1], torch.roll(tour, shifts=-1)[:-1]] += 1.0 / costs[a]
    if min_max:
        pheromone[(pheromone > 1e-9) * (pheromone) < ph_min] = ph_min
        pheromone[pheromone > ph_max] = ph_max
    pheromone[pheromone < 1e-10] = 1e-10                                      # :130
    return pheromone


class CvrpColony:
    """cvrp/aco.py:72-104 with adaptive=False (the adaptive branch is out of scope)."""

    def __init__(self, distances, demand, n_ants, heuristic=None, pheromone=None, decay=0.9,
                 alpha=1, beta=1, elitist=False, capacity=50):
        self.distances, self.demand, self.capacity = distances, demand, capacity
        self.n = distances.shape[0]
        self.n_ants = n_ants
        self.decay, self.alpha, self.beta, self.elitist = decay, alpha, beta, elitist
        self.pheromone = torch.ones_like(distances) if pheromone is None else pheromone
        self.heuristic = 1 / distances if heuristic is None else heuristic
        self.lowest_cost = float("inf")
        self.shortest_path = None

    @torch.no_grad()
    def run(self, n_iterations):
        for _ in range(n_iterations):
            paths = cvrp_gen_path(self.pheromone, self.heuristic, self.demand, self.capacity, self.n_ants,
                                  self.alpha, self.beta)
            costs = cvrp_path_costs(self.distances, paths)
            best_cost, best_idx = costs.min(dim=0)
            if best_cost < self.lowest_cost:
                self.shortest_path = paths[:, best_idx]
                self.lowest_cost = best_cost
            self.pheromone = cvrp_update_pheromone(self.pheromone, paths, costs, self.decay, self.elitist)
        return self.lowest_cost
