"""The construction step for ANY mask rule: the hook for the reference's other problem directories.

`op/aco.py:190-197` (`pick_node(mask, cur_node, require_prob)`), `sop/aco.py:158-170` (`pick_move(prev, mask1, mask2,
require_prob)`), `pctsp/`, `smtwtp/` ... all take the same step as `tsp/aco.py:165-177`:

    dist = pheromone[prev] ** alpha * heuristic[prev] ** beta * mask (* mask2);  Categorical(dist).sample() / .log_prob()

and differ only in how the caller updates its masks between steps (orienteering budget, precedence constraints,
prize thresholds).  `deepaco_pick_move` (csrc/pick_move.cuh) is that step with caller-held masks -- K1's mask functor
turned inside out: the rule stays in the caller's tensor code, the gather / product / normalise / draw / log-prob of the
step is one launch that consumes the default CUDA generator exactly as `Categorical.sample()` would, so a loop driven
through it reproduces the reference's actions bit for bit under the same seed (tests/test_gpu_pick_move.py)."""
from __future__ import annotations

import torch

from . import _engine as E
from ._lib import generator_state


def pick_move(pheromone, heuristic, prev, mask, mask2=None, *, alpha=1, beta=1, require_prob=False):
    """One construction step for `n_ants` ants: prev int64 [n_ants]; mask (and mask2) [n_ants, n] with 0 for inadmissible
    nodes -> (actions int64 [n_ants], log_probs [n_ants] | None).  Advances the default CUDA generator of the tensors'
    device like `Categorical(...).sample()`.  Differentiable log-probs (w.r.t. pheromone / heuristic) when they require
    grad, as in `ACO.pick_move` of the tsp / cvrp classes."""

    class _View:                                     # the attributes pick_move_for reads
        device = pheromone.device

        @staticmethod
        def _weights():
            ph = pheromone if alpha == 1 else pheromone ** alpha
            heu = heuristic if beta == 1 else heuristic ** beta
            return ph, heu

    return E.pick_move_for(_View, prev, mask, mask2, require_prob)


def generator_offset(device):
    """(seed, offset) of the default CUDA generator -- for callers that checkpoint / replay a construction."""
    _, seed, offset = generator_state(device)
    return seed, offset
