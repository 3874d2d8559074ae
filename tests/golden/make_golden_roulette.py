"""Golden statistics of the reference's inference sampler (tsp_nls/aco.py:260-297 `inference_batch_sample`, numba):
first-step and directed-edge usage counts over many tours on a fixed probability matrix.  Its RNG is numba's private
generator, so the fixture pins DISTRIBUTIONS, not tours.

    python tests/golden/make_golden_roulette.py      ->  tests/golden/roulette_n12_stats.npz     (build container only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_ref  # noqa: E402


def main():
    aco = load_ref("tsp_nls", "aco")
    rng = np.random.default_rng(7)
    n, count = 12, 40000
    probmat = (rng.random((n, n)) ** 3 + 1e-3).astype(np.float32)      # skewed rows, like pheromone (.) heuristic
    np.fill_diagonal(probmat, 0.0)
    routes = aco.inference_batch_sample(probmat, count, 0).astype(np.int64)       # [count, n], start node 0
    assert (np.sort(routes, axis=1) == np.arange(n)).all()
    first = np.bincount(routes[:, 1], minlength=n)
    edges = np.zeros((n, n), dtype=np.int64)
    np.add.at(edges, (routes[:, :-1].ravel(), routes[:, 1:].ravel()), 1)
    last = np.bincount(routes[:, -1], minlength=n)
    np.savez_compressed(os.path.join(HERE, "roulette_n12_stats.npz"), probmat=probmat, count=count, first=first, edges=edges,
                        last=last)
    print("first-step counts", first, "\nexact", np.round(probmat[0] / probmat[0].sum() * count, 1))


if __name__ == "__main__":
    main()
