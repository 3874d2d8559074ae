"""Roulette-wheel construction (deepaco_tsp_roulette_sample) against the reference's inference sampler
(tsp_nls/aco.py:260-297 `_inference_sample` / `inference_batch_sample`).  The reference draws from numba's private
generator, so parity is statistical: on the probability matrix of tests/golden/roulette_n12_stats.npz (generated from
the unmodified reference, 40,000 tours) the first-step distribution must fit the exact law, and the directed-edge usage
and last-node frequencies must be homogeneous with the reference's counts (chi-square, p > 1e-4).  Plus the
deterministic parts: valid permutations, forced chains, fixed / random start, the class surface."""
import numpy as np
import pytest
import torch
from scipy import stats

from deepaco_b200 import _engine as E

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _two_sample_chi2(a, b):
    """Homogeneity of two count vectors (cells with small expectation pooled)."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    keep = (a + b) >= 20
    a = np.append(a[keep], a[~keep].sum())
    b = np.append(b[keep], b[~keep].sum())
    if a[-1] + b[-1] == 0:
        a, b = a[:-1], b[:-1]
    return stats.chi2_contingency(np.stack([a, b]))[1]


def test_distribution_matches_reference_sampler(golden):
    g = golden("roulette_n12_stats")
    prob = torch.from_numpy(g["probmat"]).to(DEV)
    n, count = prob.shape[0], int(g["count"])
    paths, tours = E.tsp_roulette_sample(prob, count, start_node=0, seed=2024, offset=0, want_tours=True)
    t = tours.cpu().numpy().astype(np.int64)
    assert np.array_equal(paths.cpu().numpy().T, t)
    assert (t[:, 0] == 0).all() and (np.sort(t, axis=1) == np.arange(n)).all()
    # first step: exact law prob[0] / sum (goodness of fit), for us and -- as a sanity check of the fixture -- for the reference
    exact = g["probmat"][0].astype(np.float64) / g["probmat"][0].astype(np.float64).sum()
    first = np.bincount(t[:, 1], minlength=n)
    for counts in (first, g["first"]):
        p = stats.chisquare(counts[1:], exact[1:] * count)[1]
        assert p > 1e-4, (p, counts)
    # later steps depend on the visited set: compare with the reference's empirical counts
    edges = np.zeros((n, n), dtype=np.int64)
    np.add.at(edges, (t[:, :-1].ravel(), t[:, 1:].ravel()), 1)
    assert _two_sample_chi2(edges, g["edges"]) > 1e-4
    assert _two_sample_chi2(np.bincount(t[:, -1], minlength=n), g["last"]) > 1e-4


def test_forced_chain_and_random_start():
    n = 70
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(3))
    prob = torch.zeros(n, n)
    prob[perm, torch.roll(perm, -1)] = 1.0                      # exactly one successor with mass: the cycle `perm`
    prob = prob.to(DEV)
    _, tours = E.tsp_roulette_sample(prob, 33, start_node=int(perm[0]), seed=1, want_paths=False, want_tours=True)
    assert torch.equal(tours.cpu().to(torch.int64), perm.expand(33, n))
    _, tours = E.tsp_roulette_sample(prob, 4000, start_node=-1, seed=5, want_paths=False, want_tours=True)
    t = tours.cpu().to(torch.int64)
    pos = torch.argsort(perm)
    assert torch.equal(t, perm[(pos[t[:, :1]] + torch.arange(n)) % n])          # the cycle from whichever start was drawn
    starts = np.bincount(t[:, 0].numpy(), minlength=n)
    assert stats.chisquare(starts)[1] > 1e-4                                    # uniform start


@pytest.mark.parametrize("n", [33, 200, 500, 1000])
def test_valid_tours_at_every_size_and_batched(n):
    g = torch.Generator().manual_seed(n)
    prob = (torch.rand((2, n, n), generator=g) ** 4 + 1e-6).to(DEV)
    _, tours = E.tsp_roulette_sample(prob, 40, start_node=0, seed=9, offsets=[0, 4096], want_paths=False, want_tours=True)
    t = tours.cpu().to(torch.int64)
    assert torch.equal(torch.sort(t, dim=2).values, torch.arange(n).expand(2, 40, n))
    assert not torch.equal(t[0], t[1])
    again = E.tsp_roulette_sample(prob, 40, start_node=0, seed=9, offsets=[0, 4096], want_paths=False, want_tours=True)[1]
    assert torch.equal(again, tours)                                            # a function of (seed, offset)


def test_class_surface_inference_paths():
    """sample(inference=True), run(.., inference=True) and the module function inference_batch_sample of tsp_nls."""
    from deepaco_b200.tsp_nls import aco as A
    n = 60
    torch.manual_seed(0)
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    aco = A.ACO(d, n_ants=24, device=DEV, local_search="2opt")
    costs, logp, paths = aco.sample(inference=True)
    assert logp is None and paths.shape == (n, 24) and (paths[0] == 0).all()
    assert torch.equal(torch.sort(paths, dim=0).values, torch.arange(n, device=DEV)[:, None].expand(n, 24))
    assert torch.allclose(costs, aco.gen_path_costs(paths))
    off = torch.cuda.default_generators[0].get_offset()
    low = aco.run(3, inference=True)
    assert isinstance(low, float) and low < float(costs.min()) + 1e-6          # 2-opt to convergence on top of the sampler
    assert torch.cuda.default_generators[0].get_offset() == off + 3 * E.tsp_roulette_offset_increment(n, 24)
    assert aco.distances_numpy.dtype == np.float32 and aco.heuristic_numpy.shape == (n, n)
    routes = A.inference_batch_sample((1.0 / d).cpu().numpy(), count=5, startnode=None)
    assert routes.dtype == np.uint16 and routes.shape == (5, n) and (np.sort(routes, axis=1) == np.arange(n)).all()


def test_sparsify_matches_reference_ops():
    """ACO.sparsify(k) (tsp/aco.py:51-67): the k nearest per row keep 1 / distance, everything else 1 / 1e10."""
    from deepaco_b200.tsp.aco import ACO
    from deepaco_b200.tsp_nls.aco import ACO as NlsACO
    n, k = 50, 7
    torch.manual_seed(4)
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    # the reference's statements, verbatim semantics
    _, topk = torch.topk(d, k=k, dim=1, largest=False)
    u = torch.repeat_interleave(torch.arange(n, device=DEV), repeats=k)
    v = torch.flatten(topk)
    sparse = torch.ones_like(d) * 1e10
    sparse[u, v] = d[u, v]
    want = 1 / sparse
    for cls in (ACO, NlsACO):
        aco = cls(d, n_ants=4, device=DEV)
        assert torch.equal(aco.heuristic, 1 / d)
        assert aco.sparsify(k) is None
        assert torch.equal(aco.heuristic, want)
        assert int((aco.heuristic > 1e-9).sum()) == n * k
