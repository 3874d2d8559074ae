"""deepaco_knn_graph (csrc/knn_graph.cuh): the instance -> graph front end in one launch, against the reference's op chain
on the same GPU -- torch.norm(c[:, None] - c, dim=2, p=2) + diagonal (tsp/utils.py:4-14, cvrp/utils.py:18-22), torch.topk(...,
largest=False), edge_index (tsp/utils.py:24-34).  Bit-exact distances, identical neighbour lists (random coordinates: no
bit-equal distances inside a row, asserted), through the C ABI and through the utils modules that carry the reference's names."""
import pytest
import torch

from deepaco_b200 import _engine as E

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _reference_chain(c, k, diag):
    n = len(c)
    d = torch.norm(c[:, None] - c, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = diag
    if k == 0:
        return d, None, None, None
    tv, ti = torch.topk(d, k=k, dim=1, largest=False)
    ei = torch.stack([torch.repeat_interleave(torch.arange(n).to(ti.device), repeats=k), torch.flatten(ti)])
    return d, tv, ti, ei


@pytest.mark.parametrize("n,k", [(20, 10), (100, 20), (200, 20), (500, 50), (1000, 100), (33, 33), (2500, 20)])
def test_knn_graph_equals_reference_op_chain(n, k):
    g = torch.Generator().manual_seed(7 * n + k)
    c = torch.rand((n, 2), generator=g).to(DEV)
    d, tv, ti, ei = _reference_chain(c, k, 1e9)
    assert not (tv[:, 1:] == tv[:, :-1]).any()
    dist, idx, val, edges = E.knn_graph(coords=c, k=k, diag=1e9, want_edge_index=True)
    assert torch.equal(dist, d)
    assert torch.equal(val, tv) and torch.equal(idx.long(), ti) and torch.equal(edges, ei)
    # selection from a given matrix (what the batched network front end passes)
    _, idx2, val2, _ = E.knn_graph(distances=d, k=k)
    assert torch.equal(idx2, idx) and torch.equal(val2, val)


def test_batched_graphs_and_cvrp_diagonal():
    B, n, k = 64, 100, 20
    g = torch.Generator().manual_seed(3)
    c = torch.rand((B, n, 2), generator=g).to(DEV)
    dist, idx, val, edges = E.knn_graph(coords=c, k=k, want_edge_index=True)
    for b in (0, 17, 63):
        d, tv, ti, ei = _reference_chain(c[b], k, 1e9)
        assert torch.equal(dist[b], d) and torch.equal(val[b], tv) and torch.equal(idx[b].long(), ti) and torch.equal(edges[b], ei)
    d10, _, _, _ = _reference_chain(c[5], 0, 1e-10)
    assert torch.equal(E.knn_graph(coords=c[5], k=0, diag=1e-10)[0], d10)


def test_utils_modules_use_the_kernel_and_keep_the_reference_results():
    from deepaco_b200.cvrp import utils as cvrp_utils
    from deepaco_b200.tsp import utils as tsp_utils
    from deepaco_b200.tsp_nls import utils as nls_utils
    g = torch.Generator().manual_seed(11)
    c = torch.rand((150, 2), generator=g).to(DEV)
    d, tv, ti, ei = _reference_chain(c, 15, 1e9)
    launches = E.lib().deepaco_kernel_launches()
    pyg, dist = tsp_utils.gen_pyg_data(c, 15)
    assert E.lib().deepaco_kernel_launches() == launches + 1             # one launch, no eager op chain
    assert torch.equal(dist, d) and torch.equal(pyg.edge_index, ei) and torch.equal(pyg.edge_attr, tv.reshape(-1, 1))
    assert pyg.x is c
    pyg2, dist2 = nls_utils.gen_pyg_data(c, 15, start_node=3)
    assert torch.equal(dist2, d) and torch.equal(pyg2.edge_index, ei) and pyg2.x.shape == (150, 1) and float(pyg2.x[3]) == 1.0
    assert torch.equal(tsp_utils.gen_distance_matrix(c), d)
    assert torch.equal(cvrp_utils.gen_distance_matrix(c), _reference_chain(c, 0, 1e-10)[0])


def test_bad_arguments_are_reported():
    from deepaco_b200._lib import DeepAcoError
    c = torch.rand((10, 2), device=DEV)
    with pytest.raises(DeepAcoError):
        E.knn_graph(coords=c, k=11)
    with pytest.raises(DeepAcoError):
        E.knn_graph(coords=c, distances=torch.rand((10, 10), device=DEV), k=3)
    with pytest.raises(DeepAcoError):
        E.knn_graph(coords=torch.rand((10, 2)), k=3)                      # host tensor: no silent fallback behind the C ABI
