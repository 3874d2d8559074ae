"""Dataset loaders of the utils modules (reference tsp/utils.py:38-54, tsp_nls/utils.py:47-70, cvrp/utils.py:35-40):
file names, relative locations and the structure of what they return.  CPU tensors: the loaders are set-up code."""
import os

import torch


def test_tsp_and_tsp_nls_loaders(tmp_path, monkeypatch):
    from deepaco_b200.tsp import utils as T
    from deepaco_b200.tsp_nls import utils as N
    (tmp_path / "data" / "tsp").mkdir(parents=True)
    (tmp_path / "tsp").mkdir()
    torch.manual_seed(0)
    coords = torch.rand(3, 12, 2)
    torch.save(coords, tmp_path / "data" / "tsp" / "valDataset-12.pt")
    torch.save(coords, tmp_path / "data" / "tsp" / "testDataset-12.pt")
    monkeypatch.chdir(tmp_path / "tsp")                      # the reference drivers run from the problem directory
    for loader in (T.load_val_dataset, T.load_test_dataset):
        items = loader(12, 4, "cpu")
        assert len(items) == 3
        pyg, dist = items[1]
        assert torch.equal(pyg.x, coords[1]) and pyg.edge_index.shape == (2, 48) and pyg.edge_attr.shape == (48, 1)
        assert dist.shape == (12, 12) and float(dist[0, 0]) == 1e9
        want = torch.norm(coords[1][:, None] - coords[1], dim=2, p=2)
        assert torch.equal(dist[0, 1:], want[0, 1:])
        # k nearest by distance, grouped by source node
        assert torch.equal(pyg.edge_index[0], torch.arange(12).repeat_interleave(4))
        assert torch.equal(pyg.edge_attr.view(12, 4), torch.topk(dist, 4, dim=1, largest=False).values)
    items = N.load_test_dataset(12, 4, "cpu", start_node=0)
    assert items[0][0].x.shape == (12, 1) and float(items[0][0].x[0, 0]) == 1.0 and float(items[0][0].x.sum()) == 1.0
    other = tmp_path / "elsewhere.pt"
    torch.save(coords[:2], other)
    assert len(N.load_test_dataset(12, 4, "cpu", filename=str(other))) == 2
    # the tsp_nls validation set is created on first use (50 instances) and reused afterwards
    made = N.load_val_dataset(7, 3, "cpu", start_node=0)
    assert len(made) == 50 and os.path.isfile(tmp_path / "data" / "tsp" / "valDataset-7.pt")
    again = N.load_val_dataset(7, 3, "cpu", start_node=0)
    assert torch.equal(made[5][1], again[5][1])


def test_cvrp_loader(tmp_path, monkeypatch):
    from deepaco_b200.cvrp import utils as C
    (tmp_path / "data" / "cvrp").mkdir(parents=True)
    torch.manual_seed(1)
    insts = []
    for _ in range(4):                                       # cvrp/utils.py:46-52 layout: demands row on top of the matrix
        demands, dist = C.gen_instance(9, "cpu")
        insts.append(torch.cat((demands.unsqueeze(0), dist), dim=0))
    torch.save(torch.stack(insts), tmp_path / "data" / "cvrp" / "testDataset-9.pt")
    monkeypatch.chdir(tmp_path)                              # cvrp/test.py runs from the repository root
    items = C.load_test_dataset(9, "cpu")
    assert len(items) == 4
    demands, dist = items[2]
    assert demands.shape == (10,) and dist.shape == (10, 10) and float(demands[0]) == 0.0
    assert torch.equal(torch.cat((demands.unsqueeze(0), dist)), insts[2])
    assert dist[3, 3] == torch.tensor(1e-10)


def test_complete_graph_arrays_written_directly_equal_the_sorted_edge_list():
    """deepaco_b200.net.dense_graph (the batched CVRP front end) == train_graph of the edge list cvrp/utils.py builds."""
    from deepaco_b200 import net as N
    from deepaco_b200.cvrp import utils as C
    torch.manual_seed(4)
    insts = [C.gen_instance(8, "cpu") for _ in range(3)]
    dist = torch.stack([d for _, d in insts])
    got = N.dense_graph(dist)
    for b, (demands, d) in enumerate(insts):
        pyg = C.gen_pyg_data(demands, d, "cpu")
        want = N.train_graph(pyg.edge_index[None], pyg.edge_attr[None], 9, backward=False)
        for k in ("row_ptr", "src", "dst", "attr", "order"):
            assert torch.equal(got[k][b], want[k][0]), k
    assert got["E"] == 81 and got["n"] == 9 and got["B"] == 3


def test_graph_cache_on_data_objects_follows_the_tensors(monkeypatch):
    """Net.forward remembers the sorted edge arrays on this package's Data objects; the entry must be dropped when the
    edge tensors are replaced or written, extended when the backward arrays are needed, and never put on foreign objects."""
    from deepaco_b200 import net as N
    from deepaco_b200.tsp import utils as T
    torch.manual_seed(0)
    pyg, _ = T.gen_pyg_data(torch.rand(10, 2), 3)
    g1 = N._graph_of(pyg, backward=False)
    assert N._graph_of(pyg, backward=False) is g1 and g1["col_ptr"] is None
    g2 = N._graph_of(pyg, backward=True)
    assert g2 is not g1 and g2["col_ptr"] is not None and N._graph_of(pyg, backward=False) is g2
    pyg.edge_attr.mul_(2.0)                                  # in-place write -> version bump
    g3 = N._graph_of(pyg, backward=False)
    assert g3 is not g2 and torch.equal(g3["attr"], 2 * g2["attr"])
    pyg.edge_index = pyg.edge_index.flip(1).contiguous()     # replaced tensor
    g4 = N._graph_of(pyg, backward=False)
    assert g4 is not g3
    moved = pyg.to("cpu")
    assert "_deepaco_graph" not in vars(moved)

    class Foreign:
        pass
    f = Foreign()
    f.x, f.edge_index, f.edge_attr = pyg.x, pyg.edge_index, pyg.edge_attr
    N._graph_of(f, backward=False)
    assert "_deepaco_graph" not in vars(f)
