"""Instance -> graph helpers and dataset loaders with the reference's names (reference tsp/utils.py:4-54).
One-off set-up per instance."""
import torch

from ..net import Data


def gen_distance_matrix(tsp_coordinates):
    '''[n, 2] coordinates -> [n, n] Euclidean distances with 1e9 on the diagonal (tsp/utils.py:4-14).
    fp32 CUDA coordinates: one launch of deepaco_knn_graph (bit-identical to the op chain below, tests/test_gpu_graph.py);
    host (or non-fp32) coordinates: the reference's op chain where the tensor lives, as there.'''
    if tsp_coordinates.is_cuda and tsp_coordinates.dtype == torch.float32:
        from .. import _engine as E
        return E.knn_graph(coords=tsp_coordinates, k=0, diag=1e9)[0]
    n = len(tsp_coordinates)
    d = torch.norm(tsp_coordinates[:, None] - tsp_coordinates, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    return d


def knn_graph(tsp_coordinates, k_sparse, start_node=None):
    '''k-nearest-neighbour graph (tsp/utils.py:16-36; start_node one-hot node feature as tsp_nls/utils.py:37-43).
    CUDA coordinates: distances, topk and edge_index come out of one deepaco_knn_graph launch.'''
    n = len(tsp_coordinates)
    if tsp_coordinates.is_cuda and tsp_coordinates.dtype == torch.float32:
        from .. import _engine as E
        distances, _, near_d, edge_index = E.knn_graph(coords=tsp_coordinates, k=k_sparse, diag=1e9, want_edge_index=True)
    else:
        distances = gen_distance_matrix(tsp_coordinates)
        near_d, near_i = torch.topk(distances, k=k_sparse, dim=1, largest=False)
        src = torch.arange(n, device=near_i.device).repeat_interleave(k_sparse)
        edge_index = torch.stack([src, near_i.flatten()])
    if start_node is None:
        x = tsp_coordinates
    else:
        x = torch.zeros((n, 1), device=tsp_coordinates.device, dtype=tsp_coordinates.dtype)
        x[start_node, 0] = 1.0
    return Data(x=x, edge_index=edge_index, edge_attr=near_d.reshape(-1, 1)), distances


def gen_pyg_data(tsp_coordinates, k_sparse):
    '''tsp/utils.py:16-36 -> (pyg_data, distances).'''
    return knn_graph(tsp_coordinates, k_sparse)


def instances_to_graphs(coordinates, k_sparse, device, start_node=None):
    '''[(pyg_data, distances)] for a [count, n, 2] tensor of instances, built on `device`.'''
    return [knn_graph(instance.to(device), k_sparse, start_node) for instance in coordinates]


def load_val_dataset(n_node, k_sparse, device):
    '''tsp/utils.py:38-45: the shipped validation instances (path relative to the problem directory, as there).'''
    return instances_to_graphs(torch.load(f'../data/tsp/valDataset-{n_node}.pt'), k_sparse, device)


def load_test_dataset(n_node, k_sparse, device):
    '''tsp/utils.py:47-54.'''
    return instances_to_graphs(torch.load(f'../data/tsp/testDataset-{n_node}.pt'), k_sparse, device)
