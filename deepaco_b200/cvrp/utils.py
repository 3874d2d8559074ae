"""Instance helpers and the test-set loader with the reference's names (reference cvrp/utils.py:4-40)."""
import torch

from ..net import Data

CAPACITY = 50
DEMAND_LOW = 1
DEMAND_HIGH = 9
DEPOT_COOR = [0.5, 0.5]


def gen_distance_matrix(tsp_coordinates):
    '''cvrp/utils.py:18-22: Euclidean distances with 1e-10 on the diagonal.  CUDA coordinates: one deepaco_knn_graph launch
    (same bits); host coordinates: the op chain on the host, as in the reference.'''
    if tsp_coordinates.is_cuda and tsp_coordinates.dtype == torch.float32:
        from .. import _engine as E
        return E.knn_graph(coords=tsp_coordinates, k=0, diag=1e-10)[0]
    n = len(tsp_coordinates)
    d = torch.norm(tsp_coordinates[:, None] - tsp_coordinates, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e-10      # cvrp/utils.py:21
    return d


def gen_instance(n, device):
    '''Random instance (cvrp/utils.py:9-16): -> (demands [n+1] with depot 0 first, distances [n+1, n+1]).'''
    locations = torch.rand(size=(n, 2), device=device)
    demands = torch.randint(low=DEMAND_LOW, high=DEMAND_HIGH + 1, size=(n,), device=device)
    all_locations = torch.cat((torch.tensor([DEPOT_COOR], device=device), locations), dim=0)
    all_demands = torch.cat((torch.zeros((1,), device=device), demands))
    return all_demands, gen_distance_matrix(all_locations)


def gen_pyg_data(demands, distances, device):
    '''Dense graph (cvrp/utils.py:24-33): edge e has src = e mod N, dst = e div N, attr = distances.flatten()[e].'''
    n = demands.size(0)
    nodes = torch.arange(n, device=device)
    edge_index = torch.stack((nodes.repeat(n), torch.repeat_interleave(nodes, n)))
    return Data(x=demands.unsqueeze(1), edge_attr=distances.reshape((n * n, 1)), edge_index=edge_index)


def load_test_dataset(problem_size, device):
    '''cvrp/utils.py:35-40: rows of the saved [count, n+2, n+1] tensor -> [(demands [n+1], distances [n+1, n+1])]
    (the file is written by the reference's own `python utils.py`, seed 123456; path relative to the repo root).'''
    dataset = torch.load(f'./data/cvrp/testDataset-{problem_size}.pt', map_location=device)
    return [(inst[0], inst[1:]) for inst in dataset]
