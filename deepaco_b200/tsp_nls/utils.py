"""tsp_nls/utils.py:5-70 equivalents: same graph construction as tsp/ with the optional one-hot start-node feature,
and the dataset loaders (the validation set is created on first use, as the reference does)."""
import os

import torch

from ..tsp.utils import gen_distance_matrix, instances_to_graphs, knn_graph  # noqa: F401


def gen_pyg_data(tsp_coordinates, k_sparse, start_node=None):
    '''tsp_nls/utils.py:18-45 -> (pyg_data, distances).'''
    return knn_graph(tsp_coordinates, k_sparse, start_node)


def load_val_dataset(n_node, k_sparse, device, start_node=None):
    '''tsp_nls/utils.py:47-59: 50 random instances saved next to the shipped data sets if the file is missing.'''
    path = f'../data/tsp/valDataset-{n_node}.pt'
    if os.path.isfile(path):
        coordinates = torch.load(path)
    else:
        coordinates = torch.rand((50, n_node, 2))
        torch.save(coordinates, path)
    return instances_to_graphs(coordinates, k_sparse, device, start_node)


def load_test_dataset(n_node, k_sparse, device, start_node=None, filename=None):
    '''tsp_nls/utils.py:61-70.'''
    return instances_to_graphs(torch.load(filename or f'../data/tsp/testDataset-{n_node}.pt'), k_sparse, device, start_node)
