"""tsp_nls/utils.py:5-45 equivalents (same graph construction, optional one-hot start node feature)."""
from ..tsp.utils import gen_distance_matrix, gen_pyg_data  # noqa: F401
