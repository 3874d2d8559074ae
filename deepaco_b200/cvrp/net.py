"""`Net` for cvrp/ (feats = 1: demand, reference cvrp/net.py:9; no par_net_phe)."""
from ..net import Data, EmbNet, MLP, Net as _Net, ParNet, load_npz_state_dict  # noqa: F401


class Net(_Net):
    FEATS = 1
    HAS_PHE_HEAD = False
