"""`Net` for tsp/ (feats = 2: node coordinates; has the unused par_net_phe head like reference tsp/net.py:78-83)."""
from ..net import Data, EmbNet, MLP, Net as _Net, ParNet, load_npz_state_dict  # noqa: F401


class Net(_Net):
    FEATS = 2
    HAS_PHE_HEAD = True
