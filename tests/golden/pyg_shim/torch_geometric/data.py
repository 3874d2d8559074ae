"""torch_geometric.data stand-in: `Data` is used as an attribute bag only (`tsp/net.py:85`)."""


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if hasattr(v, "to"):
                setattr(self, k, v.to(device))
        return self
