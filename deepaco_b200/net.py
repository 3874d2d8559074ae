"""Heuristic network `Net` with the reference's module tree and state_dict keys (reference tsp/net.py:8-102;
tsp_nls/net.py and cvrp/net.py differ only in `feats` and the unused `par_net_phe`), so the checkpoints in
pretrained/{tsp,tsp_nls,cvrp}/*.pt load unchanged.

eval mode  -> one launch of the sm_100a kernel deepaco_gnn_forward (csrc/gnn.cu) per batch of graphs.
train mode -> autograd-capable tensor ops (training of the network is listed under "next" in SURVEY.md 8f and
              is not yet native).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

UNITS, DEPTH = 32, 12


class _WrappedBatchNorm(nn.Module):
    """torch_geometric.nn.BatchNorm keeps its nn.BatchNorm1d under `.module` (checkpoint keys `*.module.*`)."""

    def __init__(self, channels):
        super().__init__()
        self.module = nn.BatchNorm1d(channels)

    def forward(self, t):
        return self.module(t)


class EmbNet(nn.Module):
    def __init__(self, depth=DEPTH, feats=2, units=UNITS, act_fn='silu', agg_fn='mean'):
        super().__init__()
        assert act_fn == 'silu' and agg_fn == 'mean' and units == UNITS and depth == DEPTH
        self.depth, self.feats, self.units = depth, feats, units
        mk = lambda: nn.ModuleList([nn.Linear(units, units) for _ in range(depth)])
        self.v_lin0 = nn.Linear(feats, units)
        self.v_lins1, self.v_lins2, self.v_lins3, self.v_lins4 = mk(), mk(), mk(), mk()
        self.v_bns = nn.ModuleList([_WrappedBatchNorm(units) for _ in range(depth)])
        self.e_lin0 = nn.Linear(1, units)
        self.e_lins0 = mk()
        self.e_bns = nn.ModuleList([_WrappedBatchNorm(units) for _ in range(depth)])

    def forward(self, x, edge_index, edge_attr):
        """Tensor-op formulation (used in train mode; tsp/net.py:27-45)."""
        src, dst = edge_index[0], edge_index[1]
        n = x.shape[0]
        deg = torch.zeros(n, device=x.device, dtype=x.dtype).index_add_(0, src, torch.ones_like(src, dtype=x.dtype))
        x = F.silu(self.v_lin0(x))
        w = F.silu(self.e_lin0(edge_attr))
        for i in range(self.depth):
            msg = torch.sigmoid(w) * self.v_lins2[i](x)[dst]
            agg = torch.zeros_like(x).index_add_(0, src, msg) / deg.clamp(min=1).unsqueeze(-1)
            x_new = x + F.silu(self.v_bns[i](self.v_lins1[i](x) + agg))
            w = w + F.silu(self.e_bns[i](self.e_lins0[i](w) + self.v_lins3[i](x)[src] + self.v_lins4[i](x)[dst]))
            x = x_new
        return w


class MLP(nn.Module):
    @property
    def device(self):
        return self._dummy.device

    def __init__(self, units_list, act_fn):
        super().__init__()
        self._dummy = nn.Parameter(torch.empty(0), requires_grad=False)
        self.units_list = units_list
        self.depth = len(units_list) - 1
        self.lins = nn.ModuleList([nn.Linear(units_list[i], units_list[i + 1]) for i in range(self.depth)])

    def forward(self, t):
        for i, lin in enumerate(self.lins):
            t = lin(t)
            t = F.silu(t) if i < self.depth - 1 else torch.sigmoid(t)
        return t


class ParNet(MLP):
    def __init__(self, depth=3, units=UNITS, preds=1, act_fn='silu'):
        self.units, self.preds = units, preds
        super().__init__([units] * depth + [preds], act_fn)

    def forward(self, t):
        return super().forward(t).squeeze(dim=-1)


def pack_weights(net: "Net") -> torch.Tensor:
    """Flat fp32 tensor in the layout csrc/gnn.cu expects:
    v_lin0 W[32][feats] b[32] | e_lin0 W[32] b[32] | 12 x { 4 x (W[32][32] b[32]) node linears 1..4 |
    e_lins0 W b | v_bn gamma beta mean invstd | e_bn gamma beta mean invstd } | head lin0 W b | lin1 W b | lin2 W[32] b[1]"""
    e = net.emb_net
    parts = [e.v_lin0.weight, e.v_lin0.bias, e.e_lin0.weight.reshape(-1), e.e_lin0.bias]

    def bn(m):
        m = m.module
        return [m.weight, m.bias, m.running_mean, torch.rsqrt(m.running_var + m.eps)]

    for i in range(e.depth):
        for lins in (e.v_lins1, e.v_lins2, e.v_lins3, e.v_lins4):
            parts += [lins[i].weight, lins[i].bias]
        parts += [e.e_lins0[i].weight, e.e_lins0[i].bias] + bn(e.v_bns[i]) + bn(e.e_bns[i])
    h = net.par_net_heu.lins
    parts += [h[0].weight, h[0].bias, h[1].weight, h[1].bias, h[2].weight.reshape(-1), h[2].bias]
    flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in parts]).contiguous()
    assert flat.numel() == lib().deepaco_gnn_weight_count(e.feats)
    return flat


def csr_by_source(edge_index, n_nodes):
    """(row_ptr int32 [n+1], order int32 [E]) grouping edges by edge_index[0] (stable)."""
    src = edge_index[0]
    order = torch.argsort(src, stable=True)
    counts = torch.bincount(src, minlength=n_nodes)
    row_ptr = torch.zeros(n_nodes + 1, dtype=torch.int32, device=src.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return row_ptr, order.to(torch.int32)


def _launch_gnn(weights, feats, xin, row_ptr, dst_s, attr_s, order, want_vec, dense_eps):
    B, n = xin.shape[0], xin.shape[1]
    E = dst_s.shape[1]
    dev = xin.device
    node_ws = torch.empty((B, n, 6 * UNITS), dtype=torch.float32, device=dev)
    edge_ws = torch.empty((B, E, UNITS), dtype=torch.float32, device=dev)
    out = torch.empty((B, E), dtype=torch.float32, device=dev) if want_vec else None
    dense = torch.empty((B, n, n), dtype=torch.float32, device=dev) if dense_eps is not None else None
    with torch.cuda.device(dev):
        check(lib().deepaco_gnn_forward(ptr(xin), ptr(row_ptr), ptr(dst_s), ptr(attr_s), ptr(order), ptr(weights), n, E, feats,
                                        B, ptr(node_ws), ptr(edge_ws), ptr(out), ptr(dense),
                                        float(dense_eps if dense_eps is not None else 0.0), stream_ptr(dev)), "deepaco_gnn_forward")
    return out, dense


def gnn_forward(weights, feats, x, edge_index, edge_attr, dense_eps=None):
    """deepaco_gnn_forward for one graph or a batch ([B, ...] tensors with identical n and E).
    Returns the edge vector; with dense_eps also the dense heuristic matrix Net.reshape(...) + dense_eps."""
    batched = x.dim() == 3
    if not batched:
        x, edge_index, edge_attr = x[None], edge_index[None], edge_attr[None]
    B, n = x.shape[0], x.shape[1]
    E = edge_index.shape[-1]
    _lib.require_cuda(x, "pyg.x")
    rps, orders = zip(*(csr_by_source(edge_index[b], n) for b in range(B)))
    row_ptr, order = torch.stack(rps).contiguous(), torch.stack(orders).contiguous()
    ol = order.long()
    dst_s = torch.gather(edge_index[:, 1], 1, ol).to(torch.int32).contiguous()
    attr_s = torch.gather(edge_attr.reshape(B, E).to(torch.float32), 1, ol).contiguous()
    out, dense = _launch_gnn(weights, feats, x.to(torch.float32).contiguous(), row_ptr, dst_s, attr_s, order, True, dense_eps)
    if dense_eps is None:
        return out if batched else out[0]
    return (out, dense) if batched else (out[0], dense[0])


def knn_heuristic_matrices(weights, feats, node_features, distances, k_sparse, eps=1e-10):
    """Batched instance -> graph -> network -> dense heuristic front end for k-nearest-neighbour TSP graphs
    (tsp/utils.py:16-36 + tsp/net.py:84-102 + the `+ EPS` of tsp/test.ipynb cell 1) without any per-instance
    Python: node_features [B, n, feats], distances [B, n, n] -> heuristic [B, n, n]."""
    B, n = distances.shape[0], distances.shape[1]
    dev = distances.device
    near_d, near_i = torch.topk(distances, k=k_sparse, dim=2, largest=False)       # sorted by source, constant degree
    E = n * k_sparse
    row_ptr = (torch.arange(n + 1, device=dev, dtype=torch.int32) * k_sparse).expand(B, n + 1).contiguous()
    order = torch.arange(E, device=dev, dtype=torch.int32).expand(B, E).contiguous()
    _, dense = _launch_gnn(weights, feats, node_features.to(torch.float32).contiguous(), row_ptr,
                           near_i.reshape(B, E).to(torch.int32).contiguous(), near_d.reshape(B, E).to(torch.float32).contiguous(),
                           order, False, eps)
    return dense


class Net(nn.Module):
    FEATS = 2          # tsp/net.py:9; subclasses in tsp_nls/ and cvrp/ use 1
    HAS_PHE_HEAD = True

    def __init__(self):
        super().__init__()
        self.emb_net = EmbNet(feats=self.FEATS)
        if self.HAS_PHE_HEAD:
            self.par_net_phe = ParNet()
        self.par_net_heu = ParNet()
        self._packed = None
        self._packed_key = None

    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.state_dict().values())
        if self._packed is None or key != self._packed_key:
            self._packed, self._packed_key = pack_weights(self), key
        return self._packed

    def forward(self, pyg):
        x, edge_index, edge_attr = pyg.x, pyg.edge_index, pyg.edge_attr
        if self.training or torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            return self.par_net_heu(self.emb_net(x, edge_index, edge_attr))
        return gnn_forward(self._weights(), self.FEATS, x, edge_index, edge_attr)

    @torch.no_grad()
    def heuristic_matrices(self, node_features, distances, k_sparse, eps=1e-10):
        """[B, n, n] heuristic matrices of a batch of k-NN TSP instances in one launch (eval mode)."""
        return knn_heuristic_matrices(self._weights(), self.FEATS, node_features, distances, k_sparse, eps)

    def freeze_gnn(self):
        for p in self.emb_net.parameters():
            p.requires_grad = False

    @staticmethod
    def reshape(pyg, vector):
        '''Edge vector -> dense [n, n] matrix, zero off-graph (tsp/net.py:95-102).'''
        n = pyg.x.shape[0]
        dense = torch.zeros((n, n), device=pyg.x.device, dtype=vector.dtype)
        dense[pyg.edge_index[0], pyg.edge_index[1]] = vector
        return dense


def load_npz_state_dict(path, device="cpu"):
    """State dict from the fixture format of tests/golden/make_golden.py (keys with '.' -> '__')."""
    z = np.load(path)
    return {k.replace("__", "."): torch.from_numpy(z[k]).to(device) for k in z.files}


class Data:
    """Attribute bag standing in for torch_geometric.data.Data: Net only reads .x, .edge_index, .edge_attr."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if hasattr(v, "to"):
                setattr(self, k, v.to(device))
        return self
