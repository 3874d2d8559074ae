"""Heuristic network `Net` with the reference's module tree and state_dict keys (reference tsp/net.py:8-102;
tsp_nls/net.py and cvrp/net.py differ only in `feats` and the unused `par_net_phe`), so the checkpoints in
pretrained/{tsp,tsp_nls,cvrp}/*.pt load unchanged.

eval mode  -> one launch of the sm_100a kernel deepaco_gnn_forward (csrc/gnn.cu) per batch of graphs.
train mode -> deepaco_gnn_train_forward / deepaco_gnn_train_backward (csrc/gnn_train.cuh) behind a
              torch.autograd.Function: batch-statistics BatchNorm, analytic backward, running statistics updated
              as nn.BatchNorm1d does.  There is no tensor-op / CPU formulation in this package (the torch fp32
              restatement used by the tests lives in oracle/net_torch.py).
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
        raise _lib.DeepAcoError("EmbNet has no stand-alone forward in deepaco_b200: call Net.forward (fused kernels)")


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


def pack_weights(net: "Net", for_training: bool = False) -> torch.Tensor:
    """Flat fp32 tensor in the layout csrc/gnn.cu expects:
    v_lin0 W[32][feats] b[32] | e_lin0 W[32] b[32] | 12 x { 4 x (W[32][32] b[32]) node linears 1..4 |
    e_lins0 W b | v_bn gamma beta mean invstd | e_bn gamma beta mean invstd } | head lin0 W b | lin1 W b | lin2 W[32] b[1]
    for_training: built with autograd-tracked ops (torch.cat routes the packed gradient back to the parameters);
    the mean / invstd slots are zeros (training mode uses batch statistics)."""
    e = net.emb_net
    parts = [e.v_lin0.weight, e.v_lin0.bias, e.e_lin0.weight.reshape(-1), e.e_lin0.bias]
    zeros = torch.zeros(2 * UNITS, dtype=torch.float32, device=e.v_lin0.weight.device) if for_training else None

    def bn(m):
        m = m.module
        if for_training:
            return [m.weight, m.bias, zeros]
        return [m.weight, m.bias, m.running_mean, torch.rsqrt(m.running_var + m.eps)]

    for i in range(e.depth):
        layer = []
        for lins in (e.v_lins1, e.v_lins2, e.v_lins3, e.v_lins4):
            layer += [lins[i].weight, lins[i].bias]
        layer += [e.e_lins0[i].weight, e.e_lins0[i].bias] + bn(e.v_bns[i]) + bn(e.e_bns[i])
        if for_training and i == e.depth - 1:
            # the last layer's node update never reaches the output (EmbNet.forward returns w, tsp/net.py:45): autograd
            # leaves .grad of v_lins1/2[11] and v_bns[11] at None in the reference, so they must not receive zeros here
            # (an optimiser with weight decay treats None and 0 differently)
            dead = {id(e.v_lins1[i].weight), id(e.v_lins1[i].bias), id(e.v_lins2[i].weight), id(e.v_lins2[i].bias),
                    id(e.v_bns[i].module.weight), id(e.v_bns[i].module.bias)}
            layer = [t.detach() if id(t) in dead else t for t in layer]
        parts += layer
    h = net.par_net_heu.lins
    parts += [h[0].weight, h[0].bias, h[1].weight, h[1].bias, h[2].weight.reshape(-1), h[2].bias]
    if for_training:
        flat = torch.cat([t.reshape(-1).to(torch.float32) for t in parts])
    else:
        flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in parts]).contiguous()
    assert flat.numel() == weight_count(e.feats)
    return flat


def weight_count(feats):
    return UNITS * feats + 3 * UNITS + DEPTH * (5 * (UNITS * UNITS + UNITS) + 8 * UNITS) + 2 * (UNITS * UNITS + UNITS) + UNITS + 1


def _launch_gnn(weights, feats, xin, row_ptr, dst_s, attr_s, order, want_vec, dense_eps, src_s=None):
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
                                        float(dense_eps if dense_eps is not None else 0.0), ptr(src_s), stream_ptr(dev)),
              "deepaco_gnn_forward")
    return out, dense


def gnn_forward(weights, feats, x, edge_index, edge_attr, dense_eps=None, graph=None):
    """deepaco_gnn_forward for one graph or a batch ([B, ...] tensors with identical n and E).
    Returns the edge vector; with dense_eps also the dense heuristic matrix Net.reshape(...) + dense_eps.
    `graph`: train_graph(...) of these edges when the caller already has it."""
    batched = x.dim() == 3
    if not batched:
        x, edge_index, edge_attr = x[None], edge_index[None], edge_attr[None]
    B, n = x.shape[0], x.shape[1]
    E = edge_index.shape[-1]
    _lib.require_cuda(x, "pyg.x")
    if dense_eps is None:
        ctas = group_ctas(E, B)
        if ctas > 1:                          # few instances: several CTAs per graph instead of one (latency)
            out = gnn_forward_group(weights, feats, x, edge_index, edge_attr, ctas, graph=graph)
            return out if batched else out[0]
    if graph is None:                       # CSR by source for the whole batch at once (no per-instance Python)
        graph = train_graph(edge_index, edge_attr, n, backward=False)
    out, dense = _launch_gnn(weights, feats, x.to(torch.float32).contiguous(), graph["row_ptr"], graph["dst"], graph["attr"],
                             graph["order"], True, dense_eps, graph["src"])
    if dense_eps is None:
        return out if batched else out[0]
    return (out, dense) if batched else (out[0], dense[0])


def knn_heuristic_matrices(weights, feats, node_features, distances, k_sparse, eps=1e-10):
    """Batched instance -> graph -> network -> dense heuristic front end for k-nearest-neighbour TSP graphs
    (tsp/utils.py:16-36 + tsp/net.py:84-102 + the `+ EPS` of tsp/test.ipynb cell 1) without any per-instance
    Python: node_features [B, n, feats], distances [B, n, n] -> heuristic [B, n, n]."""
    from . import _engine
    B, n = distances.shape[0], distances.shape[1]
    dev = distances.device
    # torch.topk(distances, k, dim=2, largest=False) by deepaco_knn_graph: edges sorted by source, constant degree
    _, near_i, near_d, _ = _engine.knn_graph(distances=distances, k=k_sparse)
    E = n * k_sparse
    row_ptr = (torch.arange(n + 1, device=dev, dtype=torch.int32) * k_sparse).expand(B, n + 1).contiguous()
    order = torch.arange(E, device=dev, dtype=torch.int32).expand(B, E).contiguous()
    src_s = (order // k_sparse).contiguous()                                          # constant degree: source = e // k
    _, dense = _launch_gnn(weights, feats, node_features.to(torch.float32).contiguous(), row_ptr,
                           near_i.reshape(B, E), near_d.reshape(B, E),
                           order, False, eps, src_s)
    return dense


def dense_graph(distances):
    """train_graph(...) of the complete graph the CVRP drivers build (cvrp/utils.py:24-33: edge e = i * N + j has source
    j, destination i and attribute distances[i, j]) for a batch [B, N, N], written down directly: sorted by source the
    edges of source s are e = i * N + s, i = 0 .. N-1."""
    B, N = distances.shape[0], distances.shape[1]
    dev = distances.device
    E = N * N
    pos = torch.arange(E, device=dev, dtype=torch.int32)
    i32 = lambda t: t.to(torch.int32).expand(B, -1).contiguous()
    return {"row_ptr": i32((torch.arange(N + 1, device=dev) * N)[None]), "src": i32((pos // N)[None]), "dst": i32((pos % N)[None]),
            "attr": distances.to(torch.float32).transpose(1, 2).reshape(B, E).contiguous(),
            "order": i32(((pos % N) * N + pos // N)[None]), "col_ptr": None, "in_edges": None, "n": N, "E": E, "B": B}


def dense_heuristic_matrices(weights, feats, node_features, distances, eps=1e-10):
    """Batched instance -> complete graph -> network -> heuristic matrix front end for CVRP (cvrp/utils.py:24-33 +
    cvrp/test.py:17-19: heu_vec.reshape(N, N) + EPS) without per-instance Python: node_features [B, N, feats],
    distances [B, N, N] -> [B, N, N]."""
    B, N = distances.shape[0], distances.shape[1]
    g = dense_graph(distances)
    x = node_features.to(torch.float32).contiguous()
    ctas = group_ctas(g["E"], B)
    if ctas > 1:
        vec = gnn_forward_group(weights, feats, x, None, None, ctas, graph=g)
    else:
        vec, _ = _launch_gnn(weights, feats, x, g["row_ptr"], g["dst"], g["attr"], g["order"], True, None, g["src"])
    return vec.view(B, N, N) + eps


# ---------------------------------------------------------------------------------------------------------------
# training mode (deepaco_gnn_train_forward / _backward)
# ---------------------------------------------------------------------------------------------------------------
def train_graph(edge_index, edge_attr, n_nodes, backward=True):
    """Graph arrays of deepaco_gnn_train_args for edge_index [B, 2, E], edge_attr [B, E(,1)]: edges sorted by source
    (stable) with their CSR, and (backward=True) the same edges grouped by destination (stable) with their CSC."""
    B, E = edge_index.shape[0], edge_index.shape[-1]
    dev = edge_index.device
    src, dst = edge_index[:, 0], edge_index[:, 1]
    order = torch.argsort(src, dim=1, stable=True)
    src_s, dst_s = torch.gather(src, 1, order), torch.gather(dst, 1, order)
    attr_s = torch.gather(edge_attr.reshape(B, E).to(torch.float32), 1, order)

    def ptrs(keys):
        counts = torch.zeros((B, n_nodes), dtype=torch.int64, device=dev).scatter_add_(1, keys, torch.ones_like(keys))
        out = torch.zeros((B, n_nodes + 1), dtype=torch.int32, device=dev)
        out[:, 1:] = torch.cumsum(counts, 1)
        return out

    i32 = lambda t: t.to(torch.int32).contiguous()
    g = {"row_ptr": ptrs(src_s), "src": i32(src_s), "dst": i32(dst_s), "attr": attr_s.contiguous(), "order": i32(order),
         "col_ptr": None, "in_edges": None, "n": n_nodes, "E": E, "B": B}
    if backward:
        g["col_ptr"], g["in_edges"] = ptrs(dst_s), i32(torch.argsort(dst_s, dim=1, stable=True))
    return g


def train_buffers(B, n, E, device):
    """Activations the forward saves for the backward, BatchNorm batch statistics and scratch (sizes: include/deepaco_b200.h)."""
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
    return {"xs": f(B, DEPTH + 1, n, UNITS), "ws": f(B, DEPTH + 1, E, UNITS), "zv": f(B, DEPTH, n, UNITS),
            "ze": f(B, DEPTH, E, UNITS), "stats": f(B, DEPTH, 6, UNITS), "node_ws": f(B, n, 7 * UNITS),
            "edge_ws": f(B, E, 3 * UNITS), "red": f(B, 36, 64, 128),
            "sync_ws": torch.zeros(B, dtype=torch.int32, device=device)}


_EVAL_SCRATCH = {}        # (B, n, E, device, stream) -> buffers: one set per device AND stream (stream-ordered reuse)
_EVAL_SCRATCH_MAX = 8


def eval_buffers(B, n, E, device):
    """Scratch of deepaco_gnn_forward_group: two-layer ping-pong node / edge state, the four node linears, counters."""
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
    return {"xs": f(B, 2, n, UNITS), "ws": f(B, 2, E, UNITS), "node_ws": f(B, n, 4 * UNITS),
            "sync_ws": torch.zeros(B, dtype=torch.int32, device=device)}


def train_args(x, graph, weights, bufs, feats, ctas, bn_eps, heu_out=None, grad_heu=None, grad_weights=None):
    """Fill a deepaco_gnn_train_args struct; returns (struct, keep-alive list)."""
    p = lambda t: None if t is None else t.data_ptr()
    a = _lib.GnnTrainArgs(graph["n"], graph["E"], feats, graph["B"], ctas, bn_eps, p(x), p(graph["row_ptr"]), p(graph["src"]),
                          p(graph["dst"]), p(graph["attr"]), p(graph["order"]), p(graph["col_ptr"]), p(graph["in_edges"]),
                          p(weights), p(bufs["xs"]), p(bufs["ws"]), p(bufs.get("zv")), p(bufs.get("ze")), p(bufs.get("stats")),
                          p(bufs["node_ws"]), p(bufs.get("edge_ws")), p(bufs.get("red")), p(bufs["sync_ws"]), p(heu_out),
                          p(grad_heu), p(grad_weights))
    return a, [x, graph, weights, bufs, heu_out, grad_heu, grad_weights]


def group_ctas(n_edges, n_instances):
    """CTAs per graph for the eval-mode group forward: what the training kernels use for this graph size, halved until
    the whole batch fits the device twice over (296 CTAs); 1 means "use the one-CTA-per-instance kernel"."""
    ctas = default_train_ctas(n_edges)
    while ctas > 1 and ctas * n_instances > 296:
        ctas //= 2
    return ctas


def gnn_forward_group(weights, feats, x, edge_index, edge_attr, ctas, graph=None):
    """deepaco_gnn_forward_group: eval-mode Net.forward of [B, ...] graphs (identical n and E) by `ctas` CTAs per graph."""
    import ctypes
    B, n = x.shape[0], x.shape[1]
    if graph is None:
        graph = train_graph(edge_index, edge_attr, n, backward=False)
    dev = x.device
    key = (B, n, graph["E"], dev, torch.cuda.current_stream(dev).cuda_stream)
    bufs = _EVAL_SCRATCH.get(key)                # reused call after call on the SAME device and stream only
    if bufs is None:
        if len(_EVAL_SCRATCH) >= _EVAL_SCRATCH_MAX:
            _EVAL_SCRATCH.pop(next(iter(_EVAL_SCRATCH)))
        bufs = _EVAL_SCRATCH[key] = eval_buffers(B, n, graph["E"], dev)
    heu = torch.empty((B, graph["E"]), dtype=torch.float32, device=dev)
    a, keep = train_args(x.to(torch.float32).contiguous(), graph, weights, bufs, feats, ctas, 1e-5, heu_out=heu)
    with torch.cuda.device(dev):
        check(lib().deepaco_gnn_forward_group(ctypes.byref(a), stream_ptr(dev)), "deepaco_gnn_forward_group")
    return heu


def default_train_ctas(n_edges):
    """CTAs cooperating on one graph (DEEPACO_GNN_CTAS overrides): a thread-block cluster up to 8, a cooperative
    launch of 32 / 64 CTAs once the per-layer edge work outweighs the slower arrival-counter barrier."""
    import os
    env = os.environ.get("DEEPACO_GNN_CTAS")
    if env:
        return int(env)
    return 1 if n_edges < 512 else (8 if n_edges < 4096 else (32 if n_edges < 16384 else 64))


def _packed_entries(net):
    """[(tensor-or-None, numel)] in the packed weight order of pack_weights (None = the unused mean / invstd slots)."""
    e = net.emb_net
    out = [e.v_lin0.weight, e.v_lin0.bias, e.e_lin0.weight, e.e_lin0.bias]
    for i in range(e.depth):
        for lins in (e.v_lins1, e.v_lins2, e.v_lins3, e.v_lins4):
            out += [lins[i].weight, lins[i].bias]
        out += [e.e_lins0[i].weight, e.e_lins0[i].bias]
        for bn in (e.v_bns[i].module, e.e_bns[i].module):
            out += [bn.weight, bn.bias, None]
    h = net.par_net_heu.lins
    out += [h[0].weight, h[0].bias, h[1].weight, h[1].bias, h[2].weight, h[2].bias]
    return [(t, 2 * UNITS if t is None else t.numel()) for t in out]


class _FlatState:
    """The parameters of emb_net / par_net_heu re-homed as views of ONE fp32 buffer in the packed weight order, and the
    BatchNorm running statistics as rows of two [12, 2, 32] buffers (the flat-parameter idiom of DDP / FSDP): the
    training kernels read the buffer directly -- no per-call packing -- and the packed gradient is handed back to
    autograd as views.  Parameter / buffer objects, names, shapes and state_dict keys are unchanged."""

    def __init__(self, net):
        e = net.emb_net
        entries = _packed_entries(net)
        dev = e.v_lin0.weight.device
        for t, _ in entries:
            if t is not None and (t.dtype != torch.float32 or t.device != dev):
                raise _lib.DeepAcoError("deepaco_b200.Net: parameters must be fp32 on one device")
        self.flat = torch.zeros(sum(k for _, k in entries), dtype=torch.float32, device=dev)
        assert self.flat.numel() == weight_count(e.feats)
        self.slots = []                                           # (param, offset, numel, shape)
        dead = {id(t) for t in (e.v_lins1[-1].weight, e.v_lins1[-1].bias, e.v_lins2[-1].weight, e.v_lins2[-1].bias,
                                e.v_bns[-1].module.weight, e.v_bns[-1].module.bias)}
        off = 0
        with torch.no_grad():
            for t, k in entries:
                if t is not None:
                    view = self.flat[off:off + k].view(t.shape)
                    view.copy_(t.data)
                    t.data = view
                    # the last layer's node update never reaches the output (EmbNet.forward returns w, tsp/net.py:45):
                    # the reference's autograd leaves those .grad at None, so they are not autograd inputs here either
                    self.slots.append((t, off, k, tuple(t.shape), id(t) not in dead))
                off += k
            bns = [m.module for l in range(e.depth) for m in (e.v_bns[l], e.e_bns[l])]
            self.bns = bns
            self.track = all(m.track_running_stats and m.running_mean is not None for m in bns)
            if self.track:
                self.rm = torch.stack([m.running_mean.to(torch.float32) for m in bns]).view(e.depth, 2, UNITS).contiguous()
                self.rv = torch.stack([m.running_var.to(torch.float32) for m in bns]).view(e.depth, 2, UNITS).contiguous()
                self.nbt = torch.stack([m.num_batches_tracked for m in bns])
                for j, m in enumerate(bns):
                    m.running_mean.data = self.rm.view(-1, UNITS)[j]
                    m.running_var.data = self.rv.view(-1, UNITS)[j]
                    m.num_batches_tracked.data = self.nbt[j]
        self.base = self.flat.data_ptr()

    def intact(self):
        """False once something re-allocated a parameter / buffer (net.to(...), net.float(), a manual `.data =`)."""
        base = self.base
        if self.flat.data_ptr() != base:
            return False
        for t, off, _, _, _ in self.slots:
            if t.data_ptr() != base + 4 * off:
                return False
        if self.track:
            rm, rv = self.rm.data_ptr(), self.rv.data_ptr()
            for j, m in enumerate(self.bns):
                if m.running_mean is None or m.running_mean.data_ptr() != rm + 4 * UNITS * j or m.running_var.data_ptr() != rv + 4 * UNITS * j:
                    return False
        return True


def flat_state(net):
    st = net.__dict__.get("_flat_train_state")
    if st is None or not st.intact():
        st = _FlatState(net)
        net.__dict__["_flat_train_state"] = st
    return st


class _GnnTrain(torch.autograd.Function):
    """heu [B, E], stats [B, 12, 6, 32] = f(parameters, graph).  `flat` is the packed weight buffer the parameters alias;
    `params` are the live, trainable parameters (autograd inputs), `slots[j]` = (offset, numel, shape) of params[j]."""

    @staticmethod
    def forward(ctx, flat, x, graph, feats, ctas, bn_eps, slots, *params):
        import ctypes
        dev = x.device
        bufs = train_buffers(graph["B"], graph["n"], graph["E"], dev)
        heu = torch.empty((graph["B"], graph["E"]), dtype=torch.float32, device=dev)
        a, keep = train_args(x, graph, flat, bufs, feats, ctas, bn_eps, heu_out=heu)
        with torch.cuda.device(dev):
            check(lib().deepaco_gnn_train_forward(ctypes.byref(a), stream_ptr(dev)), "deepaco_gnn_train_forward")
        ctx.state = (flat, x, graph, bufs, feats, ctas, bn_eps, slots)
        ctx.save_for_backward(*params)          # the backward re-reads the live weights (the parameters alias `flat`):
                                                # autograd's version check on the saved parameters catches in-place updates
        ctx.mark_non_differentiable(bufs["stats"])
        return heu, bufs["stats"]

    @staticmethod
    def backward(ctx, g_heu, _g_stats):
        import ctypes
        flat, x, graph, bufs, feats, ctas, bn_eps, slots = ctx.state
        _ = ctx.saved_tensors                   # raises "... modified by an inplace operation" if a parameter changed since
        dev = x.device
        g_heu = g_heu.to(torch.float32).contiguous()
        grad = torch.zeros((graph["B"] * ctas, flat.numel()), dtype=torch.float32, device=dev)
        a, keep = train_args(x, graph, flat, bufs, feats, ctas, bn_eps, grad_heu=g_heu, grad_weights=grad)
        with torch.cuda.device(dev):
            check(lib().deepaco_gnn_train_backward(ctypes.byref(a), stream_ptr(dev)), "deepaco_gnn_train_backward")
        total = grad.sum(dim=0) if grad.shape[0] > 1 else grad[0]
        need = ctx.needs_input_grad[7:]
        grads = tuple(total[off:off + k].view(shape) if nd else None for (off, k, shape), nd in zip(slots, need))
        return (None,) * 7 + grads


def gnn_train_forward(net, x, edge_index, edge_attr, ctas=None, graph=None):
    """Training-mode Net.forward for one graph or a batch ([B, ...] tensors with identical n and E; every graph is its
    own forward call as far as BatchNorm is concerned).  Differentiable w.r.t. the parameters of `net`; updates the
    BatchNorm running statistics like the sequence of per-graph calls would (tsp/net.py:41-44 with PyG BatchNorm)."""
    batched = x.dim() == 3
    if not batched:
        x, edge_index, edge_attr = x[None], edge_index[None], edge_attr[None]
    _lib.require_cuda(x, "pyg.x")
    n = x.shape[1]
    if graph is None:
        graph = train_graph(edge_index, edge_attr, n)
    e = net.emb_net
    st = flat_state(net)
    live = [(t, off, k, shape) for t, off, k, shape, alive in st.slots if alive and t.requires_grad]
    heu, stats = _GnnTrain.apply(st.flat, x.to(torch.float32).contiguous(), graph, e.feats,
                                 ctas or default_train_ctas(graph["E"]), float(st.bns[0].eps),
                                 [(off, k, shape) for _, off, k, shape in live], *[t for t, _, _, _ in live])
    _update_running_stats(st, stats, n, graph["E"])
    return heu if batched else heu[0]


@torch.no_grad()
def _update_running_stats(st, stats, n, E):
    """running_mean / running_var / num_batches_tracked exactly as nn.BatchNorm1d.forward does in training mode
    (exponential moving average with the module's momentum, unbiased variance), one update per graph, on the stacked
    [12, 2, 32] buffers the modules' running statistics alias."""
    if not st.track:
        return
    s5 = stats.view(stats.shape[0], DEPTH, 2, 3, UNITS)            # [B][layer][node | edge][mean, invstd, biased var][32]
    scale = torch.tensor([n / max(n - 1, 1), E / max(E - 1, 1)], dtype=torch.float32).view(1, 1, 2, 1).to(stats.device, non_blocking=True)
    uvar = s5[:, :, :, 2] * scale
    moms = {m.momentum for m in st.bns}
    for b in range(stats.shape[0]):
        if len(moms) == 1 and None not in moms:
            mom = next(iter(moms))
            st.rm.mul_(1.0 - mom).add_(s5[b, :, :, 0], alpha=mom)
            st.rv.mul_(1.0 - mom).add_(uvar[b], alpha=mom)
        else:
            for j, m in enumerate(st.bns):
                mom = m.momentum if m.momentum is not None else 1.0 / float(m.num_batches_tracked + 1)   # cumulative average
                l, k = divmod(j, 2)
                m.running_mean.mul_(1.0 - mom).add_(s5[b, l, k, 0], alpha=mom)
                m.running_var.mul_(1.0 - mom).add_(uvar[b, l, k], alpha=mom)
        st.nbt += 1


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
        """Eval-mode packed weights, re-packed when any tensor that goes into them was written or re-allocated.  The
        check walks the _parameters / _buffers dicts of the modules that matter directly (state_dict() costs ~1 ms of
        Python per call, more than the network itself)."""
        mods = self.__dict__.get("_pack_modules")
        if mods is None:
            mods = [m for part in (self.emb_net, self.par_net_heu) for m in part.modules() if m._parameters or m._buffers]
            self.__dict__["_pack_modules"] = mods
        key = tuple((t.data_ptr(), t._version) for m in mods for d in (m._parameters, m._buffers) for t in d.values()
                    if t is not None)
        st = self.__dict__.get("_flat_train_state")
        if st is not None and st.track:      # training mode updates the running statistics through the stacked buffers
            key += (st.rm._version, st.rv._version)     # they alias (a `.data` view keeps its own version counter)
        if self._packed is None or key != self._packed_key:
            self._packed, self._packed_key = pack_weights(self), key
        return self._packed

    def forward(self, pyg):
        x, edge_index, edge_attr = pyg.x, pyg.edge_index, pyg.edge_attr
        graph = _graph_of(pyg, backward=self.training) if x.dim() == 2 and x.is_cuda else None
        if self.training:
            return gnn_train_forward(self, x, edge_index, edge_attr, graph=graph)
        return gnn_forward(self._weights(), self.FEATS, x, edge_index, edge_attr, graph=graph)

    @torch.no_grad()
    def heuristic_matrices(self, node_features, distances, k_sparse, eps=1e-10):
        """[B, n, n] heuristic matrices of a batch of k-NN TSP instances in one launch (eval mode)."""
        return knn_heuristic_matrices(self._weights(), self.FEATS, node_features, distances, k_sparse, eps)

    @torch.no_grad()
    def dense_heuristic_matrices(self, node_features, distances, eps=1e-10):
        """[B, N, N] heuristic matrices of a batch of complete-graph (CVRP) instances in one launch (eval mode)."""
        return dense_heuristic_matrices(self._weights(), self.FEATS, node_features, distances, eps)

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


def _graph_of(pyg, backward):
    """train_graph(...) of one pyg instance, remembered on the instance when it is this package's own Data (validation
    and training loops present the same instances every epoch; a foreign Data object is never written to)."""
    ei, ea = pyg.edge_index, pyg.edge_attr
    own = isinstance(pyg, Data)
    if own:
        hit = pyg.__dict__.get("_deepaco_graph")
        # the entry keeps the tensors it was built from alive, so identity + version is a sound "unchanged" test
        if (hit is not None and hit[0] is ei and hit[1] is ea and hit[2] == (ei._version, ea._version, pyg.x.shape[0])
                and (hit[3]["col_ptr"] is not None or not backward)):
            return hit[3]
    graph = train_graph(ei[None], ea[None], pyg.x.shape[0], backward=backward)
    if own:
        pyg.__dict__["_deepaco_graph"] = (ei, ea, (ei._version, ea._version, pyg.x.shape[0]), graph)
    return graph


class Data:
    """Attribute bag standing in for torch_geometric.data.Data: Net only reads .x, .edge_index, .edge_attr."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if k == "_deepaco_graph":
                del self.__dict__[k]
                continue
            if hasattr(v, "to"):
                setattr(self, k, v.to(device))
        return self
