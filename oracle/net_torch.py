"""Torch fp32 restatement of the heuristic network (reference tsp/net.py:27-45 EmbNet.forward with
`global_mean_pool` written as index_add / clamped degree, :62-75 MLP / ParNet.forward), autograd-capable and
device agnostic.

TEST INFRASTRUCTURE: the checker for csrc/gnn.cu (eval mode) and csrc/gnn_train.cuh (training mode: heuristic,
parameter gradients, BatchNorm running statistics).  Pinned to the unmodified reference by tests/test_oracle_golden.py
(eval- and train-mode outputs and parameter gradients of the reference's Net on the pretrained checkpoint).
Only tests/, tools/ probes and bench legs may import it; nothing under deepaco_b200/ does.
"""
import torch
from torch.nn import functional as F


def emb_forward(emb, x, edge_index, edge_attr):
    """`emb`: a deepaco_b200.net.EmbNet (module tree == reference EmbNet).  Returns the edge embedding w [E, 32]."""
    src, dst = edge_index[0], edge_index[1]
    n = x.shape[0]
    deg = torch.zeros(n, device=x.device, dtype=x.dtype).index_add_(0, src, torch.ones_like(src, dtype=x.dtype))
    x = F.silu(emb.v_lin0(x))
    w = F.silu(emb.e_lin0(edge_attr))
    for i in range(emb.depth):
        msg = torch.sigmoid(w) * emb.v_lins2[i](x)[dst]
        agg = torch.zeros_like(x).index_add_(0, src, msg) / deg.clamp(min=1).unsqueeze(-1)
        x_new = x + F.silu(emb.v_bns[i](emb.v_lins1[i](x) + agg))
        w = w + F.silu(emb.e_bns[i](emb.e_lins0[i](w) + emb.v_lins3[i](x)[src] + emb.v_lins4[i](x)[dst]))
        x = x_new
    return w


def net_forward(net, pyg):
    """Net.forward (tsp/net.py:84-88) through torch ops: BatchNorm follows net.training like the reference."""
    return net.par_net_heu(emb_forward(net.emb_net, pyg.x, pyg.edge_index, pyg.edge_attr))
