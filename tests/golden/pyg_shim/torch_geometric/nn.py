"""torch_geometric.nn stand-ins: `global_mean_pool` and `BatchNorm` (see package docstring)."""
import torch


def global_mean_pool(x, batch, size=None):
    # PyG 2.0.4: scatter(x, batch, dim=0, dim_size=size, reduce='mean') with size = batch.max()+1
    size = int(batch.max().item()) + 1 if size is None else size
    out = torch.zeros((size, x.shape[1]), dtype=x.dtype, device=x.device)
    out.index_add_(0, batch, x)
    cnt = torch.zeros((size,), dtype=x.dtype, device=x.device)
    cnt.index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp(min=1).unsqueeze(-1)


class BatchNorm(torch.nn.Module):
    # PyG wraps nn.BatchNorm1d as `.module` (checkpoint keys `*.module.running_mean`)
    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = torch.nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)
