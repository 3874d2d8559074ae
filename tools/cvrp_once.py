"""A few CVRP-100 x 512-ant ACO iterations for 64 colonies (config C4 batched) -- for ncu captures of the list kernel.
    python tools/cvrp_once.py [colonies] [iterations]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepaco_b200 import _engine as E

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(1)
N = 101
loc = torch.cat((torch.tensor([[0.5, 0.5]], device=dev), torch.rand(100, 2, device=dev)))
dc = torch.norm(loc[:, None] - loc, dim=2, p=2)
dc[torch.arange(N), torch.arange(N)] = 1e-10
dem = torch.cat((torch.zeros(1, device=dev), torch.randint(1, 10, (100,), device=dev).float()))
hc = torch.rand(N, N, device=dev) * 0.98 + 1e-10
r = E.CvrpRunner(dc.expand(B, N, N).contiguous(), dem.expand(B, N).contiguous(), hc.expand(B, N, N).contiguous(),
                 torch.ones(B, N, N, device=dev), 512)
for _ in range(T):
    r.run(1, 7, [4000 * b for b in range(B)])
torch.cuda.synchronize()
print("ok cvrp", B, float(r.lowest_cost.min()))
