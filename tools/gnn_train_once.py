"""A few native training-mode forward + backward passes on one graph (for ncu captures).
    python tools/gnn_train_once.py {C2|C3|C4} [ctas] [repeats]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_gnn_train import graph  # noqa: E402

kind = {"C2": "C2 tsp n=100 k=20", "C3": "C3 tsp_nls n=500 k=50", "C4": "C4 cvrp N=101 dense"}[sys.argv[1] if len(sys.argv) > 1 else "C2"]
if len(sys.argv) > 2:
    os.environ["DEEPACO_GNN_CTAS"] = sys.argv[2]
Net, pyg = graph(kind)
torch.manual_seed(0)
net = Net().to("cuda").train()
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    net.zero_grad()
    net(pyg).sum().backward()
torch.cuda.synchronize()
print("ok", kind)
