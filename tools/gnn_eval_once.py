"""A few eval-mode forwards of one instance (group kernel) -- for ncu captures.   python tools/gnn_eval_once.py {C2|C3|C4}"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_gnn_train import graph  # noqa: E402

kind = {"C2": "C2 tsp n=100 k=20", "C3": "C3 tsp_nls n=500 k=50", "C4": "C4 cvrp N=101 dense"}[sys.argv[1] if len(sys.argv) > 1 else "C3"]
Net, pyg = graph(kind)
torch.manual_seed(0)
net = Net().to("cuda").eval()
with torch.no_grad():
    for _ in range(4):
        out = net(pyg)
torch.cuda.synchronize()
print("ok", kind, float(out.mean()))
