import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch
from bench_legs import timed, tsp_instances
from deepaco_b200 import _engine as E
from deepaco_b200.heuristics import tsp_heuristic
dev = torch.device("cuda:0")
B, n, A = 256, 100, 512
coords, d = tsp_instances(B, n, 1234, dev)
heu, _ = tsp_heuristic(coords, d, 20)
d_h, heu_h = d.cpu().pin_memory(), heu.cpu().pin_memory()
ph_h = torch.ones_like(d_h).pin_memory()
print("pinned:", d_h.is_pinned(), heu_h.is_pinned(), ph_h.is_pinned())
low_h = torch.empty(B).pin_memory(); sp_h = torch.empty((B, n), dtype=torch.int64).pin_memory()
r = E.TspRunner(d, heu, torch.ones_like(d), A)
offs = torch.tensor([b * 4_000_000 for b in range(B)], dtype=torch.int64, device=dev)
os.environ["DEEPACO_HOST_DEBUG"] = "1"
for chunks in ("1", "3"):
    os.environ["DEEPACO_HOST_CHUNKS"] = chunks
    for ph in (None, ph_h):
        print("chunks", chunks, "ph", ph is not None, flush=True)
        for i in range(4):
            r.run_host(1, 1234, d_h, heu_h, ph, low_h, sp_h, i * r.increment, offs, copy_back_pheromone=ph is not None)
