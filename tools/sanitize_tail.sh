#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the one-launch iteration tail (tsp_tail_kernel) and the
# ant-sequential update (tsp_update_seq_kernel) at small sizes: even / odd n (TMA bulk copies / plain loads), the plain
# and the vectorised ATen plan (n < 128 / n >= 128), full and partial 16-ant blocks, elitist and min-max variants.
mkdir -p gpurun_out
cat > /tmp/san_tail.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from deepaco_b200 import _engine as E
dev = "cuda"
torch.manual_seed(0)
for n, A, kw in ((40, 40, {}), (61, 24, {"elitist": True}), (100, 48, {"min_max": True, "ph_min": 0.1}), (136, 32, {})):
    B = 64
    xy = torch.rand(B, n, 2, device=dev)
    d = torch.cdist(xy, xy); i = torch.arange(n, device=dev); d[:, i, i] = 1e9
    _, idx = torch.topk(d, 8, dim=2, largest=False)
    heu = torch.full_like(d, 1e-10).scatter_(2, idx, torch.rand(B, n, 8, device=dev) * 0.9 + 0.05)
    r = E.TspRunner(d, heu, torch.ones_like(d) * (0.1 if kw.get("min_max") else 1.0), A, **kw)
    r.run(3, 5)                                          # B >= 64, n <= 128: tail kernel; n = 136: separate kernels
    tours = r.tours.clone(); costs = r.costs.clone()
    ph = torch.ones_like(d)
    E.tsp_update_tours_(ph, tours, costs, decay=0.9)    # ant-sequential update from compact tours
    torch.cuda.synchronize()
print("sanitize tail script done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 150 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_tail.py > gpurun_out/sanitize_tail_$tool.log 2>&1
  tail -3 gpurun_out/sanitize_tail_$tool.log
done
