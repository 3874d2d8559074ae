#!/bin/bash
# compute-sanitizer passes over the hot-path kernels (run on the GPU box): memcheck + racecheck + synccheck on a
# script that exercises every kernel family at small sizes.
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
from deepaco_b200 import _engine as E
from deepaco_b200.tsp.aco import ACO
from deepaco_b200.tsp_nls.aco import ACO as NlsACO
from deepaco_b200.cvrp.aco import ACO as CvrpACO
dev = "cuda"
torch.manual_seed(0)
for n, k in ((40, 8), (100, 20), (260, 20)):
    xy = torch.rand(n, 2, device=dev)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2); d[torch.arange(n), torch.arange(n)] = 1e9
    _, idx = torch.topk(d, k, dim=1, largest=False)
    heu = torch.full_like(d, 1e-10).scatter_(1, idx, torch.rand(n, k, device=dev) * 0.9 + 0.05)
    for h in (heu, 1 / d):
        aco = ACO(d, n_ants=24, heuristic=h, device=dev)
        aco.run(2)
        h2 = h.clone().requires_grad_(True)
        c, lp = ACO(d, n_ants=8, heuristic=h2, device=dev).sample()
        (c.detach() * lp.sum(0)).sum().backward()
    if n <= 100:
        nls = NlsACO(d, n_ants=8, heuristic=heu, device=dev, local_search="nls")
        nls.run(1)
N = 31
loc = torch.rand(N, 2, device=dev); dd = torch.norm(loc[:, None] - loc, dim=2, p=2); dd[torch.arange(N), torch.arange(N)] = 1e-10
dem = torch.cat((torch.zeros(1, device=dev), torch.randint(1, 10, (N - 1,), device=dev).float()))
cv = CvrpACO(dd, dem, n_ants=16, device=dev)
cv.run(2); cv.sample()
from deepaco_b200.tsp.net import Net
from deepaco_b200.tsp.utils import gen_pyg_data
net = Net().to(dev).eval()
pyg, _ = gen_pyg_data(torch.rand(50, 2, device=dev), 10)
with torch.no_grad():
    net(pyg)
# round 2: batched front end (tensor-core tiles), training-mode network, roulette sampler, many-ants update + general
# Philox geometry, candidate-list refresh, device-sharded run with virtual ranks, host-buffer pipeline
net.heuristic_matrices(torch.rand(3, 50, 2, device=dev), torch.cdist(torch.rand(3, 50, 2, device=dev), torch.rand(3, 50, 2, device=dev)) + 0.01, 10)
net.train()
net(pyg).sum().backward()
nl = NlsACO(d[:40, :40].contiguous() + 0.01, n_ants=8, device=dev, local_search="2opt")
nl.sample(inference=True); nl.run(1, inference=True)
n = 64
xy = torch.rand(n, 2, device=dev)
d2 = torch.norm(xy[:, None] - xy, dim=2, p=2); d2[torch.arange(n), torch.arange(n)] = 1e9
_, idx = torch.topk(d2, 8, dim=1, largest=False)
h2 = torch.full_like(d2, 1e-10).scatter_(1, idx, torch.rand(n, 8, device=dev) * 0.9 + 0.05)
big = E.TspRunner(d2, h2, torch.ones_like(d2), 6000)           # > 303104 / 64 ants: general geometry + row update kernel
big.run(5, 3)
from deepaco_b200.dist import DeviceShardedColony, local_peer_memory, warm_virtual_ranks
warm_virtual_ranks(lambda: E.TspRunner(d2, h2, torch.ones_like(d2), 96), 2)
peers = local_peer_memory(1, 96, n, torch.device(dev, 0), 2)
cols = [DeviceShardedColony(E.TspRunner(d2, h2, torch.ones_like(d2), 96), peers[r], timeout_ms=20000) for r in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize()
for r in range(2):
    with torch.cuda.stream(streams[r]):
        cols[r].run(2, 5)
torch.cuda.synchronize()
for c in cols:
    c.check()
B = 16
dh = d2.expand(B, n, n).contiguous().cpu().pin_memory(); hh = h2.expand(B, n, n).contiguous().cpu().pin_memory()
ph = torch.ones(B, n, n).pin_memory(); lo = torch.empty(B).pin_memory(); sp = torch.empty((B, n), dtype=torch.int64).pin_memory()
rh = E.TspRunner(torch.zeros(B, n, n, device=dev), torch.zeros(B, n, n, device=dev), torch.zeros(B, n, n, device=dev), 32)
rh.run_host(2, 7, dh, hh, ph, lo, sp)
torch.cuda.synchronize()
print("sanitize script done")
PY
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -4 gpurun_out/sanitize_$tool.log
done
