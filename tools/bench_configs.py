"""Timing of the other BASELINE.json configs (C3-C5) on one GPU -- informational, not the contract bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepaco_b200 import _engine as E

dev = "cuda"


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def tsp_inst(B, n, k):
    torch.manual_seed(0)
    xy = torch.rand(B, n, 2, device=dev)
    d = torch.cdist(xy, xy)
    d[:, torch.arange(n), torch.arange(n)] = 1e9
    _, idx = torch.topk(d, k, dim=2, largest=False)
    heu = torch.full_like(d, 1e-10).scatter_(2, idx, torch.rand(B, n, k, device=dev) * 0.9 + 0.05)
    return d, heu


# C3: TSP-NLS n=500, 256 ants: sample / 2-opt (125 passes) / NLS
d, heu = tsp_inst(1, 500, 50)
ph = torch.ones_like(d)
out = {}
def samp():
    out["t"] = E.tsp_sample(ph[0], heu[0], 256, start_node=0, double_norm=True, seed=1, want_paths=False, want_tours=True)[2]
print(f"C3 TSP-500 x256: sample {timeit(samp):8.2f} ms", flush=True)
hd = (1 / (heu[0] / heu[0].max(-1, keepdim=True).values + 1e-5)).contiguous()
base = out["t"].clone()
def topt():
    t = base.clone(); E.two_opt_(d[0], t, 125)
def nls():
    t = base.clone(); E.tsp_nls_(d[0], hd, t, 125)
print(f"C3 2-opt(125 passes) {timeit(topt, 3, 1):8.2f} ms   NLS(T_nls=10,T_p=20) {timeit(nls, 2, 1):8.2f} ms", flush=True)

# C4: CVRP-100 x 512 ants
torch.manual_seed(1)
N = 101
loc = torch.cat((torch.tensor([[0.5, 0.5]], device=dev), torch.rand(100, 2, device=dev)))
dc = torch.norm(loc[:, None] - loc, dim=2, p=2); dc[torch.arange(N), torch.arange(N)] = 1e-10
dem = torch.cat((torch.zeros(1, device=dev), torch.randint(1, 10, (100,), device=dev).float()))
hc = torch.rand(N, N, device=dev) * 0.98 + 1e-10
for B in (1, 64):
    r = E.CvrpRunner(dc.expand(B, N, N).contiguous(), dem.expand(B, N).contiguous(), hc.expand(B, N, N).contiguous(),
                     torch.ones(B, N, N, device=dev), 512)
    t = timeit(lambda: r.run(1, 7, [4000 * b for b in range(B)]), 5, 2)
    print(f"C4 CVRP-100 x512 x{B} colonies: {t:8.3f} ms/iteration -> {B * 512 / t / 1e3:8.2f} M routes-sets/s", flush=True)

# C5: 64 x TSP-200 x 256 ants (one GPU's view: all 64, and the 8-per-GPU shard)
for B in (64, 8):
    d, heu = tsp_inst(B, 200, 20)
    r = E.TspRunner(d, heu, torch.ones_like(d), 256)
    t = timeit(lambda: r.run(1, 3, 0, [100000 * b for b in range(B)]), 5, 2)
    print(f"C5 TSP-200 x256 x{B} colonies: {t:8.3f} ms/iteration -> {B * 256 / t / 1e3:8.2f} M tours/s", flush=True)

# heuristic network, eval mode (one-off per instance): single instance at the C2 / C3 / C4 graph sizes, and the batched
# k-NN front end (instance -> graph -> network -> dense heuristic) the C2 / C5 drivers use
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from bench_gnn_train import graph as _graph  # noqa: E402
for kind in ("C2 tsp n=100 k=20", "C3 tsp_nls n=500 k=50", "C4 cvrp N=101 dense"):
    Net, pyg = _graph(kind)
    torch.manual_seed(0)
    net = Net().to(dev).eval()
    with torch.no_grad():
        t = timeit(lambda: net(pyg), 20, 3)
        os.environ["DEEPACO_GNN_CTAS"] = "1"
        t1 = timeit(lambda: net(pyg), 10, 2)
        os.environ.pop("DEEPACO_GNN_CTAS")
    print(f"GNN eval forward {kind}: {t:8.3f} ms / instance on a CTA group, {t1:8.3f} ms on one CTA (Python front end included)", flush=True)
from deepaco_b200.tsp.net import Net as _TspNet  # noqa: E402
net = _TspNet().to(dev).eval()
for B, n, k in ((256, 100, 20), (64, 200, 20)):
    xy = torch.rand(B, n, 2, device=dev)
    d = torch.cdist(xy, xy)
    d[:, torch.arange(n), torch.arange(n)] = 1e9
    t = timeit(lambda: net.heuristic_matrices(xy, d, k), 10, 3)
    print(f"GNN batched front end {B} x TSP-{n} (k={k}): {t:8.3f} ms -> {t / B * 1e3:8.1f} us / instance", flush=True)

# the reference's inference driver (tsp/test.ipynb cell 1: infer_instance) on one TSP-100 instance, 512 ants, T = 10
from deepaco_b200.tsp.aco import ACO as _ACO  # noqa: E402
from deepaco_b200.tsp.utils import gen_pyg_data as _gen  # noqa: E402
pyg, dist = _gen(torch.rand(100, 2, device=dev), 20)
def infer():
    with torch.no_grad():
        heu = net.reshape(pyg, net(pyg)) + 1e-10
    return _ACO(dist, 512, heuristic=heu, device=dev).run(10)
print(f"infer_instance TSP-100, 512 ants, T=10: {timeit(infer, 10, 3):8.3f} ms", flush=True)
