"""Where the time of deepaco_tsp_run_host goes (informational): PCIe bandwidth, device-resident step, host-entry step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
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
low_h = torch.empty(B).pin_memory()
sp_h = torch.empty((B, n), dtype=torch.int64).pin_memory()
buf = torch.empty_like(d)
t = timed(lambda: buf.copy_(d_h, non_blocking=True), 20)
print(f"H2D {d_h.numel() * 4 / 1e6:.1f} MB pinned: {t * 1e3:.0f} us -> {d_h.numel() * 4 / t / 1e6:.1f} GB/s")
t = timed(lambda: ph_h.copy_(buf, non_blocking=True), 20)
print(f"D2H same: {t * 1e3:.0f} us")
r = E.TspRunner(d, heu, torch.ones_like(d), A)
offs = torch.tensor([b * 4_000_000 for b in range(B)], dtype=torch.int64, device=dev)
st = {"it": 0}


def dev_step():
    r.run(1, 1234, st["it"] * r.increment, offs)
    st["it"] += 1


print(f"device-resident step: {timed(dev_step, 20) * 1e3:.0f} us")
for chunks in (("",) if os.environ.get("E2E_PROBE_CUTS") else ("1", "2", "3", "4", "8", "")):
    if chunks:
        os.environ["DEEPACO_HOST_CHUNKS"] = chunks
    else:
        os.environ.pop("DEEPACO_HOST_CHUNKS", None)
    for ph in (None, ph_h):
        def host_step():
            r.run_host(1, 1234, d_h, heu_h, ph, low_h, sp_h, st["it"] * r.increment, offs, copy_back_pheromone=ph is not None)
            st["it"] += 1
        t0 = time.perf_counter()
        t = timed(host_step, 20)
        print(f"run_host chunks={chunks or 'default'} pheromone={'host' if ph is not None else 'ones'}: {t * 1e3:.0f} us/step")

os.environ.pop("DEEPACO_HOST_CHUNKS", None)
CUTS = os.environ.get("E2E_PROBE_CUTS", "32,128,224;16,72,224;16,64,144,232;16,56,128,200,240;24,104,232;16,80,160,240;8,40,120,200,248").split(";")
for cuts in CUTS:
    os.environ["DEEPACO_HOST_CUTS"] = cuts
    def host_step():
        r.run_host(1, 1234, d_h, heu_h, ph_h, low_h, sp_h, st["it"] * r.increment, offs, copy_back_pheromone=True)
        st["it"] += 1
    t = timed(host_step, 30)
    print(f"run_host cuts={cuts} pheromone=host: {t * 1e3:.0f} us/step")
os.environ.pop("DEEPACO_HOST_CUTS", None)
