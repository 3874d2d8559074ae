"""Multi-GPU check (run under torchrun with the NCCL backend): one TSP-100 colony with 512 ants sharded by ants
over the ranks, ONE NCCL collective (all-gather of the compact tours) per ACO iteration; the result must be
bit-identical to the single-GPU run and to the reference op sequence.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ant_shard_nccl.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from deepaco_b200.dist import AntShardedColony, CudaTspBackend

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, A, T = 100, 512, 10
g = torch.Generator().manual_seed(5)
xy = torch.rand((n, 2), generator=g).to(dev)
d = torch.norm(xy[:, None] - xy, dim=2, p=2)
d[torch.arange(n), torch.arange(n)] = 1e9
_, idx = torch.topk(d, 20, dim=1, largest=False)
heu = torch.full_like(d, 1e-10).scatter_(1, idx, (torch.rand((n, 20), generator=g) * 0.9 + 0.05).to(dev))

results = {}
for mode in ("nccl", "p2p"):
    try:
        col = AntShardedColony(CudaTspBackend(d, heu), torch.ones(n, n, device=dev), A, exchange=mode)
        col.run(3, seed=11, offset=0)                       # warm-up (channels, kernels)
        col = AntShardedColony(CudaTspBackend(d, heu), torch.ones(n, n, device=dev), A, exchange=mode)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        low = col.run(T, seed=11, offset=0)
        torch.cuda.synchronize(); dist.barrier()
        results[mode] = (col, low, (time.perf_counter() - t0) / T)
    except Exception as exc:                                # e.g. no P2P / symmetric memory on this box
        if rank == 0:
            print({"exchange": mode, "error": repr(exc)[:300]})
if rank == 0:
    for mode, (col, low, dt) in results.items():
        results[mode] = (col, low, dt)
col, low, dt = results.get("p2p", results.get("nccl"))
if rank == 0:
    from deepaco_b200.tsp.aco import ACO
    from oracle import aco_torch as O
    torch.manual_seed(11)
    single = ACO(d, n_ants=A, heuristic=heu, device=dev)
    single.run(T)
    torch.manual_seed(11)
    ref = O.TspColony(d, A, heuristic=heu)
    ref.run(T)
    for mode, (c2, l2, dt2) in results.items():
        print({"exchange": mode, "ms_per_iteration": dt2 * 1e3, "collectives_per_iteration": c2.collectives / T,
               "identical_to_single_gpu": bool(torch.equal(c2.pheromone, single.pheromone))})
    print({"world": world, "collectives_per_iteration": col.collectives / T, "ms_per_iteration": dt * 1e3,
           "identical_to_single_gpu": bool(torch.equal(col.pheromone, single.pheromone)),
           "identical_to_reference_ops": bool(torch.equal(col.pheromone, ref.pheromone)),
           "best_cost": float(low), "single_gpu_best_cost": float(single.lowest_cost)})
dist.destroy_process_group()
