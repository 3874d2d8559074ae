#!/usr/bin/env python
"""Contract benchmark: ant-tours/s of the ACO iteration (construction + cost + best tracking + pheromone
update) on TSP-100 with 512 ants per colony (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--colonies B] [--impl reference]

One step = one ACO iteration over a batch of B independent colonies per GPU (weak scaling: every rank owns
its own B colonies, no data-path collective).  Prints ONE JSON line on rank 0.

  value      device-resident throughput (inputs in HBM), CUDA events, max over ranks
  e2e        same step through the C-ABI host entry (deepaco_tsp_run_host): pinned host matrices in,
             pheromone / best cost / best tour out, copies inside the timed region
  roofline   sampling kernel: algorithmic bytes (80,000 B per TSP-100 tour) / its live
             CUDA-event duration, against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the reference's PyTorch-CPU path (oracle port, op for op) on the host cores of this box
  configs    the other BASELINE.json configs (tools/bench_legs.py): c2_dense, c3, c4 on rank 0's GPU; c5 = 64 x TSP-200 x
             256 ants split over the N ranks with the result gather inside the timed region (strong scaling)
  ant_sharded  one TSP-200 colony of 16384 ants with its ants split over the N ranks (deepaco_tsp_run_shard: fused
             NVLink peer stores + flag barrier, no host sync), timed against the same colony on one GPU, bits compared
  reference_cuda  the reference op sequence with device='cuda' on this GPU
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES, N_ANTS, K_SPARSE = 100, 512, 20
ALG_BYTES_PER_TOUR = (N_NODES - 1) * 2 * 4 * N_NODES + 8 * N_NODES      # S*R*4n + 8(S+1) = 80,000 B


def make_instances(B, seed, device):
    """Synthetic TSP-100 colonies: uniform coordinates (tsp/train.ipynb cell 2), distance matrix with 1e9
    diagonal (tsp/utils.py:4-14), heuristic from the heuristic network on the k=20 nearest-neighbour graph
    (+1e-10 elsewhere, tsp/test.ipynb cell 1)."""
    import torch
    from deepaco_b200.heuristics import tsp_heuristic
    g = torch.Generator(device="cpu").manual_seed(seed)
    coords = torch.rand((B, N_NODES, 2), generator=g).to(device)
    dist = torch.cdist(coords, coords)
    idx = torch.arange(N_NODES, device=device)
    dist[:, idx, idx] = 1e9
    heu, how = tsp_heuristic(coords, dist, K_SPARSE)
    return dist.contiguous(), heu.contiguous(), how


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (pynvml, 10 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def finish(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_throughput(steps, warmup, seed=1234, budget_s=25.0):
    """ant-tours/s of the reference's CPU path (oracle/aco_torch.py, same ATen ops) on ONE colony of the
    workload; every step is one ACO iteration (512 tours)."""
    import torch
    from oracle import aco_torch as O
    ncpu = os.cpu_count() or 1
    g = torch.Generator().manual_seed(seed)
    coords = torch.rand((N_NODES, 2), generator=g)
    dist = torch.norm(coords[:, None] - coords, dim=2, p=2)
    dist[torch.arange(N_NODES), torch.arange(N_NODES)] = 1e9
    _, idx = torch.topk(dist, K_SPARSE, dim=1, largest=False)
    heu = torch.full_like(dist, 1e-10)
    heu.scatter_(1, idx, torch.rand((N_NODES, K_SPARSE), generator=g) * 0.9 + 0.05)
    torch.manual_seed(seed)
    col = O.TspColony(dist, N_ANTS, heuristic=heu)
    # The path is ~10^4 small ATen ops per iteration: intra-op threading beyond a few cores only adds
    # synchronisation cost.  Use the thread count (out of 8 / 16 / 32 / all) that is fastest on this host.
    best_t, best_dt = None, None
    for t in [c for c in (8, 16, 32, ncpu) if c <= ncpu] or [ncpu]:
        torch.set_num_threads(t)
        col.run(1)
        t0 = time.perf_counter()
        col.run(1)
        dt1 = time.perf_counter() - t0
        if best_dt is None or dt1 < best_dt:
            best_t, best_dt = t, dt1
        elif dt1 > 1.5 * best_dt:
            break
    threads = best_t
    torch.set_num_threads(threads)
    for _ in range(warmup):
        col.run(1)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        col.run(1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": N_ANTS * done / dt, "unit": "ant-tours/s", "cores": threads, "kind": "port",
            "sample": f"{done} ACO iterations of one TSP-{N_NODES} colony x {N_ANTS} ants, torch {torch.__version__} CPU, "
                      f"{threads} of {ncpu} host threads (fastest of 8/16/32/all), {dt / done * 1e3:.1f} ms/iteration"}, dt / done * 1e3, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--colonies", type=int, default=256, help="colonies per GPU")
    ap.add_argument("--impl", default="deepaco_b200", choices=["deepaco_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"TSP-{N_NODES} x {N_ANTS} ants per colony, heuristic on k={K_SPARSE} sparse graph, 1 ACO iteration per step"

    if args.impl == "reference":
        if rank != 0:
            return
        cb, ms, done = cpu_reference_throughput(max(args.steps, 1), args.warmup, budget_s=150.0)
        line = {"impl": "reference", "metric": "ant-tours/sec TSP-100 n_ants=512", "value": cb["value"], "unit": "ant-tours/s",
                "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "colonies_per_step": 1, "device": "cpu",
                           "same_config_as_gpu_arm": False,
                           "note": "same colony (TSP-100 x 512 ants, k=20 sparse heuristic) and same unit (ant-tours/s) as the "
                                   "GPU arm, but ONE colony per step: the reference processes instances one after another "
                                   "(tsp/test.ipynb cell 1), the GPU arm runs 256 per launch; the GPU arm's `single_colony` "
                                   "is the one-colony-per-step figure.  Heuristic here: synthetic values on the k-NN graph "
                                   "(CPU speed does not depend on the values)"},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "ant-tours/s", "h2d_bytes_per_step": 0,
                                            "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist_pg
    from deepaco_b200 import _engine as E

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    if world > 1:
        # one process per GPU: run on (and take pinned memory from) the CPUs / NUMA node next to this rank's GPU
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        except Exception:
            pass
    if world > 1:
        dist_pg.init_process_group("nccl", device_id=dev)
    B, K, W = args.colonies, args.steps, max(args.warmup, 3)
    dist, heu, heu_how = make_instances(B, 1234 + rank, dev)
    ph0 = torch.ones_like(dist)
    runner = E.TspRunner(dist, heu, ph0, N_ANTS)
    seed = 1234
    inc = runner.increment
    base = [(rank * B + b) * 1_000_000 * 4 for b in range(B)]          # disjoint Philox ranges per global colony
    offsets = torch.tensor(base, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)

    def barrier():
        if world > 1:
            dist_pg.barrier()
        torch.cuda.synchronize()

    it = 0
    for _ in range(W):
        runner.run(1, seed, it * inc, offsets)
        it += 1
    # ---- device-resident timing: per-step event pairs, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    from deepaco_b200._lib import lib as _daco_lib
    barrier()
    sampler.start()
    launches0 = _daco_lib().deepaco_kernel_launches()
    for k in range(K):
        flush.zero_()
        ev[k][0].record()
        runner.run(1, seed, it * inc, offsets, sample_events=evs[k])
        ev[k][1].record()
        it += 1
    launches = _daco_lib().deepaco_kernel_launches() - launches0
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    samp_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    samp_mean = torch.tensor([sum(samp_ms) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist_pg.all_reduce(total_ms, op=dist_pg.ReduceOp.MAX)
        dist_pg.all_reduce(samp_mean, op=dist_pg.ReduceOp.MAX)
    total_ms, samp_mean = float(total_ms), float(samp_mean)
    tours_per_step = B * N_ANTS * world
    value = tours_per_step * K / (total_ms * 1e-3)

    # ---- the same iteration for ONE colony (512 ants): latency-bound view of the workload, reported beside `value`
    r1 = E.TspRunner(dist[:1], heu[:1], ph0[:1], N_ANTS)
    for _ in range(W):
        r1.run(1, seed, it * inc)
        it += 1
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    r1.run(K, seed, it * inc)
    s1.record()
    torch.cuda.synchronize()
    it += K
    single_ms = s0.elapsed_time(s1) / K

    # ---- end to end through the host-buffer C-ABI entry (pinned matrices in, results out)
    dist_h, heu_h = dist.cpu().pin_memory(), heu.cpu().pin_memory()
    ph_h = runner.pheromone.cpu().pin_memory()          # continue the colonies where the device-resident steps left them
    low_h = torch.empty(B, dtype=torch.float32).pin_memory()
    sp_h = torch.empty((B, N_NODES), dtype=torch.int64).pin_memory()
    r2 = E.TspRunner(dist, heu, ph0, N_ANTS)
    for _ in range(2):
        r2.run_host(1, seed, dist_h, heu_h, ph_h, low_h, sp_h, it * inc, offsets)
        it += 1
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        r2.run_host(1, seed, dist_h, heu_h, ph_h, low_h, sp_h, it * inc, offsets)
        it += 1
    e1.record()
    barrier()
    clocks = sampler.finish()          # sampled across both timed regions (device-resident steps and e2e steps)
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist_pg.all_reduce(e2e_ms, op=dist_pg.ReduceOp.MAX)
    e2e_value = tours_per_step * K / (float(e2e_ms) * 1e-3)
    mat = B * N_NODES * N_NODES * 4
    h2d, d2h = 3 * mat, mat + B * 4 + B * N_NODES * 8

    # ---- multi-GPU data-path legs (every rank takes part): C5 colony-sharded with the result gather in the timed
    #      region, and one big colony with its ANTS sharded (deepaco_tsp_run_shard)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_legs as L
    del runner, r1, r2, flush
    torch.cuda.empty_cache()
    leg_steps = max(3, min(K, 20))
    c5 = L.guarded("c5", L.leg_c5, dev, leg_steps)
    ant = L.guarded("ant_sharded", L.leg_ant_sharded, dev, leg_steps)

    if rank != 0:
        if world > 1:
            dist_pg.barrier()          # rank 0 enters this barrier after its single-rank legs and the CPU baseline
            dist_pg.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    achieved = ALG_BYTES_PER_TOUR * B * N_ANTS / (samp_mean * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    line = {
        "metric": "ant-tours/sec TSP-100 n_ants=512", "value": value, "unit": "ant-tours/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "colonies_per_gpu": B, "tours_per_step": tours_per_step, "heuristic": heu_how,
                   "l2": "flushed (256 MiB memset) between steps, outside the per-step CUDA-event pairs",
                   "parallelism": f"{world} x independent colony batches (no data-path collective)"},
        "roofline": {"bound": "hbm", "kernel": "K1 tour construction (aco_knn_kernel; aco_list_kernel for dense heuristics)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_TOUR * B * N_ANTS, "kernel_ms": samp_mean,
                     "kernel_share_of_step": samp_mean / (total_ms / K),
                     "note": "matrices are L2/SMEM resident: a throughput-normalised figure, not DRAM utilisation"},
        "e2e": {"value": e2e_value, "unit": "ant-tours/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "deepaco_tsp_run_host (C ABI): pinned host distances + heuristic + pheromone in, pheromone + best cost + "
                       "best tour out, every step; chunked upload / compute / download pipeline over four streams",
                "ms_per_step": float(e2e_ms) / K},
        "single_colony": {"value": N_ANTS / (single_ms * 1e-3), "unit": "ant-tours/s", "ms_per_iteration": single_ms,
                          "note": "one colony of 512 ants, K back-to-back iterations in one deepaco_tsp_run call (rank 0)"},
        "gpu_launches": int(launches), "clocks": clocks,
        # the other BASELINE.json configs on this GPU (C2 with the dense 1/dist heuristic, C3, C4) and C5 over all ranks
        "configs": {"c2_dense": L.guarded("c2_dense", L.leg_c2_dense, dev, leg_steps),
                    "c3": L.guarded("c3", L.leg_c3, dev, leg_steps),
                    "c4": L.guarded("c4", L.leg_c4, dev, leg_steps),
                    "gnn_front_end": L.guarded("gnn_front_end", L.leg_gnn_front_end, dev, leg_steps),
                    "c5": c5},
        "ant_sharded": ant,
    }
    # best-cost gap vs the reference on this GPU: same torch seed, one colony of the workload, oracle = the
    # reference's op sequence run with device='cuda' (outside every timed region)
    try:
        from deepaco_b200.tsp.aco import ACO
        from oracle import aco_torch as O
        T_par = 5
        torch.manual_seed(2024)
        mine = ACO(dist[0], n_ants=N_ANTS, heuristic=heu[0], device=dev)
        low = float(mine.run(T_par))
        torch.manual_seed(2024)
        ref = O.TspColony(dist[0], N_ANTS, heuristic=heu[0])
        ref_low = float(ref.run(T_par))
        line["parity"] = {"iterations": T_par, "best_cost": low, "reference_best_cost": ref_low,
                          "best_cost_gap": abs(low - ref_low), "pheromone_identical": bool(torch.equal(mine.pheromone, ref.pheromone)),
                          "best_tour_identical": bool(torch.equal(mine.shortest_path, ref.shortest_path))}
    except Exception as exc:      # the oracle is test infrastructure; its absence must not break the bench
        line["parity"] = {"error": str(exc)[:200]}
    try:
        os.sched_setaffinity(0, all_cpus)       # the CPU baseline may use every host core
    except Exception:
        pass
    if not args.no_cpu_baseline:
        line["reference_cuda"] = L.guarded("reference_cuda", L.leg_reference_cuda, dev)
        cb, _, _ = cpu_reference_throughput(100, 2)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist_pg.barrier()
        dist_pg.destroy_process_group()


if __name__ == "__main__":
    main()
