"""Additional legs of bench.py: the other BASELINE.json configs (C3, C4, C5, dense-heuristic C2), the ant-sharded
colony (north_star: ants shard across the GPUs, one exchange per ACO iteration) and the reference op sequence with
device='cuda' on the same GPU.  Every leg returns a dict that goes into the one JSON line under `configs` /
`ant_sharded` / `reference_cuda`; a leg that fails returns {"error": ...} on every rank instead of taking the line down.

All timings: CUDA events on the launching stream, >= 3 warm-up calls, max over ranks for the multi-GPU legs.  At these
sizes every matrix is L2-resident whether or not L2 is flushed (the roofline fractions are throughput-normalised
figures); the legs run their steps back to back.
"""
import json
import os
import time

import torch
import torch.distributed as dist_pg

from deepaco_b200 import _engine as E
from deepaco_b200 import dist as D
from deepaco_b200.heuristics import load_net, tsp_heuristic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s"


def k1_bytes_per_tour(n, steps, rows_per_step):
    """S * R * 4n + 8 (S + 1)."""
    return steps * rows_per_step * 4 * n + 8 * (steps + 1)


def timed(fn, iters, warm=3):
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


def roofline(kernel, bytes_per_launch, kernel_ms, note=None):
    peak, src = hbm_peak()
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": None, "peak_source": src, "algorithmic_bytes_per_launch": int(bytes_per_launch), "kernel_ms": kernel_ms}
    if note:
        out["note"] = note
    return out


def tsp_instances(B, n, seed, dev):
    g = torch.Generator(device="cpu").manual_seed(seed)
    coords = torch.rand((B, n, 2), generator=g).to(dev)
    d = torch.cdist(coords, coords)
    idx = torch.arange(n, device=dev)
    d[:, idx, idx] = 1e9
    return coords, d.contiguous()


def _world():
    return (dist_pg.get_rank(), dist_pg.get_world_size()) if dist_pg.is_initialized() else (0, 1)


def _max_over_ranks(ms, dev):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist_pg.is_initialized() and dist_pg.get_world_size() > 1:
        dist_pg.all_reduce(t, op=dist_pg.ReduceOp.MAX)
    return float(t)


def _sync_all():
    if dist_pg.is_initialized() and dist_pg.get_world_size() > 1:
        dist_pg.barrier()
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------------------------
def leg_c2_dense(dev, steps):
    """C2 with the vanilla heuristic 1/dist (dense product -> list kernel): 64 colonies x 512 ants x TSP-100."""
    B, n, A = 64, 100, 512
    _, d = tsp_instances(B, n, 77, dev)
    r = E.TspRunner(d, (1.0 / d).contiguous(), torch.ones_like(d), A)
    offs = [b * 4_000_000 for b in range(B)]
    evs = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    state = {"it": 0, "k1": []}

    def step(measure=False):
        r.run(1, 1234, state["it"] * r.increment, offs, sample_events=evs if measure else None)
        state["it"] += 1
        if measure:
            torch.cuda.synchronize()
            state["k1"].append(evs[0].elapsed_time(evs[1]))

    ms = timed(step, steps)
    for _ in range(5):
        step(True)
    k1 = sorted(state["k1"])[len(state["k1"]) // 2]
    return {"workload": f"{B} colonies x TSP-{n} x {A} ants, heuristic 1/dist (dense), 1 ACO iteration per step",
            "ms_per_iteration": ms, "value": B * A / (ms * 1e-3), "unit": "ant-tours/s",
            "roofline": roofline("K1 aco_list_kernel (dense product, unvisited list in registers)",
                                 k1_bytes_per_tour(n, n - 1, 2) * B * A, k1,
                                 "4 Philox4x32-10 per lane per step are the floor of a dense row (one per candidate column)")}


def leg_c3(dev, steps):
    """C3: TSP-NLS n=500, 256 ants, pretrained tsp_nls500 heuristic on the k=50 graph; construction, batched 2-opt
    (n/4 = 125 passes), NLS (T_nls=10, T_p=20) and the whole ACO iteration with NLS (tsp_nls/aco.py:104-129)."""
    n, A, k = 500, 256, 50
    coords, d = tsp_instances(1, n, 500, dev)
    heu, how = tsp_heuristic(coords, d, k, kind="tsp_nls")
    ph = torch.ones_like(d)
    knn = E.sparse_candidates(heu)
    out = {}

    def samp():
        out["t"] = E.tsp_sample(ph, heu, A, start_node=0, double_norm=True, seed=1, want_paths=False, want_tours=True, knn=knn)[2]

    it = max(3, min(steps, 10))
    ms_s = timed(samp, it)
    hd = (1 / (heu[0] / heu[0].max(-1, keepdim=True).values + 1e-5)).contiguous()
    base = out["t"].clone()
    passes = {}

    def topt():
        t = base.clone()
        _, passes["p"] = E.two_opt_(d, t, n // 4, want_passes=True)

    def nls():
        t = base.clone()
        E.tsp_nls_(d, hd[None], t, n // 4)

    ms_t = timed(topt, max(2, it // 2), 1)
    ms_n = timed(nls, 2, 1)
    total_passes = int(passes["p"].sum().item())
    cand = (n - 2) * (n - 1) // 2
    r = E.TspRunner(d, heu, ph, A, start_node=0, double_norm=True)
    r.set_local_search("nls", n // 4, hd[None])
    st = {"it": 0}

    def iteration():
        r.run(1, 1234, st["it"] * r.increment)
        st["it"] += 1

    ms_i = timed(iteration, 3, 1)
    peak, src = hbm_peak()
    return {"workload": f"TSP-{n} x {A} ants, {how}, start node 0", "sample_ms": ms_s, "two_opt_ms": ms_t, "nls_ms": ms_n,
            "ms_per_iteration": ms_i, "value": A / (ms_i * 1e-3), "unit": "ant-tours/s (construction + NLS + cost + update)",
            "roofline": {"sample": roofline("K1 construction at n=500 (product rows from L2)", k1_bytes_per_tour(n, n - 1, 1) * A, ms_s),
                         "two_opt": {"bound": "issue", "kernel": "K4 two_opt_kernel", "passes": total_passes,
                                     "candidates_per_s": total_passes * cand / (ms_t * 1e-3),
                                     "nominal_gbs": total_passes * cand * 16 / (ms_t * 1e-3) / 1e9, "peak": peak, "peak_source": src,
                                     "note": "register-carry kernel, one shared-memory gather per candidate: the 16 B / candidate "
                                             "model of SURVEY 8d exceeds HBM peak by construction; instructions per candidate and "
                                             "issue utilisation are in profiles/"}}}


def cvrp_instances(B, n_customers, seed, dev):
    """cvrp/utils.py:9-22 gen_instance: depot (0.5, 0.5), demand U{1..9}, distances with 1e-10 diagonal."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    loc = torch.rand((B, n_customers, 2), generator=g)
    dem = torch.randint(1, 10, (B, n_customers), generator=g).float()
    loc = torch.cat((torch.full((B, 1, 2), 0.5), loc), dim=1).to(dev)
    dem = torch.cat((torch.zeros(B, 1), dem), dim=1).to(dev)
    N = n_customers + 1
    d = torch.cdist(loc, loc)
    idx = torch.arange(N, device=dev)
    d[:, idx, idx] = 1e-10
    return dem.contiguous(), d.contiguous()


def leg_c4(dev, steps):
    """C4: CVRP-100 x 512 ants, 64 colonies per step, pretrained cvrp100 heuristic on the complete graph."""
    B, nc, A = 64, 100, 512
    N = nc + 1
    dem, d = cvrp_instances(B, nc, 4040, dev)
    net = load_net("cvrp", dev)
    with torch.no_grad():
        heu = net.dense_heuristic_matrices(dem[:, :, None], d, 1e-10).contiguous()
    r = E.CvrpRunner(d, dem, heu, torch.ones_like(d), A)
    offs = torch.tensor([b * 4_000_000 for b in range(B)], dtype=torch.int64, device=dev)

    def step():
        r.run(1, 1234, offs)

    ms = timed(step, steps)
    tm = r.tmax.to(torch.float64)
    bytes_launch = float((A * (tm * 2 * 4 * N + 8 * (tm + 1))).sum().item())
    return {"workload": f"{B} colonies x CVRP-{nc} x {A} ants, Net(pretrained cvrp100) heuristic, capacity 50, 1 ACO iteration per step",
            "ms_per_iteration": ms, "value": B * A / (ms * 1e-3), "unit": "ant-routes/s", "mean_path_rows": float(tm.mean().item()) + 1,
            "roofline": roofline("K1 aco_list_kernel<CVRP> (+ cost, best, update: whole iteration timed)", bytes_launch, ms,
                                 "denominator is the WHOLE iteration (CvrpRunner has no per-kernel events): a lower bound for K1")}


def leg_c5(dev, steps, T=10):
    """C5: 64 x TSP-200 x 256 ants split over the ranks by colonies (strong scaling), T ACO iterations per step and the
    gather of every colony's best cost / tour (one NCCL collective) inside the timed region."""
    rank, world = _world()
    Btot, n, A, k = 64, 200, 256, 20
    coords, d = tsp_instances(Btot, n, 2005, dev)
    b0, nb = D.shard_range(Btot, world, rank)
    counts = [D.shard_range(Btot, world, r)[1] for r in range(world)]
    heu, how = tsp_heuristic(coords[b0:b0 + nb], d[b0:b0 + nb], k)
    r = E.TspRunner(d[b0:b0 + nb].contiguous(), heu.contiguous(), torch.ones((nb, n, n), device=dev), A)
    offs = torch.tensor([(b0 + b) * 40_000_000 for b in range(nb)], dtype=torch.int64, device=dev)
    st = {"it": 0, "out": None}

    def step():
        r.run(T, 1234, st["it"] * r.increment, offs)
        st["it"] += T
        st["out"] = D.gather_colony_results_packed(r.lowest_cost, r.shortest_path, counts)

    for _ in range(3):
        step()
    _sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    _sync_all()
    ms = _max_over_ranks(e0.elapsed_time(e1), dev) / steps
    low, sp = st["out"]
    assert low.shape[0] == Btot and sp.shape == (Btot, n)
    return {"workload": f"{Btot} x TSP-{n} x {A} ants ({how}), colonies split over {world} GPU(s): {nb} per GPU; "
                        f"{T} ACO iterations + gather of best costs / tours per step",
            "scaling": "strong", "n_gpus": world, "iterations_per_step": T, "collectives_per_step": 1 if world > 1 else 0,
            "ms_per_step": ms, "ms_per_iteration": ms / T, "value": Btot * A * T / (ms * 1e-3), "unit": "ant-tours/s",
            "mean_best_cost": float(low.mean().item())}


def leg_ant_sharded(dev, steps, T=10, n=200, A=16384, k=20):
    """One colony with enough ants to fill one GPU several times over (TSP-200 x 16384 ants), ants split over the ranks:
    deepaco_tsp_run_shard (fused peer stores over NVLink + flag barrier, no NCCL on the data path, no host sync), against
    the same colony on ONE GPU (rank 0, deepaco_tsp_run) -- time and bits."""
    rank, world = _world()
    coords, d = tsp_instances(1, n, 8192, dev)
    heu, how = tsp_heuristic(coords, d, k)
    ph0 = torch.ones_like(d)
    total_T = T * (steps + 3)
    # single GPU (every rank computes it: keeps the ranks in lockstep and gives each the reference bits)
    single = E.TspRunner(d, heu, ph0, A)
    st = {"it": 0}

    def step1():
        single.run(T, 1234, st["it"] * single.increment)
        st["it"] += T

    for _ in range(3):
        step1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step1()
    e1.record()
    torch.cuda.synchronize()
    ms_single = e0.elapsed_time(e1) / (steps * T)
    out = {"workload": f"one colony TSP-{n} x {A} ants ({how}), {T} ACO iterations per step", "n_gpus": world,
           "single_gpu_ms_per_iteration": ms_single, "single_gpu_value": A / (ms_single * 1e-3), "unit": "ant-tours/s"}
    peer = D.symmetric_peer_memory(1, A, n, dev) if world > 1 else D.local_peer_memory(1, A, n, dev, 1)[0]
    col = D.DeviceShardedColony(E.TspRunner(d, heu, ph0, A), peer, timeout_ms=20000)
    st2 = {"it": 0}

    def step2():
        col.run(T, 1234, st2["it"] * col.runner.increment)
        st2["it"] += T

    for _ in range(3):
        step2()
    _sync_all()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(steps):
        step2()
    s1.record()
    _sync_all()
    col.check()
    ms = _max_over_ranks(s0.elapsed_time(s1), dev) / (steps * T)
    same = bool(torch.equal(col.runner.pheromone, single.pheromone) and torch.equal(col.runner.lowest_cost, single.lowest_cost)
                and torch.equal(col.runner.shortest_path, single.shortest_path)) and st["it"] == st2["it"] == total_T
    flag = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist_pg.all_reduce(flag, op=dist_pg.ReduceOp.MIN)
    out.update({"ms_per_iteration": ms, "value": A / (ms * 1e-3), "speedup_vs_single_gpu": ms_single / ms,
                "ants_per_gpu": col.count, "collectives_per_iteration": 0,
                "exchange": "tours stored by the sampling kernel into every rank's peer-mapped buffer (NVLink st.global), "
                            "1 flag barrier (release/acquire.sys) per iteration, replay of cost + ordered deposit on every rank",
                "exchange_bytes_per_iteration_per_gpu": col.count * n * 2 * (world - 1),
                "identical_to_single_gpu": bool(int(flag.item()))})
    return out


def leg_reference_cuda(dev, iters=3):
    """The reference's op sequence (oracle/aco_torch.py = tsp/aco.py op for op) with device='cuda' on this GPU: one C2
    colony (TSP-100 x 512 ants).  A reported baseline like cpu_baseline; rank 0 only."""
    from oracle import aco_torch as O
    coords, d = tsp_instances(1, 100, 1234, dev)
    heu, how = tsp_heuristic(coords, d, 20)
    torch.manual_seed(1234)
    col = O.TspColony(d[0], 512, heuristic=heu[0])
    col.run(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    col.run(iters)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    return {"value": 512 / dt, "unit": "ant-tours/s", "ms_per_iteration": dt * 1e3, "kind": "port",
            "sample": f"{iters} ACO iterations of one TSP-100 colony x 512 ants, reference op sequence on cuda (torch {torch.__version__}), {how}"}


def leg_gnn_front_end(dev, steps):
    """K3: instance -> kNN graph -> heuristic network (eval) -> dense heuristic matrix for the batch the ACO iteration
    consumes (256 x TSP-100, k = 20), one launch of deepaco_gnn_forward (tensor-core tiles) after a batched topk."""
    B, n, k = 256, 100, 20
    coords, d = tsp_instances(B, n, 1234, dev)
    net = load_net("tsp", dev)
    ms = timed(lambda: net.heuristic_matrices(coords, d, k), max(3, min(steps, 10)))
    E = n * k
    flops = B * (12 * (2 * E * 32 * 32 + 4 * 2 * n * 32 * 32) + 2 * 2 * E * 32 * 32)
    bytes_alg = B * 12 * (2 * 128 * E + 3 * 128 * E + 20 * 1024)          # edge state r/w + gathers + weights
    out = {"workload": f"{B} x TSP-{n}, k={k}: k-nearest-neighbour graph (deepaco_knn_graph) + Net.forward (eval) + reshape + EPS", "ms_per_batch": ms,
           "us_per_instance": ms / B * 1e3, "tflops_linear": flops / (ms * 1e-3) / 1e12,
           "roofline": roofline("K3 gnn_forward_kernel (mma.sync TF32 split tiles)", bytes_alg, ms,
                                "edge state streams through L2 once per layer; whole front end timed (graph launch included)")}
    return out


def guarded(name, fn, *args, **kw):
    """Run a leg; an exception becomes {"error": ...}.  Collective-bearing legs raise symmetrically on all ranks or not
    at all (their failure modes -- missing checkpoint, symmetric memory unavailable -- do not depend on the rank)."""
    try:
        return fn(*args, **kw)
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
