"""The cost / neighbour-table / evaporate + deposit kernel SOURCE (deepaco_b200/csrc/tsp_update.cuh) without a GPU:
tests/cpu_emu compiles the same text for the host.  The ordered deposit is arithmetic whose result does not depend on
the device -- the reference adds 1/cost per ant in ant order with unique indices per statement -- so the kernel source
must give the oracle's pheromone matrix bit for bit (all variants: plain, elitist, min-max); tour costs follow ATen's
CUDA summation order and are compared with the CPU sum to fp32 rounding, and the three cost kernels with each other
exactly.  The sm_100a build of the same source is checked on the GPU by tests/test_gpu_tsp.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import aco_torch as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cpu_emu")
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


@pytest.fixture(scope="session")
def emu_u():
    subprocess.run(["make", "-C", EMU_DIR, "-s", "_build/libtsp_update_emu.so"], check=True)
    h = ctypes.CDLL(os.path.join(EMU_DIR, "_build", "libtsp_update_emu.so"))
    h.emu_tsp_cost.restype = ctypes.c_char_p
    h.emu_tsp_cost.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp, ci]
    h.emu_tsp_update.restype = ctypes.c_char_p
    h.emu_tsp_update.argtypes = [vp, vp, vp, ci, ci, cf, ci, ci, cf, vp, vp, vp, vp]
    return h


def _ptr(t):
    return None if t is None else t.data_ptr()


def _instance(n, A, seed):
    torch.manual_seed(seed)
    xy = torch.rand(n, 2)
    dist = torch.norm(xy[:, None] - xy, dim=2, p=2)
    dist[torch.arange(n), torch.arange(n)] = 1e9
    paths = torch.stack([torch.randperm(n) for _ in range(A)], dim=1).contiguous()          # [n, A] int64
    return dist.contiguous(), paths


def _costs(emu_u, dist, paths, mode):
    from deepaco_b200 import _engine as E
    n, A = paths.shape
    bw, vec, _ = E.aten_sum_plan(n, A)
    lbw = int(np.log2(min(bw, 32)))
    costs = torch.full((A,), float("nan"))
    nbr = torch.full((n, A), -1, dtype=torch.int32)
    tours = paths.T.contiguous().to(torch.int16)
    err = emu_u.emu_tsp_cost(_ptr(dist), _ptr(paths) if mode == "paths" else None, None if mode == "paths" else _ptr(tours), n, A, lbw,
                             int(vec), _ptr(costs), _ptr(nbr), int(mode == "tile"))
    assert err is None, err
    return costs, nbr


@pytest.mark.parametrize("n,A", [(20, 8), (100, 40), (130, 33), (200, 16)])
def test_cost_kernels_on_host(emu_u, n, A):
    dist, paths = _instance(n, A, n)
    want = O.tsp_path_costs(dist, paths)
    got = {m: _costs(emu_u, dist, paths, m) for m in ("paths", "tours", "tile")}
    for m, (c, nbr) in got.items():
        assert torch.allclose(c, want, rtol=1e-6), m
        assert torch.equal(c, got["paths"][0]), m                         # same summation order in all three kernels
        assert torch.equal(nbr, got["paths"][1]), m
    nbr = got["tile"][1].to(torch.int64) & 0xffffffff
    prev, nxt = torch.roll(paths, 1, dims=0), torch.roll(paths, -1, dims=0)
    cols = torch.arange(A).expand(n, A)
    assert torch.equal(nbr[paths, cols], (prev << 16) | nxt)              # N[u][a] = pred << 16 | succ of node u in tour a


@pytest.mark.parametrize("kw", [{}, {"elitist": True}, {"min_max": True}, {"min_max": True, "scale": 1.7}])
@pytest.mark.parametrize("n,A", [(20, 8), (100, 48), (61, 130)])
def test_update_kernel_on_host_is_bit_identical_to_the_reference_ops(emu_u, n, A, kw):
    dist, paths = _instance(n, A, 7 * n + A)
    torch.manual_seed(1)
    ph0 = (torch.rand(n, n) + 0.2).contiguous()
    costs = O.tsp_path_costs(dist, paths).contiguous()
    _, nbr = _costs(emu_u, dist, paths, "tile")
    elitist, min_max, scale = kw.get("elitist", False), kw.get("min_max", False), kw.get("scale")
    ph_min = 0.1
    ph_max = torch.tensor([float(n / costs.min())])
    ref_in = ph0 * scale if scale else ph0                                 # MMAS rescale happens before the update (tsp/aco.py:86-87)
    want = O.tsp_update_pheromone(ref_in.clone(), paths, costs, decay=0.9, elitist=elitist, min_max=min_max, ph_min=ph_min,
                                  ph_max=float(ph_max))
    ph = ph0.clone()
    heu = torch.rand(n, n).contiguous()
    prod = torch.full((n, n), float("nan"))
    sc = torch.tensor([scale], dtype=torch.float32) if scale else None
    err = emu_u.emu_tsp_update(_ptr(ph), _ptr(nbr), _ptr(costs), n, A, 0.9, int(elitist), int(min_max), ph_min,
                               _ptr(ph_max) if min_max else None, _ptr(sc), _ptr(heu), _ptr(prod))
    assert err is None, err
    assert torch.equal(ph, want)
    assert torch.equal(prod, ph * heu)                                     # fused product for the next construction


@pytest.mark.parametrize("kw", [{}, {"min_max": True}, {"min_max": True, "scale": 1.7}])
@pytest.mark.parametrize("n,A,CH,W", [(20, 8, 8, 2), (61, 130, 48, 8), (40, 300, 64, 3), (100, 257, 8192, 8)])
def test_row_update_kernel_on_host_is_bit_identical_to_the_reference_ops(emu_u, n, A, CH, W, kw):
    """tsp_update_row_kernel (many ants: one CTA per row, ants in chunks of CH split over W warps): same bits as the
    reference's sequential per-ant deposit, incl. ragged chunk / warp splits."""
    emu_u.emu_tsp_update_rows.restype = ctypes.c_char_p
    emu_u.emu_tsp_update_rows.argtypes = [vp, vp, vp, ci, ci, ci, ci, cf, ci, cf, vp, vp, vp, vp]
    dist, paths = _instance(n, A, 3 * n + A)
    torch.manual_seed(2)
    ph0 = (torch.rand(n, n) + 0.2).contiguous()
    costs = O.tsp_path_costs(dist, paths).contiguous()
    _, nbr = _costs(emu_u, dist, paths, "tile")
    min_max, scale = kw.get("min_max", False), kw.get("scale")
    ph_max = torch.tensor([float(n / costs.min())])
    ref_in = ph0 * scale if scale else ph0
    want = O.tsp_update_pheromone(ref_in.clone(), paths, costs, decay=0.9, elitist=False, min_max=min_max, ph_min=0.1,
                                  ph_max=float(ph_max))
    ph = ph0.clone()
    heu = torch.rand(n, n).contiguous()
    prod = torch.full((n, n), float("nan"))
    sc = torch.tensor([scale], dtype=torch.float32) if scale else None
    err = emu_u.emu_tsp_update_rows(_ptr(ph), _ptr(nbr), _ptr(costs), n, A, CH, W, 0.9, int(min_max), 0.1,
                                    _ptr(ph_max) if min_max else None, _ptr(sc), _ptr(heu), _ptr(prod))
    assert err is None, err
    assert torch.equal(ph, want)
    assert torch.equal(prod, ph * heu)


@pytest.mark.parametrize("kw", [{}, {"elitist": True}, {"min_max": True}, {"min_max": True, "scale": 1.7}])
@pytest.mark.parametrize("n,A,threads", [(5, 3, 32), (20, 8, 32), (20, 9, 64), (100, 48, 128), (100, 21, 256), (61, 130, 64), (130, 19, 256)])
def test_seq_update_kernel_on_host_is_bit_identical_to_the_reference_ops(emu_u, n, A, threads, kw):
    """tsp_update_seq_kernel (matrix in shared memory, ants one after another with a barrier each, tours prefetched eight
    ahead): same bits as the reference's per-ant index_put loop, incl. elitist / min-max / MMAS rescale, partial
    prefetch windows (A not a multiple of 8) and thread counts not equal to n."""
    emu_u.emu_tsp_update_seq.restype = ctypes.c_char_p
    emu_u.emu_tsp_update_seq.argtypes = [vp, vp, vp, ci, ci, ci, cf, ci, ci, cf, vp, vp, vp, vp]
    dist, paths = _instance(n, A, 5 * n + A)
    torch.manual_seed(3)
    ph0 = (torch.rand(n, n) + 0.2).contiguous()
    costs = O.tsp_path_costs(dist, paths).contiguous()
    elitist, min_max, scale = kw.get("elitist", False), kw.get("min_max", False), kw.get("scale")
    ph_max = torch.tensor([float(n / costs.min())])
    ref_in = ph0 * scale if scale else ph0
    want = O.tsp_update_pheromone(ref_in.clone(), paths, costs, decay=0.9, elitist=elitist, min_max=min_max, ph_min=0.1,
                                  ph_max=float(ph_max))
    ph = ph0.clone()
    heu = torch.rand(n, n).contiguous()
    prod = torch.full((n, n), float("nan"))
    sc = torch.tensor([scale], dtype=torch.float32) if scale else None
    tours = paths.T.contiguous().to(torch.int16)
    err = emu_u.emu_tsp_update_seq(_ptr(ph), _ptr(tours), _ptr(costs), n, A, threads, 0.9, int(elitist), int(min_max), 0.1,
                                   _ptr(ph_max) if min_max else None, _ptr(sc), _ptr(heu), _ptr(prod))
    assert err is None, err
    assert torch.equal(ph, want)
    assert torch.equal(prod, ph * heu)


@pytest.mark.parametrize("kw", [{}, {"elitist": True}, {"min_max": True}])
@pytest.mark.parametrize("n,A", [(20, 8), (100, 40), (61, 130), (136, 40), (33, 16), (32, 17), (127, 33), (128, 20)])
def test_tail_kernel_on_host_equals_cost_best_update_sequence(emu_u, n, A, kw):
    """tsp_tail_kernel (cost + best tracking + MMAS bookkeeping + ant-sequential update in one launch) over three
    iterations with fresh tours each: costs equal to the cost kernel's bits, lowest cost / best tour / MMAS max and the
    pheromone equal to the reference's run loop (tsp/aco.py:76-90) restated with the oracle ops."""
    emu_u.emu_tsp_tail.restype = ctypes.c_char_p
    emu_u.emu_tsp_tail.argtypes = [vp] * 9 + [ci, ci, cf, ci, ci, cf, ci, ci]
    from deepaco_b200 import _engine as E
    elitist, min_max = kw.get("elitist", False), kw.get("min_max", False)
    torch.manual_seed(n + A)
    heu = (torch.rand(n, n) + 0.1).contiguous()
    ph = (torch.ones(n, n) * (0.1 if min_max else 1.0)).contiguous()
    ref_ph = ph.clone()
    lowest, shortest, ph_max = torch.tensor([float("inf")]), torch.zeros(n, dtype=torch.int64), torch.zeros(1)
    ref_low, ref_short, ref_max = float("inf"), None, None
    bw, vec, _ = E.aten_sum_plan(n, A)
    lbw = int(np.log2(min(bw, 32)))
    for it in range(3):
        dist, paths = _instance(n, A, 31 * it + n)
        tours = paths.T.contiguous().to(torch.int16)
        want_costs, _ = _costs(emu_u, dist, paths, "tile")
        costs = torch.full((A,), float("nan"))
        prod = torch.full((n, n), float("nan"))
        err = emu_u.emu_tsp_tail(_ptr(ph), _ptr(tours), _ptr(dist), _ptr(heu), _ptr(prod), _ptr(costs), _ptr(lowest), _ptr(shortest),
                                 _ptr(ph_max), n, A, 0.9, int(elitist), int(min_max), 0.1, lbw, int(vec))
        assert err is None, err
        assert torch.equal(costs, want_costs)
        # the reference's run loop on the same costs (tsp/aco.py:79-90)
        best_cost, best_idx = want_costs.min(dim=0)
        if best_cost < ref_low:
            ref_short, ref_low = paths[:, best_idx].clone(), best_cost.clone()
            if min_max:
                new_max = n / ref_low                                   # Tensor.__rtruediv__: reciprocal * n
                if ref_max is None:
                    ref_ph = ref_ph * (new_max / ref_ph.max())
                ref_max = new_max
        ref_ph = O.tsp_update_pheromone(ref_ph, paths, want_costs, decay=0.9, elitist=elitist, min_max=min_max, ph_min=0.1,
                                        ph_max=ref_max)
        assert torch.equal(ph, ref_ph), it
        assert float(lowest) == float(ref_low) and torch.equal(shortest, ref_short)
        assert torch.equal(prod, ph * heu)
        if min_max:
            assert float(ph_max) == float(ref_max)


@pytest.mark.parametrize("n,rows", [(33, 5), (100, 100), (256, 19)])
def test_knn_refresh_kernel_on_host_selects_the_32_largest_per_row(emu_u, n, rows):
    """knn_refresh_kernel: per row the columns of the 32 largest product entries, ties to the lower column (the floor
    entries of a sparse product are all equal), 32 distinct columns in any order."""
    emu_u.emu_knn_refresh.restype = ctypes.c_char_p
    emu_u.emu_knn_refresh.argtypes = [vp, vp, ci, ci]
    torch.manual_seed(n)
    prod = torch.full((rows, n), 1e-10)
    for r in range(rows):                                   # 5 .. 40 distinct large entries per row, the rest tied at the floor
        k = 5 + (7 * r) % 36
        prod[r, torch.randperm(n)[:k]] = torch.rand(k) + 0.01
    prod = prod.contiguous()
    knn = torch.full((rows, 32), 255, dtype=torch.uint8)
    assert emu_u.emu_knn_refresh(_ptr(prod), _ptr(knn), n, rows) is None
    order = torch.argsort(prod, dim=1, descending=True, stable=True)[:, :32]        # stable: ties keep column order
    assert torch.equal(torch.sort(knn.long(), dim=1).values, torch.sort(order, dim=1).values)


# ---- CVRP: open-path cost, one-directional deposit, repeated (0, 0) pairs count once, 1e-10 floor ------------------
@pytest.fixture(scope="session")
def emu_c(emu_u):
    emu_u.emu_cvrp_cost.restype = ctypes.c_char_p
    emu_u.emu_cvrp_cost.argtypes = [vp, vp, vp, ci, ci, ci, ci, vp, vp]
    emu_u.emu_cvrp_update.restype = ctypes.c_char_p
    emu_u.emu_cvrp_update.argtypes = [vp, vp, vp, ci, ci, cf, ci, ci, cf, vp, vp]
    return emu_u


def _cvrp_paths(customers, A, seed):
    torch.manual_seed(seed)
    loc = torch.cat((torch.tensor([[0.5, 0.5]]), torch.rand(customers, 2)))
    dist = torch.norm(loc[:, None] - loc, dim=2, p=2)
    N = customers + 1
    dist[torch.arange(N), torch.arange(N)] = 1e-10
    demand = torch.cat((torch.zeros(1), torch.randint(1, 10, (customers,)).float()))
    heu = torch.rand(N, N) * 0.98 + 1e-10
    paths = O.cvrp_gen_path(torch.ones_like(dist), heu, demand, 50, A)      # the reference's construction ops on CPU
    return dist.contiguous(), paths.contiguous()


@pytest.mark.parametrize("kw", [{}, {"elitist": True}, {"min_max": True}])
@pytest.mark.parametrize("customers,A", [(12, 8), (40, 33)])
def test_cvrp_cost_and_update_kernels_on_host(emu_c, customers, A, kw):
    dist, paths = _cvrp_paths(customers, A, 100 + customers)
    N, rows = customers + 1, paths.shape[0]
    T = rows - 1
    assert bool((paths[-1] == 0).all()) and bool(((paths[:-1] == 0) & (paths[1:] == 0)).any())   # padded (0, 0) pairs present
    want_c = O.cvrp_path_costs(dist, paths)
    costs = torch.full((A,), float("nan"))
    nbr = torch.zeros((N, A), dtype=torch.int32)
    assert emu_c.emu_cvrp_cost(_ptr(dist), _ptr(paths), None, N, A, rows, T, _ptr(costs), _ptr(nbr)) is None
    assert torch.allclose(costs, want_c, rtol=1e-6)
    tours = paths.T.contiguous().to(torch.int16)
    costs2, nbr2 = torch.full((A,), float("nan")), torch.zeros((N, A), dtype=torch.int32)
    assert emu_c.emu_cvrp_cost(_ptr(dist), None, _ptr(tours), N, A, rows, T, _ptr(costs2), _ptr(nbr2)) is None
    assert torch.equal(costs2, costs) and torch.equal(nbr2, nbr)

    elitist, min_max = kw.get("elitist", False), kw.get("min_max", False)
    torch.manual_seed(2)
    ph0 = (torch.rand(N, N) * 0.5 + 1e-11).contiguous()                     # some cells below the 1e-10 floor after decay
    ph0[3, 4] = 5e-11
    ph_min, ph_max = 0.05, torch.tensor([0.9])
    c_in = want_c.contiguous()
    want = O.cvrp_update_pheromone(ph0.clone(), paths, c_in, decay=0.9, elitist=elitist, min_max=min_max, ph_min=ph_min,
                                   ph_max=float(ph_max))
    ph = ph0.clone()
    err = emu_c.emu_cvrp_update(_ptr(ph), _ptr(nbr), _ptr(c_in), N, A, 0.9, int(elitist), int(min_max), ph_min,
                                _ptr(ph_max) if min_max else None, None)
    assert err is None, err
    assert torch.equal(ph, want)


# ---- analytic gradient of the sampled log-probabilities (REINFORCE) vs autograd through the reference's ops ----------
@pytest.mark.parametrize("problem", ["tsp", "cvrp"])
def test_logp_backward_kernel_on_host_matches_autograd_through_the_reference_ops(emu_u, problem):
    emu_u.emu_logp_backward.restype = ctypes.c_char_p
    emu_u.emu_logp_backward.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, cf, vp, vp]
    torch.manual_seed(5)
    n, A = 30, 16
    ph = (torch.rand(n, n) + 0.5).requires_grad_(True)
    heu0 = torch.rand(n, n) * 0.9 + 0.05
    heu0[torch.rand(n, n) < 0.3] = 1e-10                   # the clamp region (p < eps) passes no gradient
    heu = heu0.clone().requires_grad_(True)
    demand = None
    torch.manual_seed(9)
    if problem == "tsp":
        paths, logp = O.tsp_gen_path(ph, heu, A, require_prob=True)
    else:
        demand = torch.cat((torch.zeros(1), torch.randint(1, 10, (n - 1,)).float())).contiguous()
        paths, logp = O.cvrp_gen_path(ph, heu, demand, 30, A, require_prob=True)
    g = torch.randn_like(logp)
    (logp * g).sum().backward()
    g_heu, g_ph = torch.zeros(n, n), torch.zeros(n, n)
    paths_c, g_c = paths.contiguous(), g.contiguous()
    err = emu_u.emu_logp_backward(_ptr(ph.detach().contiguous()), _ptr(heu.detach().contiguous()), _ptr(paths_c), _ptr(g_c), n, A,
                                  paths.shape[0], _ptr(demand), 30.0, _ptr(g_heu), _ptr(g_ph))
    assert err is None, err
    scale = float(heu.grad.abs().max())
    assert torch.allclose(g_heu, heu.grad, rtol=1e-4, atol=1e-6 * scale)
    assert torch.allclose(g_ph, ph.grad, rtol=1e-4, atol=1e-6 * float(ph.grad.abs().max()))
