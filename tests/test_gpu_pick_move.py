"""ACO.pick_move and the CVRP step helpers (reference tsp/aco.py:165-177, cvrp/aco.py:167-205): driving the
construction step by step through them reproduces, bit for bit, what the reference's gen_path does with the same torch
seed on this GPU -- checked against the oracle (reference ops on the device) and against the fused construction kernel."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tsp_instance(n, k, seed):
    torch.manual_seed(seed)
    xy = torch.rand(n, 2, device=DEV)
    dist = torch.norm(xy[:, None] - xy, dim=2, p=2)
    dist[torch.arange(n), torch.arange(n)] = 1e9
    _, idx = torch.topk(dist, k, dim=1, largest=False)
    heu = torch.full_like(dist, 1e-10).scatter_(1, idx, torch.rand(n, k, device=DEV) * 0.9 + 0.05)
    return dist, heu


@pytest.mark.parametrize("n,A", [(20, 32), (100, 64), (200, 16)])
def test_tsp_stepwise_pick_move_equals_reference_gen_path(n, A):
    from deepaco_b200.tsp.aco import ACO
    from oracle import aco_torch as O
    dist, heu = _tsp_instance(n, 10, n)
    ph = torch.rand(n, n, device=DEV) + 0.5
    aco = ACO(dist, n_ants=A, heuristic=heu, pheromone=ph, device=DEV)

    torch.manual_seed(5)
    start = torch.randint(low=0, high=n, size=(A,), device=DEV)                     # tsp/aco.py:141
    mask = torch.ones(A, n, device=DEV)
    mask[torch.arange(A, device=DEV), start] = 0
    rows, logps, prev = [start], [], start
    for _ in range(n - 1):
        actions, lp = aco.pick_move(prev, mask, True)
        rows.append(actions)
        logps.append(lp)
        mask = mask.clone()
        mask[torch.arange(A, device=DEV), actions] = 0
        prev = actions
    paths, logp = torch.stack(rows), torch.stack(logps)
    offset_after = torch.cuda.default_generators[0].get_offset()

    torch.manual_seed(5)
    want_paths, want_logp = O.tsp_gen_path(ph, heu, A, require_prob=True)          # the reference's ops on this GPU
    assert torch.cuda.default_generators[0].get_offset() == offset_after            # same generator consumption
    assert torch.equal(paths, want_paths)
    from deepaco_b200 import _engine as E
    _, _, exact = E.aten_sum_plan(n, A)          # False where ATen widens the reduction block past one warp
    if exact:
        assert torch.equal(logp, want_logp)
    else:
        assert torch.allclose(logp, want_logp, rtol=0, atol=1e-6)

    torch.manual_seed(5)                                                           # and the fused kernel agrees
    fused_paths, fused_logp = aco.gen_path(require_prob=True)
    assert torch.equal(paths, fused_paths)
    assert torch.allclose(logp, fused_logp, rtol=0, atol=0 if exact else 1e-6)


def test_cvrp_stepwise_helpers_equal_fused_gen_path():
    from deepaco_b200.cvrp.aco import ACO
    from deepaco_b200.cvrp.utils import gen_instance
    from oracle import aco_torch as O
    torch.manual_seed(3)
    demand, dist = gen_instance(40, DEV)
    N, A = 41, 48
    heu = torch.rand(N, N, device=DEV) * 0.9 + 0.05
    aco = ACO(dist, demand, n_ants=A, heuristic=heu, device=DEV)

    torch.manual_seed(9)                                                           # cvrp/aco.py:138-165, step by step
    actions = torch.zeros(A, dtype=torch.long, device=DEV)
    visit = aco.update_visit_mask(torch.ones(A, N, device=DEV), actions)
    used = torch.zeros(A, device=DEV)
    used, cap = aco.update_capacity_mask(actions, used)
    rows, logps = [actions], []
    while not aco.check_done(visit, actions):
        actions, lp = aco.pick_move(actions, visit, cap, True)
        rows.append(actions)
        logps.append(lp)
        visit = aco.update_visit_mask(visit.clone(), actions)
        used, cap = aco.update_capacity_mask(actions, used)
    paths, logp = torch.stack(rows), torch.stack(logps)

    torch.manual_seed(9)
    want_paths, want_logp = O.cvrp_gen_path(torch.ones_like(dist), heu, demand, 50, A, require_prob=True)
    assert torch.equal(paths, want_paths)
    assert torch.allclose(logp, want_logp, rtol=0, atol=1e-6)
    torch.manual_seed(9)
    fused_paths, fused_logp = aco.gen_path(require_prob=True)
    assert torch.equal(paths, fused_paths)
    assert torch.allclose(logp, fused_logp, rtol=0, atol=1e-6)


def test_pick_move_log_probs_carry_the_gradient_to_the_heuristic():
    from torch.distributions import Categorical

    from deepaco_b200.tsp.aco import ACO
    n, A = 30, 16
    dist, heu0 = _tsp_instance(n, 8, 1)
    prev = torch.randint(0, n, (A,), device=DEV)
    mask = (torch.rand(A, n, device=DEV) > 0.3).float()
    mask[:, 0] = 1
    heu = heu0.clone().requires_grad_(True)
    aco = ACO(dist, n_ants=A, heuristic=heu, device=DEV)
    torch.manual_seed(2)
    actions, lp = aco.pick_move(prev, mask, True)
    lp.sum().backward()
    ref_heu = heu0.clone().requires_grad_(True)
    torch.manual_seed(2)
    d = Categorical(torch.ones_like(dist)[prev] * ref_heu[prev] * mask)            # tsp/aco.py:171-176
    ref_actions = d.sample()
    d.log_prob(ref_actions).sum().backward()
    assert torch.equal(actions, ref_actions)
    assert torch.allclose(heu.grad, ref_heu.grad, rtol=1e-6, atol=1e-9)
    _, none = aco.pick_move(prev, mask, False)
    assert none is None
    with pytest.raises(IndexError):
        aco.pick_move(torch.full((A,), n, device=DEV), mask, False)


def _categorical_step(ph, heu, prev, mask, mask2, alpha, beta, require_prob):
    """The statements every reference pick_move / pick_node consists of (op/aco.py:190-197, sop/aco.py:158-170)."""
    from torch.distributions import Categorical
    d = (ph[prev] ** alpha) * (heu[prev] ** beta) * mask
    if mask2 is not None:
        d = d * mask2
    dist = Categorical(d)
    item = dist.sample()
    return item, (dist.log_prob(item) if require_prob else None)


def _op_update_mask(mask, cur, travel, dist, max_len):
    """op/aco.py:199-219 (orienteering): visited nodes, nodes from which the depot cannot be reached any more, and the
    dummy node n that absorbs finished ants -- vectorised, same result as the reference's per-ant loop."""
    A, n1 = mask.shape
    n = n1 - 1
    mask = mask.clone()
    mask[torch.arange(A, device=mask.device), cur] = 0
    at_real = cur != n
    trails = travel[:, None] + dist[cur][:, :n] + dist[:n, 0][None, :]
    mask[:, :n] = torch.where(at_real[:, None] & (trails > max_len), torch.zeros_like(mask[:, :n]), mask[:, :n])
    mask[:, -1] = 0
    mask[(mask[:, :-1] == 0).all(dim=1), -1] = 1
    return mask


@pytest.mark.parametrize("style", ["op", "sop"])
def test_other_problem_directories_run_their_own_mask_rules_through_the_step(style):
    """The step of op/ (one mask, dummy node, travel budget) and sop/ (two masks: visited + precedence) through
    deepaco_b200.masked.pick_move: same actions as Categorical on this GPU under the same seed, step after step, with
    the masks updated by the caller's own rule; log-probs to fp32 rounding; same generator consumption."""
    from deepaco_b200 import masked
    torch.manual_seed(11)
    n, A = 40, 24
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    g = torch.cuda.default_generators[0]
    if style == "op":
        N = n + 1                                                   # + dummy node (op/aco.py:35-44)
        dist = torch.zeros(N, N, device=DEV)
        dist[:n, :n] = d
        ph = torch.rand(N, N, device=DEV) + 0.5
        heu = torch.rand(N, N, device=DEV) + 0.05
        max_len = 3.0

        def run(step):
            cur = torch.zeros(A, dtype=torch.long, device=DEV)
            mask = _op_update_mask(torch.ones(A, N, device=DEV), cur, torch.zeros(A, device=DEV), dist, max_len)
            travel = torch.zeros(A, device=DEV)
            sol, lps = [cur], []
            for _ in range(N):
                if bool((mask[:, :-1] == 0).all()):
                    break
                nxt, lp = step(ph, heu, cur, mask, None)
                travel = travel + dist[cur, nxt]
                sol.append(nxt)
                lps.append(lp)
                cur = nxt
                mask = _op_update_mask(mask, cur, travel, dist, max_len)
            return torch.stack(sol), torch.stack(lps)
    else:
        N = n
        ph = torch.rand(N, N, device=DEV) + 0.5
        heu = 1.0 / (d + 0.1)
        prec = torch.triu(torch.rand(N, N, device=DEV) < 0.05, diagonal=1)      # prec[i, j]: i must precede j

        def run(step):
            cur = torch.zeros(A, dtype=torch.long, device=DEV)
            visited = torch.zeros(A, N, dtype=torch.bool, device=DEV)
            visited[:, 0] = True
            sol, lps = [cur], []
            for _ in range(N - 1):
                mask1 = (~visited).float()
                pending = (prec[None, :, :] & ~visited[:, :, None]).any(dim=1)   # j still has an unvisited predecessor
                mask2 = (~pending).float()
                mask2[(mask1 * mask2).sum(1) == 0] = 1.0                          # never an all-zero row
                nxt, lp = step(ph, heu, cur, mask1, mask2)
                visited[torch.arange(A, device=DEV), nxt] = True
                sol.append(nxt)
                lps.append(lp)
                cur = nxt
            return torch.stack(sol), torch.stack(lps)

    torch.manual_seed(77)
    ref_sol, ref_lp = run(lambda *a: _categorical_step(*a, 1, 1, True))
    ref_off = g.get_offset()
    torch.manual_seed(77)
    sol, lp = run(lambda p, h, prev, m1, m2: masked.pick_move(p, h, prev, m1, m2, require_prob=True))
    assert g.get_offset() == ref_off
    assert torch.equal(sol, ref_sol)
    assert torch.allclose(lp, ref_lp, rtol=0, atol=2e-6)
    assert sol.shape[0] > 3
