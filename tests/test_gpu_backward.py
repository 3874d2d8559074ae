"""sample() backward: analytic gradient kernel vs autograd through the op-for-op oracle on the same
GPU, same seed (hence same paths).  fp32 tolerance against autograd (whose scatter is atomic); the kernel itself
accumulates in a fixed order and is bit-identical run to run."""
import pytest
import torch

from oracle import aco_torch as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tsp(n, seed):
    torch.manual_seed(seed)
    xy = torch.rand(n, 2, device=DEV)
    d = torch.norm(xy[:, None] - xy, dim=2, p=2)
    d[torch.arange(n), torch.arange(n)] = 1e9
    heu = torch.rand(n, n, device=DEV) * 0.9 + 0.05
    heu[torch.rand(n, n, device=DEV) < 0.3] = 1e-10          # off-graph entries as in tsp/train.ipynb cell 1
    return d, heu


@pytest.mark.parametrize("nls", [False, True])
def test_tsp_sample_backward_matches_autograd(nls):
    from deepaco_b200.tsp.aco import ACO as TspACO
    from deepaco_b200.tsp_nls.aco import ACO as NlsACO
    n, A = 30, 16
    d, heu0 = _tsp(n, 3)
    ph = torch.rand(n, n, device=DEV) + 0.5
    # reference: autograd through the oracle ops
    h_ref = heu0.clone().requires_grad_(True)
    torch.manual_seed(99)
    paths_ref, logp_ref = O.tsp_gen_path(ph, h_ref, A, require_prob=True, nls_variant=nls)
    costs = O.tsp_path_costs(d, paths_ref)
    loss_ref = ((costs - costs.mean()) * logp_ref.sum(0)).sum() / A          # REINFORCE loss of tsp/train.ipynb cell 1
    loss_ref.backward()
    # ours
    h = heu0.clone().requires_grad_(True)
    torch.manual_seed(99)
    aco = (NlsACO if nls else TspACO)(d, n_ants=A, heuristic=h, pheromone=ph, device=DEV)
    if nls:
        c, logp, paths = aco.sample()
    else:
        c, logp = aco.sample()
        paths = None
    loss = ((c - c.mean()) * logp.sum(0)).sum() / A
    loss.backward()
    assert torch.allclose(loss, loss_ref, rtol=1e-5)
    assert torch.allclose(h.grad, h_ref.grad, rtol=2e-4, atol=1e-6 * float(h_ref.grad.abs().max()))
    assert (h.grad != 0).any()


def test_cvrp_sample_backward_matches_autograd():
    from deepaco_b200.cvrp.aco import ACO
    n, A = 24, 12
    torch.manual_seed(5)
    loc = torch.rand(n, 2, device=DEV)
    demand = torch.cat((torch.zeros(1, device=DEV), torch.randint(1, 10, (n,), device=DEV).float()))
    allc = torch.cat((torch.tensor([[0.5, 0.5]], device=DEV), loc))
    d = torch.norm(allc[:, None] - allc, dim=2, p=2)
    d[torch.arange(n + 1), torch.arange(n + 1)] = 1e-10
    heu0 = torch.rand(n + 1, n + 1, device=DEV) * 0.9 + 0.05
    ph = torch.ones_like(d)
    h_ref = heu0.clone().requires_grad_(True)
    torch.manual_seed(31)
    paths_ref, logp_ref = O.cvrp_gen_path(ph, h_ref, demand, 50, A, require_prob=True)
    w = torch.linspace(-1, 1, A, device=DEV)
    (w * logp_ref.sum(0)).sum().backward()
    h = heu0.clone().requires_grad_(True)
    torch.manual_seed(31)
    aco = ACO(d, demand, n_ants=A, heuristic=h, device=DEV)
    costs, logp = aco.sample()
    (w * logp.sum(0)).sum().backward()
    assert torch.equal(logp.detach(), logp_ref.detach())
    assert torch.allclose(h.grad, h_ref.grad, rtol=2e-4, atol=1e-6 * float(h_ref.grad.abs().max()))


def test_train_instance_parameter_gradients_match_reference_pipeline():
    """One REINFORCE step as in tsp/train.ipynb cell 1 (`train_instance`): Net in train mode -> heuristic matrix ->
    ACO.sample() -> loss.backward().  The gradients of every network parameter equal those of the same pipeline with
    the reference's sampling ops (oracle) in place of the kernels."""
    import copy
    from deepaco_b200.tsp.aco import ACO
    from deepaco_b200.tsp.net import Net
    from deepaco_b200.tsp.utils import gen_pyg_data
    n, A, EPS = 40, 16, 1e-10
    torch.manual_seed(21)
    net = Net().to(DEV).train()
    ref_net = copy.deepcopy(net)
    coords = torch.rand(n, 2, device=DEV)
    pyg, distances = gen_pyg_data(coords, k_sparse=8)

    def loss_of(model, sampler, forward=lambda m: m(pyg)):
        heu_vec = forward(model)
        heu_mat = model.reshape(pyg, heu_vec) + EPS
        costs, log_probs = sampler(heu_mat)
        baseline = costs.mean()
        return torch.sum((costs - baseline) * log_probs.sum(dim=0)) / A

    def mine(heu_mat):
        torch.manual_seed(77)
        return ACO(n_ants=A, heuristic=heu_mat, distances=distances, device=DEV).sample()

    def reference(heu_mat):
        torch.manual_seed(77)
        paths, logp = O.tsp_gen_path(torch.ones_like(distances), heu_mat, A, require_prob=True)
        return O.tsp_path_costs(distances, paths), logp

    l1 = loss_of(net, mine)
    l1.backward()
    from oracle import net_torch
    l2 = loss_of(ref_net, reference, forward=lambda m: net_torch.net_forward(m, pyg))   # torch autograd through the restated ops
    l2.backward()
    assert torch.allclose(l1, l2, rtol=1e-4, atol=1e-5)     # native training-mode network vs torch ops: fp32 rounding only
    checked = 0
    # biases feeding a train-mode BatchNorm have an exactly-zero true gradient (pure rounding noise on both sides):
    # the absolute tolerance is therefore set from the largest gradient of the whole network
    gmax = max(float(p.grad.abs().max()) for p in ref_net.parameters() if p.grad is not None)
    for (name, p1), (_, p2) in zip(net.named_parameters(), ref_net.named_parameters()):
        if p2.grad is None:
            assert p1.grad is None, name
            continue
        if name.endswith(".bias") and any(t in name for t in ("v_lins1.", "v_lins3.", "v_lins4.", "e_lins0.")):
            # a bias in front of a train-mode BatchNorm: the true gradient is exactly zero, both sides hold rounding noise
            assert float(p1.grad.abs().max()) <= 1e-3 * gmax and float(p2.grad.abs().max()) <= 1e-3 * gmax, name
            continue
        assert torch.allclose(p1.grad, p2.grad, rtol=5e-3, atol=2e-5 * gmax), (name, float((p1.grad - p2.grad).abs().max()), gmax)
        checked += 1
    assert checked > 50
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4)
    opt.step()                                              # the optimiser step of train_instance runs


@pytest.mark.parametrize("problem", ["tsp", "cvrp"])
def test_logp_backward_is_bit_identical_run_to_run(problem):
    """deepaco_logp_backward adds every matrix element's contributions in a fixed order (ants ascending, steps
    ascending; no atomics): many ants, repeated calls -> identical bits, for the heuristic and the pheromone gradient."""
    from deepaco_b200 import _engine as E
    if problem == "tsp":
        n, A = 100, 512
        d, heu = _tsp(n, 11)
        ph = torch.rand(n, n, device=DEV) + 0.5
        torch.manual_seed(3)
        paths = O.tsp_gen_path(ph, heu, A)
        demand, cap = None, 0.0
    else:
        n, A = 61, 256
        torch.manual_seed(6)
        allc = torch.cat((torch.tensor([[0.5, 0.5]], device=DEV), torch.rand(n - 1, 2, device=DEV)))
        d = torch.norm(allc[:, None] - allc, dim=2, p=2)
        d[torch.arange(n), torch.arange(n)] = 1e-10
        demand = torch.cat((torch.zeros(1, device=DEV), torch.randint(1, 10, (n - 1,), device=DEV).float()))
        heu = torch.rand(n, n, device=DEV) * 0.98 + 1e-10
        ph = torch.rand(n, n, device=DEV) + 0.5
        paths = O.cvrp_gen_path(ph, heu, demand, 50, A)
        cap = 50.0
    g = torch.randn(paths.shape[0] - 1, A, device=DEV)
    runs = [E.logp_backward(ph, heu, paths, g, demand=demand, capacity=cap, want_pheromone_grad=True) for _ in range(4)]
    for gh, gp in runs[1:]:
        assert torch.equal(gh, runs[0][0]) and torch.equal(gp, runs[0][1])
    assert torch.isfinite(runs[0][0]).all() and (runs[0][0] != 0).any()
