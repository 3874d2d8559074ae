"""sample() backward: analytic gradient kernel vs autograd through the op-for-op oracle on the same
GPU, same seed (hence same paths).  fp32 tolerance: both sides accumulate with atomics."""
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
