"""Training-mode heuristic network without a GPU.

1. The torch restatement (oracle/net_torch.py) is pinned to the UNMODIFIED reference: train-mode output, the
   gradient of every parameter and the BatchNorm running statistics equal the goldens written by
   tests/golden/make_golden_gnn_train.py from /root/reference/{tsp,cvrp}/net.py.
2. The CUDA kernel SOURCE (deepaco_b200/csrc/gnn_train.cuh) is compiled for the host by tests/cpu_emu (one OS thread
   per CUDA thread, pthread barriers) and must reproduce the same goldens for every cluster size -- this checks the
   kernels' indexing, phase ordering and reduction trees here; the sm_100a build of the same source is checked against
   the same goldens by tests/test_gpu_gnn_train.py on the B200.
The emulation is a checker only: the product (deepaco_b200.net) has no CPU path.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cpu_emu")


@pytest.fixture(scope="session")
def emu():
    out = os.path.join(EMU_DIR, "_build", "libgnn_train_emu.so")
    subprocess.run(["make", "-C", EMU_DIR, "-s"], check=True)
    h = ctypes.CDLL(out)
    from deepaco_b200._lib import GnnTrainArgs
    for fn in (h.emu_gnn_train_forward, h.emu_gnn_train_backward, h.emu_gnn_forward_group):
        fn.restype = ctypes.c_char_p
        fn.argtypes = [ctypes.POINTER(GnnTrainArgs), ctypes.c_int]
    return h


def _net(kind):
    from deepaco_b200.net import load_npz_state_dict
    if kind == "tsp":
        from deepaco_b200.tsp.net import Net
        weights = "weights_tsp100"
    else:
        from deepaco_b200.cvrp.net import Net
        weights = "weights_cvrp100"
    net = Net()
    r = net.load_state_dict(load_npz_state_dict(os.path.join(ROOT, "deepaco_b200", "data", weights + ".npz")))
    assert not r.missing_keys and not r.unexpected_keys
    return net.train()


def _pyg(g):
    from deepaco_b200.net import Data
    return Data(x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"]), edge_attr=torch.from_numpy(g["edge_attr"]))


CASES = [("tsp", "tsp_n40_gnn_train_grads"), ("cvrp", "cvrp_n14_gnn_train_grads")]


def _check_against_golden(net, g, heu, grads, bufs, tag):
    """heu / parameter gradients / running statistics vs the reference's.  Tolerances: fp32 with a different
    summation order; gradients additionally get an absolute floor relative to the largest gradient of the network
    (the biases in front of a train-mode BatchNorm have an exactly-zero true gradient: rounding noise on both sides)."""
    assert torch.allclose(heu, torch.from_numpy(g["heu_vec"]), rtol=2e-4, atol=1e-7), tag
    gmax = max(float(np.abs(v).max()) for k, v in g.items() if k.startswith("grad__"))
    checked = 0
    for name, _ in net.named_parameters():
        key = "grad__" + name.replace(".", "__")
        if key not in g:
            assert grads.get(name) is None, name                   # None in the reference -> None here (not zeros)
            continue
        assert torch.allclose(grads[name], torch.from_numpy(g[key]), rtol=2e-3, atol=2e-5 * gmax), (tag, name)
        checked += 1
    assert checked >= 70
    for name, want in ((k[5:].replace("__", "."), v) for k, v in g.items() if k.startswith("buf__")):
        got = bufs[name]
        if name.endswith("num_batches_tracked"):
            assert int(got) == int(want), name
        else:
            assert torch.allclose(got, torch.from_numpy(want), rtol=1e-4, atol=1e-6), (tag, name)


@pytest.mark.parametrize("kind,fixture", CASES)
def test_torch_restatement_is_pinned_to_the_reference(golden, kind, fixture):
    from oracle import net_torch
    g = golden(fixture)
    net = _net(kind)
    heu = net_torch.net_forward(net, _pyg(g))
    (heu * torch.from_numpy(g["c"])).sum().backward()
    grads = {k: p.grad for k, p in net.named_parameters()}
    _check_against_golden(net, g, heu.detach(), grads, dict(net.named_buffers()), "oracle")


def _emu_run(emu, net, pyg, c, ctas, threads_f, threads_b, n_copies=1):
    """Forward + backward of L = sum(c * heu) through the host build of the kernels; returns heu, {param: grad}, stats."""
    from deepaco_b200 import net as N
    x = pyg.x.to(torch.float32)[None].repeat(n_copies, 1, 1).contiguous()
    ei = pyg.edge_index[None].repeat(n_copies, 1, 1)
    ea = pyg.edge_attr[None].repeat(n_copies, 1, 1)
    n = x.shape[1]
    graph = N.train_graph(ei, ea, n)
    flat = N.pack_weights(net, for_training=True)
    w = flat.detach().contiguous()
    bufs = N.train_buffers(n_copies, n, graph["E"], "cpu")
    for t in bufs.values():
        if t.is_floating_point():
            t.fill_(float("nan"))                      # anything read before it is written poisons the result
    heu = torch.full((n_copies, graph["E"]), float("nan"))
    eps = float(net.emb_net.v_bns[0].module.eps)
    a, keep = N.train_args(x, graph, w, bufs, net.emb_net.feats, ctas, eps, heu_out=heu)
    err = emu.emu_gnn_train_forward(ctypes.byref(a), threads_f)
    assert err is None, err
    g_heu = c[None].repeat(n_copies, 1).contiguous()
    grad = torch.zeros((n_copies, ctas, w.numel()))
    a, keep = N.train_args(x, graph, w, bufs, net.emb_net.feats, ctas, eps, grad_heu=g_heu, grad_weights=grad)
    err = emu.emu_gnn_train_backward(ctypes.byref(a), threads_b)
    assert err is None, err
    flat.backward(grad.sum(dim=(0, 1)))            # torch.cat routes the packed gradient to the parameters
    return heu, {k: p.grad for k, p in net.named_parameters()}, bufs["stats"], graph


@pytest.mark.parametrize("kind,fixture", CASES)
@pytest.mark.parametrize("ctas,threads_f,threads_b", [(1, 256, 256), (2, 128, 128), (4, 64, 192), (8, 64, 128), (16, 64, 128)])
def test_kernel_source_on_host_matches_the_reference(emu, golden, kind, fixture, ctas, threads_f, threads_b):
    from deepaco_b200 import net as N
    g = golden(fixture)
    net = _net(kind)
    pyg = _pyg(g)
    heu, grads, stats, graph = _emu_run(emu, net, pyg, torch.from_numpy(g["c"]), ctas, threads_f, threads_b)
    assert torch.isfinite(heu).all() and torch.isfinite(stats).all()
    N._update_running_stats(N.flat_state(net), stats, graph["n"], graph["E"])
    _check_against_golden(net, g, heu[0], grads, dict(net.named_buffers()), f"emu ctas={ctas}")


def test_cluster_size_does_not_change_the_forward_bits_much_and_batches_are_independent(emu, golden):
    """Two copies of the same graph in one launch give identical results (instances are independent forward calls), and
    the gradient of the batch is twice the single-instance gradient."""
    g = golden("tsp_n40_gnn_train_grads")
    c = torch.from_numpy(g["c"])
    net1, net2 = _net("tsp"), _net("tsp")
    heu1, grads1, stats1, _ = _emu_run(emu, net1, _pyg(g), c, 2, 128, 128)
    heu2, grads2, stats2, _ = _emu_run(emu, net2, _pyg(g), c, 2, 128, 128, n_copies=2)
    assert torch.equal(heu2[0], heu2[1]) and torch.equal(heu2[0], heu1[0])
    assert torch.equal(stats2[0], stats2[1])
    for k in grads1:
        if grads1[k] is not None:
            assert torch.allclose(grads2[k], 2 * grads1[k], rtol=1e-6, atol=0), k


def test_argument_checks_match_the_library(emu):
    from deepaco_b200._lib import GnnTrainArgs
    a = GnnTrainArgs()
    assert emu.emu_gnn_train_forward(ctypes.byref(a), 128) == b"NULL argument"


def test_python_wiring_autograd_function_and_running_stats(emu, golden, monkeypatch):
    """deepaco_b200.net's training path (autograd Function, packed-gradient routing, running-stat update, batch API)
    exercised end to end with the two C-ABI calls redirected to the host build of the same kernels."""
    import contextlib

    from deepaco_b200 import _lib
    from deepaco_b200 import net as N
    from oracle import net_torch

    class HostLib:
        @staticmethod
        def deepaco_gnn_train_forward(a, stream):
            return 0 if emu.emu_gnn_train_forward(a, 128) is None else -1

        @staticmethod
        def deepaco_gnn_train_backward(a, stream):
            return 0 if emu.emu_gnn_train_backward(a, 128) is None else -1

    monkeypatch.setattr(N, "lib", lambda: HostLib)
    monkeypatch.setattr(N, "stream_ptr", lambda dev=None: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: t)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setenv("DEEPACO_GNN_CTAS", "2")
    g = golden("tsp_n40_gnn_train_grads")
    net = _net("tsp")
    heu = net(_pyg(g))
    assert heu.requires_grad
    (heu * torch.from_numpy(g["c"])).sum().backward()
    _check_against_golden(net, g, heu.detach(), {k: p.grad for k, p in net.named_parameters()}, dict(net.named_buffers()), "wiring")
    # frozen backbone (tsp/net.py:90-92) and a batch of two graphs == two sequential forward calls of the torch restatement
    net, ref = _net("tsp"), _net("tsp")
    net.freeze_gnn(); ref.freeze_gnn()
    pyg = _pyg(g)
    hb = N.gnn_train_forward(net, torch.stack([pyg.x, pyg.x * 0.5]), torch.stack([pyg.edge_index] * 2), torch.stack([pyg.edge_attr, pyg.edge_attr * 1.5]))
    hb.sum().backward()
    from deepaco_b200.net import Data
    wants = [net_torch.net_forward(ref, pyg), net_torch.net_forward(ref, Data(x=pyg.x * 0.5, edge_index=pyg.edge_index, edge_attr=pyg.edge_attr * 1.5))]
    (wants[0].sum() + wants[1].sum()).backward()
    for b in range(2):
        assert torch.allclose(hb[b].detach(), wants[b].detach(), rtol=2e-4, atol=1e-7)
    assert all(p.grad is None for p in net.emb_net.parameters())
    for (k, p), (_, q) in zip(net.par_net_heu.named_parameters(), ref.par_net_heu.named_parameters()):
        if q.grad is not None:
            assert torch.allclose(p.grad, q.grad, rtol=2e-3, atol=1e-5 * float(q.grad.abs().max())), k
    for (k, v), (_, w) in zip(net.named_buffers(), ref.named_buffers()):
        assert torch.allclose(v.float(), w.float(), rtol=1e-4, atol=1e-6), k


def test_barrier_placement_under_thread_sanitizer():
    """`make tsan`: both kernels on an irregular synthetic graph with 1, 2, 4 (cluster barrier) and 16 (arrival-counter barrier) CTAs per graph under ThreadSanitizer;
    any cross-thread dependency not ordered by __syncthreads / the cluster barrier fails the run."""
    import shutil
    if not (os.path.exists("/usr/bin/g++") and shutil.which("make")):
        pytest.skip("no distribution g++ with libtsan")
    probe = subprocess.run(["/usr/bin/g++", "-fsanitize=thread", "-x", "c++", "-", "-o", "/dev/null"], input=b"int main(){}",
                           capture_output=True)
    if probe.returncode != 0:
        pytest.skip("libtsan not installed")
    r = subprocess.run(["make", "-C", EMU_DIR, "tsan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("tsan run ok") == 4


@pytest.mark.parametrize("kind,fixture", CASES)
@pytest.mark.parametrize("ctas,threads", [(1, 256), (2, 128), (8, 64), (16, 64)])
def test_eval_mode_group_forward_on_host_matches_the_restatement(emu, golden, kind, fixture, ctas, threads):
    """deepaco_gnn_forward_group (eval mode: running-statistics BatchNorm, two-phase layers, ping-pong state) through
    the host build of the kernel source, against the torch restatement in eval mode with the pretrained checkpoint
    (itself pinned to the reference's eval output by tests/test_oracle_golden.py)."""
    from deepaco_b200 import net as N
    from oracle import net_torch
    g = golden(fixture)
    net = _net(kind).eval()
    pyg = _pyg(g)
    with torch.no_grad():
        want = net_torch.net_forward(net, pyg)
    n_copies = 2
    x = pyg.x.to(torch.float32)[None].repeat(n_copies, 1, 1).contiguous()
    ei = pyg.edge_index[None].repeat(n_copies, 1, 1)
    ea = pyg.edge_attr[None].repeat(n_copies, 1, 1)
    n = x.shape[1]
    graph = N.train_graph(ei, ea, n, backward=False)
    bufs = N.eval_buffers(n_copies, n, graph["E"], "cpu")
    for t in bufs.values():
        if t.is_floating_point():
            t.fill_(float("nan"))                      # anything read before it is written poisons the result
    heu = torch.full((n_copies, graph["E"]), float("nan"))
    a, keep = N.train_args(x, graph, N.pack_weights(net), bufs, net.emb_net.feats, ctas, 1e-5, heu_out=heu)
    err = emu.emu_gnn_forward_group(ctypes.byref(a), threads)
    assert err is None, err
    assert torch.isfinite(heu).all()
    assert torch.equal(heu[0], heu[1])
    assert torch.allclose(heu[0], want, rtol=2e-4, atol=1e-7), float((heu[0] - want).abs().max())


def test_eval_weights_follow_running_statistics_updated_in_training_mode():
    """Net._weights() (the eval-mode packed weights, cached) must notice BatchNorm running statistics that a
    training-mode forward updated through the stacked buffers they alias, even when no parameter changed."""
    from deepaco_b200 import net as N
    net = _net("tsp")
    w0 = net.eval()._weights().clone()
    st = N.flat_state(net)
    assert st.track
    stats = torch.rand(1, N.DEPTH, 6, N.UNITS) + 0.5
    N._update_running_stats(st, stats, 40, 400)
    w1 = net._weights()
    assert not torch.equal(w0, w1)
    assert torch.equal(w1, N.pack_weights(net))
    assert net._weights() is w1                                   # unchanged state -> cached tensor


@pytest.fixture(scope="session")
def emu_g():
    out = os.path.join(EMU_DIR, "_build", "libgnn_emu.so")
    subprocess.run(["make", "-C", EMU_DIR, "-s", "_build/libgnn_emu.so"], check=True)
    h = ctypes.CDLL(out)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    h.emu_gnn_forward.restype = ctypes.c_char_p
    h.emu_gnn_forward.argtypes = [vp] * 6 + [ci] * 4 + [vp] * 4 + [ctypes.c_float, ci]
    return h


@pytest.mark.parametrize("kind,fixture", CASES)
def test_one_cta_eval_kernel_on_host_matches_the_restatement_and_the_group_kernel(emu, emu_g, golden, kind, fixture):
    """deepaco_gnn_forward (csrc/gnn.cuh: one CTA per instance, 16-row tiles through tile_linear -- the tensor-core MMA on
    the device, its plain-loop stand-in here --, fused reshape + EPS) through its host build: equal to the eval-mode torch
    restatement and to deepaco_gnn_forward_group's host build to fp32 rounding (different summation order inside the
    32-term dot products), and the dense output is Net.reshape(...) + EPS."""
    from deepaco_b200 import net as N
    from oracle import net_torch
    g = golden(fixture)
    net = _net(kind).eval()
    pyg = _pyg(g)
    with torch.no_grad():
        want = net_torch.net_forward(net, pyg)
    n = pyg.x.shape[0]
    graph = N.train_graph(pyg.edge_index[None], pyg.edge_attr[None], n, backward=False)
    E = graph["E"]
    w = N.pack_weights(net)
    x = pyg.x.to(torch.float32)[None].contiguous()
    node_ws = torch.full((1, n, 6 * N.UNITS), float("nan"))
    edge_ws = torch.full((1, E, N.UNITS), float("nan"))
    heu = torch.full((1, E), float("nan"))
    dense = torch.full((1, n, n), float("nan"))
    err = emu_g.emu_gnn_forward(x.data_ptr(), graph["row_ptr"].data_ptr(), graph["dst"].data_ptr(), graph["attr"].data_ptr(),
                                graph["order"].data_ptr(), w.data_ptr(), n, E, net.emb_net.feats, 1, node_ws.data_ptr(),
                                edge_ws.data_ptr(), heu.data_ptr(), dense.data_ptr(), 1e-10, 256)
    assert err is None, err
    assert torch.allclose(heu[0], want, rtol=2e-4, atol=1e-7)
    assert torch.equal(dense[0], N.Net.reshape(pyg, heu[0]) + 1e-10)
    # group kernel, 4 CTAs per graph
    bufs = N.eval_buffers(1, n, E, "cpu")
    heu_g = torch.full((1, E), float("nan"))
    a, keep = N.train_args(x, graph, w, bufs, net.emb_net.feats, 4, 1e-5, heu_out=heu_g)
    assert emu.emu_gnn_forward_group(ctypes.byref(a), 128) is None
    assert torch.allclose(heu_g, heu, rtol=1e-4, atol=1e-7)
