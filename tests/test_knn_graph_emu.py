"""deepaco_b200/csrc/knn_graph.cuh (distance matrix + k-nearest-neighbour edges, the instance -> graph front end) compiled
for the host (tests/cpu_emu) against the reference's op chain (tsp/utils.py:4-36, cvrp/utils.py:18-22).  The distance bits
are checked against the formula the B200 probe pinned for ATen's CUDA norm (tools/probes/norm_formula.py):
sqrt(fl(fl(dx * dx) + fl(dy * dy))) -- restated here in numpy float32 -- and the selection against torch.topk on the same
matrix (random coordinates: no bit-equal distances inside a row, asserted)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cpu_emu")
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", EMU_DIR, "-s", "_build/libtsp_update_emu.so"], check=True)
    h = ctypes.CDLL(os.path.join(EMU_DIR, "_build", "libtsp_update_emu.so"))
    h.emu_knn_graph.restype = ctypes.c_char_p
    h.emu_knn_graph.argtypes = [vp, vp, ci, ci, ci, cf, vp, vp, vp, vp]
    return h


def _ptr(t):
    return None if t is None else t.data_ptr()


def _ref_distances(coords, diag):
    c = coords.numpy()
    d = c[:, :, None, :] - c[:, None, :, :]                                  # float32, each op rounded on its own
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)
    i = np.arange(c.shape[1])
    dist[:, i, i] = np.float32(diag)
    return torch.from_numpy(dist)


@pytest.mark.parametrize("B,n,k,diag", [(1, 5, 4, 1e9), (3, 20, 10, 1e9), (2, 100, 20, 1e9), (1, 200, 20, 1e9), (1, 500, 50, 1e9),
                                         (2, 33, 33, 1e9), (2, 21, 0, 1e-10)])
def test_knn_graph_kernel_on_host(emu, B, n, k, diag):
    g = torch.Generator().manual_seed(100 * n + k)
    coords = torch.rand((B, n, 2), generator=g)
    dist = torch.full((B, n, n), float("nan"))
    idx = torch.full((B, n, max(k, 1)), -1, dtype=torch.int32)
    val = torch.full((B, n, max(k, 1)), float("nan"))
    ei = torch.full((B, 2, n * max(k, 1)), -1, dtype=torch.int64)
    err = emu.emu_knn_graph(_ptr(coords), None, n, B, k, diag, _ptr(dist), _ptr(idx) if k else None, _ptr(val) if k else None,
                            _ptr(ei) if k else None)
    assert err is None, err
    want = _ref_distances(coords, diag)
    assert torch.equal(dist, want)
    if k == 0:
        return
    tv, ti = torch.topk(want, k=k, dim=2, largest=False)
    assert not (tv[:, :, 1:] == tv[:, :, :-1]).any(), "instance has bit-equal distances inside a row: tie order is unspecified"
    assert torch.equal(val, tv) and torch.equal(idx.long(), ti)
    src = torch.arange(n).repeat_interleave(k)
    for b in range(B):                                                       # tsp/utils.py:29-34
        assert torch.equal(ei[b], torch.stack([src, ti[b].flatten()]))
    # same selection from a given matrix (the batched front end passes distances)
    idx2 = torch.empty_like(idx)
    val2 = torch.empty_like(val)
    assert emu.emu_knn_graph(None, _ptr(want), n, B, k, 0.0, None, _ptr(idx2), _ptr(val2), None) is None
    assert torch.equal(idx2, idx) and torch.equal(val2, val)


def test_equal_distances_go_lowest_column_first(emu):
    """Grid coordinates: many bit-equal distances per row.  Values equal torch.topk's; among equal values the kernel lists
    the lower column first (torch's order there is unspecified)."""
    c = torch.stack(torch.meshgrid(torch.arange(6.), torch.arange(6.), indexing="ij"), dim=-1).reshape(1, -1, 2) / 8
    n, k = 36, 12
    dist = torch.empty((1, n, n))
    idx = torch.empty((1, n, k), dtype=torch.int32)
    val = torch.empty((1, n, k))
    assert emu.emu_knn_graph(_ptr(c), None, n, 1, k, 1e9, _ptr(dist), _ptr(idx), _ptr(val), None) is None
    tv, _ = torch.topk(dist, k=k, dim=2, largest=False)
    assert torch.equal(val, tv)
    assert torch.equal(torch.gather(dist, 2, idx.long()), val)
    same = val[:, :, 1:] == val[:, :, :-1]
    assert same.any() and bool((idx[:, :, 1:][same] > idx[:, :, :-1][same]).all())
    assert all(len(set(r.tolist())) == k for r in idx[0])


def test_selection_reproduces_the_unmodified_references_graph(emu):
    """tests/golden/tsp_n100_a32_gnn.npz holds gen_pyg_data(coords, 20) of the UNMODIFIED reference (tsp/utils.py:16-36, run
    on the host by tests/golden/make_golden.py).  From the reference's own distance matrix the kernel's selection is the
    reference's edge_index / edge_attr exactly.  (The distance bits themselves are device arithmetic: the host build of
    ATen's norm rounds 9 % of the entries differently from its CUDA build, so they are pinned on the GPU,
    tests/test_gpu_graph.py, against the same ops on the same device.)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "tsp_n100_a32_gnn.npz"))
    dist = torch.from_numpy(z["dist"]).contiguous()[None]
    n, k = dist.shape[1], z["edge_index"].shape[1] // dist.shape[1]
    idx = torch.empty((1, n, k), dtype=torch.int32)
    val = torch.empty((1, n, k))
    ei = torch.empty((1, 2, n * k), dtype=torch.int64)
    assert emu.emu_knn_graph(None, _ptr(dist), n, 1, k, 0.0, None, _ptr(idx), _ptr(val), _ptr(ei)) is None
    assert np.array_equal(ei[0].numpy(), z["edge_index"])
    assert np.array_equal(val.reshape(-1, 1).numpy(), z["edge_attr"])
    # and the distances from the coordinates agree with the reference's host result to fp32 rounding
    d2 = torch.empty((1, n, n))
    c = torch.from_numpy(z["coords"]).contiguous()[None]
    assert emu.emu_knn_graph(_ptr(c), None, n, 1, 0, 1e9, _ptr(d2), None, None, None) is None
    assert np.allclose(d2[0].numpy(), z["dist"], rtol=2e-7, atol=0)
