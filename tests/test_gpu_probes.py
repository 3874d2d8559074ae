"""On-device confirmation of the torch-parity assumptions:
our regenerated Philox streams equal torch's `exponential_` / `randint`, and our row sums equal
`torch.sum(x, -1)` bit for bit, for every shape class the kernels rely on."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gen():
    g = torch.cuda.default_generators[0]
    return g, int(g.initial_seed()), int(g.get_offset())


@pytest.mark.parametrize("shape", [(8, 20), (512, 100), (512, 101), (256, 200), (256, 500), (2048, 200), (1024, 1000)])
def test_exponential_stream(shape):
    from deepaco_b200 import _engine as E
    torch.manual_seed(1234)
    torch.empty(3, device="cuda").exponential_(1)           # move the offset off zero
    g, seed, off = _gen()
    ref = torch.empty(shape, device="cuda").exponential_(1)
    mine = E.debug_exponential(seed, off, ref.numel(), "cuda").reshape(shape)
    assert torch.equal(ref, mine)
    import ctypes as C
    from deepaco_b200._lib import lib
    thr, inc = C.c_uint32(), C.c_uint64()
    lib().deepaco_torch_draw_geometry(ref.numel(), C.byref(thr), C.byref(inc))
    assert g.get_offset() - off == inc.value


@pytest.mark.parametrize("numel,high", [(8, 20), (512, 100), (256, 200), (20, 5), (400000, 1000)])
def test_randint_stream(numel, high):
    from deepaco_b200 import _engine as E
    torch.manual_seed(99)
    g, seed, off = _gen()
    ref = torch.randint(low=0, high=high, size=(numel,), device="cuda")
    mine = E.debug_randint(seed, off, numel, high, "cuda")
    assert torch.equal(ref, mine)


@pytest.mark.parametrize("rows,cols", [(8, 20), (20, 20), (16, 50), (512, 100), (32, 100), (16, 100), (512, 101),
                                       (256, 200), (64, 128), (256, 500), (256, 1000), (100, 119), (512, 64),
                                       (50, 500), (20, 21), (16, 21)])
def test_row_sum_order(rows, cols):
    from deepaco_b200 import _engine as E
    bw, vec, exact = E.aten_sum_plan(cols, rows)
    torch.manual_seed(rows * 1000 + cols)
    for scale in (1.0, 1e-6):
        x = torch.rand(rows, cols, device="cuda") * scale
        x[x < 0.3 * scale] = 0.0                       # masked entries, like visited nodes
        ref = torch.sum(x, dim=-1)
        mine = E.debug_row_sum(x)
        if exact:
            assert torch.equal(ref, mine), f"bw={bw} vec={vec}: {(ref - mine).abs().max().item()}"
        else:
            assert torch.allclose(ref, mine, rtol=1e-6)


def test_shortened_exponential_transform_is_bit_identical():
    """common.cuh exp1_from_word drops the explicit u >= 1 - 2^-25 guard in favour of max(., 2^-24); the two forms are
    compared exhaustively over the top 2^20 Philox words (the only place they can differ) and a stride sample."""
    from deepaco_b200 import _engine as E
    assert E.debug_exp_guard("cuda") == 0
