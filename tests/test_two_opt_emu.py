"""The 2-opt / NLS kernel SOURCE (deepaco_b200/csrc/two_opt.cuh) without a GPU: tests/cpu_emu compiles the same text for
the host (one OS thread per CUDA thread, pthread barriers, warp shuffles through an exchange buffer, TMA / cp.async /
mbarrier stand-ins) and it must reproduce the C oracle -- itself pinned to the reference's numba output -- for the
register-carry variants (KMAX = 4 / 8 / 16), the band kernel, both row-staging paths, NLS and tours that are not
permutations.  The sm_100a build of the same source is checked against the same oracle by tests/test_gpu_two_opt.py.
The emulation is a checker only: the product has no CPU path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import two_opt as T2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "cpu_emu")


@pytest.fixture(scope="session")
def emu2():
    subprocess.run(["make", "-C", EMU_DIR, "-s", "_build/libtwo_opt_emu.so"], check=True)
    h = ctypes.CDLL(os.path.join(EMU_DIR, "_build", "libtwo_opt_emu.so"))
    h.emu_two_opt.restype = ctypes.c_char_p
    vp = ctypes.c_void_p
    h.emu_two_opt.argtypes = [vp, vp, vp] + [ctypes.c_int] * 7 + [vp, vp]
    return h


def _instance(n, count, seed, asymmetric=False):
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2), dtype=np.float32)
    dist = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1)).astype(np.float32)
    np.fill_diagonal(dist, 1e9)
    if asymmetric:
        dist = (dist * (0.5 + rng.random((n, n), dtype=np.float32))).astype(np.float32)
    tours = np.stack([np.concatenate(([0], 1 + rng.permutation(n - 1))) for _ in range(count)]).astype(np.uint16)
    return np.ascontiguousarray(dist), np.ascontiguousarray(tours), rng


def _run(emu2, dist, tours, it, variant=-1, heu_dist=None, T_nls=0, T_p=0):
    out = np.ascontiguousarray(tours.copy())
    passes = np.zeros(len(tours), np.int32)
    err = emu2.emu_two_opt(dist.ctypes.data, None if heu_dist is None else heu_dist.ctypes.data, out.ctypes.data, dist.shape[0],
                           len(tours), 0 if heu_dist is None else 1, it, T_nls, T_p, variant, None, passes.ctypes.data)
    assert err is None, err
    return out, passes


@pytest.mark.parametrize("n,count,it,variant", [
    (4, 2, 10, -1), (5, 3, 50, -1), (33, 3, 40, -1), (100, 2, 12, -1), (127, 1, 6, -1),      # KMAX 4; n = 100: TMA rows
    (128, 1, 6, -1), (200, 1, 6, -1), (255, 1, 4, -1),                                     # KMAX 8
    (256, 1, 3, -1), (261, 1, 3, -1),                                                      # KMAX 16 (TMA / cp.async rows)
    (60, 2, 20, 8), (60, 2, 20, 16),                                                       # a larger KMAX than needed
    (60, 2, 20, 0), (131, 1, 5, 0),                                                        # band kernel
])
def test_two_opt_kernel_source_on_host_matches_the_oracle(emu2, n, count, it, variant):
    dist, tours, _ = _instance(n, count, seed=n + max(variant, 0), asymmetric=bool(n % 2))
    got, passes = _run(emu2, dist, tours, it, variant)
    ref = T2.batched_two_opt(dist, tours, it)
    assert np.array_equal(got, ref)
    assert (passes >= 1).all() and (passes <= it).all()


def test_nls_kernel_source_on_host_matches_the_oracle(emu2):
    n, count = 48, 3
    dist, tours, rng = _instance(n, count, seed=11)
    heu = (rng.random((n, n), dtype=np.float32) * 0.9 + 0.05).astype(np.float32)
    heu_dist = np.ascontiguousarray((1 / (heu / heu.max(-1, keepdims=True) + np.float32(1e-5))).astype(np.float32))
    got, _ = _run(emu2, dist, tours, n // 4, heu_dist=heu_dist, T_nls=3, T_p=5)
    ref = T2.nls(dist, heu_dist, tours, n // 4, 3, 5)
    assert np.array_equal(got, ref)


def test_tours_with_repeated_nodes_take_the_band_kernel_inside_the_launch(emu2):
    n, count = 40, 4
    dist, tours, rng = _instance(n, count, seed=7)
    for a in range(0, count, 2):
        pos = rng.choice(np.arange(2, n), size=3, replace=False)
        tours[a, pos] = tours[a, pos - 2]
    got, _ = _run(emu2, dist, tours, 15)
    assert np.array_equal(got, T2.batched_two_opt(dist, tours, 15))


def test_argument_checks(emu2):
    dist, tours, _ = _instance(8, 1, seed=1)
    assert emu2.emu_two_opt(dist.ctypes.data, None, tours.ctypes.data, 8, 1, 1, 3, 1, 1, -1, None, None) == b"heuristic_dist is NULL"
    assert emu2.emu_two_opt(dist.ctypes.data, None, tours.ctypes.data, 200, 1, 0, 3, 0, 0, 4, None, None) == b"n does not fit this KMAX"


def test_barrier_placement_under_thread_sanitizer():
    """`make tsan2opt`: register-carry kernel (TMA and cp.async rows), band kernel and NLS under ThreadSanitizer; any
    cross-thread shared-memory dependency not ordered by a barrier / warp sync point / mbarrier wait fails the run.
    (The stand-in copies complete at issue time, so this checks barrier placement, not the async-copy wait counts.)"""
    import shutil
    if not (os.path.exists("/usr/bin/g++") and shutil.which("make")):
        pytest.skip("no distribution g++ with libtsan")
    probe = subprocess.run(["/usr/bin/g++", "-fsanitize=thread", "-x", "c++", "-", "-o", "/dev/null"], input=b"int main(){}",
                           capture_output=True)
    if probe.returncode != 0:
        pytest.skip("libtsan not installed")
    r = subprocess.run(["make", "-C", EMU_DIR, "tsan2opt"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("tsan run ok") == 5
