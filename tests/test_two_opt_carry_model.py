"""The data flow of the register-carry 2-opt kernel (csrc/two_opt.cu: two_opt_call_v2), restated lane by lane in numpy
and checked against the C oracle (itself pinned to the reference's numba output).  What this pins on CPU: the identity
change(i, j) = G(i-1, j) + G(i, j+1) - edge[i] - edge[j+1] with G(a, m) = d[tour[a], tour[m]] in the reference's fp32
operation order, the position ownership m = 32k + lane with the wrap position m = n, the neighbour-lane hand-over of
g_i[m+1], the mirrored row pairing over warps (every row exactly once, two contiguous runs per warp) and the
(value, i*n+j) arg-min == first strict minimum rule.  The kernel itself is compared with the oracle on the GPU
(tests/test_gpu_two_opt.py)."""
import numpy as np
import pytest

from oracle import two_opt as T2

f32 = np.float32


def _pass(D, tour, n, W, KMAX):
    text = np.concatenate([tour, tour[:1]]).astype(np.int64)      # tour[n] mirrors tour[0]
    edge = np.zeros(n + 1, f32)
    for k in range(1, n + 1):
        edge[k] = D[text[k - 1], text[k]]
    lane = np.arange(32)
    cand = []
    rows_seen = []
    F = (n - 2) // 2
    for warp in range(W):
        lo1, hi1 = 1 + F * warp // W, 1 + F * (warp + 1) // W
        lo2, hi2 = n - hi1, n - lo1
        if warp == W - 1 and ((n - 2) & 1):
            hi1 = F + 2
        node = np.zeros((KMAX, 32), np.int64)
        en = np.zeros((KMAX, 32), f32)
        for k in range(KMAX):
            m = 32 * k + lane
            node[k] = np.where(m <= n, text[np.minimum(m, n)], 0)
            en[k] = np.where(m <= n - 1, edge[np.minimum(m + 1, n)], -np.inf)
        best = np.zeros((2, 32), f32)
        key = np.full((2, 32), 0xffffffff, np.int64)
        for lo, hi in ((lo1, hi1), (lo2, hi2)):
            if lo >= hi:
                continue
            prev = np.zeros((KMAX + 1, 32), f32)
            for r in range(lo - 1, hi):
                row = D[text[r]]
                k0 = (r + 1) >> 5
                cur = prev.copy()                                  # stale below the first gathered block, like the registers
                cur[KMAX] = 0
                for kb in range(0, KMAX, 4):
                    if kb + 3 >= k0:
                        for k in range(kb, kb + 4):
                            cur[k] = row[node[k]]
                e_i = edge[r] if r >= lo else f32(-np.inf)
                if r >= lo:
                    rows_seen.append(r)
                for kb in range(0, KMAX, 4):
                    if kb + 3 >= k0:
                        for k in range(kb, kb + 4):
                            mine = np.where(lane == 0, cur[k + 1], cur[k])
                            nb = mine[(lane + 1) & 31]
                            with np.errstate(all="ignore"):
                                change = ((prev[k] + nb).astype(f32) - e_i).astype(f32) - en[k]
                            take = (32 * k + lane > r) & (change < best[k & 1])
                            best[k & 1] = np.where(take, change, best[k & 1])
                            key[k & 1] = np.where(take, r * n + 32 * k + lane, key[k & 1])
                prev = cur
        cand += [(float(best[h, l]), int(key[h, l])) for h in range(2) for l in range(32)]
    assert sorted(rows_seen) == list(range(1, n - 1))             # every row exactly once
    b, k = min(cand)
    if not (b < -1e-6):
        return False
    i, j = k // n, k % n
    tour[i:j + 1] = tour[i:j + 1][::-1].copy()
    return True


@pytest.mark.parametrize("n,count,it", [(4, 3, 10), (5, 5, 50), (33, 4, 40), (64, 3, 20), (100, 2, 12), (127, 2, 8),
                                        (128, 2, 8), (200, 1, 6), (255, 1, 4), (256, 1, 4)])
def test_register_carry_dataflow_equals_reference_two_opt(n, count, it):
    rng = np.random.default_rng(n)
    xy = rng.random((n, 2), dtype=np.float32)
    dist = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1)).astype(np.float32)
    np.fill_diagonal(dist, 1e9)
    if n % 2:                                                      # asymmetric matrix (the NLS heuristic "distance" is one)
        dist = (dist * rng.random((n, n), dtype=np.float32)).astype(np.float32)
    tours = np.stack([np.concatenate(([0], 1 + rng.permutation(n - 1))) for _ in range(count)]).astype(np.uint16)
    ref = T2.batched_two_opt(dist, tours, it)
    KMAX = 4 if n + 1 <= 128 else (8 if n + 1 <= 256 else 16)
    for a in range(count):
        t = tours[a].copy()
        for _ in range(it):
            if not _pass(dist, t, n, 8, KMAX):
                break
        assert np.array_equal(t, ref[a]), (n, a)
