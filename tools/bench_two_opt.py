"""K4 timing on one B200: 2-opt (n//4 passes) and NLS (T_nls=10, T_p=20) on sampled tours at the reference's tsp_nls
sizes, for the register-carry kernel (default) and the legacy band kernel (DEEPACO_2OPT_LEGACY=1), with a result
equality check between the two.  CUDA events; informational.   python tools/bench_two_opt.py [--sizes 500x256,...]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepaco_b200 import _engine as E

dev = "cuda"


def timeit(fn, iters, warm=1):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="500x256,200x256,100x512")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--variants", default="v2,legacy")
    args = ap.parse_args()
    for spec in args.sizes.split(","):
        n, A = (int(v) for v in spec.split("x"))
        k = max(10, n // 10)
        torch.manual_seed(0)
        xy = torch.rand(n, 2, device=dev)
        d = torch.cdist(xy[None], xy[None])[0]
        d[torch.arange(n), torch.arange(n)] = 1e9
        _, idx = torch.topk(d, k, dim=1, largest=False)
        heu = torch.full_like(d, 1e-10).scatter_(1, idx, torch.rand(n, k, device=dev) * 0.9 + 0.05)
        base = E.tsp_sample(torch.ones_like(d), heu, A, start_node=0, double_norm=True, seed=1, want_paths=False, want_tours=True)[2]
        hd = (1 / (heu / heu.max(-1, keepdim=True).values + 1e-5)).contiguous()
        row = {"n": n, "ants": A}
        outs = {}
        for tag in args.variants.split(","):
            os.environ["DEEPACO_2OPT_LEGACY"] = "1" if tag == "legacy" else "0"
            t = base.clone()
            _, passes = E.two_opt_(d, t, n // 4, want_passes=True)
            t2 = base.clone()
            _, passes_nls = E.tsp_nls_(d, hd, t2, n // 4, want_passes=True)
            outs[tag] = (t, t2)
            row[f"{tag}_two_opt_ms"] = round(timeit(lambda: E.two_opt_(d, base.clone(), n // 4), args.iters), 3)
            row[f"{tag}_nls_ms"] = round(timeit(lambda: E.tsp_nls_(d, hd, base.clone(), n // 4), max(1, args.iters // 2)), 3)
            row[f"{tag}_passes_two_opt"] = int(passes.sum())
            row[f"{tag}_passes_nls"] = int(passes_nls.sum())
            row[f"{tag}_us_per_pass"] = round(1e3 * row[f"{tag}_two_opt_ms"] / max(1, int(passes.max())), 2)
        os.environ.pop("DEEPACO_2OPT_LEGACY")
        if len(outs) == 2:
            a, b = outs.values()
            row["identical"] = bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
