"""Training-mode heuristic network on one B200: forward + backward time of the native kernels (per cluster size) next
to torch autograd through the restated ops (oracle/net_torch.py, the formulation the reference runs through PyG) on the
same GPU, at the graph sizes of BASELINE.json's configs C2 / C3 / C4, plus one full `train_instance` step
(tsp/train.ipynb cell 1: Net -> heuristic matrix -> ACO.sample -> REINFORCE loss -> backward -> AdamW step).
CUDA-event timing, informational (not bench.py's metric).   python tools/bench_gnn_train.py [--iters 20]"""
import argparse
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = "cuda"


def timed(fn, iters, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def graph(kind):
    torch.manual_seed(11)
    if kind == "C2 tsp n=100 k=20":
        from deepaco_b200.tsp.net import Net
        from deepaco_b200.tsp.utils import gen_pyg_data
        return Net, gen_pyg_data(torch.rand(100, 2, device=DEV), 20)[0]
    if kind == "C3 tsp_nls n=500 k=50":
        from deepaco_b200.tsp_nls.net import Net
        from deepaco_b200.tsp_nls.utils import gen_pyg_data
        return Net, gen_pyg_data(torch.rand(500, 2, device=DEV), 50, start_node=0)[0]
    from deepaco_b200.cvrp.net import Net
    from deepaco_b200.cvrp.utils import gen_instance, gen_pyg_data
    demand, dist = gen_instance(100, DEV)
    return Net, gen_pyg_data(demand, dist, DEV)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from oracle import net_torch
    out = []
    for kind in ("C2 tsp n=100 k=20", "C3 tsp_nls n=500 k=50", "C4 cvrp N=101 dense"):
        Net, pyg = graph(kind)
        torch.manual_seed(0)
        net = Net().to(DEV).train()
        E = pyg.edge_index.shape[1]
        c = torch.rand(E, device=DEV)
        row = {"graph": kind, "n": pyg.x.shape[0], "E": E}

        def torch_step():
            net.zero_grad(set_to_none=True)
            (net_torch.net_forward(net, pyg) * c).sum().backward()

        row["torch_autograd_ms"] = round(timed(torch_step, args.iters), 4)
        for ctas in (1, 4, 8, 16, 32, 64):
            os.environ["DEEPACO_GNN_CTAS"] = str(ctas)

            def native_step():
                net.zero_grad(set_to_none=True)
                (net(pyg) * c).sum().backward()

            def native_fwd():
                with torch.no_grad():
                    net(pyg)

            row[f"native_ctas{ctas}_fwd_bwd_ms"] = round(timed(native_step, args.iters), 4)
            row[f"native_ctas{ctas}_fwd_only_ms"] = round(timed(native_fwd, args.iters), 4)
        os.environ.pop("DEEPACO_GNN_CTAS")
        row["native_default_ms"] = round(timed(native_step, args.iters), 4)
        out.append(row)
        print(json.dumps(row), flush=True)

    # one train_instance step, TSP-100, 20 ants (tsp/train.ipynb cells 1-2)
    from deepaco_b200.tsp.aco import ACO
    from deepaco_b200.tsp.net import Net
    from deepaco_b200.tsp.utils import gen_pyg_data
    torch.manual_seed(1)
    net = Net().to(DEV).train()
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4)
    pyg, dist = gen_pyg_data(torch.rand(100, 2, device=DEV), 20)

    def train_instance(forward):
        heu_vec = forward()
        heu_mat = net.reshape(pyg, heu_vec) + 1e-10
        aco = ACO(n_ants=20, heuristic=heu_mat, distances=dist, device=DEV)
        costs, log_probs = aco.sample()
        loss = torch.sum((costs - costs.mean()) * log_probs.sum(dim=0)) / 20
        opt.zero_grad()
        loss.backward()
        opt.step()

    row = {"step": "train_instance TSP-100 20 ants",
           "native_ms": round(timed(lambda: train_instance(lambda: net(pyg)), args.iters), 4),
           "torch_ops_network_ms": round(timed(lambda: train_instance(lambda: net_torch.net_forward(net, pyg)), args.iters), 4)}
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
