"""Snapshot of the reference's drop-in boundary: for tsp/, tsp_nls/ and cvrp/ the public methods of
`ACO` with their parameter names and defaults, the attributes `ACO.__init__` sets, the `Net` surface and the
module-level functions of utils.py -- taken by importing the UNMODIFIED reference (build container only).

    python tests/golden/make_api_surface.py      ->  tests/golden/api_surface.json
"""
import inspect
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_ref  # noqa: E402


def params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        if p.name == "self":
            continue
        out.append([p.name, None if p.default is inspect._empty else repr(p.default)])
    return out


def methods(cls, skip=()):
    return {k: params(v) for k, v in vars(cls).items()
            if callable(v) and not k.startswith("__") and k not in skip} | {"__init__": params(cls.__init__)}


ADAPTIVE = ("improvement_phase", "intensification_phase", "diversification_phase", "get_top_solutions", "two_opt",
            "insertion_single", "rearrange", "select_customers", "route_swap", "cross_exchange", "N1_neighbourhood",
            "N2_neighbourhood", "get_subroutes", "insertion", "merge_subroutes")   # adaptive elitist AS, cvrp/aco.py:207-384


def main():
    out = {}
    for sub in ("tsp", "tsp_nls", "cvrp"):
        aco = load_ref(sub, "aco")
        net = load_ref(sub, "net")
        utils = load_ref(sub, "utils")
        n = 6
        torch.manual_seed(0)
        dist = torch.rand(n, n) + 0.1
        if sub == "cvrp":
            inst = aco.ACO(dist, torch.cat((torch.zeros(1), torch.ones(n - 1))), n_ants=2)
        else:
            inst = aco.ACO(dist, n_ants=2)
        skip = ADAPTIVE if sub == "cvrp" else ()
        out[sub] = {
            "ACO": {k: v for k, v in methods(aco.ACO).items() if k not in skip},
            "ACO_attributes": sorted(k for k in vars(inst) if not k.startswith("_")),
            # properties / cached properties of the class (tsp_nls: distances_numpy, heuristic_numpy, heuristic_dist)
            "ACO_properties": sorted(k for k, v in vars(aco.ACO).items()
                                     if not k.startswith("_") and (isinstance(v, property) or type(v).__name__ == "cached_property")),
            # public module-level functions of aco.py (tsp_nls: inference_batch_sample)
            "aco_functions": {k: params(v) for k, v in vars(aco).items()
                              if not k.startswith("_") and callable(v) and getattr(v, "__module__", None) == aco.__name__
                              and not inspect.isclass(v)},
            "Net": methods(net.Net),
            "Net_state_dict_keys": sorted(net.Net().state_dict().keys()),
            "utils": {k: params(v) for k, v in vars(utils).items()
                      if inspect.isfunction(v) and v.__module__ == utils.__name__},
        }
    with open(os.path.join(HERE, "api_surface.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for sub, d in out.items():
        print(sub, "ACO:", sorted(d["ACO"]), "| utils:", sorted(d["utils"]))


if __name__ == "__main__":
    main()
