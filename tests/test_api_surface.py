"""Drop-in boundary: the classes and functions under deepaco_b200/{tsp,tsp_nls,cvrp} expose the
reference's public surface -- same method names, same parameter names in the same order with the same defaults, same
public attributes after construction, same Net state_dict keys, same utils functions.  The expectation is a snapshot
taken from the unmodified reference (tests/golden/api_surface.json, made by tests/golden/make_api_surface.py); nothing
here touches a GPU: signatures only."""
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "api_surface.json")) as f:
    SURFACE = json.load(f)


def _params(fn):
    return [[p.name, None if p.default is inspect._empty else repr(p.default)]
            for p in inspect.signature(fn).parameters.values() if p.name != "self"]


def _lookup(cls, name):
    for klass in cls.__mro__:
        if name in vars(klass):
            v = vars(klass)[name]
            return v.__func__ if isinstance(v, (staticmethod, classmethod)) else v
    raise AssertionError(f"{cls.__module__}.{cls.__name__} has no attribute {name}")


@pytest.mark.parametrize("sub", ["tsp", "tsp_nls", "cvrp"])
def test_aco_methods_and_signatures(sub):
    ACO = importlib.import_module(f"deepaco_b200.{sub}.aco").ACO
    for name, want in SURFACE[sub]["ACO"].items():
        assert _params(_lookup(ACO, name)) == want, (sub, name)


@pytest.mark.parametrize("sub", ["tsp", "tsp_nls", "cvrp"])
def test_aco_public_attributes_are_declared(sub):
    """Every attribute the reference's __init__ sets is readable on the class here (instance attribute assigned in
    __init__ or a property) -- checked on the source, no device needed."""
    mod = importlib.import_module(f"deepaco_b200.{sub}.aco")
    src = "".join(inspect.getsource(k) for k in mod.ACO.__mro__ if k is not object)
    for attr in SURFACE[sub]["ACO_attributes"]:
        declared = f"self.{attr} =" in src or f"self.{attr}," in src or f"def {attr}(self" in src
        assert declared, (sub, attr)


@pytest.mark.parametrize("sub", ["tsp", "tsp_nls", "cvrp"])
def test_net_surface_and_checkpoint_keys(sub):
    Net = importlib.import_module(f"deepaco_b200.{sub}.net").Net
    for name, want in SURFACE[sub]["Net"].items():
        assert _params(_lookup(Net, name)) == want, (sub, name)
    assert sorted(Net().state_dict().keys()) == SURFACE[sub]["Net_state_dict_keys"]


@pytest.mark.parametrize("sub", ["tsp", "tsp_nls", "cvrp"])
def test_utils_functions(sub):
    utils = importlib.import_module(f"deepaco_b200.{sub}.utils")
    for name, want in SURFACE[sub]["utils"].items():
        assert hasattr(utils, name), (sub, name)
        assert _params(getattr(utils, name)) == want, (sub, name)


@pytest.mark.parametrize("sub", ["tsp", "tsp_nls", "cvrp"])
def test_aco_properties_and_module_functions(sub):
    """Properties of the reference class (tsp_nls: distances_numpy / heuristic_numpy / heuristic_dist) and the public
    module-level functions of aco.py (tsp_nls: inference_batch_sample) exist with the reference's parameters."""
    mod = importlib.import_module(f"deepaco_b200.{sub}.aco")
    for name in SURFACE[sub]["ACO_properties"]:
        assert isinstance(_lookup(mod.ACO, name), property) or type(_lookup(mod.ACO, name)).__name__ == "cached_property", (sub, name)
    for name, want in SURFACE[sub]["aco_functions"].items():
        assert _params(getattr(mod, name)) == want, (sub, name)
