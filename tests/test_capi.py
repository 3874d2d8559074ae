"""CPU-side checks of the boundary: the shared library loads and exports every symbol the header declares."""
import os
import re

from deepaco_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "deepaco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(deepaco_[a-z0-9_]+)\s*\(", text))


def test_library_loads_and_exports_header_symbols():
    h = _lib.lib()
    declared = _header_symbols()
    assert declared, "no symbols parsed from the header"
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/deepaco_b200.h but not exported"
    assert set(_lib.exported_symbols()) == declared, "ctypes signature table out of sync with the header"
    assert h.deepaco_version() >= 100


def test_sum_plan_matches_aten_rules():
    from deepaco_b200 import _engine as E
    assert E.aten_sum_plan(100, 512) == (32, False, True)
    assert E.aten_sum_plan(20, 8) == (16, False, True)
    assert E.aten_sum_plan(200, 256) == (32, True, True)
    assert E.aten_sum_plan(500, 256) == (32, True, True)
    assert E.aten_sum_plan(101, 512) == (32, False, True)
    bw, vec, exact = E.aten_sum_plan(100, 8)      # few rows -> ATen widens the block: not reproduced
    assert bw == 64 and not exact
