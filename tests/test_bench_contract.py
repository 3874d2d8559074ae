"""The parts of the measurement contract that can be checked without a GPU: the reference arm of bench.py prints one
JSON line with the agreed keys (it times the CPU restatement of the reference path on the host cores), and the ctypes
stub shown in INTEGRATION.md has the arity of the real entry point."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ant-tours/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ant-tours/sec TSP-100 n_ants=512")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "ant-tours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_integration_stub_matches_the_entry_point():
    from deepaco_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"_L\.deepaco_tsp_sample\.argtypes = \[(.*?)\]\n", text, re.S)
    assert m, "INTEGRATION.md no longer shows the ctypes stub"
    n_args = len([a for a in m.group(1).replace("\n", " ").split(",") if a.strip()])
    assert n_args == len(_lib._SIGNATURES["deepaco_tsp_sample"][1])
    call = re.search(r"rc = _L\.deepaco_tsp_sample\((.*?)\)\n    assert rc == 0", text, re.S)
    assert call and len([a for a in re.sub(r"\([^()]*\)", "", call.group(1)).split(",") if a.strip()]) == n_args
